#!/usr/bin/env python
"""Development: kernel-time breakdown of one encoder fwd+bwd on a workload (torch profiler)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import profile, ProfilerActivity
import gmlm_b200 as G
from gmlm_b200 import synth

key = sys.argv[1] if len(sys.argv) > 1 else "c4"
dev = torch.device("cuda:0")
w = synth.WORKLOADS[key]
dtype = torch.bfloat16 if w.dtype == "bf16" else torch.float32
ei = synth.make_graph(w, device=dev)
x = synth.make_features(w.num_nodes, w.feat, device=dev, dtype=dtype)
et = G.edge_type_from_degree(ei, w.num_nodes)
enc = G.GraphEncoder(w.feat, w.hidden, 768, dropout_rate=0.0, act_dtype=dtype).to(dev)
if dtype == torch.bfloat16:
    enc.residual_proj1.to(dtype), enc.residual_proj2.to(dtype), enc.multi_scale_fusion.to(dtype)
xg = x.detach().requires_grad_(True)


_ones = {}


def _ones_like(y):
    k = (tuple(y.shape), y.dtype)
    if k not in _ones:
        _ones[k] = torch.ones_like(y)
    return _ones[k]


def step():
    with torch.amp.autocast("cuda", enabled=dtype == torch.float32):     # as bench.py: fp32 workloads under autocast
        y = enc.get_graph_embeddings(xg, ei, et)
    y.backward(_ones_like(y))
    xg.grad = None
    enc.zero_grad(set_to_none=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=70))
# device kernels only, full names
from torch.autograd import DeviceType
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count)
        for e in prof.key_averages() if getattr(e, "device_type", None) == DeviceType.CUDA]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"--- device kernels: total {tot / 1e3:.2f} ms")
for name, t, c in rows[:45]:
    print(f"{t / 1e3:8.3f} ms  {100 * t / tot:5.1f}%  x{c:<3d} {name[:150]}")
