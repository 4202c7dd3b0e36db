#!/usr/bin/env python
"""Development: which elementwise adds does one C4 encoder step launch (shapes / strides)?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import profile, ProfilerActivity
import gmlm_b200 as G
from gmlm_b200 import synth

dev = torch.device("cuda:0")
w = synth.WORKLOADS["c4"]
dtype = torch.bfloat16
ei = synth.make_graph(w, device=dev)
x = synth.make_features(w.num_nodes, w.feat, device=dev, dtype=dtype)
et = G.edge_type_from_degree(ei, w.num_nodes)
enc = G.GraphEncoder(w.feat, w.hidden, 768, dropout_rate=0.0, act_dtype=dtype).to(dev)
enc.residual_proj1.to(dtype), enc.residual_proj2.to(dtype), enc.multi_scale_fusion.to(dtype)
xg = x.detach().requires_grad_(True)


def step():
    y = enc.get_graph_embeddings(xg, ei, et)
    y.backward(torch.ones_like(y))
    xg.grad = None
    enc.zero_grad(set_to_none=True)


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=False) as prof:
    step()
    torch.cuda.synchronize()
rows = []
for e in prof.events():
    if e.name in ("aten::add", "aten::add_", "aten::fill_", "aten::zeros_like", "aten::zero_", "aten::copy_", "aten::cat") and e.device_time_total > 50:
        rows.append((e.device_time_total, e.name, str(e.input_shapes)))
rows.sort(reverse=True)
for t, n, sh in rows[:40]:
    print(f"{t / 1e3:7.3f} ms  {n:16s} {sh}")
