#!/usr/bin/env python
"""Development: encoder fwd+bwd time on a workload with the segment-compact formulation forced on / off per layer
(RGCNConv.segment_compact; None = the layer's own FLOP / byte model).  Usage: try_segment_compact.py [c4|c2|c3]"""
import itertools
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
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
ones = None


def step():
    global ones
    with torch.amp.autocast("cuda", enabled=dtype == torch.float32):
        y = enc.get_graph_embeddings(xg, ei, et)
    if ones is None:
        ones = torch.ones_like(y)
    y.backward(ones)
    xg.grad = None
    enc.zero_grad(set_to_none=True)


def timed(n=8):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


convs = [enc.rgcn2, enc.rgcn3, enc.rgcn4]
print("default (model):", [c._use_segment_compact(G.get_rel_graph(ei, et, w.num_nodes, 5), x) for c in convs],
      f"{timed():.3f} ms", flush=True)
for combo in itertools.product([False, True], repeat=3):
    for c, v in zip(convs, combo):
        c.segment_compact = v
    print("layers 2,3,4 compact =", combo, f"{timed():.3f} ms", flush=True)
