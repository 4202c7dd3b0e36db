#!/usr/bin/env python
"""Development: end-of-round ncu targets -- the four-stage fusion GEMM, the 3xTF32 GEMM, the narrow-row kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import gmlm_b200 as G
from gmlm_b200 import _lib, synth
from gmlm_b200.ops import gemm_nt

dev = torch.device("cuda:0")
m = 2_000_000
xs = [torch.randn(m, k, device=dev).bfloat16() for k in (64, 128, 256, 512)]
wf = torch.randn(768, 960, device=dev).bfloat16()
for _ in range(2):
    y = gemm_nt(xs, wf, bias=torch.zeros(768, device=dev))
del xs, y
# fp32 operands (3xTF32): the C2 layer-1 shape at the reference's widths, scaled to 400k rows
a = torch.randn(400_000, 1200, device=dev)
x = torch.randn(400_000, 300, device=dev)
w = torch.randn(512, 1500, device=dev)
for _ in range(2):
    y = gemm_nt([a, x[:, :300]], w, bias=torch.zeros(512, device=dev))
del a, x, y
wl = synth.WORKLOADS["c4"]
ei = synth.make_graph(wl, device=dev)
et = G.edge_type_from_degree(ei, wl.num_nodes)
g = G.get_rel_graph(ei, et, wl.num_nodes, 5)
xf = synth.make_features(wl.num_nodes, 64, device=dev, dtype=torch.bfloat16)
gh = synth.make_features(wl.num_nodes * g.num_slots, 64, device=dev, seed=7, dtype=torch.bfloat16)
for _ in range(2):
    G.spmm(xf, g.fwd, _lib.AGG_MEAN)
    G.spmm(gh, g.bwd, _lib.AGG_WEIGHTED)
torch.cuda.synchronize()
print("done")
