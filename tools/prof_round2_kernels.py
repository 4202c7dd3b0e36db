#!/usr/bin/env python
"""Development: run the kernels added in the second half of round 2 once each at their production shapes, for an
ncu capture: gemm_tn (weight gradients), basis_compose fwd/bwd, gemm_nt with a residual addend / four sources."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gmlm_b200.ops import basis_compose, basis_compose_bwd, gemm_nt, gemm_tn

dev = torch.device("cuda:0")
m = 2_000_000
# C4 layer 4 (H = 64: 256 -> 512): dW | droot = [h | x]^T g
h = torch.randn(m, 1024, device=dev).bfloat16()
x = torch.randn(m, 256, device=dev).bfloat16()
g = torch.randn(m, 512, device=dev).bfloat16()
for _ in range(2):
    d = gemm_tn([h, x], g)
# C4 layer 2: [256 | 64]^T g[128]
g2 = torch.randn(m, 128, device=dev).bfloat16()
for _ in range(2):
    d = gemm_tn([h[:, :256], x[:, :64]], g2)
del h, g, g2
# residual projection with the add in the epilogue: x1 + Linear(256 -> 64)(x_feat)
w = torch.randn(64, 256, device=dev).bfloat16()
acc = torch.randn(m, 64, device=dev).bfloat16()
for _ in range(2):
    y = gemm_nt(x, w, bias=torch.zeros(64, device=dev), addend=acc)
del x, acc, y
# MultiScaleFusion: four sources, never concatenated
xs = [torch.randn(m, k, device=dev).bfloat16() for k in (64, 128, 256, 512)]
wf = torch.randn(768, 960, device=dev).bfloat16()
for _ in range(2):
    y = gemm_nt(xs, wf, bias=torch.zeros(768, device=dev))
del xs, y
# basis composition at the reference's widths (H = 512, layer 4: 2048 -> 4096, 30 bases = 1.0 GB fp32)
weight = torch.randn(30, 2048, 4096, device=dev)
comp = torch.randn(5, 30, device=dev)
root = torch.randn(2048, 4096, device=dev)
for _ in range(2):
    wn, wt = basis_compose(weight, comp, root, (0, 1, 2, 3), torch.float16, "agg")
dw = torch.randn(4 * 2048, 4096, device=dev)
for _ in range(2):
    dweight, dcomp = basis_compose_bwd(weight, comp, dw, 2048 * 4096, 4096, (0, 1, 2, 3))
torch.cuda.synchronize()
print("done")
