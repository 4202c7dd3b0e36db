#!/usr/bin/env python
"""Development: run the dense-side kernels once (tcgen05 GEMM, GraphNorm fwd/bwd) for an ncu capture."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import gmlm_b200 as G
from gmlm_b200.ops import gemm_nt

dev = torch.device("cuda:0")
m = 2_000_000
a1 = torch.randn(m, 1024, device=dev).bfloat16()
a2 = torch.randn(m, 256, device=dev).bfloat16()
b = torch.randn(64, 1280, device=dev).bfloat16()
bias = torch.randn(64, device=dev)
for _ in range(3):
    out = gemm_nt(a1, b, bias=bias, a2=a2)
del a1, a2
for c in (64, 512):
    norm = G.GraphNorm(c).to(dev)
    x = torch.randn(m, c, device=dev).bfloat16().requires_grad_(True)
    for _ in range(2):
        y = norm(x, fuse_gelu=True)
        y.backward(torch.ones_like(y))
        x.grad = None
    del x, y
torch.cuda.synchronize()
print("done")
