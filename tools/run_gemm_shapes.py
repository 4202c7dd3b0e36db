"""Development: the layer shapes of the tcgen05 GEMM, three calls each (ncu target)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gmlm_b200.ops import gemm_nt

dev = torch.device("cuda:0")
m = 2_000_000
for (n, k1, k2) in [(64, 1024, 256), (1280, 64, 0), (320, 256, 0)]:
    a1 = torch.randn(m, k1, device=dev).bfloat16()
    a2 = torch.randn(m, k2, device=dev).bfloat16() if k2 else None
    b = torch.randn(n, k1 + k2, device=dev).bfloat16()
    bias = torch.randn(n, device=dev)
    for _ in range(3):
        c = gemm_nt(a1, b, bias=bias, a2=a2)
    torch.cuda.synchronize()
    del a1, a2, b, c
print("ok")
