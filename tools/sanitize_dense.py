#!/usr/bin/env python
"""Development (GPU): small invocations of the dense-side kernels for a compute-sanitizer pass
(`compute-sanitizer --tool memcheck python tools/sanitize_dense.py`)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gmlm_b200.ops import basis_compose, basis_compose_bwd, gemm_nt, gemm_tn

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def rnd(*shape, dt=torch.bfloat16):
    return torch.randn(*shape, generator=g).to(dt).to(dev)


for dt in (torch.bfloat16, torch.float16, torch.float32):
    for (m, n, ks) in [(300, 512, (1200, 300)), (129, 36, (72, 20)), (1000, 768, (64, 128, 256, 512)), (70, 8, (8,))]:
        srcs = [rnd(m, k, dt=dt) for k in ks]
        b = rnd(n, sum(ks), dt=dt)
        out = gemm_nt(srcs, b, bias=rnd(n, dt=torch.float32), out_dtype=torch.float32)
        ref = torch.cat([s.float() for s in srcs], 1) @ b.float().t()
        assert torch.isfinite(out).all()
    a = rnd(513, 320, dt=dt)
    b = rnd(1500, 320, dt=dt)
    c1, c2 = gemm_nt(a, b, out_dtype=torch.float32, split=1200)
    add = rnd(513, 1500, dt=torch.float32)
    gemm_nt(a, b, out_dtype=torch.float32, addend=add)
for dt in (torch.bfloat16, torch.float16):
    for (m, ks, n) in [(1000, (64,), 64), (5000, (1200, 300, 8), 512), (777, (100, 36), 200), (37, (130,), 70)]:
        d = gemm_tn([rnd(m, k, dt=dt) for k in ks], rnd(m, n, dt=dt))
        assert torch.isfinite(d).all()
for (fi, fo) in [(300, 512), (33, 20), (64, 64)]:
    w, c, r = rnd(30, fi, fo, dt=torch.float32), rnd(5, 30, dt=torch.float32), rnd(fi, fo, dt=torch.float32)
    for op in (torch.float16, torch.float32):
        for layout in ("agg", "tf"):
            basis_compose(w, c, r, (0, 1, 3), op, layout)
    if fo % 4 == 0:
        basis_compose_bwd(w, c, rnd(3 * fi, fo, dt=torch.float32), fi * fo, fo, (0, 1, 3))
torch.cuda.synchronize()
print("sanitize_dense done")
