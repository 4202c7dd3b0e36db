#!/usr/bin/env python
"""Development micro-benchmark (GPU): tcgen05 gemm_nt vs torch.matmul (cuBLAS) on the layer shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gmlm_b200.ops import gemm_nt


def t(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


dev = torch.device("cuda:0")
for (m, n, k1, k2) in [(2_000_000, 64, 1024, 256), (2_000_000, 128, 256, 64), (2_000_000, 256, 512, 128),
                       (2_000_000, 1024 + 256, 64, 0), (2_000_000, 320, 256, 0), (2_000_000, 256, 320, 0),
                       (2_000_000, 768, 960, 0)]:
    a1 = torch.randn(m, k1, device=dev).bfloat16()
    a2 = torch.randn(m, k2, device=dev).bfloat16() if k2 else None
    b = torch.randn(n, k1 + k2, device=dev).bfloat16()
    bias = torch.randn(n, device=dev)
    ms_tc = t(lambda: gemm_nt(a1, b, bias=bias, a2=a2))
    bt = b.t().contiguous()
    if a2 is not None:
        ms_cb = t(lambda: torch.addmm(bias.bfloat16(), a1, bt[:k1]) + a2 @ bt[k1:])
    else:
        ms_cb = t(lambda: torch.addmm(bias.bfloat16(), a1, bt))
    bytes_ = (m * (k1 + k2) + m * n) * 2
    print(f"M={m} N={n} K={k1}+{k2}: tcgen05 {ms_tc:.3f} ms ({bytes_/ms_tc/1e6:.0f} GB/s, "
          f"{2*m*n*(k1+k2)/ms_tc/1e9:.0f} TFLOP/s)  cuBLAS {ms_cb:.3f} ms ({bytes_/ms_cb/1e6:.0f} GB/s)", flush=True)
    del a1, a2, b

# ---- weight-gradient reductions: D = [A_0 | A_1]^T G (tcgen05, MN-major operands) vs cuBLAS split-K
from gmlm_b200.ops import gemm_tn  # noqa: E402
for (m, ks, n, dt) in [(2_000_000, (1024, 256), 64, torch.bfloat16), (2_000_000, (256, 64), 128, torch.bfloat16),
                       (2_000_000, (512, 128), 256, torch.bfloat16), (2_000_000, (1024, 256), 512, torch.bfloat16),
                       (2_000_000, (256,), 320, torch.bfloat16), (2_000_000, (64, 128, 256, 512), 768, torch.bfloat16),
                       (22_662, (8192, 2048), 4096, torch.float16), (22_662, (1200, 300), 512, torch.float16)]:
    srcs = [torch.randn(m, k, device=dev).to(dt) for k in ks]
    g = torch.randn(m, n, device=dev).to(dt)
    ms_tc = t(lambda: gemm_tn(srcs, g))
    ms_cb = t(lambda: [torch.mm(s_.t(), g, out_dtype=torch.float32) for s_ in srcs])
    bytes_ = (m * sum(ks) + m * n) * 2
    fl = 2 * m * n * sum(ks)
    print(f"TN M={m} K={ks} N={n} {dt}: tcgen05 {ms_tc:.3f} ms ({bytes_/ms_tc/1e6:.0f} GB/s, {fl/ms_tc/1e9:.0f} TFLOP/s)"
          f"  cuBLAS {ms_cb:.3f} ms", flush=True)
    del srcs, g

# ---- basis composition (A4) at the reference's widths: HBM floor = one pass over the fp32 bases
from gmlm_b200.ops import basis_compose, basis_compose_bwd  # noqa: E402
for (fi, fo) in [(2048, 4096), (1024, 2048), (300, 512), (256, 64)]:
    weight = torch.randn(30, fi, fo, device=dev)
    comp = torch.randn(5, 30, device=dev)
    root = torch.randn(fi, fo, device=dev)
    dw = torch.randn(4 * fi, fo, device=dev)
    ms_f = t(lambda: basis_compose(weight, comp, root, (0, 1, 2, 3), torch.float16, "agg"))
    ms_b = t(lambda: basis_compose_bwd(weight, comp, dw, fi * fo, fo, (0, 1, 2, 3)))
    gb = 30 * fi * fo * 4 / 1e9
    print(f"compose Fi={fi} Fo={fo}: fwd {ms_f:.3f} ms ({(gb + 5 * fi * fo * 4 / 1e9) / ms_f * 1e3:.0f} GB/s)  "
          f"bwd {ms_b:.3f} ms ({(2 * gb + 4 * fi * fo * 4 / 1e9) / ms_b * 1e3:.0f} GB/s)", flush=True)
    del weight, root, dw
