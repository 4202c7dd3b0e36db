"""Development: GB/s of the pointer row gathers on ONE GPU (local addresses): the LDG kernel vs the bulk-copy
(TMA) kernel at several CTA counts and ring sizes.  Tells how many bytes per SM the bulk-copy path sustains
for 512-byte rows (the figure that decides whether it can feed an aggregation or only a halo pull)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from gmlm_b200 import _lib
from gmlm_b200.graph import _ptr, _stream

dev = torch.device("cuda:0")
lib = _lib.load()
for feat in (256, 64):
    n_src, n = 4_000_000, 4_000_000
    src = torch.randn(n_src, feat, device=dev).to(torch.bfloat16)
    ids = torch.randint(0, n_src, (n,), device=dev)
    ptrs = (src.data_ptr() + ids * (feat * 2)).contiguous()
    out = torch.empty((n, feat), device=dev, dtype=torch.bfloat16)
    nbytes = 2 * n * feat * 2

    def timeit(fn, iters=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    ms = timeit(lambda: _lib.check(lib.gmlm_gather_rows_ptr(_ptr(ptrs), None, _lib.BF16, feat, n, _ptr(out), feat, _stream(dev))))
    print(f"feat {feat}: LDG gather        {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s (read+write)")
    for ctas in (16, 32, 74, 148):
        for warps in (1, 2, 3, 4, 6, 8):
            if lib.gmlm_gather_rows_ptr_tma(_ptr(ptrs), None, _lib.BF16, feat, n, _ptr(out), feat, ctas, warps, 0, 0,
                                            _stream(dev)):
                continue                                            # ring too small for that many warps
            ms = timeit(lambda: _lib.check(lib.gmlm_gather_rows_ptr_tma(_ptr(ptrs), None, _lib.BF16, feat, n, _ptr(out), feat,
                                                                        ctas, warps, 0, 0, _stream(dev))))
            print(f"feat {feat}: TMA ctas={ctas:3d} warps={warps}  {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s  "
                  f"({nbytes / 2 / ms / 1e6 / ctas:.1f} GB/s gathered per CTA)")
    assert torch.equal(out, src[ids])
