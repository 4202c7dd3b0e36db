"""Single-GPU checks of the halo row movers of include/gmlm_b200.h.  On one device the
"peer" addresses are plain local addresses, so the pointer-gather kernels — the LDG pull and the bulk-copy
(TMA, cp.async.bulk) pull — can be checked bit for bit against torch indexing; the multi-GPU run of the same
kernels over NVLink is tests/multi_gpu_check.py."""
import ctypes as C

import pytest
import torch

from gmlm_b200 import _lib
from gmlm_b200.graph import _ptr, _stream

pytestmark = pytest.mark.gpu


def _row_ptrs(src, ids):
    return (src.data_ptr() + ids * (src.size(1) * src.element_size())).contiguous()


@pytest.mark.parametrize("dtype,feat", [(torch.bfloat16, 256), (torch.bfloat16, 64), (torch.float32, 256),
                                        (torch.float32, 20), (torch.bfloat16, 1024)])
@pytest.mark.parametrize("n", [1, 31, 32, 33, 5000])
@pytest.mark.parametrize("impl", ["ldg", "tma1", "tma7", "tma_all", "tma_1warp", "tma_8warps", "tma_rb16", "tma_rb8"])
def test_gather_rows_ptr_matches_indexing(cuda_dev, dtype, feat, n, impl):
    lib = _lib.load()
    g = torch.Generator(device="cpu").manual_seed(n * 131 + feat)
    src = torch.randn(7000, feat, generator=g).to(cuda_dev).to(dtype)
    ids = torch.randint(0, 7000, (n,), generator=g).to(cuda_dev)
    out_ids = torch.randperm(n + 5, generator=g)[:n].to(cuda_dev)            # scattered destinations
    out = torch.full((n + 5, feat), -7.0, device=cuda_dev, dtype=dtype)
    ptrs = _row_ptrs(src, ids)
    code = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[dtype]
    if impl == "ldg":
        rc = lib.gmlm_gather_rows_ptr(_ptr(ptrs), _ptr(out_ids), code, feat, n, _ptr(out), feat, _stream(cuda_dev))
    else:
        ctas, warps, kb, rb = {"tma1": (1, 0, 0, 0), "tma7": (7, 0, 0, 0), "tma_all": (0, 0, 0, 0), "tma_1warp": (5, 1, 0, 0),
                               "tma_8warps": (3, 8, 0, 0), "tma_rb16": (9, 2, 100, 16), "tma_rb8": (4, 2, 64, 8)}[impl]
        row_bytes = feat * src.element_size()
        if warps == 8 and row_bytes > 256:
            warps = 2 if row_bytes <= 1024 else 1                     # 8 rings of wide rows do not fit one CTA
        if rb and row_bytes > 512:
            kb = 200                                                  # wide rows need the full ring
        rc = lib.gmlm_gather_rows_ptr_tma(_ptr(ptrs), _ptr(out_ids), code, feat, n, _ptr(out), feat, ctas, warps, kb, rb,
                                          _stream(cuda_dev))
    _lib.check(rc, "gather_rows_ptr")
    want = torch.full_like(out, -7.0)
    want[out_ids] = src[ids]
    assert torch.equal(out, want)


def test_gather_rows_ptr_tma_small_ring_and_identity_destinations(cuda_dev):
    """A one-warp 24 KiB ring (3 slots of 32 x 256 B) wraps many times; out_ids = NULL means identity."""
    lib = _lib.load()
    n, feat = 20_000, 128
    src = torch.randn(30_000, feat, device=cuda_dev).to(torch.bfloat16)
    ids = torch.randint(0, 30_000, (n,), device=cuda_dev)
    out = torch.empty((n, feat), device=cuda_dev, dtype=torch.bfloat16)
    ptrs = _row_ptrs(src, ids)
    rc = lib.gmlm_gather_rows_ptr_tma(_ptr(ptrs), C.c_void_p(0), _lib.BF16, feat, n, _ptr(out), feat, 3, 1, 29, 0,
                                      _stream(cuda_dev))
    _lib.check(rc, "gather_rows_ptr_tma")
    assert torch.equal(out, src[ids])


def test_gather_rows_ptr_tma_rejects_wide_rows(cuda_dev):
    lib = _lib.load()
    src = torch.zeros(4, 4096, device=cuda_dev)                               # 16 KiB rows: no ring fits
    ids = torch.arange(4, device=cuda_dev)
    out = torch.empty_like(src)
    ptrs = _row_ptrs(src, ids)
    rc = lib.gmlm_gather_rows_ptr_tma(_ptr(ptrs), C.c_void_p(0), _lib.F32, 4096, 4, _ptr(out), 4096, 0, 0, 0, 0,
                                      _stream(cuda_dev))
    assert rc != 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_reduce_rows_ptr_adds_entries_in_order(cuda_dev, dtype):
    lib = _lib.load()
    feat, n_dst, n_src = 64, 300, 2000
    g = torch.Generator().manual_seed(3)
    dst = torch.randn(n_dst, feat, generator=g).to(cuda_dev).to(dtype)
    src = torch.randn(n_src, feat, generator=g).to(cuda_dev).to(dtype)
    rows = torch.randperm(n_dst, generator=g)[:200].sort().values.to(cuda_dev)
    counts = torch.randint(1, 8, (200,), generator=g)
    rowptr = torch.zeros(201, dtype=torch.int32)
    rowptr[1:] = torch.cumsum(counts, 0)
    entries = torch.randint(0, n_src, (int(rowptr[-1]),), generator=g).to(cuda_dev)
    want = dst.float().clone()
    rp = rowptr.tolist()
    for r in range(200):
        acc = want[rows[r]].clone()
        for e in range(rp[r], rp[r + 1]):
            acc = acc + src[entries[e]].float()
        want[rows[r]] = acc
    code = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[dtype]
    rowptr_d, ptrs = rowptr.to(cuda_dev), _row_ptrs(src, entries)      # named: they must outlive the launch
    rc = lib.gmlm_reduce_rows_ptr(_ptr(dst), code, feat, feat, _ptr(rows), _ptr(rowptr_d), _ptr(ptrs), 200,
                                  _stream(cuda_dev))
    _lib.check(rc, "reduce_rows_ptr")
    assert torch.equal(dst, want.to(dtype))
