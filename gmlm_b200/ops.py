"""Thin torch layer over the C ABI: tensor checks, output allocation from the caching
allocator, launch on the current stream, autograd registration.

Registered ops (``torch.ops.gmlm.*``, CUDA only — calling them with CPU tensors raises from
the dispatcher; there is no fallback), every one with a fake (meta) implementation:

    degree_i32, edge_type_bucket, csr_build, spmm_csr, colstats, gemm_nt, graphnorm_bwd, layernorm_bwd, soft_mask_bwd
    rgcn_aggregate, plan_aggregate, graphnorm_fwd, layernorm_fwd, soft_mask_fwd      <- differentiable
                                                       (torch.library.register_autograd: backward = the ops above)

Thin functional wrappers used by the modules in ``gmlm_b200.nn``:

    rgcn_aggregate(x, graph)            A5 forward / A14 backward
    graph_norm(x, weight, bias, mean_scale, eps, fuse_gelu)   A7
    soft_mask(x, mask, token, beta)     A11
    layer_norm(x, weight, bias, eps)    A13 (the LayerNorm closing MultiScaleFusion)
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .graph import CSR, RelGraph, _ptr, _require_cuda, _stream

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _dtype_code(t: torch.Tensor, what: str) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise _lib.GmlmError(f"{what}: dtype {t.dtype} unsupported (float32 and bfloat16 only)") from None


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor with unit inner stride (a leading dimension is fine, anything else is copied)."""
    if t.dim() != 2:
        raise _lib.GmlmError(f"expected a 2-D tensor, got shape {tuple(t.shape)}")
    if t.size(1) > 0 and t.stride(1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return int(t.stride(0)) if t.size(0) > 1 else int(t.size(1))


_LIB = torch.library.Library("gmlm", "DEF")
_LIB.define("degree_i32(Tensor index, int num_nodes) -> Tensor")
_LIB.define("edge_type_bucket(Tensor src, Tensor deg, int[] bounds) -> Tensor")
_LIB.define("spmm_csr(Tensor x, Tensor rowptr, Tensor col, Tensor? w, int num_rows, int mode, Tensor? grp_row, "
            "int hub_thresh, Tensor? hub_row, Tensor? hub_chunk_ptr, Tensor? chunk_beg, Tensor? chunk_end) -> Tensor")
_LIB.define("colstats(Tensor x) -> (Tensor, Tensor)")
_LIB.define("graphnorm_fwd(Tensor x, Tensor weight, Tensor bias, Tensor mean_scale, float eps, bool fuse_gelu) "
            "-> (Tensor, Tensor, Tensor)")
_LIB.define("graphnorm_bwd(Tensor x, Tensor gy, Tensor mean, Tensor rstd, Tensor weight, Tensor bias, "
            "Tensor mean_scale, bool fuse_gelu, bool need_gx) -> (Tensor, Tensor, Tensor, Tensor)")
_LIB.define("layernorm_fwd(Tensor x, Tensor weight, Tensor bias, float eps) -> (Tensor, Tensor, Tensor)")
_LIB.define("layernorm_bwd(Tensor x, Tensor gy, Tensor mean, Tensor rstd, Tensor weight, bool need_gx) "
            "-> (Tensor, Tensor, Tensor)")
_LIB.define("soft_mask_fwd(Tensor x, Tensor mask, Tensor token, float beta) -> Tensor")
_LIB.define("soft_mask_bwd(Tensor gy, Tensor mask, float beta, bool need_gx) -> (Tensor, Tensor)")
# differentiable ops (autograd registered below with torch.library.register_autograd)
# a CSR travels as eight (optional) tensors: rowptr, col, w, grp_row, hub_row, hub_chunk_ptr, chunk_beg, chunk_end
_CSR_ARGS = ("Tensor {p}rowptr, Tensor {p}col, Tensor? {p}w, Tensor? {p}grp, Tensor? {p}hub_row, "
             "Tensor? {p}hub_chunk_ptr, Tensor? {p}chunk_beg, Tensor? {p}chunk_end")
_LIB.define("rgcn_aggregate(Tensor x, " + _CSR_ARGS.format(p="f_") + ", " + _CSR_ARGS.format(p="b_") +
            ", int[] meta) -> Tensor")
_LIB.define("plan_aggregate(Tensor rows, " + _CSR_ARGS.format(p="f_") + ", " + _CSR_ARGS.format(p="b_") +
            ", int[] meta) -> Tensor")
_LIB.define("gemm_nt(Tensor a1, Tensor b, Tensor? bias, Tensor? a2, ScalarType out_dtype) -> Tensor")
_LIB.define("csr_build(Tensor row, Tensor col, Tensor? rel, Tensor? keep, int num_rows, int num_cols, int num_relations, "
            "int[] slot_of_rel, int num_slots) -> (Tensor, Tensor, Tensor, Tensor)")


# ------------------------------------------------------------------------------ A1 / A2
def _degree_i32(index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    lib = _lib.load()
    index = index.reshape(-1).contiguous()
    if index.dtype != torch.int64:
        index = index.long()
    with torch.cuda.device(index.device):
        deg = torch.empty(num_nodes, dtype=torch.int32, device=index.device)
        _lib.check(lib.gmlm_degree_i32(_ptr(index), index.numel(), num_nodes, _ptr(deg), 0, _stream(index.device)),
                   "degree")
    return deg


def _edge_type_bucket(src: torch.Tensor, deg: torch.Tensor, bounds: Sequence[int]) -> torch.Tensor:
    lib = _lib.load()
    src = src.reshape(-1).contiguous()
    if src.dtype != torch.int64:
        src = src.long()
    nb = len(bounds)
    arr = (C.c_int32 * max(nb, 1))(*[int(b) for b in bounds])
    with torch.cuda.device(src.device):
        out = torch.empty(src.numel(), dtype=torch.int64, device=src.device)
        _lib.check(lib.gmlm_edge_type_bucket(_ptr(src), src.numel(), _ptr(deg), deg.numel(), arr, nb, _ptr(out),
                                             _stream(src.device)), "edge_type_bucket")
    return out


# ------------------------------------------------------------------------------ A5 / A14
def _spmm_csr(x, rowptr, col, w, num_rows, mode, grp_row, hub_thresh, hub_row, hub_chunk_ptr, chunk_beg, chunk_end,
              out=None):
    lib = _lib.load()
    x = _rowmajor(x)
    feat = int(x.size(1))
    dev = x.device
    n_hub = int(hub_row.numel()) if hub_row is not None else 0
    n_chunks = int(chunk_beg.numel()) if chunk_beg is not None else 0
    n_groups = int(grp_row.numel()) - 1 if grp_row is not None else 0
    w_heads = int(w.size(1)) if (w is not None and w.dim() == 2) else 1
    if col.numel() == 0 and mode == _lib.AGG_WEIGHTED:
        mode, w, w_heads = _lib.AGG_SUM, None, 1          # no edges: every row is empty, the result is zeros
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((num_rows, feat), dtype=x.dtype, device=dev)
        elif out.shape != (num_rows, feat) or out.dtype != x.dtype or not out.is_contiguous():
            raise _lib.GmlmError("spmm: `out` must be a contiguous [num_rows, feat] tensor of x's dtype")
        hub_ws = torch.empty(n_chunks * feat, dtype=torch.float32, device=dev) if n_hub else None
        _lib.check(lib.gmlm_spmm_csr(_ptr(x), _dtype_code(x, "spmm_csr"), feat, _ld(x), _ptr(rowptr), _ptr(col),
                                     _ptr(w), w_heads, num_rows, mode, _ptr(grp_row), n_groups, hub_thresh, n_hub,
                                     n_chunks,
                                     _ptr(hub_row),
                                     _ptr(hub_chunk_ptr), _ptr(chunk_beg), _ptr(chunk_end), _ptr(hub_ws), _ptr(out),
                                     feat, _stream(dev)), "spmm_csr")
    return out


def spmm(x: torch.Tensor, csr: CSR, mode: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[r] = reduce over row r of csr (no autograd).  ``out`` lets the result land in a caller-owned
    buffer (e.g. NVLink-peer-visible symmetric memory for the halo exchange)."""
    _require_cuda(x, "x")
    if out is not None:
        return _spmm_csr(x, csr.rowptr, csr.col, csr.w if mode == _lib.AGG_WEIGHTED else None, csr.num_rows, mode,
                         csr.grp_row, csr.hub_thresh, csr.hub_row, csr.hub_chunk_ptr, csr.chunk_beg, csr.chunk_end,
                         out=out)
    return torch.ops.gmlm.spmm_csr(x, csr.rowptr, csr.col, csr.w if mode == _lib.AGG_WEIGHTED else None,
                                   csr.num_rows, mode, csr.grp_row, csr.hub_thresh, csr.hub_row, csr.hub_chunk_ptr,
                                   csr.chunk_beg, csr.chunk_end)


# ------------------------------------------------------------------------------ A7
def _colstats(x: torch.Tensor):
    lib = _lib.load()
    x = _rowmajor(x)
    n, c = x.shape
    dev = x.device
    with torch.cuda.device(dev):
        colsum = torch.empty(c, dtype=torch.float64, device=dev)
        colsq = torch.empty(c, dtype=torch.float64, device=dev)
        ws_bytes = lib.gmlm_colstats_workspace_bytes(n, c)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gmlm_colstats(_ptr(x), _dtype_code(x, "colstats"), n, c, _ld(x), _ptr(colsum), _ptr(colsq),
                                     _ptr(ws), ws_bytes, _stream(dev)), "colstats")
    return colsum, colsq


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def graphnorm_apply_stats(x, colsum, colsq, weight, bias, mean_scale, eps, fuse_gelu, stat_rows=0):
    """Normalise pass of A7 given the fp64 column sums; ``stat_rows`` = rows the sums cover (0 = x's rows;
    the whole-graph count when x holds one rank's rows of a partition)."""
    lib = _lib.load()
    x = _rowmajor(x)
    n, c = x.shape
    dev = x.device
    weight, bias, mean_scale = _f32c(weight), _f32c(bias), _f32c(mean_scale)
    with torch.cuda.device(dev):
        y = torch.empty((n, c), dtype=x.dtype, device=dev)
        mean = torch.empty(c, dtype=torch.float32, device=dev)
        rstd = torch.empty(c, dtype=torch.float32, device=dev)
        _lib.check(lib.gmlm_graphnorm_fwd(_ptr(x), _dtype_code(x, "graphnorm"), n, c, _ld(x), _ptr(colsum),
                                          _ptr(colsq), _ptr(weight), _ptr(bias), _ptr(mean_scale), float(eps),
                                          int(bool(fuse_gelu)), _ptr(y), c, _ptr(mean), _ptr(rstd), int(stat_rows),
                                          _stream(dev)), "graphnorm_fwd")
    return y, mean, rstd


def graphnorm_bwd_stats(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu):
    """Column sums the backward needs (sum dn, sum dn*xhat), fp64, local rows only."""
    lib = _lib.load()
    x = _rowmajor(x)
    gy = _rowmajor(gy)
    if gy.dtype != x.dtype:
        gy = gy.to(x.dtype)
    n, c = x.shape
    dev = x.device
    weight, bias, mean_scale = _f32c(weight), _f32c(bias), _f32c(mean_scale)
    with torch.cuda.device(dev):
        s1 = torch.empty(c, dtype=torch.float64, device=dev)
        s2 = torch.empty(c, dtype=torch.float64, device=dev)
        ws_bytes = lib.gmlm_colstats_workspace_bytes(n, c)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gmlm_graphnorm_bwd_stats(_ptr(x), _ptr(gy), _dtype_code(x, "graphnorm_bwd"), n, c, _ld(x),
                                                _ld(gy), _ptr(mean), _ptr(rstd), _ptr(weight), _ptr(bias),
                                                _ptr(mean_scale), int(bool(fuse_gelu)), _ptr(s1), _ptr(s2), _ptr(ws),
                                                ws_bytes, _stream(dev)), "graphnorm_bwd_stats")
    return s1, s2


def graphnorm_bwd_apply(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, s1, s2, need_gx, stat_rows=0):
    lib = _lib.load()
    x = _rowmajor(x)
    gy = _rowmajor(gy)
    if gy.dtype != x.dtype:
        gy = gy.to(x.dtype)
    n, c = x.shape
    dev = x.device
    weight, bias, mean_scale = _f32c(weight), _f32c(bias), _f32c(mean_scale)
    with torch.cuda.device(dev):
        gx = torch.empty((n, c), dtype=x.dtype, device=dev) if need_gx else torch.empty(0, dtype=x.dtype, device=dev)
        gw = torch.empty(c, dtype=torch.float32, device=dev)
        gb = torch.empty(c, dtype=torch.float32, device=dev)
        gms = torch.empty(c, dtype=torch.float32, device=dev)
        _lib.check(lib.gmlm_graphnorm_bwd_apply(_ptr(x), _ptr(gy), _dtype_code(x, "graphnorm_bwd"), n, c, _ld(x),
                                                _ld(gy), _ptr(mean), _ptr(rstd), _ptr(weight), _ptr(bias),
                                                _ptr(mean_scale), int(bool(fuse_gelu)), _ptr(s1), _ptr(s2),
                                                _ptr(gx) if need_gx else C.c_void_p(0), c, _ptr(gw), _ptr(gb),
                                                _ptr(gms), int(stat_rows), _stream(dev)), "graphnorm_bwd_apply")
    return gx, gw, gb, gms


def _graphnorm_fwd(x, weight, bias, mean_scale, eps, fuse_gelu):
    colsum, colsq = _colstats(x)
    return graphnorm_apply_stats(x, colsum, colsq, weight, bias, mean_scale, eps, fuse_gelu)


def _graphnorm_bwd(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, need_gx):
    s1, s2 = graphnorm_bwd_stats(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu)
    return graphnorm_bwd_apply(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, s1, s2, need_gx)


# ------------------------------------------------------------------------------ A13 LayerNorm
def layer_norm_ok(x: torch.Tensor) -> bool:
    """Shapes the row-wise kernel takes: CUDA fp32/bf16 [N, C], C a multiple of the 16-byte pack, C <= 1024."""
    if not (x.is_cuda and x.dim() == 2 and x.dtype in _DT and x.size(0) >= 1):
        return False
    vec = 4 if x.dtype == torch.float32 else 8
    return x.size(1) % vec == 0 and 0 < x.size(1) <= _lib.load().gmlm_layernorm_max_channels(_DT[x.dtype])


def _layernorm_fwd(x, weight, bias, eps):
    lib = _lib.load()
    x = _rowmajor(x)
    n, c = x.shape
    dev = x.device
    weight, bias = _f32c(weight), _f32c(bias)
    with torch.cuda.device(dev):
        y = torch.empty((n, c), dtype=x.dtype, device=dev)
        mean = torch.empty(n, dtype=torch.float32, device=dev)
        rstd = torch.empty(n, dtype=torch.float32, device=dev)
        _lib.check(lib.gmlm_layernorm_fwd(_ptr(x), _dtype_code(x, "layernorm"), n, c, _ld(x), _ptr(weight), _ptr(bias),
                                          float(eps), _ptr(y), c, _ptr(mean), _ptr(rstd), _stream(dev)),
                   "layernorm_fwd")
    return y, mean, rstd


def _layernorm_bwd(x, gy, mean, rstd, weight, need_gx):
    lib = _lib.load()
    x = _rowmajor(x)
    gy = _rowmajor(gy)
    if gy.dtype != x.dtype:
        gy = gy.to(x.dtype)
    n, c = x.shape
    dev = x.device
    weight = _f32c(weight)
    with torch.cuda.device(dev):
        gx = torch.empty((n, c), dtype=x.dtype, device=dev) if need_gx else torch.empty(0, dtype=x.dtype, device=dev)
        gw = torch.empty(c, dtype=torch.float32, device=dev)
        gb = torch.empty(c, dtype=torch.float32, device=dev)
        ws_bytes = lib.gmlm_layernorm_bwd_workspace_bytes(n, c)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gmlm_layernorm_bwd(_ptr(x), _ptr(gy), _dtype_code(x, "layernorm_bwd"), n, c, _ld(x), _ld(gy),
                                          _ptr(weight), _ptr(mean), _ptr(rstd),
                                          _ptr(gx) if need_gx else C.c_void_p(0), c, _ptr(gw), _ptr(gb), _ptr(ws),
                                          ws_bytes, _stream(dev)), "layernorm_bwd")
    return gx, gw, gb


# ------------------------------------------------------------------------------ A11
def _mask_u8(mask: torch.Tensor, n: int) -> torch.Tensor:
    mask = mask.reshape(-1)
    if mask.numel() != n:
        raise _lib.GmlmError(f"mask has {mask.numel()} entries for {n} rows")
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return (mask != 0).view(torch.uint8)


def _soft_mask_fwd(x, mask, token, beta):
    lib = _lib.load()
    x = _rowmajor(x)
    n, f = x.shape
    dev = x.device
    m = _mask_u8(mask.to(dev), n)
    token = _f32c(token.to(dev)).reshape(-1)
    if token.numel() != f:
        raise _lib.GmlmError(f"mask token has {token.numel()} features, x has {f}")
    with torch.cuda.device(dev):
        y = torch.empty((n, f), dtype=x.dtype, device=dev)
        _lib.check(lib.gmlm_soft_mask_fwd(_ptr(x), _dtype_code(x, "soft_mask"), n, f, _ld(x), _ptr(m), _ptr(token),
                                          float(beta), _ptr(y), f, _stream(dev)), "soft_mask_fwd")
    return y


def _soft_mask_bwd(gy, mask, beta, need_gx):
    lib = _lib.load()
    gy = _rowmajor(gy)
    n, f = gy.shape
    dev = gy.device
    m = _mask_u8(mask.to(dev), n)
    with torch.cuda.device(dev):
        g_token = torch.empty(f, dtype=torch.float32, device=dev)
        gx = torch.empty((n, f), dtype=gy.dtype, device=dev) if need_gx else torch.empty(0, dtype=gy.dtype, device=dev)
        ws_bytes = lib.gmlm_soft_mask_bwd_workspace_bytes(n, f)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gmlm_soft_mask_bwd(_ptr(gy), _dtype_code(gy, "soft_mask_bwd"), n, f, _ld(gy), _ptr(m),
                                          float(beta), _ptr(g_token), _ptr(gx) if need_gx else C.c_void_p(0), f,
                                          _ptr(ws), ws_bytes, _stream(dev)), "soft_mask_bwd")
    return g_token, gx


_LIB.impl("degree_i32", _degree_i32, "CUDA")
_LIB.impl("edge_type_bucket", _edge_type_bucket, "CUDA")
_LIB.impl("spmm_csr", _spmm_csr, "CUDA")
_LIB.impl("colstats", _colstats, "CUDA")
_LIB.impl("graphnorm_fwd", _graphnorm_fwd, "CUDA")
_LIB.impl("graphnorm_bwd", _graphnorm_bwd, "CUDA")
_LIB.impl("layernorm_fwd", _layernorm_fwd, "CUDA")
_LIB.impl("layernorm_bwd", _layernorm_bwd, "CUDA")
_LIB.impl("soft_mask_fwd", _soft_mask_fwd, "CUDA")
_LIB.impl("soft_mask_bwd", _soft_mask_bwd, "CUDA")


# ------------------------------------------------------------------------------ differentiable custom ops
# torch.ops.gmlm.{rgcn_aggregate, plan_aggregate, graphnorm_fwd, layernorm_fwd, soft_mask_fwd} carry their
# backward (torch.library.register_autograd) and a fake implementation (register_fake), so they can be called,
# traced and differentiated as ordinary torch operators; the thin functions further down only validate arguments.
def csr_pack(csr: CSR):
    """CSR -> (Tensor?[8], [num_rows, hub_thresh]): rowptr, col, w, grp_row, hub_row, hub_chunk_ptr, chunk_beg, chunk_end."""
    return ([csr.rowptr, csr.col, csr.w, csr.grp_row, csr.hub_row, csr.hub_chunk_ptr, csr.chunk_beg, csr.chunk_end],
            [int(csr.num_rows), int(csr.hub_thresh)])


def _spmm_list(x, c, rows, thresh, mode):
    return _spmm_csr(x, c[0], c[1], c[2] if mode == _lib.AGG_WEIGHTED else None, rows, mode, c[3], thresh, c[4], c[5],
                     c[6], c[7])


def _rgcn_aggregate_impl(x, *a):
    fwd, meta = a[0:8], a[16]
    rows_f, thresh_f, _, _, n_nodes, n_slots = meta
    h = _spmm_list(x, fwd, rows_f, thresh_f, _lib.AGG_MEAN)                 # [N*S, F]
    return h.view(n_nodes, n_slots * x.size(1))


def _plan_aggregate_impl(rows, *a):
    fwd, meta = a[0:8], a[16]
    return _spmm_list(rows, fwd, meta[0], meta[1], _lib.AGG_WEIGHTED)


def _gemm_nt_op(a1, b, bias, a2, out_dtype):
    return gemm_nt(a1, b, bias=bias, a2=a2, out_dtype=out_dtype)


def _csr_build_op(row, col, rel, keep, num_rows, num_cols, num_relations, slot_of_rel, num_slots):
    """(rowptr int32 [num_rows*num_slots+1], col int32 [E], perm int32 [E], seg_of_edge int32 [E]) of the CSR keyed
    on ``row*num_slots + slot_of_rel[rel]`` (stable: original edge order inside a segment); A3.  ``keep`` (uint8 /
    bool [E], optional) is an edge-dropout mask fused into the build: only the first rowptr[-1] entries of col / perm
    are then the CSR."""
    lib = _lib.load()
    dev = row.device
    E = int(row.numel())
    row, col = row.contiguous(), col.contiguous()
    rel = rel.contiguous() if rel is not None else None
    rows_total = num_rows * num_slots
    with torch.cuda.device(dev):
        rowptr = torch.empty(rows_total + 1, dtype=torch.int32, device=dev)
        colv = torch.empty(E, dtype=torch.int32, device=dev)
        perm = torch.empty(E, dtype=torch.int32, device=dev)
        seg = torch.empty(E, dtype=torch.int32, device=dev)
        ws_bytes = lib.gmlm_csr_workspace_bytes(max(E, 1), rows_total)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        slots = (C.c_int32 * max(num_relations, 1))(*[int(v) for v in slot_of_rel]) if rel is not None else None
        keep_u8 = _mask_u8(keep, E) if keep is not None else None
        nnz = C.c_int64(0)
        # csr_build keys on `dst`; here the CSR row plays that role and `col` is the gathered id
        _lib.check(lib.gmlm_csr_build(_ptr(col), _ptr(row), _ptr(rel), _ptr(keep_u8), E, num_rows, num_cols,
                                      num_relations, slots, num_slots, _ptr(rowptr), _ptr(colv), _ptr(perm), _ptr(seg),
                                      C.byref(nnz) if keep is not None else None, _ptr(ws), ws_bytes, _stream(dev)),
                   "csr_build")
    return rowptr, colv, perm, seg


_LIB.impl("rgcn_aggregate", _rgcn_aggregate_impl, "CUDA")
_LIB.impl("plan_aggregate", _plan_aggregate_impl, "CUDA")
_LIB.impl("gemm_nt", _gemm_nt_op, "CUDA")
_LIB.impl("csr_build", _csr_build_op, "CUDA")


# ---- fake (meta) implementations: shapes and dtypes only
@torch.library.register_fake("gmlm::degree_i32")
def _(index, num_nodes):
    return index.new_empty((num_nodes,), dtype=torch.int32)


@torch.library.register_fake("gmlm::edge_type_bucket")
def _(src, deg, bounds):
    return src.new_empty((src.numel(),), dtype=torch.int64)


@torch.library.register_fake("gmlm::spmm_csr")
def _(x, rowptr, col, w, num_rows, mode, grp_row, hub_thresh, hub_row, hub_chunk_ptr, chunk_beg, chunk_end):
    return x.new_empty((num_rows, x.size(1)))


@torch.library.register_fake("gmlm::colstats")
def _(x):
    return x.new_empty((x.size(1),), dtype=torch.float64), x.new_empty((x.size(1),), dtype=torch.float64)


@torch.library.register_fake("gmlm::graphnorm_fwd")
def _(x, weight, bias, mean_scale, eps, fuse_gelu):
    c = x.size(1)
    return torch.empty_like(x), x.new_empty((c,), dtype=torch.float32), x.new_empty((c,), dtype=torch.float32)


@torch.library.register_fake("gmlm::graphnorm_bwd")
def _(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, need_gx):
    c = x.size(1)
    f = lambda: x.new_empty((c,), dtype=torch.float32)   # noqa: E731
    return (torch.empty_like(x) if need_gx else x.new_empty((0,))), f(), f(), f()


@torch.library.register_fake("gmlm::layernorm_fwd")
def _(x, weight, bias, eps):
    n = x.size(0)
    return torch.empty_like(x), x.new_empty((n,), dtype=torch.float32), x.new_empty((n,), dtype=torch.float32)


@torch.library.register_fake("gmlm::layernorm_bwd")
def _(x, gy, mean, rstd, weight, need_gx):
    c = x.size(1)
    return ((torch.empty_like(x) if need_gx else x.new_empty((0,))), x.new_empty((c,), dtype=torch.float32),
            x.new_empty((c,), dtype=torch.float32))


@torch.library.register_fake("gmlm::soft_mask_fwd")
def _(x, mask, token, beta):
    return torch.empty_like(x)


@torch.library.register_fake("gmlm::soft_mask_bwd")
def _(gy, mask, beta, need_gx):
    return gy.new_empty((gy.size(1),), dtype=torch.float32), (torch.empty_like(gy) if need_gx else gy.new_empty((0,)))


@torch.library.register_fake("gmlm::rgcn_aggregate")
def _(x, *a):
    meta = a[16]
    return x.new_empty((meta[4], meta[5] * x.size(1)))


@torch.library.register_fake("gmlm::plan_aggregate")
def _(rows, *a):
    return rows.new_empty((a[16][0], rows.size(1)))


@torch.library.register_fake("gmlm::gemm_nt")
def _(a1, b, bias, a2, out_dtype):
    return a1.new_empty((a1.size(0), b.size(0)), dtype=out_dtype)


@torch.library.register_fake("gmlm::csr_build")
def _(row, col, rel, keep, num_rows, num_cols, num_relations, slot_of_rel, num_slots):
    e = row.numel()
    i32 = lambda k: row.new_empty((k,), dtype=torch.int32)   # noqa: E731
    return i32(num_rows * num_slots + 1), i32(e), i32(e), i32(e)


# ---- autograd formulas
def _rgcn_aggregate_setup(ctx, inputs, output):
    x = inputs[0]
    ctx.bwd, ctx.meta, ctx.x_dtype = inputs[9:17], inputs[17], x.dtype


def _rgcn_aggregate_backward(ctx, gh):
    _, _, rows_b, thresh_b, n_nodes, n_slots = ctx.meta
    gh = gh.contiguous()
    if gh.dtype != ctx.x_dtype:
        gh = gh.to(ctx.x_dtype)
    feat = gh.size(1) // n_slots
    b = ctx.bwd
    gx = torch.ops.gmlm.spmm_csr(gh.view(n_nodes * n_slots, feat), b[0], b[1], b[2], rows_b, _lib.AGG_WEIGHTED, b[3],
                                 thresh_b, b[4], b[5], b[6], b[7])            # A14: gather on the transposed CSR
    return (gx,) + (None,) * 17


def _plan_aggregate_setup(ctx, inputs, output):
    ctx.bwd, ctx.meta, ctx.in_dtype = inputs[9:17], inputs[17], inputs[0].dtype


def _plan_aggregate_backward(ctx, g):
    g = g.contiguous()
    if g.dtype != ctx.in_dtype:
        g = g.to(ctx.in_dtype)
    b = ctx.bwd
    return (torch.ops.gmlm.spmm_csr(g, b[0], b[1], b[2], ctx.meta[2], _lib.AGG_WEIGHTED, b[3], ctx.meta[3], b[4], b[5],
                                    b[6], b[7]),) + (None,) * 17


def _graphnorm_setup(ctx, inputs, output):
    x, weight, bias, mean_scale, eps, fuse_gelu = inputs
    _, mean, rstd = output
    ctx.save_for_backward(x, mean, rstd, weight, bias, mean_scale)
    ctx.fuse_gelu = bool(fuse_gelu)


def _graphnorm_backward(ctx, gy, g_mean, g_rstd):
    x, mean, rstd, weight, bias, mean_scale = ctx.saved_tensors
    gx, gw, gb, gms = torch.ops.gmlm.graphnorm_bwd(x, gy, mean, rstd, weight, bias, mean_scale, ctx.fuse_gelu,
                                                   ctx.needs_input_grad[0])
    return (gx if ctx.needs_input_grad[0] else None,
            gw.to(weight.dtype) if ctx.needs_input_grad[1] else None,
            gb.to(bias.dtype) if ctx.needs_input_grad[2] else None,
            gms.to(mean_scale.dtype) if ctx.needs_input_grad[3] else None, None, None)


def _layernorm_setup(ctx, inputs, output):
    x, weight, bias, eps = inputs
    _, mean, rstd = output
    ctx.save_for_backward(x, mean, rstd, weight)
    ctx.bias_dtype = bias.dtype


def _layernorm_backward(ctx, gy, g_mean, g_rstd):
    x, mean, rstd, weight = ctx.saved_tensors
    gx, gw, gb = torch.ops.gmlm.layernorm_bwd(x, gy, mean, rstd, weight, ctx.needs_input_grad[0])
    return (gx if ctx.needs_input_grad[0] else None, gw.to(weight.dtype) if ctx.needs_input_grad[1] else None,
            gb.to(ctx.bias_dtype) if ctx.needs_input_grad[2] else None, None)


def _soft_mask_setup(ctx, inputs, output):
    x, mask, token, beta = inputs
    ctx.save_for_backward(mask)
    ctx.beta, ctx.token_shape, ctx.token_dtype = float(beta), token.shape, token.dtype


def _soft_mask_backward(ctx, gy):
    (mask,) = ctx.saved_tensors
    g_token, gx = torch.ops.gmlm.soft_mask_bwd(gy, mask, ctx.beta, ctx.needs_input_grad[0])
    return (gx if ctx.needs_input_grad[0] else None, None,
            g_token.view(ctx.token_shape).to(ctx.token_dtype) if ctx.needs_input_grad[2] else None, None)


torch.library.register_autograd("gmlm::rgcn_aggregate", _rgcn_aggregate_backward, setup_context=_rgcn_aggregate_setup)
torch.library.register_autograd("gmlm::plan_aggregate", _plan_aggregate_backward, setup_context=_plan_aggregate_setup)
torch.library.register_autograd("gmlm::graphnorm_fwd", _graphnorm_backward, setup_context=_graphnorm_setup)
torch.library.register_autograd("gmlm::layernorm_fwd", _layernorm_backward, setup_context=_layernorm_setup)
torch.library.register_autograd("gmlm::soft_mask_fwd", _soft_mask_backward, setup_context=_soft_mask_setup)


# ------------------------------------------------------------------------------ A6 dense transform (tcgen05)
def gemm_nt(a1: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor] = None, a2: Optional[torch.Tensor] = None,
            out_dtype: Optional[torch.dtype] = None, split: int = 0):
    """``[a1 | a2] @ b.T + bias`` on the tcgen05 tensor cores (bf16 or fp16 in, fp32 accumulate).

    a1 [M,K1], a2 [M,K2] (optional), b [N,K1+K2], all bf16 (or all fp16) with unit inner stride; returns C [M,N]
    (bf16 or fp32), or (C[:, :split], C[:, split:]) as two contiguous tensors when ``split`` > 0."""
    lib = _lib.load()
    _require_cuda(a1, "a1")
    if a1.dtype not in (torch.bfloat16, torch.float16) or b.dtype != a1.dtype or (a2 is not None and a2.dtype != a1.dtype):
        raise _lib.GmlmError("gemm_nt: operands must all be bfloat16 or all be float16")
    a1, b = _rowmajor(a1), _rowmajor(b)
    a2 = _rowmajor(a2) if a2 is not None else None
    m, k1 = a1.shape
    k2 = int(a2.size(1)) if a2 is not None else 0
    n = int(b.size(0))
    if b.size(1) != k1 + k2 or (a2 is not None and a2.size(0) != m):
        raise _lib.GmlmError(f"gemm_nt: shape mismatch a1 {tuple(a1.shape)} a2 {None if a2 is None else tuple(a2.shape)} "
                             f"b {tuple(b.shape)}")
    out_dtype = out_dtype or (torch.bfloat16 if a1.dtype == torch.bfloat16 else torch.float32)
    if out_dtype not in _DT:
        raise _lib.GmlmError("gemm_nt: output must be float32 or bfloat16")
    code = _DT[out_dtype]
    in_code = _lib.BF16 if a1.dtype == torch.bfloat16 else _lib.F16
    dev = a1.device
    bias32 = bias.detach().float().contiguous() if bias is not None else None
    with torch.cuda.device(dev):
        if split and 0 < split < n:
            c1 = torch.empty((m, split), dtype=out_dtype, device=dev)
            c2 = torch.empty((m, n - split), dtype=out_dtype, device=dev)
        else:
            split = 0
            c1 = torch.empty((m, n), dtype=out_dtype, device=dev)
            c2 = None
        _lib.check(lib.gmlm_gemm_nt(_ptr(a1), _ld(a1), k1, _ptr(a2), _ld(a2) if a2 is not None else 0, k2,
                                    _ptr(b), _ld(b), _ptr(bias32), m, n, _ptr(c1), c1.size(1), split,
                                    _ptr(c2), c2.size(1) if c2 is not None else 0, in_code, code, _stream(dev)),
                   "gemm_nt")
    return (c1, c2) if c2 is not None else c1


class _RGCNTransform(torch.autograd.Function):
    """out = [h | x] @ [w ; root] + bias as ONE tcgen05 GEMM (A6; 0.81 ms vs 1.03 ms for the two cuBLAS calls at
    M=2M, K=1280, N=64).  Backward: [dh | dx] = g @ [w ; root]^T is one more launch of the same persistent kernel,
    its two outputs written through two tensor maps so that dh lands contiguous for the transposed aggregation;
    dW = h^T g and droot = x^T g are reductions over all nodes and are produced in fp32.

    ``op_dtype`` is the operand type of the GEMMs: bf16 for the bf16 pipeline, fp16 under ``torch.amp.autocast``
    (what the reference's matmuls run in, main.py:446,543); h and x are cast once and saved in that type."""

    @staticmethod
    def forward(ctx, h, x, w, root, bias, out_dtype, op_dtype):
        hq = h if h.dtype == op_dtype else h.to(op_dtype)
        xq = x if x.dtype == op_dtype else x.to(op_dtype)
        wc = torch.cat([w, root], dim=0).detach().to(op_dtype)          # [K1+K2, Fo]
        out = gemm_nt(hq, wc.t().contiguous(), bias=bias, a2=xq, out_dtype=out_dtype)
        ctx.save_for_backward(hq, xq, wc)
        ctx.k1 = h.size(1)
        ctx.dtypes = (w.dtype, root.dtype, None if bias is None else bias.dtype, h.dtype, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        h, x, wc = ctx.saved_tensors
        k1, k2 = ctx.k1, x.size(1)
        gb = g.to(wc.dtype).contiguous()
        fo = gb.size(1)
        dh = dx = dw = droot = dbias = None
        need_h, need_x = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_h or need_x:
            out_dt = torch.bfloat16 if ctx.dtypes[3] == torch.bfloat16 else torch.float32
            if fo % 64 == 0 and k1 % 32 == 0 and k2 % 32 == 0 and ctx.dtypes[3] == ctx.dtypes[4]:
                if need_h and need_x:
                    dh, dx = gemm_nt(gb, wc, out_dtype=out_dt, split=k1)      # [M, Fo] x [K1+K2, Fo]^T
                elif need_h:
                    dh = gemm_nt(gb, wc[:k1], out_dtype=out_dt)
                else:
                    dx = gemm_nt(gb, wc[k1:], out_dtype=out_dt)
            else:
                if need_h:
                    dh = (gb @ wc[:k1].t()).to(ctx.dtypes[3])
                if need_x:
                    dx = (gb @ wc[k1:].t()).to(ctx.dtypes[4])
        if ctx.needs_input_grad[2]:
            dw = _mm_f32(h.t(), gb).to(ctx.dtypes[0])
        if ctx.needs_input_grad[3]:
            droot = _mm_f32(x.t(), gb).to(ctx.dtypes[1])
        if ctx.needs_input_grad[4]:
            dbias = torch.ops.gmlm.colstats(gb.float() if gb.dtype == torch.float16 else gb)[0].to(ctx.dtypes[2])
        return dh, dx, dw, droot, dbias, None, None


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a @ b with an fp32 result for a parameter gradient (a reduction over up to millions of rows must not be
    rounded to 8 mantissa bits before it reaches the fp32 parameter)."""
    if a.dtype == torch.float32:
        return a @ b
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return (a @ b).float()


def linear_nt_ok(x: torch.Tensor, n_out: int) -> bool:
    """Shapes ``linear_nt`` runs on the tcgen05 GEMM: bf16 CUDA activations, K a multiple of 64, N of 32."""
    return (x.is_cuda and x.dim() == 2 and x.dtype == torch.bfloat16 and x.size(1) % 64 == 0 and x.size(1) > 0
            and n_out % 32 == 0 and n_out > 0)


class _LinearNT(torch.autograd.Function):
    """y = x @ wt.T + bias on the tcgen05 GEMM (bf16 operands, fp32 accumulation).  Backward: dx on the same
    kernel when its shape fits (K = n_out a multiple of 64), weight gradient as an fp32 reduction."""

    @staticmethod
    def forward(ctx, x, wt, bias, out_dtype):
        wtb = wt.detach().to(torch.bfloat16).contiguous()                 # [N_out, K]
        y = gemm_nt(x, wtb, bias=bias, out_dtype=out_dtype)
        ctx.save_for_backward(x, wtb)
        ctx.dtypes = (wt.dtype, None if bias is None else bias.dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        x, wtb = ctx.saved_tensors
        gb = g.to(torch.bfloat16).contiguous()
        dx = dwt = dbias = None
        if ctx.needs_input_grad[0]:
            if wtb.size(0) % 64 == 0 and wtb.size(1) % 32 == 0:
                dx = gemm_nt(gb, wtb.t().contiguous())                     # [M, N_out] x [K, N_out]^T
            else:
                dx = gb @ wtb
        if ctx.needs_input_grad[1]:
            dwt = _mm_f32(gb.t(), x).to(ctx.dtypes[0])
        if ctx.needs_input_grad[2]:
            dbias = torch.ops.gmlm.colstats(gb)[0].to(ctx.dtypes[1])       # one pass, fp64 accumulate, deterministic
        return dx, dwt, dbias, None


def linear_nt(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor] = None,
              out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``F.linear(x, wt, bias)`` for the bf16 pipeline (``linear_nt_ok``) on the tcgen05 GEMM."""
    return _LinearNT.apply(x, wt, bias, out_dtype or x.dtype)


def rgcn_transform_first(x: torch.Tensor, graph: RelGraph, w_live: torch.Tensor, root: torch.Tensor,
                         bias: Optional[torch.Tensor], out_dtype: torch.dtype) -> torch.Tensor:
    """RGCNConv as TRANSFORM-then-aggregate (A5+A6 fused by linearity of the mean; include/gmlm_b200.h
    ``gmlm_dst_plan``):  Z = x @ [W_0 | .. | W_{S-1} | root] (+ bias on the root slab),
    out[i] = sum_e w_e Z[src_e, slot_e] + Z[i, S].  The [N, S*Fi] matrix H of the aggregate-first form is never
    materialised and the gather moves Fo-wide rows instead of Fi-wide ones.

    x [num_src, Fi]; w_live [S, Fi, Fo] (composed weights of the populated relations); root [Fi, Fo]."""
    _require_cuda(x, "x")
    S, fi, fo = w_live.shape
    if S != graph.num_slots or x.size(0) != graph.num_src or x.size(1) != fi:
        raise _lib.GmlmError("rgcn_transform_first: shape mismatch")
    fplan, bplan = graph.dst_plan()
    wcat_t = torch.cat([w_live.permute(0, 2, 1).reshape(S * fo, fi), root.t()], dim=0)     # [(S+1)*Fo, Fi]
    bias_cat = None
    if bias is not None:
        bias_cat = torch.cat([bias.new_zeros(S * fo), bias])
    if linear_nt_ok(x, (S + 1) * fo) and not torch.is_autocast_enabled("cuda"):
        z = linear_nt(x, wcat_t, bias_cat)
    else:
        z = torch.nn.functional.linear(x, wcat_t if torch.is_autocast_enabled("cuda") else wcat_t.to(x.dtype),
                                       None if bias_cat is None else
                                       (bias_cat if torch.is_autocast_enabled("cuda") else bias_cat.to(x.dtype)))
        if z.dtype not in _DT:
            z = z.float()                                   # autocast produced fp16: the kernels take fp32 / bf16
    fl, fm = csr_pack(fplan)
    bl, bm = csr_pack(bplan)
    out = torch.ops.gmlm.plan_aggregate(z.view(graph.num_src * (S + 1), fo), *fl, *bl, fm + bm)
    return out if out.dtype == out_dtype else out.to(out_dtype)


def rgcn_transform_ok(h: torch.Tensor, x: torch.Tensor, fo: int, op_dtype: Optional[torch.dtype] = None) -> bool:
    """Shapes the tcgen05 path covers: K blocks of 64, Fo a multiple of 32; activations bf16 (bf16 pipeline) or
    anything castable when an explicit fp16 / bf16 operand type is given (autocast)."""
    if not (h.is_cuda and h.size(1) % 64 == 0 and x.size(1) % 64 == 0 and fo % 32 == 0 and h.size(1) > 0):
        return False
    if op_dtype is None:
        return h.dtype == torch.bfloat16 and x.dtype == torch.bfloat16
    return op_dtype in (torch.bfloat16, torch.float16)


def rgcn_transform(h, x, w, root, bias, out_dtype, op_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    return _RGCNTransform.apply(h, x, w, root, bias, out_dtype, op_dtype or h.dtype)


# ------------------------------------------------------------------------------ halo pack / unpack
def gather_rows(x: torch.Tensor, ids: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[k] = x[ids[k]] (pack the rows a peer needs)."""
    lib = _lib.load()
    _require_cuda(x, "x")
    x = _rowmajor(x)
    ids = ids.contiguous()
    n, feat = int(ids.numel()), int(x.size(1))
    with torch.cuda.device(x.device):
        if out is None:
            out = torch.empty((n, feat), dtype=x.dtype, device=x.device)
        _lib.check(lib.gmlm_gather_rows(_ptr(x), _dtype_code(x, "gather_rows"), feat, _ld(x), _ptr(ids), n, _ptr(out),
                                        _ld(out) if n > 1 else feat, _stream(x.device)), "gather_rows")
    return out


def scatter_add_rows_(dst: torch.Tensor, ids: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """dst[ids[k]] += src[k] in place; ``ids`` must be unique (no atomics, deterministic)."""
    lib = _lib.load()
    _require_cuda(dst, "dst")
    if dst.dim() != 2 or dst.stride(1) != 1 or src.dim() != 2 or src.stride(1) != 1:
        raise _lib.GmlmError("scatter_add_rows_: row-major 2-D tensors required")
    ids = ids.contiguous()
    n, feat = int(ids.numel()), int(dst.size(1))
    with torch.cuda.device(dst.device):
        _lib.check(lib.gmlm_scatter_add_rows(_ptr(dst), _dtype_code(dst, "scatter_add_rows"), feat, _ld(dst), _ptr(ids),
                                             n, _ptr(src), _ld(src) if n > 1 else feat, _stream(dst.device)),
                   "scatter_add_rows")
    return dst


# ------------------------------------------------------------------------------ public functional API
def degree(index: torch.Tensor, num_nodes: Optional[int] = None, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Drop-in for ``torch_geometric.utils.degree`` (``/root/reference/main.py:7``; called
    ``main.py:65,256``): float32 [N] occurrence counts of ``index``."""
    _require_cuda(index, "index")
    if num_nodes is None:  # upstream maybe_num_nodes (one host sync)
        num_nodes = int(index.max()) + 1 if index.numel() > 0 else 0
    deg = torch.ops.gmlm.degree_i32(index, int(num_nodes))
    return deg.to(dtype or torch.get_default_dtype())


def edge_type_from_degree(edge_index: torch.Tensor, num_nodes: int, bounds: Sequence[int] = (2, 5, 10)) -> torch.Tensor:
    """The reference's per-edge loop ``main.py:253-267`` as one kernel: relation id =
    bucket of the source node's out-degree (<=2, <=5, <=10, else)."""
    _require_cuda(edge_index, "edge_index")
    src = edge_index[0]
    deg = torch.ops.gmlm.degree_i32(src, int(num_nodes))
    return torch.ops.gmlm.edge_type_bucket(src, deg, list(bounds))


def rgcn_aggregate(x: torch.Tensor, graph: RelGraph) -> torch.Tensor:
    """Per-(dst, relation) mean of source rows: ``[N, F] -> [N, S*F]`` (S = populated relations).
    Forward = A5, backward = A14 (gather on the transposed CSR with 1/count folded in)."""
    _require_cuda(x, "x")
    if x.size(0) != graph.num_src:
        raise _lib.GmlmError(f"x has {x.size(0)} rows, graph has {graph.num_src} source nodes")
    fl, fm = csr_pack(graph.fwd)
    bl, bm = csr_pack(graph.bwd)
    return torch.ops.gmlm.rgcn_aggregate(x, *fl, *bl, fm + bm + [int(graph.num_nodes), int(graph.num_slots)])


def graph_norm(x, weight, bias, mean_scale, eps: float = 1e-5, fuse_gelu: bool = False) -> torch.Tensor:
    _require_cuda(x, "x")
    return torch.ops.gmlm.graphnorm_fwd(x, weight, bias, mean_scale, float(eps), bool(fuse_gelu))[0]


def layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """``nn.LayerNorm(C)`` over the last dimension of a 2-D CUDA tensor (the ``self.layer_norm(fused)`` of
    ``MultiScaleFusion``, ``/root/reference/main.py:171,180``): one read + one write forward, two reads +
    one write backward, deterministic parameter gradients."""
    _require_cuda(x, "x")
    if not layer_norm_ok(x):
        raise _lib.GmlmError(f"layer_norm: unsupported input {tuple(x.shape)} {x.dtype} (see layer_norm_ok)")
    return torch.ops.gmlm.layernorm_fwd(x, weight, bias, float(eps))[0]


def soft_masking_gnn_input(x: torch.Tensor, gnn_perturb_mask: torch.Tensor, mask_token_embed: torch.Tensor,
                           beta: float = 0.7) -> torch.Tensor:
    """Drop-in for ``soft_masking_gnn_input`` (``/root/reference/main.py:92-99``): one fused pass,
    no clone + boolean-index round trip and no ``.any()`` host sync."""
    _require_cuda(x, "x")
    return torch.ops.gmlm.soft_mask_fwd(x, gnn_perturb_mask, mask_token_embed, float(beta))
