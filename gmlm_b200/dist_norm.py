"""GraphNorm over a row-partitioned node matrix ("GraphNorm (A7) = all_reduce(SUM) of the
column sums"; the reference has no distributed code — ``/root/reference/main.py:43`` is one device).

Every rank holds the rows of its own nodes.  The statistics are whole-graph quantities, so the two
column-sum vectors each pass of the single-GPU implementation reduces locally are summed over the ranks
before they are used — one ``all_reduce`` of ``[2C]`` fp64 values in the forward (Σx, Σx²) and one in the
backward (Σ dn, Σ dn·ô):

    forward :  (s, q) = colstats(x_local);  all_reduce;  y = normalise(x_local; s, q, N_global)
    backward:  (s1, s2) = bwd_stats(x_local, gy_local);  all_reduce;  gx = bwd_apply(...; s1, s2, N_global)

The C ABI takes ONE row count (it is both the loop bound and the divisor of the statistics), so the reduced
sums are rescaled by ``N_local / N_global`` before the call: ``s' / N_local == s / N_global``, every formula in
``csrc/graphnorm.cu`` is linear in the sums, and the parameter gradients (which the kernels return as plain
sums) are scaled back.  The parameter gradients that come out are the WHOLE-GRAPH gradients, identical on
every rank: they must not be all-reduced again by a data-parallel wrapper.

``backend`` abstracts the four kernel entry points so that the rank logic can be exercised on CPU with gloo
against the oracle (tests/test_partition.py injects an fp64 restatement of the kernels' formulas); the product
backend is the CUDA library and refuses CPU tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist


class CudaGraphNormBackend:
    """The four entry points of include/gmlm_b200.h behind A7 (no CPU path)."""

    @staticmethod
    def colstats(x):
        return torch.ops.gmlm.colstats(x)

    @staticmethod
    def fwd(x, colsum, colsq, weight, bias, mean_scale, eps, fuse_gelu):
        from . import _lib
        from .ops import _dtype_code, _f32c, _ld, _ptr, _rowmajor, _stream
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
                                              int(bool(fuse_gelu)), _ptr(y), c, _ptr(mean), _ptr(rstd), _stream(dev)),
                       "graphnorm_fwd")
        return y, mean, rstd

    @staticmethod
    def bwd_stats(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu):
        from . import _lib
        from .ops import _dtype_code, _f32c, _ld, _ptr, _rowmajor, _stream
        lib = _lib.load()
        x, gy = _rowmajor(x), _rowmajor(gy)
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
                                                    _ptr(mean_scale), int(bool(fuse_gelu)), _ptr(s1), _ptr(s2),
                                                    _ptr(ws), ws_bytes, _stream(dev)), "graphnorm_bwd_stats")
        return s1, s2

    @staticmethod
    def bwd_apply(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, s1, s2, need_gx):
        from . import _lib
        from .ops import _dtype_code, _f32c, _ld, _ptr, _rowmajor, _stream
        lib = _lib.load()
        x, gy = _rowmajor(x), _rowmajor(gy)
        n, c = x.shape
        dev = x.device
        weight, bias, mean_scale = _f32c(weight), _f32c(bias), _f32c(mean_scale)
        with torch.cuda.device(dev):
            gx = torch.empty((n, c), dtype=x.dtype, device=dev) if need_gx else None
            gw = torch.empty(c, dtype=torch.float32, device=dev)
            gb = torch.empty(c, dtype=torch.float32, device=dev)
            gms = torch.empty(c, dtype=torch.float32, device=dev)
            _lib.check(lib.gmlm_graphnorm_bwd_apply(_ptr(x), _ptr(gy), _dtype_code(x, "graphnorm_bwd"), n, c, _ld(x),
                                                    _ld(gy), _ptr(mean), _ptr(rstd), _ptr(weight), _ptr(bias),
                                                    _ptr(mean_scale), int(bool(fuse_gelu)), _ptr(s1), _ptr(s2),
                                                    _ptr(gx) if need_gx else C.c_void_p(0), c, _ptr(gw), _ptr(gb),
                                                    _ptr(gms), _stream(dev)), "graphnorm_bwd_apply")
        return gx, gw, gb, gms


class _PartitionedGraphNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mean_scale, n_global, eps, fuse_gelu, group, backend):
        n_local = x.size(0)
        s, q = backend.colstats(x)
        sums = torch.stack([s, q]).to(torch.float64)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        scale = float(n_local) / float(n_global)
        y, mean, rstd = backend.fwd(x, (sums[0] * scale).contiguous(), (sums[1] * scale).contiguous(), weight, bias,
                                    mean_scale, eps, fuse_gelu)
        ctx.save_for_backward(x, mean, rstd, weight, bias, mean_scale)
        ctx.fuse_gelu, ctx.group, ctx.backend, ctx.scale = bool(fuse_gelu), group, backend, scale
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mean, rstd, weight, bias, mean_scale = ctx.saved_tensors
        be = ctx.backend
        if gy.dtype != x.dtype:
            gy = gy.to(x.dtype)
        s1, s2 = be.bwd_stats(x, gy, mean, rstd, weight, bias, mean_scale, ctx.fuse_gelu)
        sums = torch.stack([s1, s2]).to(torch.float64)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=ctx.group)
        gx, gw, gb, gms = be.bwd_apply(x, gy, mean, rstd, weight, bias, mean_scale, ctx.fuse_gelu,
                                       (sums[0] * ctx.scale).contiguous(), (sums[1] * ctx.scale).contiguous(),
                                       ctx.needs_input_grad[0])
        inv = 1.0 / ctx.scale                       # the kernels return plain sums: undo the N_local/N_global rescale
        return (gx if ctx.needs_input_grad[0] else None,
                (gw * inv).to(weight.dtype) if ctx.needs_input_grad[1] else None,
                (gb * inv).to(bias.dtype) if ctx.needs_input_grad[2] else None,
                (gms * inv).to(mean_scale.dtype) if ctx.needs_input_grad[3] else None,
                None, None, None, None, None)


def partitioned_graph_norm(x_local: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                           mean_scale: torch.Tensor, num_nodes_global: int, eps: float = 1e-5,
                           fuse_gelu: bool = False, group=None, backend: Optional[object] = None) -> torch.Tensor:
    """GraphNorm(batch=None) (``main.py:273``) of the whole graph, computed on this rank's rows.
    Every rank must call it (two collectives: one forward, one backward).  Returned parameter gradients are the
    whole-graph gradients on every rank."""
    if x_local.size(0) < 1:
        raise ValueError("partitioned_graph_norm: every rank needs at least one row")
    return _PartitionedGraphNorm.apply(x_local, weight, bias, mean_scale, int(num_nodes_global), float(eps),
                                       bool(fuse_gelu), group, backend or CudaGraphNormBackend)
