"""GraphNorm over a row-partitioned node matrix ("GraphNorm (A7) = all_reduce(SUM) of the
column sums"; the reference has no distributed code — ``/root/reference/main.py:43`` is one device).

Every rank holds the rows of its own nodes.  The statistics are whole-graph quantities, so the two
column-sum vectors each pass of the single-GPU implementation reduces locally are summed over the ranks
before they are used — one ``all_reduce`` of ``[2C]`` fp64 values in the forward (Σx, Σx²) and one in the
backward (Σ dn, Σ dn·ô):

    forward :  (s, q) = colstats(x_local);  all_reduce;  y = normalise(x_local; s, q, N_global)
    backward:  (s1, s2) = bwd_stats(x_local, gy_local);  all_reduce;  gx = bwd_apply(...; s1, s2, N_global)

The C ABI takes the row count of the local block and, separately, ``stat_rows`` = the row count the sums
cover (include/gmlm_b200.h), so the reduced sums go in unchanged.  The parameter gradients that come out are
the WHOLE-GRAPH gradients, identical on every rank: they must not be all-reduced again by a data-parallel
wrapper.  A rank with zero rows takes part in both collectives with zero sums and returns an empty block, so
a degenerate partition cannot dead-lock the others.

``backend`` abstracts the four kernel entry points so that the rank logic can be exercised on CPU with gloo
against the oracle (tests/test_partition.py injects an fp64 restatement of the kernels' formulas); the product
backend is the CUDA library and refuses CPU tensors.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class CudaGraphNormBackend:
    """The four entry points of include/gmlm_b200.h behind A7 (no CPU path): the same wrappers the
    single-GPU op uses (gmlm_b200/ops.py)."""

    @staticmethod
    def colstats(x):
        return torch.ops.gmlm.colstats(x)

    @staticmethod
    def fwd(x, colsum, colsq, weight, bias, mean_scale, eps, fuse_gelu, stat_rows):
        from .ops import graphnorm_apply_stats
        return graphnorm_apply_stats(x, colsum, colsq, weight, bias, mean_scale, eps, fuse_gelu, stat_rows)

    @staticmethod
    def bwd_stats(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu):
        from .ops import graphnorm_bwd_stats
        return graphnorm_bwd_stats(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu)

    @staticmethod
    def bwd_apply(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, s1, s2, need_gx, stat_rows):
        from .ops import graphnorm_bwd_apply
        gx, gw, gb, gms = graphnorm_bwd_apply(x, gy, mean, rstd, weight, bias, mean_scale, fuse_gelu, s1, s2, need_gx,
                                              stat_rows)
        return (gx if need_gx else None), gw, gb, gms


class _PartitionedGraphNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, mean_scale, n_global, eps, fuse_gelu, group, backend):
        s, q = backend.colstats(x)
        sums = torch.stack([s, q]).to(torch.float64)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        y, mean, rstd = backend.fwd(x, sums[0].contiguous(), sums[1].contiguous(), weight, bias, mean_scale, eps,
                                    fuse_gelu, n_global)
        ctx.save_for_backward(x, mean, rstd, weight, bias, mean_scale)
        ctx.fuse_gelu, ctx.group, ctx.backend, ctx.n_global = bool(fuse_gelu), group, backend, n_global
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
                                       sums[0].contiguous(), sums[1].contiguous(), ctx.needs_input_grad[0],
                                       ctx.n_global)
        return (gx if ctx.needs_input_grad[0] else None,
                gw.to(weight.dtype) if ctx.needs_input_grad[1] else None,
                gb.to(bias.dtype) if ctx.needs_input_grad[2] else None,
                gms.to(mean_scale.dtype) if ctx.needs_input_grad[3] else None,
                None, None, None, None, None)


def partitioned_graph_norm(x_local: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                           mean_scale: torch.Tensor, num_nodes_global: int, eps: float = 1e-5,
                           fuse_gelu: bool = False, group=None, backend: Optional[object] = None) -> torch.Tensor:
    """GraphNorm(batch=None) (``main.py:273``) of the whole graph, computed on this rank's rows.
    Every rank must call it (two collectives: one forward, one backward), a rank without rows included.
    Returned parameter gradients are the whole-graph gradients on every rank."""
    if int(num_nodes_global) < max(1, x_local.size(0)):
        raise ValueError("partitioned_graph_norm: num_nodes_global must cover the local rows (and be >= 1)")
    return _PartitionedGraphNorm.apply(x_local, weight, bias, mean_scale, int(num_nodes_global), float(eps),
                                       bool(fuse_gelu), group, backend or CudaGraphNormBackend)
