"""The two aggregation variants north_star names beside the RGCN mean: GCN symmetric-degree
normalisation and GAT edge-softmax (row A9, BASELINE.json configs[2]).
They have **no counterpart in /root/reference** (it uses only RGCNConv); semantics follow upstream
``GCNConv`` / ``GATConv`` defaults as restated in ``oracle/pyg_ref.py`` ("parity unpinned").

Both run on one dst-keyed CSR with self-loops normalised the upstream way (existing self-loops
removed, one loop per node appended) and reuse the aggregation kernel ``gmlm_spmm_csr`` in
weighted mode — scalar weights for GCN, one weight column per head for GAT — forward on the
CSR, backward on its transpose.  The per-edge scalars come from ``csrc/gat.cu``.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .graph import CSR, _ptr, _require_cuda, _stream, _tensor_key, build_csr, transpose_csr
from .nn import glorot_
from .ops import _dtype_code, _ld, _rowmajor, spmm


@dataclass
class LoopGraph:
    """dst-keyed CSR of ``edge_index`` with upstream's self-loop normalisation, plus its transpose."""
    num_nodes: int
    fwd: CSR                       # rows = dst, col = src
    bwd: CSR                       # rows = src, col = dst (row to gather the output gradient from)
    t2f: torch.Tensor              # int64 [nnz]: transposed position -> forward CSR position
    gcn_w: Optional[torch.Tensor] = None      # float32 [nnz] in forward CSR order
    gcn_w_t: Optional[torch.Tensor] = None    # same weights in transposed order

    @staticmethod
    def build(edge_index: torch.Tensor, num_nodes: int, hub_thresh=None, quantum=None) -> "LoopGraph":
        lib = _lib.load()
        _require_cuda(edge_index, "edge_index")
        ei = edge_index.long()
        keep = ei[0] != ei[1]
        loops = torch.arange(num_nodes, dtype=torch.int64, device=ei.device)
        src = torch.cat([ei[0][keep], loops]).contiguous()         # remove_self_loops + add_self_loops
        dst = torch.cat([ei[1][keep], loops]).contiguous()
        fwd, _ = build_csr(dst, src, num_nodes, num_nodes, hub_thresh=hub_thresh, quantum=quantum)
        bwd = transpose_csr(src, dst.to(torch.int32), num_nodes, hub_thresh=hub_thresh, quantum=quantum)
        inv = torch.empty_like(fwd.perm, dtype=torch.int64)
        inv[fwd.perm.long()] = torch.arange(fwd.perm.numel(), dtype=torch.int64, device=ei.device)
        t2f = inv[bwd.perm.long()].contiguous()
        g = LoopGraph(num_nodes=num_nodes, fwd=fwd, bwd=bwd, t2f=t2f)
        with torch.cuda.device(ei.device):
            dis = torch.empty(num_nodes, dtype=torch.float32, device=ei.device)
            w = torch.empty(fwd.nnz, dtype=torch.float32, device=ei.device)
            _lib.check(lib.gmlm_gcn_edge_weights(_ptr(fwd.rowptr), _ptr(fwd.col), num_nodes, _ptr(dis), _ptr(w),
                                                 _stream(ei.device)), "gcn_edge_weights")
        g.gcn_w = w
        g.gcn_w_t = w[t2f].contiguous()
        return g


_LOOP_CACHE: "OrderedDict[tuple, tuple]" = OrderedDict()


def get_loop_graph(edge_index: torch.Tensor, num_nodes: int) -> LoopGraph:
    key = (_tensor_key(edge_index), int(num_nodes))
    hit = _LOOP_CACHE.get(key)
    if hit is not None:
        _LOOP_CACHE.move_to_end(key)
        return hit[0]
    g = LoopGraph.build(edge_index, num_nodes)
    _LOOP_CACHE[key] = (g, edge_index)
    while len(_LOOP_CACHE) > 4:
        _LOOP_CACHE.popitem(last=False)
    return g


def _with_w(csr: CSR, w: torch.Tensor) -> CSR:
    """A shallow view of ``csr`` carrying different per-edge weights."""
    c = CSR(**{k: getattr(csr, k) for k in csr.__dataclass_fields__})
    c.w = w
    return c


# ------------------------------------------------------------------------------ GCN (A8)
class _GCNAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, graph: LoopGraph):
        ctx.graph = graph
        return spmm(z, _with_w(graph.fwd, graph.gcn_w), _lib.AGG_WEIGHTED)

    @staticmethod
    def backward(ctx, g):
        graph: LoopGraph = ctx.graph
        return spmm(g.contiguous(), _with_w(graph.bwd, graph.gcn_w_t), _lib.AGG_WEIGHTED), None


class GCNConv(nn.Module):
    """Upstream ``GCNConv`` defaults: ``out = D^-1/2 (A + I) D^-1/2 (x W) + b`` (row A8)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.lin.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, edge_index) -> torch.Tensor:
        graph = edge_index if isinstance(edge_index, LoopGraph) else get_loop_graph(edge_index, x.size(0))
        out = _GCNAggregate.apply(self.lin(x), graph)
        return out + self.bias if self.bias is not None else out


# ------------------------------------------------------------------------------ GAT (A9)
class _GATAggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, a_src, a_dst, graph: LoopGraph, slope: float):
        lib = _lib.load()
        z = _rowmajor(z)
        n, heads = a_src.shape
        dev = z.device
        a_src32, a_dst32 = a_src.detach().float().contiguous(), a_dst.detach().float().contiguous()
        with torch.cuda.device(dev):
            alpha = torch.empty((graph.fwd.nnz, heads), dtype=torch.float32, device=dev)
            _lib.check(lib.gmlm_gat_alpha_fwd(_ptr(graph.fwd.rowptr), _ptr(graph.fwd.col), graph.num_nodes,
                                              _ptr(a_src32), _ptr(a_dst32), heads, float(slope), _ptr(alpha),
                                              _stream(dev)), "gat_alpha_fwd")
        out = spmm(z, _with_w(graph.fwd, alpha), _lib.AGG_WEIGHTED)
        ctx.graph, ctx.slope = graph, float(slope)
        ctx.save_for_backward(z, a_src32, a_dst32, alpha)
        ctx.in_dtypes = (a_src.dtype, a_dst.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        z, a_src, a_dst, alpha = ctx.saved_tensors
        graph: LoopGraph = ctx.graph
        g = _rowmajor(g)
        if g.dtype != z.dtype:
            g = g.to(z.dtype)
        n, heads = a_src.shape
        head_dim = z.size(1) // heads
        dev = z.device
        with torch.cuda.device(dev):
            d_score = torch.empty_like(alpha)
            da_dst = torch.empty((n, heads), dtype=torch.float32, device=dev)
            _lib.check(lib.gmlm_gat_alpha_bwd(_ptr(graph.fwd.rowptr), _ptr(graph.fwd.col), n, _ptr(z), _ld(z), _ptr(g),
                                              _ld(g), _dtype_code(z, "gat_alpha_bwd"), heads, head_dim, _ptr(a_src),
                                              _ptr(a_dst), _ptr(alpha), ctx.slope, _ptr(d_score), _ptr(da_dst),
                                              _stream(dev)), "gat_alpha_bwd")
            da_src = torch.empty((n, heads), dtype=torch.float32, device=dev)
            _lib.check(lib.gmlm_segment_sum_f32(_ptr(d_score), _ptr(graph.t2f), _ptr(graph.bwd.rowptr), n, heads,
                                                _ptr(da_src), _stream(dev)), "segment_sum")
        dz = spmm(g, _with_w(graph.bwd, alpha[graph.t2f].contiguous()), _lib.AGG_WEIGHTED)
        return dz, da_src.to(ctx.in_dtypes[0]), da_dst.to(ctx.in_dtypes[1]), None, None


def gat_aggregate(z, a_src, a_dst, graph: LoopGraph, negative_slope: float = 0.2) -> torch.Tensor:
    """``out[i,h,:] = sum_j softmax_j(leaky_relu(a_src[j,h] + a_dst[i,h])) * z[j,h,:]`` over the in-edges."""
    _require_cuda(z, "z")
    return _GATAggregate.apply(z, a_src, a_dst, graph, negative_slope)


class GATConv(nn.Module):
    """Upstream ``GATConv`` defaults (self-loops, LeakyReLU 0.2, concat heads, bias; attention
    dropout is not implemented — the upstream default is 0).  Row A9 / BASELINE configs[2]."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, bias: bool = True):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError("gmlm_b200.GATConv: attention dropout is not implemented (upstream default 0)")
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negative_slope = negative_slope
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.lin.weight)
        glorot_(self.att_src)
        glorot_(self.att_dst)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, edge_index) -> torch.Tensor:
        n, h, c = x.size(0), self.heads, self.out_channels
        graph = edge_index if isinstance(edge_index, LoopGraph) else get_loop_graph(edge_index, n)
        z = self.lin(x)
        zv = z.view(n, h, c)
        a_src = (zv * self.att_src).sum(-1)
        a_dst = (zv * self.att_dst).sum(-1)
        out = gat_aggregate(z, a_src, a_dst, graph, self.negative_slope)
        if not self.concat:
            out = out.view(n, h, c).mean(dim=1)
        return out + self.bias if self.bias is not None else out
