"""Pure-PyTorch restatement of the torch_geometric operators the reference imports
(``/root/reference/main.py:6-7``).  TEST INFRASTRUCTURE — see oracle/__init__.py.

Everything here runs on CPU tensors of any float dtype (fp64 for the parity
checks), uses only ``index_select`` / ``index_add_`` / ``matmul`` and follows the
*per-relation loop* that upstream RGCNConv executes, so that the CUDA path's
different formulation (one (dst,rel)-keyed CSR, one concatenated GEMM, gather-based
backward) is checked against an independent statement of the same maths.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- A1
def degree_ref(index: torch.Tensor, num_nodes: int, dtype=None) -> torch.Tensor:
    """``torch_geometric.utils.degree`` as called at main.py:65 and main.py:256.

    Upstream: ``zeros(N, dtype).scatter_add_(0, index, ones(E))`` with the default
    dtype (float32).  Counts are integer-valued and exact below 2**24.
    """
    out = torch.zeros((num_nodes,), dtype=dtype or torch.get_default_dtype(), device=index.device)
    one = torch.ones((index.numel(),), dtype=out.dtype, device=out.device)
    return out.scatter_add_(0, index.reshape(-1).long(), one)


# --------------------------------------------------------------------------- A2
def edge_type_loop_ref(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """The reference's per-edge Python loop, main.py:253-267, restated literally.

    O(E) interpreter work — call only on small graphs.
    """
    edge_index = edge_index.long()
    num_edges = edge_index.size(1)
    edge_type = torch.zeros(num_edges, dtype=torch.long, device=edge_index.device)
    deg = degree_ref(edge_index[0], num_nodes)
    for i in range(num_edges):
        d = deg[edge_index[0, i]]
        if d <= 2:
            edge_type[i] = 0
        elif d <= 5:
            edge_type[i] = 1
        elif d <= 10:
            edge_type[i] = 2
        else:
            edge_type[i] = 3
    return edge_type


def edge_type_bucket_ref(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """Vectorised equivalent of main.py:253-267."""
    edge_index = edge_index.long()
    deg = degree_ref(edge_index[0], num_nodes)
    bounds = torch.tensor([2.0, 5.0, 10.0], dtype=deg.dtype, device=deg.device)
    return torch.bucketize(deg[edge_index[0]], bounds, right=False)


# ------------------------------------------------------------------ initialisers
def glorot_(t: torch.Tensor) -> torch.Tensor:
    """``torch_geometric.nn.inits.glorot``: U(-a, a), a = sqrt(6 / (size(-2)+size(-1)))."""
    if t is not None:
        a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        with torch.no_grad():
            t.uniform_(-a, a)
    return t


# --------------------------------------------------------------------------- A5
def rgcn_propagate_mean_ref(x: torch.Tensor, src: torch.Tensor, dst: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """One relation's ``propagate`` with aggr='mean' (upstream MessagePassing +
    ``scatter(..., reduce='mean')``): sum of x[src] into dst, divided by the
    in-count clamped to >= 1.  Duplicate edges count twice, self-loops are kept."""
    x_j = x.index_select(0, src)
    summed = torch.zeros((num_nodes, x.size(1)), dtype=x.dtype, device=x.device).index_add_(0, dst, x_j)
    cnt = torch.zeros((num_nodes,), dtype=x.dtype, device=x.device).index_add_(
        0, dst, torch.ones(dst.numel(), dtype=x.dtype, device=x.device))
    return summed / cnt.clamp(min=1).unsqueeze(-1)


# ----------------------------------------------------------------------- A3-A6
class RGCNConvRef(nn.Module):
    """Upstream ``RGCNConv`` (basis-decomposition branch, aggr='mean', root_weight,
    bias) as constructed at main.py:189 and called at main.py:272.

    State-dict ABI: ``weight [num_bases,in,out]``, ``comp
    [num_relations,num_bases]``, ``root [in,out]``, ``bias [out]``; glorot, glorot,
    glorot, zeros.
    """

    def __init__(self, in_channels: int, out_channels: int, num_relations: int, num_bases: int | None = None):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_relations = num_relations
        self.num_bases = num_bases
        if num_bases is not None:
            self.weight = nn.Parameter(torch.empty(num_bases, in_channels, out_channels))
            self.comp = nn.Parameter(torch.empty(num_relations, num_bases))
        else:
            self.weight = nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
            self.register_parameter("comp", None)
        self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        glorot_(self.weight)
        glorot_(self.comp)
        glorot_(self.root)
        nn.init.zeros_(self.bias)

    def composed_weight(self) -> torch.Tensor:
        w = self.weight
        if self.num_bases is not None:  # A4
            w = (self.comp @ w.view(self.num_bases, -1)).view(
                self.num_relations, self.in_channels, self.out_channels)
        return w

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_type: torch.Tensor) -> torch.Tensor:
        n = x.size(0)
        # upstream: out = torch.zeros(N, out_channels, device=...)  (default dtype);
        # the oracle keeps the parameter dtype so it can run in fp64.
        out = torch.zeros((n, self.out_channels), dtype=self.root.dtype, device=x.device)
        w = self.composed_weight()
        for r in range(self.num_relations):  # A3: boolean-mask compaction per relation
            m = edge_type == r
            src, dst = edge_index[0, m], edge_index[1, m]
            h = rgcn_propagate_mean_ref(x, src, dst, n)          # A5
            out = out + h @ w[r]                                 # A6
        out = out + x @ self.root
        out = out + self.bias
        return out


# --------------------------------------------------------------------------- A7
class GraphNormRef(nn.Module):
    """Upstream ``GraphNorm`` with ``batch=None`` (whole graph = one segment), as
    built at main.py:190 and called at main.py:273."""

    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.in_channels = in_channels
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))
        self.mean_scale = nn.Parameter(torch.ones(in_channels))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        mean = x.mean(dim=0, keepdim=True)
        out = x - mean * self.mean_scale
        var = out.pow(2).mean(dim=0, keepdim=True)
        std = (var + self.eps).sqrt()
        return self.weight * out / std + self.bias


# -------------------------------------------------------------------------- A11
def soft_masking_ref(x: torch.Tensor, mask: torch.Tensor, mask_token_embed: torch.Tensor, beta: float = 0.7) -> torch.Tensor:
    """main.py:92-99."""
    x_masked = x.clone()
    if mask.any():
        x_masked[mask] = (1 - beta) * x[mask] + beta * mask_token_embed
    return x_masked


# --------------------------------------------------------------------------- A8
def _remove_then_add_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """``remove_self_loops`` followed by ``add_self_loops`` (GATConv) — and, for
    unit weights, also what ``add_remaining_self_loops`` (GCNConv) amounts to:
    non-loop edges in their original order, then one loop per node 0..N-1."""
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)


class GCNConvRef(nn.Module):
    """Upstream ``GCNConv`` defaults (add_self_loops, normalize, bias; no edge
    weights): extension A8 — no counterpart in /root/reference (parity unpinned)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        glorot_(self.lin.weight)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        n = x.size(0)
        ei = _remove_then_add_self_loops(edge_index.long(), n)
        row, col = ei[0], ei[1]
        w = torch.ones(ei.size(1), dtype=x.dtype, device=x.device)
        deg = torch.zeros(n, dtype=x.dtype, device=x.device).index_add_(0, col, w)
        dis = deg.pow(-0.5)
        dis = dis.masked_fill(dis == float("inf"), 0)
        norm = dis[row] * w * dis[col]
        z = self.lin(x)
        out = torch.zeros_like(z).index_add_(0, col, norm.unsqueeze(-1) * z.index_select(0, row))
        return out + self.bias


# --------------------------------------------------------------------------- A9
class GATConvRef(nn.Module):
    """Upstream ``GATConv`` defaults (concat heads, negative_slope 0.2, self-loops,
    bias, attention dropout p): extension A9 — no counterpart in /root/reference
    (parity unpinned)."""

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0):
        super().__init__()
        self.heads, self.out_channels, self.concat = heads, out_channels, concat
        self.negative_slope, self.dropout = negative_slope, dropout
        self.lin = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))
        glorot_(self.lin.weight)
        glorot_(self.att_src)
        glorot_(self.att_dst)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        n, hd, c = x.size(0), self.heads, self.out_channels
        z = self.lin(x).view(n, hd, c)
        a_src = (z * self.att_src).sum(-1)
        a_dst = (z * self.att_dst).sum(-1)
        ei = _remove_then_add_self_loops(edge_index.long(), n)
        src, dst = ei[0], ei[1]
        e = F.leaky_relu(a_src[src] + a_dst[dst], self.negative_slope)          # [E', hd]
        # torch_geometric.utils.softmax: max-shift, exp, sum + 1e-16
        emax = torch.full((n, hd), float("-inf"), dtype=e.dtype, device=e.device)
        emax = emax.scatter_reduce(0, dst.unsqueeze(-1).expand_as(e), e.detach(), reduce="amax", include_self=True)
        ex = (e - emax[dst]).exp()
        den = torch.zeros((n, hd), dtype=e.dtype, device=e.device).index_add_(0, dst, ex) + 1e-16
        alpha = ex / den[dst]
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = torch.zeros((n, hd, c), dtype=z.dtype, device=z.device).index_add_(
            0, dst, alpha.unsqueeze(-1) * z.index_select(0, src))
        out = out.reshape(n, hd * c) if self.concat else out.mean(dim=1)
        return out + self.bias
