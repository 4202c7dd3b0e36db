"""The GNN encoder body (``get_graph_embeddings``, ``/root/reference/main.py:250-320``) for ONE rank of a
destination-row partition.  The reference is single-device; this is the composition the
multi-GPU pieces of this package are for:

    per layer:  X = halo exchange of the layer input (rows of remote sources appended)      partition.py
                y = RGCNConv on the rank's rectangular (dst,rel) CSR, root term on the local rows   nn.py
                y = GELU(GraphNorm(y)) with the column sums all-reduced over the ranks       dist_norm.py
    residual projections and MultiScaleFusion (+ LayerNorm) are row-wise: local, unchanged.

Parameters are replicated (the wrapped encoder keeps the reference's names: ``rgcn1..4``, ``gnorm1..4``,
``residual_proj1..3``, ``multi_scale_fusion``); with a loss that is a sum over nodes, every rank's backward
yields its PARTIAL gradient of the replicated parameters, so ``sync_gradients`` all-reduces them — except the
GraphNorm parameters, whose gradients already come out of ``partitioned_graph_norm`` as whole-graph values.

``ops`` abstracts the three device operations so that the rank logic can be run on CPU with gloo against the
oracle encoder (tests/test_partition.py supplies oracle-backed ops); the product ops are the CUDA paths and
refuse CPU tensors.
"""
from __future__ import annotations


import torch
import torch.distributed as dist
import torch.nn as nn


class CudaPartitionOps:
    """Product operations: NCCL halo exchange with the pack/unpack kernels, the CUDA RGCNConv on the rank's
    rectangular RelGraph, GraphNorm+GELU with all-reduced column sums."""

    def __init__(self, part, graph, num_nodes_global: int, group=None):
        self.part, self.graph, self.n_global, self.group = part, graph, int(num_nodes_global), group

    def exchange(self, x_local):
        from .partition import halo_exchange
        return halo_exchange(x_local, self.part, self.group)

    def conv(self, conv_module, X):
        return conv_module(X, self.graph)

    def norm_gelu(self, gn, y):
        from .dist_norm import partitioned_graph_norm
        return partitioned_graph_norm(y, gn.weight, gn.bias, gn.mean_scale, self.n_global, gn.eps, True, self.group)


class PartitionedGraphEncoder(nn.Module):
    def __init__(self, encoder: nn.Module, ops):
        super().__init__()
        self.encoder = encoder
        self.ops = ops

    def _block(self, k: int, x_local: torch.Tensor) -> torch.Tensor:
        enc = self.encoder
        conv, gn, drop = getattr(enc, f"rgcn{k}"), getattr(enc, f"gnorm{k}"), getattr(enc, f"dropout{k}")
        y = self.ops.conv(conv, self.ops.exchange(x_local))
        return drop(self.ops.norm_gelu(gn, y))            # the whole graph has > 1 node here (main.py:273)

    def _residual(self, lin: nn.Linear, x: torch.Tensor, acc: torch.Tensor) -> torch.Tensor:
        enc = self.encoder
        if acc.is_cuda and hasattr(enc, "_residual"):
            return enc._residual(lin, x, acc)              # the add folded into the projection's epilogue
        if x.dtype == lin.weight.dtype or torch.is_autocast_enabled(x.device.type):
            return acc + lin(x).to(acc.dtype)
        return acc + torch.nn.functional.linear(x, lin.weight.to(x.dtype), lin.bias.to(x.dtype)).to(acc.dtype)

    def forward(self, x_local: torch.Tensor, return_layers: bool = False):
        enc = self.encoder
        outs = []
        x1 = self._block(1, x_local)
        outs.append(x1)                                   # pre-residual outputs feed the fusion (main.py:279)
        x1 = self._residual(enc.residual_proj1, x_local, x1)
        x2 = self._block(2, x1)
        outs.append(x2)
        x2 = self._residual(enc.residual_proj2, x1, x2)
        x3 = self._block(3, x2)
        outs.append(x3)
        x4 = self._block(4, x3)
        outs.append(x4)                                   # the dead x4 + residual_proj3(x2) is not computed
        fused = enc.multi_scale_fusion(outs)
        return (fused, outs) if return_layers else fused


def sync_gradients(encoder: nn.Module, group=None) -> None:
    """All-reduce (SUM) the gradients of the replicated parameters after a backward over a node-sum loss.
    GraphNorm parameters are skipped: their gradients are already whole-graph values on every rank.
    Which parameters take part is agreed collectively first (a parameter that received a gradient on ANY rank is
    reduced on every rank, a missing local gradient counting as zeros), so that ranks whose backward skipped a
    parameter cannot leave the others waiting in a collective."""
    params = [(n, p) for n, p in encoder.named_parameters()
              if p.requires_grad and not n.split(".")[0].startswith("gnorm")]
    if not params:
        return
    has = torch.tensor([0 if p.grad is None else 1 for _, p in params], dtype=torch.int32, device=params[0][1].device)
    dist.all_reduce(has, op=dist.ReduceOp.MAX, group=group)
    for flag, (_, p) in zip(has.tolist(), params):
        if not flag:
            continue
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
