"""Restatement of the reference's GNN-encoder body, ``GraphTextLM.get_graph_embeddings``
(``/root/reference/main.py:250-320``) and ``MultiScaleFusion`` (``main.py:167-180``),
built from the oracle operators.  TEST INFRASTRUCTURE — see oracle/__init__.py.

Submodule / parameter names equal the reference's (``rgcn1..4``, ``gnorm1..4``,
``dropout1..4``, ``residual_proj1..3``, ``multi_scale_fusion.*``) so a state dict
can be moved between this oracle and the CUDA-backed encoder in tests.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

from .pyg_ref import GraphNormRef, RGCNConvRef, edge_type_bucket_ref


class MultiScaleFusionRef(nn.Module):
    """main.py:167-180: LayerNorm(sum_l softmax(s)_l * Linear_l(x_l))."""

    def __init__(self, hidden_dims, output_dim):
        super().__init__()
        self.scale_weights = nn.Parameter(torch.ones(len(hidden_dims)) / len(hidden_dims))
        self.projections = nn.ModuleList([nn.Linear(d, output_dim) for d in hidden_dims])
        self.layer_norm = nn.LayerNorm(output_dim)

    def forward(self, embeddings_list):
        w = F.softmax(self.scale_weights, dim=0)
        acc = 0
        for wi, proj, emb in zip(w, self.projections, embeddings_list):
            acc = acc + wi * proj(emb)
        return self.layer_norm(acc)


class EncoderRef(nn.Module):
    """The four conv→(GraphNorm if N>1)→GELU→Dropout blocks with the two live
    residual projections, the dead third one, and the multi-scale fusion."""

    def __init__(self, gnn_in_channels: int, hidden_channels: int, out_dim: int, num_relations: int = 5,
                 num_bases: int = 30, dropout_rate: float = 0.3, use_checkpoint: bool = True):
        super().__init__()
        h = hidden_channels
        dims = [gnn_in_channels, h, 2 * h, 4 * h, 8 * h]
        for k in range(4):                                   # main.py:189-203
            setattr(self, f"rgcn{k+1}", RGCNConvRef(dims[k], dims[k + 1], num_relations, num_bases))
            setattr(self, f"gnorm{k+1}", GraphNormRef(dims[k + 1]))
            setattr(self, f"dropout{k+1}", nn.Dropout(dropout_rate))
        self.residual_proj1 = nn.Linear(gnn_in_channels, h)  # main.py:205-207
        self.residual_proj2 = nn.Linear(h, 2 * h)
        self.residual_proj3 = nn.Linear(2 * h, 8 * h)
        self.multi_scale_fusion = MultiScaleFusionRef(dims[1:], out_dim)
        self.use_checkpoint = use_checkpoint

    def _block(self, k: int):
        conv, norm, drop = getattr(self, f"rgcn{k}"), getattr(self, f"gnorm{k}"), getattr(self, f"dropout{k}")

        def run(x, edge_index, edge_type):                   # main.py:271-276
            y = conv(x, edge_index, edge_type)
            if y.size(0) > 1:
                y = norm(y)
            return drop(F.gelu(y))
        return run

    def _run(self, k, x, edge_index, edge_type):
        if self.use_checkpoint:                              # main.py:278
            return checkpoint(self._block(k), x, edge_index, edge_type, use_reentrant=False)
        return self._block(k)(x, edge_index, edge_type)

    def forward(self, x_feat, edge_index, edge_type=None, return_layers: bool = False):
        edge_index = edge_index.long()
        if edge_type is None:                                # main.py:253-267
            edge_type = edge_type_bucket_ref(edge_index, x_feat.size(0))
        outs = []
        x1 = self._run(1, x_feat, edge_index, edge_type)
        outs.append(x1)                                      # appended BEFORE the residual (main.py:279,282)
        x1 = x1 + self.residual_proj1(x_feat)
        x2 = self._run(2, x1, edge_index, edge_type)
        outs.append(x2)
        x2 = x2 + self.residual_proj2(x1)
        x3 = self._run(3, x2, edge_index, edge_type)
        outs.append(x3)                                      # no residual on layer 3 (main.py:297-305)
        x4 = self._run(4, x3, edge_index, edge_type)
        outs.append(x4)
        _ = x4 + self.residual_proj3(x2)                     # computed then discarded (main.py:317-318)
        fused = self.multi_scale_fusion(outs)
        return (fused, outs) if return_layers else fused
