"""The step in front of the hot path in both training loops: the degree-weighted active-node mask of
``generate_active_node_mask`` (``/root/reference/main.py:47-89``; called twice per pre-training epoch,
``main.py:439-440``, and once per fine-tuning epoch, ``main.py:532``) — SURVEY §8f N2.

Same signature, same branches, same random stream as the reference: ``torch.multinomial(probs, k,
replacement=False)`` IS "k largest of probs / q with q ~ Exp(1)" (ATen's implementation), so drawing
``q`` from the same generator and taking ``topk`` returns the very same nodes for the same seed on the
same device — checked bit for bit against the reference's own function (tests/test_dropin_reference.py).
What changes:

  * the out-degree comes from the cached ``gmlm_degree`` kernel result (``deg=`` lets a caller pass it in;
    the reference recounts it with a scatter on every call);
  * ``multinomial``'s three validity checks (each a device->host sync) and the reference's three
    ``.sum() == 0`` tests collapse into ONE host read of a 2-element tensor;
  * no ``2^24`` category limit (``multinomial`` refuses more; the 10M-node graph has 10^7 base nodes).

Host logic on stock torch ops (``exponential_``, ``topk``): there is no arithmetic here worth a kernel.
"""
from __future__ import annotations

from typing import Optional

import torch


def weighted_sample_without_replacement(weights: torch.Tensor, k: int,
                                        generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Indices of ``k`` items drawn without replacement with probability proportional to ``weights`` —
    the exponential-race form ATen's ``multinomial`` uses (bit-identical to it for the same generator
    state), without its host-synchronising validity checks.  ``weights`` must be non-negative with at
    least ``k`` positive entries (the reference's call would raise otherwise; here zero-weight items are
    simply ranked last)."""
    q = torch.empty_like(weights).exponential_(1, generator=generator)
    vals = weights / q
    if k == 1:
        return vals.argmax(dim=-1, keepdim=True)
    return vals.topk(k).indices


def generate_active_node_mask(data, mask_ratio, split_edge=None, base_mask_name="train_mask", *,
                              deg: Optional[torch.Tensor] = None,
                              generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Drop-in for ``generate_active_node_mask`` (``/root/reference/main.py:47-89``)."""
    device = data.x.device
    num_nodes = data.num_nodes
    if base_mask_name and hasattr(data, base_mask_name) and getattr(data, base_mask_name) is not None:
        base_nodes_idx = getattr(data, base_mask_name).nonzero(as_tuple=False).reshape(-1)
        if base_nodes_idx.numel() == 0:
            return torch.zeros(num_nodes, dtype=torch.bool, device=device)
    elif split_edge is not None and "train" in split_edge and "edge" in split_edge["train"]:
        base_nodes_idx = torch.unique(split_edge["train"]["edge"].flatten())
    else:
        base_nodes_idx = torch.arange(num_nodes, device=device)

    num_base_nodes = base_nodes_idx.size(0)
    if num_base_nodes == 0:
        return torch.zeros(num_nodes, dtype=torch.bool, device=device)
    num_select = max(1, min(int(mask_ratio * num_base_nodes), num_base_nodes))

    if deg is None:
        from .ops import degree                      # CUDA kernel; CPU callers pass `deg`
        deg = degree(data.edge_index[0].to(torch.long), num_nodes=num_nodes)
    degrees_of_base_nodes = deg[base_nodes_idx]

    # main.py:67-83 in one host read: [sum of degrees, sum of probabilities]
    total = degrees_of_base_nodes.sum()
    probs = degrees_of_base_nodes.float() / total.float()
    probs = torch.nan_to_num(probs, nan=1.0 / num_base_nodes)
    total_h, psum_h = torch.stack([total.float(), probs.sum()]).tolist()
    if total_h == 0 or psum_h == 0:
        # main.py:68-70 / 77-79: uniform choice (same randperm call, same stream)
        sampled = base_nodes_idx[torch.randperm(num_base_nodes, device=device, generator=generator)[:num_select]]
    else:
        sampled = base_nodes_idx[weighted_sample_without_replacement(probs, num_select, generator)]

    output_mask = torch.zeros(num_nodes, dtype=torch.bool, device=device)
    output_mask[sampled] = True
    return output_mask
