"""numpy restatement of the (dst,rel)-keyed CSR the CUDA path builds once per graph
in place of the five boolean-mask compactions upstream RGCNConv does per call
(call sites main.py:272,285,298,308).  TEST INFRASTRUCTURE.

Contract the CUDA ``csr_build`` must match bit-exactly:
  * forward CSR: segments ``s = dst*S + slot`` (S = number of relation slots),
    ``rowptr int32[N*S+1]``, ``col int32[E]`` = source node of each edge,
    ``perm int32[E]`` = original edge position; edges inside a segment keep their
    original order (stable sort) so the fp summation order is deterministic.
  * transposed CSR (for the backward gather, row A14): rows = source node,
    ``rowptr_t int32[N+1]``, ``seg_t int32[E]`` = forward segment of each edge,
    ``w_t float32[E]`` = 1 / |segment| (the mean's divisor folded in),
    ``perm_t int32[E]``; stable in original edge order.
"""
from __future__ import annotations

import numpy as np


def rel_csr_ref(src: np.ndarray, dst: np.ndarray, slot: np.ndarray, num_nodes: int, num_slots: int):
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    slot = np.asarray(slot, dtype=np.int64)
    key = dst * num_slots + slot
    perm = np.argsort(key, kind="stable")
    counts = np.bincount(key, minlength=num_nodes * num_slots)
    rowptr = np.zeros(num_nodes * num_slots + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr.astype(np.int32), src[perm].astype(np.int32), perm.astype(np.int32)


def transposed_csr_ref(src: np.ndarray, dst: np.ndarray, slot: np.ndarray, num_nodes: int, num_slots: int):
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    slot = np.asarray(slot, dtype=np.int64)
    seg = dst * num_slots + slot
    seg_cnt = np.bincount(seg, minlength=num_nodes * num_slots)
    perm_t = np.argsort(src, kind="stable")
    counts = np.bincount(src, minlength=num_nodes)
    rowptr_t = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr_t[1:])
    seg_t = seg[perm_t]
    w_t = (np.float32(1.0) / seg_cnt[seg_t].astype(np.float32)).astype(np.float32)
    return rowptr_t.astype(np.int32), seg_t.astype(np.int32), w_t, perm_t.astype(np.int32)
