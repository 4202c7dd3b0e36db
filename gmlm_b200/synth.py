"""Seeded synthetic graphs of the five shapes BASELINE.json names.

The reference's datasets (``collapse/data/*.npz``, ``/root/reference/main.py:841-845``) are not
shipped, so every measurement and parity test runs on these.  Edges are directed; duplicates
and self-loops are kept (the reference never cleans them).  Works on CPU or CUDA generators
(the two produce different streams: generate once, then copy, when both sides need the
same graph).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch


@dataclass(frozen=True)
class Workload:
    key: str
    title: str
    num_nodes: int
    num_edges: int
    feat: int
    num_classes: int
    generator: str          # "uniform" | "rmat"
    dtype: str = "f32"
    hidden: int = 512       # hidden_channels (main.py:1003); C4/C5 use 64


WORKLOADS = {
    "c1": Workload("c1", "Cornell-shaped (183 nodes, 300 edges, 1703 feats)", 183, 300, 1703, 5, "uniform", "f32", 512),
    "c2": Workload("c2", "Roman-empire-shaped (22.7k nodes, 32.9k edges, 300 feats)", 22662, 32927, 300, 18,
                   "uniform", "f32", 512),
    "c3": Workload("c3", "Amazon-ratings-shaped (24.5k nodes, 93k edges, 300 feats)", 24492, 93050, 300, 5,
                   "uniform", "f32", 512),
    "c4": Workload("c4", "power-law R-MAT (2M nodes, 40M edges, 256 feats, bf16)", 2_000_000, 40_000_000, 256, 16,
                   "rmat", "bf16", 64),
    "c5": Workload("c5", "power-law R-MAT (10M nodes, 200M edges, 256 feats, bf16)", 10_000_000, 200_000_000, 256,
                   16, "rmat", "bf16", 64),
}

RMAT_ABCD = (0.57, 0.19, 0.19, 0.05)


def _gen(device, seed: int) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def uniform_edges(num_nodes: int, num_edges: int, device="cpu", seed: int = 42) -> torch.Tensor:
    g = _gen(device, seed)
    return torch.randint(0, num_nodes, (2, num_edges), generator=g, device=device, dtype=torch.int64)


def rmat_edges(num_nodes: int, num_edges: int, device="cpu", seed: int = 42, abcd=RMAT_ABCD,
               chunk: int = 1 << 24) -> torch.Tensor:
    """R-MAT (a,b,c,d) with scale = ceil(log2 N); ids are folded into [0, N) by modulo.
    Natural id order is kept (low ids are the hubs), so destination-row ranges keep R-MAT's
    block locality."""
    scale = max(1, (num_nodes - 1).bit_length())
    a, b, c, _ = abcd
    g = _gen(device, seed)
    out = torch.empty((2, num_edges), dtype=torch.int64, device=device)
    for lo in range(0, num_edges, chunk):
        n = min(chunk, num_edges - lo)
        src = torch.zeros(n, dtype=torch.int64, device=device)
        dst = torch.zeros(n, dtype=torch.int64, device=device)
        for _ in range(scale):
            u = torch.rand(n, generator=g, device=device)
            src_bit = (u >= a + b).to(torch.int64)
            dst_bit = ((u >= a) & (u < a + b) | (u >= a + b + c)).to(torch.int64)
            src = (src << 1) | src_bit
            dst = (dst << 1) | dst_bit
        out[0, lo:lo + n] = src % num_nodes
        out[1, lo:lo + n] = dst % num_nodes
    return out


def make_graph(w: Workload, device="cpu", seed: int = 42, num_nodes: Optional[int] = None,
               num_edges: Optional[int] = None) -> torch.Tensor:
    n = num_nodes or w.num_nodes
    e = num_edges or w.num_edges
    if w.generator == "rmat":
        return rmat_edges(n, e, device=device, seed=seed)
    return uniform_edges(n, e, device=device, seed=seed)


def make_features(num_nodes: int, feat: int, device="cpu", seed: int = 42, dtype=torch.float32,
                  chunk_rows: int = 1 << 20) -> torch.Tensor:
    """x ~ N(0,1), generated in fp32 row chunks then cast (bounds temp memory at 10M x 256)."""
    g = _gen(device, seed + 1)
    x = torch.empty((num_nodes, feat), dtype=dtype, device=device)
    for lo in range(0, num_nodes, chunk_rows):
        n = min(chunk_rows, num_nodes - lo)
        x[lo:lo + n] = torch.randn((n, feat), generator=g, device=device, dtype=torch.float32).to(dtype)
    return x
