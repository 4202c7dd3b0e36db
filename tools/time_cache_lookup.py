#!/usr/bin/env python
"""Development (GPU): what a graph-cache lookup costs at C4 size (40M edges) -- identity hit, content hit (a fresh
edge_type tensor per call, the import-swap pattern of main.py:255) -- against the CSR build it avoids."""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import gmlm_b200 as G
from gmlm_b200 import graph as gg

dev = torch.device("cuda:0")
n, e = 2_000_000, 40_000_000
g = torch.Generator(device=dev).manual_seed(0)
ei = torch.randint(0, n, (2, e), device=dev, generator=g)
et = G.edge_type_from_degree(ei, n)


def timed(fn, reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


build_ms = timed(lambda: G.get_rel_graph(ei, et, n, 5), 1)
ident_ms = timed(lambda: G.get_rel_graph(ei, et, n, 5), 20)
fresh = [et.clone() for _ in range(5)]
it = iter(fresh)
content_ms = timed(lambda: G.get_rel_graph(ei, next(it), n, 5), 5)
print(f"C4-size graph ({n} nodes, {e} edges): build {build_ms:.1f} ms, identity hit {ident_ms * 1e3:.1f} us, "
      f"content hit (fresh edge_type tensor) {content_ms:.2f} ms; stats {gg.cache_stats}")
