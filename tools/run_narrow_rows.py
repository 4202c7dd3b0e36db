#!/usr/bin/env python
"""Development (GPU): the aggregation pair on the C4 graph at the encoder's row widths (F = 64, 128, 256 bf16 ->
8-, 16-, 32-lane groups) with event timing; run under ncu (-k regex:rows_kernel|chunk_kernel) for the LPR table."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import gmlm_b200 as G
from gmlm_b200 import _lib, synth
from bench import algorithmic_bytes

dev = torch.device("cuda:0")
w = synth.WORKLOADS["c4"]
n, e = w.num_nodes, w.num_edges
ei = synth.make_graph(w, device=dev)
et = G.edge_type_from_degree(ei, n)
g = G.get_rel_graph(ei, et, n, 5)
S = g.num_slots
a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
import os
for feat in [int(v) for v in os.environ.get("NARROW_FEATS", "64,128,256").split(",")]:
    x = synth.make_features(n, feat, device=dev, dtype=torch.bfloat16)
    gh = synth.make_features(n * S, feat, device=dev, seed=7, dtype=torch.bfloat16)
    for _ in range(2):
        G.spmm(x, g.fwd, _lib.AGG_MEAN)
        G.spmm(gh, g.bwd, _lib.AGG_WEIGHTED)
    torch.cuda.synchronize()
    f = bw = 0.0
    for _ in range(5):
        a.record(); G.spmm(x, g.fwd, _lib.AGG_MEAN); b.record(); G.spmm(gh, g.bwd, _lib.AGG_WEIGHTED); c.record()
        torch.cuda.synchronize()
        f += a.elapsed_time(b) / 5; bw += b.elapsed_time(c) / 5
    fb, bb = algorithmic_bytes(n, e, feat, 2, S)
    print(f"F={feat}: fwd {f:.3f} ms ({fb / f / 1e6:.0f} GB/s algorithmic)  bwd {bw:.3f} ms ({bb / bw / 1e6:.0f} GB/s)", flush=True)
    del x, gh
print("done")
