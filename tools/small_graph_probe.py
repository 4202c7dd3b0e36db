#!/usr/bin/env python
"""Development probe: forward/backward aggregation time on the small reference-shaped graphs for a
few group quanta; prints per-iteration times to expose outliers."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import gmlm_b200 as G
from gmlm_b200 import _lib, synth

dev = torch.device("cuda:0")
for key in ("c2", "c3"):
    w = synth.WORKLOADS[key]
    ei = synth.make_graph(w, device=dev)
    x = synth.make_features(w.num_nodes, w.feat, device=dev)
    et = G.edge_type_from_degree(ei, w.num_nodes)
    for q in (512, 128, 32, 0):
        g = G.RelGraph.build(ei, et, w.num_nodes, 5, quantum=q)
        gh = torch.randn(w.num_nodes * g.num_slots, w.feat, device=dev)
        for name, fn in (("fwd", lambda: G.spmm(x, g.fwd, _lib.AGG_MEAN)), ("bwd", lambda: G.spmm(gh, g.bwd, _lib.AGG_WEIGHTED))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(round(a.elapsed_time(b) * 1e3))
            print(key, "quantum", q, "slots", g.num_slots, "groups", g.fwd.n_groups, name, "us:", ts, flush=True)
