#!/usr/bin/env python
"""Development sweep (GPU): time the aggregation kernel pair on a workload for each
(variant, hub_thresh) and print achieved algorithmic GB/s.  Not part of the product or bench."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import gmlm_b200 as G  # noqa: E402
from gmlm_b200 import _lib, synth  # noqa: E402
from bench import algorithmic_bytes  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--thresh", type=int, nargs="*", default=[256, 1024, 4096])
    ap.add_argument("--variants", type=int, nargs="*", default=[1])
    ap.add_argument("--quantum", type=int, nargs="*", default=[0, 128, 256, 512])
    ap.add_argument("--unroll", type=int, nargs="*", default=[8, 4])
    ap.add_argument("--uniform", action="store_true", help="uniform random edges instead of the workload's generator")
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    w = synth.WORKLOADS[args.workload]
    n, e, feat = int(w.num_nodes * args.scale), int(w.num_edges * args.scale), w.feat
    dtype = torch.bfloat16 if w.dtype == "bf16" else torch.float32
    esz = 2 if dtype == torch.bfloat16 else 4
    ei = synth.uniform_edges(n, e, device=dev) if args.uniform else synth.make_graph(w, device=dev, num_nodes=n, num_edges=e)
    x = synth.make_features(n, feat, device=dev, dtype=dtype)
    et = G.edge_type_from_degree(ei, n)
    Path(args.out).parent.mkdir(exist_ok=True)
    rows = []
    for th, qn in [(t, q) for t in args.thresh for q in args.quantum]:
        g = G.RelGraph.build(ei, et, n, 5, hub_thresh=th, quantum=qn)
        S = g.num_slots
        gh = synth.make_features(n * S, feat, device=dev, seed=7, dtype=dtype)
        fb, bb = algorithmic_bytes(n, e, feat, esz, S)
        lens = (g.fwd.rowptr[1:] - g.fwd.rowptr[:-1])
        lens_t = (g.bwd.rowptr[1:] - g.bwd.rowptr[:-1])
        for v, un in [(v, u) for v in args.variants for u in args.unroll]:
            G.set_tuning("spmm_variant", v)
            G.set_tuning("spmm_unroll", un)
            f_ms = timeit(lambda: G.spmm(x, g.fwd, _lib.AGG_MEAN))
            b_ms = timeit(lambda: G.spmm(gh, g.bwd, _lib.AGG_WEIGHTED))
            row = {"thresh": th, "quantum": qn, "variant": v, "unroll": un, "fwd_ms": f_ms, "bwd_ms": b_ms, "fwd_gbs": fb / f_ms / 1e6,
                   "bwd_gbs": bb / b_ms / 1e6, "edges_per_s": e / ((f_ms + b_ms) * 1e-3),
                   "hub_fwd": g.fwd.n_hub, "chunks_fwd": g.fwd.n_chunks, "hub_bwd": g.bwd.n_hub,
                   "chunks_bwd": g.bwd.n_chunks, "max_len_fwd": int(lens.max()), "max_len_bwd": int(lens_t.max()),
                   "empty_frac_fwd": float((lens == 0).float().mean()), "slots": S}
            rows.append(row)
            print(json.dumps(row), flush=True)
        del g, gh
    with open(args.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
