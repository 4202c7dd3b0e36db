#!/usr/bin/env python
"""Development (GPU): brute-force the parameter space of tests/test_gpu_spmm.py::test_aggregate_property_random_graphs
(boundary-heavy sampling) and print every failing example."""
import random
import sys
import traceback
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
import gmlm_b200 as G
from oracle import edge_type_bucket_ref
from test_gpu_spmm import oracle_aggregate, BF16_TOL, FP32_TOL
from conftest import rel_err

dev = torch.device("cuda:0")
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
bad = 0
for it in range(iters):
    n = rng.choice([1, 2, 3, 31, 32, 33, 255, 256, 257, 300, rng.randint(1, 300)])
    e = rng.choice([0, 1, 2, 16, 17, 255, 256, 257, 512, 3000, rng.randint(0, 3000)])
    feat = rng.choice([3, 8, 20, 64, 72, 256])
    seed = rng.randint(0, 10000)
    hub_thresh = rng.choice([2, 17, 256])
    quantum = rng.choice([0, 5, 64])
    bf16 = rng.random() < 0.5
    try:
        g_ = torch.Generator().manual_seed(seed)
        ei = torch.randint(0, n, (2, e), generator=g_)
        et = edge_type_bucket_ref(ei, n)
        dtype = torch.bfloat16 if bf16 else torch.float32
        tol = BF16_TOL if bf16 else max(FP32_TOL, 0.5 * max(e, 1) / max(n, 1) * 2.0 ** -24 * 4)
        x = torch.randn(n, feat, generator=g_).to(dtype)
        g = G.RelGraph.build(ei.to(dev), et.to(dev), n, 5, hub_thresh=hub_thresh, quantum=quantum)
        assert int(g.fwd.rowptr[-1]) == e and int(g.bwd.rowptr[-1]) == e
        xg = x.to(dev).requires_grad_(True)
        out = G.rgcn_aggregate(xg, g)
        gh = torch.randn(out.shape, generator=g_).to(dtype)
        out.backward(gh.to(dev))
        x64 = x.double().requires_grad_(True)
        ref = oracle_aggregate(x64, ei, et, n, g.live_rels)
        ref.backward(gh.double())
        if e == 0:
            assert torch.count_nonzero(out) == 0 and torch.count_nonzero(xg.grad) == 0
        else:
            e1, e2 = rel_err(out, ref), rel_err(xg.grad, x64.grad)
            assert e1 <= tol and e2 <= tol, (e1, e2)
    except Exception as ex:  # noqa: BLE001
        bad += 1
        print(f"FAIL n={n} e={e} feat={feat} seed={seed} hub_thresh={hub_thresh} quantum={quantum} bf16={bf16}: "
              f"{type(ex).__name__}: {str(ex)[:300]}", flush=True)
        if bad <= 3:
            traceback.print_exc()
        if bad > 20:
            break
print(f"done: {iters} examples, {bad} failures")
