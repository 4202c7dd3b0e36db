"""Development: multi-GPU check of PeerHalo.backward_fetched (copy-engine halo-gradient return) against the
whole-graph gradient and the pushed variant.  torchrun --nproc-per-node N tools/check_fetch_multi.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

import gmlm_b200 as G
from gmlm_b200 import synth
from gmlm_b200.partition import PeerHalo, build_local_part, random_relabel


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        n, e, feat = 60_000, 900_000, 64
        ei = synth.rmat_edges(n, e, device="cpu", seed=11).to(dev)
        ei, ranges, _ = random_relabel(ei, n, world)
        et = G.edge_type_from_degree(ei, n)
        live = sorted(torch.unique(et).tolist())
        x = synth.make_features(n, feat, device="cpu", seed=5).to(dev).to(dtype)
        g_full = G.RelGraph.build(ei, et, n, 5, live_rels=live)
        S = g_full.num_slots
        gh = synth.make_features(n, S * feat, device="cpu", seed=6).to(dev).to(dtype)
        xg = x.clone().requires_grad_(True)
        G.rgcn_aggregate(xg, g_full).backward(gh)
        part = build_local_part(ei, et, ranges, rank)
        lo, hi = ranges[rank]
        g = G.RelGraph.build(part.edge_index, part.edge_type, part.n_local, 5, num_src=part.n_src, live_rels=live,
                             keep_seg=True)
        peer = PeerHalo(part, feat, dtype)
        peer.build_backward_push(g)
        ghl = gh[lo:hi].reshape(part.n_local * S, feat).contiguous()
        f1 = peer.backward_fetched(ghl).clone()
        f2 = peer.backward_fetched(ghl).clone()
        p1 = peer.backward_pushed(ghl).clone()
        torch.cuda.synchronize()
        assert torch.equal(f1, f2), "fetched backward is not deterministic"
        assert torch.equal(f1, p1), "fetched and pushed backward differ (same plan, same order: must be bit-equal)"
        err = rel(f1, xg.grad[lo:hi])
        assert err <= tol, err
        print(f"[rank {rank}] {dtype}: fetched backward ok, err {err:.2e}", flush=True)
        del peer
    dist.barrier()
    if rank == 0:
        print("FETCH_CHECK_OK", flush=True)
finally:
    dist.destroy_process_group()
