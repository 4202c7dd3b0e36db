"""Development: multi-GPU check of gmlm_b200.PartitionedGraphEncoder (CUDA ops) against the single-GPU
GraphEncoder on the whole graph.  torchrun --nproc-per-node N tools/check_dist_encoder_multi.py"""
import copy
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

import gmlm_b200 as G
from gmlm_b200 import synth
from gmlm_b200.dist_encoder import CudaPartitionOps, PartitionedGraphEncoder, sync_gradients
from gmlm_b200.partition import build_local_part, random_relabel


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    n, e, fin, hidden, out_dim = 50_000, 700_000, 64, 16, 48
    ei = synth.rmat_edges(n, e, device="cpu", seed=13).to(dev)
    ei, ranges, _ = random_relabel(ei, n, world)
    et = G.edge_type_from_degree(ei, n)
    live = sorted(torch.unique(et).tolist())
    x = synth.make_features(n, fin, device="cpu", seed=2).to(dev)
    gout = synth.make_features(n, out_dim, device="cpu", seed=3).to(dev)
    torch.manual_seed(0)
    enc_full = G.GraphEncoder(fin, hidden, out_dim, dropout_rate=0.0).to(dev)
    enc_rank = copy.deepcopy(enc_full)
    xf = x.clone().requires_grad_(True)
    fused_full = enc_full.get_graph_embeddings(xf, ei, et)
    (fused_full * gout).sum().backward()
    part = build_local_part(ei, et, ranges, rank)
    lo, hi = ranges[rank]
    g = G.RelGraph.build(part.edge_index, part.edge_type, part.n_local, 5, num_src=part.n_src, live_rels=live)
    model = PartitionedGraphEncoder(enc_rank, CudaPartitionOps(part, g, n))
    xl = x[lo:hi].clone().requires_grad_(True)
    fused = model(xl)
    (fused * gout[lo:hi]).sum().backward()
    sync_gradients(enc_rank)
    errs = {"fused": rel(fused, fused_full[lo:hi]), "grad_x": rel(xl.grad, xf.grad[lo:hi])}
    ref = dict(enc_full.named_parameters())
    for name, p in enc_rank.named_parameters():
        if ref[name].grad is None:
            assert p.grad is None, name
            continue
        errs[name] = rel(p.grad, ref[name].grad)
    worst = max(errs, key=errs.get)
    print(f"[rank {rank}] worst {worst}: {errs[worst]:.2e}; fused {errs['fused']:.2e}, grad_x {errs['grad_x']:.2e}", flush=True)
    assert errs["fused"] <= 5e-5 and max(errs.values()) <= 2e-3, errs     # fp32, four stacked GraphNorm backward passes
    dist.barrier()
    if rank == 0:
        print("DIST_ENCODER_CHECK_OK", flush=True)
finally:
    dist.destroy_process_group()
