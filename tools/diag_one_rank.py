"""Diagnose PartitionedGraphEncoder(CudaPartitionOps) on ONE rank against GraphEncoder (prints every error)."""
import copy, os, socket, sys, traceback
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist
import gmlm_b200 as G
from gmlm_b200 import synth
from gmlm_b200.partition import build_local_part

dev = torch.device("cuda:0")
s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


try:
    n, e, fin, hidden, out_dim = 2000, 24000, 32, 8, 24
    ei = synth.rmat_edges(n, e, seed=13).to(dev)
    et = G.edge_type_from_degree(ei, n)
    live = sorted(torch.unique(et).tolist())
    x = synth.make_features(n, fin, seed=2).to(dev)
    gout = synth.make_features(n, out_dim, seed=3).to(dev)
    torch.manual_seed(0)
    enc_full = G.GraphEncoder(fin, hidden, out_dim, dropout_rate=0.0).to(dev)
    enc_rank = copy.deepcopy(enc_full)
    xf = x.clone().requires_grad_(True)
    fused_full = enc_full.get_graph_embeddings(xf, ei, et)
    (fused_full * gout).sum().backward()
    part = build_local_part(ei, et, [(0, n)], 0)
    print("part", part.n_halo, part.n_local, part.n_src)
    g = G.RelGraph.build(part.edge_index, part.edge_type, part.n_local, 5, num_src=part.n_src, live_rels=live)
    model = G.PartitionedGraphEncoder(enc_rank, G.CudaPartitionOps(part, g, n))
    xl = x.clone().requires_grad_(True)
    fused = model(xl)
    (fused * gout).sum().backward()
    G.sync_gradients(enc_rank)
    print("fused", rel(fused, fused_full))
    print("gx", rel(xl.grad, xf.grad))
    ref = dict(enc_full.named_parameters())
    for name, p in enc_rank.named_parameters():
        if ref[name].grad is None:
            print(name, "ref grad None; rank grad", None if p.grad is None else float(p.grad.abs().max()))
        elif p.grad is None:
            print(name, "rank grad None but ref has grad", float(ref[name].grad.abs().max()))
        else:
            print(name, rel(p.grad, ref[name].grad))
except Exception:
    traceback.print_exc()
finally:
    dist.destroy_process_group()
