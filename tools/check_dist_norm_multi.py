"""Development: multi-GPU check of gmlm_b200.partitioned_graph_norm (CUDA backend) against the single-GPU
GraphNorm on the whole matrix.  torchrun --nproc-per-node N tools/check_dist_norm_multi.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

import gmlm_b200 as G


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp(min=1e-30))


rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    torch.manual_seed(0)
    n, c = 100_003, 256
    x = (torch.randn(n, c, device=dev) * 2 + 1)
    gout = torch.randn(n, c, device=dev)
    norm = G.GraphNorm(c).to(dev)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5), norm.bias.uniform_(-0.5, 0.5), norm.mean_scale.uniform_(0.2, 1.2)
    cuts = [n * r // world for r in range(world)] + [n]
    lo, hi = cuts[rank], cuts[rank + 1]
    for dtype, tol in ((torch.float32, 2e-5), (torch.bfloat16, 2e-2)):
        for fuse in (False, True):
            xf = x.to(dtype).requires_grad_(True)
            norm.zero_grad()
            norm(xf, fuse_gelu=fuse).backward(gout.to(dtype))
            want = (norm(xf, fuse_gelu=fuse).detach(), xf.grad, norm.weight.grad.clone(), norm.bias.grad.clone(),
                    norm.mean_scale.grad.clone())
            w = norm.weight.detach().clone().requires_grad_(True)
            b = norm.bias.detach().clone().requires_grad_(True)
            ms = norm.mean_scale.detach().clone().requires_grad_(True)
            xl = x[lo:hi].to(dtype).requires_grad_(True)
            y = G.partitioned_graph_norm(xl, w, b, ms, n, norm.eps, fuse)
            y.backward(gout[lo:hi].to(dtype))
            errs = (rel(y, want[0][lo:hi]), rel(xl.grad, want[1][lo:hi]), rel(w.grad, want[2]), rel(b.grad, want[3]),
                    rel(ms.grad, want[4]))
            assert max(errs) <= tol, (dtype, fuse, errs)
            print(f"[rank {rank}] {dtype} gelu={fuse}: errs {['%.1e' % e for e in errs]}", flush=True)
    dist.barrier()
    if rank == 0:
        print("DIST_NORM_CHECK_OK", flush=True)
finally:
    dist.destroy_process_group()
