"""Single-GPU smoke of the multi-GPU composition modules through their CUDA backends, with a world of ONE
rank (NCCL on one device): with one rank the all-reduces are identities and N_local/N_global = 1, so
``partitioned_graph_norm`` must equal ``graph_norm`` bit for bit and ``PartitionedGraphEncoder`` must equal
``GraphEncoder.get_graph_embeddings``.  The rank logic itself (world 2 and 3, uneven shards) is covered on CPU
by tests/test_partition.py; the real multi-GPU run is tools/check_dist_{norm,encoder}_multi.py.

The file name sorts last so that it also RUNS last (it initialises a process group).

Round-1 note: this file ran behind a non-strict xfail and the encoder case failed on hardware.  Cause (found in
round 2, tools/diag_one_rank.py): ``rgcnK.bias`` feeds GraphNorm with ``mean_scale`` = 1, whose mean subtraction
cancels a bias shift exactly, so the bias gradient is mathematically ZERO and both encoders return ~1e-6 of
rounding noise there; a ratio of two noises failed the 1e-5 gate.  Every other tensor agreed to 4e-7.  The
gate now measures parameter-gradient error against the largest gradient of the same layer (a real error in a
bias gradient is O(1) against a scale of O(100); the noise is 1e-7 of it)."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist

import gmlm_b200 as G
from gmlm_b200 import synth

from conftest import rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180, method="thread")]


@pytest.fixture(scope="module")
def one_rank_group(cuda_dev):
    if dist.is_initialized():
        yield None
        return
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda_dev)
    try:
        yield None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fuse_gelu", [False, True])
def test_partitioned_graph_norm_one_rank_equals_graph_norm(cuda_dev, one_rank_group, dtype, fuse_gelu):
    torch.manual_seed(0)
    n, c = 3001, 96
    x = (torch.randn(n, c, device=cuda_dev) * 2 + 1).to(dtype)
    gout = torch.randn(n, c, device=cuda_dev).to(dtype)
    params = [torch.rand(c, device=cuda_dev) + 0.5, torch.rand(c, device=cuda_dev) - 0.5, torch.rand(c, device=cuda_dev)]
    outs = []
    for fn in (lambda xx, w, b, ms: G.graph_norm(xx, w, b, ms, 1e-5, fuse_gelu),
               lambda xx, w, b, ms: G.partitioned_graph_norm(xx, w, b, ms, n, 1e-5, fuse_gelu)):
        xx = x.clone().requires_grad_(True)
        w, b, ms = (p.clone().requires_grad_(True) for p in params)
        y = fn(xx, w, b, ms)
        y.backward(gout)
        outs.append((y.detach(), xx.grad, w.grad, b.grad, ms.grad))
    for a, bb in zip(*outs):
        assert torch.equal(a, bb)


def test_partitioned_encoder_one_rank_equals_encoder(cuda_dev, one_rank_group):
    from gmlm_b200.partition import build_local_part
    n, e, fin, hidden, out_dim = 2000, 24000, 32, 8, 24
    ei = synth.rmat_edges(n, e, seed=13).to(cuda_dev)
    et = G.edge_type_from_degree(ei, n)
    live = sorted(torch.unique(et).tolist())
    x = synth.make_features(n, fin, seed=2).to(cuda_dev)
    gout = synth.make_features(n, out_dim, seed=3).to(cuda_dev)
    torch.manual_seed(0)
    enc_full = G.GraphEncoder(fin, hidden, out_dim, dropout_rate=0.0).to(cuda_dev)
    enc_rank = copy.deepcopy(enc_full)
    xf = x.clone().requires_grad_(True)
    fused_full = enc_full.get_graph_embeddings(xf, ei, et)
    (fused_full * gout).sum().backward()
    part = build_local_part(ei, et, [(0, n)], 0)
    assert part.n_halo == 0 and part.n_local == n
    g = G.RelGraph.build(part.edge_index, part.edge_type, part.n_local, 5, num_src=part.n_src, live_rels=live)
    model = G.PartitionedGraphEncoder(enc_rank, G.CudaPartitionOps(part, g, n))
    xl = x.clone().requires_grad_(True)
    fused = model(xl)
    (fused * gout).sum().backward()
    G.sync_gradients(enc_rank)
    # same kernels on the same data; only the order in which autograd adds a tensor's several gradient
    # contributions may differ, so compare to rounding instead of bit for bit
    assert rel_err(fused, fused_full) <= 1e-6
    assert rel_err(xl.grad, xf.grad) <= 1e-5
    ref = dict(enc_full.named_parameters())
    scale = {}                                   # largest gradient magnitude per layer prefix (rgcn1, gnorm1, ...)
    for name, p in ref.items():
        if p.grad is not None:
            k = name.split(".")[0]
            scale[k] = max(scale.get(k, 0.0), float(p.grad.abs().max()))
    for name, p in enc_rank.named_parameters():
        if ref[name].grad is None:
            assert p.grad is None, name
        else:
            # a gradient that is exactly zero in exact arithmetic (rgcnK.bias under GraphNorm) holds only
            # rounding noise: measure it against the layer's gradient scale, not against itself
            assert rel_err(p.grad, ref[name].grad, floor=scale[name.split(".")[0]]) <= 1e-5, name
