"""The torch custom-op layer: every op is registered under ``torch.ops.gmlm`` with a schema, a CUDA
implementation and a fake (meta) implementation; the differentiable ones carry their backward through
``torch.library.register_autograd``.  ``torch.library.opcheck`` verifies schema, fake-tensor agreement and the
autograd registration."""
import pytest
import torch

import gmlm_b200 as G
from gmlm_b200 import _lib, synth
from gmlm_b200.ops import csr_pack

pytestmark = pytest.mark.gpu

CHECKS = ("test_schema", "test_faketensor", "test_autograd_registration")


def _graph(dev, n=300, e=4000):
    ei = synth.rmat_edges(n, e, seed=1).to(dev)
    et = G.edge_type_from_degree(ei, n)
    return n, ei, et, G.RelGraph.build(ei, et, n, 5, hub_thresh=32)


def test_registered_ops_are_listed(cuda_dev):
    for name in ("degree_i32", "edge_type_bucket", "csr_build", "spmm_csr", "colstats", "gemm_nt", "graphnorm_fwd",
                 "graphnorm_bwd", "layernorm_fwd", "layernorm_bwd", "soft_mask_fwd", "soft_mask_bwd", "rgcn_aggregate",
                 "plan_aggregate"):
        assert hasattr(torch.ops.gmlm, name), name


def test_opcheck_rgcn_aggregate_and_plan_aggregate(cuda_dev):
    n, ei, et, g = _graph(cuda_dev)
    x = torch.randn(n, 32, device=cuda_dev, requires_grad=True)
    fl, fm = csr_pack(g.fwd)
    bl, bm = csr_pack(g.bwd)
    torch.library.opcheck(torch.ops.gmlm.rgcn_aggregate.default, (x, *fl, *bl, fm + bm + [n, g.num_slots]),
                          test_utils=CHECKS)
    fp, bp = g.dst_plan()
    z = torch.randn(n * (g.num_slots + 1), 16, device=cuda_dev, requires_grad=True)
    fl2, fm2 = csr_pack(fp)
    bl2, bm2 = csr_pack(bp)
    torch.library.opcheck(torch.ops.gmlm.plan_aggregate.default, (z, *fl2, *bl2, fm2 + bm2), test_utils=CHECKS)
    # and the gradient itself: d/dx of sum(h * G) is the transposed aggregation
    gh = torch.randn(n, g.num_slots * 32, device=cuda_dev)
    h = torch.ops.gmlm.rgcn_aggregate(x, *fl, *bl, fm + bm + [n, g.num_slots])
    (gx,) = torch.autograd.grad((h * gh).sum(), x)
    want = G.spmm(gh.view(n * g.num_slots, 32), g.bwd, _lib.AGG_WEIGHTED)
    assert torch.equal(gx, want)


def test_opcheck_norms_and_masking(cuda_dev):
    x = torch.randn(257, 64, device=cuda_dev, requires_grad=True)
    w = torch.rand(64, device=cuda_dev, requires_grad=True)
    b = torch.rand(64, device=cuda_dev, requires_grad=True)
    ms = torch.rand(64, device=cuda_dev, requires_grad=True)
    torch.library.opcheck(torch.ops.gmlm.graphnorm_fwd.default, (x, w, b, ms, 1e-5, True), test_utils=CHECKS)
    torch.library.opcheck(torch.ops.gmlm.layernorm_fwd.default, (x, w, b, 1e-5), test_utils=CHECKS)
    mask = (torch.rand(257, device=cuda_dev) < 0.3)
    tok = torch.randn(1, 64, device=cuda_dev, requires_grad=True)
    torch.library.opcheck(torch.ops.gmlm.soft_mask_fwd.default, (x, mask, tok, 0.7), test_utils=CHECKS)


def test_opcheck_gemm_and_csr_build(cuda_dev):
    a = torch.randn(300, 128, device=cuda_dev).bfloat16()
    bm = torch.randn(64, 128, device=cuda_dev).bfloat16()
    bias = torch.randn(64, device=cuda_dev)
    torch.library.opcheck(torch.ops.gmlm.gemm_nt.default, (a, bm, bias, None, torch.float32),
                          test_utils=("test_schema", "test_faketensor"))
    n, ei, et, g = _graph(cuda_dev)
    args = (ei[1].contiguous(), ei[0].contiguous(), et, None, n, n, 5, [0, 1, 2, 3, -1], 4)
    torch.library.opcheck(torch.ops.gmlm.csr_build.default, args, test_utils=("test_schema", "test_faketensor"))
    rowptr, col, perm, seg = torch.ops.gmlm.csr_build(*args)
    assert int(rowptr[-1]) == ei.size(1) and torch.equal(col, ei[0][perm.long()].int())
