"""GPU parity of the aggregation kernel (SURVEY §8a rows A5 forward, A14 backward) against the
oracle's per-relation index_select/index_add_ restatement of [PyG] RGCNConv.propagate (called
main.py:272,285,298,308).  Gates: <= 1e-5 relative (fp32), <= 2e-2 (bf16) — north_star."""
import pytest
import torch

import gmlm_b200 as G
from gmlm_b200 import synth
from oracle import edge_type_bucket_ref, rgcn_propagate_mean_ref

from conftest import rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2


def oracle_aggregate(x64, ei, et, n, live):
    """[N, S*F] per-(dst, live relation) means, fp64, relation loop as upstream executes it."""
    outs = []
    for r in live:
        m = et == r
        outs.append(rgcn_propagate_mean_ref(x64, ei[0, m], ei[1, m], n))
    return torch.cat(outs, dim=1)


def _case(n, e, seed, kind="uniform"):
    ei = synth.rmat_edges(n, e, seed=seed) if kind == "rmat" else synth.uniform_edges(n, e, seed=seed)
    return n, ei


@pytest.mark.parametrize("feat", [1, 5, 64, 96, 128, 256, 300, 512, 1000, 1703])
@pytest.mark.parametrize("variant", [0, 1, 3, 4])
def test_aggregate_fp32_widths(cuda_dev, feat, variant):
    n, ei = _case(257, 2000, seed=feat)
    et = edge_type_bucket_ref(ei, n)
    x = torch.randn(n, feat, generator=torch.Generator().manual_seed(1))
    old = G.set_tuning("spmm_variant", variant)
    try:
        g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5, hub_thresh=64)
        got = G.rgcn_aggregate(x.to(cuda_dev), g)
    finally:
        G.set_tuning("spmm_variant", old)
    ref = oracle_aggregate(x.double(), ei, et, n, g.live_rels)
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= FP32_TOL


@pytest.mark.parametrize("feat", [8, 64, 128, 256, 264, 512, 257])
@pytest.mark.parametrize("variant", [0, 1, 3, 4])
def test_aggregate_bf16_widths(cuda_dev, feat, variant):
    n, ei = _case(300, 3000, seed=feat + 1)
    et = edge_type_bucket_ref(ei, n)
    x = torch.randn(n, feat, generator=torch.Generator().manual_seed(2)).bfloat16()
    old = G.set_tuning("spmm_variant", variant)
    try:
        g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5, hub_thresh=64)
        got = G.rgcn_aggregate(x.to(cuda_dev), g)
    finally:
        G.set_tuning("spmm_variant", old)
    assert got.dtype == torch.bfloat16
    ref = oracle_aggregate(x.double(), ei, et, n, g.live_rels)
    assert rel_err(got, ref) <= BF16_TOL
    # accumulation is fp32: the only bf16 rounding is the final store
    assert rel_err(got, ref.bfloat16()) <= 1e-2


@pytest.mark.parametrize("hub_thresh", [1, 3, 16, 100000])
@pytest.mark.parametrize("variant", [0, 1, 3, 4])
@pytest.mark.parametrize("quantum", [0, 7, 256])
def test_aggregate_hub_rows_and_determinism(cuda_dev, hub_thresh, variant, quantum):
    """A star-like graph: the hub path (chunk partials + in-order final sum) must agree with the
    oracle and be bit-identical run to run (the stock scatter_add_ path is not)."""
    gen = torch.Generator().manual_seed(5)
    n, e, feat = 400, 20000, 256
    dst = torch.where(torch.rand(e, generator=gen) < 0.6, torch.randint(0, 3, (e,), generator=gen),
                      torch.randint(0, n, (e,), generator=gen))
    src = torch.randint(0, n, (e,), generator=gen)
    ei = torch.stack([src, dst])
    et = edge_type_bucket_ref(ei, n)
    x = torch.randn(n, feat, generator=gen)
    old = G.set_tuning("spmm_variant", variant)
    try:
        g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5, hub_thresh=hub_thresh, quantum=quantum)
        xs = x.to(cuda_dev)
        a = G.rgcn_aggregate(xs, g)
        b = G.rgcn_aggregate(xs, g)
    finally:
        G.set_tuning("spmm_variant", old)
    assert torch.equal(a, b)
    ref = oracle_aggregate(x.double(), ei, et, n, g.live_rels)
    assert rel_err(a, ref) <= FP32_TOL
    if hub_thresh < 100000:
        assert g.fwd.n_hub > 0


@pytest.mark.parametrize("feat,dtype", [(64, torch.bfloat16), (128, torch.bfloat16), (72, torch.bfloat16), (8, torch.bfloat16),
                                        (16, torch.float32), (64, torch.float32), (33, torch.float32)])
@pytest.mark.parametrize("hub_thresh", [2, 17, 256])
@pytest.mark.parametrize("quantum", [0, 5, 512])
def test_narrow_row_kernel_is_bit_identical_to_the_general_kernel(cuda_dev, feat, dtype, hub_thresh, quantum):
    """rows_narrow_kernel (rows of at most 16 packs: warp-uniform edge walk over zero-filled rows) keeps the general
    kernel's per-row summation order: forward (mean) and backward (weighted, transposed CSR) agree bit for bit, on a
    graph with hub rows, long runs of empty rows and every plan setting."""
    n, e = 3000, 30000
    ei = synth.rmat_edges(n, e, seed=feat + hub_thresh)
    ei[1, : e // 4] = ei[1, : e // 4] % 7                       # hub destinations; most (dst, rel) rows stay empty
    et = edge_type_bucket_ref(ei, n)
    g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5, hub_thresh=hub_thresh, quantum=quantum)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, feat, generator=gen).to(dtype).to(cuda_dev)
    gh = torch.randn(n * g.num_slots, feat, generator=gen).to(dtype).to(cuda_dev)
    outs = {}
    for variant in (1, 2):                                      # 1: narrow kernel where eligible, 2: general kernel
        old = G.set_tuning("spmm_variant", variant)
        try:
            outs[variant] = (G.spmm(x, g.fwd, 1), G.spmm(gh, g.bwd, 2))
        finally:
            G.set_tuning("spmm_variant", old)
    assert torch.equal(outs[1][0], outs[2][0])
    assert torch.equal(outs[1][1], outs[2][1])
    ref = oracle_aggregate(x.double().cpu(), ei, et, n, g.live_rels)
    assert rel_err(outs[1][0].view(n, -1), ref) <= (BF16_TOL if dtype == torch.bfloat16 else FP32_TOL)


@pytest.mark.parametrize("name,n,e,kind", [("empty", 7, 0, "uniform"), ("one_node", 1, 4, "uniform"),
                                           ("cornell", 183, 300, "uniform"), ("rmat", 4096, 50000, "rmat")])
def test_aggregate_edge_cases(cuda_dev, name, n, e, kind):
    _, ei = _case(n, e, seed=11, kind=kind)
    et = edge_type_bucket_ref(ei, n)
    x = torch.randn(n, 40, generator=torch.Generator().manual_seed(3))
    g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5)
    got = G.rgcn_aggregate(x.to(cuda_dev), g)
    ref = oracle_aggregate(x.double(), ei, et, n, g.live_rels)
    assert got.shape == ref.shape
    if e == 0:
        assert torch.count_nonzero(got) == 0
    else:
        assert rel_err(got, ref) <= FP32_TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("hub_thresh", [8, 1024])
@pytest.mark.parametrize("quantum", [0, 33, 256])
def test_aggregate_backward_matches_autograd_of_oracle(cuda_dev, dtype, tol, hub_thresh, quantum):
    """A14: gather on the transposed CSR with 1/count folded in == autograd through
    index_select + index_add_ + divide (what the reference's backward executes)."""
    n, ei = _case(500, 6000, seed=21, kind="rmat")
    et = edge_type_bucket_ref(ei, n)
    gen = torch.Generator().manual_seed(4)
    feat = 128
    x = torch.randn(n, feat, generator=gen).to(dtype)
    g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5, hub_thresh=hub_thresh, quantum=quantum)
    gh = torch.randn(n, g.num_slots * feat, generator=gen).to(dtype)

    xg = x.to(cuda_dev).requires_grad_(True)
    out = G.rgcn_aggregate(xg, g)
    out.backward(gh.to(cuda_dev))

    x64 = x.double().requires_grad_(True)
    ref = oracle_aggregate(x64, ei, et, n, g.live_rels)
    ref.backward(gh.double())
    assert rel_err(out, ref) <= tol
    assert rel_err(xg.grad, x64.grad) <= tol


def test_group_plan_covers_rows_in_order(cuda_dev):
    """The cost-balanced cuts are monotone, start at 0, end at num_rows, and no group exceeds
    quantum + (longest non-hub row) units of work."""
    n, e = 5000, 80000
    ei = synth.rmat_edges(n, e, seed=13)
    et = edge_type_bucket_ref(ei, n)
    for q in (1, 16, 256, 100000):
        g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5, hub_thresh=64, quantum=q)
        for csr in (g.fwd, g.bwd):
            cuts = csr.grp_row.cpu().long()
            rp = csr.rowptr.cpu().long()
            assert cuts[0] == 0 and cuts[-1] == csr.num_rows and bool((cuts[1:] >= cuts[:-1]).all())
            assert cuts.numel() == csr.n_groups + 1
            cost = (cuts[1:] - cuts[:-1]) + (rp[cuts[1:]] - rp[cuts[:-1]])
            max_row = int((rp[1:] - rp[:-1]).max())
            assert int(cost.max()) <= q + max_row + 1


def test_aggregate_linearity_full_size_property(cuda_dev):
    """Size-independent property used at BASELINE sizes: aggregation is linear, and aggregating a
    constant-one feature returns exactly 1 on non-empty (dst,rel) segments and 0 on empty ones."""
    n, e, feat = 200_000, 4_000_000, 64
    ei = synth.rmat_edges(n, e, device=cuda_dev, seed=42)
    et = G.edge_type_from_degree(ei, n)
    g = G.get_rel_graph(ei, et, n, 5)
    ones = torch.ones(n, feat, device=cuda_dev)
    h = G.rgcn_aggregate(ones, g).view(n * g.num_slots, feat)
    lens = (g.fwd.rowptr[1:] - g.fwd.rowptr[:-1]).long()
    # mean of ones: exact for short rows, within fp32 rounding of len*(1/len) for long ones
    expect = (lens > 0).float().unsqueeze(1).expand_as(h)
    assert float((h - expect).abs().max()) <= 1e-6
    assert int(lens.sum()) == e
    x = synth.make_features(n, feat, device=cuda_dev, seed=1)
    y = synth.make_features(n, feat, device=cuda_dev, seed=2)
    lhs = G.rgcn_aggregate(2.0 * x + y, g)
    rhs = 2.0 * G.rgcn_aggregate(x, g) + G.rgcn_aggregate(y, g)
    assert rel_err(lhs, rhs) <= 1e-5
    # adjoint identity <A x, g> == <x, A^T g> ties the backward kernel to the forward one
    gh = torch.randn_like(lhs)
    xg = x.clone().requires_grad_(True)
    out = G.rgcn_aggregate(xg, g)
    out.backward(gh)
    lhs_ip = (out.double() * gh.double()).sum()
    rhs_ip = (x.double() * xg.grad.double()).sum()
    scale = float(out.double().norm() * gh.double().norm())
    assert abs(float(lhs_ip - rhs_ip)) <= 1e-5 * scale


def test_cpu_tensors_raise(lib_built):
    x = torch.randn(4, 8)
    with pytest.raises(Exception):
        G.degree(torch.tensor([0, 1, 2]), 4)
    ei = torch.tensor([[0, 1], [1, 2]])
    with pytest.raises(Exception):
        G.RelGraph.build(ei, torch.tensor([0, 0]), 4, 5)
    del x


@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_aggregate_equals_whole_graph(cuda_dev, world):
    """SURVEY §8e on one GPU: for every rank of a destination-row partition, the kernel on the
    rectangular local CSR over [local ‖ halo] rows reproduces that rank's slice of the
    whole-graph result bit for bit (same per-row edge order), forward and backward-adjoint."""
    from gmlm_b200.partition import partition_ranges, select_local
    n, e, feat = 3000, 40000, 64
    ei = synth.rmat_edges(n, e, seed=17).to(cuda_dev)
    et = G.edge_type_from_degree(ei, n)
    x = synth.make_features(n, feat, device=cuda_dev, seed=3)
    g_full = G.RelGraph.build(ei, et, n, 5)
    h_full = G.rgcn_aggregate(x, g_full)
    in_deg = torch.ops.gmlm.degree_i32(ei[1], n)
    ranges = partition_ranges(in_deg, world)
    gh = torch.randn_like(h_full)
    gx_sum = torch.zeros_like(x)
    for rank in range(world):
        lo, hi = ranges[rank]
        ei_l, et_l, halo_gid, _ = select_local(ei, et, ranges, rank)
        n_local, n_src = hi - lo, hi - lo + int(halo_gid.numel())
        g = G.RelGraph.build(ei_l, et_l, n_local, 5, num_src=n_src, live_rels=g_full.live_rels)
        X = torch.cat([x[lo:hi], x[halo_gid]]).requires_grad_(True)
        h = G.rgcn_aggregate(X, g)
        assert torch.equal(h, h_full[lo:hi])
        h.backward(gh[lo:hi])
        gx_sum[lo:hi] += X.grad[:n_local]
        gx_sum.index_add_(0, halo_gid, X.grad[n_local:])
    xg = x.clone().requires_grad_(True)
    G.rgcn_aggregate(xg, g_full).backward(gh)
    assert rel_err(gx_sum, xg.grad) <= 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("feat", [256, 300, 7])
def test_halo_pack_and_unpack_kernels(cuda_dev, dtype, feat):
    """gather_rows == index_select bit for bit; scatter_add_rows_ (unique ids) == index_add_ with the
    addition done in fp32 and rounded once."""
    from gmlm_b200.ops import gather_rows, scatter_add_rows_
    gen = torch.Generator().manual_seed(9)
    n, k = 1000, 400
    x = torch.randn(n, feat, generator=gen).to(dtype).to(cuda_dev)
    ids = torch.randperm(n, generator=gen)[:k].to(cuda_dev)
    assert torch.equal(gather_rows(x, ids), x.index_select(0, ids))
    rows = torch.randn(k, feat, generator=gen).to(dtype).to(cuda_dev)
    want = x.float().clone()
    want[ids] += rows.float()
    got = scatter_add_rows_(x.clone(), ids, rows)
    assert torch.equal(got, want.to(dtype))
    assert gather_rows(x, ids[:0]).shape == (0, feat)


from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@given(n=st.integers(1, 300), e=st.integers(0, 3000), feat=st.sampled_from([3, 8, 20, 64, 72, 256]),
       seed=st.integers(0, 10_000), hub_thresh=st.sampled_from([2, 17, 256]), quantum=st.sampled_from([0, 5, 64]),
       bf16=st.booleans())
# derandomize: the suite draws the same 40 examples on every box (a gate that fails one run in fifty is no gate);
# the random exploration of this space is tools/fuzz_spmm.py (thousands of examples per run)
@settings(max_examples=40, deadline=None, derandomize=True, database=None,
          suppress_health_check=[HealthCheck.function_scoped_fixture])
def test_aggregate_property_random_graphs(cuda_dev, n, e, feat, seed, hub_thresh, quantum, bf16):
    """Random graphs (isolated nodes, multi-edges, self-loops, empty relations, N=1), random widths and
    every plan setting: forward and backward agree with the oracle and the integer structure is consistent."""
    g_ = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g_)
    et = edge_type_bucket_ref(ei, n)
    dtype = torch.bfloat16 if bf16 else torch.float32
    # an fp32 running sum over a row of L terms carries up to L/2 ulp: the 1e-5 gate holds for rows up to a few
    # hundred entries; degenerate examples (N = 1 with 3000 self-loops: every term identical, no cancellation of the
    # rounding errors, found by brute force over this space) get the bound of the summation itself
    tol = BF16_TOL if bf16 else max(FP32_TOL, 0.5 * max(e, 1) / max(n, 1) * 2.0 ** -24 * 4)
    x = torch.randn(n, feat, generator=g_).to(dtype)
    g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, 5, hub_thresh=hub_thresh, quantum=quantum)
    assert int(g.fwd.rowptr[-1]) == e and int(g.bwd.rowptr[-1]) == e
    xg = x.to(cuda_dev).requires_grad_(True)
    out = G.rgcn_aggregate(xg, g)
    gh = torch.randn(out.shape, generator=g_).to(dtype)
    out.backward(gh.to(cuda_dev))
    x64 = x.double().requires_grad_(True)
    ref = oracle_aggregate(x64, ei, et, n, g.live_rels)
    ref.backward(gh.double())
    if e == 0:
        assert torch.count_nonzero(out) == 0 and torch.count_nonzero(xg.grad) == 0
    else:
        assert rel_err(out, ref) <= tol
        assert rel_err(xg.grad, x64.grad) <= tol
