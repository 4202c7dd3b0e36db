"""GPU parity, integer side (bit-exact gates of north_star): degree, edge typing, the
(dst,rel)-keyed CSR, its transpose and the hub plan — CUDA path (through the C ABI) vs the
oracle restatement.  Reference call sites: main.py:65,256 (degree), main.py:253-267 (edge
typing), main.py:272,285,298,308 (per-relation split inside RGCNConv)."""
import numpy as np
import pytest
import torch

import gmlm_b200 as G
from gmlm_b200 import synth
from oracle import degree_ref, edge_type_bucket_ref, edge_type_loop_ref, rel_csr_ref, transposed_csr_ref

pytestmark = pytest.mark.gpu


def _graphs():
    g = torch.Generator().manual_seed(7)
    out = {
        "uniform_small": (50, torch.randint(0, 50, (2, 400), generator=g)),
        "multi_self": (6, torch.tensor([[0, 0, 0, 1, 2, 2, 5, 5, 3], [1, 1, 0, 1, 4, 4, 5, 0, 3]])),
        "isolated": (40, torch.randint(0, 10, (2, 100), generator=g)),          # nodes 10..39 isolated
        "single_node": (1, torch.zeros((2, 3), dtype=torch.long)),
        "no_edges": (5, torch.zeros((2, 0), dtype=torch.long)),
        "cornell": (183, synth.uniform_edges(183, 300, seed=42)),
        "rmat": (5000, synth.rmat_edges(5000, 60000, seed=3)),
        "star_hub": (300, torch.stack([torch.randint(0, 300, (5000,), generator=g),
                                       torch.where(torch.rand(5000, generator=g) < 0.7, 7,
                                                   torch.randint(0, 300, (5000,), generator=g))])),
    }
    return out


GRAPHS = _graphs()


@pytest.mark.parametrize("name", list(GRAPHS))
def test_degree_bit_exact(cuda_dev, name):
    n, ei = GRAPHS[name]
    for row in (0, 1):
        got = G.degree(ei[row].to(cuda_dev), num_nodes=n)
        ref = degree_ref(ei[row], n)
        assert got.dtype == torch.float32 and got.shape == (n,)
        assert torch.equal(got.cpu(), ref)


def test_degree_infers_num_nodes(cuda_dev):
    idx = torch.tensor([0, 3, 3, 9])
    got = G.degree(idx.to(cuda_dev))
    assert got.shape == (10,)
    assert torch.equal(got.cpu(), degree_ref(idx, 10))


@pytest.mark.parametrize("name", list(GRAPHS))
def test_edge_type_bit_exact(cuda_dev, name):
    n, ei = GRAPHS[name]
    got = G.edge_type_from_degree(ei.to(cuda_dev), n)
    ref = edge_type_bucket_ref(ei, n)
    assert got.dtype == torch.int64
    assert torch.equal(got.cpu(), ref)
    if ei.size(1) <= 500:   # the verbatim per-edge loop of main.py:257-267
        assert torch.equal(got.cpu(), edge_type_loop_ref(ei, n))


@pytest.mark.parametrize("name", list(GRAPHS))
@pytest.mark.parametrize("hub_thresh", [4, 1024])
def test_rel_csr_bit_exact(cuda_dev, name, hub_thresh):
    n, ei = GRAPHS[name]
    R = 5
    et = edge_type_bucket_ref(ei, n)
    g = G.RelGraph.build(ei.to(cuda_dev), et.to(cuda_dev), n, R, hub_thresh=hub_thresh)
    live = sorted(set(et.tolist())) or [0]
    assert g.live_rels == live
    slot_of = {r: s for s, r in enumerate(live)}
    slot = np.array([slot_of[int(t)] for t in et.tolist()], dtype=np.int64)
    src, dst = ei[0].numpy(), ei[1].numpy()
    rowptr, col, perm = rel_csr_ref(src, dst, slot, n, len(live))
    assert np.array_equal(g.fwd.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.fwd.col.cpu().numpy(), col)
    assert np.array_equal(g.fwd.perm.cpu().numpy(), perm)
    rowptr_t, seg_t, w_t, perm_t = transposed_csr_ref(src, dst, slot, n, len(live))
    assert np.array_equal(g.bwd.rowptr.cpu().numpy(), rowptr_t)
    assert np.array_equal(g.bwd.col.cpu().numpy(), seg_t)
    assert np.array_equal(g.bwd.perm.cpu().numpy(), perm_t)
    assert np.array_equal(g.bwd.w.cpu().numpy().view(np.uint32), w_t.view(np.uint32))   # bit-exact floats
    # hub plan: exactly the rows longer than the threshold, chunks tile each hub row in order
    for csr, rp in ((g.fwd, rowptr), (g.bwd, rowptr_t)):
        lens = np.diff(rp.astype(np.int64))
        hubs = np.nonzero(lens > hub_thresh)[0]
        assert csr.n_hub == len(hubs)
        if len(hubs):
            assert np.array_equal(csr.hub_row.cpu().numpy(), hubs.astype(np.int32))
            cptr = csr.hub_chunk_ptr.cpu().numpy()
            cb, ce = csr.chunk_beg.cpu().numpy(), csr.chunk_end.cpu().numpy()
            assert cptr[0] == 0 and cptr[-1] == csr.n_chunks == len(cb)
            for h, r in enumerate(hubs):
                b = cb[cptr[h]:cptr[h + 1]]
                e = ce[cptr[h]:cptr[h + 1]]
                assert b[0] == rp[r] and e[-1] == rp[r + 1]
                assert np.array_equal(b[1:], e[:-1])
                assert np.all(e - b <= hub_thresh) and np.all(e - b > 0)


@pytest.mark.parametrize("n,e,hub_thresh", [(3_000_000, 12_000_000, 256), (700_001, 5_000_003, 16)])
def test_rel_csr_bit_exact_at_scale(cuda_dev, n, e, hub_thresh):
    """The library's own scan / stable radix sort / selection at sizes with several scan levels (12 M keys = 2930 sort
    tiles, a 750 k-entry digit table, 12 M segment counters) against torch's stable sort on the same device."""
    ei = synth.rmat_edges(n, e, seed=13).to(cuda_dev)
    et = G.edge_type_from_degree(ei, n)
    g = G.RelGraph.build(ei, et, n, 5, hub_thresh=hub_thresh)
    live = torch.tensor(g.live_rels, device=cuda_dev)
    slot = torch.searchsorted(live, et)
    S = len(g.live_rels)
    key = ei[1] * S + slot
    order = torch.sort(key, stable=True).indices
    counts = torch.bincount(key, minlength=n * S)
    rowptr = torch.zeros(n * S + 1, dtype=torch.int64, device=cuda_dev)
    rowptr[1:] = torch.cumsum(counts, 0)
    assert torch.equal(g.fwd.rowptr.long(), rowptr)
    assert torch.equal(g.fwd.perm.long(), order)
    assert torch.equal(g.fwd.col.long(), ei[0][order])
    order_t = torch.sort(ei[0], stable=True).indices                    # transposed CSR: rows = source nodes
    rowptr_t = torch.zeros(n + 1, dtype=torch.int64, device=cuda_dev)
    rowptr_t[1:] = torch.cumsum(torch.bincount(ei[0], minlength=n), 0)
    assert torch.equal(g.bwd.rowptr.long(), rowptr_t)
    assert torch.equal(g.bwd.perm.long(), order_t)
    assert torch.equal(g.bwd.col.long(), key[order_t])
    hubs = torch.nonzero(counts > hub_thresh).flatten()
    assert g.fwd.n_hub == hubs.numel()
    if hubs.numel():
        assert torch.equal(g.fwd.hub_row.long(), hubs)


def test_csr_rejects_bad_indices(cuda_dev):
    ei = torch.tensor([[0, 1, 9], [1, 2, 0]]).to(cuda_dev)
    et = torch.tensor([0, 1, 2]).to(cuda_dev)
    with pytest.raises(G.GmlmError):
        G.RelGraph.build(ei, et, 5, 5)                      # node 9 out of range
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]]).to(cuda_dev)
    with pytest.raises(G.GmlmError):
        G.RelGraph.build(ei, torch.tensor([0, 1, 7]).to(cuda_dev), 5, 5)   # relation 7 out of range


def test_graph_cache_identity_and_content(cuda_dev):
    """Identity is the fast path; CONTENT is what the one-line import swap needs: the reference re-creates its
    edge_type tensor on every call (main.py:255), which must not rebuild the CSR."""
    from gmlm_b200 import graph as gg
    G.clear_graph_cache()
    n, ei = GRAPHS["cornell"]
    ei = ei.to(cuda_dev)
    et = G.edge_type_from_degree(ei, n)
    a = G.get_rel_graph(ei, et, n, 5)
    before = dict(gg.cache_stats)
    b = G.get_rel_graph(ei, et, n, 5)
    assert a is b and gg.cache_stats["identity_hits"] == before["identity_hits"] + 1
    et2 = et.clone()                   # same contents, new tensor: no rebuild
    c = G.get_rel_graph(ei.clone(), et2, n, 5)
    assert c is a and gg.cache_stats["content_hits"] == before["content_hits"] + 1
    assert gg.cache_stats["builds"] == before["builds"]
    et2[0] = (et2[0] + 1) % 4          # in-place edit bumps _version AND changes the contents -> new graph
    d = G.get_rel_graph(ei, et2, n, 5)
    assert d is not a and gg.cache_stats["builds"] == before["builds"] + 1
    perm = torch.randperm(ei.size(1), device=cuda_dev)            # same multiset of edges, different order:
    e = G.get_rel_graph(ei[:, perm].contiguous(), et[perm].contiguous(), n, 5)   # a different CSR edge order
    assert e is not a


def test_reference_style_calls_build_the_csr_once(cuda_dev):
    """What main.py:250-267 does under the import swap: a fresh edge_type per get_graph_embeddings call and four
    conv calls with it.  One CSR build in total."""
    from gmlm_b200 import graph as gg
    G.clear_graph_cache()
    n, ei = GRAPHS["star_hub"]
    ei = ei.to(cuda_dev)
    conv = G.RGCNConv(16, 8, 5, 30).to(cuda_dev)
    x = torch.randn(n, 16, device=cuda_dev)
    builds0 = gg.cache_stats["builds"]
    outs = []
    for _ in range(3):
        et = G.edge_type_from_degree(ei, n)      # fresh tensor every call, as the reference's loop produces
        for _ in range(4):
            outs.append(conv(x, ei, et))
    assert gg.cache_stats["builds"] == builds0 + 1
    assert all(torch.equal(o, outs[0]) for o in outs)
