"""N3 on the GPU: the reference's edge dropout (``augment_graph``, main.py:832-837) fused into the degree
histogram, the edge typing and the CSR build as a keep MASK must give, array for array, the graph built from the
filtered ``edge_index`` -- and the saved / reloaded CSR must aggregate identically."""
import pytest
import torch

import gmlm_b200 as G
from gmlm_b200 import _lib, synth
from gmlm_b200.ingest import build_dropped_graph, edge_dropout_mask, load_rel_graph, save_rel_graph

pytestmark = pytest.mark.gpu


def _same_csr(a, b):
    for k in ("rowptr", "col", "perm", "w", "grp_row", "hub_row", "hub_chunk_ptr", "chunk_beg", "chunk_end"):
        x, y = getattr(a, k), getattr(b, k)
        assert (x is None) == (y is None), k
        if x is not None and k != "perm":
            assert torch.equal(x, y), k
    assert (a.num_rows, a.n_hub, a.n_chunks, a.n_groups) == (b.num_rows, b.n_hub, b.n_chunks, b.n_groups)


@pytest.mark.parametrize("n,e,p", [(500, 8000, 0.1), (3000, 60000, 0.5), (40, 300, 0.9), (10, 50, 0.0)])
def test_fused_edge_dropout_equals_filtered_build(cuda_dev, n, e, p):
    ei = synth.rmat_edges(n, e, seed=n)
    keep = edge_dropout_mask(e, p, generator=torch.Generator().manual_seed(3))
    ei_d = ei.to(cuda_dev)
    fused = build_dropped_graph(ei_d, keep, n)
    ei_f = ei[:, keep].to(cuda_dev)                                         # what the reference materialises
    et_f = G.edge_type_from_degree(ei_f, n)
    want = G.RelGraph.build(ei_f, et_f, n, 5)
    assert fused.num_edges == want.num_edges == int(keep.sum()) and fused.live_rels == want.live_rels
    _same_csr(fused.fwd, want.fwd)
    _same_csr(fused.bwd, want.bwd)
    # perm refers to positions in the UNFILTERED edge list: map through the kept positions
    kept_pos = torch.nonzero(keep).squeeze(1).to(cuda_dev)
    assert torch.equal(fused.fwd.perm.long(), kept_pos[want.fwd.perm.long()])
    x = torch.randn(n, 32, device=cuda_dev)
    assert torch.equal(G.rgcn_aggregate(x, fused), G.rgcn_aggregate(x, want))


def test_rel_graph_file_round_trip(cuda_dev, tmp_path):
    n, e = 2000, 40000
    ei = synth.rmat_edges(n, e, seed=2).to(cuda_dev)
    g = G.RelGraph.build(ei, G.edge_type_from_degree(ei, n), n, 5, hub_thresh=64)
    path = str(tmp_path / "graph.pt")
    save_rel_graph(g, path)
    g2 = load_rel_graph(path, cuda_dev)
    _same_csr(g.fwd, g2.fwd)
    _same_csr(g.bwd, g2.bwd)
    x = torch.randn(n, 64, device=cuda_dev, requires_grad=True)
    h1, h2 = G.rgcn_aggregate(x, g), G.rgcn_aggregate(x, g2)
    assert torch.equal(h1, h2)
    gh = torch.randn_like(h1)
    assert torch.equal(torch.autograd.grad((h1 * gh).sum(), x)[0], torch.autograd.grad((h2 * gh).sum(), x)[0])
    with pytest.raises(G.GmlmError):
        load_rel_graph(path, "cpu")
