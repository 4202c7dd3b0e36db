"""CUDA path vs the committed golden vectors (tests/golden/*.npz, generated from the oracle by
tests/golden/make_golden.py).  Needs nothing but numpy + the CUDA library on the GPU box."""
from pathlib import Path

import numpy as np
import pytest
import torch

import gmlm_b200 as G

from conftest import rel_err

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", ["cornell_shaped", "multi_edge_self_loop", "hub"])
@pytest.mark.parametrize("hub_thresh", [16, 1024])
def test_cuda_path_matches_golden(cuda_dev, name, hub_thresh):
    z = np.load(GOLDEN / f"{name}.npz")
    n, e, fin, hidden, out_dim = (int(v) for v in z["dims"])
    ei = torch.from_numpy(z["edge_index"]).to(cuda_dev)
    x = torch.from_numpy(z["x"]).to(cuda_dev)

    # integers: bit-exact
    assert np.array_equal(G.degree(ei[0], n).cpu().numpy(), z["degree_src"])
    et = G.edge_type_from_degree(ei, n)
    assert np.array_equal(et.cpu().numpy(), z["edge_type"])
    g = G.RelGraph.build(ei, et, n, 5, hub_thresh=hub_thresh)
    assert g.live_rels == z["live_rels"].tolist()
    for got, key in ((g.fwd.rowptr, "rowptr"), (g.fwd.col, "col"), (g.fwd.perm, "perm"), (g.bwd.rowptr, "rowptr_t"),
                     (g.bwd.col, "seg_t"), (g.bwd.perm, "perm_t")):
        assert np.array_equal(got.cpu().numpy(), z[key]), key
    assert np.array_equal(g.bwd.w.cpu().numpy().view(np.uint32), z["w_t"].view(np.uint32))

    # soft masking: bit-exact fp32
    mask = torch.from_numpy(z["mask"]).to(cuda_dev)
    token = torch.from_numpy(z["token"]).to(cuda_dev)
    xm = G.soft_masking_gnn_input(x, mask, token, 0.7)
    assert np.array_equal(xm.cpu().numpy(), z["x_masked_f32"])

    # floats: <= 1e-5 per op, compounding over the four stacked layers
    enc = G.GraphEncoder(fin, hidden, out_dim, dropout_rate=0.0)
    sd = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    enc.load_state_dict(sd, strict=False)
    with torch.no_grad():
        enc.gnn_mask_token_embed.copy_(torch.from_numpy(z["token"]))
    enc = enc.to(cuda_dev).eval()
    with torch.no_grad():
        conv1 = enc.rgcn1(x, g)
        assert rel_err(conv1, torch.from_numpy(z["conv1"])) <= 1e-5
        assert rel_err(enc.gnorm1(conv1), torch.from_numpy(z["norm1"])) <= 1e-5
        fused, layers = enc.get_graph_embeddings(x, ei, return_layers=True)
        for i, l in enumerate(layers):
            assert rel_err(l, torch.from_numpy(z[f"layer{i+1}"])) <= 5e-5, i
        assert rel_err(fused, torch.from_numpy(z["fused"])) <= 5e-5
        fused_m = enc(x, ei, gnn_perturb_mask=mask)
        assert rel_err(fused_m, torch.from_numpy(z["fused_masked"])) <= 5e-5


def test_nt_xent_matches_reference_golden(cuda_dev):
    """§8f N4: the batched NT-Xent on CUDA against vectors produced by the reference's own function
    (tests/golden/make_nt_xent_golden.py), value and input gradients, fp32 tolerance."""
    z = np.load(GOLDEN / "nt_xent.npz")
    for i, ((n, d, b), t) in enumerate(zip(z["cases"].tolist(), z["temperature"].tolist())):
        z1 = torch.from_numpy(z[f"z1_{i}"]).to(cuda_dev).requires_grad_(True)
        z2 = torch.from_numpy(z[f"z2_{i}"]).to(cuda_dev).requires_grad_(True)
        loss = G.nt_xent_loss(z1, z2, temperature=t, batch_size=None if b < 0 else b)
        loss.backward()
        assert abs(float(loss.detach()) - float(z[f"loss_{i}"])) <= 1e-5 * abs(float(z[f"loss_{i}"])), i
        assert rel_err(z1.grad, torch.from_numpy(z[f"g1_{i}"])) <= 1e-4, i      # fp32 (TF32 off) vs the fp64 reference
        assert rel_err(z2.grad, torch.from_numpy(z[f"g2_{i}"])) <= 1e-4, i
