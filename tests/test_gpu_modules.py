"""GPU parity of the drop-in modules against the oracle restatement of the reference path:
RGCNConv (main.py:189/272), GraphNorm (main.py:190/273), soft masking (main.py:92-99) and the
whole encoder body (main.py:250-320), forward and parameter/input gradients."""
import pytest
import torch
import torch.nn.functional as F

import gmlm_b200 as G
from gmlm_b200 import synth
from oracle import (EncoderRef, GraphNormRef, RGCNConvRef, edge_type_bucket_ref, soft_masking_ref)

from conftest import rel_err

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2


def _copy_params(dst: torch.nn.Module, src: torch.nn.Module):
    missing, unexpected = dst.load_state_dict(src.state_dict(), strict=False)
    return missing, unexpected


@pytest.mark.parametrize("n,e,fi,fo", [(183, 300, 1703, 64), (500, 4000, 300, 96), (64, 900, 32, 256), (1, 3, 16, 8)])
def test_rgcn_conv_matches_oracle(cuda_dev, n, e, fi, fo):
    torch.manual_seed(0)
    ei = synth.uniform_edges(n, e, seed=n)
    et = edge_type_bucket_ref(ei, n)
    ref = RGCNConvRef(fi, fo, num_relations=5, num_bases=30).double()
    with torch.no_grad():
        ref.bias.uniform_(-0.1, 0.1)
    mod = G.RGCNConv(fi, fo, num_relations=5, num_bases=30)
    assert set(mod.state_dict()) == set(ref.state_dict()) == {"weight", "comp", "root", "bias"}
    for k, v in mod.state_dict().items():
        assert v.shape == ref.state_dict()[k].shape
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mod = mod.to(cuda_dev)
    x = torch.randn(n, fi)
    gout = torch.randn(n, fo)

    x64 = x.double().requires_grad_(True)
    y_ref = ref(x64, ei, et)
    y_ref.backward(gout.double())

    xg = x.to(cuda_dev).requires_grad_(True)
    y = mod(xg, ei.to(cuda_dev), et.to(cuda_dev))
    assert y.dtype == torch.float32
    y.backward(gout.to(cuda_dev))
    assert rel_err(y, y_ref) <= FP32_TOL
    assert rel_err(xg.grad, x64.grad) <= FP32_TOL
    for name in ("weight", "comp", "root", "bias"):
        assert rel_err(getattr(mod, name).grad, getattr(ref, name).grad) <= 2e-5, name
    # relation 4 is never produced by the degree buckets: exact-zero comp gradient
    assert torch.count_nonzero(mod.comp.grad[4]) == 0


def test_rgcn_conv_autocast(cuda_dev):
    """Callers run the encoder under torch.amp.autocast (main.py:446,543): fp32 in, fp32 out,
    half-precision GEMMs; tolerance 2e-2."""
    n, e, fi, fo = 300, 2500, 128, 64
    ei = synth.uniform_edges(n, e, seed=9)
    et = edge_type_bucket_ref(ei, n)
    ref = RGCNConvRef(fi, fo, 5, 30).double()
    mod = G.RGCNConv(fi, fo, 5, 30)
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mod = mod.to(cuda_dev)
    x = torch.randn(n, fi)
    with torch.amp.autocast("cuda"):
        y = mod(x.to(cuda_dev), ei.to(cuda_dev), et.to(cuda_dev))
    assert y.dtype == torch.float32
    assert rel_err(y, ref(x.double(), ei, et)) <= BF16_TOL


@pytest.mark.parametrize("n,c", [(2, 8), (183, 512), (1000, 300), (4097, 64), (333, 1703)])
@pytest.mark.parametrize("fuse_gelu", [False, True])
def test_graph_norm_matches_oracle(cuda_dev, n, c, fuse_gelu):
    torch.manual_seed(1)
    ref = GraphNormRef(c).double()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.uniform_(-0.5, 0.5)
        ref.mean_scale.uniform_(0.2, 1.2)
    mod = G.GraphNorm(c)
    assert set(mod.state_dict()) == {"weight", "bias", "mean_scale"}
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mod = mod.to(cuda_dev)
    x = torch.randn(n, c) * 2.0 + 3.0          # non-zero mean: exercises the shifted-variance identity
    gout = torch.randn(n, c)
    x64 = x.double().requires_grad_(True)
    y_ref = ref(x64)
    if fuse_gelu:
        y_ref = F.gelu(y_ref)
    y_ref.backward(gout.double())
    xg = x.to(cuda_dev).requires_grad_(True)
    y = mod(xg, fuse_gelu=fuse_gelu)
    y.backward(gout.to(cuda_dev))
    assert rel_err(y, y_ref) <= FP32_TOL
    assert rel_err(xg.grad, x64.grad) <= 2e-5
    for name in ("weight", "bias", "mean_scale"):
        assert rel_err(getattr(mod, name).grad, getattr(ref, name).grad) <= 2e-5, name


def test_graph_norm_bf16(cuda_dev):
    n, c = 2000, 256
    ref = GraphNormRef(c).double()
    mod = G.GraphNorm(c).to(cuda_dev)
    x = (torch.randn(n, c) + 1.0).bfloat16()
    y = mod(x.to(cuda_dev), fuse_gelu=True)
    assert y.dtype == torch.bfloat16
    assert rel_err(y, F.gelu(ref(x.double()))) <= BF16_TOL


@pytest.mark.parametrize("n,c", [(1, 8), (183, 768), (1000, 300), (4097, 64), (50, 1024), (300000, 96)])
def test_layer_norm_matches_torch_fp64(cuda_dev, n, c):
    """A13: the LayerNorm closing MultiScaleFusion (main.py:171,180) is nn.LayerNorm in the reference, so
    nn.LayerNorm in fp64 is the oracle; forward, input gradient and both parameter gradients."""
    torch.manual_seed(2)
    ref = torch.nn.LayerNorm(c).double()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5)
        ref.bias.uniform_(-0.5, 0.5)
    x = torch.randn(n, c) * 2.0 + 3.0
    gout = torch.randn(n, c)
    x64 = x.double().requires_grad_(True)
    ref(x64).backward(gout.double())
    w = ref.weight.detach().float().to(cuda_dev).requires_grad_(True)
    b = ref.bias.detach().float().to(cuda_dev).requires_grad_(True)
    xg = x.to(cuda_dev).requires_grad_(True)
    y = G.layer_norm(xg, w, b, ref.eps)
    y.backward(gout.to(cuda_dev))
    assert rel_err(y, ref(x64)) <= FP32_TOL
    assert rel_err(xg.grad, x64.grad) <= 2e-5
    assert rel_err(w.grad, ref.weight.grad) <= 2e-5
    assert rel_err(b.grad, ref.bias.grad) <= 2e-5
    # deterministic parameter gradients (fixed-order two-stage sums)
    w2 = w.detach().clone().requires_grad_(True)
    G.layer_norm(xg.detach(), w2, b.detach(), ref.eps).backward(gout.to(cuda_dev))
    assert torch.equal(w2.grad, w.grad)


def test_layer_norm_bf16_and_fusion_module(cuda_dev):
    n, c = 3000, 768
    x = (torch.randn(n, c) + 0.5).bfloat16()
    ref = torch.nn.LayerNorm(c).double()
    y = G.layer_norm(x.to(cuda_dev), ref.weight.float().to(cuda_dev), ref.bias.float().to(cuda_dev), ref.eps)
    assert y.dtype == torch.bfloat16
    assert rel_err(y, ref(x.double())) <= BF16_TOL
    # the module routes through the kernel on CUDA and equals its stock formulation
    fus = G.MultiScaleFusion([16, 32], 64).to(cuda_dev)
    xs = [torch.randn(40, 16, device=cuda_dev), torch.randn(40, 32, device=cuda_dev)]
    w = F.softmax(fus.scale_weights, dim=0)
    want = fus.layer_norm(sum(w[i] * fus.projections[i](xs[i]) for i in range(2)))
    assert rel_err(fus(xs), want) <= FP32_TOL
    with pytest.raises(G.GmlmError):
        G.layer_norm(torch.randn(4, 6, device=cuda_dev), torch.ones(6, device=cuda_dev),
                     torch.zeros(6, device=cuda_dev))


@pytest.mark.parametrize("n,f", [(183, 1703), (1000, 300), (50, 256)])
@pytest.mark.parametrize("ratio", [0.0, 0.3, 1.0])
def test_soft_mask_bit_exact_and_grad(cuda_dev, n, f, ratio):
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(n, f, generator=gen)
    mask = torch.rand(n, generator=gen) < ratio
    token = torch.randn(1, f, generator=gen) * 0.1
    ref = soft_masking_ref(x, mask, token, beta=0.7)
    tok = token.to(cuda_dev).requires_grad_(True)
    xg = x.to(cuda_dev).requires_grad_(True)
    got = G.soft_masking_gnn_input(xg, mask.to(cuda_dev), tok, beta=0.7)
    assert torch.equal(got.cpu(), ref)                       # same rounding sequence as main.py:98
    gout = torch.randn(n, f, generator=gen)
    got.backward(gout.to(cuda_dev))
    t64 = token.double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    soft_masking_ref(x64, mask, t64, beta=0.7).backward(gout.double())
    if t64.grad is None:
        assert torch.count_nonzero(tok.grad) == 0
    else:
        assert rel_err(tok.grad, t64.grad) <= FP32_TOL
    assert rel_err(xg.grad, x64.grad) <= FP32_TOL


def _encoder_pair(fin, hidden, out_dim, dropout=0.0):
    torch.manual_seed(5)
    ref = EncoderRef(fin, hidden, out_dim, dropout_rate=dropout, use_checkpoint=False).double()
    with torch.no_grad():
        for k in range(1, 5):
            getattr(ref, f"gnorm{k}").mean_scale.uniform_(0.5, 1.0)
            getattr(ref, f"rgcn{k}").bias.uniform_(-0.1, 0.1)
    mod = G.GraphEncoder(fin, hidden, out_dim, dropout_rate=dropout)
    missing, unexpected = mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()}, strict=False)
    assert unexpected == [] and missing == ["gnn_mask_token_embed"]
    return ref, mod


@pytest.mark.parametrize("n,e,fin,hidden", [(183, 300, 1703, 32), (2000, 9000, 300, 16), (1, 2, 24, 8)])
@pytest.mark.parametrize("use_checkpoint", [False, True])
def test_encoder_matches_oracle(cuda_dev, n, e, fin, hidden, use_checkpoint):
    """Whole get_graph_embeddings body (main.py:250-320), eval-free (dropout p=0), fp32."""
    ref, mod = _encoder_pair(fin, hidden, 48)
    mod = mod.to(cuda_dev)
    mod.use_checkpoint = use_checkpoint
    ei = synth.uniform_edges(n, e, seed=n + 1)
    x = torch.randn(n, fin)
    fused_ref, layers_ref = ref(x.double(), ei, return_layers=True)
    fused, layers = mod.get_graph_embeddings(x.to(cuda_dev), ei.to(cuda_dev), return_layers=True)
    for a, b in zip(layers, layers_ref):
        assert rel_err(a, b) <= 5e-5     # four stacked layers: per-layer 1e-5 compounds
    assert rel_err(fused, fused_ref) <= 5e-5
    gout = torch.randn(n, 48)
    fused_ref.backward(gout.double())
    fused.backward(gout.to(cuda_dev))
    # yardstick for the gradients: the SAME oracle run in fp32 on the CPU.  Back-propagating
    # through four GraphNorms is ill-conditioned enough that fp32 itself sits near 2e-4 here.
    import copy
    ref32 = copy.deepcopy(ref).float()
    ref32.zero_grad()
    ref32(x, ei).backward(gout)
    sd_ref = dict(ref.named_parameters())
    sd32 = dict(ref32.named_parameters())
    for name, p in mod.named_parameters():
        if name == "gnn_mask_token_embed":
            continue
        if sd_ref[name].grad is None:
            # residual_proj3 is dead (main.py:317-318); GraphNorm is skipped when N == 1 (main.py:273)
            assert p.grad is None, name
            continue
        assert p.grad is not None, name
        noise = rel_err(sd32[name].grad, sd_ref[name].grad)
        assert rel_err(p.grad, sd_ref[name].grad) <= max(2e-5, 3.0 * noise), (name, noise)


def test_encoder_with_soft_mask_and_autocast(cuda_dev):
    """Caller #1/#2 shape (main.py:443-448): soft-masked input, autocast region."""
    n, e, fin, hidden = 600, 3000, 300, 16
    ref, mod = _encoder_pair(fin, hidden, 32)
    mod = mod.to(cuda_dev)
    ei = synth.uniform_edges(n, e, seed=77)
    gen = torch.Generator().manual_seed(8)
    x = torch.randn(n, fin, generator=gen)
    mask = torch.rand(n, generator=gen) < 0.3
    tok = mod.gnn_mask_token_embed.detach().cpu()
    fused_ref = ref(soft_masking_ref(x.double(), mask, tok.double()), ei)
    with torch.amp.autocast("cuda"):
        fused = mod(x.to(cuda_dev), ei.to(cuda_dev), gnn_perturb_mask=mask.to(cuda_dev))
    assert rel_err(fused, fused_ref) <= BF16_TOL
    fused.float().sum().backward()
    assert mod.gnn_mask_token_embed.grad is not None
    assert torch.isfinite(mod.gnn_mask_token_embed.grad).all()


# ------------------------------------------------------------------ extensions A8 / A9 (parity unpinned by the reference)
from oracle import GATConvRef, GCNConvRef  # noqa: E402


def _loopy_graph(n, e, seed):
    """random graph with duplicate edges and a few self-loops (upstream removes then re-adds them)"""
    ei = synth.uniform_edges(n, e, seed=seed)
    ei[1, : e // 20] = ei[0, : e // 20]
    return ei


@pytest.mark.parametrize("n,e,fi,fo", [(183, 300, 64, 32), (2000, 15000, 300, 64), (3, 0, 8, 8)])
def test_gcn_conv_matches_oracle(cuda_dev, n, e, fi, fo):
    torch.manual_seed(2)
    ei = _loopy_graph(n, e, seed=n)
    ref = GCNConvRef(fi, fo).double()
    with torch.no_grad():
        ref.bias.uniform_(-0.1, 0.1)
    mod = G.GCNConv(fi, fo)
    assert set(mod.state_dict()) == set(ref.state_dict()) == {"lin.weight", "bias"}
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mod = mod.to(cuda_dev)
    x, gout = torch.randn(n, fi), torch.randn(n, fo)
    x64 = x.double().requires_grad_(True)
    y_ref = ref(x64, ei)
    y_ref.backward(gout.double())
    xg = x.to(cuda_dev).requires_grad_(True)
    y = mod(xg, ei.to(cuda_dev))
    y.backward(gout.to(cuda_dev))
    assert rel_err(y, y_ref) <= FP32_TOL
    assert rel_err(xg.grad, x64.grad) <= FP32_TOL
    assert rel_err(mod.lin.weight.grad, ref.lin.weight.grad) <= 2e-5
    assert rel_err(mod.bias.grad, ref.bias.grad) <= 2e-5


@pytest.mark.parametrize("n,e,fi,c,heads,concat", [(183, 300, 64, 16, 4, True), (2449, 9305, 300, 64, 8, True),
                                                   (500, 6000, 32, 8, 3, False), (1, 2, 8, 4, 2, True)])
def test_gat_conv_matches_oracle(cuda_dev, n, e, fi, c, heads, concat):
    """Amazon-ratings-shaped config (BASELINE configs[2], 1/10 scale here) with 8 heads, plus odd
    head counts / widths (scalar-pack path) and the mean-over-heads variant."""
    torch.manual_seed(3)
    ei = _loopy_graph(n, e, seed=n + 7)
    ref = GATConvRef(fi, c, heads=heads, concat=concat).double()
    with torch.no_grad():
        ref.bias.uniform_(-0.1, 0.1)
    mod = G.GATConv(fi, c, heads=heads, concat=concat)
    assert set(mod.state_dict()) == set(ref.state_dict()) == {"lin.weight", "att_src", "att_dst", "bias"}
    mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
    mod = mod.to(cuda_dev)
    out_dim = heads * c if concat else c
    x, gout = torch.randn(n, fi), torch.randn(n, out_dim)
    x64 = x.double().requires_grad_(True)
    y_ref = ref(x64, ei)
    y_ref.backward(gout.double())
    xg = x.to(cuda_dev).requires_grad_(True)
    y = mod(xg, ei.to(cuda_dev))
    y.backward(gout.to(cuda_dev))
    assert rel_err(y, y_ref) <= FP32_TOL
    assert rel_err(xg.grad, x64.grad) <= 2e-5
    # the floor covers gradients that are exactly zero in exact arithmetic (a node whose only in-edge is its
    # self-loop has alpha = 1: d_score = alpha (d_alpha - t) cancels to rounding noise ~1e-8, not to 0.0)
    for name in ("att_src", "att_dst", "bias"):
        assert rel_err(getattr(mod, name).grad, getattr(ref, name).grad, floor=1e-2) <= 2e-5, name
    assert rel_err(mod.lin.weight.grad, ref.lin.weight.grad) <= 2e-5


def test_gat_attention_rows_sum_to_one_and_bf16(cuda_dev):
    n, e, heads, c = 3000, 40000, 8, 32
    ei = synth.rmat_edges(n, e, seed=4).to(cuda_dev)
    g = G.get_loop_graph(ei, n)
    z = torch.ones(n, heads * c, device=cuda_dev)
    a_s = torch.randn(n, heads, device=cuda_dev)
    a_d = torch.randn(n, heads, device=cuda_dev)
    out = G.gat_aggregate(z, a_s, a_d, g)        # softmax-weighted mean of identical rows == the row
    assert float((out - 1).abs().max()) <= 1e-5
    zb = torch.randn(n, heads * c, device=cuda_dev)
    o32 = G.gat_aggregate(zb, a_s, a_d, g)
    o16 = G.gat_aggregate(zb.bfloat16(), a_s, a_d, g)
    assert o16.dtype == torch.bfloat16 and rel_err(o16, o32) <= BF16_TOL


def test_rgcn_conv_arbitrary_relation_ids_and_no_basis(cuda_dev):
    """General RGCNConv use beyond the reference's degree buckets: all five relations populated,
    and the num_bases=None (one weight per relation) variant of upstream."""
    torch.manual_seed(4)
    n, e, fi, fo = 400, 5000, 48, 40
    ei = synth.uniform_edges(n, e, seed=5)
    et = torch.randint(0, 5, (e,), generator=torch.Generator().manual_seed(6))
    for num_bases in (30, None):
        ref = RGCNConvRef(fi, fo, 5, num_bases).double()
        mod = G.RGCNConv(fi, fo, 5, num_bases)
        mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
        mod = mod.to(cuda_dev)
        x = torch.randn(n, fi)
        gout = torch.randn(n, fo)
        x64 = x.double().requires_grad_(True)
        y_ref = ref(x64, ei, et)
        y_ref.backward(gout.double())
        xg = x.to(cuda_dev).requires_grad_(True)
        y = mod(xg, ei.to(cuda_dev), et.to(cuda_dev))
        y.backward(gout.to(cuda_dev))
        assert rel_err(y, y_ref) <= FP32_TOL
        assert rel_err(xg.grad, x64.grad) <= FP32_TOL
        assert rel_err(mod.weight.grad, ref.weight.grad) <= 2e-5
        if num_bases is not None:
            assert rel_err(mod.comp.grad, ref.comp.grad) <= 2e-5 and torch.count_nonzero(mod.comp.grad[4]) > 0


@pytest.mark.parametrize("mode", ["fp32", "autocast", "bf16"])
@pytest.mark.parametrize("root_weight,bias", [(False, True), (True, False), (False, False)])
def test_rgcn_conv_without_root_or_bias(cuda_dev, mode, root_weight, bias):
    """Upstream's root_weight=False / bias=False (not used by the reference, part of the constructor it calls): the
    single-source GEMM, the weight-gradient reduction without the x source and without the ones source."""
    torch.manual_seed(9)
    n, e, fi, fo = 600, 7000, 64, 128
    ei = synth.rmat_edges(n, e, seed=8)
    et = edge_type_bucket_ref(ei, n)
    ref = RGCNConvRef(fi, fo, 5, 30).double()
    with torch.no_grad():
        ref.bias.uniform_(-0.1, 0.1)
        if not root_weight:
            ref.root.zero_()
        if not bias:
            ref.bias.zero_()
    dt = torch.bfloat16 if mode == "bf16" else torch.float32
    mod = G.RGCNConv(fi, fo, 5, 30, root_weight=root_weight, bias=bias, out_dtype=dt if mode == "bf16" else None)
    sd = {k: v.float() for k, v in ref.state_dict().items()}
    if not root_weight:
        sd.pop("root")
    if not bias:
        sd.pop("bias")
    mod.load_state_dict(sd)
    mod = mod.to(cuda_dev)
    x = torch.randn(n, fi).to(dt)
    gout = torch.randn(n, fo).to(dt)
    x64 = x.double().requires_grad_(True)
    y_ref = ref(x64, ei, et)
    y_ref.backward(gout.double())
    xg = x.to(cuda_dev).requires_grad_(True)
    with torch.amp.autocast("cuda", enabled=mode == "autocast"):
        y = mod(xg, ei.to(cuda_dev), et.to(cuda_dev))
    y.backward(gout.to(cuda_dev).to(y.dtype))
    tol = {"fp32": 1e-5, "autocast": 2e-3, "bf16": 2e-2}[mode]
    gtol = {"fp32": 2e-5, "autocast": 5e-3, "bf16": 2e-2}[mode]
    assert rel_err(y, y_ref) <= tol
    assert rel_err(xg.grad, x64.grad) <= tol
    assert rel_err(mod.weight.grad, ref.weight.grad) <= gtol and rel_err(mod.comp.grad, ref.comp.grad) <= gtol
    if root_weight:
        assert rel_err(mod.root.grad, ref.root.grad) <= gtol
    if bias:
        assert rel_err(mod.bias.grad, ref.bias.grad) <= gtol


def test_mask_sampler_on_cuda(cuda_dev):
    """§8f N2 on the device: the exponential-race sampler reproduces torch.multinomial for the same seed, and
    generate_active_node_mask (degrees from the gmlm_degree kernel) selects exactly num_select base nodes with
    positive out-degree."""
    from types import SimpleNamespace
    n, e = 5000, 60000
    ei = synth.rmat_edges(n, e, seed=4).to(cuda_dev)
    w = G.degree(ei[0], n)
    p = w / w.sum()
    torch.manual_seed(11)
    want = torch.multinomial(p, 500, replacement=False)
    torch.manual_seed(11)
    got = G.weighted_sample_without_replacement(p, 500)
    assert torch.equal(got, want)
    train_mask = torch.rand(n, device=cuda_dev) < 0.5
    data = SimpleNamespace(x=torch.zeros(n, 1, device=cuda_dev), edge_index=ei, num_nodes=n, train_mask=train_mask)
    m = G.generate_active_node_mask(data, 0.3)
    k = max(1, int(0.3 * int(train_mask.sum())))
    assert m.dtype == torch.bool and int(m.sum()) == k
    assert bool(train_mask[m].all()) and bool((w[m] > 0).all())


def _gat_reference_csr(graph, z, a_s, a_d, slope, keep=None, p=0.0):
    """fp64 restatement of the GAT aggregation in forward-CSR order (edge k of the CSR), optional keep mask."""
    rowptr, col = graph.fwd.rowptr.cpu().long(), graph.fwd.col.cpu().long()
    n, heads = a_d.shape
    c = z.size(1) // heads
    z64 = z.detach().double().cpu().view(-1, heads, c)
    dst = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    e = torch.nn.functional.leaky_relu(a_s.double().cpu()[col] + a_d.double().cpu()[dst], slope)
    emax = torch.full((n, heads), float("-inf"), dtype=torch.float64).scatter_reduce(
        0, dst.unsqueeze(-1).expand_as(e), e, reduce="amax", include_self=True)
    ex = (e - emax[dst]).exp()
    den = torch.zeros((n, heads), dtype=torch.float64).index_add_(0, dst, ex) + 1e-16
    alpha = ex / den[dst]
    if keep is not None:
        alpha = alpha * keep.cpu().double() / (1.0 - p)
    out = torch.zeros((n, heads, c), dtype=torch.float64).index_add_(0, dst, alpha.unsqueeze(-1) * z64[col])
    return out.view(n, heads * c)


def test_gat_hub_rows_split_softmax_matches_reference_and_is_deterministic(cuda_dev):
    """A star-shaped graph: destination 7 has thousands of in-edges, so the row is cut into chunks whose partial
    (max, normaliser, accumulator) triples are merged by the split-softmax identity; forward AND backward."""
    from gmlm_b200.attn import LoopGraph
    n, e, heads, c = 600, 30000, 4, 16
    g = torch.Generator().manual_seed(5)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.where(torch.rand(e, generator=g) < 0.6, torch.tensor(7), torch.randint(0, n, (e,), generator=g))
    ei = torch.stack([src, dst]).to(cuda_dev)
    graph = LoopGraph.build(ei, n, hub_thresh=64)
    assert graph.fwd.n_hub >= 1 and graph.fwd.n_chunks > graph.fwd.n_hub
    z = torch.randn(n, heads * c, generator=g).to(cuda_dev).requires_grad_(True)
    a_s = (torch.randn(n, heads, generator=g) * 3).to(cuda_dev).requires_grad_(True)     # wide score range
    a_d = (torch.randn(n, heads, generator=g) * 3).to(cuda_dev).requires_grad_(True)
    gout = torch.randn(n, heads * c, generator=g).to(cuda_dev)
    out = G.gat_aggregate(z, a_s, a_d, graph)
    out.backward(gout)
    z64 = z.detach().double().cpu().requires_grad_(True)
    as64, ad64 = a_s.detach().double().cpu().requires_grad_(True), a_d.detach().double().cpu().requires_grad_(True)
    # autograd through the fp64 restatement (edge list = the CSR)
    rowptr, col = graph.fwd.rowptr.cpu().long(), graph.fwd.col.cpu().long()
    dstv = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    sc = torch.nn.functional.leaky_relu(as64[col] + ad64[dstv], 0.2)
    emax = torch.full((n, heads), float("-inf"), dtype=torch.float64).scatter_reduce(
        0, dstv.unsqueeze(-1).expand_as(sc), sc.detach(), reduce="amax", include_self=True)
    ex = (sc - emax[dstv]).exp()
    alpha = ex / (torch.zeros((n, heads), dtype=torch.float64).index_add_(0, dstv, ex) + 1e-16)[dstv]
    ref = torch.zeros((n, heads, c), dtype=torch.float64).index_add_(0, dstv, alpha.unsqueeze(-1) * z64.view(n, heads, c)[col])
    ref = ref.view(n, heads * c)
    ref.backward(gout.double().cpu())
    assert rel_err(out, ref) <= FP32_TOL
    assert rel_err(z.grad, z64.grad) <= 2e-5
    assert rel_err(a_s.grad, as64.grad) <= 2e-5 and rel_err(a_d.grad, ad64.grad) <= 2e-5
    out2 = G.gat_aggregate(z.detach(), a_s.detach(), a_d.detach(), graph)
    assert torch.equal(out.detach(), out2)                       # no atomics: bit-identical run to run


def test_gat_attention_dropout_uses_the_hash_mask(cuda_dev):
    """Training-mode attention dropout (upstream: F.dropout on alpha).  The keep mask is a counter-based hash of
    (seed, CSR position, head); with that mask the output equals the reference exactly, the kept fraction is
    1 - p, eval mode ignores dropout and a fixed seed reproduces the call."""
    from gmlm_b200.attn import gat_dropout_mask
    n, e, heads, c, p = 1500, 20000, 8, 16, 0.4
    ei = synth.rmat_edges(n, e, seed=9).to(cuda_dev)
    graph = G.get_loop_graph(ei, n)
    gen = torch.Generator().manual_seed(2)
    z = torch.randn(n, heads * c, generator=gen).to(cuda_dev)
    a_s, a_d = torch.randn(n, heads, generator=gen).to(cuda_dev), torch.randn(n, heads, generator=gen).to(cuda_dev)
    seed = 123456789
    keep = gat_dropout_mask(seed, graph.fwd.nnz, heads, p, cuda_dev)
    assert abs(float(keep.float().mean()) - (1 - p)) < 0.01
    out = G.gat_aggregate(z, a_s, a_d, graph, 0.2, p, seed)
    ref = _gat_reference_csr(graph, z, a_s, a_d, 0.2, keep=keep, p=p)
    assert rel_err(out, ref) <= FP32_TOL
    assert torch.equal(out, G.gat_aggregate(z, a_s, a_d, graph, 0.2, p, seed))
    assert not torch.equal(out, G.gat_aggregate(z, a_s, a_d, graph, 0.2, p, seed + 1))
    # gradient through the dropped attention: finite differences are noisy, so compare with autograd of the reference
    zg = z.clone().requires_grad_(True)
    og = G.gat_aggregate(zg, a_s, a_d, graph, 0.2, p, seed)
    og.sum().backward()
    rowptr, col = graph.fwd.rowptr.cpu().long(), graph.fwd.col.cpu().long()
    # d(sum out)/dz[j,h,:] = sum over out-edges of j of alpha_eff  (broadcast over the head's channels)
    dstv = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    sc = torch.nn.functional.leaky_relu(a_s.double().cpu()[col] + a_d.double().cpu()[dstv], 0.2)
    emax = torch.full((n, heads), float("-inf"), dtype=torch.float64).scatter_reduce(
        0, dstv.unsqueeze(-1).expand_as(sc), sc, reduce="amax", include_self=True)
    ex = (sc - emax[dstv]).exp()
    alpha = ex / (torch.zeros((n, heads), dtype=torch.float64).index_add_(0, dstv, ex) + 1e-16)[dstv]
    alpha = alpha * keep.cpu().double() / (1 - p)
    want = torch.zeros((n, heads), dtype=torch.float64).index_add_(0, col, alpha)
    assert rel_err(zg.grad.view(n, heads, c)[:, :, 0], want) <= 2e-5
    mod = G.GATConv(32, c, heads=heads, dropout=p).to(cuda_dev)
    x = torch.randn(n, 32, device=cuda_dev)
    mod.eval()
    assert torch.equal(mod(x, ei), mod(x, ei))                   # eval: dropout off
    mod.train()
    torch.manual_seed(0)
    y1 = mod(x, ei)
    torch.manual_seed(0)
    y2 = mod(x, ei)
    assert torch.equal(y1, y2) and not torch.equal(y1, mod(x, ei))



def test_graphed_encoder_step_equals_eager(cuda_dev):
    """GraphedEncoderStep (forward + backward captured once as a CUDA graph) replays to the same
    output and parameter gradients as the eager step, for new inputs copied into the captured buffers."""
    import gmlm_b200 as G
    from gmlm_b200 import synth
    torch.manual_seed(3)
    n, e, f, h = 3000, 9000, 300, 64
    ei = synth.rmat_edges(n, e, seed=11).to(cuda_dev)
    enc = G.GraphEncoder(f, h, 96, dropout_rate=0.0).to(cuda_dev)
    xs = [torch.randn(n, f, device=cuda_dev) for _ in range(3)]
    gy = torch.randn(n, 96, device=cuda_dev)
    step = G.GraphedEncoderStep(enc, xs[0], ei, autocast=True, x_requires_grad=True)
    replayed = []
    for x in xs[1:]:                                      # replays first: the captured .grad tensors stay attached
        y = step(x, gy).clone()
        grads = {k: p.grad.clone() for k, p in enc.named_parameters() if p.grad is not None}
        replayed.append((y, step.input_grad.clone(), grads))
    assert len(replayed[0][2]) >= 40
    for x, (y, gx, grads) in zip(xs[1:], replayed):       # then the same steps eagerly
        enc.zero_grad(set_to_none=True)
        xe = x.clone().requires_grad_(True)
        with torch.amp.autocast("cuda"):
            ye = enc.get_graph_embeddings(xe, ei)
        ye.backward(gy)
        assert torch.equal(y, ye)
        assert torch.equal(gx, xe.grad)
        for k, p in enc.named_parameters():
            if p.grad is not None:
                assert torch.equal(grads[k], p.grad), k
