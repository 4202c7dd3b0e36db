"""tcgen05 GEMM (the dense transform of [PyG] RGCNConv.forward, main.py:272)
against a plain PyTorch fp32 reference of the same op on the same bf16 inputs."""
import pytest
import torch

from gmlm_b200.ops import gemm_nt, gemm_tn

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ref(a1, b, bias, a2):
    a = a1.float() if a2 is None else torch.cat([a1.float(), a2.float()], dim=1)
    out = a @ b.float().t()
    return out + bias if bias is not None else out


@pytest.mark.parametrize("m,n,k1,k2", [(128, 64, 64, 0), (300, 64, 128, 0), (389, 256, 64, 0), (1000, 128, 256, 64),
                                       (4097, 64, 1024, 256), (77, 32, 192, 0), (513, 512, 128, 128), (5, 96, 64, 0),
                                       (40000, 320, 256, 0), (30011, 1280, 64, 0), (70000, 64, 64, 0), (9000, 160, 128, 0),
                                       (20000, 640, 64, 64)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("use_bias", [False, True])
def test_gemm_nt_matches_fp32_reference(cuda_dev, m, n, k1, k2, out_dtype, use_bias):
    g = torch.Generator().manual_seed(m + n + k1)
    a1 = torch.randn(m, k1, generator=g).bfloat16().to(cuda_dev)
    a2 = torch.randn(m, k2, generator=g).bfloat16().to(cuda_dev) if k2 else None
    b = (torch.randn(n, k1 + k2, generator=g) / (k1 + k2) ** 0.5).bfloat16().to(cuda_dev)
    bias = torch.randn(n, generator=g).to(cuda_dev) if use_bias else None
    got = gemm_nt(a1, b, bias=bias, a2=a2, out_dtype=out_dtype)
    ref = _ref(a1, b, bias, a2)
    assert got.shape == (m, n) and got.dtype == out_dtype
    tol = 1e-5 if out_dtype == torch.float32 else 1e-2     # fp32 accumulate; bf16 only rounds the store
    assert rel_err(got, ref) <= tol


def test_gemm_nt_split_outputs_and_strided_operands(cuda_dev):
    """backward use: [dH | dx] = g @ [W ; root]^T written to two contiguous tensors; A given as a
    column slice (leading dimension > K)."""
    g = torch.Generator().manual_seed(1)
    m, n, k = 777, 1024 + 256, 64
    big = torch.randn(m, 192, generator=g).bfloat16().to(cuda_dev)
    a = big[:, 64:128]                                   # lda = 192
    b = (torch.randn(n, k, generator=g) / 8).bfloat16().to(cuda_dev)
    c1, c2 = gemm_nt(a, b, split=1024)
    ref = a.float() @ b.float().t()
    assert c1.shape == (m, 1024) and c2.shape == (m, 256) and c1.is_contiguous() and c2.is_contiguous()
    assert rel_err(c1, ref[:, :1024]) <= 1e-2 and rel_err(c2, ref[:, 1024:]) <= 1e-2


def test_gemm_nt_rejects_bad_operands(cuda_dev):
    from gmlm_b200 import GmlmError
    a = torch.zeros(64, 100, dtype=torch.bfloat16, device=cuda_dev)
    b = torch.zeros(64, 100, dtype=torch.bfloat16, device=cuda_dev)
    with pytest.raises(GmlmError):
        gemm_nt(a, b.half())                                              # mixed operand types
    with pytest.raises(GmlmError):
        gemm_nt(a.float(), b.float(), out_dtype=torch.bfloat16)           # fp32 operands give fp32
    with pytest.raises(GmlmError):
        gemm_nt(a, b[:, :96])                                             # K mismatch
    with pytest.raises(GmlmError):
        gemm_nt(a, b, addend=torch.zeros(64, 63, device=cuda_dev))        # addend shape


@pytest.mark.parametrize("op", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,ks", [(300, 512, (1200, 300)), (1000, 100, (100,)), (257, 1500, (512,)), (129, 36, (70, 30, 20)),
                                    (5000, 768, (64, 128, 256, 512)), (64, 8, (8,)), (999, 200, (72, 200))])
def test_gemm_nt_ragged_shapes_and_sources(cuda_dev, m, n, ks, op):
    """Any K, N: partial k-blocks of every source are zero-filled by the TMA, partial output tiles clipped by the
    TMA store (C2/C3: Fi = 300, S*Fi = 1200); up to four A sources (the MultiScaleFusion inputs, never concatenated)."""
    g = torch.Generator().manual_seed(m + n)
    srcs = [torch.randn(m, k, generator=g).to(op).to(cuda_dev) for k in ks]
    b = (torch.randn(n, sum(ks), generator=g) / sum(ks) ** 0.5).to(op).to(cuda_dev)
    bias = torch.randn(n, generator=g).to(cuda_dev)
    got = gemm_nt(srcs, b, bias=bias, out_dtype=torch.float32)
    ref = torch.cat([t.float() for t in srcs], dim=1) @ b.float().t() + bias
    assert got.shape == (m, n) and got.dtype == torch.float32
    assert torch.isfinite(got).all()
    assert rel_err(got, ref) <= 1e-5


@pytest.mark.parametrize("m,n,ks", [(300, 512, (1200, 300)), (4097, 64, (1024, 256)), (1000, 100, (100,)), (50000, 128, (256, 64)),
                                    (777, 36, (72, 20)), (2000, 4096, (512,)), (64, 8, (4,)), (22662, 512, (1200, 300)),
                                    (9000, 320, (1703,))])
@pytest.mark.parametrize("use_bias", [False, True])
def test_gemm_nt_fp32_operands_3xtf32(cuda_dev, m, n, ks, use_bias):
    """fp32 operands (the reference's eval mode, main.py:603) on the tf32 tensor cores as 3xTF32 -- hi/lo split in
    shared memory, three MMAs per k-step -- against an fp64 product: the 1e-5 gate plain TF32 (2^-11) cannot meet."""
    g = torch.Generator().manual_seed(m + n)
    srcs = [torch.randn(m, k, generator=g).to(cuda_dev) for k in ks]
    b = (torch.randn(n, sum(ks), generator=g) / sum(ks) ** 0.5).to(cuda_dev)
    bias = torch.randn(n, generator=g).to(cuda_dev) if use_bias else None
    got = gemm_nt(srcs, b, bias=bias)
    ref = torch.cat([t.double() for t in srcs], dim=1) @ b.double().t()
    if bias is not None:
        ref = ref + bias.double()
    assert got.shape == (m, n) and got.dtype == torch.float32
    assert rel_err(got, ref) <= 2e-6
    # per element, against |ref| + rms(ref): a dot product's rounding error scales with sum |a_k b_k| (~ rms of the
    # outputs), not with a result that cancelled to nearly zero
    from conftest import elementwise_err
    assert elementwise_err(got, ref, atol_frac=1.0) <= 1e-5


def test_gemm_nt_fp32_split_and_addend(cuda_dev):
    g = torch.Generator().manual_seed(9)
    m, k, n = 3000, 512, 1500
    a = torch.randn(m, k, generator=g).to(cuda_dev)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda_dev)
    c1, c2 = gemm_nt(a, b, split=1200)
    ref = a.double() @ b.double().t()
    assert rel_err(c1, ref[:, :1200]) <= 2e-6 and rel_err(c2, ref[:, 1200:]) <= 2e-6
    add = torch.randn(m, n, generator=g).to(cuda_dev)
    got = gemm_nt(a, b, addend=add)
    assert rel_err(got, ref + add.double()) <= 2e-6


@pytest.mark.parametrize("n,split", [(1500, 1200), (1280, 1024), (330, 30), (96, 64)])
def test_gemm_nt_split_at_any_column(cuda_dev, n, split):
    g = torch.Generator().manual_seed(n)
    m, k = 1111, 512
    a = torch.randn(m, k, generator=g).half().to(cuda_dev)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(cuda_dev)
    c1, c2 = gemm_nt(a, b, out_dtype=torch.float32, split=split)
    ref = a.float() @ b.float().t()
    assert c1.shape == (m, split) and c2.shape == (m, n - split)
    assert rel_err(c1, ref[:, :split]) <= 1e-5 and rel_err(c2, ref[:, split:]) <= 1e-5


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("widths", [(64, 128, 256, 512), (128, 64, 40), (256, 256, 256, 192), (64, 64, 64)])
def test_gemm_nt_up_to_four_routed_outputs(cuda_dev, widths, out_dtype):
    """More than two outputs: tiles straddle the cuts (multiples of 64 columns), every epilogue chunk is routed to its
    output's tensor map -- the input gradients of MultiScaleFusion land contiguous per layer."""
    g = torch.Generator().manual_seed(sum(widths))
    m, k, n = 3001, 768, sum(widths)
    a = torch.randn(m, k, generator=g).bfloat16().to(cuda_dev)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).bfloat16().to(cuda_dev)
    outs = gemm_nt(a, b, out_dtype=out_dtype, split=list(widths))
    ref = a.float() @ b.float().t()
    assert len(outs) == len(widths)
    c0 = 0
    for o, w in zip(outs, widths):
        assert o.shape == (m, w) and o.dtype == out_dtype
        assert o.is_contiguous() or w % (16 // o.element_size()) != 0
        assert rel_err(o, ref[:, c0:c0 + w]) <= (1e-5 if out_dtype == torch.float32 else 1e-2)
        c0 += w


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("m,n,k", [(1000, 512, 300), (4099, 64, 256), (70, 128, 64), (2500, 1024, 512)])
def test_gemm_nt_addend_in_the_epilogue(cuda_dev, m, n, k, out_dtype):
    """C = A B^T + bias + addend: the residual add of main.py:281-282 folded into the projection."""
    g = torch.Generator().manual_seed(k)
    a = torch.randn(m, k, generator=g).bfloat16().to(cuda_dev)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).bfloat16().to(cuda_dev)
    bias = torch.randn(n, generator=g).to(cuda_dev)
    add = torch.randn(m, n, generator=g).to(out_dtype).to(cuda_dev)
    got = gemm_nt(a, b, bias=bias, out_dtype=out_dtype, addend=add)
    ref = a.float() @ b.float().t() + bias + add.float()
    assert rel_err(got, ref) <= (1e-5 if out_dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("op", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,ks,n", [(1000, (64,), 64), (100000, (1024, 256), 64), (5000, (1200, 300), 512),
                                    (70000, (256,), 320), (777, (100, 36), 200), (30000, (64, 128, 256, 512), 768),
                                    (200, (8,), 8), (37, (130,), 70), (22662, (2560,), 1024), (64, (128,), 256),
                                    (300000, (320,), 64)])
def test_gemm_tn_weight_gradient_reduction(cuda_dev, m, ks, n, op):
    """D = [A_0 | ..]^T G (dW = H^T g, droot = x^T g in one launch): MN-major tcgen05 operands, split over the rows,
    against an fp64 product of the same 16-bit inputs; deterministic."""
    g = torch.Generator().manual_seed(m + n)
    srcs = [torch.randn(m, k, generator=g).to(op).to(cuda_dev) for k in ks]
    gg = torch.randn(m, n, generator=g).to(op).to(cuda_dev)
    got = gemm_tn(srcs, gg)
    ref = torch.cat([t.double() for t in srcs], dim=1).t() @ gg.double()
    assert got.shape == (sum(ks), n) and got.dtype == torch.float32
    assert rel_err(got, ref) <= 1e-4          # an fp32 tensor-core accumulation up to 22 662 rows deep (one split)
    assert torch.equal(gemm_tn(srcs, gg), got)


def test_gemm_tn_strided_operands(cuda_dev):
    g = torch.Generator().manual_seed(3)
    big = torch.randn(5000, 512, generator=g).bfloat16().to(cuda_dev)
    a, gg = big[:, 64:192], big[:, 256:320]                  # column slices: leading dimension 512
    got = gemm_tn([a], gg)
    assert rel_err(got, a.double().t() @ gg.double()) <= 2e-5


def test_rgcn_conv_bf16_tcgen05_path_matches_cublas_path_and_oracle(cuda_dev):
    """bf16 pipeline (BASELINE configs[3]/[4]): the layer with the tcgen05 transform vs the same layer
    with cuBLAS matmuls vs the fp64 oracle, forward and all gradients."""
    import gmlm_b200 as G
    from gmlm_b200 import synth
    from oracle import RGCNConvRef, edge_type_bucket_ref
    n, e, fi, fo = 3000, 40000, 128, 64
    torch.manual_seed(0)
    ei = synth.rmat_edges(n, e, seed=2)
    et = edge_type_bucket_ref(ei, n)
    ref = RGCNConvRef(fi, fo, 5, 30).double()
    with torch.no_grad():
        ref.bias.uniform_(-0.1, 0.1)
    x = torch.randn(n, fi).bfloat16()
    gout = torch.randn(n, fo).bfloat16()
    x64 = x.double().requires_grad_(True)
    y_ref = ref(x64, ei, et)
    y_ref.backward(gout.double())
    results = {}
    for use_tc in (True, False):
        mod = G.RGCNConv(fi, fo, 5, 30, out_dtype=torch.bfloat16)
        mod.load_state_dict({k: v.float() for k, v in ref.state_dict().items()})
        mod = mod.to(cuda_dev)
        mod.use_tcgen05 = use_tc
        xg = x.to(cuda_dev).requires_grad_(True)
        y = mod(xg, ei.to(cuda_dev), et.to(cuda_dev))
        y.backward(gout.to(cuda_dev))
        assert y.dtype == torch.bfloat16
        assert rel_err(y, y_ref) <= 2e-2
        assert rel_err(xg.grad, x64.grad) <= 2e-2
        for name in ("weight", "comp", "root", "bias"):
            assert rel_err(getattr(mod, name).grad, getattr(ref, name).grad) <= 2e-2, (use_tc, name)
        results[use_tc] = (y.float(), xg.grad.float())
    assert rel_err(results[True][0], results[False][0]) <= 2e-2
