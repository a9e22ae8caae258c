"""Tensor-core GEMM / implicit conv / attention kernels against a plain PyTorch fp64 reference (through the C ABI)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

DT = {0: torch.float32, 1: torch.bfloat16}
# max|err| relative to max|ref|: TF32 carries a 10-bit mantissa (2^-11 per operand), bf16 a 7-bit one
TOL = {0: 2e-3, 1: 1.2e-2}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _rna_tf32(x):
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _rel(got, ref):
    return ((got.double() - ref).abs().max() / ref.abs().max()).item()


@pytest.mark.parametrize("dt", [0, 1])
@pytest.mark.parametrize("M,N,K,relu,res,res_mod", [
    (128, 64, 64, 0, 0, 0), (1, 64, 64, 0, 0, 0), (300, 256, 512, 1, 1, 0), (6272, 64, 256, 1, 0, 0),
    (2352, 768, 256, 0, 1, 784), (40, 256, 2048, 0, 1, 0), (12544, 64, 192, 1, 0, 0), (1000, 2048, 256, 1, 0, 0),
    (784 * 16, 128, 1152, 1, 0, 0),
    # CTA-pair (cta_group::2) path: N % 256 == 0, deep K; odd number of 128-row tiles, ragged tail, wrapped residual
    (12837, 512, 1024, 1, 1, 0), (12288, 256, 2048, 0, 1, 784), (784 * 9, 256, 512, 1, 0, 0),
    # A-resident CTA-pair path (K = 256, N % 256 == 0, >= 74 pair tiles): three column blocks with a wrapped addend
    # and a ragged, odd tile count; one column block with a plain residual + ReLU; many tiles per pair
    (19001, 768, 256, 0, 1, 784), (18944, 256, 256, 1, 1, 0), (50176, 768, 256, 0, 1, 784), (50176, 256, 256, 0, 1, 0)])
def test_gemm(lib, cuda_dev, dt, M, N, K, relu, res, res_mod):
    torch.manual_seed(M + N + K)
    tdt = DT[dt]
    A = torch.randn(M, K, device=cuda_dev).to(tdt)
    W = (torch.randn(N, K, device=cuda_dev) / K ** 0.5).to(tdt)
    scale = torch.rand(N, device=cuda_dev) + 0.5
    bias = torch.randn(N, device=cuda_dev)
    rows = res_mod if res_mod else M
    R = torch.randn(rows, N, device=cuda_dev).to(tdt) if res else None
    out = torch.full((M, N), float("nan"), device=cuda_dev).to(tdt)
    assert lib.spe_debug_gemm(dt, _p(A), _p(W), M, N, K, _p(scale), _p(bias), _p(R), res_mod, relu, _p(out), None) == 0
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() * scale.double() + bias.double()
    if res:
        r = R.double()
        ref = ref + (r.repeat(M // rows + 1, 1)[:M] if res_mod else r)
    if relu:
        ref = ref.clamp_min(0)
    assert not torch.isnan(out.float()).any()
    assert _rel(out, ref) < TOL[dt]


@pytest.mark.parametrize("dt", [0, 1])
@pytest.mark.parametrize("NB,H,Cin,Cout", [(2, 28, 64, 64), (3, 56, 64, 64), (2, 14, 256, 256), (1, 28, 1024, 256),
                                            (1, 7, 64, 128), (5, 32, 128, 128), (16, 28, 256, 256), (15, 25, 512, 256),
                                            # tap-reuse kernel (Cout <= 128): odd extents, widest grid, many tiles per CTA
                                            (3, 28, 128, 128), (2, 30, 64, 128), (1, 126, 64, 64), (40, 56, 64, 64),
                                            # im2col-TMA pair path (tiles of 128 consecutive output pixels that wrap over
                                            # rows and images): ragged last tile, odd tile count, two column blocks
                                            (20, 30, 256, 256), (33, 28, 256, 512), (64, 28, 512, 512),
                                            (2, 13, 128, 64)])
def test_implicit_conv3x3(lib, cuda_dev, dt, NB, H, Cin, Cout):
    """TMA out-of-bounds zero fill == the convolution's zero padding; ragged last row-tile (H % hrows != 0)."""
    torch.manual_seed(H * Cin)
    tdt = DT[dt]
    x = torch.randn(NB, H, H, Cin, device=cuda_dev).to(tdt)
    w = (torch.randn(Cout, Cin, 3, 3, device=cuda_dev) / (9 * Cin) ** 0.5).to(tdt)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    bias = torch.randn(Cout, device=cuda_dev)
    out = torch.full((NB, H, H, Cout), float("nan"), device=cuda_dev).to(tdt)
    assert lib.spe_debug_conv(dt, _p(x), _p(wk), NB, H, H, Cin, Cout, 3, 3, 1, 1, None, _p(bias), 1, _p(out), None) == 0
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), bias.double(), padding=1)
    ref = ref.clamp_min(0).permute(0, 2, 3, 1)
    assert not torch.isnan(out.float()).any()
    assert _rel(out, ref) < TOL[dt]


@pytest.mark.parametrize("dt", [0, 1])
@pytest.mark.parametrize("NB,H,Cin,Cout,R", [(2, 56, 128, 128, 3), (3, 28, 256, 256, 3), (2, 56, 256, 512, 1),
                                              (2, 28, 512, 1024, 1), (1, 64, 128, 128, 3), (16, 56, 256, 256, 3),
                                              (13, 50, 512, 1024, 1)])
def test_implicit_conv_stride2(lib, cuda_dev, dt, NB, H, Cin, Cout, R):
    """Stride-2 convolutions (layerN.0.conv2, downsample) through the TMA traversal stride: no im2col buffer."""
    torch.manual_seed(H * Cin + R)
    tdt = DT[dt]
    x = torch.randn(NB, H, H, Cin, device=cuda_dev).to(tdt)
    w = (torch.randn(Cout, Cin, R, R, device=cuda_dev) / (R * R * Cin) ** 0.5).to(tdt)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, R * R * Cin).contiguous()
    Ho = (H + 2 * (R // 2) - R) // 2 + 1
    out = torch.full((NB, Ho, Ho, Cout), float("nan"), device=cuda_dev).to(tdt)
    assert lib.spe_debug_conv(dt, _p(x), _p(wk), NB, H, H, Cin, Cout, R, R, R // 2, 2, None, None, 0, _p(out), None) == 0
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), stride=2, padding=R // 2)
    ref = ref.permute(0, 2, 3, 1)
    assert not torch.isnan(out.float()).any()
    assert _rel(out, ref) < TOL[dt]


@pytest.mark.parametrize("dt", [0, 1])
@pytest.mark.parametrize("B,Lq,Lk", [(2, 784, 784), (3, 40, 40), (2, 40, 784), (1, 100, 1024), (1, 1, 5), (3, 1024, 1024),
                                     (2, 200, 300)])
def test_attention(lib, cuda_dev, dt, B, Lq, Lk):
    torch.manual_seed(Lq + Lk)
    tdt = DT[dt]
    q = torch.randn(B, Lq, 256, device=cuda_dev).to(tdt)
    k = torch.randn(B, Lk, 256, device=cuda_dev).to(tdt)
    v = torch.randn(B, Lk, 256, device=cuda_dev).to(tdt)
    if dt == 0:   # in the pipeline Q/K/V are GEMM outputs already rounded to TF32 (what the tensor core consumes)
        q, k, v = _rna_tf32(q), _rna_tf32(k), _rna_tf32(v)
    out = torch.full((B, Lq, 256), float("nan"), device=cuda_dev).to(tdt)
    assert lib.spe_debug_attention(dt, _p(q), _p(k), _p(v), _p(out), B, 8, Lq, Lk, 256, 256, 256, 256, None) == 0
    torch.cuda.synchronize()
    qh = q.double().view(B, Lq, 8, 32).transpose(1, 2)
    kh = k.double().view(B, Lk, 8, 32).transpose(1, 2)
    vh = v.double().view(B, Lk, 8, 32).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) / 32 ** 0.5, -1) @ vh).transpose(1, 2).reshape(B, Lq, 256)
    assert not torch.isnan(out.float()).any()
    assert _rel(out, ref) < (2e-3 if dt == 0 else 8e-3)


@pytest.mark.parametrize("late_boost", [0.0, 40.0, 4000.0])
def test_attention_fixed_reference_softmax_and_its_fallback(lib, cuda_dev, late_boost):
    """The tcgen05 attention kernel takes its softmax reference from the first chunk of 112 keys and keeps it for the
    row; scores far above it later in the row (here: keys 500.. boosted so that q.k grows by `late_boost`) make
    probabilities > 1 -- still exact -- and, beyond ~350 logits, overflow a row sum, which the kernel detects and
    answers by repeating the tile with the classical online softmax.  All three regimes must give the same softmax."""
    torch.manual_seed(7)
    B, L = 2, 784
    q = torch.randn(B, L, 256, device=cuda_dev)
    k = torch.randn(B, L, 256, device=cuda_dev)
    v = torch.randn(B, L, 256, device=cuda_dev)
    if late_boost:
        # add a multiple of q's own direction to the late keys of head 0 and head 5 for the first 200 queries' mean
        # direction: every query's score on those keys rises by roughly late_boost * |component|
        d = torch.nn.functional.normalize(q[:, :, :32].mean(1, keepdim=True), dim=-1)
        k[:, 500:, :32] += late_boost * d * 32 ** 0.5 / q[:, :, :32].norm(dim=-1).mean()
    q, k, v = _rna_tf32(q), _rna_tf32(k), _rna_tf32(v)
    out = torch.full((B, L, 256), float("nan"), device=cuda_dev)
    assert lib.spe_debug_attention(0, _p(q), _p(k), _p(v), _p(out), B, 8, L, L, 256, 256, 256, 256, None) == 0
    torch.cuda.synchronize()
    qh = q.double().view(B, L, 8, 32).transpose(1, 2)
    kh = k.double().view(B, L, 8, 32).transpose(1, 2)
    vh = v.double().view(B, L, 8, 32).transpose(1, 2)
    ref = (torch.softmax(qh @ kh.transpose(-1, -2) / 32 ** 0.5, -1) @ vh).transpose(1, 2).reshape(B, L, 256)
    assert not torch.isnan(out).any() and not torch.isinf(out).any()
    assert _rel(out, ref) < 2e-3


@pytest.mark.parametrize("M,N,K,relu,res_mod", [(40, 256, 256, 0, 0), (2560, 768, 256, 0, 40), (2560, 2048, 256, 1, 0),
                                                (300, 256, 2048, 0, 0), (784 * 8, 2048, 256, 0, 784)])
def test_gemm_3xtf32_reaches_fp32_accuracy(lib, cuda_dev, M, N, K, relu, res_mod):
    """Error-compensated 3xTF32 (decoder + heads): ~1e-6 relative, vs ~5e-4 for plain TF32."""
    torch.manual_seed(M + N)
    A = torch.randn(M, K, device=cuda_dev)
    W = torch.randn(N, K, device=cuda_dev) / K ** 0.5
    hi = _rna_tf32(W)
    W2 = torch.cat([hi, _rna_tf32(W - hi)], 1).contiguous()
    bias = torch.randn(N, device=cuda_dev)
    rows = res_mod if res_mod else M
    R = torch.randn(rows, N, device=cuda_dev)
    out = torch.full((M, N), float("nan"), device=cuda_dev)
    assert lib.spe_debug_gemm(2, _p(A), _p(W2), M, N, K, None, _p(bias), _p(R), res_mod, relu, _p(out), None) == 0
    torch.cuda.synchronize()
    r = R.double()
    ref = A.double() @ W.double().t() + bias.double() + (r.repeat(M // rows + 1, 1)[:M] if res_mod else r)
    if relu:
        ref = ref.clamp_min(0)
    assert _rel(out, ref) < 2e-5      # fp32 accumulation over up to 3 x 2048 products; plain TF32 sits at ~5e-4


@pytest.mark.parametrize("M,hidden,mode", [(128, 256, 0), (784, 2048, 0), (3 * 784, 2048, 2), (1000, 512, 1),
                                           (50176, 2048, 0), (1000, 2048, 3), (50176, 2048, 3)])
def test_fused_ffn_layernorm(lib, cuda_dev, M, hidden, mode):
    """LayerNorm(X + relu(X W1^T + b1) W2^T + b2) in one kernel (hidden tile in tensor memory) vs fp64 on the same
    TF32-rounded operands; ragged last tile, several tiles per CTA, in-place output, 3xTF32 operand form."""
    torch.manual_seed(M + hidden)
    X = _rna_tf32(torch.randn(M, 256, device=cuda_dev))
    W1 = _rna_tf32(torch.randn(hidden, 256, device=cuda_dev) / 16)
    W2 = _rna_tf32(torch.randn(256, hidden, device=cuda_dev) / hidden ** 0.5)
    b1, b2 = torch.randn(hidden, device=cuda_dev) * 0.5, torch.randn(256, device=cuda_dev)
    g, be = torch.rand(256, device=cuda_dev) + 0.5, torch.randn(256, device=cuda_dev)
    h = _rna_tf32(torch.relu(X.double() @ W1.double().t() + b1.double()).float()).double()
    v = X.double() + h @ W2.double().t() + b2.double()
    ref = (v - v.mean(1, keepdim=True)) / torch.sqrt(v.var(1, unbiased=False, keepdim=True) + 1e-5) * g.double() + be.double()
    if mode == 2:
        out = torch.full((M, 768), float("nan"), device=cuda_dev)
    elif mode == 3:                                       # bf16 output (bf16 storage with the fp32 / TF32 feed-forward)
        out = torch.full((M, 256), float("nan"), device=cuda_dev, dtype=torch.bfloat16)
    else:
        out = X.clone()                                   # in place, like the encoder schedule
    src = X if mode in (2, 3) else out
    assert lib.spe_debug_ffn(_p(src), M, _p(W1), _p(b1), _p(W2), _p(b2), _p(g), _p(be), hidden, mode, _p(out), None) == 0
    torch.cuda.synchronize()
    assert not torch.isnan(out.float()).any()
    if mode == 3:
        got = out.double()
        assert ((got - ref).abs() <= 1e-3 + 2.0 ** -8 * ref.abs()).all()      # bf16 rounding of the result
        return
    if mode == 2:
        hi, lo, hi2 = out[:, :256], out[:, 256:512], out[:, 512:]
        assert torch.equal(hi, hi2) and torch.equal(hi, _rna_tf32(hi))
        got = hi.double() + lo.double()
    else:
        got = out.double()
        if mode == 0:
            assert torch.equal(out, _rna_tf32(out))
    # H is rounded to TF32 (rel 2^-11) once, everything else is fp32; mode 0 also rounds the result itself
    bound = 1e-3 + (6e-4 if mode == 0 else 0.0) * ref.abs()
    assert ((got - ref).abs() <= bound).all(), ((got - ref).abs() - bound).max().item()


def test_gemm_rejects_bad_shapes(lib, cuda_dev):
    a = torch.zeros(8, 48, device=cuda_dev)
    assert lib.spe_debug_gemm(0, _p(a), _p(a), 8, 8, 48, None, None, None, 0, 0, _p(a), None) != 0
    assert b"gemm" in lib.spe_global_last_error()


@pytest.mark.parametrize("dt", [0, 1])
@pytest.mark.parametrize("NB,H,K1,K2,N,stride", [(3, 56, 64, 64, 256, 1), (5, 56, 128, 256, 512, 2), (7, 28, 256, 512, 1024, 2),
                                                   (2, 13, 64, 128, 128, 2)])
def test_gemm_with_a_second_operand_source(lib, cuda_dev, dt, NB, H, K1, K2, N, stride):
    """The first block of a ResNet layer: conv3 and the (strided) 1x1 downsample branch as ONE GEMM over the concatenated
    operand [t | x sampled at the stride] (GemmDesc::A2, model.cu Bottleneck::c3d).  Reference: the two matrix products
    in fp64 on the same (TF32- / bf16-rounded) operands."""
    torch.manual_seed(NB * 100 + H)
    Ho = (H - 1) // stride + 1
    M = NB * Ho * Ho
    tdt = torch.float32 if dt == 0 else torch.bfloat16
    rnd = (lambda t: _rna_tf32(t)) if dt == 0 else (lambda t: t.to(torch.bfloat16))
    t = rnd(torch.randn(M, K1, device=cuda_dev))
    x = rnd(torch.randn(NB, H, H, K2, device=cuda_dev))
    w = rnd(torch.randn(N, K1 + K2, device=cuda_dev) * 0.05)
    bias = torch.randn(N, device=cuda_dev)
    out = torch.full((M, N), float("nan"), device=cuda_dev, dtype=tdt)
    assert lib.spe_debug_gemm2(dt, _p(t), K1, _p(x), K2, stride, NB, H, H, _p(w), M, N, _p(bias), 1, _p(out), None) == 0, \
        lib.spe_global_last_error()
    torch.cuda.synchronize()
    xs = x[:, ::stride, ::stride, :].reshape(M, K2)
    ref = torch.relu(t.double() @ w[:, :K1].double().T + xs.double() @ w[:, K1:].double().T + bias.double())
    assert not torch.isnan(out.float()).any()
    assert _rel(out.float(), ref) < (2e-3 if dt == 0 else 8e-3)


def _split_tf32(w):
    hi = _rna_tf32(w)
    return torch.cat([hi, _rna_tf32(w - hi)], 1).contiguous()


@pytest.mark.parametrize("NB,H,Cin,Cout,stride", [(2, 128, 64, 64, 1), (2, 64, 128, 128, 2), (2, 32, 128, 128, 1), (2, 16, 256, 256, 1),
                                                  (2, 16, 512, 512, 2), (3, 8, 128, 128, 1), (2, 8, 512, 512, 1)])
def test_implicit_conv3x3_3xtf32(lib, cuda_dev, NB, H, Cin, Cout, stride):
    """Error-compensated 3xTF32 on the implicit-convolution modes (row-block boxes, im2col tensor maps, stride 2; the SA
    predictor's shapes down to its 8 x 8 maps) with folded-BN scale / bias + ReLU, against fp64: fp32-grade results."""
    torch.manual_seed(H * Cin + stride)
    x = torch.randn(NB, H, H, Cin, device=cuda_dev)
    w = torch.randn(Cout, Cin, 3, 3, device=cuda_dev) / (9 * Cin) ** 0.5
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    sc, bi = torch.rand(Cout, device=cuda_dev) + 0.5, torch.randn(Cout, device=cuda_dev)
    Ho = (H + 2 - 3) // stride + 1
    out = torch.full((NB, Ho, Ho, Cout), float("nan"), device=cuda_dev)
    assert lib.spe_debug_conv(2, _p(x), _p(_split_tf32(wk)), NB, H, H, Cin, Cout, 3, 3, 1, stride, _p(sc), _p(bi), 1, _p(out), None) == 0
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, stride=stride, padding=1)
    ref = (ref * sc.double()[None, :, None, None] + bi.double()[None, :, None, None]).clamp_min(0).permute(0, 2, 3, 1)
    assert not torch.isnan(out).any()
    assert _rel(out, ref) < 6e-5       # plain TF32 sits at ~5e-4; fp32 accumulation over up to 4608 products
