"""T2: every sm_100a kernel, called through the C ABI (ctypes), against a plain PyTorch fp32 reference of the same
op on the same (bf16-rounded) inputs.  Run on a B200:  python -m pytest tests -m gpu -x -q"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from transformer_b200 import _lib as L  # noqa: E402


def dev():
    assert torch.cuda.is_available(), "gpu tests need a GPU"
    return torch.device("cuda:0")


def sp():
    return L.stream_ptr()


def rnd(shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).to(dev())


def padded(rows, cols, ld, dtype, fill=float("nan")):
    """[rows, ld] buffer whose [:, :cols] view is the matrix (ld in elements)."""
    buf = torch.full((rows, ld), fill, dtype=dtype, device=dev())
    return buf


def ld_for(cols, dtype):
    q = 8 if dtype == torch.bfloat16 else 4
    return (cols + q - 1) // q * q


# ------------------------------------------------------------------------------------------------
# GEMM
# ------------------------------------------------------------------------------------------------
def run_gemm(M, N, K, a_major, b_major, out_f32, bias=False, relu=False, aux=False, accum=False, splits=1, seed=0):
    A = rnd((M, K), 1.0, seed)          # logical A [M,K]
    Bm = rnd((K, N), 1.0, seed + 1)     # logical B [K,N]
    # storage
    if a_major == 0:
        lda = ld_for(K, torch.bfloat16); As = padded(M, K, lda, torch.bfloat16, 0.0); As[:, :K] = A
    else:
        lda = ld_for(M, torch.bfloat16); As = padded(K, M, lda, torch.bfloat16, 0.0); As[:, :M] = A.t()
    if b_major == 0:
        ldb = ld_for(K, torch.bfloat16); Bs = padded(N, K, ldb, torch.bfloat16, 0.0); Bs[:, :K] = Bm.t()
    else:
        ldb = ld_for(N, torch.bfloat16); Bs = padded(K, N, ldb, torch.bfloat16, 0.0); Bs[:, :N] = Bm
    odt = torch.float32 if out_f32 else torch.bfloat16
    ldd = ld_for(N, odt)
    D = torch.full((M, ldd), 7.0, dtype=odt, device=dev())
    ref = A.float() @ Bm.float()
    bias_t = None
    if bias:
        bias_t = rnd((ldd,), 1.0, seed + 2, torch.float32)
        ref = ref + bias_t[:N]
    if relu:
        ref = torch.relu(ref)
    aux_t = None
    ld_aux = 0
    if aux:
        ld_aux = ld_for(N, torch.bfloat16)
        aux_t = rnd((M, ld_aux), 1.0, seed + 3)
        ref = ref * (aux_t[:, :N].float() > 0)
    flags = (1 if relu else 0) | (2 if accum else 0) | (4 if aux else 0)
    if accum:
        init = rnd((M, ldd), 1.0, seed + 4, torch.float32)
        D.copy_(init)
        ref = ref + init[:, :N]
    L.call("b200xf_gemm_bf16", L.ptr(As), lda, a_major, L.ptr(Bs), ldb, b_major, L.ptr(D), ldd, int(out_f32),
           L.ptr(bias_t), L.ptr(aux_t), ld_aux, flags, splits, M, N, K, sp())
    torch.cuda.synchronize()
    got = D[:, :N].float()
    # untouched padding columns must keep their sentinel
    if ldd > N and not accum:
        assert torch.all(D[:, N:] == 7.0), "GEMM wrote outside [M,N]"
    tol = 1e-3 * math.sqrt(K) + (2e-2 if not out_f32 else 0)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * max(1.0, scale / 8), f"gemm M{M} N{N} K{K} maj{a_major}{b_major} f32={out_f32} err {err} (scale {scale})"
    return err


@pytest.fixture(params=[1, 2], ids=["cta1", "ctapair"])
def cta_group(request):
    """run every GEMM test on both tile engines: single-CTA 128xBN tiles and cta_group::2 pair tiles (256x256)"""
    L.load().b200xf_gemm_force_cta_group(request.param)
    yield request.param
    L.load().b200xf_gemm_force_cta_group(0)


@pytest.mark.parametrize("a_major,b_major", [(0, 1), (0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("out_f32", [False, True])
def test_gemm_majors_basic(a_major, b_major, out_f32, cta_group):
    run_gemm(256, 256, 128, a_major, b_major, out_f32)
    run_gemm(128, 64, 64, a_major, b_major, out_f32)          # BN=128 path, single k-block


@pytest.mark.parametrize("a_major,b_major", [(0, 1), (0, 0), (1, 1)])
def test_gemm_ragged_shapes(a_major, b_major, cta_group):
    run_gemm(300, 200, 136, a_major, b_major, True)           # ragged M, N, K
    run_gemm(643, 520, 200, a_major, b_major, False)
    run_gemm(77, 40, 72, a_major, b_major, True)


def test_gemm_epilogues(cta_group):
    run_gemm(384, 512, 256, 0, 1, False, bias=True, relu=True)            # FFN dense0 (point_ffn.py:5)
    run_gemm(384, 512, 256, 0, 1, True, bias=True)                        # dense (Attention.py:77)
    run_gemm(384, 512, 256, 0, 0, False, aux=True)                        # ReLU-backward dgrad
    run_gemm(300, 200, 136, 0, 1, False, bias=True, relu=True)            # ragged + bias tail path
    run_gemm(384, 256, 256, 0, 0, True, accum=True)                       # residual accumulate
    run_gemm(256, 512, 4096, 1, 1, True, accum=True, splits=8)            # wgrad split-K
    run_gemm(512, 2048, 3000, 1, 1, True, accum=True, splits=5)


def test_gemm_relu_bitmask_roundtrip(cta_group):
    """forward FFN epilogue stores a 1-bit (relu(x.W+b) > 0) gate per element; the dgrad epilogue applies it"""
    M, N, K = 700, 512, 192
    A = rnd((M, K), 1.0, 1); W = rnd((K, N), 1.0, 2); bias = rnd((N,), 1.0, 3, torch.float32)
    H = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
    bits = torch.zeros((M, N // 64), dtype=torch.int64, device=dev())
    L.call("b200xf_gemm_bf16", L.ptr(A), K, 0, L.ptr(W), N, 1, L.ptr(H), N, 0, L.ptr(bias), L.ptr(bits), N // 64, 1 | 8, 1, M, N, K, sp())
    torch.cuda.synchronize()
    pre = A.float() @ W.float() + bias
    assert (H.float() - torch.relu(pre)).abs().max().item() < 0.3
    got = ((bits[:, :, None] >> torch.arange(64, device=dev())) & 1).reshape(M, N).bool()
    sure = pre.abs() > 1e-3
    assert torch.equal(got[sure], (pre > 0)[sure])
    # backward: dh = (dy . W2^T) o gate, with W2 [N, K2] keras layout
    K2 = 256
    dy = rnd((M, K2), 1.0, 4); W2 = rnd((N, K2), 1.0, 5)
    dh = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
    L.call("b200xf_gemm_bf16", L.ptr(dy), K2, 0, L.ptr(W2), K2, 0, L.ptr(dh), N, 0, None, L.ptr(bits), N // 64, 16, 1, M, N, K2, sp())
    torch.cuda.synchronize()
    ref = (dy.float() @ W2.float().t()) * got
    assert (dh.float() - ref).abs().max().item() < 0.02 * max(1.0, ref.abs().max().item())


def test_gemm_multi_wave_persistent(cta_group):
    run_gemm(4096, 2048, 512, 0, 1, False, bias=True, relu=True)          # 512 tiles > 148 SMs: ring + TMEM double buffer
    run_gemm(4096, 512, 2048, 0, 1, True, bias=True)
    run_gemm(9000, 1000, 520, 0, 0, True, accum=True)                      # odd pair count, ragged everything
    run_gemm(512, 2048, 8192, 1, 1, True, accum=True, splits=0)            # wgrad, library-chosen split-K


# ------------------------------------------------------------------------------------------------
# attention
# ------------------------------------------------------------------------------------------------
def ref_attention(q, k, v, mask):
    """Attention.py:3-34 in fp32.  q,k,v: (B,H,S,dk) fp32; mask broadcastable float or None."""
    logits = (q @ k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    if mask is not None:
        logits = logits + mask * -1e9
    p = torch.softmax(logits, -1)
    return p @ v, p, torch.logsumexp(logits, -1)


def build_mask(B, Sq, Sk, kpad, causal):
    m = torch.zeros(B, 1, Sq, Sk, device=dev())
    if kpad is not None:
        m = torch.maximum(m, kpad.float()[:, None, None, :])
    if causal:
        m = torch.maximum(m, 1 - torch.tril(torch.ones(Sq, Sk, device=dev())))
    return m


def attn_case(B, H, Sq, Sk, dk, kind, seed=0, backward=True):
    d = H * dk
    ldq = 3 * d
    qkv_q = rnd((B * Sq, ldq), 1.0, seed)            # packed layout like the QKV GEMM output
    qkv_k = qkv_q if Sq == Sk else rnd((B * Sk, ldq), 1.0, seed + 1)
    Qs = qkv_q[:, 0:d]; Ks = qkv_k[:, d:2 * d]; Vs = qkv_k[:, 2 * d:3 * d]
    kpad = None; causal = 0; maskf = None
    g = torch.Generator().manual_seed(seed + 5)
    if kind in ("pad", "causal_pad"):
        lens = torch.randint(max(1, Sk // 2), Sk + 1, (B,), generator=g)
        kpad = (torch.arange(Sk)[None, :] >= lens[:, None]).to(torch.uint8).to(dev())
    if kind == "lead_pad":
        kpad = torch.zeros(B, Sk, dtype=torch.uint8); kpad[0, :3] = 1; kpad[-1, :] = 1; kpad = kpad.to(dev())   # fully-masked rows
    if kind == "all_pad":
        kpad = torch.zeros(B, Sk, dtype=torch.uint8); kpad[0, :] = 1; kpad = kpad.to(dev())
    if kind in ("causal", "causal_pad", "lead_pad"):
        causal = 1
    if kind == "float_mask":
        maskf = (torch.rand(B, 1, Sq, Sk, generator=g) < 0.3).float().to(dev())
    mask = maskf if maskf is not None else build_mask(B, Sq, Sk, kpad, causal)
    q4 = Qs.float().reshape(B, Sq, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    k4 = Ks.float().reshape(B, Sk, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    v4 = Vs.float().reshape(B, Sk, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    o_ref, p_ref, lse_ref = ref_attention(q4, k4, v4, mask)

    O = torch.full((B * Sq, d), float("nan"), dtype=torch.bfloat16, device=dev())
    lse = torch.full((B, H, Sq), float("nan"), dtype=torch.float32, device=dev())
    msb = Sq * Sk if maskf is not None else 0
    msq = Sk if maskf is not None else 0
    L.call("b200xf_attn_fwd", L.ptr(Qs), ldq, L.ptr(Ks), ldq, L.ptr(Vs), ldq, L.ptr(O), d, L.ptr(lse), L.ptr(kpad), causal,
           L.ptr(maskf), msb, msq, B, H, Sq, Sk, dk, sp())
    torch.cuda.synchronize()
    o_got = O.float().reshape(B, Sq, H, dk).permute(0, 2, 1, 3)
    assert torch.isfinite(o_got).all(), f"{kind}: non-finite attention output"
    err_o = (o_got - o_ref).abs().max().item()
    assert err_o < 3e-2, f"attn fwd {kind} B{B} H{H} Sq{Sq} Sk{Sk} dk{dk}: |dO| {err_o}"
    # lse: for fully-masked rows fp32 absorbs ln(Sk) (-1e9 + 6 == -1e9), same as the reference's own fp32
    err_l = ((lse - lse_ref) / lse_ref.abs().clamp(min=1)).abs().max().item()
    assert err_l < 2e-2, f"attn fwd {kind}: lse rel err {err_l}"

    # attention_weights recompute
    P = torch.empty(B, H, Sq, Sk, dtype=torch.float32, device=dev())
    L.call("b200xf_attn_probs", L.ptr(Qs), ldq, L.ptr(Ks), ldq, L.ptr(lse), L.ptr(P), L.ptr(kpad), causal, L.ptr(maskf), msb, msq,
           B, H, Sq, Sk, dk, sp())
    torch.cuda.synchronize()
    err_p = (P - p_ref).abs().max().item()
    assert err_p < 2e-2, f"attn probs {kind}: {err_p}"
    if not backward:
        return
    dO = rnd((B * Sq, d), 1.0, seed + 9)
    do4 = dO.float().reshape(B, Sq, H, dk).permute(0, 2, 1, 3)
    (o_ref * do4).sum().backward()
    delta = torch.empty(B, H, Sq, dtype=torch.float32, device=dev())
    L.call("b200xf_attn_delta", L.ptr(O), d, L.ptr(dO), d, L.ptr(delta), B, Sq, H, dk, sp())
    dq_acc = torch.zeros(B * Sq, d, dtype=torch.float32, device=dev())
    dKV = torch.full((B * Sk, 2 * d), float("nan"), dtype=torch.bfloat16, device=dev())
    dKs = dKV[:, :d]; dVs = dKV[:, d:]
    f32_dq = dk <= 64                       # head depth 128 accumulates dQ in bf16 only (two CTAs per kv tile, see attention.cu)
    if f32_dq:
        L.call("b200xf_attn_bwd", L.ptr(Qs), ldq, L.ptr(Ks), ldq, L.ptr(Vs), ldq, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dq_acc), d, 0,
               L.ptr(dKs), 2 * d, L.ptr(dVs), 2 * d, L.ptr(kpad), causal, L.ptr(maskf), msb, msq, B, H, Sq, Sk, dk, sp())
    else:
        with pytest.raises(L.B200xfError, match="bf16 only"):
            L.call("b200xf_attn_bwd", L.ptr(Qs), ldq, L.ptr(Ks), ldq, L.ptr(Vs), ldq, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dq_acc), d, 0,
                   L.ptr(dKs), 2 * d, L.ptr(dVs), 2 * d, L.ptr(kpad), causal, L.ptr(maskf), msb, msq, B, H, Sq, Sk, dk, sp())
    # the same with dQ reduce-added in bf16 into a strided (packed-QKV-gradient style) tensor
    dq16 = torch.zeros(B * Sq, 3 * d, dtype=torch.bfloat16, device=dev())
    dKV2 = torch.empty_like(dKV)
    L.call("b200xf_attn_bwd", L.ptr(Qs), ldq, L.ptr(Ks), ldq, L.ptr(Vs), ldq, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dq16[:, d:2 * d]), 3 * d, 1,
           L.ptr(dKV2[:, :d]), 2 * d, L.ptr(dKV2[:, d:]), 2 * d, L.ptr(kpad), causal, L.ptr(maskf), msb, msq, B, H, Sq, Sk, dk, sp())
    torch.cuda.synchronize()
    if f32_dq:
        assert torch.equal(dKV2, dKV)
    else:
        dKV, dKs, dVs = dKV2, dKV2[:, :d], dKV2[:, d:]
        dq_acc = dq16[:, d:2 * d].float()
    assert (dq16[:, :d] == 0).all() and (dq16[:, 2 * d:] == 0).all()            # only the addressed column block is touched
    delta_ref = (o_ref.detach() * do4).sum(-1)
    assert (delta - delta_ref).abs().max().item() < 0.15
    for name, got, ref, S in (("dq", dq_acc, q4.grad, Sq), ("dq_bf16", dq16[:, d:2 * d].float(), q4.grad, Sq), ("dk", dKs.float(), k4.grad, Sk),
                              ("dv", dVs.float(), v4.grad, Sk)):
        got4 = got.reshape(B, S, H, dk).permute(0, 2, 1, 3)
        assert torch.isfinite(got4).all(), f"{kind}: non-finite {name}"
        err = (got4 - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err < 4e-2 * max(1.0, scale), f"attn bwd {kind} {name} B{B} H{H} Sq{Sq} Sk{Sk} dk{dk}: err {err} (scale {scale})"


@pytest.mark.parametrize("dk", [64, 32, 128])
@pytest.mark.parametrize("kind", ["none", "pad", "causal", "causal_pad"])
def test_attention_self(dk, kind):
    attn_case(2, 2, 256, 256, dk, kind)
    attn_case(2, 3, 200, 200, dk, kind, seed=3)      # ragged tail
    attn_case(3, 2, 40, 40, dk, kind, seed=5)        # single partial tile (config C1 length)


@pytest.mark.parametrize("dk", [64, 32, 128])
def test_attention_long_and_cross(dk):
    attn_case(1, 2, 640, 640, dk, "causal", seed=7)          # 5 kv tiles, odd slot count
    attn_case(2, 2, 300, 520, dk, "pad", seed=8)             # cross attention: Sq != Sk
    attn_case(2, 2, 130, 70, dk, "none", seed=9)


@pytest.mark.parametrize("kind", ["lead_pad", "all_pad", "float_mask"])
def test_attention_reference_mask_semantics(kind):
    attn_case(2, 2, 200, 200, 64, kind, seed=11)     # fully-masked rows are uniform (never NaN); arbitrary float mask
    attn_case(2, 2, 200, 200, 128, kind, seed=12)    # head depth 128 (the reference's default flags)


# ------------------------------------------------------------------------------------------------
# elementwise / reductions
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,g1_f32,two", [(64, 0, True), (128, 1, False), (512, 0, True), (512, 1, False), (512, 0, False), (1024, 0, True)])
def test_residual_layernorm_fwd_bwd(d, g1_f32, two):
    """bf16 streams in HBM, fp32 math: forward against torch LayerNorm of the fp32 sum; backward against autograd through
    LayerNorm of the SAVED (bf16-rounded) z, with the incoming gradient given as g1 (+ g2)."""
    rows = 333
    y = rnd((rows, d), 1.0, 1); x = rnd((rows, d), 1.0, 2)                      # bf16
    gamma = 1 + rnd((d,), 0.1, 3, torch.float32); beta = rnd((d,), 0.1, 4, torch.float32)
    z_ref = x.float() + y.float()
    out_ref = torch.nn.functional.layer_norm(z_ref, (d,), gamma, beta, 1e-6)
    z16 = torch.empty_like(y); o16 = torch.empty_like(y)
    mean = torch.empty(rows, device=dev()); rstd = torch.empty(rows, device=dev())
    L.call("b200xf_residual_ln_fwd", L.ptr(y), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z16), L.ptr(o16), L.ptr(mean), L.ptr(rstd), None,
           rows, d, 1e-6, 0.0, 0, 0, sp())
    torch.cuda.synchronize()
    assert torch.equal(z16, z_ref.bfloat16())                                     # z is the correctly rounded fp32 sum
    assert (mean - z_ref.mean(-1)).abs().max().item() < 1e-5
    assert (rstd - (z_ref.var(-1, unbiased=False) + 1e-6).rsqrt()).abs().max().item() < 1e-4
    assert (o16.float() - out_ref).abs().max().item() < 3e-2
    # backward
    zs = z16.float().clone().requires_grad_(True)
    out2 = torch.nn.functional.layer_norm(zs, (d,), gamma, beta, 1e-6)
    g1 = rnd((rows, d), 1.0, 5, torch.float32) if g1_f32 else rnd((rows, d), 1.0, 5)
    g2 = rnd((rows, d), 1.0, 6) if two else None
    dout = g1.float() + (g2.float() if two else 0)
    out2.backward(dout)
    dz = torch.empty_like(y); dy16 = torch.full_like(y, float("nan"))
    dg = torch.zeros(d, device=dev()); db = torch.zeros(d, device=dev()); dbias = torch.zeros(d, device=dev())
    L.call("b200xf_residual_ln_bwd", L.ptr(g1), g1_f32, L.ptr(g2) if two else None, L.ptr(z16), L.ptr(mean), L.ptr(rstd), L.ptr(gamma),
           L.ptr(dz), L.ptr(dy16) if two else None, L.ptr(dg), L.ptr(db), L.ptr(dbias), None, rows, d, 0.0, 0, 0, sp())
    torch.cuda.synchronize()
    scale = max(1.0, zs.grad.abs().max().item())
    assert (dz.float() - zs.grad).abs().max().item() < 1.2e-2 * scale           # bf16 output rounding + rounded-z statistics
    if two:
        assert torch.equal(dy16, dz)                                              # no dropout: dy_sub == dz
    assert (dbias - zs.grad.sum(0)).abs().max().item() < 5e-3 * max(1.0, zs.grad.sum(0).abs().max().item())
    g_ref = (dout * (out2.detach() - beta) / gamma).sum(0)
    assert (dg - g_ref).abs().max().item() < 5e-3 * max(1.0, g_ref.abs().max().item())
    assert (db - dout.sum(0)).abs().max().item() < 2e-3 * max(1.0, dout.sum(0).abs().max().item())


def test_dropout_fwd_bwd_consistent():
    rows, d, rate = 257, 512, 0.1
    y = torch.ones(rows, d, device=dev(), dtype=torch.bfloat16); x = torch.zeros_like(y)
    gamma = torch.ones(d, device=dev()); beta = torch.zeros(d, device=dev())
    z = torch.empty_like(y); o16 = torch.empty_like(y)
    mean = torch.empty(rows, device=dev()); rstd = torch.empty(rows, device=dev())
    bits = torch.zeros(rows, d // 8, dtype=torch.uint8, device=dev())
    L.call("b200xf_residual_ln_fwd", L.ptr(y), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z), L.ptr(o16), L.ptr(mean), L.ptr(rstd), L.ptr(bits),
           rows, d, 1e-6, rate, 1234, 7, sp())
    torch.cuda.synchronize()
    keep = z != 0
    unpacked = ((bits[:, :, None].int() >> torch.arange(8, device=dev())) & 1).bool().view(rows, d)
    assert torch.equal(unpacked, keep)                               # stored keep mask: bit e of byte j = element 8j+e
    frac = keep.float().mean().item()
    assert abs(frac - (1 - rate)) < 0.01, frac                      # keep probability
    assert torch.allclose(z[keep].float(), torch.full_like(z[keep].float(), 1 / (1 - rate)), rtol=4e-3)    # inverted-dropout scale
    # backward regenerates the same mask: dy_sub == dz * mask/(1-rate)
    dout = (torch.ones(rows, d, device=dev()) * (torch.arange(d, device=dev()) % 64)).bfloat16()
    dz = torch.empty_like(y); dy16 = torch.empty_like(y)
    dg = torch.zeros(d, device=dev()); db = torch.zeros(d, device=dev())
    for kb in (None, bits):       # mask regenerated from (seed, site), or read back from the bits the forward stored
        dy16.fill_(float("nan"))
        L.call("b200xf_residual_ln_bwd", L.ptr(dout), 0, None, L.ptr(z), L.ptr(mean), L.ptr(rstd), L.ptr(gamma), L.ptr(dz), L.ptr(dy16),
               L.ptr(dg), L.ptr(db), None, L.ptr(kb) if kb is not None else None, rows, d, rate, 1234, 7, sp())
        torch.cuda.synchronize()
        want = dz.float() * keep / (1 - rate)
        assert (dy16.float() - want).abs().max().item() <= 0.02 * want.abs().max().item() + 1e-3
    # a different site gives a different mask
    z2 = torch.empty_like(y)
    L.call("b200xf_residual_ln_fwd", L.ptr(y), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z2), L.ptr(o16), L.ptr(mean), L.ptr(rstd), None,
           rows, d, 1e-6, rate, 1234, 8, sp())
    torch.cuda.synchronize()
    assert (z2 != z).any()


def test_embed_fwd_bwd():
    B, S, d, V = 5, 37, 128, 1000
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, V, (B, S), generator=g).to(dev())
    table = rnd((V, d), 0.05, 1, torch.float32); pos = rnd((64, d), 1.0, 2, torch.float32)
    o32 = torch.empty(B * S, d, device=dev()); o16 = torch.empty(B * S, d, dtype=torch.bfloat16, device=dev())
    L.call("b200xf_embed_fwd", L.ptr(ids), S, 0, L.ptr(table), L.ptr(pos), L.ptr(o32), L.ptr(o16), B * S, S, d, V, 0.0, 0, 0, sp())
    torch.cuda.synchronize()
    ref = table[ids] * math.sqrt(d) + pos[:S][None]
    assert (o32.view(B, S, d) - ref).abs().max().item() < 1e-6
    assert (o16.float().view(B, S, d) - ref).abs().max().item() < 2e-2
    o16b = torch.empty_like(o16)
    L.call("b200xf_embed_fwd", L.ptr(ids), S, 0, L.ptr(table), L.ptr(pos), None, L.ptr(o16b), B * S, S, d, V, 0.0, 0, 0, sp())   # bf16 only
    torch.cuda.synchronize()
    assert torch.equal(o16b, o16)
    g1 = rnd((B * S, d), 1.0, 3); g2 = rnd((B * S, d), 1.0, 4)                  # the gradient arrives as two bf16 streams
    for second in (g2, None):
        dt = torch.zeros(V, d, device=dev())
        L.call("b200xf_embed_bwd", L.ptr(ids), S, 0, L.ptr(g1), L.ptr(second) if second is not None else None, L.ptr(dt), B * S, S, d, V,
               0.0, 0, 0, sp())
        torch.cuda.synchronize()
        dx = g1.float() + (second.float() if second is not None else 0)
        dref = torch.zeros(V, d, device=dev()).index_add_(0, ids.view(-1), dx * math.sqrt(d))
        assert (dt - dref).abs().max().item() < 1e-3


@pytest.mark.parametrize("logits_f32,V,ld,inplace", [(1, 1003, 1008, False), (0, 1003, 1008, False), (0, 1003, 1008, True),
                                                     (0, 8190, 8192, True), (0, 32768, 32768, True), (0, 20001, 20008, False),
                                                     (0, 40000, 40000, True)])
def test_xent_loss_grad_accuracy(logits_f32, V, ld, inplace):
    """two-pass kernel (fp32 logits, V > 32768) and the register-resident row kernel (bf16, KV 4 / 16), also in place."""
    B, T = 4, 9
    g = torch.Generator().manual_seed(1)
    logits = torch.randn(B * T, ld, generator=g).to(dev()) * 3
    tar = torch.randint(1, V, (B, T + 1), generator=g); tar[1, 5:] = 0; tar[3, 1:] = 0
    tar = tar.to(dev())
    real = tar[:, 1:]
    lg = logits if logits_f32 else logits.bfloat16()
    lref = lg[:, :V].float().view(B, T, V).clone().requires_grad_(True)
    lse = torch.logsumexp(lref, -1)
    per = (lse - torch.gather(lref, -1, real[..., None]).squeeze(-1)) * (real != 0)
    loss = per.sum() / B
    loss.backward()
    dl = lg if inplace else torch.full((B * T, ld), float("nan"), dtype=torch.bfloat16, device=dev())
    stats = torch.zeros(4, dtype=torch.float64, device=dev())
    row_loss = torch.empty(B * T, device=dev())
    L.call("b200xf_xent", L.ptr(lg), logits_f32, ld, L.ptr(tar), T + 1, 1, T, L.ptr(dl), ld, L.ptr(row_loss), L.ptr(stats), B * T, V, 1.0 / B, sp())
    torch.cuda.synchronize()
    assert abs(stats[0].item() / B - loss.item()) < 1e-4 * abs(loss.item())
    assert (row_loss.view(B, T) - per.detach()).abs().max().item() < 1e-3
    correct = (lref.argmax(-1) == real).sum().item()
    assert stats[1].item() == correct and stats[2].item() == B * T and stats[3].item() == (real != 0).sum().item()
    assert (dl[:, :V].float().view(B, T, V) - lref.grad).abs().max().item() < 2e-3
    assert torch.all(dl[:, V:] == 0)


def test_colsum_adam_casts_padmask():
    rows, N = 1000, 520
    X = rnd((rows, 528), 1.0, 0)
    out = torch.zeros(N, device=dev())
    L.call("b200xf_colsum_bf16", L.ptr(X), 528, L.ptr(out), rows, N, sp())
    torch.cuda.synchronize()
    assert (out - X[:, :N].float().sum(0)).abs().max().item() < 1e-2
    n = 4096 * 3
    p = rnd((n,), 1.0, 1, torch.float32); m = rnd((n,), 0.1, 2, torch.float32); v = rnd((n,), 0.1, 3, torch.float32).abs()
    g = rnd((n,), 1.0, 4, torch.float32)
    p0, m0, v0 = p.clone(), m.clone(), v.clone()
    sh = torch.empty(n, dtype=torch.bfloat16, device=dev())
    alpha, b1, b2, eps = 1e-3, 0.9, 0.98, 1e-9
    L.call("b200xf_adam_step", L.ptr(p), L.ptr(m), L.ptr(v), L.ptr(g), L.ptr(sh), n, alpha, b1, b2, eps, sp())
    torch.cuda.synchronize()
    m1 = b1 * m0 + (1 - b1) * g; v1 = b2 * v0 + (1 - b2) * g * g; p1 = p0 - alpha * m1 / (v1.sqrt() + eps)
    assert (m - m1).abs().max().item() < 1e-6 and (v - v1).abs().max().item() < 1e-6 and (p - p1).abs().max().item() < 1e-6
    assert torch.equal(sh, p.bfloat16())
    ids = torch.tensor([[3, 0, 5, 0], [0, 7, 0, 0]], device=dev()); mask = torch.empty(6, dtype=torch.uint8, device=dev())
    L.call("b200xf_pad_mask", L.ptr(ids), 4, 1, L.ptr(mask), 2, 3, sp())      # columns 1..3 of each row (tar[:, 1:])
    torch.cuda.synchronize()
    assert mask.tolist() == [1, 0, 1, 0, 1, 1]


def test_attention_delta_from_output_residual():
    """dS = P o (dP - delta) cancels when the value vectors share a common component (a bias / LayerNorm beta upstream): delta
    formed from the bf16-ROUNDED attention output then puts percent-level errors into dQ and dK.  The training forward keeps
    the rounding residual O_lo (b200xf_attn_fwd_residual) and delta is formed from O + O_lo: errors back at the operand-rounding
    floor.  (fp64 torch reference.)"""
    B, H, S, dk = 2, 4, 512, 64
    d = H * dk
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B * S, 3 * d, generator=g).to(dev())
    qkv[:, :2 * d] *= 0.3
    qkv[:, 2 * d:] = 0.3 * qkv[:, 2 * d:] + 3.0 * torch.randn(1, d, generator=g).to(dev())      # common component >> per-key part
    qkv = qkv.bfloat16()
    Q, K, V = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    q4, k4, v4 = (t.double().reshape(B, S, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True) for t in (Q, K, V))
    o_ref = torch.softmax((q4 @ k4.transpose(-1, -2)) / math.sqrt(dk), -1) @ v4
    dO = (torch.randn(B * S, d, generator=g).to(dev()) * 0.01).bfloat16()
    (o_ref * dO.double().reshape(B, S, H, dk).permute(0, 2, 1, 3)).sum().backward()
    O = torch.empty(B * S, d, dtype=torch.bfloat16, device=dev()); O_lo = torch.full_like(O, float("nan"))
    lse = torch.empty(B, H, S, device=dev())
    L.call("b200xf_attn_fwd_residual", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(O), L.ptr(O_lo), d, L.ptr(lse), None, 0, None, 0, 0,
           B, H, S, S, dk, sp())
    torch.cuda.synchronize()
    # the residual is the rounding error of O: at most half a bf16 ulp of the element it corrects, and not identically zero
    assert torch.isfinite(O_lo.float()).all() and (O_lo.float().abs() <= O.float().abs() * 2.0 ** -8 + 1e-30).all()
    assert O_lo.float().abs().max().item() > 0
    errs = {}
    for name, lo in (("rounded", None), ("residual", O_lo)):
        delta = torch.empty(B, H, S, device=dev())
        if lo is None:
            L.call("b200xf_attn_delta", L.ptr(O), d, L.ptr(dO), d, L.ptr(delta), B, S, H, dk, sp())
        else:
            L.call("b200xf_attn_delta_residual", L.ptr(O), L.ptr(lo), d, L.ptr(dO), d, L.ptr(delta), B, S, H, dk, sp())
        dq = torch.zeros(B * S, d, dtype=torch.float32, device=dev())
        dKV = torch.empty(B * S, 2 * d, dtype=torch.bfloat16, device=dev())
        L.call("b200xf_attn_bwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dq), d, 0,
               L.ptr(dKV[:, :d]), 2 * d, L.ptr(dKV[:, d:]), 2 * d, None, 0, None, 0, 0, B, H, S, S, dk, sp())
        torch.cuda.synchronize()
        for nm, got, ref in (("dq", dq, q4.grad), ("dk", dKV[:, :d], k4.grad)):
            got4 = got.double().reshape(B, S, H, dk).permute(0, 2, 1, 3)
            errs[(name, nm)] = (got4 - ref).norm().item() / ref.norm().item()
    print(errs)
    assert errs[("residual", "dq")] < 4e-3 and errs[("residual", "dk")] < 4e-3
    assert errs[("rounded", "dk")] > 3 * errs[("residual", "dk")]          # the effect the residual removes


@pytest.mark.parametrize("M,N,K,rate", [(512, 512, 512, 0.0), (1000, 512, 2048, 0.1), (300, 256, 384, 0.1), (257, 128, 128, 0.0), (4096, 512, 512, 0.1)])
def test_gemm_bias_dropout_residual_layernorm_fused(M, N, K, rate):
    """K3/K4: the LayerNorm-in-the-epilogue GEMM against (a) fp32 torch on the same bf16 operands and (b) the stand-alone
    GEMM + residual_ln_fwd pair of this library: same dropout mask (bit-identical keep bytes), same z, same statistics."""
    A = rnd((M, K), 1.0, 1); W = rnd((K, N), K ** -0.5, 2)
    bias = rnd((N,), 0.2, 3, torch.float32); x = rnd((M, N), 1.0, 4)
    gamma = 1 + rnd((N,), 0.1, 5, torch.float32); beta = rnd((N,), 0.1, 6, torch.float32)
    z = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev()); out = torch.full_like(z, float("nan"))
    mean = torch.full((M,), float("nan"), device=dev()); rstd = torch.full((M,), float("nan"), device=dev())
    keep = torch.zeros(M, N // 8, dtype=torch.uint8, device=dev())
    seed, site = 1234, 7
    L.call("b200xf_gemm_bias_dropout_residual_ln", L.ptr(A), K, L.ptr(W), N, L.ptr(bias), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z), L.ptr(out),
           L.ptr(mean), L.ptr(rstd), L.ptr(keep), M, N, K, 1e-6, rate, seed, site, sp())
    # the two-kernel path
    y2 = torch.empty_like(z); z2 = torch.empty_like(z); out2 = torch.empty_like(z)
    mean2 = torch.empty_like(mean); rstd2 = torch.empty_like(rstd); keep2 = torch.zeros_like(keep)
    L.call("b200xf_gemm_bf16", L.ptr(A), K, 0, L.ptr(W), N, 1, L.ptr(y2), N, 0, L.ptr(bias), None, 0, 0, 1, M, N, K, sp())
    L.call("b200xf_residual_ln_fwd", L.ptr(y2), L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(z2), L.ptr(out2), L.ptr(mean2), L.ptr(rstd2), L.ptr(keep2),
           M, N, 1e-6, rate, seed, site, sp())
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.isfinite(z.float()).all()
    if rate > 0:
        assert torch.equal(keep, keep2)                                       # the same Philox mask, bit for bit
        frac = ((keep[:, :, None].int() >> torch.arange(8, device=dev())) & 1).float().mean().item()
        assert abs(frac - (1 - rate)) < 0.01
    # fp32 reference on the same operands (and the mask the kernel reported)
    y = A.float() @ W.float() + bias
    if rate > 0:
        kb = ((keep[:, :, None].int() >> torch.arange(8, device=dev())) & 1).bool().view(M, N)
        y = torch.where(kb, y / (1 - rate), torch.zeros_like(y))
    z_ref = x.float() + y
    out_ref = torch.nn.functional.layer_norm(z_ref, (N,), gamma, beta, 1e-6)
    assert (z.float() - z_ref).abs().max().item() < 2e-2 * max(1.0, z_ref.abs().max().item())
    assert (mean - z_ref.mean(-1)).abs().max().item() < 1e-4
    assert ((rstd - (z_ref.var(-1, unbiased=False) + 1e-6).rsqrt()) / rstd).abs().max().item() < 1e-3
    assert (out.float() - out_ref).abs().max().item() < 4e-2
    # the fused kernel normalises the UNROUNDED fp32 GEMM output, the two-kernel path a bf16-rounded y: agreement to bf16 rounding
    assert (z.float() - z2.float()).abs().max().item() < 4e-2 * max(1.0, z_ref.abs().max().item())
    assert (out.float() - out2.float()).abs().max().item() < 6e-2


@pytest.mark.parametrize("M,N,K,flags", [(1000, 2048, 512, 16), (4096 + 40, 320, 256, 0), (130, 64, 128, 0)])
def test_gemm_output_column_sums(M, N, K, flags):
    """The dgrad GEMM that produces dy for the layer below also accumulates colsum(dy) = that layer's bias gradient from its
    staging tile (ragged M / N tiles, the ReLU bit-mask epilogue of the FFN backward): must equal the column sums of the bf16
    D it wrote, added to what the buffer held."""
    A = rnd((M, K), 1.0, 1); B = rnd((N, K), K ** -0.5, 2)          # dgrad form: both operands K-major
    D = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
    cs = torch.full((N,), 3.0, device=dev())
    aux, ld_aux = None, 0
    if flags & 16:
        bits = torch.randint(-2 ** 62, 2 ** 62, (M, N // 64), dtype=torch.int64, device=dev())
        aux, ld_aux = L.ptr(bits), N // 64
    L.call("b200xf_gemm_bf16_colsum", L.ptr(A), K, 0, L.ptr(B), K, 0, L.ptr(D), N, None, aux, ld_aux, flags, M, N, K, L.ptr(cs), sp())
    D2 = torch.empty_like(D)
    L.call("b200xf_gemm_bf16", L.ptr(A), K, 0, L.ptr(B), K, 0, L.ptr(D2), N, 0, None, aux, ld_aux, flags, 1, M, N, K, sp())
    torch.cuda.synchronize()
    assert torch.equal(D, D2)                                        # the product itself is untouched
    ref = 3.0 + D.double().sum(0)
    assert (cs.double() - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item())


def test_gemm_chunked_schedule_matches_persistent():
    """b200xf_gemm_set_chunks_per_unit: hardware-scheduled chunks of tiles (data-parallel backward, NCCL sharing the SMs) must
    produce the persistent grid's results -- bit-identical without split-K, to fp32 reduction order with it."""
    lib = L.load()
    M, N, K = 8192 + 96, 1024 + 64, 512
    A = rnd((M, K), 1.0, 1); B = rnd((K, N), K ** -0.5, 2); bias = rnd((N,), 0.1, 3, torch.float32)
    outs = []
    for chunks in (0, 4, 1):
        lib.b200xf_gemm_set_chunks_per_unit(chunks)
        D = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=dev())
        L.call("b200xf_gemm_bf16", L.ptr(A), K, 0, L.ptr(B), N, 1, L.ptr(D), N, 0, L.ptr(bias), None, 0, 0, 1, M, N, K, sp())
        # wgrad form: split-K over the tokens with fp32 reduce-add
        W = torch.zeros(K, N, dtype=torch.float32, device=dev())
        dY = rnd((M, N), 1.0, 4)
        L.call("b200xf_gemm_bf16", L.ptr(A), K, 1, L.ptr(dY), N, 1, L.ptr(W), N, 1, None, None, 0, 2, 0, K, N, M, sp())
        torch.cuda.synchronize()
        outs.append((D, W))
    lib.b200xf_gemm_set_chunks_per_unit(0)
    ref = A.float() @ B.float() + bias
    assert (outs[0][0].float() - ref).abs().max().item() < 3e-2 * max(1.0, ref.abs().max().item())
    for D, W in outs[1:]:
        assert torch.equal(D, outs[0][0])
        assert (W - outs[0][1]).abs().max().item() < 1e-3 * outs[0][1].abs().max().item()
