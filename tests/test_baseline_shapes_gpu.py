"""Oracle parity AT THE BASELINE SHAPES (BASELINE.json configs): the CUDA path against oracle/restatement.py on identical
weights and tokens, with the tolerances of tests/test_model_gpu.py unchanged:
  logits  max|diff| <= 3e-2 * max|logits_oracle| ; loss |L - L_o|/|L_o| <= 1e-3 ;
  gradients per variable <= max(3e-2, 2 x the bf16-emulated oracle's own error), whole-gradient cosine >= 0.995.
Shapes: (a) transformer-base 6L/512/8h/2048, V 32768, S=T=512 (config C3) at batch 2; (b) base S=T=128 (C2) at batch 4;
(c) transformer-big 6L/1024/16h/4096, S=T=256 (C4) at batch 1; (d) the reference's DEFAULT flags (utils.py:26-29: 4L, d 512,
dff 1024, 4 heads => head depth 128); (e) attention forward+backward at S = 1024 and 2048 (C5's long end), incl. the error of
the bf16 red.global dQ accumulation against the fp32 TMA-reduce variant as the number of kv tiles grows.
The batch is reduced (the CPU oracle materialises P and fp32 logits); every other dimension is the BASELINE one."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restatement as R  # noqa: E402  (test infrastructure)
from transformer_b200 import _lib as L  # noqa: E402
from transformer_b200.Transformer import Transformer  # noqa: E402
from transformer_b200.train import Train  # noqa: E402


def _check_gradients(model, grads_o, grads_emu, floor=0.05):
    norms = {v.name: grads_o[v.name].norm().item() for v in model.trainable_variables}
    rms = math.sqrt(sum(n * n for n in norms.values()) / len(norms))
    rows, dot, n1, n2 = [], 0.0, 0.0, 0.0
    for v in model.trainable_variables:
        go, g = grads_o[v.name], v.grad.cpu()
        den = max(norms[v.name], floor * rms)
        rows.append(((g - go).norm().item() / den, (grads_emu[v.name] - go).norm().item() / den, v.name))
        dot += (g * go).sum().item(); n1 += (g * g).sum().item(); n2 += (go * go).sum().item()
    rows.sort(reverse=True)
    print("gradient rel-L2 errors vs fp32 oracle (worst 8): cuda | bf16-emulated oracle")
    for e, ee, n in rows[:8]:
        print(f"   {e:.4f} | {ee:.4f}  {n}")
    cos = dot / math.sqrt(n1 * n2)
    print("cosine(whole gradient) =", cos)
    assert cos >= 0.995, cos
    for e, ee, n in rows:
        tol = max(3e-2, 2.0 * ee)
        assert e <= tol, f"grad {n}: rel err {e} > {tol} (bf16-emulated oracle: {ee})"


def _model_parity(L_, d, H, f, V, B, S, seed=42, pad_tail=False):
    cfg = R.make_cfg(L_, d, H, f, V, V, rate=0.0, max_position=S + 1)
    P = R.init_params(L_, d, f, V, V, seed=seed, perturb=True)
    model = Transformer(L_, d, H, f, V, V, rate=0.0, pos_rows=S + 1, max_batch=B, max_src_len=S, max_tgt_len=S)
    model.load_variables(P)
    src, tar = R.synthetic_batch(B, S, S + 1, V, V, seed=1234, pad_tail=pad_tail)
    # ---- forward logits (inference mode), both through the public call ----
    want, _ = R.transformer(P, cfg, src, tar[:, :-1], training=False)
    got, _ = model((src.cuda(), tar[:, :-1].cuda()), training=False)
    torch.cuda.synchronize()
    scale = want.abs().max().item()
    err = (got.cpu() - want).abs().max().item()
    print(f"logits: max|diff| {err:.4e} (scale {scale:.3f}, ratio {err / scale:.2e})")
    assert err <= 3e-2 * scale, f"logits err {err} scale {scale}"
    del got, want
    # ---- loss + every gradient ----
    loss_o, grads_o, correct_o, total_o = R.loss_and_grads(P, cfg, src, tar, B, training=True)
    R.EMULATE_BF16 = True
    try:
        _, grads_emu, _, _ = R.loss_and_grads(P, cfg, src, tar, B, training=True)
    finally:
        R.EMULATE_BF16 = False
    tr = Train(1, True, model, None, None, B, None, None, 5, "ckpt", d)
    model.zero_grads()
    tr._step_core(src, tar, True)
    L.call("b200xf_model_backward", model._handle, -1, L.stream_ptr())
    torch.cuda.synchronize()
    stats = tr._stats.tolist()
    loss = stats[0] / B
    rel = abs(loss - float(loss_o)) / abs(float(loss_o))
    print(f"loss {loss:.5f} oracle {float(loss_o):.5f} rel {rel:.2e}")
    assert rel <= 1e-3, (loss, float(loss_o))
    assert stats[2] == total_o and stats[3] == int((tar[:, 1:] != 0).sum())
    assert abs(stats[1] - correct_o) <= max(2, int(2e-3 * total_o))          # argmax near-ties under bf16 (V = 32768 classes)
    _check_gradients(model, grads_o, grads_emu)


def test_base_s512_matches_oracle():
    """BASELINE C3: transformer-base (6L, d 512, 8 heads, dff 2048), V 32768, S = T = 512; batch 2 of 64"""
    _model_parity(6, 512, 8, 2048, 32768, B=2, S=512)


def test_base_s128_matches_oracle():
    """BASELINE C2: transformer-base, S = T = 128; batch 4 of 256, ragged (zero-padded) rows"""
    _model_parity(6, 512, 8, 2048, 32768, B=4, S=128, pad_tail=True)


def test_big_s256_matches_oracle():
    """BASELINE C4: transformer-big (6L, d 1024, 16 heads, dff 4096), S = T = 256; batch 1 of 16 per GPU"""
    _model_parity(6, 1024, 16, 4096, 32768, B=1, S=256)


def test_reference_default_flags_shape_matches_oracle():
    """the reference's default CLI configuration (utils.py:26-29): 4 layers, d_model 512, dff 1024, 4 heads -> head depth 128"""
    _model_parity(4, 512, 4, 1024, 8000, B=3, S=50, pad_tail=True)


# ------------------------------------------------------------------------------------------------------------------
# attention at the long end of the C5 sweep
# ------------------------------------------------------------------------------------------------------------------
def _attn_long(S, causal, dk=64, B=1, H=2, seed=0):
    dev = torch.device("cuda:0")
    d = H * dk
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(B * S, 3 * d, generator=g).to(dev).bfloat16()
    Q, K, V = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
    q4 = Q.float().reshape(B, S, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    k4 = K.float().reshape(B, S, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    v4 = V.float().reshape(B, S, H, dk).permute(0, 2, 1, 3).contiguous().requires_grad_(True)
    logits = (q4 @ k4.transpose(-1, -2)) / math.sqrt(dk)
    if causal:
        logits = logits + (1 - torch.tril(torch.ones(S, S, device=dev))) * -1e9
    o_ref = torch.softmax(logits, -1) @ v4
    O = torch.empty(B * S, d, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, S, device=dev)
    sp = L.stream_ptr()
    L.call("b200xf_attn_fwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(O), d, L.ptr(lse), None, causal, None, 0, 0,
           B, H, S, S, dk, sp)
    torch.cuda.synchronize()
    o_got = O.float().reshape(B, S, H, dk).permute(0, 2, 1, 3)
    err_o = (o_got - o_ref).abs().max().item()
    assert err_o < 3e-2, f"attention fwd S={S}: {err_o}"
    dO = torch.randn(B * S, d, generator=g).to(dev).bfloat16()
    (o_ref * dO.float().reshape(B, S, H, dk).permute(0, 2, 1, 3)).sum().backward()
    delta = torch.empty(B, H, S, device=dev)
    L.call("b200xf_attn_delta", L.ptr(O), d, L.ptr(dO), d, L.ptr(delta), B, S, H, dk, sp)
    out = {}
    for name, bf in (("fp32_tma_reduce", 0), ("bf16_red_global", 1)):
        dq = torch.zeros(B * S, d, dtype=torch.bfloat16 if bf else torch.float32, device=dev)
        dKV = torch.empty(B * S, 2 * d, dtype=torch.bfloat16, device=dev)
        L.call("b200xf_attn_bwd", L.ptr(Q), 3 * d, L.ptr(K), 3 * d, L.ptr(V), 3 * d, L.ptr(dO), d, L.ptr(lse), L.ptr(delta), L.ptr(dq), d, bf,
               L.ptr(dKV[:, :d]), 2 * d, L.ptr(dKV[:, d:]), 2 * d, None, causal, None, 0, 0, B, H, S, S, dk, sp)
        torch.cuda.synchronize()
        for nm, got, ref in (("dq", dq.float(), q4.grad), ("dk", dKV[:, :d].float(), k4.grad), ("dv", dKV[:, d:].float(), v4.grad)):
            got4 = got.reshape(B, S, H, dk).permute(0, 2, 1, 3)
            e = (got4 - ref).norm().item() / ref.norm().item()
            m = (got4 - ref).abs().max().item() / max(1.0, ref.abs().max().item())
            out[(name, nm)] = (e, m)
            assert m < 4e-2, f"attention bwd S={S} {name} {nm}: max err ratio {m}"
    return out


@pytest.mark.parametrize("causal", [0, 1])
def test_attention_long_sequences_and_dq_accumulation_error(causal):
    """S = 1024 and 2048 (8 and 16 kv tiles): forward, dK/dV, and dQ accumulated over the kv tiles in bf16 (red.global.add
    .bf16x2, the training path) against the fp32 TMA reduce-add variant -- relative L2 error per variant and tile count."""
    rows = []
    for S in (256, 1024, 2048):
        o = _attn_long(S, causal)
        rows.append((S, (S + 127) // 128, o[("fp32_tma_reduce", "dq")][0], o[("bf16_red_global", "dq")][0]))
    print("S | kv tiles | dQ rel-L2 err fp32 accumulate | bf16 accumulate")
    for r in rows:
        print("  %5d | %2d | %.3e | %.3e" % r)
    for S, nt, e32, e16 in rows:
        assert e32 < 1e-2, (S, e32)          # bf16 operands, fp32 accumulate
        assert e16 < 2.5e-2, (S, e16)        # + one bf16 rounding per kv tile and per add
