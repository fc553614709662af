"""T3: the whole drop-in (Transformer / Train through the C ABI) against the CPU oracle (oracle/restatement.py) on
identical weights and tokens.  Tolerances (bf16 tensor-core operands, fp32 accumulate, fp32 master weights):
  forward logits   : max |diff| <= 3e-2 * max|logits_oracle|  (+ per-layer activation taps to localise a failure)
  loss (train.py:88): |L - L_oracle| / |L_oracle| <= 1e-3   (the north-star bound, read as RELATIVE; SURVEY §0 trap 2)
  gradients        : per variable  ||g - g_oracle|| / ||g_oracle|| <= max(3e-2, 2 x the bf16-emulated oracle's own error)
  100-step curve   : every step's loss within 1e-3 relative of the fp32 oracle trained in lock-step
Run on a B200:  python -m pytest tests -m gpu -x -q"""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restatement as R  # noqa: E402  (test infrastructure)
from transformer_b200.Transformer import Transformer  # noqa: E402
from transformer_b200.train import Train, CustomSchedule  # noqa: E402
from transformer_b200.positionalencoding import create_masks  # noqa: E402


def make_pair(L=2, d=128, H=2, f=256, V=1003, rate=0.0, seed=7, perturb=True, pos_rows=128):
    cfg = R.make_cfg(L, d, H, f, V, V + 6, rate=rate, max_position=pos_rows)
    P = R.init_params(L, d, f, V, V + 6, seed=seed, perturb=perturb)
    model = Transformer(L, d, H, f, V, V + 6, rate=rate, pos_rows=pos_rows)
    model.load_variables(P)
    return cfg, P, model


def rel(a, b):
    return (a - b).norm().item() / max(b.norm().item(), 1e-12)


@pytest.mark.parametrize("H", [2, 4])
def test_forward_matches_oracle(H):
    cfg, P, model = make_pair(H=H)
    src, tar = R.synthetic_batch(3, 37, 30, 1003, 1009, pad_tail=True)
    tar_inp = tar[:, :-1]
    taps = {}
    want, w_attn = R.transformer(P, cfg, src, tar_inp, training=False, taps=taps)
    got, g_attn = model((src.cuda(), tar_inp.cuda()), training=False)
    torch.cuda.synchronize()
    for name in ["enc_embed", "enc_layer0", "enc_layer1", "dec_embed", "dec_layer0", "dec_layer1"]:
        t = model.tap(name).cpu()
        err = (t - taps[name]).abs().max().item()
        assert err <= 4e-2 * max(1.0, taps[name].abs().max().item()), f"{name}: {err}"
    scale = want.abs().max().item()
    err = (got.cpu() - want).abs().max().item()
    assert err <= 3e-2 * scale, f"logits err {err} scale {scale}"
    assert set(g_attn.keys()) == set(w_attn.keys())
    for k in w_attn:
        e = (g_attn[k].cpu() - w_attn[k]).abs().max().item()
        assert e < 2e-2, f"{k}: {e}"


def test_six_argument_form():
    cfg, P, model = make_pair()
    src, tar = R.synthetic_batch(2, 21, 18, 1003, 1009, pad_tail=True)
    tar_inp = tar[:, :-1]
    want, _ = R.transformer(P, cfg, src, tar_inp, training=False)
    masks = create_masks(src.cuda(), tar_inp.cuda())
    got, _ = model(src.cuda(), tar_inp.cuda(), False, *masks)            # canonical masks -> fast path
    assert (got.cpu() - want).abs().max().item() <= 3e-2 * want.abs().max().item()
    # non-canonical float masks -> generic-mask kernels
    g = torch.Generator().manual_seed(3)
    em = (torch.rand(2, 1, 1, 21, generator=g) < 0.2).float()
    la = (torch.rand(2, 1, 17, 17, generator=g) < 0.3).float()
    dm = (torch.rand(2, 1, 1, 21, generator=g) < 0.2).float()
    want2, _ = R.transformer(P, cfg, src, tar_inp, training=False, masks=(em, la, dm))
    got2, _ = model(src.cuda(), tar_inp.cuda(), False, em.cuda(), la.cuda(), dm.cuda())
    assert (got2.cpu() - want2).abs().max().item() <= 3e-2 * want2.abs().max().item()
    assert (want2 - want).abs().max().item() > 1e-3                       # the masks actually mattered


def make_trainer(model, batch_size, warmup=60000):
    return Train(1, True, model, None, None, batch_size, None, None, 5, "ckpt", model.d_model, warmup_steps=warmup)


@pytest.mark.parametrize("H,pad", [(2, True), (4, False)])
def test_loss_and_gradients_match_oracle(H, pad):
    cfg, P, model = make_pair(H=H)
    B = 4
    src, tar = R.synthetic_batch(B, 33, 27, 1003, 1009, pad_tail=pad, seed=5)
    loss_o, grads_o, correct_o, total_o = R.loss_and_grads(P, cfg, src, tar, B, training=True)
    R.EMULATE_BF16 = True
    try:
        _, grads_emu, _, _ = R.loss_and_grads(P, cfg, src, tar, B, training=True)
    finally:
        R.EMULATE_BF16 = False
    tr = make_trainer(model, B)
    model.zero_grads()
    tr._step_core(src, tar, True)
    from transformer_b200 import _lib as L
    L.call("b200xf_model_backward", model._handle, -1, L.stream_ptr())
    torch.cuda.synchronize()
    stats = tr._stats.tolist()
    loss = stats[0] / B
    assert abs(loss - float(loss_o)) / abs(float(loss_o)) <= 1e-3, (loss, float(loss_o))
    assert stats[2] == total_o and stats[3] == int((tar[:, 1:] != 0).sum())
    assert abs(stats[1] - correct_o) <= 2                                  # argmax ties/near-ties under bf16
    check_gradients(model, grads_o, grads_emu)


def check_gradients(model, grads_o, grads_emu):
    """per-variable relative L2 error against the fp32 oracle.
    * denominator floored at 5% of the RMS variable-gradient norm: some gradients are mathematically ZERO (key biases:
      softmax is shift-invariant), their oracle value is rounding noise;
    * tolerance = max(3e-2, 2 x the error the oracle ITSELF shows when only its matmul operands are rounded to bf16)
      (R.EMULATE_BF16): on this tiny model bf16 operand rounding alone moves FFN dense0 / LayerNorm-gamma gradients by
      5-6% (ReLU gates of near-zero pre-activations flip); a real kernel bug shows up far above that;
    * the whole flattened gradient must have cosine similarity >= 0.995 with the oracle's."""
    norms = {v.name: grads_o[v.name].norm().item() for v in model.trainable_variables}
    rms = math.sqrt(sum(n * n for n in norms.values()) / len(norms))
    rows, dot, n1, n2 = [], 0.0, 0.0, 0.0
    for v in model.trainable_variables:
        go, g = grads_o[v.name], v.grad.cpu()
        den = max(norms[v.name], 0.05 * rms)
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


def test_test_step_and_loss_function():
    cfg, P, model = make_pair()
    B = 3
    src, tar = R.synthetic_batch(B, 20, 16, 1003, 1009, pad_tail=True, seed=9)
    want = R.test_step(P, cfg, src, tar, B)
    tr = make_trainer(model, B)
    tr.test_step((src, tar))
    assert abs(tr.test_loss.result() - want["loss"]) / want["loss"] <= 1e-3
    assert abs(tr.test_accuracy.result() - want["correct"] / want["total"]) <= 2 / want["total"]
    logits, _ = model((src.cuda(), tar[:, :-1].cuda()), training=False)
    lf = tr.loss_function(tar[:, 1:], logits).item()
    assert abs(lf - want["loss"]) / want["loss"] <= 1e-3


def _lockstep(warmup, steps, H=2, tol=1e-3, calibrate=False):
    """product and oracle take the same train steps side by side.  ``calibrate``: a second oracle whose matmul operands
    are rounded to bf16 (R.EMULATE_BF16) runs the same steps; the gate becomes max(tol, 2 x ITS worst deviation from
    the fp32 oracle) -- the noise floor of any bf16-operand implementation on this trajectory."""
    cfg, P, model = make_pair(H=H, perturb=False)
    B = 4
    batches = [R.synthetic_batch(B, 24, 21, 1003, 1009, pad_tail=(i % 2 == 0), seed=100 + i) for i in range(5)]
    st = R.AdamState(P)
    if calibrate:
        P2 = {k: v.clone() for k, v in P.items()}
        st2 = R.AdamState(P2)
    tr = make_trainer(model, B, warmup)
    worst, emu_worst, worst_at = 0.0, 0.0, -1
    for i in range(steps):
        src, tar = batches[i % len(batches)]
        o = R.train_step(P, cfg, st, src, tar, B, warmup_steps=warmup)
        if calibrate:
            R.EMULATE_BF16 = True
            try:
                o2 = R.train_step(P2, cfg, st2, src, tar, B, warmup_steps=warmup)
            finally:
                R.EMULATE_BF16 = False
            emu_worst = max(emu_worst, abs(o2["loss"] - o["loss"]) / abs(o["loss"]))
        tr.train_loss.reset_states()
        tr.train_step((src, tar))
        loss = tr.train_loss.result()
        r = abs(loss - o["loss"]) / abs(o["loss"])
        if r > worst:
            worst, worst_at = r, i
        if not calibrate:
            assert r <= tol, f"step {i}: loss {loss} oracle {o['loss']} rel {r}"
        assert abs(tr.last_lr - o["lr"]) <= 1e-7 * max(1.0, abs(o["lr"]))
    if calibrate:
        gate = max(tol, 2.0 * emu_worst)
        print(f"bf16-emulated oracle deviates {emu_worst:.3e}; gate {gate:.3e}")
        assert worst <= gate, f"step {worst_at}: rel loss err {worst} > {gate} (bf16-emulated oracle: {emu_worst})"
    return worst, P, model


def test_100_step_loss_curve_reference_schedule():
    """north-star gate: 100 identical steps, dropout 0, warmup 60000 (lr ~ 0: forward+loss parity with frozen weights)"""
    worst, P, model = _lockstep(60000, 100)
    print("worst rel loss err (warmup 60000):", worst)


def test_loss_curve_with_moving_weights():
    """second gate (SURVEY §0 trap 3): warmup_steps=200 -> lr reaches 1.2e-3 by step 40, so Adam really moves the weights
    (the reference's 60000 leaves them frozen for the first 100 steps); optimizer parity.  Trajectories of a bf16 and an
    fp32 run separate chaotically once weights move: the oracle with bf16-rounded matmul operands (R.EMULATE_BF16)
    deviates from its own fp32 self by ~1e-2 relative loss over these 40 steps, so the gate is calibrated in the test:
    max(1e-2, 2 x the deviation of that bf16-emulated oracle run on the same steps).  The north-star 1e-3 bound applies
    to the reference schedule above."""
    worst, P, model = _lockstep(200, 40, tol=1e-2, calibrate=True)
    print("worst rel loss err (warmup 200):", worst)
    ours = torch.cat([v.value.cpu().reshape(-1) for v in model.trainable_variables])
    theirs = torch.cat([P[v.name].reshape(-1) for v in model.trainable_variables])
    drift = rel(ours, theirs)                      # global relative L2 distance between the two trained weight sets
    print("weight drift after 40 moving steps:", drift)
    assert drift <= 2e-2, f"weights drifted from the oracle's: {drift}"


def test_dropout_training_runs_and_is_seeded():
    cfg, P, model = make_pair(rate=0.1)
    B = 4
    src, tar = R.synthetic_batch(B, 24, 21, 1003, 1009, seed=1)
    tr = make_trainer(model, B)
    tr._step_core(src, tar, True); a = tr._stats.clone()
    tr._step_core(src, tar, True); b = tr._stats.clone()
    tr._step_core(src, tar, False); c = tr._stats.clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b)                      # same seed -> same mask
    assert abs(a[0].item() - c[0].item()) > 1e-6  # dropout changes the loss
    tr.train_step((src, tar))                      # full step with dropout regenerates masks in backward
    torch.cuda.synchronize()
    assert math.isfinite(tr.train_loss.result())


def test_base_config_full_size_properties():
    """BASELINE config C3 shapes (transformer-base, S=T=512) at reduced batch: size-independent properties --
    finite loss close to ln(V) per token at init, gradient arena finite and non-zero in every variable."""
    V = 32768
    model = Transformer(6, 512, 8, 2048, V, V, rate=0.0, pos_rows=1024)
    B, S = 4, 512
    src, tar = R.synthetic_batch(B, S, S + 1, V, V, seed=1234)
    tr = make_trainer(model, B)
    tr.train_step((src, tar))
    torch.cuda.synchronize()
    per_tok = tr.train_loss.result() * B / (B * S)
    assert abs(per_tok - math.log(V)) < 0.5, per_tok
    for v in model.trainable_variables:
        assert torch.isfinite(v.grad).all(), v.name
        if not v.name.endswith("embedding"):
            assert v.grad.abs().max().item() > 0, v.name


def test_checkpoint_roundtrip_and_predict(tmp_path):
    """checkpoint/resume (train.py:77-80,159-164) and greedy predict (train.py:91-121) on the drop-in"""
    from transformer_b200.utils import WordTokenizer
    cfg, P, model = make_pair()
    tok = WordTokenizer([f"w{i}" for i in range(1003 - 4)])           # vocab_size + 2 == model vocab
    assert tok.vocab_size + 2 == 1003
    B = 4
    tr = Train(1, True, model, tok, tok, B, None, None, 2, str(tmp_path / "ck"), model.d_model, warmup_steps=100)
    src, tar = R.synthetic_batch(B, 12, 11, 1003, 1009, seed=3)
    for _ in range(3):
        tr.train_step((src, tar))
    p1 = tr.save_checkpoint()
    want = model.params.clone()
    for _ in range(2):
        tr.train_step((src, tar))
    tr.save_checkpoint(); tr.save_checkpoint()
    import os
    assert len(os.listdir(tmp_path / "ck")) == 2 and not os.path.exists(p1)     # max_ckpt_keep = 2
    # restore into a fresh model/trainer and continue: identical to the uninterrupted run
    cfg2, P2, model2 = make_pair(seed=99)
    tr2 = Train(1, True, model2, tok, tok, B, None, None, 2, str(tmp_path / "ck"), model2.d_model, warmup_steps=100)
    assert tr2.load_ckpt() and tr2.optimizer.iterations == 5
    assert torch.equal(model2.params, model.params)
    tr.train_step((src, tar)); tr2.train_step((src, tar))
    torch.cuda.synchronize()
    assert torch.allclose(model2.params, model.params, rtol=0, atol=1e-6)
    assert not torch.equal(want, model.params)
    out = tr.predict("w1 w2 w3")
    assert out.shape == (11,) and out[0].item() == tok.vocab_size and out.dtype == torch.int64
    assert torch.equal(out, tr.predict("w1 w2 w3", use_cache=False))      # key/value cache == the reference's full recomputation
    # greedy decode is a fixed point of the model: re-scoring the produced prefix reproduces each next token
    ids = torch.tensor([[tok.vocab_size] + tok.encode("w1 w2 w3") + [tok.vocab_size + 1]], device="cuda")
    logits, _ = model((ids, out[None, :-1]), training=False)
    assert torch.equal(logits.argmax(-1)[0], out[1:])


def test_forward_matches_reference_code_on_numpy_stub(golden_dir):
    """The CUDA forward against logits / attention weights produced by the reference's OWN Transformer code (run on
    oracle/tf_numpy_stub.py, fixture tests/golden/ref_model_on_numpy_stub.npz): bf16-operand tolerance."""
    import numpy as np
    z = np.load(os.path.join(golden_dir, "ref_model_on_numpy_stub.npz"))
    L_, d, H, f, Vin, Vt, B = (int(v) for v in z["cfg"])
    P = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    model = Transformer(L_, d, H, f, Vin, Vt, rate=0.1, pos_rows=64)
    model.load_variables(P)
    src, tar = torch.from_numpy(z["src"]), torch.from_numpy(z["tar"])
    tar_inp = tar[:, :-1]
    logits, attn = model((src.cuda(), tar_inp.cuda()), training=False)
    torch.cuda.synchronize()
    want = torch.from_numpy(z["logits"])
    assert (logits.cpu() - want).abs().max().item() <= 3e-2 * max(1.0, want.abs().max().item())
    for k in (k[5:] for k in z.files if k.startswith("attn/")):
        assert (attn[k].cpu() - torch.from_numpy(z["attn/" + k])).abs().max().item() <= 2e-2, k
    tr = make_trainer(model, B)
    lf = tr.loss_function(tar[:, 1:], logits).item()
    assert abs(lf - float(z["loss"])) <= 1e-3 * abs(float(z["loss"]))


def test_incremental_decode_matches_full_forward():
    """KV-cached decoding (b200xf_model_decode_begin/_step) against the full forward on the same prefix, batch of 3 with padding in the
    source and a pad id inside the target prefix: logits of every position agree to bf16 rounding."""
    cfg, P, model = make_pair(H=4)
    src, tar = R.synthetic_batch(3, 29, 14, 1003, 1009, pad_tail=True, seed=21)
    tar = tar.clone(); tar[1, 4] = 0                                      # a padding id as a decoder input (masked as a key)
    T = 12
    full, _ = model((src.cuda(), tar[:, :T].cuda()), training=False)
    model.decode_begin(src.cuda())
    for t in range(T):
        step = model.decode_step(tar[:, t].cuda())
        torch.cuda.synchronize()
        err = (step - full[:, t]).abs().max().item()
        assert err <= 2e-2 * max(1.0, full[:, t].abs().max().item()), (t, err)


def test_shape_bucketing_is_neutral():
    """Train(shape_bucket=8): ragged batches are padded to bucket lengths (one CUDA graph per bucket); loss, every metric counter and
    the updated weights must equal the unpadded step's (dropout 0)."""
    B = 4
    out = {}
    for bucket in (0, 8):
        cfg, P, model = make_pair(H=2)
        tr = Train(1, True, model, None, None, B, None, None, 5, "ckpt", model.d_model, warmup_steps=20, shape_bucket=bucket)
        stats = []
        for i, (S, T1) in enumerate([(21, 18), (19, 23), (21, 18), (22, 17), (21, 18)]):        # (21,18) and (22,17) share the bucket (24, 24)
            src, tar = R.synthetic_batch(B, S, T1, 1003, 1009, pad_tail=True, seed=30 + i)
            tr.train_step((src, tar))
            stats.append(tr._stats.clone())
        src, tar = R.synthetic_batch(B, 20, 16, 1003, 1009, pad_tail=True, seed=77)
        tr.test_step((src, tar))
        torch.cuda.synchronize()
        out[bucket] = (torch.stack(stats).cpu(), model.params.clone().cpu(), tr.test_accuracy.result(), tr.train_accuracy.result(), len(tr._graphs))
    s0, p0, ta0, tra0, _ = out[0]
    s8, p8, ta8, tra8, ng = out[8]
    assert torch.equal(s0[:, 1:], s8[:, 1:]), (s0, s8)                     # correct / positions / non-pad counters: exact
    assert torch.allclose(s0[:, 0], s8[:, 0], rtol=2e-3)                   # loss sums (bf16 kernels on a different tile shape)
    assert abs(ta0 - ta8) < 1e-12 and abs(tra0 - tra8) < 1e-12
    assert (p0 - p8).norm().item() <= 2e-3 * p0.norm().item()
    assert ng <= 2                                                          # two buckets seen -> at most two captured graphs
