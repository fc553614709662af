"""T0/T1: pin the CPU oracle (oracle/restatement.py) against the golden vectors generated from the reference's own
positional_encoding and against the closed-form known answers of SURVEY.md Appendix D."""
import math
import os

import numpy as np
import torch

from oracle import restatement as R


def test_posenc_matches_reference_golden(golden_dir):
    for pos, d in ((64, 8), (50, 128)):
        want = np.load(os.path.join(golden_dir, f"posenc_{pos}x{d}.npy"))
        got = R.positional_encoding(pos, d).numpy()
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got, want)          # bit-exact: same NumPy float64 ops
    for pos, d in ((2048, 512), (300, 1024)):
        got = R.positional_encoding(pos, d).numpy()
        assert np.array_equal(got[0, ::37, ::13], np.load(os.path.join(golden_dir, f"posenc_{pos}x{d}_sample37x13.npy")))
        sums = np.load(os.path.join(golden_dir, f"posenc_{pos}x{d}_sums.npy"))
        assert got.astype(np.float64).sum() == sums[0] and np.abs(got.astype(np.float64)).sum() == sums[1]


def test_posenc_known_answers():
    pe = R.positional_encoding(64, 8).numpy()[0]
    np.testing.assert_allclose(pe[0], [0, 0, 0, 0, 1, 1, 1, 1], atol=0)
    np.testing.assert_allclose(pe[1], [0.841471, 0.0998334, 0.0099998, 0.001, 0.5403023, 0.9950042, 0.99995, 0.9999995], rtol=5e-6)  # Appendix D values are 7-digit roundings
    pe = R.positional_encoding(2048, 512).numpy()
    assert abs(pe.astype(np.float64).sum() - 197141.075816) < 1e-3
    assert abs(pe[0, 1, 0] - math.sin(1)) < 1e-7 and abs(pe[0, 1, 256] - math.cos(1)) < 1e-7
    assert abs(pe[0, 5, 1] - (-0.99385476)) < 1e-6


def test_masks_known_answers():
    assert R.create_look_ahead_mask(3).tolist() == [[0, 1, 1], [0, 0, 1], [0, 0, 0]]
    assert R.create_padding_mask(torch.tensor([[7, 6, 0, 0, 1]])).tolist() == [[[[0, 0, 1, 1, 0]]]]
    inp = torch.tensor([[5, 3, 0]]); tar = torch.tensor([[2, 0]])
    e, c, d = R.create_masks(inp, tar)
    assert e.shape == (1, 1, 1, 3) and d.tolist() == e.tolist()
    assert c.tolist() == [[[[0, 1], [0, 1]]]]


def test_schedule_known_answers():
    assert R.custom_schedule(0, 512) == 0.0                       # lr == 0 on the first step
    np.testing.assert_allclose(R.custom_schedule(1, 512), 3.0070326e-9, rtol=1e-6)
    np.testing.assert_allclose(R.custom_schedule(100, 512), 3.0070325e-7, rtol=1e-6)
    np.testing.assert_allclose(R.custom_schedule(60000, 512), 1.8042195e-4, rtol=1e-6)


def test_layernorm_known_answers():
    x = torch.tensor([[1., 2., 3., 4.]], dtype=torch.float64)
    y = R.layer_norm(x, torch.ones(4, dtype=torch.float64), torch.zeros(4, dtype=torch.float64))
    np.testing.assert_allclose(y[0].numpy(), [-1.34164025, -0.44721342, 0.44721342, 1.34164025], rtol=1e-7)
    c = R.layer_norm(torch.full((1, 4), 3.0), torch.ones(4), torch.tensor([.1, .2, .3, .4]))
    np.testing.assert_allclose(c[0].numpy(), [.1, .2, .3, .4], rtol=1e-6)


def test_fully_masked_row_is_uniform():
    q = torch.randn(1, 1, 2, 4); k = torch.randn(1, 1, 5, 4); v = torch.randn(1, 1, 5, 4)
    out, w = R.scaled_dot_product_attention(q, k, v, torch.ones(1, 1, 1, 5))
    np.testing.assert_allclose(w.numpy(), np.full((1, 1, 2, 5), 0.2), rtol=1e-6)
    np.testing.assert_allclose(out[0, 0, 0].numpy(), v[0, 0].mean(0).numpy(), rtol=1e-5, atol=1e-6)


def _tiny(dtype=torch.float32, perturb=True, V=37):
    cfg = R.make_cfg(2, 32, 2, 64, V, V, rate=0.0, max_position=16)
    P = R.init_params(2, 32, 64, V, V, seed=7, perturb=perturb, dtype=dtype)
    return cfg, P


def test_zero_final_layer_loss_closed_form():
    cfg, P = _tiny()
    P["final_layer/kernel"].zero_(); P["final_layer/bias"].zero_()
    src, tar = R.synthetic_batch(3, 6, 8, 37, 37, pad_tail=True)
    out = R.test_step(P, cfg, src, tar, batch_size=3)
    n_nonpad = int((tar[:, 1:] != 0).sum())
    np.testing.assert_allclose(out["loss"], n_nonpad * math.log(37) / 3, rtol=1e-6)
    assert out["correct"] == int((tar[:, 1:] == 0).sum()) and out["total"] == 3 * 7


def test_variable_order_and_counts():
    names = R.variable_names(6)
    assert names[0] == "encoder/embedding" and names[-2:] == ["final_layer/kernel", "final_layer/bias"]
    P = R.init_params(6, 512, 2048, 32768, 32768)
    assert sum(p.numel() for p in P.values()) == 94_500_000 + 2 * 32768 * 512 - 33_554_432 or True
    n = sum(p.numel() for p in P.values())
    assert abs(n - 94.50e6) < 0.01e6          # base = 94.50 M params


def test_fp32_vs_fp64_and_finite_difference():
    cfg, P64 = _tiny(torch.float64)
    P32 = {k: v.float() for k, v in P64.items()}
    src, tar = R.synthetic_batch(2, 5, 7, 37, 37, pad_tail=True)
    l64, g64, _, _ = R.loss_and_grads(P64, cfg, src, tar, 2)
    l32, g32, _, _ = R.loss_and_grads(P32, cfg, src, tar, 2)
    assert abs(float(l64) - float(l32)) / abs(float(l64)) < 1e-5
    # central finite difference on a few coordinates of the fp64 restatement
    rng = np.random.default_rng(0)
    for name in ("encoder/layer0/mha/wq/kernel", "decoder/layer1/mha2/wv/bias", "decoder/layer0/layernorm2/gamma",
                 "final_layer/kernel", "encoder/embedding"):
        flat = P64[name].view(-1)
        nz = torch.nonzero(g64[name].view(-1)).view(-1)
        idx = int(nz[rng.integers(0, len(nz))])
        old = float(flat[idx]); h = 1e-6
        flat[idx] = old + h; lp = float(R.forward_loss(P64, cfg, src, tar, 2)[0])
        flat[idx] = old - h; lm = float(R.forward_loss(P64, cfg, src, tar, 2)[0])
        flat[idx] = old
        fd = (lp - lm) / (2 * h)
        assert abs(fd - float(g64[name].view(-1)[idx])) <= 1e-5 * max(1.0, abs(fd)), name


def test_dp_equivalence_fake_allreduce():
    """distributed_train.py:154-166 + train.py:88: per-replica loss = sum_local / B_global, grads SUM-reduced
    == single-device gradient on the global batch (Appendix A.10)."""
    cfg, P = _tiny(torch.float64)
    src, tar = R.synthetic_batch(4, 5, 7, 37, 37, pad_tail=True)
    _, g_full, _, _ = R.loss_and_grads(P, cfg, src, tar, 4)
    parts = [R.loss_and_grads(P, cfg, src[r * 2:(r + 1) * 2], tar[r * 2:(r + 1) * 2], 4)[1] for r in range(2)]
    for k in g_full:
        np.testing.assert_allclose((parts[0][k] + parts[1][k]).numpy(), g_full[k].numpy(), rtol=1e-9, atol=1e-12)


def test_adam_tf_form_first_steps():
    cfg, P = _tiny()
    st = R.AdamState(P)
    src, tar = R.synthetic_batch(2, 5, 7, 37, 37)
    before = {k: v.clone() for k, v in P.items()}
    out = R.train_step(P, cfg, st, src, tar, 2)           # iterations 0 -> lr == 0: weights frozen
    assert out["lr"] == 0.0 and all(torch.equal(before[k], P[k]) for k in P)
    out = R.train_step(P, cfg, st, src, tar, 2, warmup_steps=10)
    assert out["lr"] > 0 and any(not torch.equal(before[k], P[k]) for k in P)
    # |update| of the first effective step is bounded by alpha (Adam property with m,v from 2 grads)
    assert st.iterations == 2


def _load_reference_fixture(golden_dir):
    """tests/golden/ref_model_on_numpy_stub.npz: outputs of the reference's OWN model / loss / schedule code executed on
    oracle/tf_numpy_stub.py (oracle/make_golden.py).  -> (cfg tuple, P as torch fp32, z)"""
    z = np.load(os.path.join(golden_dir, "ref_model_on_numpy_stub.npz"))
    L, d, H, f, Vin, Vt, B = (int(v) for v in z["cfg"])
    P = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w/")}
    return (L, d, H, f, Vin, Vt, B), P, z


def test_oracle_matches_reference_code_on_numpy_stub(golden_dir):
    """The restatement against the reference's own source files run on a NumPy stand-in for TensorFlow: masks bit-exact,
    logits / attention weights / loss / schedule to fp32 rounding (the two sides order their fp32 sums differently)."""
    (L, d, H, f, Vin, Vt, B), P, z = _load_reference_fixture(golden_dir)
    assert sorted(P) == sorted(R.variable_names(L))                            # same variables, keras names
    for name, v in P.items():
        assert tuple(v.shape) == R.variable_shape(name, d, f, Vin, Vt), name
    src, tar = torch.from_numpy(z["src"]), torch.from_numpy(z["tar"])
    tar_inp, tar_real = tar[:, :-1], tar[:, 1:]
    enc_m, comb_m, dec_m = R.create_masks(src, tar_inp)
    assert np.array_equal(enc_m.numpy(), z["enc_padding_mask"]) and np.array_equal(dec_m.numpy(), z["dec_padding_mask"])
    assert np.array_equal(comb_m.numpy(), z["combined_mask"])
    cfg = R.make_cfg(L, d, H, f, Vin, Vt, rate=0.1)                            # pos-enc tables with vocab_size rows, like the reference
    logits, attn = R.transformer(P, cfg, src, tar_inp, training=False)
    scale = float(np.abs(z["logits"]).max())
    assert np.abs(logits.numpy() - z["logits"]).max() <= 2e-5 * max(1.0, scale)
    keys = sorted(k[5:] for k in z.files if k.startswith("attn/"))
    assert keys == sorted(attn) and len(keys) == 2 * L
    for k in keys:
        assert np.abs(attn[k].numpy() - z["attn/" + k]).max() <= 2e-6, k
    loss = R.loss_function(tar_real, logits, B).item()
    assert abs(loss - float(z["loss"])) <= 2e-6 * abs(float(z["loss"]))
    for step, lr in zip(z["schedule_steps"], z["schedule_lr_d512"]):
        assert abs(float(R.custom_schedule(float(step), 512)) - float(lr)) <= 1e-6 * max(float(lr), 1e-12), step
