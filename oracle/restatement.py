"""CPU restatement of kuetuofa/Transformer's training hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the *oracle*: an op-for-op restatement, in torch-CPU (fp32 or fp64), of the TF2 graph the
reference builds.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path (``transformer_b200``) never does.

PARITY PARTLY PINNED.  The reference ships no tests, golden vectors or fixtures and TensorFlow
is not installable here, so nothing can be checked against the reference running on real TensorFlow: in that sense
parity is UNPINNED for TensorFlow's own kernels, whose numerics are restated from the TF 2.x public API (third-party
``tensorflow`` package, unpinned: ``assert tf.__version__.startswith('2')``, reference ``train.py:19``).  What IS
pinned, against the reference's own source files executed unmodified on a NumPy stand-in for the TF symbols they touch
(``oracle/tf_numpy_stub.py``, ``oracle/make_golden.py`` -> ``tests/golden/``):
  * ``positional_encoding`` (bit-exact, ``posenc_*.npy``);
  * ``create_masks`` (bit-exact), the whole inference-mode forward of a 2-layer ``Transformer`` -- logits and all four
    attention-weight tensors -- ``Train.loss_function`` and ``CustomSchedule`` (fp32 rounding,
    ``ref_model_on_numpy_stub.npz``, ``tests/test_oracle.py::test_oracle_matches_reference_code_on_numpy_stub``): i.e.
    everything the reference itself wrote -- wiring, mask routing, scales, residual order, head split/merge, loss
    masking and normalisation, the schedule formula;
plus the closed-form known answers of SURVEY.md Appendix D (``tests/test_oracle.py``).  Not pinned: training-mode
dropout streams, TF's autodiff (backward here is torch autograd over the restated graph) and Keras' Adam kernel.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------
# positionalencoding.py
# --------------------------------------------------------------------------------------------------


def get_angles(pos, i, d_model):
    """positionalencoding.py:4-6 -- ``pos / 10000^(2*(i//2)/d_model)``, float64 NumPy."""
    angle_rates = 1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model))
    return pos * angle_rates


def positional_encoding(position, d_model):
    """positionalencoding.py:8-23 -- [all sines || all cosines] (NOT interleaved), (1, position, d) fp32."""
    angle_rads = get_angles(np.arange(position)[:, np.newaxis], np.arange(d_model)[np.newaxis, :], d_model)
    sines = np.sin(angle_rads[:, 0::2])
    cosines = np.cos(angle_rads[:, 1::2])
    pos_encoding = np.concatenate([sines, cosines], axis=-1)[np.newaxis, ...]
    return torch.from_numpy(pos_encoding.astype(np.float32))


def create_padding_mask(seq):
    """positionalencoding.py:25-30 -- (B,1,1,S) float, 1.0 where token id == 0."""
    return (seq == 0).to(torch.float32)[:, None, None, :]


def create_look_ahead_mask(size):
    """positionalencoding.py:32-34 -- ``1 - band_part(ones, -1, 0)``: strictly upper triangular ones."""
    return 1 - torch.tril(torch.ones((size, size), dtype=torch.float32))


def create_masks(inp, tar):
    """positionalencoding.py:37-52 -- both padding masks come from ``inp``."""
    enc_padding_mask = create_padding_mask(inp)
    dec_padding_mask = create_padding_mask(inp)
    look_ahead_mask = create_look_ahead_mask(tar.shape[1])
    dec_target_padding_mask = create_padding_mask(tar)
    combined_mask = torch.maximum(dec_target_padding_mask, look_ahead_mask)
    return enc_padding_mask, combined_mask, dec_padding_mask


# --------------------------------------------------------------------------------------------------
# parameters (Keras variable order)
# --------------------------------------------------------------------------------------------------


def variable_names(num_layers):
    """``transformer.trainable_variables`` order (Keras layer-tracking order, Appendix A.8)."""
    names = ["encoder/embedding"]

    def mha(prefix):
        out = []
        for w in ("wq", "wk", "wv", "dense"):
            out += [f"{prefix}/{w}/kernel", f"{prefix}/{w}/bias"]
        return out

    def ffn(prefix):
        return [f"{prefix}/dense0/kernel", f"{prefix}/dense0/bias", f"{prefix}/dense1/kernel", f"{prefix}/dense1/bias"]

    def ln(prefix, n):
        out = []
        for j in range(1, n + 1):
            out += [f"{prefix}/layernorm{j}/gamma", f"{prefix}/layernorm{j}/beta"]
        return out

    for i in range(num_layers):
        p = f"encoder/layer{i}"
        names += mha(p + "/mha") + ffn(p + "/ffn") + ln(p, 2)
    names.append("decoder/embedding")
    for i in range(num_layers):
        p = f"decoder/layer{i}"
        names += mha(p + "/mha1") + mha(p + "/mha2") + ffn(p + "/ffn") + ln(p, 3)
    names += ["final_layer/kernel", "final_layer/bias"]
    return names


def variable_shape(name, d_model, dff, input_vocab_size, target_vocab_size):
    if name == "encoder/embedding":
        return (input_vocab_size, d_model)
    if name == "decoder/embedding":
        return (target_vocab_size, d_model)
    if name == "final_layer/kernel":
        return (d_model, target_vocab_size)
    if name == "final_layer/bias":
        return (target_vocab_size,)
    leaf = name.split("/")[-1]
    layer = name.split("/")[-2]
    if leaf in ("gamma", "beta"):
        return (d_model,)
    if layer == "dense0":
        return (d_model, dff) if leaf == "kernel" else (dff,)
    if layer == "dense1":
        return (dff, d_model) if leaf == "kernel" else (d_model,)
    return (d_model, d_model) if leaf == "kernel" else (d_model,)


def init_params(num_layers, d_model, dff, input_vocab_size, target_vocab_size, seed=42, perturb=False,
                dtype=torch.float32):
    """Keras default initialisers (Appendix A.1): Dense kernel glorot_uniform, bias zeros;
    Embedding U(-0.05,0.05); LayerNormalization gamma ones / beta zeros.
    ``perturb=True`` gives non-trivial biases / gamma / beta so tests exercise them."""
    rng = np.random.default_rng(seed)
    params = OrderedDict()
    for name in variable_names(num_layers):
        shape = variable_shape(name, d_model, dff, input_vocab_size, target_vocab_size)
        leaf = name.split("/")[-1]
        if leaf == "embedding":
            v = rng.uniform(-0.05, 0.05, shape)
        elif leaf == "kernel":
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            v = rng.uniform(-lim, lim, shape)
        elif leaf == "gamma":
            v = np.ones(shape) + (rng.normal(0, 0.1, shape) if perturb else 0)
        else:  # bias / beta
            v = rng.normal(0, 0.05, shape) if perturb else np.zeros(shape)
        params[name] = torch.tensor(np.asarray(v, dtype=np.float32)).to(dtype)
    return params


# --------------------------------------------------------------------------------------------------
# Attention.py / point_ffn.py / Encoder.py / Decoder.py / Transformer.py
# --------------------------------------------------------------------------------------------------


# Test-only switch: round every matmul OPERAND to bfloat16 (products/accumulation stay fp32), i.e. what a bf16
# tensor-core implementation of the same graph computes.  Used by tests to calibrate tolerances ("how far does bf16
# operand rounding alone move this quantity?"); the reference semantics are EMULATE_BF16 = False.
EMULATE_BF16 = False


def _op(x):
    return x.bfloat16().to(x.dtype) if EMULATE_BF16 else x


def dense(x, kernel, bias):
    """keras Dense: ``x . kernel + bias`` over the last axis (Appendix A.1)."""
    return _op(x) @ _op(kernel) + bias


def layer_norm(x, gamma, beta, eps=1e-6):
    """keras LayerNormalization(epsilon=1e-6), non-fused branch (Appendix A.4): biased variance."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def dropout(x, rate, training, generator=None):
    """keras Dropout (Appendix A.5): inverted dropout; identity when not training or rate == 0."""
    if not training or rate == 0.0:
        return x
    keep = (torch.rand(x.shape, generator=generator) >= rate).to(x.dtype)
    return x * keep / (1.0 - rate)


def scaled_dot_product_attention(q, k, v, mask):
    """Attention.py:3-34."""
    matmul_qk = _op(q) @ _op(k).transpose(-1, -2)                         # Attention.py:20
    dk = torch.tensor(float(k.shape[-1]), dtype=q.dtype)                  # Attention.py:21
    scaled_attention_logits = matmul_qk / torch.sqrt(dk)                  # Attention.py:22
    if mask is not None:
        scaled_attention_logits = scaled_attention_logits + (mask.to(q.dtype) * -1e9)   # Attention.py:25-26
    attention_weights = torch.softmax(scaled_attention_logits, dim=-1)    # Attention.py:30
    output = _op(attention_weights) @ _op(v)                              # Attention.py:32
    return output, attention_weights


def split_heads(x, batch_size, num_heads, depth):
    """Attention.py:52-57."""
    return x.reshape(batch_size, -1, num_heads, depth).permute(0, 2, 1, 3)


def multi_head_attention(P, prefix, num_heads, v, k, q, mask):
    """Attention.py:59-78 -- note the (v, k, q) argument order."""
    batch_size = q.shape[0]
    d_model = P[prefix + "/wq/kernel"].shape[1]
    depth = d_model // num_heads
    q = dense(q, P[prefix + "/wq/kernel"], P[prefix + "/wq/bias"])
    k = dense(k, P[prefix + "/wk/kernel"], P[prefix + "/wk/bias"])
    v = dense(v, P[prefix + "/wv/kernel"], P[prefix + "/wv/bias"])
    q = split_heads(q, batch_size, num_heads, depth)
    k = split_heads(k, batch_size, num_heads, depth)
    v = split_heads(v, batch_size, num_heads, depth)
    scaled_attention, attention_weights = scaled_dot_product_attention(q, k, v, mask)
    scaled_attention = scaled_attention.permute(0, 2, 1, 3)
    concat_attention = scaled_attention.reshape(batch_size, -1, d_model)
    output = dense(concat_attention, P[prefix + "/dense/kernel"], P[prefix + "/dense/bias"])
    return output, attention_weights


def point_wise_feed_forward_network(P, prefix, x):
    """point_ffn.py:3-7."""
    h = torch.relu(dense(x, P[prefix + "/dense0/kernel"], P[prefix + "/dense0/bias"]))
    return dense(h, P[prefix + "/dense1/kernel"], P[prefix + "/dense1/bias"])


def encoder_layer(P, prefix, num_heads, x, training, mask, rate, gen=None):
    """Encoder.py:19-29."""
    attn_output, _ = multi_head_attention(P, prefix + "/mha", num_heads, x, x, x, mask)
    attn_output = dropout(attn_output, rate, training, gen)
    out1 = layer_norm(x + attn_output, P[prefix + "/layernorm1/gamma"], P[prefix + "/layernorm1/beta"])
    ffn_output = point_wise_feed_forward_network(P, prefix + "/ffn", out1)
    ffn_output = dropout(ffn_output, rate, training, gen)
    return layer_norm(out1 + ffn_output, P[prefix + "/layernorm2/gamma"], P[prefix + "/layernorm2/beta"])


def encoder(P, cfg, x, training, mask, gen=None, taps=None):
    """Encoder.py:48-60."""
    d_model = cfg["d_model"]
    seq_len = x.shape[1]
    dt = P["encoder/embedding"].dtype
    x = P["encoder/embedding"][x]
    x = x * torch.sqrt(torch.tensor(float(d_model), dtype=torch.float32)).to(dt)
    x = x + cfg["pos_encoding_in"][:, :seq_len, :].to(dt)
    x = dropout(x, cfg["rate"], training, gen)
    if taps is not None:
        taps["enc_embed"] = x
    for i in range(cfg["num_layers"]):
        x = encoder_layer(P, f"encoder/layer{i}", cfg["num_heads"], x, training, mask, cfg["rate"], gen)
        if taps is not None:
            taps[f"enc_layer{i}"] = x
    return x


def decoder_layer(P, prefix, num_heads, x, enc_output, training, look_ahead_mask, padding_mask, rate, gen=None):
    """Decoder.py:25-42."""
    attn1, w1 = multi_head_attention(P, prefix + "/mha1", num_heads, x, x, x, look_ahead_mask)
    attn1 = dropout(attn1, rate, training, gen)
    out1 = layer_norm(attn1 + x, P[prefix + "/layernorm1/gamma"], P[prefix + "/layernorm1/beta"])
    attn2, w2 = multi_head_attention(P, prefix + "/mha2", num_heads, enc_output, enc_output, out1, padding_mask)
    attn2 = dropout(attn2, rate, training, gen)
    out2 = layer_norm(attn2 + out1, P[prefix + "/layernorm2/gamma"], P[prefix + "/layernorm2/beta"])
    ffn_output = point_wise_feed_forward_network(P, prefix + "/ffn", out2)
    ffn_output = dropout(ffn_output, rate, training, gen)
    out3 = layer_norm(ffn_output + out2, P[prefix + "/layernorm3/gamma"], P[prefix + "/layernorm3/beta"])
    return out3, w1, w2


def decoder(P, cfg, x, enc_output, training, look_ahead_mask, padding_mask, gen=None, taps=None):
    """Decoder.py:59-79."""
    d_model = cfg["d_model"]
    seq_len = x.shape[1]
    attention_weights = {}
    dt = P["decoder/embedding"].dtype
    x = P["decoder/embedding"][x]
    x = x * torch.sqrt(torch.tensor(float(d_model), dtype=torch.float32)).to(dt)
    x = x + cfg["pos_encoding_tgt"][:, :seq_len, :].to(dt)
    x = dropout(x, cfg["rate"], training, gen)
    if taps is not None:
        taps["dec_embed"] = x
    for i in range(cfg["num_layers"]):
        x, b1, b2 = decoder_layer(P, f"decoder/layer{i}", cfg["num_heads"], x, enc_output, training,
                                  look_ahead_mask, padding_mask, cfg["rate"], gen)
        attention_weights["decoder_layer{}_block1".format(i + 1)] = b1
        attention_weights["decoder_layer{}_block2".format(i + 1)] = b2
        if taps is not None:
            taps[f"dec_layer{i}"] = x
    return x, attention_weights


def make_cfg(num_layers, d_model, num_heads, dff, input_vocab_size, target_vocab_size, rate=0.1,
             max_position=None):
    """Static (non-trainable) state of ``Transformer.__init__`` (Transformer.py:6-16).  The reference builds
    pos-enc tables with ``position = vocab_size`` rows (Encoder.py:40, Decoder.py:53); ``max_position`` lets
    tests build fewer rows (rows are independent of the row count)."""
    assert d_model % num_heads == 0                                        # Attention.py:42
    return dict(num_layers=num_layers, d_model=d_model, num_heads=num_heads, dff=dff,
                input_vocab_size=input_vocab_size, target_vocab_size=target_vocab_size, rate=rate,
                pos_encoding_in=positional_encoding(max_position or input_vocab_size, d_model),
                pos_encoding_tgt=positional_encoding(max_position or target_vocab_size, d_model))


def transformer(P, cfg, inp, tar, training=False, masks=None, gen=None, taps=None):
    """Transformer.py:21-32 (``masks`` = the north-star six-argument form's three explicit masks)."""
    if masks is None:
        enc_padding_mask, look_ahead_mask, dec_padding_mask = create_masks(inp, tar)
    else:
        enc_padding_mask, look_ahead_mask, dec_padding_mask = masks
    enc_output = encoder(P, cfg, inp, training, enc_padding_mask, gen, taps)
    dec_output, attention_weights = decoder(P, cfg, tar, enc_output, training, look_ahead_mask,
                                            dec_padding_mask, gen, taps)
    final_output = dense(dec_output, P["final_layer/kernel"], P["final_layer/bias"])
    return final_output, attention_weights


# --------------------------------------------------------------------------------------------------
# train.py
# --------------------------------------------------------------------------------------------------


def custom_schedule(step, d_model, warmup_steps=60000):
    """train.py:21-34, evaluated in fp32 like TF.  ``step`` is Keras' 0-based ``iterations`` (Appendix A.8/A.9)."""
    step = np.float32(step)
    with np.errstate(divide="ignore"):
        arg1 = np.float32(1.0) / np.sqrt(step, dtype=np.float32)
    arg2 = step * np.float32(float(warmup_steps) ** -1.5)
    return np.float32(np.float32(1.0) / np.sqrt(np.float32(d_model))) * np.minimum(arg1, arg2)


def loss_function(real, pred, batch_size):
    """train.py:83-88 -- per-token sparse softmax xent, masked where real == 0, SUM / batch_size."""
    lse = torch.logsumexp(pred, dim=-1)
    picked = torch.gather(pred, -1, real[..., None]).squeeze(-1)
    loss_ = lse - picked
    mask = (real != 0).to(loss_.dtype)
    return (loss_ * mask).sum() * (1.0 / batch_size)


def accuracy_counts(real, pred):
    """train.py:141 SparseCategoricalAccuracy: argmax == label over ALL positions (pads included)."""
    correct = (pred.argmax(-1) == real).sum().item()
    return correct, real.numel()


class AdamState:
    """keras Adam(beta_1=0.9, beta_2=0.98, epsilon=1e-9) slots + iterations (train.py:65-66, Appendix A.8)."""

    def __init__(self, params):
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in params.items())
        self.iterations = 0


def adam_apply(params, grads, state, d_model, warmup_steps=60000, beta1=0.9, beta2=0.98, eps=1e-9):
    """TF-form Adam: lr_t = schedule(iterations) (0-based); alpha = lr_t*sqrt(1-b2^t)/(1-b1^t), t=iterations+1;
    p -= alpha * m / (sqrt(v) + eps)  (eps OUTSIDE the bias correction)."""
    t = state.iterations + 1
    lr = float(custom_schedule(state.iterations, d_model, warmup_steps))
    alpha = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    with torch.no_grad():
        for k in params:
            g = grads[k]
            state.m[k].mul_(beta1).add_(g, alpha=1 - beta1)
            state.v[k].mul_(beta2).addcmul_(g, g, value=1 - beta2)
            params[k].sub_(alpha * state.m[k] / (state.v[k].sqrt() + eps))
    state.iterations += 1
    return lr


def forward_loss(P, cfg, src, tar, batch_size, training=True, taps=None):
    """train.py:129-135: shift targets, forward, loss."""
    tar_inp = tar[:, :-1]
    tar_real = tar[:, 1:]
    predictions, _ = transformer(P, cfg, src, tar_inp, training=training, taps=taps)
    loss = loss_function(tar_real, predictions, batch_size)
    return loss, predictions, tar_real


def loss_and_grads(P, cfg, src, tar, batch_size, training=True):
    """train.py:133-137 (GradientTape -> torch autograd on the restated graph)."""
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in P.items())
    loss, predictions, tar_real = forward_loss(leaves, cfg, src, tar, batch_size, training)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=False)
    grads = OrderedDict(zip(leaves.keys(), grads))
    correct, total = accuracy_counts(tar_real, predictions.detach())
    return loss.detach(), grads, correct, total


def train_step(P, cfg, state, src, tar, batch_size, warmup_steps=60000, training=True):
    """train.py:124-141 on the restated graph.  Mutates ``P`` and ``state`` in place; returns the metrics."""
    loss, grads, correct, total = loss_and_grads(P, cfg, src, tar, batch_size, training)
    lr = adam_apply(P, grads, state, cfg["d_model"], warmup_steps)
    return dict(loss=float(loss), correct=correct, total=total, lr=lr, grads=grads)


def test_step(P, cfg, src, tar, batch_size):
    """train.py:144-156."""
    with torch.no_grad():
        loss, predictions, tar_real = forward_loss(P, cfg, src, tar, batch_size, training=False)
    correct, total = accuracy_counts(tar_real, predictions)
    return dict(loss=float(loss), correct=correct, total=total)


test_step.__test__ = False  # not a pytest test


def synthetic_batch(batch, src_len, tgt_len_plus1, vocab_in, vocab_tgt, seed=1234, pad_tail=False):
    """SURVEY.md §8(d) synthetic inputs: ids in [1, V) (no padding) -- or, with ``pad_tail``, per-row
    lengths ~U[len/2, len] with the tail zero-filled (padded_batch semantics, utils.py:154-155)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(1, vocab_in, (batch, src_len))
    tar = rng.integers(1, vocab_tgt, (batch, tgt_len_plus1))
    if pad_tail:
        for b in range(batch):
            ls = rng.integers(max(1, src_len // 2), src_len + 1)
            lt = rng.integers(max(2, tgt_len_plus1 // 2), tgt_len_plus1 + 1)
            src[b, ls:] = 0
            tar[b, lt:] = 0
    return torch.from_numpy(src.astype(np.int64)), torch.from_numpy(tar.astype(np.int64))
