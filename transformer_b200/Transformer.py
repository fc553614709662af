"""``Transformer`` -- drop-in for the reference's keras model (Transformer.py:5-32) on one B200.

Same constructor, same two call forms, same return value:

    model = Transformer(num_layers, d_model, num_heads, dff, input_vocab_size, target_vocab_size, rate=0.1)
    logits, attention_weights = model((inp, tar), training=False)                         # reference form (train.py:134,152)
    logits, attention_weights = model(inp, tar, training, enc_padding_mask,
                                      look_ahead_mask, dec_padding_mask)                  # six-argument form

All arithmetic runs in libb200xf.so's sm_100a kernels (ctypes, C ABI); torch only owns the device buffers.
There is no CPU path: constructing the model without a CUDA device or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np
import torch

from . import _lib as L
from .positionalencoding import create_masks, positional_encoding


class LazyAttentionWeights(Mapping):
    """``attention_weights`` dict of Decoder.call (Decoder.py:75-76): keys 'decoder_layer{i}_block{1,2}', values fp32
    (batch, heads, T, Sk).  The reference materialises 2L such tensors every step and ``train_step`` throws them away
    (train.py:134); here each one is recomputed from the saved Q/K/log-sum-exp only when it is read.  Valid until the
    model's next forward."""

    def __init__(self, model, num_layers, B, H, T, S, generation):
        self._m, self._L, self._B, self._H, self._T, self._S, self._gen = model, num_layers, B, H, T, S, generation

    def __iter__(self):
        for i in range(self._L):
            yield "decoder_layer{}_block1".format(i + 1)
            yield "decoder_layer{}_block2".format(i + 1)

    def __len__(self):
        return 2 * self._L

    def __getitem__(self, key):
        if not key.startswith("decoder_layer") or "_block" not in key:
            raise KeyError(key)
        layer = int(key[len("decoder_layer"):key.index("_block")]) - 1
        block = int(key[key.index("_block") + 6:])
        if not (0 <= layer < self._L and block in (1, 2)):
            raise KeyError(key)
        if self._gen != self._m._generation:
            raise RuntimeError("attention_weights of an earlier forward are no longer available")
        sk = self._T if block == 1 else self._S
        out = torch.empty((self._B, self._H, self._T, sk), dtype=torch.float32, device=self._m.device)
        with torch.cuda.device(self._m.device):
            L.call("b200xf_model_attention_weights", self._m._handle, layer, block, L.ptr(out), L.stream_ptr())
        return out


class Variable:
    """A named (possibly strided) fp32 view into the parameter arena -- stands in for a tf.Variable."""

    def __init__(self, name, value, grad, model=None):
        self.name, self.value, self.grad = name, value, grad
        self._model = model                 # assign() must invalidate the model's bf16 operand copy

    @property
    def shape(self):
        return tuple(self.value.shape)

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def assign(self, t):
        self.value.copy_(torch.as_tensor(t).to(self.value.device, self.value.dtype))
        if self._model is not None:
            self._model._shadow_dirty = True


def _round_up(x, m):
    return (x + m - 1) // m * m


class Transformer:
    def __init__(self, num_layers, d_model, num_heads, dff, input_vocab_size, target_vocab_size, rate=0.1, *,
                 device=None, seed=42, max_batch=0, max_src_len=0, max_tgt_len=0, pos_rows=None):
        assert d_model % num_heads == 0                                       # Attention.py:42
        if not torch.cuda.is_available():
            raise L.B200xfError("transformer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        L.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_layers, self.d_model, self.num_heads, self.dff = num_layers, d_model, num_heads, dff
        self.input_vocab_size, self.target_vocab_size, self.rate = input_vocab_size, target_vocab_size, float(rate)
        # Encoder.py:40 / Decoder.py:53 build tables with `position = vocab_size` rows; rows are independent of the
        # row count, so only the first pos_rows are kept resident (default: all of them, like the reference).
        self._pos_rows = int(pos_rows) if pos_rows else max(input_vocab_size, target_vocab_size)
        self.pos_encoding_in = positional_encoding(min(self._pos_rows, input_vocab_size), d_model, self.device)
        self.pos_encoding_tgt = positional_encoding(min(self._pos_rows, target_vocab_size), d_model, self.device)
        self._pos_rows = min(self.pos_encoding_in.shape[1], self.pos_encoding_tgt.shape[1])
        self._cfg = L.ModelConfig(num_layers, d_model, num_heads, dff, input_vocab_size, target_vocab_size, 1, 1, 1,
                                  self._pos_rows, self.rate, 1e-6)
        n = L.load().b200xf_model_param_count(C.byref(self._cfg))
        if n <= 0:
            L.check(1, "b200xf_model_param_count")
        self._n = int(n)
        with torch.cuda.device(self.device):
            self.params = torch.zeros(self._n, dtype=torch.float32, device=self.device)
            self.grads = torch.zeros(self._n, dtype=torch.float32, device=self.device)
            self.adam_m = torch.zeros(self._n, dtype=torch.float32, device=self.device)
            self.adam_v = torch.zeros(self._n, dtype=torch.float32, device=self.device)
            self.shadow = torch.zeros(self._n, dtype=torch.bfloat16, device=self.device)
        self._handle = None
        self._workspace = None
        self._generation = 0
        self._last_shape = None
        self.trainable_variables = self._make_variables()
        self._init_weights(seed)
        self._ensure_workspace(max(1, max_batch), max(1, max_src_len), max(1, max_tgt_len))

    # ------------------------------------------------------------------ parameters
    def _make_variables(self):
        out = []
        lib = L.load()
        name = C.create_string_buffer(128)
        off, rows, cols, ld = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        i = 0
        while lib.b200xf_model_variable(C.byref(self._cfg), i, name, 128, C.byref(off), C.byref(rows), C.byref(cols), C.byref(ld)) == 0:
            def view(arena):
                if cols.value == 0:
                    return arena[off.value:off.value + rows.value]
                return torch.as_strided(arena, (rows.value, cols.value), (ld.value, 1), off.value)
            out.append(Variable(name.value.decode(), view(self.params), view(self.grads), self))
            i += 1
        return out

    def _init_weights(self, seed):
        """keras default initialisers: Dense kernel glorot_uniform / bias zeros,
        Embedding U(-0.05, 0.05), LayerNormalization gamma ones / beta zeros."""
        rng = np.random.default_rng(seed)
        for v in self.trainable_variables:
            leaf = v.name.split("/")[-1]
            if leaf == "embedding":
                a = rng.uniform(-0.05, 0.05, v.shape)
            elif leaf == "kernel":
                lim = math.sqrt(6.0 / (v.shape[0] + v.shape[1]))
                a = rng.uniform(-lim, lim, v.shape)
            elif leaf == "gamma":
                a = np.ones(v.shape)
            else:
                a = np.zeros(v.shape)
            v.assign(torch.from_numpy(np.asarray(a, dtype=np.float32)))
        self._shadow_dirty = True

    def load_variables(self, named):
        """copy {keras-order name: array/tensor} into the arena (e.g. the oracle's parameter dict)."""
        byname = {v.name: v for v in self.trainable_variables}
        for k, t in named.items():
            byname[k].assign(t)
        self._shadow_dirty = True

    def variables_dict(self):
        return OrderedDict((v.name, v.value.detach().clone()) for v in self.trainable_variables)

    def sync_shadow(self):
        """refresh the bf16 operand copy after the fp32 weights were written from outside"""
        with torch.cuda.device(self.device):
            L.call("b200xf_model_sync_shadow", self._handle, L.stream_ptr())
        self._shadow_dirty = False

    # ------------------------------------------------------------------ workspace / handle
    def _ensure_workspace(self, B, S, T):
        c = self._cfg
        if self._handle is not None and B <= c.max_batch and S <= c.max_src_len and T <= c.max_tgt_len:
            return
        if max(S, T) > self._pos_rows:
            raise ValueError(f"sequence length {max(S, T)} exceeds the positional-encoding table ({self._pos_rows} rows)")
        torch.cuda.synchronize(self.device)
        if self._handle is not None:
            L.load().b200xf_model_destroy(self._handle)
            self._handle = None
            self._workspace = None
        c.max_batch, c.max_src_len, c.max_tgt_len = max(B, c.max_batch), max(S, c.max_src_len), max(T, c.max_tgt_len)
        nbytes = L.load().b200xf_model_workspace_bytes(C.byref(c))
        if nbytes <= 0:
            L.check(1, "b200xf_model_workspace_bytes")
        with torch.cuda.device(self.device):
            raw = torch.empty(int(nbytes) + 1024, dtype=torch.uint8, device=self.device)
        shift = (-raw.data_ptr()) % 1024
        self._workspace = raw[shift:shift + int(nbytes)]
        h = C.c_void_p()
        with torch.cuda.device(self.device):          # the library creates its second stream / events on the current device
          L.call("b200xf_model_create", C.byref(c), L.ptr(self.params), L.ptr(self.grads), L.ptr(self.adam_m), L.ptr(self.adam_v),
               L.ptr(self.shadow), L.ptr(self.pos_encoding_in), L.ptr(self.pos_encoding_tgt), L.ptr(self._workspace), C.byref(h))
        self._handle = h
        self._tvalid_set = False                       # (train.Train: device-side valid target length for shape bucketing)
        self._handle_generation = getattr(self, "_handle_generation", 0) + 1      # captured CUDA graphs of older handles are stale
        self._shadow_dirty = True

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                L.load().b200xf_model_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ forward
    def _forward(self, *a, **kw):
        with torch.cuda.device(self.device):          # library calls launch on the current device's current stream
            return self._forward_on_device(*a, **kw)

    def _forward_on_device(self, inp, tar, tar_offset, T, training, seed=0, masks=None, want_logits=True, seed_dev=None):
        """inp int64 (B,S); tar int64 (B, >= tar_offset+T) -- the decoder input is tar[:, tar_offset:tar_offset+T].
        ``seed_dev``: optional device uint64 holding the dropout seed (graph-replayable steps); else ``seed``."""
        if inp.dtype != torch.int64 or tar.dtype != torch.int64:
            inp, tar = inp.to(torch.int64), tar.to(torch.int64)
        inp = inp.to(self.device).contiguous()
        tar = tar.to(self.device).contiguous()
        B, S = inp.shape
        assert tar.shape[0] == B and tar.shape[1] >= tar_offset + T
        self._ensure_workspace(B, S, T)
        if self._shadow_dirty:
            self.sync_shadow()
        enc_m = la_m = dec_m = None
        la_bs = 0
        if masks is not None:
            enc_m, la_m, dec_m = masks
            enc_m = torch.broadcast_to(enc_m.to(self.device, torch.float32), (B, 1, 1, S)).contiguous()
            dec_m = torch.broadcast_to(dec_m.to(self.device, torch.float32), (B, 1, 1, S)).contiguous()
            la_m = la_m.to(self.device, torch.float32)
            if la_m.dim() == 2:
                la_m = la_m[None, None]
            assert la_m.shape[-2:] == (T, T), "look_ahead_mask must be (.., T, T)"
            if la_m.shape[0] not in (1, B):
                raise ValueError("look_ahead_mask batch dimension must be 1 or batch_size")
            la_m = la_m.reshape(la_m.shape[0], T, T).contiguous()
            la_bs = T * T if la_m.shape[0] == B and B > 1 else 0      # batch stride (0 = broadcast over the batch)
        logits = None
        ld = 0
        if want_logits:
            ld = _round_up(self.target_vocab_size, 4)
            logits = torch.empty((B * T, ld), dtype=torch.float32, device=self.device)
        self._generation += 1
        self._last_shape = (B, S, T)
        self._keepalive = (inp, tar, enc_m, la_m, dec_m, seed_dev)      # the library keeps raw pointers until the next forward
        L.load().b200xf_model_set_seed_ptr(self._handle, L.ptr(seed_dev))
        L.call("b200xf_model_forward", self._handle, L.ptr(inp), L.ptr(tar), tar.shape[1], tar_offset, B, S, T, int(bool(training)),
               int(seed), L.ptr(enc_m), L.ptr(la_m), la_bs, L.ptr(dec_m), L.ptr(logits), ld, L.stream_ptr())
        weights = LazyAttentionWeights(self, self.num_layers, B, self.num_heads, T, S, self._generation)
        if logits is not None:
            logits = logits[:, :self.target_vocab_size].reshape(B, T, self.target_vocab_size)
        return logits, weights

    def call(self, input_, training=False, *extra, seed=0):
        """Transformer.call (Transformer.py:21-32).  Accepts ``call((inp, tar), training)`` and
        ``call(inp, tar, training, enc_padding_mask, look_ahead_mask, dec_padding_mask)``."""
        if isinstance(input_, (tuple, list)) and not extra:
            inp, tar = input_
            return self._forward(inp, tar, 0, tar.shape[1], training, seed)
        if len(extra) == 4:                      # (inp, tar, training, enc_mask, look_ahead_mask, dec_mask)
            inp, tar = input_, training
            training, enc_m, la_m, dec_m = extra
            canonical = create_masks(inp.to(self.device), tar.to(self.device))
            same = all(m is not None and tuple(torch.broadcast_shapes(m.shape, c.shape)) == tuple(c.shape) and
                       torch.equal(torch.broadcast_to(m.to(self.device, torch.float32), c.shape), c)
                       for m, c in zip((enc_m, la_m, dec_m), canonical))
            masks = None if same else (enc_m, la_m, dec_m)       # canonical masks -> in-kernel fast path
            return self._forward(inp, tar, 0, tar.shape[1], training, seed, masks=masks)
        raise TypeError("Transformer expects ((inp, tar), training) or (inp, tar, training, enc_mask, look_ahead_mask, dec_mask)")

    def __call__(self, input_, training=False, *extra, **kw):
        return self.call(input_, training, *extra, **kw)

    # ------------------------------------------------------------------ incremental decoding (used by Train.predict)
    def decode_begin(self, inp):
        """run the encoder on ``inp`` (B,S) int64 and the cross-attention K/V projections of every decoder layer once"""
        inp = inp.to(self.device, torch.int64).contiguous()
        B, S = inp.shape
        self._ensure_workspace(B, S, self._cfg.max_tgt_len)
        with torch.cuda.device(self.device):
            if self._shadow_dirty:
                self.sync_shadow()
            self._generation += 1
            self._keepalive = (inp,)
            L.call("b200xf_model_decode_begin", self._handle, L.ptr(inp), B, S, L.stream_ptr())
        self._decode_batch = B

    def decode_step(self, tokens):
        """feed ONE decoder-input token per sequence (``tokens`` (B,) int64); returns the logits (B, V) fp32 of that position.
        The self-attention keys/values of all earlier positions come from the cache."""
        tokens = tokens.to(self.device, torch.int64).contiguous().reshape(-1)
        assert tokens.numel() == self._decode_batch
        ld = _round_up(self.target_vocab_size, 4)
        logits = torch.empty((self._decode_batch, ld), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.call("b200xf_model_decode_step", self._handle, L.ptr(tokens), L.ptr(logits), ld, L.stream_ptr())
        return logits[:, :self.target_vocab_size]

    # ------------------------------------------------------------------ training-step pieces (used by train.Train)
    def zero_grads(self):
        self.grads.zero_()

    def tap(self, name):
        """debug/parity: fp32 copy of an internal activation of the last forward ('enc_layer0', 'dec_embed', ...)."""
        B, S, T = self._last_shape
        rows = B * (S if name.startswith("enc") else T)
        out = torch.empty((rows, self.d_model), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.call("b200xf_model_tap", self._handle, name.encode(), L.ptr(out), L.stream_ptr())
        return out.reshape(B, -1, self.d_model)
