"""Generate tests/golden/* from the REAL reference code (test infrastructure; run in the build container).

TensorFlow is not installable here, so the reference's source files are imported and executed UNMODIFIED on top of
``oracle/tf_numpy_stub.py`` -- a NumPy stand-in, installed as ``tensorflow``, for exactly the TF / Keras primitives they
touch (documented semantics, float32).  Two kinds of fixtures come out:

* ``posenc_*.npy``: ``positionalencoding.positional_encoding`` (positionalencoding.py:4-23) -- its body is NumPy up to
  a final ``tf.cast``;
* ``ref_model_on_numpy_stub.npz``: a 2-layer (d_model 64, 2 heads, dff 128) model built by the reference's ``Transformer`` class (Transformer.py,
  Encoder.py, Decoder.py, Attention.py, point_ffn.py) with seeded weights: its masks (``create_masks``), logits and
  attention weights in inference mode, ``Train.loss_function`` (train.py:83-88) on those logits, and
  ``CustomSchedule`` (train.py:21-34) values.  This pins everything the reference itself wrote (wiring, masks,
  scales, loss normalisation, schedule); TF's own kernels are restated by the stub and therefore NOT pinned.

/root/reference is read at generation time only; tests read the committed files.

    python oracle/make_golden.py        # needs /root/reference
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _rnd(rng, shape, scale):
    return (rng.standard_normal(shape) * scale).astype(np.float32)


def model_goldens():
    """Run the reference's own Transformer / loss / schedule code on the NumPy stand-in."""
    import Transformer as ref_transformer      # the reference's files (REF is first on sys.path)
    import train as ref_train
    import positionalencoding as ref_pe
    tf = sys.modules["tensorflow"]

    L, d, H, f, Vin, Vt, B = 2, 64, 2, 128, 37, 41, 3      # head depth 32: a shape the CUDA path runs too
    model = ref_transformer.Transformer(L, d, H, f, Vin, Vt, rate=0.1)
    rng = np.random.default_rng(20260921)
    src = rng.integers(1, Vin, (B, 11)); src[0, 8:] = 0; src[2, 5:] = 0
    tar = rng.integers(1, Vt, (B, 10)); tar[1, 6:] = 0
    tar_inp, tar_real = tar[:, :-1], tar[:, 1:]                     # train.py:130-131
    model((src, tar_inp), False)                                    # first call: keras layers build their variables

    W = {}                                                          # oracle / keras-order names -> seeded values

    def dense(name, layer):
        layer.kernel = _rnd(rng, layer.kernel.shape, layer.kernel.shape[0] ** -0.5); W[name + "/kernel"] = layer.kernel
        layer.bias = _rnd(rng, layer.bias.shape, 0.1); W[name + "/bias"] = layer.bias

    def ln(name, layer):
        layer.gamma = (1 + _rnd(rng, layer.gamma.shape, 0.1)).astype(np.float32); W[name + "/gamma"] = layer.gamma
        layer.beta = _rnd(rng, layer.beta.shape, 0.1); W[name + "/beta"] = layer.beta

    def mha(name, m):
        for part in ("wq", "wk", "wv", "dense"):
            dense(f"{name}/{part}", getattr(m, part))

    def ffn(name, seq):
        dense(name + "/dense0", seq.layers[0]); dense(name + "/dense1", seq.layers[1])

    enc, dec = model.encoder, model.decoder
    enc.embedding.embeddings = _rnd(rng, enc.embedding.embeddings.shape, 0.3); W["encoder/embedding"] = enc.embedding.embeddings
    for i, lay in enumerate(enc.enc_layers):
        p = f"encoder/layer{i}"
        mha(p + "/mha", lay.mha); ffn(p + "/ffn", lay.ffn); ln(p + "/layernorm1", lay.layernorm1); ln(p + "/layernorm2", lay.layernorm2)
    dec.embedding.embeddings = _rnd(rng, dec.embedding.embeddings.shape, 0.3); W["decoder/embedding"] = dec.embedding.embeddings
    for i, lay in enumerate(dec.dec_layers):
        p = f"decoder/layer{i}"
        mha(p + "/mha1", lay.mha1); mha(p + "/mha2", lay.mha2); ffn(p + "/ffn", lay.ffn)
        ln(p + "/layernorm1", lay.layernorm1); ln(p + "/layernorm2", lay.layernorm2); ln(p + "/layernorm3", lay.layernorm3)
    dense("final_layer", model.final_layer)

    logits, attn = model((src, tar_inp), False)
    enc_mask, combined_mask, dec_mask = ref_pe.create_masks(src, tar_inp)
    trainer = object.__new__(ref_train.Train)                       # Train.__init__ needs checkpoints / summary writers: bypass it
    trainer.loss_object = tf.keras.losses.SparseCategoricalCrossentropy(from_logits=True, reduction="none")   # train.py:61-62
    trainer.batch_size = B
    loss = trainer.loss_function(tar_real, logits)
    steps = np.array([0, 1, 2, 100, 59999, 60000, 60001, 1000000], dtype=np.float32)
    sched = ref_train.CustomSchedule(512)
    with np.errstate(divide="ignore"):
        lrs = np.array([sched(s) for s in steps], dtype=np.float32)

    out = {"cfg": np.array([L, d, H, f, Vin, Vt, B]), "src": src, "tar": tar, "logits": logits.astype(np.float32),
           "enc_padding_mask": enc_mask.astype(np.float32), "combined_mask": combined_mask.astype(np.float32),
           "dec_padding_mask": dec_mask.astype(np.float32), "loss": np.float32(loss), "schedule_steps": steps,
           "schedule_lr_d512": lrs}
    for k, v in attn.items():
        out["attn/" + k] = v.astype(np.float32)
    for k, v in W.items():
        out["w/" + k] = v
    np.savez_compressed(os.path.join(OUT, "ref_model_on_numpy_stub.npz"), **out)
    print("reference-on-stub: logits", logits.shape, "loss", float(loss), "lr", lrs)


def main():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import tf_numpy_stub
    tf_numpy_stub.install()                     # before ANY reference module is imported
    sys.path.insert(0, REF)
    import positionalencoding as ref  # the reference's own file

    os.makedirs(OUT, exist_ok=True)
    for pos, d in ((64, 8), (50, 128), (2048, 512), (300, 1024)):
        pe = ref.positional_encoding(pos, d)
        assert pe.shape == (1, pos, d) and pe.dtype == np.float32
        if pos * d > 64 * 1024:   # keep fixtures small: strided sample + checksums
            sample = pe[0, ::37, ::13].copy()
            np.save(os.path.join(OUT, f"posenc_{pos}x{d}_sample37x13.npy"), sample)
            np.save(os.path.join(OUT, f"posenc_{pos}x{d}_sums.npy"),
                    np.array([pe.astype(np.float64).sum(), np.abs(pe.astype(np.float64)).sum()]))
        else:
            np.save(os.path.join(OUT, f"posenc_{pos}x{d}.npy"), pe)
    model_goldens()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
