"""Generate tests/golden/*.npy from the REAL reference code (test infrastructure; run in the build container).

Only ``positionalencoding.get_angles/positional_encoding`` (positionalencoding.py:4-23) is executable without
TensorFlow: its body is NumPy up to a final ``tf.cast``.  We import the reference module under a stub ``tensorflow``
exposing exactly the two symbols it touches at import/call time (``tf.float32``, ``tf.cast``) and save its outputs.
/root/reference is read at generation time only; tests read the committed .npy files.

    python oracle/make_golden.py        # needs /root/reference
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    tf = types.ModuleType("tensorflow")
    tf.float32 = np.float32
    tf.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    sys.modules["tensorflow"] = tf
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
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
