"""Host-side mirror of the reference's positionalencoding.py (same names, same argument meaning).

The sinusoid table is computed exactly like the reference (float64 NumPy, [all sines || all cosines],
positionalencoding.py:4-23) and uploaded once; it stays resident on the device.  The mask helpers return the same
float masks (1.0 = masked) for callers of the six-argument ``Transformer`` form; the fast path never needs them --
the attention kernel derives the canonical masks from compact key-padding bytes and a causal flag.
"""
import numpy as np
import torch


def get_angles(pos, i, d_model):
    """positionalencoding.py:4-6."""
    angle_rates = 1 / np.power(10000, (2 * (i // 2)) / np.float32(d_model))
    return pos * angle_rates


def positional_encoding(position, d_model, device=None):
    """positionalencoding.py:8-23 -> torch fp32 tensor of shape (1, position, d_model)."""
    angle_rads = get_angles(np.arange(position)[:, np.newaxis], np.arange(d_model)[np.newaxis, :], d_model)
    sines = np.sin(angle_rads[:, 0::2])
    cosines = np.cos(angle_rads[:, 1::2])
    pos_encoding = np.concatenate([sines, cosines], axis=-1)[np.newaxis, ...]
    t = torch.from_numpy(pos_encoding.astype(np.float32))
    return t.to(device) if device is not None else t


def create_padding_mask(seq):
    """positionalencoding.py:25-30: (batch, 1, 1, seq_len) float32, 1.0 where the token id is 0."""
    return (seq == 0).to(torch.float32)[:, None, None, :]


def create_look_ahead_mask(size, device=None):
    """positionalencoding.py:32-34: strictly-upper-triangular ones (size, size)."""
    return 1 - torch.tril(torch.ones((size, size), dtype=torch.float32, device=device))


def create_masks(inp, tar):
    """positionalencoding.py:37-52: (enc_padding_mask, combined look-ahead|target-padding mask, dec_padding_mask);
    both padding masks are built from ``inp``."""
    enc_padding_mask = create_padding_mask(inp)
    dec_padding_mask = create_padding_mask(inp)
    look_ahead_mask = create_look_ahead_mask(tar.shape[1], device=tar.device)
    dec_target_padding_mask = create_padding_mask(tar)
    combined_mask = torch.maximum(dec_target_padding_mask, look_ahead_mask)
    return enc_padding_mask, combined_mask, dec_padding_mask
