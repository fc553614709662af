"""Mirror of the reference's utils.py CLI surface (absl flags with the same names and defaults, utils.py:17-62) plus
the synthetic-token data source used on the GPU box (the reference's text pipeline, utils.py:65-161, is host-side
TensorFlow/tfds I/O and out of scope -- SURVEY.md §8f)."""
from __future__ import annotations

import numpy as np
import torch
from absl import flags

FLAGS = flags.FLAGS


def transformer_flags():
    """utils.py:17-33, names and defaults verbatim; extra flags for the synthetic data source."""
    flags.DEFINE_string('dataset_path', 'data/', ' Dataset Folder')
    flags.DEFINE_integer('buffer_size', 100000, 'Shuffle buffer size')
    flags.DEFINE_string('src_vocab_file', 'src_vocab.txt', 'Source Vocabulary file')
    flags.DEFINE_string('tgt_vocab_file', 'tgt_vocab.txt', 'Target Vocabulary file')
    flags.DEFINE_integer('sequence_length', 50, 'Maxinum number of words in a sequence')
    flags.DEFINE_integer('epochs', 4, 'Number of Epochs')
    flags.DEFINE_integer('batch_size', 64, 'Batch Size')
    flags.DEFINE_integer('per_replica_batch_size', 16, 'Batch Size')
    flags.DEFINE_integer('num_layers', 4, 'Nnmber of Encoder/Decoder Stack')
    flags.DEFINE_integer('d_model', 512, 'Output dimesion of all sublayers including Embedding layer')
    flags.DEFINE_integer('dff', 1024, 'Dimetionality of inner layer')
    flags.DEFINE_integer('num_heads', 4, 'Number of Attention Head')
    flags.DEFINE_boolean('enable_function', True, 'Enable Function')
    flags.DEFINE_integer('max_ckpt_keep', 5, 'Maximum Number of Checkpoint to keep')
    flags.DEFINE_string('ckpt_path', 'model_dist', 'Checkpoint Path')
    flags.DEFINE_float('dropout_rate', 0.1, 'Dropout Probability')
    # additions (synthetic data)
    flags.DEFINE_integer('vocab', 32768, 'Synthetic vocabulary size (source and target)')
    flags.DEFINE_integer('steps_per_epoch', 20, 'Synthetic batches per epoch')
    flags.DEFINE_integer('warmup_steps', 60000, 'CustomSchedule warm-up steps (train.py:22)')


def flags_dict():
    """utils.py:36-62."""
    return {
        'dataset_path': FLAGS.dataset_path, 'enable_function': FLAGS.enable_function, 'buffer_size': FLAGS.buffer_size,
        'src_vocab_file': FLAGS.src_vocab_file, 'tgt_vocab_file': FLAGS.tgt_vocab_file, 'batch_size': FLAGS.batch_size,
        'per_replica_batch_size': FLAGS.per_replica_batch_size, 'sequence_length': FLAGS.sequence_length,
        'epochs': FLAGS.epochs, 'num_layers': FLAGS.num_layers, 'd_model': FLAGS.d_model, 'dff': FLAGS.dff,
        'num_heads': FLAGS.num_heads, 'max_ckpt_keep': FLAGS.max_ckpt_keep, 'ckpt_path': FLAGS.ckpt_path,
        'dropout_rate': FLAGS.dropout_rate,
    }


def synthetic_batch(batch, src_len, tgt_len_plus1, vocab_in, vocab_tgt, seed=1234, pad_tail=False):
    """ids uniform in [1, V) (no padding); ``pad_tail`` zero-fills a random tail per row like
    padded_batch (utils.py:154-155).  int64, host tensors."""
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


class SyntheticDataset:
    """Iterable of (src, tgt) int64 batches of a fixed shape, pinned on the host; ``tgt`` has seq_len + 1 columns
    (train_step shifts it, train.py:130-131)."""

    def __init__(self, steps, batch, seq_len, vocab, seed=1234, pad_tail=False, pin=True):
        self.batches = []
        for i in range(steps):
            s, t = synthetic_batch(batch, seq_len, seq_len + 1, vocab, vocab, seed + i, pad_tail)
            if pin and torch.cuda.is_available():
                s, t = s.pin_memory(), t.pin_memory()
            self.batches.append((s, t))

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)
