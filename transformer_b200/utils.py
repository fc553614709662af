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


# ------------------------------------------------------------------------------------------------------------------
# Text pipeline stand-in (SURVEY.md §8f "next" #1).  The reference's utils.read_data / _load_or_create_tokenier /
# load_dataset (utils.py:65-161) are tf.data + tensorflow_datasets' SubwordTextEncoder -- neither library exists here
# and the subword algorithm lives in the unvendored tfds package (parity unpinned).  This keeps the SAME function
# names, arguments, return values and batching semantics on plain Python/torch, with a deterministic word-level
# tokenizer standing in for SubwordTextEncoder.  Host-side I/O only; nothing here runs on the GPU.
# ------------------------------------------------------------------------------------------------------------------
import collections
import os
import random


class WordTokenizer:
    """Stand-in for tfds SubwordTextEncoder: ids 1..vocab_size-1 are words by descending corpus frequency, id 0 is the
    padding id (never produced), the last id is <unk>.  Same surface: ``vocab_size``, ``encode``, ``decode``,
    ``save_to_file`` / ``load_from_file`` (file suffix '.subwords' like tfds, so the reference's existence check,
    utils.py:96-97, keeps working)."""

    def __init__(self, words):
        self.words = list(words)
        self.index = {w: i + 1 for i, w in enumerate(self.words)}
        self.vocab_size = len(self.words) + 2          # + pad (0) + <unk> (last)

    @classmethod
    def build_from_corpus(cls, corpus, target_vocab_size=2 ** 15):
        cnt = collections.Counter()
        for line in corpus:
            cnt.update(line.split())
        words = [w for w, _ in sorted(cnt.items(), key=lambda kv: (-kv[1], kv[0]))[:max(0, target_vocab_size - 2)]]
        return cls(words)

    def encode(self, s):
        if isinstance(s, bytes):
            s = s.decode("utf-8")
        unk = self.vocab_size - 1
        return [self.index.get(w, unk) for w in s.split()]

    def decode(self, ids):
        return " ".join(self.words[i - 1] if 1 <= i <= len(self.words) else "<unk>" for i in ids if i != 0)

    def save_to_file(self, prefix):
        with open(prefix + ".subwords", "w", encoding="utf-8") as f:
            f.write("\n".join(self.words))

    @classmethod
    def load_from_file(cls, prefix):
        with open(prefix + ".subwords", encoding="utf-8") as f:
            return cls([w for w in f.read().split("\n") if w])


def read_data(src_file, tgt_file):
    """utils.py:65-80: zip of the two line files -> list of (src_line, tgt_line)."""
    with open(src_file, encoding="utf-8") as fs, open(tgt_file, encoding="utf-8") as ft:
        return list(zip((l.rstrip("\n") for l in fs), (l.rstrip("\n") for l in ft)))


def _load_or_create_tokenier(dataset, src_vocab_file, tgt_vocab_file):
    """utils.py:83-111 (name kept, typo included)."""
    if not os.path.exists(src_vocab_file + '.subwords') and not os.path.exists(tgt_vocab_file + '.subwords'):
        src_tokenizer = WordTokenizer.build_from_corpus((s for s, _ in dataset), target_vocab_size=2 ** 15)
        src_tokenizer.save_to_file(src_vocab_file)
        tgt_tokenizer = WordTokenizer.build_from_corpus((t for _, t in dataset), target_vocab_size=2 ** 15)
        tgt_tokenizer.save_to_file(tgt_vocab_file)
    else:
        print("Vocabulary exists Loading...")
        src_tokenizer = WordTokenizer.load_from_file(src_vocab_file)
        tgt_tokenizer = WordTokenizer.load_from_file(tgt_vocab_file)
    return src_tokenizer, tgt_tokenizer


def _padded_batches(examples, batch_size):
    """tf.data padded_batch(batch_size, ([-1],[-1])): ragged per batch, pad id 0, last batch partial (utils.py:154-159)."""
    out = []
    for i in range(0, len(examples), batch_size):
        chunk = examples[i:i + batch_size]
        ls = max(len(s) for s, _ in chunk)
        lt = max(len(t) for _, t in chunk)
        src = torch.zeros((len(chunk), ls), dtype=torch.int64)
        tgt = torch.zeros((len(chunk), lt), dtype=torch.int64)
        for r, (s, t) in enumerate(chunk):
            src[r, :len(s)] = torch.tensor(s, dtype=torch.int64)
            tgt[r, :len(t)] = torch.tensor(t, dtype=torch.int64)
        out.append((src, tgt))
    return out


def prefetch_to_device(batches, device):
    """tf.data's prefetch (utils.py:157-160) on this side of the boundary: yields device-resident (src, tgt) pairs, with the NEXT
    batch's host -> device copy (from pinned memory, on its own stream) in flight while the current one trains."""
    dev = torch.device(device)
    if dev.type != "cuda":
        yield from batches
        return
    copy_stream = torch.cuda.Stream(device=dev)

    def stage(b):
        with torch.cuda.stream(copy_stream):
            out = tuple(t if t.is_cuda else (t if t.is_pinned() else t.pin_memory()).to(dev, non_blocking=True) for t in b)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return out, ev
    it = iter(batches)
    nxt = None
    for b in it:
        cur, nxt = nxt, stage(b)
        if cur is not None:
            torch.cuda.current_stream(dev).wait_event(cur[1])
            for t in cur[0]:
                t.record_stream(torch.cuda.current_stream(dev))
            yield cur[0]
    if nxt is not None:
        torch.cuda.current_stream(dev).wait_event(nxt[1])
        for t in nxt[0]:
            t.record_stream(torch.cuda.current_stream(dev))
        yield nxt[0]


def make_synthetic_corpus(dataset_path, lines=10000, words=6000, seed=7):
    """The reference trains on data/src-train.txt -> data/tgt-train.txt (10000 sentence pairs shipped with its repo).  Those
    files are not part of THIS repo; when ``dataset_path`` has no train files, a synthetic parallel corpus of the same size
    (Zipf-distributed word ids, 3..40 words per line, the target a deterministic word-by-word 'translation' with a length
    jitter) is written there so that the default command line runs anywhere."""
    os.makedirs(dataset_path, exist_ok=True)
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, words + 1)
    p /= p.sum()
    with open(os.path.join(dataset_path, 'src-train.txt'), 'w', encoding='utf-8') as fs, \
            open(os.path.join(dataset_path, 'tgt-train.txt'), 'w', encoding='utf-8') as ft:
        for _ in range(lines):
            n = int(rng.integers(3, 41))
            ids = rng.choice(words, size=n, p=p)
            fs.write(" ".join(f"s{i}" for i in ids) + "\n")
            keep = ids[rng.random(n) > 0.08]
            ft.write(" ".join(f"t{(int(i) * 7 + 3) % words}" for i in keep) + (" t0" if rng.random() < 0.3 else "") + "\n")


def load_dataset(dataset_path, sequence_length, src_vocab_file, tgt_vocab_file, batch_size, buffer_size, seed=0,
                 holdout=0.05):
    """utils.py:114-161.  The reference opens src-test.txt / tgt-test.txt, which it does not ship (SURVEY §0 trap 6):
    when they are missing, the last ``holdout`` fraction of the train files is held out as the test set."""
    if not os.path.exists(os.path.join(dataset_path, 'src-train.txt')):
        print('{}src-train.txt not found: writing a synthetic parallel corpus there'.format(dataset_path))
        make_synthetic_corpus(dataset_path)
    pairs = read_data(os.path.join(dataset_path, 'src-train.txt'), os.path.join(dataset_path, 'tgt-train.txt'))
    st, tt = os.path.join(dataset_path, 'src-test.txt'), os.path.join(dataset_path, 'tgt-test.txt')
    if os.path.exists(st) and os.path.exists(tt):
        test_pairs = read_data(st, tt)
    else:
        n_test = max(1, int(len(pairs) * holdout))
        pairs, test_pairs = pairs[:-n_test], pairs[-n_test:]
    src_tokenizer, tgt_tokenizer = _load_or_create_tokenier(pairs, src_vocab_file, tgt_vocab_file)

    def encode(src, tgt):      # start / end ids = vocab_size, vocab_size + 1 (utils.py:137-143)
        return ([src_tokenizer.vocab_size] + src_tokenizer.encode(src) + [src_tokenizer.vocab_size + 1],
                [tgt_tokenizer.vocab_size] + tgt_tokenizer.encode(tgt) + [tgt_tokenizer.vocab_size + 1])

    train = [encode(s, t) for s, t in pairs]
    train = [(s, t) for s, t in train if len(s) <= sequence_length and len(t) <= sequence_length]   # train only (utils.py:153)
    random.Random(seed).shuffle(train)                       # shuffle(buffer_size) with buffer >= dataset
    test = [encode(s, t) for s, t in test_pairs]
    return _padded_batches(train, batch_size), _padded_batches(test, batch_size), src_tokenizer, tgt_tokenizer
