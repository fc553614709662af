"""Mirror of the reference's train.py hot path: CustomSchedule (train.py:21-34) and Train (train.py:37-213) with the
same constructor, method names and metric attributes -- but ``train_step`` is one fused device step:

    zero grads -> forward (training=True) -> masked sparse-softmax-xent (+ accuracy counters, + dlogits)
    -> hand-written backward -> [data-parallel: SUM all-reduce of the gradient arena] -> fused TF-form Adam

Logits and attention weights are never materialised inside ``train_step`` (the reference discards the latter,
train.py:134).  No host synchronisation happens inside a step; metrics are read back only by ``.result()``.
"""
from __future__ import annotations

import collections
import math
import os
import time

import numpy as np
import torch

from . import _lib as L


class CustomSchedule:
    """train.py:21-34: ``rsqrt(d_model) * min(rsqrt(step), step * warmup_steps**-1.5)`` evaluated in float32."""

    def __init__(self, d_model, warmup_steps=60000):
        self.d_model = np.float32(d_model)
        self.warmup_steps = warmup_steps

    def __call__(self, step):
        step = np.float32(step)
        with np.errstate(divide="ignore"):
            arg1 = np.float32(1.0) / np.sqrt(step, dtype=np.float32)
        arg2 = step * np.float32(float(self.warmup_steps) ** -1.5)
        return np.float32(np.float32(1.0) / np.sqrt(self.d_model)) * np.minimum(arg1, arg2)


class Mean:
    """keras.metrics.Mean stand-in: running mean of the values passed to ``__call__`` (kept on the device)."""

    def __init__(self, name, device):
        self.name = name
        self._acc = torch.zeros(2, dtype=torch.float64, device=device)      # [sum, count]

    def __call__(self, value):
        self._acc[0] += value
        self._acc[1] += 1

    def add_(self, total, count):
        self._acc[0] += total
        self._acc[1] += count

    def result(self):
        s, c = self._acc.tolist()
        return s / c if c else 0.0

    def reset_states(self):
        self._acc.zero_()


class SparseCategoricalAccuracy(Mean):
    """keras.metrics.SparseCategoricalAccuracy stand-in: mean of argmax==label over ALL positions (train.py:141)."""


class Adam:
    """keras Adam(lr=schedule, beta_1, beta_2, epsilon) in TF form, fused over the flat arena."""

    def __init__(self, learning_rate, beta_1=0.9, beta_2=0.98, epsilon=1e-9):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0

    def current_lr(self):
        lr = self.learning_rate
        return float(lr(self.iterations)) if callable(lr) else float(lr)

    def schedule_args(self):
        """(d_model, warmup_steps, const_lr) for the device-side schedule (b200xf_step_params)."""
        lr = self.learning_rate
        if isinstance(lr, CustomSchedule):
            return float(lr.d_model), float(lr.warmup_steps), -1.0
        if callable(lr):
            return None                      # arbitrary Python schedule: host-computed alpha only (no graph replay)
        return 1.0, 1.0, float(lr)

    def apply_gradients_dev(self, transformer, alpha_dev, begin=0, end=None):
        """same update with alpha read from the device (written by b200xf_step_params); graph-replayable.  ``[begin, end)``:
        the arena range to update (data parallel: a range is updated as soon as its all-reduce has landed)."""
        import ctypes as C
        end = transformer._n if end is None else end

        def at(t):
            return C.c_void_p(t.data_ptr() + begin * t.element_size())
        L.call("b200xf_adam_step_dev", at(transformer.params), at(transformer.adam_m), at(transformer.adam_v), at(transformer.grads),
               at(transformer.shadow), end - begin, L.ptr(alpha_dev), self.beta_1, self.beta_2, self.epsilon, L.stream_ptr())

    def apply_gradients(self, transformer):
        t = self.iterations + 1
        lr_t = self.current_lr()                                              # evaluated at the 0-based iteration count
        alpha = lr_t * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)
        L.call("b200xf_adam_step", L.ptr(transformer.params), L.ptr(transformer.adam_m), L.ptr(transformer.adam_v),
               L.ptr(transformer.grads), L.ptr(transformer.shadow), transformer._n, alpha, self.beta_1, self.beta_2, self.epsilon,
               L.stream_ptr())
        self.iterations += 1
        return lr_t


class Train(object):
    """Train class (train.py:37-80).

    Args:
        epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, train_log_dir, test_log_dir,
        max_ckpt_keep, ckpt_path, d_model: as in the reference.  ``enable_function`` is accepted and ignored (there is
        no graph tracing; every step is already a fixed kernel sequence).
    """

    def __init__(self, epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, train_log_dir,
                 test_log_dir, max_ckpt_keep, ckpt_path, d_model, warmup_steps=60000, seed=0, cuda_graph=None, shape_bucket=0):
        self.epochs = epochs
        self.enable_function = enable_function
        self.transformer = transformer
        self.src_tokenizer = src_tokenizer
        self.tgt_tokenizer = tgt_tokenizer
        self.batch_size = batch_size                                           # GLOBAL batch size (train.py:88)
        self.learning_rate = CustomSchedule(d_model, warmup_steps)
        self.optimizer = Adam(self.learning_rate, beta_1=0.9, beta_2=0.98, epsilon=1e-9)      # train.py:65-66
        dev = transformer.device
        self.train_loss = Mean("train_loss", dev)
        self.test_loss = Mean("test_loss", dev)
        self.train_accuracy = SparseCategoricalAccuracy("train_accuracy", dev)
        self.test_accuracy = SparseCategoricalAccuracy("test_accuracy", dev)
        self.train_log_dir, self.test_log_dir = train_log_dir, test_log_dir
        self.max_ckpt_keep, self.ckpt_path = max_ckpt_keep, ckpt_path
        self._stats = torch.zeros(4, dtype=torch.float64, device=dev)          # per-step {loss_sum, correct, positions, nonpad}
        self._seed = int(seed)
        self.last_lr = 0.0
        # per-step scalars live on the device (iteration counter -> dropout seed, lr, Adam alpha), so a whole step can be
        # captured ONCE per batch shape as a CUDA graph and replayed: ~400 kernel launches become one graph launch.
        self._it_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self._alpha_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        self._lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)
        if cuda_graph is None:
            cuda_graph = os.environ.get("B200XF_CUDA_GRAPH", "1") != "0"
        self.cuda_graph = bool(cuda_graph) and self.optimizer.schedule_args() is not None
        # one captured graph per batch shape, least-recently-used eviction, ONE memory pool shared by all captures
        # (replays never overlap), so ragged text batches cannot grow the cache without bound
        # shape bucketing for ragged text batches (padded_batch pads every batch to ITS longest sentence, utils.py:154-159): source and
        # target lengths are rounded up to a multiple of ``shape_bucket`` with pad id 0, so that a handful of captured graphs serve a
        # whole epoch; the real target length goes to the device for the accuracy counters (include/b200xf.h)
        self.shape_bucket = int(shape_bucket)
        self._tvalid_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._graphs = collections.OrderedDict()
        self._graph_cache_size = int(os.environ.get("B200XF_GRAPH_CACHE", "32"))
        self._graph_pool = None
        self._shape_seen = {}
        self._graph_handle_generation = None
        self._summary_writers = {}

    # -- train.py:83-88 on materialised predictions (API parity; train_step uses the fused path) ---------------------
    def loss_function(self, real, pred):
        real = real.to(self.transformer.device, torch.int64).contiguous()
        B, T = real.shape
        V = pred.shape[-1]
        ld = (V + 3) // 4 * 4
        buf = torch.zeros((B * T, ld), dtype=torch.float32, device=real.device)
        buf[:, :V] = pred.reshape(B * T, V)
        stats = torch.zeros(4, dtype=torch.float64, device=real.device)
        L.call("b200xf_xent", L.ptr(buf), 1, ld, L.ptr(real), T, 0, T, None, 0, None, L.ptr(stats), B * T, V,
               1.0 / self.batch_size, L.stream_ptr())
        return stats[0] * (1.0 / self.batch_size)

    # -- hooks for DistributedTrain -----------------------------------------------------------------------------------
    def _reduce_gradients(self, part_index):      # called after each backward part; single device: nothing to do
        pass

    def _finish_gradients(self):
        pass

    def _backward_parts(self):
        """(part, layer) in backward-completion order (include/b200xf.h: b200xf_model_backward_part): final layer, decoder layers
        L-1..0, decoder embedding, encoder layers L-1..0, encoder embedding -- each finishes one range of the gradient arena."""
        n = self.transformer.num_layers
        return ([(0, 0)] + [(1, l) for l in range(n - 1, -1, -1)] + [(2, 0)] + [(3, l) for l in range(n - 1, -1, -1)] + [(4, 0)])

    def _apply_adam_dev(self):
        """the fused Adam update over the whole arena (DistributedTrain splits it around the last gradient bucket)"""
        self._finish_gradients()
        self.optimizer.apply_gradients_dev(self.transformer, self._alpha_dev)

    def _step_core(self, src, tar, training, seed_dev=None):
        tr = self.transformer
        src = src.to(tr.device, torch.int64).contiguous()
        tar = tar.to(tr.device, torch.int64).contiguous()
        T = tar.shape[1] - 1                                                   # tar_inp = tar[:, :-1]; tar_real = tar[:, 1:]
        self._stats.zero_()
        seed = self._seed + self.optimizer.iterations
        tr._forward(src, tar, 0, T, training, seed=seed, want_logits=False, seed_dev=seed_dev)
        L.call("b200xf_model_loss", tr._handle, L.ptr(tar), tar.shape[1], 1, 1.0 / self.batch_size, L.ptr(self._stats),
               int(training), L.stream_ptr())

    def _train_step_body(self, src, tar):
        """every device operation of one training step, in launch order (this is what gets captured into a CUDA graph)"""
        tr = self.transformer
        sched = self.optimizer.schedule_args()
        tr.zero_grads()
        if sched is not None:
            L.call("b200xf_step_params", L.ptr(self._it_dev), self._seed, sched[0], sched[1], sched[2], self.optimizer.beta_1,
                   self.optimizer.beta_2, L.ptr(self._seed_dev), L.ptr(self._alpha_dev), L.ptr(self._lr_dev), L.stream_ptr())
        self._step_core(src, tar, True, seed_dev=self._seed_dev if sched is not None else None)
        for i, (part, layer) in enumerate(self._backward_parts()):             # final layer | decoder layers | ... | encoder embedding
            L.call("b200xf_model_backward_part", tr._handle, part, layer, L.stream_ptr())
            self._reduce_gradients(i)
        if sched is not None:
            self._apply_adam_dev()
        else:
            self._finish_gradients()
            self.optimizer.apply_gradients(tr)
            self.optimizer.iterations -= 1                                     # bookkeeping happens in train_step
        self.train_loss(self._stats[0] * (1.0 / self.batch_size))
        self.train_accuracy.add_(self._stats[1], self._stats[2])

    def train_step(self, inputs):
        """One training step (train.py:124-141).  inputs: (src (B,S) int64, tar (B,T+1) int64)."""
        with torch.cuda.device(self.transformer.device):        # the library launches on the CURRENT device's stream
            return self._train_step(inputs)

    def _train_step(self, inputs):
        src, tar = inputs
        tr = self.transformer
        src = src.to(tr.device, torch.int64, non_blocking=True).contiguous()
        tar = tar.to(tr.device, torch.int64, non_blocking=True).contiguous()
        src, tar = self._bucket(src, tar)
        self.last_lr = self.optimizer.current_lr()
        key = (tuple(src.shape), tuple(tar.shape))
        if self._graph_handle_generation != getattr(tr, "_handle_generation", 0):
            self._graphs.clear()                                               # workspace was re-created: old graphs are stale
            self._graph_handle_generation = getattr(tr, "_handle_generation", 0)
        entry = self._graphs.get(key) if self.cuda_graph else None
        if self.cuda_graph and entry is None and self._shape_seen.get(key, 0) >= 1:
            entry = self._capture(key, src, tar)                               # second time this shape shows up
        if entry is None:
            self._shape_seen[key] = self._shape_seen.get(key, 0) + 1
            self._train_step_body(src, tar)                                    # eager (also the warm-up before a capture)
        else:
            self._graphs.move_to_end(key)
            graph, g_src, g_tar = entry
            if tr._shadow_dirty:               # load_ckpt / load_variables / Variable.assign since the last step: the replay
                tr.sync_shadow()               # does not go through Transformer._forward, so refresh the bf16 operands here
            g_src.copy_(src, non_blocking=True)
            g_tar.copy_(tar, non_blocking=True)
            graph.replay()
        self.optimizer.iterations += 1

    def _bucket(self, src, tar):
        """pad (B,S) / (B,T+1) up to the bucket lengths with pad id 0 and publish the real target length on the device"""
        k = self.shape_bucket
        tr = self.transformer
        if k <= 1:
            if tr._tvalid_set:
                L.load().b200xf_model_set_valid_target_len_ptr(tr._handle, None)
                tr._tvalid_set = False
            return src, tar
        S, T = src.shape[1], tar.shape[1] - 1
        S2, T2 = min(-(-S // k) * k, max(S, tr._pos_rows)), min(-(-T // k) * k, max(T, tr._pos_rows))
        if S2 > S:
            src = torch.nn.functional.pad(src, (0, S2 - S))
        if T2 > T:
            tar = torch.nn.functional.pad(tar, (0, T2 - T))
        tr._ensure_workspace(src.shape[0], S2, T2)     # (a re-created handle forgets the pointer: grow the workspace first)
        self._tvalid_dev.fill_(T)
        if not tr._tvalid_set:
            L.load().b200xf_model_set_valid_target_len_ptr(tr._handle, L.ptr(self._tvalid_dev))
            tr._tvalid_set = True
        return src, tar

    def _capture(self, key, src, tar):
        tr = self.transformer
        tr._ensure_workspace(src.shape[0], src.shape[1], tar.shape[1] - 1)
        self._graph_handle_generation = getattr(tr, "_handle_generation", 0)
        g_src, g_tar = src.clone(), tar.clone()
        graph = torch.cuda.CUDAGraph()
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        try:
            torch.cuda.synchronize(tr.device)
            with torch.cuda.graph(graph, pool=self._graph_pool):
                self._train_step_body(g_src, g_tar)
        except Exception as e:                                                 # e.g. a collective that cannot be captured
            import warnings
            warnings.warn(f"CUDA-graph capture of train_step failed ({type(e).__name__}: {e}); staying eager")
            self.cuda_graph = False
            torch.cuda.synchronize(tr.device)
            return None
        self._graphs[key] = (graph, g_src, g_tar)
        while len(self._graphs) > max(1, self._graph_cache_size):
            self._graphs.popitem(last=False)
        return self._graphs[key]

    def test_step(self, inputs):
        """One test step (train.py:144-156)."""
        src, tar = inputs
        with torch.cuda.device(self.transformer.device):
            if self.transformer._tvalid_set:                   # (shape bucketing active: this batch is not padded, every position counts)
                self._tvalid_dev.fill_(tar.shape[1] - 1)
            self._step_core(src, tar, False)
        self.test_loss(self._stats[0] * (1.0 / self.batch_size))
        self.test_accuracy.add_(self._stats[1], self._stats[2])

    # -- checkpoint / resume (train.py:77-80, 159-164, 208-211): flat arenas + keras-order manifest ---------------------
    def _ckpt_files(self):
        import glob
        import re
        files = glob.glob(os.path.join(self.ckpt_path, "ckpt-*.pt"))
        return sorted(files, key=lambda f: int(re.findall(r"ckpt-(\d+)\.pt$", f)[0]))

    def save_checkpoint(self):
        """CheckpointManager.save(): params + Adam slots + iterations, newest ``max_ckpt_keep`` files kept."""
        os.makedirs(self.ckpt_path, exist_ok=True)
        tr = self.transformer
        files = self._ckpt_files()
        n = (int(files[-1].rsplit("-", 1)[1][:-3]) + 1) if files else 1
        path = os.path.join(self.ckpt_path, f"ckpt-{n}.pt")
        torch.cuda.synchronize(tr.device)
        torch.save({"params": tr.params.cpu(), "adam_m": tr.adam_m.cpu(), "adam_v": tr.adam_v.cpu(),
                    "iterations": self.optimizer.iterations,
                    "manifest": [(v.name, tuple(v.shape)) for v in tr.trainable_variables],
                    "config": dict(num_layers=tr.num_layers, d_model=tr.d_model, num_heads=tr.num_heads, dff=tr.dff,
                                   input_vocab_size=tr.input_vocab_size, target_vocab_size=tr.target_vocab_size)}, path)
        for old in self._ckpt_files()[:-self.max_ckpt_keep]:
            os.remove(old)
        return path

    def load_ckpt(self):
        """if a checkpoint exists, restore the latest checkpoint (train.py:159-164)."""
        files = self._ckpt_files() if self.ckpt_path and os.path.isdir(self.ckpt_path) else []
        if not files:
            return False
        tr = self.transformer
        ck = torch.load(files[-1], map_location="cpu")
        assert ck["manifest"] == [(v.name, tuple(v.shape)) for v in tr.trainable_variables], "checkpoint/model mismatch"
        tr.params.copy_(ck["params"]); tr.adam_m.copy_(ck["adam_m"]); tr.adam_v.copy_(ck["adam_v"])
        tr._shadow_dirty = True
        self.optimizer.iterations = int(ck["iterations"])
        self._it_dev.fill_(self.optimizer.iterations)
        print('Latest checkpoint restored!!')
        return True

    def _write_summaries(self, epoch, values=None):
        """train.py:200-206 / distributed_train.py:108-114: per-epoch 'loss' and 'accuracy' scalars of the train and test
        metrics, written as TensorBoard event files under train_log_dir / test_log_dir (no-op when the dirs are None)."""
        if values is None:
            values = [self.train_loss.result(), self.train_accuracy.result(), self.test_loss.result(), self.test_accuracy.result()]
        for log_dir, loss, acc in ((self.train_log_dir, values[0], values[1]), (self.test_log_dir, values[2], values[3])):
            if not log_dir:
                continue
            w = self._summary_writers.get(log_dir)
            if w is None:
                from torch.utils.tensorboard import SummaryWriter      # event-file writer only; nothing on the hot path
                w = self._summary_writers[log_dir] = SummaryWriter(log_dir)
            w.add_scalar('loss', loss, epoch)
            w.add_scalar('accuracy', acc, epoch)
            w.flush()

    def predict(self, input_sentence, max_steps=10, use_cache=True):
        """Greedy inference (train.py:91-121): start token, 10 decode steps, no EOS stop.  The reference re-runs the whole model
        (encoder included) for every generated token; with ``use_cache`` the encoder and the cross-attention K/V run once and each
        step feeds one token through the decoder against a key/value cache (same tokens out; ``use_cache=False`` keeps the
        reference's full recomputation).  (The reference indexes ``input_sentence[0]``, i.e. expects a 1-element batch of strings.)"""
        if not isinstance(input_sentence, str):
            input_sentence = input_sentence[0]
        tr = self.transformer
        start_token = [self.src_tokenizer.vocab_size]
        end_token = [self.src_tokenizer.vocab_size + 1]
        ids = start_token + self.src_tokenizer.encode(input_sentence) + end_token
        encoder_input = torch.tensor([ids], dtype=torch.int64, device=tr.device)
        output = torch.tensor([start_token], dtype=torch.int64, device=tr.device)
        if use_cache:
            tr._ensure_workspace(1, len(ids), max(max_steps, tr._cfg.max_tgt_len))
            tr.decode_begin(encoder_input)
            for i in range(max_steps):
                logits = tr.decode_step(output[:, i])                          # position i attends to the cached positions 0..i
                predicted_id = logits.argmax(dim=-1, keepdim=True).to(torch.int64)
                output = torch.cat([output, predicted_id], dim=-1)
            return output[0]
        for _ in range(max_steps):
            predictions, _ = tr.call((encoder_input, output), False)
            predicted_id = predictions[:, -1:, :].argmax(dim=-1).to(torch.int64)
            output = torch.cat([output, predicted_id], dim=-1)
        return output[0]

    def training_loop(self, train_dataset, test_dataset):
        """Custom training and testing loop (train.py:167-213): same cadence and print template."""
        template = 'Epoch {}  Loss {:.4f} Accuracy {:.4f}, Test Loss {:.4f}, Test Accuracy {:.4f}'
        for epoch in range(self.epochs):
            self.train_loss.reset_states()
            self.test_loss.reset_states()
            self.train_accuracy.reset_states()
            self.test_accuracy.reset_states()
            start = time.time()
            counter = 0
            from .utils import prefetch_to_device
            for src, tgt in prefetch_to_device(train_dataset, self.transformer.device):       # next batch's H2D copy under this step
                self.train_step((src, tgt))
                counter += 1
                if (counter + 1) % 100 == 0:
                    for t_src, t_tgt in test_dataset:
                        self.test_step((t_src, t_tgt))
                    print(template.format(epoch + 1, self.train_loss.result(), self.train_accuracy.result() * 100,
                                          self.test_loss.result(), self.test_accuracy.result() * 100))
            self._write_summaries(epoch)
            if (epoch + 1) % 5 or (epoch + 1 == self.epochs) == 0:      # the reference's save condition, verbatim (train.py:208)
                if self.ckpt_path:
                    ckpt_save_path = self.save_checkpoint()
                    print('Saving checkpoint for epoch {} at {}'.format(epoch + 1, ckpt_save_path))
            torch.cuda.synchronize()
            print('Time taken for {} epoch: {} secs\n'.format(epoch + 1, (time.time() - start)))


def run_main(argv):
    """train.py:216-220."""
    del argv
    from . import utils
    kwargs = utils.flags_dict()
    del kwargs['per_replica_batch_size']
    main(**kwargs)


def main(epochs, enable_function, buffer_size, batch_size, d_model, dff, num_heads, src_vocab_file, tgt_vocab_file, dataset_path,
         dropout_rate, num_layers, sequence_length, ckpt_path, max_ckpt_keep):
    """train.py:223-246 on the B200 drop-in (text pipeline: utils.load_dataset stand-in)."""
    import datetime
    from . import utils
    from .Transformer import Transformer
    current_time = datetime.datetime.now().strftime("%Y%m%d-%H%M%S")
    train_log_dir = 'logs/gradient_tape/' + current_time + '/train'
    test_log_dir = 'logs/gradient_tape/' + current_time + '/test'
    train_dataset, test_dataset, src_tokenizer, tgt_tokenizer = utils.load_dataset(
        dataset_path, sequence_length, src_vocab_file, tgt_vocab_file, batch_size, buffer_size)
    input_vocab_size = src_tokenizer.vocab_size + 2
    target_vocab_size = tgt_tokenizer.vocab_size + 2
    # the workspace is sized once for the largest batch the pipeline can emit (no re-creation -- and no stale CUDA graphs --
    # when a longer batch shows up); test sentences are not length-filtered (utils.py:153 filters train only)
    longest = max([sequence_length + 2] + [max(s.shape[1], t.shape[1]) for s, t in test_dataset])
    transformer = Transformer(num_layers, d_model, num_heads, dff, input_vocab_size, target_vocab_size, dropout_rate,
                              pos_rows=max(longest, 512), max_batch=batch_size, max_src_len=longest, max_tgt_len=longest)
    print('create training object')
    train_obj = Train(epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, train_log_dir, test_log_dir,
                      max_ckpt_keep, ckpt_path, d_model, shape_bucket=8)
    train_obj.training_loop(train_dataset, test_dataset)
    train_obj.load_ckpt()
    input_sentence = 'he go to school'
    print(tgt_tokenizer.decode(train_obj.predict(input_sentence).tolist()))
    return train_obj


if __name__ == '__main__':
    from absl import app
    from . import utils as _utils
    _utils.transformer_flags()
    app.run(run_main)
