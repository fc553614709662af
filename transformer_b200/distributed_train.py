"""Mirror of the reference's distributed_train.py: data-parallel training of the same step.

The reference runs ONE process with tf.distribute.MirroredStrategy over N GPUs (distributed_train.py:137-138): each
replica runs ``train_step`` on batch_size/N sequences, loss still divided by the GLOBAL batch size (train.py:88 via
distributed_train.py:164-166), gradients SUM-all-reduced inside ``apply_gradients``, identical Adam everywhere.

Here: one process per GPU (torchrun / torch.distributed, NCCL over NVLink 5 / NVSwitch).  The gradient arena is ONE
flat fp32 buffer; backward finishes it in three contiguous ranges (final layer -> decoder -> encoder) and each range
is all-reduced (SUM) asynchronously on NCCL's stream while the next backward stage runs on the compute stream.  No
other sharding; no collective anywhere else on the data path.  ``MirroredStrategy`` below is the tiny stand-in for the
tf.distribute object the reference passes around (``num_replicas_in_sync``, ``experimental_run``, ``scope``).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import time

import torch
import torch.distributed as dist

from . import _lib as L
from .train import Train


class MirroredStrategy:
    """Stand-in for tf.distribute.MirroredStrategy (distributed_train.py:137-141): wraps the default process group."""

    def __init__(self, devices=None):
        self.devices = devices
        self.initialized = dist.is_available() and dist.is_initialized()
        self.num_replicas_in_sync = dist.get_world_size() if self.initialized else 1
        self.rank = dist.get_rank() if self.initialized else 0

    @contextlib.contextmanager
    def scope(self):
        yield self

    def shard(self, src, tar):
        """make_dataset_iterator semantics (distributed_train.py:151-152): split the global batch evenly over replicas."""
        n, r = self.num_replicas_in_sync, self.rank
        b = src.shape[0]
        if b % n:
            raise ValueError('Batch size ({}) must be divisible by the number of replicas ({})'.format(b, n))
        per = b // n
        return src[r * per:(r + 1) * per], tar[r * per:(r + 1) * per]

    def experimental_run(self, fn, iterator):
        """strategy.experimental_run(step_fn, iterator) (distributed_train.py:59,62): run fn on this replica's shard."""
        src, tar = next(iterator)
        return fn(self.shard(src, tar))


def split_replica_batch(batch_size, num_replicas):
    """distributed_train.py:154-158."""
    local_batch_size, remainder = divmod(batch_size, num_replicas)
    if remainder:
        raise ValueError('Batch size ({}) must be divisible by the number of replicas ({})'.format(batch_size, num_replicas))
    return local_batch_size


class GradientBuckets:
    """The three contiguous ranges of the flat gradient arena, in backward-completion order, and their async SUM
    all-reduce.  Works on any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests of this host logic)."""

    def __init__(self, ranges, group=None):
        self.ranges = list(ranges)          # [(begin, end)] for stage 0, 1, 2
        self.group = group
        self.pending = []

    def reduce_stage(self, arena, stage):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        b, e = self.ranges[stage]
        self.pending.append(dist.all_reduce(arena[b:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self):
        for w in self.pending:
            w.wait()              # NCCL: makes the current stream wait on the collective; no host block
        self.pending = []


def arena_stage_ranges(transformer):
    lib = L.load()
    out = []
    for stage in range(3):
        b, e = C.c_int64(), C.c_int64()
        L.check(lib.b200xf_model_grad_range(C.byref(transformer._cfg), stage, C.byref(b), C.byref(e)), "b200xf_model_grad_range")
        out.append((b.value, e.value))
    return out


class DistributedTrain(Train):
    """DistributedTrain (distributed_train.py:26-47): ``batch_size`` is the GLOBAL batch (it scales the loss),
    ``per_replica_batch_size`` is accepted and -- exactly like the reference -- not used by the step."""

    def __init__(self, epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, per_replica_batch_size,
                 train_log_dir, test_log_dir, max_ckpt_keep, ckpt_path, d_model, warmup_steps=60000, seed=0):
        Train.__init__(self, epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, train_log_dir,
                       test_log_dir, max_ckpt_keep, ckpt_path, d_model, warmup_steps=warmup_steps, seed=seed)
        self.per_replica_batch_size = per_replica_batch_size
        self.buckets = GradientBuckets(arena_stage_ranges(transformer))
        rank = dist.get_rank() if dist.is_initialized() else 0
        self._seed = int(seed) * 1000003 + rank * 7919            # independent dropout streams per replica
        if dist.is_initialized() and dist.get_world_size() > 1:
            # initial mirror of the variables (created under strategy.scope(), distributed_train.py:141,160)
            dist.broadcast(transformer.params, src=0)
            transformer._shadow_dirty = True

    def _reduce_gradients(self, stage):
        self.buckets.reduce_stage(self.transformer.grads, stage)

    def _finish_gradients(self):
        self.buckets.wait()

    def reduce_metrics(self):
        """sync-on-read SUM aggregation of the metric accumulators across replicas (train.py:140-141 under the
        strategy).  Note the reference's reported train_loss is the MEAN over replicas of per-replica losses, i.e.
        global_loss / N -- reproduced here."""
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return
        for m in (self.train_loss, self.test_loss, self.train_accuracy, self.test_accuracy):
            dist.all_reduce(m._acc, op=dist.ReduceOp.SUM)

    def training_loop(self, train_iterator, test_iterator, strategy):
        """Custom training and testing loop (distributed_train.py:50-121): StopIteration plays tf.errors.OutOfRangeError."""
        template = ('Epoch {} Steps {} Train Loss {:.4f} Accuracy {:.4f}, Valid Loss {:.4f}, Test Accuracy {:.4f} '
                    'Times takes {: .4f} seconds')
        for epoch in range(self.epochs):
            for m in (self.train_loss, self.test_loss, self.train_accuracy, self.test_accuracy):
                m.reset_states()
            start_time = time.time()
            train_steps = 0
            it = iter(train_iterator)
            step_timer = time.time()
            while True:
                step_timer = time.time()
                try:
                    strategy.experimental_run(self.train_step, it)
                    train_steps += 1
                    if train_steps % 500 == 0:
                        for t in iter(test_iterator):
                            self.test_step(strategy.shard(*t))
                            break                                   # the reference's `if test_steps % 100: break` exits after 1 step
                except StopIteration:
                    break
            self.reduce_metrics()
            if strategy.rank == 0:
                print(template.format(epoch + 1, train_steps, self.train_loss.result(), self.train_accuracy.result() * 100,
                                      self.test_loss.result(), self.test_accuracy.result() * 100, time.time() - step_timer))
                print('Time taken for {} epoch: {} secs\n'.format(epoch + 1, (time.time() - start_time)))


def init_process_group_from_env(backend="nccl"):
    """one process per GPU, launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": torch.device(f"cuda:{local}")} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, **kw)
    return int(os.environ.get("RANK", "0")), local, world
