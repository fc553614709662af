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
        """make_dataset_iterator semantics (distributed_train.py:151-152): split the global batch over the replicas.  A full
        batch splits evenly (the divisibility of ``batch_size`` itself is checked once, distributed_train.py:154-158);
        the partial LAST batch of an epoch (padded_batch keeps it, utils.py:154-159) is split as evenly as possible --
        the first ``b % n`` replicas get one more row -- and a replica left without rows gets an EMPTY shard, which
        ``DistributedTrain`` turns into a zero-gradient step so that every rank still joins the all-reduce."""
        n, r = self.num_replicas_in_sync, self.rank
        b = src.shape[0]
        q, rem = divmod(b, n)
        lo = r * q + min(r, rem)
        hi = lo + q + (1 if r < rem else 0)
        return src[lo:hi], tar[lo:hi]

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
    """Contiguous ranges of the flat gradient arena, in backward-completion order (one per backward part: final layer, every
    decoder layer, decoder embedding, every encoder layer, encoder embedding), and their async SUM all-reduce.  Works on any
    torch.distributed backend (NCCL on GPUs; gloo in the CPU tests of this host logic)."""

    def __init__(self, ranges, group=None):
        self.ranges = list(ranges)          # [(begin, end)] per bucket, in the order the backward finishes them
        self.group = group
        self.pending = []

    def reduce_stage(self, arena, stage):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        b, e = self.ranges[stage]
        self.pending.append(dist.all_reduce(arena[b:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self, keep_last=0):
        """make the current stream wait for the pending collectives, except the newest ``keep_last`` (NCCL: no host block)"""
        n = len(self.pending) - keep_last
        for w in self.pending[:max(0, n)]:
            w.wait()
        self.pending = self.pending[max(0, n):]


def arena_stage_ranges(transformer):
    """the three coarse ranges (final layer | decoder + embedding | encoder + embedding)"""
    lib = L.load()
    out = []
    for stage in range(3):
        b, e = C.c_int64(), C.c_int64()
        L.check(lib.b200xf_model_grad_range(C.byref(transformer._cfg), stage, C.byref(b), C.byref(e)), "b200xf_model_grad_range")
        out.append((b.value, e.value))
    return out


def arena_part_ranges(transformer, parts):
    """arena range finished by each backward part (b200xf_model_grad_range_part)"""
    lib = L.load()
    out = []
    for part, layer in parts:
        b, e = C.c_int64(), C.c_int64()
        L.check(lib.b200xf_model_grad_range_part(C.byref(transformer._cfg), part, layer, C.byref(b), C.byref(e)), "b200xf_model_grad_range_part")
        out.append((b.value, e.value))
    return out


class DistributedTrain(Train):
    """DistributedTrain (distributed_train.py:26-47): ``batch_size`` is the GLOBAL batch (it scales the loss),
    ``per_replica_batch_size`` is accepted and -- exactly like the reference -- not used by the step."""

    def __init__(self, epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, per_replica_batch_size,
                 train_log_dir, test_log_dir, max_ckpt_keep, ckpt_path, d_model, warmup_steps=60000, seed=0, shape_bucket=0):
        Train.__init__(self, epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, train_log_dir,
                       test_log_dir, max_ckpt_keep, ckpt_path, d_model, warmup_steps=warmup_steps, seed=seed, shape_bucket=shape_bucket)
        self.per_replica_batch_size = per_replica_batch_size
        # one gradient bucket per backward part: 2L + 3 all-reduces per step, each issued the moment its range is final, so that
        # only the LAST one (the encoder embedding, 4 x V x d bytes) can still be in flight when the backward ends -- and Adam on
        # everything else runs underneath it (see _apply_adam_dev)
        parts = self._backward_parts()
        ranges = arena_part_ranges(transformer, parts)
        # Bucket granularity.  Every all-reduce in flight takes SMs away from the persistent GEMMs (whose tile schedule assumes all of
        # them), so MORE buckets are not better: measured on 2 B200 (profiles/r02_dp_buckets.txt) one bucket per layer costs more than it
        # hides.  Default "coarse": final layer | decoder layers + decoder embedding | encoder layers | encoder embedding -- the last
        # one is the only bucket that cannot overlap backward work, and it is the small one.  B200XF_DP_BUCKETS=layer: one per part.
        mode = os.environ.get("B200XF_DP_BUCKETS", "coarse")
        if mode == "layer":
            groups = [[i] for i in range(len(parts))]
        else:
            groups = [[i for i, (p, _) in enumerate(parts) if p == 0], [i for i, (p, _) in enumerate(parts) if p in (1, 2)],
                      [i for i, (p, _) in enumerate(parts) if p == 3], [i for i, (p, _) in enumerate(parts) if p == 4]]
        self._bucket_after_part = {g[-1]: k for k, g in enumerate(groups)}          # part index that completes bucket k
        # hardware-scheduled chunks of GEMM tiles while a collective is in flight: measured SLOWER than persistent grids on 2 B200
        # (23.7-23.9 vs 23.2 ms/step, profiles/r02_dp_buckets.txt), so off unless asked for
        self._chunks_per_unit = int(os.environ.get("B200XF_DP_GEMM_CHUNKS", "0"))
        self.buckets = GradientBuckets([(min(ranges[i][0] for i in g), max(ranges[i][1] for i in g)) for g in groups])
        rank = dist.get_rank() if dist.is_initialized() else 0
        self._seed = int(seed) * 1000003 + rank * 7919            # independent dropout streams per replica
        if dist.is_initialized() and dist.get_world_size() > 1:
            # initial mirror of the variables (created under strategy.scope(), distributed_train.py:141,160)
            dist.broadcast(transformer.params, src=0)
            transformer._shadow_dirty = True

    def _reduce_gradients(self, part_index):
        k = self._bucket_after_part.get(part_index)
        if k is not None:
            self.buckets.reduce_stage(self.transformer.grads, k)
            if self.buckets.pending and self._chunks_per_unit > 0:
                # from here to the end of the backward an all-reduce shares the GPU with the GEMMs: hardware-scheduled chunks of tiles
                # instead of persistent grids (include/b200xf.h: b200xf_gemm_set_chunks_per_unit)
                L.load().b200xf_gemm_set_chunks_per_unit(self._chunks_per_unit)

    def _with_rows(self, inputs, step, loss_metric, acc_metric):
        """run ``step`` on this replica's shard; an EMPTY shard (partial last batch smaller than the replica count) runs on
        one all-padding row instead: label 0 everywhere => zero loss, zero gradients (train.py:85-87 mask), so the
        collective still sees every rank; the row's accuracy-counter contribution is taken back out."""
        src, tar = inputs
        if src.shape[0] > 0:
            return step((src, tar))
        dummy = (torch.zeros((1, max(1, src.shape[1])), dtype=torch.int64), torch.zeros((1, max(2, tar.shape[1])), dtype=torch.int64))
        step(dummy)
        loss_metric._acc[1] -= 1                              # Mean(): this step added one (zero) sample
        acc_metric.add_(-self._stats[1], -self._stats[2])

    def train_step(self, inputs):
        return self._with_rows(inputs, lambda x: Train.train_step(self, x), self.train_loss, self.train_accuracy)

    def test_step(self, inputs):
        return self._with_rows(inputs, lambda x: Train.test_step(self, x), self.test_loss, self.test_accuracy)

    def _finish_gradients(self):
        self.buckets.wait()

    def _apply_adam_dev(self):
        L.load().b200xf_gemm_set_chunks_per_unit(0)                # the forward of the next step runs alone: persistent grids again
        return self._apply_adam_split()

    def _apply_adam_split(self):
        """identical Adam on every replica, split around the last bucket: the ranges whose all-reduce has landed are updated while
        the encoder-embedding bucket (the one nothing else can overlap) is still on the wire"""
        tr = self.transformer
        b_last, e_last = self.buckets.ranges[-1]
        if len(self.buckets.pending) < 2 or b_last != 0:
            return Train._apply_adam_dev(self)
        self.buckets.wait(keep_last=1)
        self.optimizer.apply_gradients_dev(tr, self._alpha_dev, e_last, tr._n)
        self.buckets.wait()
        self.optimizer.apply_gradients_dev(tr, self._alpha_dev, 0, e_last)

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
                        self._print_progress(template, epoch, train_steps, step_timer, strategy)     # distributed_train.py:95-98
                except StopIteration:
                    break
            self._print_progress(template, epoch, train_steps, step_timer, strategy, final=True)     # distributed_train.py:103-106
            if strategy.rank == 0:
                self._write_summaries(epoch, self._epoch_metrics)                                                    # distributed_train.py:108-114
                if (epoch + 1) % 5 or ((epoch + 1) == self.epochs) == 0:   # the reference's save condition, verbatim (distributed_train.py:116)
                    if self.ckpt_path:
                        ckpt_save_path = self.save_checkpoint()             # replicas are mirrored: rank 0's arena is THE model
                        print('Saving checkpoint for epoch {} at {}'.format(epoch + 1, ckpt_save_path))
                print('Time taken for {} epoch: {} secs\n'.format(epoch + 1, (time.time() - start_time)))

    def _print_progress(self, template, epoch, train_steps, step_timer, strategy, final=False):
        """metric read-out across replicas (sync-on-read SUM, every rank takes part) and rank 0's progress line.  The
        accumulators keep LOCAL sums, so the reduction works on copies."""
        vals = []
        for m in (self.train_loss, self.train_accuracy, self.test_loss, self.test_accuracy):
            acc = m._acc.clone()
            if dist.is_initialized() and dist.get_world_size() > 1:
                dist.all_reduce(acc, op=dist.ReduceOp.SUM)
            s, c = acc.tolist()
            vals.append(s / c if c else 0.0)
        if final:
            self._epoch_metrics = vals
        if strategy.rank == 0:
            print(template.format(epoch + 1, train_steps, vals[0], vals[1] * 100, vals[2], vals[3] * 100, time.time() - step_timer))


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


# ----------------------------------------------------------------------------------------------------------------------
# entry point: distributed_train.py:22-23,124-179  (`python -m transformer_b200.distributed_train --num_gpu=N`)
# ----------------------------------------------------------------------------------------------------------------------
def _define_num_gpu_flag():
    from absl import flags
    if 'num_gpu' not in flags.FLAGS:
        flags.DEFINE_integer('num_gpu', 4, 'Number of GPUs to use')                     # distributed_train.py:23


def run_main(argv):
    """distributed_train.py:124-128."""
    del argv
    from absl import flags
    from . import utils
    kwargs = utils.flags_dict()
    kwargs.update({'num_gpu': flags.FLAGS.num_gpu})
    main(**kwargs)


def _relaunch_one_process_per_gpu(num_gpu):
    """The reference drives N GPUs from ONE process (MirroredStrategy, distributed_train.py:137-138).  Here each GPU gets its
    own process: the command line is re-executed under torch.distributed.run with --nproc-per-node=num_gpu on this node
    (rendezvous on 127.0.0.1); every child then finds WORLD_SIZE in its environment and runs ``main`` as one replica."""
    import subprocess
    import sys
    port = os.environ.get("MASTER_PORT", str(29500 + os.getpid() % 2000))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={num_gpu}", "--master-addr", "127.0.0.1",
           "--master-port", port, "-m", "transformer_b200.distributed_train"] + sys.argv[1:]
    return subprocess.call(cmd)


def main(epochs, enable_function, buffer_size, batch_size, d_model, dff, num_heads, src_vocab_file, tgt_vocab_file, dataset_path,
         dropout_rate, num_layers, sequence_length, per_replica_batch_size, ckpt_path, max_ckpt_keep, num_gpu=1):
    """distributed_train.py:131-174 on the B200 drop-in."""
    import datetime
    from . import utils
    from .Transformer import Transformer
    if num_gpu > 1 and "WORLD_SIZE" not in os.environ:
        if num_gpu > torch.cuda.device_count():
            raise ValueError('num_gpu ({}) exceeds the visible GPUs ({})'.format(num_gpu, torch.cuda.device_count()))
        rc = _relaunch_one_process_per_gpu(num_gpu)
        if rc:
            raise SystemExit(rc)
        return None
    rank, local, world = init_process_group_from_env("nccl")
    current_time = datetime.datetime.now().strftime("%Y%m%d-%H%M%S")
    train_log_dir = 'logs/gradient_tape/' + current_time + '/train'
    test_log_dir = 'logs/gradient_tape/' + current_time + '/test'
    devices = ['/device:GPU:{}'.format(i) for i in range(world)]
    strategy = MirroredStrategy(devices)
    num_replicas = strategy.num_replicas_in_sync
    with strategy.scope():
        # every replica reads the same files in the same (seeded) order; rank 0 creates the vocabulary files first
        if rank == 0:
            data = utils.load_dataset(dataset_path, sequence_length, src_vocab_file, tgt_vocab_file, batch_size, buffer_size)
        if world > 1:
            dist.barrier()
        if rank != 0:
            data = utils.load_dataset(dataset_path, sequence_length, src_vocab_file, tgt_vocab_file, batch_size, buffer_size)
        train_dataset, test_dataset, src_tokenizer, tgt_tokenizer = data
        input_vocab_size = src_tokenizer.vocab_size + 2
        target_vocab_size = tgt_tokenizer.vocab_size + 2
        local_batch_size = split_replica_batch(batch_size, num_replicas)                 # distributed_train.py:154-158
        longest = max([sequence_length + 2] + [max(s.shape[1], t.shape[1]) for s, t in test_dataset])
        transformer = Transformer(num_layers, d_model, num_heads, dff, input_vocab_size, target_vocab_size, dropout_rate,
                                  device=torch.device(f"cuda:{local}"), pos_rows=max(longest, 512), max_batch=local_batch_size + 1,
                                  max_src_len=longest, max_tgt_len=longest)
        if rank == 0:
            print('create training object')
        train_obj = DistributedTrain(epochs, enable_function, transformer, src_tokenizer, tgt_tokenizer, batch_size, local_batch_size,
                                     train_log_dir, test_log_dir, max_ckpt_keep, ckpt_path, d_model, shape_bucket=8)
        if rank == 0:
            print('Training.....')
        train_obj.training_loop(train_dataset, test_dataset, strategy)
        if rank == 0:
            print(tgt_tokenizer.decode(train_obj.predict(['he goes to school']).tolist()))
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        train_obj._graphs.clear()                 # captured collectives hold the communicator: drop the graphs first
        dist.destroy_process_group()
    return train_obj


if __name__ == '__main__':
    from absl import app
    from . import utils as _utils
    _utils.transformer_flags()
    _define_num_gpu_flag()
    app.run(run_main)
