"""CPU-only checks of the host side: the C-ABI library loads and exports every declared symbol, the parameter-arena
layout matches keras' trainable_variables order (oracle), schedule / flags / replica-split logic, and the N>1
gradient-bucket all-reduce on world_size-2 gloo.  No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import restatement as R
from transformer_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    header = open(os.path.join(ROOT, "include", "b200xf.h")).read()
    declared = set(re.findall(r"\b(b200xf_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200xf_stream_t"}
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/b200xf.h but not exported"
    assert set(L.EXPORTED_SYMBOLS) == declared, set(L.EXPORTED_SYMBOLS) ^ declared
    assert lib.b200xf_version() >= 100 and lib.b200xf_launch_count() == 0


def _cfg(Lyr=2, d=128, H=2, f=256, Vin=1003, Vtgt=1009):
    return L.ModelConfig(Lyr, d, H, f, Vin, Vtgt, 4, 64, 64, 128, 0.1, 1e-6)


def _variables(cfg):
    lib = L.load()
    out = []
    name = C.create_string_buffer(128)
    off, rows, cols, ld = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
    i = 0
    while lib.b200xf_model_variable(C.byref(cfg), i, name, 128, C.byref(off), C.byref(rows), C.byref(cols), C.byref(ld)) == 0:
        out.append((name.value.decode(), off.value, rows.value, cols.value, ld.value))
        i += 1
    return out


def test_arena_layout_matches_keras_variable_order():
    cfg = _cfg()
    vs = _variables(cfg)
    assert [v[0] for v in vs] == R.variable_names(2)
    n = L.load().b200xf_model_param_count(C.byref(cfg))
    used = np.zeros(n, dtype=np.int32)
    for name, off, rows, cols, ld in vs:
        shape = R.variable_shape(name, 128, 256, 1003, 1009)
        assert shape == ((rows, cols) if cols else (rows,)), name
        assert off % 4 == 0 and (cols == 0 or ld % 4 == 0)
        for r in range(rows if cols else 1):
            a = off + r * ld
            used[a:a + (cols if cols else rows)] += 1
    assert used.max() == 1                                   # views never overlap
    # every logical parameter is inside exactly one of the three backward stages, and stages tile the arena
    lib = L.load()
    rng = []
    for s in range(3):
        b, e = C.c_int64(), C.c_int64()
        assert lib.b200xf_model_grad_range(C.byref(cfg), s, C.byref(b), C.byref(e)) == 0
        rng.append((b.value, e.value))
    assert sorted(rng)[0][0] == 0 and sorted(rng)[-1][1] == n
    assert sorted(rng)[0][1] == sorted(rng)[1][0] and sorted(rng)[1][1] == sorted(rng)[2][0]
    fl = [v for v in vs if v[0].startswith("final_layer")]
    assert all(rng[0][0] <= v[1] < rng[0][1] for v in fl)


def test_param_count_base_and_errors():
    cfg = L.ModelConfig(6, 512, 8, 2048, 32768, 32768, 64, 512, 512, 1024, 0.1, 1e-6)
    n = L.load().b200xf_model_param_count(C.byref(cfg))
    assert abs(n - 94.50e6) < 0.02e6                          # SURVEY Appendix B (+ alignment padding)
    assert L.load().b200xf_model_workspace_bytes(C.byref(cfg)) > 1e9
    bad = L.ModelConfig(2, 100, 3, 256, 10, 10, 1, 8, 8, 8, 0.1, 1e-6)      # d_model % num_heads != 0
    assert L.load().b200xf_model_param_count(C.byref(bad)) < 0
    assert b"num_heads" in L.load().b200xf_last_error()


def test_schedule_matches_oracle_and_reference_values():
    from transformer_b200.train import CustomSchedule
    s = CustomSchedule(512)
    for step in (0, 1, 2, 100, 59999, 60000, 60001, 10 ** 6):
        assert s(step) == R.custom_schedule(step, 512)
    assert s(0) == 0.0
    np.testing.assert_allclose(s(60000), 1.8042195e-4, rtol=1e-6)


def test_positional_encoding_and_masks_match_oracle(golden_dir):
    from transformer_b200 import positionalencoding as pe
    assert np.array_equal(pe.positional_encoding(64, 8).numpy(), np.load(os.path.join(golden_dir, "posenc_64x8.npy")))
    inp = torch.tensor([[5, 3, 0], [0, 0, 1]]); tar = torch.tensor([[2, 0], [4, 4]])
    for a, b in zip(pe.create_masks(inp, tar), R.create_masks(inp, tar)):
        assert torch.equal(a, b)


def test_replica_split_and_flags():
    from transformer_b200.distributed_train import split_replica_batch, MirroredStrategy
    assert split_replica_batch(128, 8) == 16
    with pytest.raises(ValueError, match="divisible"):
        split_replica_batch(130, 8)
    st = MirroredStrategy()
    assert st.num_replicas_in_sync == 1
    src = torch.arange(12).view(4, 3); tar = torch.arange(16).view(4, 4)
    s, t = st.shard(src, tar)
    assert torch.equal(s, src) and torch.equal(t, tar)
    from absl import flags
    from transformer_b200 import utils
    if "num_layers" not in flags.FLAGS:
        utils.transformer_flags()
    flags.FLAGS(["prog"])
    d = utils.flags_dict()
    assert d["num_layers"] == 4 and d["d_model"] == 512 and d["dff"] == 1024 and d["num_heads"] == 4 and d["batch_size"] == 64
    assert d["dropout_rate"] == 0.1 and d["sequence_length"] == 50 and d["per_replica_batch_size"] == 16


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from transformer_b200.distributed_train import GradientBuckets, MirroredStrategy
    # DP-equivalence with the ORACLE's gradients pushed through the product's bucket logic (Appendix A.10):
    cfg = R.make_cfg(1, 32, 1, 64, 37, 37, rate=0.0, max_position=16)
    P = R.init_params(1, 32, 64, 37, 37, seed=3, perturb=True, dtype=torch.float64)
    src, tar = R.synthetic_batch(4, 5, 7, 37, 37, pad_tail=True)
    st = MirroredStrategy()
    s, t = st.shard(src, tar)
    _, g, _, _ = R.loss_and_grads(P, cfg, s, t, 4)            # per-replica loss = sum_local / B_global
    flat = torch.cat([v.reshape(-1) for v in g.values()])
    n = flat.numel()
    buckets = GradientBuckets([(2 * n // 3, n), (n // 3, 2 * n // 3), (0, n // 3)])
    for stage in range(3):
        buckets.reduce_stage(flat, stage)
    buckets.wait()
    if rank == 0:
        _, g_full, _, _ = R.loss_and_grads(P, cfg, src, tar, 4)
        full = torch.cat([v.reshape(-1) for v in g_full.values()])
        q.put(float((flat - full).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_buckets_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err < 1e-10, err


def test_text_pipeline_semantics(tmp_path):
    """utils.load_dataset stand-in keeps the reference's batching semantics (utils.py:114-161)."""
    from transformer_b200 import utils
    rng = np.random.default_rng(0)
    words = [f"w{i}" for i in range(50)]
    src_lines = [" ".join(rng.choice(words, rng.integers(1, 12))) for _ in range(103)]
    tgt_lines = [" ".join(rng.choice(words, rng.integers(1, 12))) for _ in range(103)]
    (tmp_path / "src-train.txt").write_text("\n".join(src_lines) + "\n")
    (tmp_path / "tgt-train.txt").write_text("\n".join(tgt_lines) + "\n")
    sv, tv = str(tmp_path / "sv"), str(tmp_path / "tv")
    train, test, st, tt = utils.load_dataset(str(tmp_path) + "/", 8, sv, tv, 16, 1000)
    assert os.path.exists(sv + ".subwords") and os.path.exists(tv + ".subwords")
    n_test = max(1, int(103 * 0.05))
    assert sum(b[0].shape[0] for b in test) == n_test                       # held-out tail (reference ships no test split)
    kept = [(s, t) for s, t in zip(src_lines[:-n_test], tgt_lines[:-n_test])
            if len(s.split()) + 2 <= 8 and len(t.split()) + 2 <= 8]         # length filter counts start/end ids, train only
    assert sum(b[0].shape[0] for b in train) == len(kept)
    assert all(b[0].shape[0] == 16 for b in train[:-1]) and 1 <= train[-1][0].shape[0] <= 16   # last batch partial
    for src, tgt in train:
        assert src.dtype == torch.int64 and src.shape[1] <= 8 and tgt.shape[1] <= 8
        assert (src[:, 0] == st.vocab_size).all() and (tgt[:, 0] == tt.vocab_size).all()       # start id = vocab_size
        for row in src:
            nz = row[row != 0]
            assert nz[-1] == st.vocab_size + 1 and (row[len(nz):] == 0).all()                   # end id, then pad 0
    # tokenizer round trip and the "vocabulary exists" path
    s = src_lines[0]
    assert st.decode(st.encode(s)) == s
    train2, _, st2, _ = utils.load_dataset(str(tmp_path) + "/", 8, sv, tv, 16, 1000)
    assert st2.vocab_size == st.vocab_size and torch.equal(train2[0][0], train[0][0])


def _c_kind(decl):
    """coarse kind of one C parameter declaration: 'p' pointer, 'f' float, 'i64' 64-bit integer, 'i' 32-bit integer"""
    d = decl.strip()
    if "*" in d or d.startswith("b200xf_stream_t"):
        return "p"
    t = " ".join(d.split()[:-1]) if len(d.split()) > 1 else d
    if t in ("float",):
        return "f"
    if t in ("int64_t", "uint64_t", "long long", "unsigned long long"):
        return "i64"
    if t in ("int", "uint32_t", "int32_t", "unsigned", "unsigned int"):
        return "i"
    raise AssertionError(f"unclassified C parameter '{d}'")


def test_ctypes_signatures_match_the_header():
    """include/b200xf.h is the contract; transformer_b200/_lib.py must bind every entry with the same number of
    parameters and the same kind (pointer / float / 32- / 64-bit integer) in every position -- an ABI drift between the
    two would otherwise only show up as garbage arguments on the GPU box."""
    header = open(os.path.join(ROOT, "include", "b200xf.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = dict(re.findall(r"\b(b200xf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S))
    kinds = {C.c_void_p: "p", C.c_char_p: "p", C.c_float: "f", C.c_int64: "i64", C.c_uint64: "i64", C.c_longlong: "i64",
             C.c_int: "i", C.c_uint32: "i", C.c_int32: "i"}
    checked = 0
    for name, argtypes in L._SIGS.items():
        assert name in protos, f"{name} bound in _lib.py but not declared in the header"
        params = [a for a in protos[name].replace("\n", " ").split(",") if a.strip() and a.strip() != "void"]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, _lib.py binds {len(argtypes)}"
        for i, (decl, at) in enumerate(zip(params, argtypes)):
            got = kinds.get(at, "p" if hasattr(at, "contents") or issubclass(at, C._Pointer) else None)   # POINTER(struct) etc.
            assert _c_kind(decl) == got, f"{name} parameter {i} ('{decl.strip()}'): header {_c_kind(decl)} vs ctypes {got}"
        checked += 1
    assert checked >= 30


def test_bench_reference_arm_contract_and_no_cpu_fallback():
    """bench.py --impl reference (the oracle on host cores) prints ONE JSON line with the contract's keys; the product arm
    has no CPU fallback: without a GPU it must fail loudly, never print a number."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny_s40_b32", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in j, k
    assert j["impl"] == "reference" and j["value"] > 0 and j["unit"] == "tokens/s" and "workload" in j["config"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0 and j["e2e"]["value"] == j["value"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    prod = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "tiny_s40_b32", "--steps", "1", "--warmup", "1",
                           "--no-cpu-baseline"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert prod.returncode != 0 and not any(l.startswith("{") for l in prod.stdout.splitlines())


def test_partial_batch_shards_and_distributed_entry_point():
    """the text pipeline emits a partial last batch; the strategy must split it (unevenly) instead of raising, and
    `distributed_train.py --num_gpu` (distributed_train.py:22-23,124-174) must exist with the reference's signature."""
    import inspect
    from transformer_b200 import distributed_train as D
    st = D.MirroredStrategy()
    got = []
    for rank in range(3):
        st.num_replicas_in_sync, st.rank = 3, rank
        s, t = st.shard(torch.arange(10).view(5, 2), torch.arange(15).view(5, 3))
        assert s.shape[0] == t.shape[0]
        got.append(s)
    assert [g.shape[0] for g in got] == [2, 2, 1] and torch.equal(torch.cat(got), torch.arange(10).view(5, 2))
    st.num_replicas_in_sync, st.rank = 4, 3
    s, t = st.shard(torch.arange(4).view(2, 2), torch.arange(6).view(2, 3))
    assert s.shape == (0, 2) and t.shape == (0, 3)                      # fewer rows than replicas: an empty shard, not an error
    params = list(inspect.signature(D.main).parameters)
    assert params == ['epochs', 'enable_function', 'buffer_size', 'batch_size', 'd_model', 'dff', 'num_heads', 'src_vocab_file',
                      'tgt_vocab_file', 'dataset_path', 'dropout_rate', 'num_layers', 'sequence_length', 'per_replica_batch_size',
                      'ckpt_path', 'max_ckpt_keep', 'num_gpu']              # distributed_train.py:131-132
    from absl import flags
    D._define_num_gpu_flag()
    assert flags.FLAGS['num_gpu'].default == 4                             # distributed_train.py:23
    assert callable(D.run_main)


def test_synthetic_corpus_feeds_the_text_pipeline(tmp_path):
    from transformer_b200 import utils
    p = str(tmp_path / "data") + "/"
    tr, te, st, tt = utils.load_dataset(p, 50, str(tmp_path / "sv"), str(tmp_path / "tv"), 64, 1000)
    assert len(tr) > 100 and len(te) >= 1 and tr[0][0].shape[0] == 64 and tr[0][0].dtype == torch.int64
    assert all(s.shape[1] <= 50 and t.shape[1] <= 50 for s, t in tr)


def test_prefetch_to_device_passthrough_on_cpu():
    from transformer_b200.utils import prefetch_to_device
    batches = [(torch.arange(6).view(2, 3), torch.arange(8).view(2, 4)) for _ in range(3)]
    got = list(prefetch_to_device(batches, "cpu"))
    assert len(got) == 3 and all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(got, batches))
