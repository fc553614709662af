"""T4: data-parallel training on real GPUs (NCCL): an N-GPU step over a global batch must equal the 1-GPU step on the
same global batch (distributed_train.py semantics: per-replica loss = sum_local / B_global, gradients SUM-reduced,
identical Adam on every replica).  Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`);
skips on a 1-GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from oracle import restatement as R
    from transformer_b200.Transformer import Transformer
    from transformer_b200.train import Train
    from transformer_b200.distributed_train import DistributedTrain, MirroredStrategy

    L_, d, H, f, V = 2, 128, 2, 256, 1003
    P = R.init_params(L_, d, f, V, V, seed=5, perturb=True)
    Bg = 8
    batches = [R.synthetic_batch(Bg, 40, 33, V, V, seed=50 + i, pad_tail=True) for i in range(3)]
    strategy = MirroredStrategy()
    model = Transformer(L_, d, H, f, V, V, rate=0.0, pos_rows=64, device=f"cuda:{rank}")
    if rank == 0:
        model.load_variables(P)          # other ranks keep their own init: DistributedTrain must broadcast rank 0's weights
    tr = DistributedTrain(1, True, model, None, None, Bg, Bg // world, None, None, 5, "ckpt", d, warmup_steps=50)
    it = iter(batches)
    for _ in range(3):
        strategy.experimental_run(tr.train_step, it)
    tr.reduce_metrics()
    torch.cuda.synchronize()
    flat = model.params.clone()
    # replicas must hold identical weights
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(ref, flat))
    if rank == 0:
        single = Transformer(L_, d, H, f, V, V, rate=0.0, pos_rows=64, device="cuda:0")
        single.load_variables(P)
        tr1 = Train(1, True, single, None, None, Bg, None, None, 5, "ckpt", d, warmup_steps=50)
        for b in batches:
            tr1.train_step(b)
        torch.cuda.synchronize()
        num = (single.params - flat).norm().item()
        moved = (single.params.norm().item())
        # reported losses: DP train_loss is the mean over replicas of per-replica losses = global / N (Appendix A.10)
        q.put(dict(rel=num / moved, same=same, loss_dp=tr.train_loss.result(), loss_1=tr1.train_loss.result(), world=world))
    else:
        q.put(dict(same=same))
    torch.cuda.synchronize()
    dist.barrier()
    q.close(); q.join_thread()      # flush the result before leaving without destructors
    os._exit(0)       # destroy_process_group() can block once collectives were captured into CUDA graphs


def test_dp_step_equals_single_gpu_step():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0
    assert all(o["same"] for o in outs), "replicas diverged"
    o = [o for o in outs if "rel" in o][0]
    print(o)
    assert o["rel"] < 5e-3, f"DP weights differ from the single-GPU run: {o}"
    # train_loss under the strategy = sum over replicas of (local_sum / B_global) accumulated per step / (N steps * N replicas)
    assert abs(o["loss_dp"] * o["world"] - o["loss_1"]) / o["loss_1"] < 2e-3, o
