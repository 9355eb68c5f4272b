"""world_size-2 gloo tests of the multi-process host logic (runs on CPU; NCCL is exercised under gpurun):
process-group bootstrap, the autograd-aware collectives, and the data-parallel gradient averaging contract
(every rank ends a step with identical averaged gradients => identical replicas, no data-path collective
other than the gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    from src.utils.distributed import AllGather, AllReduce, AllReduceSum, init_distributed
    ws, rk = init_distributed(port=port, rank_and_world_size=(rank, world))
    assert (ws, rk) == (world, rank) and dist.get_backend() == "gloo"
    assert init_distributed() == (world, rank)  # early return once a group exists (distributed.py:20-21)
    x = torch.tensor([float(rank + 1)], requires_grad=True)
    avg = AllReduce.apply(x * 1.0)
    tot = AllReduceSum.apply(x * 1.0)
    gat = AllGather.apply(torch.full((2, 3), float(rank)))
    (avg + tot).sum().backward()
    # DDP contract on a plain module: averaged grads identical on every rank
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 2)
    ddp = torch.nn.parallel.DistributedDataParallel(lin, static_graph=True)
    ddp(torch.full((3, 4), float(rank + 1))).sum().backward()
    q.put((rank, float(avg), float(tot), gat.tolist(), float(x.grad), lin.weight.grad.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_collectives():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, avg, tot, gat, xg, wg in res:
        assert avg == 1.5 and tot == 3.0 and xg == 2.0
        assert gat == [[0.0] * 3] * 2 + [[1.0] * 3] * 2
    assert res[0][5] == res[1][5] == [[4.5] * 4] * 2  # mean over ranks of 3 * (rank + 1)


def test_init_distributed_without_slurm_returns_single_process(monkeypatch):
    from src.utils.distributed import init_distributed
    monkeypatch.delenv("SLURM_NTASKS", raising=False)
    assert init_distributed() == (1, 0)
