"""Data-parallel host logic on CPU with world_size 2 (gloo): molecule sharding = DistributedSampler(shuffle=False),
flat-buffer gradient all-reduce = sum of the per-rank gradients (reference: Lightning DDP, trainer.py:308-325,
datamodules.py:40-41)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import conan_fgw_b200 as cmp
from conan_fgw_b200.dp import FlatParameters, shard_molecules


def test_shard_matches_distributed_sampler():
    from torch.utils.data import DistributedSampler

    for n, world in ((10, 2), (11, 4), (7, 8), (128, 8), (3, 2)):
        data = list(range(n))
        for rank in range(world):
            want = list(DistributedSampler(data, num_replicas=world, rank=rank, shuffle=False))
            assert shard_molecules(n, rank, world) == want, (n, world, rank)
    assert shard_molecules(0, 0, 2) == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
        flat = FlatParameters([net])
        assert all(p.data.data_ptr() >= flat.flat.data_ptr() for p in net.parameters())
        # each rank sees its own shard of a common data set
        g = torch.Generator().manual_seed(1)
        x_all, y_all = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
        idx = shard_molecules(8, rank, world)
        flat.zero_grad()
        loss = torch.nn.functional.mse_loss(net(x_all[idx]), y_all[idx], reduction="sum")
        loss.backward()
        flat.collect_grads()
        local = flat.grad.clone()
        w = flat.all_reduce()
        assert w == world
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(flat.grad, sum(gathered), atol=1e-6)
        # equals the single-process gradient over the whole data set
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 1))
        torch.nn.functional.mse_loss(ref(x_all), y_all, reduction="sum").backward()
        ref_flat = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        assert torch.allclose(flat.grad, ref_flat, atol=1e-5)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
