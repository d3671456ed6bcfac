"""N > 1 host logic on CPU: two gloo ranks shard the instances, draw their realisations from the global
instance ids and all-gather per-instance statistics; the result must equal the unsharded run."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from rtmpc_b200 import distributed as D


def test_shard_covers_everything_once():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [D.shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def _worker(rank, world, port, total, T, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = D.shard(total, rank, world)
    hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
    p = np.array([0.1 * (i % 10) for i in range(total)])
    th, ga, w = H.device_draws(seed, np.arange(off, off + cnt), T, p[off:off + cnt], hw)
    stat = torch.as_tensor(np.c_[th.sum(0), ga.sum(0), np.abs(w).sum((0, 2))])        # [cnt, 3] per instance
    allst = D.all_gather_instances(stat, total)
    counts = D.all_reduce_sum(torch.tensor([cnt, int(th.sum()), int(ga.sum())]))
    tmax = D.all_reduce_max(torch.tensor(float(rank + 1)))
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), allst=allst.numpy(), counts=counts.numpy(), tmax=tmax.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_reproduce_the_unsharded_run(tmp_path):
    total, T, seed = 37, 12, 679          # ragged: 19 + 18
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, total, T, seed, str(tmp_path)), nprocs=2, join=True)
    g = np.load(tmp_path / "gathered.npz")
    hw = np.array([1e-4, 2.7e-3, 3e-4, 4.3e-2])
    p = np.array([0.1 * (i % 10) for i in range(total)])
    th, ga, w = H.device_draws(seed, np.arange(total), T, p, hw)
    ref = np.c_[th.sum(0), ga.sum(0), np.abs(w).sum((0, 2))]
    assert np.array_equal(g["allst"], ref)
    assert g["counts"].tolist() == [total, int(th.sum()), int(ga.sum())]
    assert float(g["tmax"]) == 2.0


def test_single_process_passthrough():
    x = torch.arange(6.0).reshape(3, 2)
    assert D.all_gather_instances(x) is x
    assert torch.equal(D.all_reduce_sum(torch.tensor([1, 2])), torch.tensor([1, 2]))
