"""N > 1 host logic on CPU (gloo, world_size 2): the landmark partition of SURVEY 8e and the identity the
multi-GPU path relies on -- the all-reduced sum of the per-rank Schur-reduced systems equals the reduced system of
the whole window."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lvio_fusion_b200 import backend, synth


def test_partition_is_exact():
    d = synth.make_ba_problem(6, 500, with_imu=True, seed=4)
    world = 3
    shards = [synth.shard_ba_problem(d, r, world) for r in range(world)]
    for kind, (c, ix) in d["factors"].items():
        got = np.concatenate([s["factors"][kind][0] for s in shards])
        assert len(got) == len(c)
        assert sorted(map(tuple, got.round(9))) == sorted(map(tuple, c.round(9)))
    for r, s in enumerate(shards):                       # every block touching rho_l lives on rank l % world
        assert ((s["factors"][0][1][:, 0] % world) == r).all()
        assert ((s["factors"][2][1][:, 0] % world) == r).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding
    orc = binding.load()
    ctx = backend.Context(orc)
    d = synth.make_ba_problem(5, 300, with_imu=True, seed=8)
    p = backend.Problem.from_dict(ctx, synth.shard_ba_problem(d, rank, world))
    S, b, cost = p.reduced_system(1e30)                  # radius -> infinity: no LM damping, pure Schur complement
    t = torch.from_numpy(np.concatenate([S.ravel(), b, [cost]]))
    dist.all_reduce(t)                                   # the one collective of an LM iteration
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


def test_allreduced_shards_equal_the_full_system(tmp_path, orc_ctx):
    world = 2
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = np.load(out)
    d = synth.make_ba_problem(5, 300, with_imu=True, seed=8)
    S, b, cost = backend.Problem.from_dict(orc_ctx, d).reduced_system(1e30)
    n = len(b)
    Sg, bg, cg = got[:n * n].reshape(n, n), got[n * n:n * n + n], got[-1]
    assert abs(cg - cost) < 1e-9 * cost
    assert np.max(np.abs(Sg - S)) < 1e-9 * np.abs(S).max()
    assert np.max(np.abs(bg - b)) < 1e-9 * np.abs(b).max()
