"""world_size-2 gloo test of the frame-sharding host logic: rank-major all-gather == global frame order, and the
search window per query reproduces the sequential reference (keyframe.cpp:274-282) for any round size."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, b, rounds, q):
    sys.path.insert(0, ROOT)
    from d_vins_b200 import sharding
    from oracle import knn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng_all = np.random.default_rng(7)
    total = rounds * world * b
    truth = rng_all.standard_normal((total, 512)).astype(np.float32)        # descriptor of global frame t
    truth /= np.linalg.norm(truth, axis=1, keepdims=True)
    bank = np.zeros((0, 512), np.float32)
    ok = True
    for R in range(rounds):
        ids = sharding.round_frame_ids(R, rank, world, b)
        mine = torch.from_numpy(truth[ids])
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)                                      # stand-in for the ncclAllGather
        bank = np.concatenate([bank] + [g.numpy() for g in gathered])
        ok &= np.array_equal(bank, truth[:len(bank)])                        # bank row == global frame index
        for t in ids:
            nb = sharding.nb_limit(int(t))
            D, I = knn.knn_ip(bank, truth[t], nb)
            Ds, Is = knn.knn_ip(truth[: t + 1], truth[t], knn.nb_limit(int(t)))   # sequential single-GPU reference
            ok &= np.array_equal(I, Is) and np.array_equal(D, Ds)
            ok &= sharding.owner_of(int(t), world, b) == rank
        if R > 0:
            prev = sharding.previous_round_ids(ids, world, b)
            ok &= all(sharding.owner_of(int(p), world, b) == rank for p in prev)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("b,rounds", [(4, 20), (32, 4)])
def test_two_rank_sharding_matches_sequential(b, rounds):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + b
    procs = [ctx.Process(target=_worker, args=(r, 2, port, b, rounds, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
