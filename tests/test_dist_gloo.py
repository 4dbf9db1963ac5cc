"""CPU tier: the N>1 host logic (sharding by problem index, gather to rank 0) with world_size 2 and 3 over gloo.
The per-shard compute is a stand-in (the CUDA ops need a GPU); what is tested is that results are invariant to the
number of ranks and arrive in problem order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from learningagileflight_se3_b200.dist import shard_range, sharded_call


def _fake_op(x0, t):
    # deterministic per-problem function with the output shapes of sol_gradient ([B,8]) and objective ([B]); like the real ops it
    # returns EMPTY outputs for an empty shard (ops.py: B == 0 returns before any workspace / launch)
    g = torch.stack([x0[:, i % 13] * (i + 1) + t for i in range(8)], 1)
    return {"grad": g, "reward": x0.sum(1) - t, "status": (x0[:, 0] > 0).to(torch.int32)}


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    inputs = {"x0": rng.normal(size=(n, 13)), "t": rng.uniform(2, 4, n)}
    out = sharded_call(_fake_op, inputs)
    if rank == 0:
        q.put({k: v.numpy() for k, v in out.items()})
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 4096, 1000003):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_results_are_invariant_to_world_size():
    n = 37          # ragged: not divisible by 2 or 3
    rng = np.random.default_rng(0)
    inputs = {"x0": rng.normal(size=(n, 13)), "t": rng.uniform(2, 4, n)}
    want = {k: v.numpy() for k, v in _fake_op(torch.as_tensor(inputs["x0"]), torch.as_tensor(inputs["t"])).items()}
    for world in (2, 3):
        got = _run(world, n)
        for k in want:
            assert np.array_equal(got[k], want[k]), (world, k)


def test_fewer_problems_than_ranks_does_not_deadlock():
    """n < world: some ranks own an empty shard; they must still take part in the gather (an exception before the collective
    would leave the other ranks blocked in it until the timeout)."""
    rng = np.random.default_rng(0)
    for n, world in ((2, 3), (1, 2)):
        inputs = {"x0": rng.normal(size=(n, 13)), "t": rng.uniform(2, 4, n)}
        rng = np.random.default_rng(0)
        want = {k: v.numpy() for k, v in _fake_op(torch.as_tensor(inputs["x0"]), torch.as_tensor(inputs["t"])).items()}
        got = _run(world, n)
        for k in want:
            assert got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), (world, n, k)
