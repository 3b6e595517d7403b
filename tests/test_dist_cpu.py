"""Host-side logic of the multi-GPU BA path on CPU: two gloo ranks shard the points the way
bench.py does (BAProblem.shard), evaluate their partial sums with the oracle and all-reduce them;
the reduced quantities must equal the single-process values (the CUDA solver all-reduces exactly
these sums over NCCL)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coslam_b200 import synth
    from oracle import orc
    prob, _ = synth.make_ba_scene(2, 4, 400, 320, 240, seed=5, m_con=2, n_con=3)
    shard, (lo, hi) = prob.shard(rank, world)
    # every observation lands in exactly one shard
    cnt = torch.tensor([shard.nobs, shard.n, shard.n_con], dtype=torch.int64)
    dist.all_reduce(cnt)
    # partial weighted cost (all weights 1) == sum of squared residuals of this shard
    part = torch.tensor([orc.ba_cost(shard)], dtype=torch.float64)
    dist.all_reduce(part)
    # per-camera partial sums (what the solver packs into its all-reduce buffer): U diag proxy
    r = shard.residuals()
    percam = np.zeros(prob.m)
    np.add.at(percam, shard.cam, (r * r).sum(1))
    pc = torch.from_numpy(percam)
    dist.all_reduce(pc)
    if rank == 0:
        q.put((cnt.tolist(), float(part.item()), pc.numpy(), orc.ba_cost(prob), prob.nobs, prob.n,
               prob.n_con))
    dist.barrier()
    dist.destroy_process_group()


def test_point_sharding_partial_sums_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cnt, cost_sum, percam, cost_full, nobs, n, ncon = out
    assert cnt == [nobs, n, ncon]
    assert abs(cost_sum - cost_full) <= 1e-9 * cost_full
    assert abs(percam.sum() - cost_full) <= 1e-9 * cost_full


def test_shard_balance_and_coverage():
    sys.path.insert(0, ROOT)
    from coslam_b200 import synth
    prob, _ = synth.make_ba_scene(4, 6, 3000, 640, 480, seed=2, m_con=4, n_con=10)
    for world in (2, 4, 8):
        ranges, work = [], []
        for r in range(world):
            sh, (lo, hi) = prob.shard(r, world)
            ranges.append((lo, hi))
            k = np.diff(sh.ptr).astype(float)
            work.append((k * k + 8 * k).sum())
            assert sh.ptr[0] == 0 and sh.ptr[-1] == sh.nobs
            assert np.array_equal(sh.X, prob.X[lo:hi])
        assert ranges[0][0] == 0 and ranges[-1][1] == prob.n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        assert max(work) < 1.25 * np.mean(work)
        assert sum(prob.shard(r, world)[0].n_con for r in range(world)) == prob.n_con
