"""World-size-2 gloo test (CPU) of the N > 1 host logic: contiguous shards partition the batch,
every rank regenerates exactly its slice of the global counter-based ensemble, the per-rank
results concatenate to the single-rank result, and the report reduction is max(time)/sum(work).
The per-shard integration uses the CPU oracle as a stand-in for the GPU (test infrastructure)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from ode_solvers_b200 import _abi as A
from ode_solvers_b200 import sharding
from ode_solvers_b200 import workloads as W

N = 301  # not divisible by 2 on purpose


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(N, world, rank)
    y0, p = W.lorenz_ensemble(hi - lo, offset=lo)
    cfg = oracle.config_new(A.DOPRI5, 0.0, 2.0, 0.1, 1e-6, 1e-6, out_type=A.OUT_SPARSE)
    o = oracle.integrate_batch("lorenz", cfg, y0, p, n_threads=1)
    t, work = sharding.reduce_report(dist, torch.device("cpu"), 10.0 + rank, [int(o.accepted.sum()), hi - lo])
    q.put((rank, lo, hi, o.y_final.copy(), o.accepted.copy(), t, work))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # shards are contiguous, disjoint and cover [0, N)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == N
    y0, p = W.lorenz_ensemble(N)
    cfg = oracle.config_new(A.DOPRI5, 0.0, 2.0, 0.1, 1e-6, 1e-6, out_type=A.OUT_SPARSE)
    full = oracle.integrate_batch("lorenz", cfg, y0, p)
    ycat = np.concatenate([r[3] for r in res], axis=1)
    acat = np.concatenate([r[4] for r in res])
    assert np.array_equal(ycat.view(np.uint64), full.y_final.view(np.uint64))
    assert np.array_equal(acat, full.accepted)
    # report reduction: max of the times, sum of the work, identical on every rank
    for r in res:
        assert r[5] == 11.0 and r[6] == [int(full.accepted.sum()), N]


def test_shard_ranges():
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 4, 8):
            rs = [sharding.shard_range(n, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert all(0 <= hi - lo <= (n + world - 1) // world for lo, hi in rs)


def test_counter_based_workloads_slice_consistently():
    for gen in (W.lorenz_ensemble, W.kepler_sweep, W.cr3bp_ensemble, W.robertson_sweep, W.bouncing_ball_ensemble,
                W.threebody18_ensemble):
        a, pa = gen(100)
        b, pb = gen(40, offset=60)
        assert np.array_equal(a[:, 60:], b)
        if pa.ndim == 2:
            assert np.array_equal(pa[:, 60:], pb)
