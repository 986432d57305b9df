"""World-size-2 test (gloo, CPU) of the data-parallel decomposition the multi-GPU solve uses: every rank lowers the
FULL problem, evaluates its time-contiguous share of every residual table with the product's own lowering + analytic Jacobians
(host-compiled, liblvi_hostcheck.so), and the all-reduced {cost, J^T r, J^T J} must equal the single-rank normal equations —
which is exactly what lvi_problem_solve does with ncclAllReduce on the GPUs.  Also covers the NCCL-id bootstrap broadcast of bench.py."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lvi_exc_b200._capi import ProblemDesc, c_double_p, ptr


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _normal_equations(pd, rank, world):
    from tests import hostcheck_binding as hc
    L = hc.lib()
    L.lvi_hostcheck_normal_equations.argtypes = [C.POINTER(ProblemDesc), C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]
    nt = hc.layout(pd)["nt"]
    cost, g, H = np.zeros(1), np.zeros(nt), np.zeros((nt, nt))
    d = pd.desc()
    assert L.lvi_hostcheck_normal_equations(C.byref(d), rank, world, ptr(cost), ptr(g), ptr(H)) == 0
    return cost, g, H


def _worker(rank, world, port, stage, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.problems import make_lvi_problem
        pd = make_lvi_problem(stage, 1.0, 300) if stage != "lvi" else make_lvi_problem(stage)
        cost, g, H = _normal_equations(pd, rank, world)
        tc, tg, tH = torch.from_numpy(cost.copy()), torch.from_numpy(g.copy()), torch.from_numpy(H.copy())
        for t in (tc, tg, tH):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        # bootstrap pattern of bench.py: rank 0 owns a 128-byte id, everyone must end up with the same bytes
        idbuf = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(idbuf, 0)
        assert idbuf.tolist() == list(range(128))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), cost=tc.numpy(), g=tg.numpy(), H=tH.numpy(), local_cost=cost)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("stage", ["surfel", "lvi"])
def test_sharded_normal_equations_sum_to_the_full_problem(stage, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), stage, str(tmp_path)), nprocs=world, join=True)
    from tests.problems import make_lvi_problem
    pd = make_lvi_problem(stage, 1.0, 300) if stage != "lvi" else make_lvi_problem(stage)
    cost, g, H = _normal_equations(pd, 0, 1)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert 0 < r0["local_cost"][0] < cost[0] and 0 < r1["local_cost"][0] < cost[0]      # both ranks really hold a share
    for r in (r0, r1):
        assert r["cost"][0] == pytest.approx(cost[0], rel=1e-12)
        assert np.abs(r["g"] - g).max() <= 1e-9 * max(1.0, np.abs(g).max())
        assert np.abs(r["H"] - H).max() <= 1e-9 * max(1.0, np.abs(H).max())
    assert np.array_equal(r0["H"], r1["H"])   # every rank factorises the same all-reduced system


def test_shard_ranges_partition_every_table():
    for n in (0, 1, 7, 12080, 110048):
        for world in (1, 2, 4, 8):
            edges = [(n * r // world, n * (r + 1) // world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))


# ---- the sharded map path (csrc/shard.cu): its bookkeeping, restated in lvi_exc_b200/shardplan.py, at world size 2 over gloo ----------------
def _map_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lvi_exc_b200 import shardplan
        rng = np.random.default_rng(7)
        n = 20000
        pts_all = (rng.random((n, 3)) * np.array([12.0, 9.0, 3.5]) - np.array([6.0, 4.5, 1.0])).astype(np.float32)   # the same cloud on every rank
        lo, hi = n * rank // world, n * (rank + 1) // world     # ranks = consecutive time chunks
        pts = pts_all[lo:hi]
        inv = np.float32(1.0) / np.float32(0.5)
        # collective 1: grid of the WHOLE cloud from all-reduced min / max
        mn, mx = torch.from_numpy(pts.min(0).copy()), torch.from_numpy(pts.max(0).copy())
        dist.all_reduce(mn, op=dist.ReduceOp.MIN); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        min_b = np.floor(mn.numpy() * inv).astype(np.int64)
        div_b = np.floor(mx.numpy() * inv).astype(np.int64) - min_b + 1
        ijk = (np.floor(pts * inv) - min_b.astype(np.float32)).astype(np.int64)
        keys = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
        ncell = int(div_b.prod())
        # ownership ranges from the all-reduced histogram
        hist = torch.from_numpy(shardplan.key_histogram(keys, ncell).astype(np.int64))
        dist.all_reduce(hist, op=dist.ReduceOp.SUM)
        split = shardplan.balanced_splitters(hist.numpy().astype(np.uint64), ncell, world)
        owner = shardplan.owner_of(keys, split)
        order = np.argsort(owner, kind="stable")
        send = [dict(xyz=pts[order][owner[order] == q], idx=(lo + order[owner[order] == q])) for q in range(world)]
        got = [None] * world
        dist.all_gather_object(got, send)                       # the all-to-all of points
        mine_xyz = np.concatenate([got[q][rank]["xyz"] for q in range(world)])   # rank order = time order
        mine_idx = np.concatenate([got[q][rank]["idx"] for q in range(world)])
        # decimation bookkeeping: random hit counts per rank
        tots = [137 + 59 * q for q in range(world)]
        first, count = shardplan.decimation_plan(tots, 10)
        off = sum(tots[:rank])
        kept = [off + first[rank] + 10 * j for j in range(count[rank])]
        np.savez(os.path.join(out_dir, f"m{rank}.npz"), xyz=mine_xyz, idx=mine_idx, split=split, min_b=min_b, div_b=div_b, kept=np.array(kept), tots=np.array(tots))
    finally:
        dist.destroy_process_group()


def test_sharded_map_bookkeeping_world2(tmp_path):
    world = 2
    mp.spawn(_map_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    n = 20000
    pts = (rng.random((n, 3)) * np.array([12.0, 9.0, 3.5]) - np.array([6.0, 4.5, 1.0])).astype(np.float32)
    inv = np.float32(2.0)
    min_b = np.floor(pts.min(0) * inv).astype(np.int64)
    div_b = np.floor(pts.max(0) * inv).astype(np.int64) - min_b + 1
    ijk = (np.floor(pts * inv) - min_b.astype(np.float32)).astype(np.int64)
    keys = ijk[:, 0] + ijk[:, 1] * div_b[0] + ijk[:, 2] * div_b[0] * div_b[1]
    r = [np.load(tmp_path / f"m{q}.npz") for q in range(world)]
    assert all(np.array_equal(x["min_b"], min_b) and np.array_equal(x["div_b"], div_b) for x in r)      # the grid of the whole cloud on every rank
    assert np.array_equal(r[0]["split"], r[1]["split"])
    split = r[0]["split"]
    seen = 0
    for q in range(world):
        own = np.nonzero((keys >= split[q]) & (keys < split[q + 1]))[0]                                 # single-process: the owner's points in cloud order
        assert np.array_equal(r[q]["idx"], own) and np.array_equal(r[q]["xyz"], pts[own])               # ... arrive in exactly that order
        assert 0.35 * n < len(own) < 0.65 * n                                                           # balanced ranges
        seen += len(own)
    assert seen == n
    tots = r[0]["tots"]
    kept = np.concatenate([x["kept"] for x in r])
    assert np.array_equal(kept, np.arange(0, tots.sum(), 10))                                           # the global every-10th decimation
