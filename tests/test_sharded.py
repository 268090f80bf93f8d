"""CPU, world_size 2 (gloo): row-block sharding + the padded all-gather of x, local product by the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cudalibrarysamples_b200.sharded import ShardedCsr, split_rows_by_nnz
from oracle import oracle as O


def _collect(q, procs, timeout=240):
    """The result rank 0 put on the queue -- or a test failure (not a hang) when a worker died or the time is up."""
    import time
    t0 = time.time()
    while q.empty():
        if any(p.exitcode not in (None, 0) for p in procs):
            for p in procs:
                p.kill()
            pytest.fail("a worker process failed: " + str([p.exitcode for p in procs]))
        if time.time() - t0 > timeout:
            for p in procs:
                p.kill()
            pytest.fail("workers timed out")
        time.sleep(0.05)
    return q.get()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_op(rows, cols, arrays):
    off, col, val = (arrays[k].numpy() for k in ("off", "col", "val"))

    def op(x, y, alpha, beta):
        y.copy_(torch.from_numpy(O.spmv_csr(off, col, val, x.numpy(), y.numpy(), alpha, beta)))
        return y
    return op


def _worker(rank, world, port, rows, q):
    os.environ["B200SPMV_PANELS_FROM"] = "2"           # column panels from 2 ranks on (default 4): the host logic under test
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        off, col, val = (torch.from_numpy(a) for a in O.rmat_csr(rows, avg_nnz=8, seed=3, val_seed=4))
        x = torch.from_numpy(O.uniform(5, rows))
        y0 = torch.from_numpy(O.uniform(6, rows))
        sh = ShardedCsr(off, col, val, rank, world, _oracle_op)
        assert sh.exchange == "allgather"                  # R-MAT: every rank reads every x block
        assert sh.panels and sh.own_nnz + sh.remote_nnz == sh.nnz   # local matrix split into own-column / remote-column panels
        assert "overlapped with the own-column panel" in sh.describe_exchange()
        xs, ys = sh.new_x_shard(x), sh.new_y_shard(y0)
        sh.spmv(xs, ys, alpha=-1.0, beta=1.0)
        # second product chained on the first (solver style: y becomes the next x -> redistribute row blocks into
        # equal blocks; here through an all_gather_object of the row blocks)
        parts = [None] * world
        dist.all_gather_object(parts, ys.numpy())
        y_full = torch.from_numpy(np.concatenate(parts))
        zs = sh.new_y_shard()
        sh.spmv(sh.new_x_shard(y_full), zs, alpha=1.0, beta=0.0)
        gathered = sh.x_full[:sh.global_rows].clone()
        # the timing-loop form bench.py uses (make_step: fixed buffers, alpha = 1, beta = 0, one call per column panel)
        zs2 = sh.new_y_shard(torch.from_numpy(O.uniform(7, rows)))        # beta = 0: whatever y held is overwritten
        step = sh.make_step(sh.new_x_shard(y_full), zs2, graph=False)
        step()
        step()
        assert len(sh.panel_calls) == len(sh.panel_groups) == min(world - 1, 3) + 1
        assert sorted(b for g in sh.panel_groups for b in g) == list(range(world))   # every x block in exactly one panel
        assert sh.panel_groups[0] == [rank] and [b for g in sh.panel_groups[1:] for b in g] == sh.pull_order
        assert torch.equal(zs2, zs)
        out = [None] * world
        dist.all_gather_object(out, dict(rank=rank, r0=sh.r0, r1=sh.r1, nnz=sh.nnz, y=ys.numpy(), z=zs.numpy(),
                                         xg=gathered.numpy(), x_block=sh.x_block))
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


def test_column_panels_partition_the_local_matrix():
    from cudalibrarysamples_b200.sharded import split_column_panels
    off, col, val = (torch.from_numpy(a) for a in O.rmat_csr(3000, avg_nnz=8, seed=9, val_seed=10))
    x, y0 = O.uniform(1, 3000), O.uniform(2, 3000)
    lo, hi = 1000, 2000
    (oo, oc, ov), (ro, rc, rv) = split_column_panels(off, col, val, lo, hi)
    assert oc.numel() + rc.numel() == col.numel()
    assert bool(((oc >= lo) & (oc < hi)).all()) and bool(((rc < lo) | (rc >= hi)).all())
    assert int(oo[-1]) == oc.numel() and int(ro[-1]) == rc.numel() and int(oo[0]) == 0 and int(ro[0]) == 0
    # own panel first (beta as given), remote panel second with beta = 1: the same product
    y = O.spmv_csr(oo.numpy(), oc.numpy(), ov.numpy(), x, y0, -0.5, 2.0)
    y = O.spmv_csr(ro.numpy(), rc.numpy(), rv.numpy(), x, y, -0.5, 1.0)
    want = O.spmv_csr(off.numpy(), col.numpy(), val.numpy(), x, y0, -0.5, 2.0)
    assert np.linalg.norm(y - want) <= 1e-13 * np.linalg.norm(want)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_spmv_matches_single_process(world):
    rows = 4000
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = _collect(q, procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    off, col, val = O.rmat_csr(rows, avg_nnz=8, seed=3, val_seed=4)
    x, y0 = O.uniform(5, rows), O.uniform(6, rows)
    y = O.spmv_csr(off, col, val, x, y0, -1.0, 1.0)
    z = O.spmv_csr(off, col, val, y)
    out.sort(key=lambda d: d["rank"])
    assert out[0]["r0"] == 0 and out[-1]["r1"] == rows
    for a, b in zip(out[:-1], out[1:]):
        assert a["r1"] == b["r0"]
    got_y = np.concatenate([d["y"] for d in out])
    got_z = np.concatenate([d["z"] for d in out])
    # every row is summed in two column panels (own block first, the rest with beta = 1): equal up to rounding
    assert np.linalg.norm(got_y - y) <= 1e-14 * np.linalg.norm(y)
    assert np.linalg.norm(got_z - z) <= 1e-13 * np.linalg.norm(z)
    assert np.array_equal(out[0]["xg"], got_y)   # the last gather carried y, reassembled bit for bit from equal blocks
    assert all(d["x_block"] == (rows + world - 1) // world for d in out)
    nnzs = [d["nnz"] for d in out]
    assert sum(nnzs) == off[-1]
    assert max(nnzs) - min(nnzs) <= np.diff(off).max() + 1   # balanced up to one row


def test_split_rows_by_nnz_edge_cases():
    off = torch.tensor([0, 0, 0, 10, 10, 10], dtype=torch.int32)
    b = split_rows_by_nnz(off, 4).tolist()
    assert b[0] == 0 and b[-1] == 5 and all(x <= y for x, y in zip(b[:-1], b[1:]))
    off = torch.zeros(8, dtype=torch.int32)
    b = split_rows_by_nnz(off, 3).tolist()
    assert b[0] == 0 and b[-1] == 7


def _cg_worker(rank, world, port, grid, iters, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cudalibrarysamples_b200.cg import conjugate_gradient
        off, col, val = (torch.from_numpy(a) for a in O.gen_stencil5(grid))
        n = grid * grid
        sh = ShardedCsr(off, col, val, rank, world, _oracle_op, balance="rows")
        assert sh.exchange == "halo"                       # 5-pt stencil: only a halo of `grid` entries per neighbour
        assert sh.exchanged_elements == 2 * (world - 1) * grid
        ones = torch.ones(n, dtype=torch.float64)
        b = sh.new_y_shard()
        sh.spmv(sh.new_x_shard(ones), b, alpha=0.75, beta=0.0)        # b = 0.75 * A * 1  (cg_example.c:405-418)
        x, norms = conjugate_gradient(sh, b, iters)
        out = [None] * world
        dist.all_gather_object(out, dict(rank=rank, x=x.numpy(), norms=norms.numpy()))
        if rank == 0:
            q.put(out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_cg_converges_like_scipy(world):
    """Config 4's solver at toy size: 5-pt Laplacian (cg_example.c:71-128), b = 0.75*A*1, x0 = 0, plain CG."""
    import scipy.sparse as sp
    grid, iters = 48, 150
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, grid, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = _collect(q, procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    out.sort(key=lambda d: d["rank"])
    x = np.concatenate([d["x"] for d in out])
    norms = out[0]["norms"]
    off, col, val = O.gen_stencil5(grid)
    n = grid * grid
    A = sp.csr_matrix((val, col, off), shape=(n, n))
    b = 0.75 * (A @ np.ones(n))
    assert abs(norms[0] - np.linalg.norm(b)) < 1e-9 * np.linalg.norm(b)
    assert norms[-1] < 1e-6 * norms[0]                      # converged (exact solution is 0.75 * ones)
    assert np.linalg.norm(x - 0.75) / np.linalg.norm(0.75 * np.ones(n)) < 1e-6
    assert np.linalg.norm(b - A @ x) <= 1.01 * norms[-1] + 1e-12   # the recurrence residual is the true residual
