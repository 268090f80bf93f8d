"""Row-sharded conjugate gradient around the SpMV path (BASELINE.json config 4: 5-pt Poisson 8192^2, fp64, 200 fixed
iterations, iterations/s at 1..8 GPUs).

Restates the loop of the reference's gpu_CG (cuSPARSE/cg/cg_example.c:132-307) without the IC(0) preconditioner: the
reference's preconditioner is a global incomplete Cholesky + two SpSV solves, which does not row-shard,
so the sharded driver is plain CG (the SpMV call per iteration, cg_example.c:220-224, is the part this project
replaces).  Same set-up as the sample: b = 0.75 * A * 1, x0 = 0 (cg_example.c:405-420).

Host logic only: the local SpMV is injected, so the gloo tests drive it with the CPU oracle and cg_bench.py with the
sm_100a operator.  Vector updates and dot products are torch ops (plumbing); the dots are all-reduced over ranks.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from .sharded import ShardedCsr


def _dot(a: torch.Tensor, b: torch.Tensor, world: int, group=None) -> torch.Tensor:
    d = torch.dot(a, b).reshape(1)
    if world > 1:
        dist.all_reduce(d, group=group)
    return d


def conjugate_gradient(sh: ShardedCsr, b_shard: torch.Tensor, iters: int, x0_shard: torch.Tensor | None = None,
                       tol: float | None = None):
    """Plain CG on this rank's row block; requires the x / y distributions of `sh` to coincide (equal row blocks), which
    is the case for the stencil matrices of config 4 when rows % world == 0.

    Returns (x_shard, residual_norms) -- residual_norms[k] = ||r_k||_2 (global), k = 0 .. iters_done.
    """
    assert sh.rows == sh.x_block or sh.world == 1, "CG needs y-rows == x-block (regular matrix, rows % world == 0)"
    world, group = sh.world, sh.group
    x = torch.zeros_like(b_shard) if x0_shard is None else x0_shard.clone()
    r = b_shard.clone()
    t = torch.zeros_like(b_shard)
    if x0_shard is not None:
        sh.spmv(x, r, alpha=-1.0, beta=1.0)            # r = b - A x0   (cg_example.c:153-160)
    p = r.clone()
    delta = _dot(r, r, world, group)
    norms = [delta.sqrt()]
    for _ in range(iters):
        sh.spmv(p, t, alpha=1.0, beta=0.0)             # T = A * P      (cg_example.c:220-224)
        denom = _dot(t, p, world, group)
        alpha = delta / denom                          # 1-element device tensors: no host sync in the loop
        x.addcmul_(p, alpha)                           # X += alpha P   (cg_example.c:236-239), one pass, no temporary
        r.addcmul_(t, alpha, value=-1.0)               # R -= alpha T   (cg_example.c:241-244)
        delta_new = _dot(r, r, world, group)
        norms.append(delta_new.sqrt())
        if tol is not None and float(norms[-1]) < tol * float(norms[0]):
            break
        beta = delta_new / delta
        torch.addcmul(r, p, beta, out=p)               # P = beta P + R (cg_example.c:280-286), one pass
        delta = delta_new
    return x, torch.cat(norms)


class CgSolver:
    """conjugate_gradient() packaged for repeated timed runs (bench.py's cg_config4 leg, scripts/cg_bench.py)."""

    def __init__(self, sh: ShardedCsr, b_shard: torch.Tensor):
        self.sh, self.b = sh, b_shard

    def run(self, iters: int):
        x, norms = conjugate_gradient(self.sh, self.b, iters)
        return x, [float(v) for v in norms.tolist()]

    def describe(self) -> str:
        return ("plain CG (cg_example.c:215-287 without the IC(0) preconditioner), SpMV through the C ABI, vector updates as "
                "single-pass torch ops (addcmul), dots all-reduced over ranks, no host synchronisation inside the loop")


def make_cg_solver(sh: ShardedCsr, b_shard: torch.Tensor) -> CgSolver:
    return CgSolver(sh, b_shard)
