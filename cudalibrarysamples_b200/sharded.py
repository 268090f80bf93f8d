"""Row-block sharding of a square CSR matrix over the GPUs of one box (one process per GPU, torch.distributed).

SpMV shards naturally by contiguous row blocks: rank g owns rows [R_g, R_{g+1}) of A -- chosen so every
rank holds ~nnz/G non-zeros, not rows/G, because R-MAT rows are skewed -- plus the matching slice of x and y.  The only
exchange step of the path is one all-gather of x (NCCL over NVLink 5 / NVSwitch) right before the local kernel; in CG the x
shard of one product is the y shard of the previous one, so the gather sits on the critical path of every iteration.

x itself is distributed in EQUAL blocks (not by the row blocks), so the gather moves exactly |x|*(G-1)/G bytes per
rank, no padding, and the gathered buffer is x: the local product is an ordinary rows_g x cols CSR SpMV through the
same C ABI as the single-GPU path, on the caller's unmodified column indices.

Host logic only -- the local kernel is injected (`make_local_op`) so the world_size-2 gloo tests on CPU can drive the
same code with the CPU oracle, while bench.py passes the sm_100a operator.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def split_rows_by_nnz(off: torch.Tensor, world: int) -> torch.Tensor:
    """Row boundaries R_0=0 <= ... <= R_world=rows with ~equal non-zeros per block (prefix search on rowOff)."""
    rows = off.numel() - 1
    o = off.to(torch.int64) - off[0].to(torch.int64)
    nnz = int(o[-1].item())
    targets = torch.tensor([(nnz * g) // world for g in range(1, world)], dtype=torch.int64, device=off.device)
    inner = torch.searchsorted(o.contiguous(), targets, right=False).clamp_(0, rows)
    b = torch.cat([torch.zeros(1, dtype=torch.int64, device=off.device), inner,
                   torch.full((1,), rows, dtype=torch.int64, device=off.device)])
    return torch.cummax(b, 0).values


class ShardedCsr:
    """This rank's row block of a global square CSR matrix plus the all-gather layout for x.

    Two distributions, both fixed at set-up:
      * A and y by contiguous row blocks with ~nnz/G non-zeros each (rows [r0, r1) on this rank);
      * x by EQUAL blocks of x_block = ceil(n/G) entries (the last one zero-padded), so the exchange step is one
        all_gather_into_tensor of equal shards with no padding traffic, and the gathered vector is x itself:
        column indices need no remapping.
    (In a solver loop y becomes the next x: with skewed matrices the two distributions differ and the hand-over is a
    redistribution of the row-block / equal-block overlap; for banded matrices they coincide up to a halo.)
    """

    def __init__(self, off: torch.Tensor, col: torch.Tensor, val: torch.Tensor, rank: int, world: int,
                 make_local_op: Callable[[int, int, dict], Callable], group=None, base: int = 0, balance: str = "nnz",
                 exchange: str = "auto"):
        assert base == 0
        self.rank, self.world, self.group = rank, world, group
        n = off.numel() - 1
        self.global_rows = n
        self.global_nnz = int(off[-1].item())
        if balance == "nnz":
            self.bounds = split_rows_by_nnz(off, world)           # [world+1], identical on every rank
        else:   # "rows": equal row blocks == the x blocks (regular matrices; lets a solver hand y over as the next x)
            blk = max(1, (n + world - 1) // world)
            self.bounds = torch.tensor([min(g * blk, n) for g in range(world + 1)], dtype=torch.int64, device=off.device)
        b = self.bounds.tolist()
        self.r0, self.r1 = b[rank], b[rank + 1]
        self.rows = self.r1 - self.r0
        self.x_block = max(1, (n + world - 1) // world)
        self.pad = self.x_block                                    # length of one x shard
        n0, n1 = int(off[self.r0].item()), int(off[self.r1].item())
        self.off = (off[self.r0:self.r1 + 1].to(torch.int64) - n0).to(torch.int32).contiguous()
        self.col = col[n0:n1].contiguous()
        self.val = val[n0:n1].contiguous()
        self.nnz = n1 - n0
        self.cols_padded = world * self.x_block
        self.x_full = torch.zeros(self.cols_padded, dtype=val.dtype, device=val.device)
        self.local_op = make_local_op(self.rows, self.cols_padded, dict(off=self.off, col=self.col, val=self.val))
        self._plan_exchange(exchange)

    def _plan_exchange(self, exchange: str):
        """Structure-only preprocessing of the exchange step: which part of x does this row block actually read?

        The local columns span [cmin, cmax].  For a banded matrix that is this rank's own block plus a halo (5-pt
        Poisson 8192^2 on 8 GPUs: 2 x 64 KB instead of 470 MB), so instead of the full all-gather every rank receives,
        from each owner, only the overlap of [cmin, cmax] with the owner's x block -- straight into x_full, no packing.
        Falls back to the all-gather when the ranks together need more than half of what the all-gather would move
        (R-MAT: every rank reads every block)."""
        world, blk = self.world, self.x_block
        self.recv_plan, self.send_plan, self.exchange = [], [], "allgather"
        if world == 1 or exchange == "allgather":
            return
        if self.nnz > 0:
            cmin, cmax = int(self.col.min().item()), int(self.col.max().item()) + 1
        else:
            cmin, cmax = 0, 0
        need = torch.tensor([[cmin, cmax]], dtype=torch.int64, device=self.col.device)
        allneed = [torch.zeros_like(need) for _ in range(world)]
        dist.all_gather(allneed, need, group=self.group)
        ranges = [tuple(int(v) for v in t.flatten().tolist()) for t in allneed]      # (cmin, cmax) of every rank
        total = 0
        for g, (lo_need, hi_need) in enumerate(ranges):
            for h in range(world):
                if h == g:
                    continue
                lo, hi = max(lo_need, h * blk), min(hi_need, (h + 1) * blk)
                if hi > lo:
                    total += hi - lo
                    if g == self.rank:
                        self.recv_plan.append((h, lo, hi))          # receive x[lo:hi] from its owner h
                    if h == self.rank:
                        self.send_plan.append((g, lo - h * blk, hi - h * blk))   # send my block[lo:hi] to g
        if exchange == "halo" or total * 2 < world * (world - 1) * blk:
            self.exchange = "halo"
            self.exchanged_elements = total

    def _halo_exchange(self, x_shard: torch.Tensor):
        blk = self.x_block
        self.x_full[self.rank * blk:(self.rank + 1) * blk].copy_(x_shard)
        ops = [dist.P2POp(dist.isend, x_shard[lo:hi], g, group=self.group) for g, lo, hi in self.send_plan]
        ops += [dist.P2POp(dist.irecv, self.x_full[lo:hi], h, group=self.group) for h, lo, hi in self.recv_plan]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def new_x_shard(self, x=None):
        """This rank's equal block of a global vector x (zero-padded at the end of the last block)."""
        t = torch.zeros(self.x_block, dtype=self.val.dtype, device=self.val.device)
        if x is not None:
            lo = self.rank * self.x_block
            hi = min(lo + self.x_block, self.global_rows)
            if hi > lo:
                t[:hi - lo] = x[lo:hi]
        return t

    def new_y_shard(self, y=None):
        """This rank's row block of a global vector y."""
        t = torch.zeros(max(self.rows, 1), dtype=self.val.dtype, device=self.val.device)[:self.rows]
        if y is not None and self.rows:
            t.copy_(y[self.r0:self.r1])
        return t

    def gather_x(self, x_shard: torch.Tensor) -> torch.Tensor:
        """The path's single exchange step: all-gather of the equal x shards, or -- when the set-up analysis found that
        this matrix only reads a narrow column range per rank -- just the needed ranges (halo exchange)."""
        if self.world == 1:
            self.x_full[:self.x_block].copy_(x_shard)
        elif self.exchange == "halo":
            self._halo_exchange(x_shard)
        else:
            dist.all_gather_into_tensor(self.x_full, x_shard, group=self.group)
        return self.x_full

    def spmv(self, x_shard: torch.Tensor, y_shard: torch.Tensor, alpha=1.0, beta=0.0) -> torch.Tensor:
        """y_shard = alpha * A[r0:r1, :] @ x + beta * y_shard   (x given as this rank's equal block)."""
        self.gather_x(x_shard)
        if self.rows > 0:
            self.local_op(self.x_full, y_shard, alpha, beta)
        return y_shard
