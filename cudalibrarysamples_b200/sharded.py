"""Row-block sharding of a square CSR matrix over the GPUs of one box (one process per GPU, torch.distributed).

SpMV shards naturally by contiguous row blocks: rank g owns rows [R_g, R_{g+1}) of A -- chosen so every
rank holds ~nnz/G non-zeros, not rows/G, because R-MAT rows are skewed -- plus the matching slice of x and y.  The only
exchange step of the path is one all-gather of x (NCCL over NVLink 5 / NVSwitch) right before the local kernel; in CG the x
shard of one product is the y shard of the previous one, so the gather sits on the critical path of every iteration.

Shards have different row counts, so the gathered vector uses a padded layout: shard g occupies
x_full[g*pad : g*pad + rows_g], pad = max_g rows_g, and the local column indices are remapped into that layout ONCE at
set-up (structure-only preprocessing, like the tile plan).  The local product is then an ordinary rows_g x (G*pad) CSR SpMV
through the same C ABI as the single-GPU path.

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
    """This rank's row block of a global square CSR matrix plus the padded all-gather layout for x."""

    def __init__(self, off: torch.Tensor, col: torch.Tensor, val: torch.Tensor, rank: int, world: int,
                 make_local_op: Callable[[int, int, dict], Callable], group=None, base: int = 0):
        assert base == 0
        self.rank, self.world, self.group = rank, world, group
        n = off.numel() - 1
        self.global_rows = n
        self.global_nnz = int(off[-1].item())
        self.bounds = split_rows_by_nnz(off, world)               # [world+1], identical on every rank
        b = self.bounds.tolist()
        self.r0, self.r1 = b[rank], b[rank + 1]
        self.rows = self.r1 - self.r0
        self.pad = max(1, max(b[g + 1] - b[g] for g in range(world)))
        n0, n1 = int(off[self.r0].item()), int(off[self.r1].item())
        self.off = (off[self.r0:self.r1 + 1].to(torch.int64) - n0).to(torch.int32).contiguous()
        gcol = col[n0:n1].to(torch.int64)
        owner = torch.searchsorted(self.bounds[1:].contiguous(), gcol, right=True)          # block that owns the column
        self.col = (owner * self.pad + (gcol - self.bounds[owner])).to(torch.int32).contiguous()
        self.val = val[n0:n1].contiguous()
        self.nnz = n1 - n0
        self.cols_padded = world * self.pad
        self.x_full = torch.zeros(self.cols_padded, dtype=val.dtype, device=val.device)
        self.local_op = make_local_op(self.rows, self.cols_padded, dict(off=self.off, col=self.col, val=self.val))

    def new_shard(self, fill=None):
        """A padded vector shard (pad entries, the first `rows` are live)."""
        t = torch.zeros(self.pad, dtype=self.val.dtype, device=self.val.device)
        if fill is not None:
            t[:self.rows] = fill[self.r0:self.r1]
        return t

    def gather_x(self, x_shard: torch.Tensor) -> torch.Tensor:
        """The path's single exchange step: all-gather of the padded x shards."""
        if self.world == 1:
            self.x_full[:self.pad].copy_(x_shard)
        else:
            dist.all_gather_into_tensor(self.x_full, x_shard, group=self.group)
        return self.x_full

    def spmv(self, x_shard: torch.Tensor, y_shard: torch.Tensor, alpha=1.0, beta=0.0) -> torch.Tensor:
        """y_shard[:rows] = alpha * A[r0:r1, :] @ x + beta * y_shard[:rows]   (x given as this rank's padded shard)."""
        self.gather_x(x_shard)
        if self.rows > 0:
            self.local_op(self.x_full, y_shard[:self.rows], alpha, beta)
        return y_shard

    def unpad(self, x_full: torch.Tensor) -> torch.Tensor:
        """Padded gathered layout -> the global vector (test helper)."""
        b = self.bounds.tolist()
        return torch.cat([x_full[g * self.pad: g * self.pad + (b[g + 1] - b[g])] for g in range(self.world)])
