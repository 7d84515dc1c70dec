"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY.md §8e.

* Sweep sharding: the cells of the reference's parameter sweep (Eval_run_DP.py:68-95: nu x batch_len x lr x M x
  theta_diff x symb_rate x flex_step x theta x SNR x iter) are independent `processing()` calls.  `shard_cells`
  deals them round-robin to the ranks, `gather_cell_results` reassembles the per-cell result tensors on rank 0.
  No data-path collective.
* Batch-split of ONE long minibatch (VAE-flex with large batch_len): `BatchSplitDP` gives every rank a contiguous
  symbol range of the window and all-reduces (a) the ELBO partial sums (8 + 2(M-1) doubles) before the backward scale
  kappa = (L-Mh)/C is known and (b) the 16*M tap-gradient floats before the replicated Adam step.  Both messages are
  latency-bound (< 1 KB); NCCL over NVLink carries them on the step's stream.
CMA does not shard (per-symbol tap recurrence): replicas only.
"""
from __future__ import annotations

import itertools

import torch
import torch.distributed as dist

TILE = 496                      # owned symbols per tile of the fast kernels (csrc/dp_fast.cu FT_T = 4 * 128 - 16)


def split_ranges(B: int, world: int, align: int = TILE):
    """Contiguous symbol ranges [lo, hi) covering [0, B), multiples of 4, sized in whole tiles where possible."""
    if B % 4:
        raise ValueError("batch-split needs batch_len % 4 == 0")
    units = (B + align - 1) // align
    out, lo = [], 0
    for r in range(world):
        n = units // world + (1 if r < units % world else 0)
        hi = min(B, lo + n * align)
        if r == world - 1:
            hi = B
        out.append((lo, hi))
        lo = hi
    if any(hi <= lo for lo, hi in out):
        raise ValueError(f"batch_len={B} is too short to split {world} ways in tiles of {align} symbols")
    return out


def shard_cells(cells, rank: int, world: int):
    """Round-robin assignment of sweep cells (any sequence) to `rank`; returns [(global_index, cell), ...]."""
    return [(i, c) for i, c in enumerate(cells) if i % world == rank]


def sweep_cells(**axes):
    """Cartesian product of named parameter lists in the reference's loop order -> list of dicts."""
    names = list(axes)
    return [dict(zip(names, vals)) for vals in itertools.product(*(axes[n] for n in names))]


def gather_cell_results(local, n_cells: int, shape, rank: int, world: int, group=None, device="cpu"):
    """`local` = {global_index: tensor(shape)} computed by this rank.  Returns the (n_cells, *shape) tensor on rank 0
    (None elsewhere).  One all_gather of a dense, zero-padded block per rank; cells are disjoint so a sum suffices."""
    block = torch.zeros((n_cells,) + tuple(shape), dtype=torch.float32, device=device)
    for i, t in local.items():
        block[i] = t.to(device=device, dtype=torch.float32)
    if world > 1:
        dist.all_reduce(block, op=dist.ReduceOp.SUM, group=group)
    return block if rank == 0 else None


def allreduce_sum_(t: torch.Tensor, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class BatchSplitDP:
    """One DP VAE-LE/VAE-flex run whose minibatch is split over the ranks of `group` (all ranks call in lockstep).

    Every rank keeps the full rx window (an input, replicated once per frame outside the step) and an identical
    copy of W / h / Adam state; the update is replicated and bit-identical because both all-reduces return the same
    bits on every rank.  Summation order differs from the single-GPU run, so parity with it is ~1e-6, not bitwise."""

    def __init__(self, eq, group=None):
        self.eq, self.group = eq, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def train_step(self, rx, lr_w, lr_h, q, out):
        B = rx.shape[-1] // self.eq.sps
        lo, hi = split_ranges(B, self.world)[self.rank]
        stats = self.eq.split_forward(rx, lo, hi, q, out)
        allreduce_sum_(stats, self.group)
        grads = self.eq.split_backward(rx, lo, hi, q, out)
        allreduce_sum_(grads, self.group)
        self.eq.split_update(rx, q, out, lr_w, lr_h)
        return self.eq.loss, self.eq.var_est, (lo, hi)
