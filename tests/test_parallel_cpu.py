"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes exercise the sweep sharder, the result gather and the
all-reduce plumbing of vae_equalizer_b200.parallel; the symbol-range partition is checked directly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vae_equalizer_b200 import parallel as par


def test_split_ranges_cover_and_align():
    for B, world in ((1 << 22, 8), (1 << 22, 4), (4032, 2), (100800, 3), (1 << 20, 1)):
        r = par.split_ranges(B, world)
        assert r[0][0] == 0 and r[-1][1] == B and len(r) == world
        for (a, b), (c, d) in zip(r, r[1:]):
            assert b == c
        assert all(lo % 4 == 0 and hi % 4 == 0 and hi > lo for lo, hi in r)
        sizes = [hi - lo for lo, hi in r]
        assert max(sizes) - min(sizes) <= 2 * par.TILE
    with pytest.raises(ValueError):
        par.split_ranges(1002, 2)
    with pytest.raises(ValueError):
        par.split_ranges(992, 8)


def test_split_ranges_balance_the_cost_of_the_kept_section():
    """Frame loops write q / out for the kept middle section only (VAEflex_DP:64-65): ranges of equal COST, still a tiling of [0, B) in whole tiles."""
    for B, world in ((8 << 22, 8), (4 << 22, 4), (2 << 22, 2), (496 * 64, 3)):
        keep = (B // 4, B // 2)
        r, e = par.split_ranges(B, world, keep=keep), par.split_ranges(B, world)

        def cost(lo, hi):
            return (hi - lo) + par.KEEP_COST * max(0, min(hi, keep[0] + keep[1]) - max(lo, keep[0]))

        assert r[0][0] == 0 and r[-1][1] == B and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % 4 == 0 and (hi - lo) % par.TILE == 0 for lo, hi in r[:-1])
        spread = max(cost(*x) for x in r) / min(cost(*x) for x in r)
        assert spread <= max(cost(*x) for x in e) / min(cost(*x) for x in e) + 1e-9
        if B >= 1 << 22:
            assert spread < 1.001
        parts = par.kept_owner_ranges(B, world, keep[0], keep[1], weighted=True)
        assert sum(b - a for a, b in parts) == keep[1]
        assert [p for p in parts if p[1] > p[0]][0][0] == 0
    assert par.split_ranges(1 << 22, 4, keep=None) == par.split_ranges(1 << 22, 4, keep=(0, 1 << 22), keep_cost=0.0)


def test_sweep_cells_order_matches_reference_loops():
    cells = par.sweep_cells(nu=[0, 0.027], lr=[2.5e-3, 2e-3, 3e-3], SNR=[20, 23], it=list(range(5)))
    assert len(cells) == 60 and cells[0] == dict(nu=0, lr=2.5e-3, SNR=20, it=0) and cells[1]["it"] == 1 and cells[5]["SNR"] == 23
    mine = [par.shard_cells(cells, r, 8) for r in range(8)]
    assert sorted(i for m in mine for i, _ in m) == list(range(60))
    assert max(len(m) for m in mine) - min(len(m) for m in mine) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cells = par.sweep_cells(SNR=[15, 17, 19], it=[0, 1, 2])
        mine = par.shard_cells(cells, rank, world)
        local = {i: torch.full((4, 3), float(100 * c["SNR"] + c["it"])) for i, c in mine}      # stands for SER_valid of a run
        full = par.gather_cell_results(local, len(cells), (4, 3), rank, world)
        stats = torch.tensor([1.0 + rank, 10.0 * (rank + 1)], dtype=torch.float64)
        par.allreduce_sum_(stats)
        ranges = par.split_ranges(16 * par.TILE, world)
        q.put((rank, None if full is None else full[:, 0, 0].tolist(), stats.tolist(), ranges[rank]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_shard_gather_allreduce():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, full0, st0, rg0), (r1, full1, st1, rg1) = res
    assert full1 is None and full0 == [1500.0, 1501.0, 1502.0, 1700.0, 1701.0, 1702.0, 1900.0, 1901.0, 1902.0]
    assert st0 == st1 == [3.0, 30.0]
    assert rg0 == (0, 8 * par.TILE) and rg1 == (8 * par.TILE, 16 * par.TILE)


def test_kept_owner_ranges_tile_the_kept_section():
    for B, world, stride in ((496 * 64, 2, 496 * 32), (1 << 22, 8, 1 << 21), (496 * 12, 5, 496 * 6), (4032, 3, 1000)):
        keep_lo = (B - stride) // 2
        parts = par.kept_owner_ranges(B, world, keep_lo, stride)
        assert len(parts) == world and sum(b - a for a, b in parts) == stride
        nonempty = [(a, b) for a, b in parts if b > a]
        assert nonempty[0][0] == 0 and nonempty[-1][1] == stride
        assert all(b == c for (_, b), (c, _) in zip(nonempty, nonempty[1:]))


class _FakeEq:
    device, sps, M, n_lev = torch.device("cpu"), 2, 25, 8


def _gather_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, stride, n_steps = 496 * 8, 496 * 4, 3
        keep_lo = (B - stride) // 2
        bs = par.BatchSplitDP(_FakeEq(), None, "nccl")             # host logic only: ranges, kept-column exchange
        lo, hi, col0, n = bs.local_columns(B)
        ot, oc = torch.zeros(2, 16, n_steps * stride), torch.zeros(2, 2, n_steps * stride)
        a, b = par.kept_owner_ranges(B, world, keep_lo, stride, weighted=True)[rank]
        for m in range(n_steps):                                   # what the forward kernel leaves: this rank's part of every kept section
            cols = torch.arange(m * stride + a, m * stride + b, dtype=torch.float32)
            ot[:, :, m * stride + a:m * stride + b] = cols
            oc[:, :, m * stride + a:m * stride + b] = -cols
        bs.gather_kept(B, stride, n_steps, ot, oc, keep_lo, stride, dst=0)
        full = torch.arange(n_steps * stride, dtype=torch.float32)
        ok = bool((ot == full).all() and (oc == -full).all()) if rank == 0 else True
        q.put((rank, ok, (lo, hi, col0, n)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_kept_columns():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == (0, 496 * 4, 0, 496 * 4 + 16) and res[1][2] == (496 * 4, 496 * 8, 496 * 4 - 16, 496 * 4 + 16)
