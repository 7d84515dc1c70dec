"""Batch-split of one long run over GPUs (BASELINE configs[2]): per-rank column buffers and kept columns with
emulated ranks on one GPU; the real thing -- processing_vaeflex_dp(group=...) over NCCL and over NVLink peer memory -- on >= 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vaeq_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("world", [2, 5])
def test_emulated_ranks_with_local_buffers_and_kept_columns(world):
    """Every rank passes q / out buffers that hold only its own columns (base pointers shifted by col0) and the frame-level keep
    buffers; the kept columns of all ranks tile the kept section exactly and equal the single-GPU frame step."""
    from vae_equalizer_b200.dp import DPEqualizer
    from vae_equalizer_b200.parallel import split_ranges, kept_owner_ranges
    M, B, stride = 25, 496 * 12, 496 * 6
    keep_lo, keep_n = (B - stride) // 2, stride
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", "64-QAM", "cpu", 0.0270955, 2, M, 23)
    rx, tx, _ = O.generate_data_shaping(B, amps, 23, h_ch, P, 2, 90e9, 2, -26e-24, 0.1e-12 * np.sqrt(1000),
                                        np.array([0.0314, 0.0314], dtype=np.complex64), np.pi / 10, "cpu", rng=np.random.default_rng(4))
    Pt = torch.tensor(P, dtype=torch.float32)
    rxd = rx.cuda()
    ref = DPEqualizer(M, 2, amp, Pt, var, nu_sc)
    ot_ref, oc_ref = torch.zeros(2, 16, stride, device="cuda"), torch.zeros(2, 2, stride, device="cuda")
    loss_ref, _ = ref.train_frame(rxd, B, stride, 1, 2.5e-3, 2.5e-3, ot_ref, oc_ref, keep_lo, keep_n)
    ranks = [DPEqualizer(M, 2, amp, Pt, var, nu_sc) for _ in range(world)]
    ranges = split_ranges(B, world)
    ot, oc = torch.zeros_like(ot_ref), torch.zeros_like(oc_ref)
    bufs, stats = [], []
    for eq, (lo, hi) in zip(ranks, ranges):
        col0, col1 = max(0, lo - 16), min(B, hi + 16)
        q, out = torch.full((2, 16, col1 - col0), float("nan"), device="cuda"), torch.full((2, 2, col1 - col0), float("nan"), device="cuda")
        bufs.append((q, out, col0))
        ot_r, oc_r = torch.zeros_like(ot_ref), torch.zeros_like(oc_ref)
        if lo == 0:                                                  # one rank without per-window q / out: kept columns only
            bufs[-1] = (None, None, 0)
            q = out = None
            col0 = 0
        stats.append(eq.split_forward(rxd, lo, hi, q, out, col0, ot_r, oc_r, keep_lo, keep_n).clone())
        assert int(((ot != 0) & (ot_r != 0)).sum()) == 0          # no column is kept by two ranks
        ot += ot_r
        oc += oc_r
    total = torch.stack(stats).sum(0)
    grads = []
    for eq, (lo, hi), (q, out, col0) in zip(ranks, ranges, bufs):
        eq._stats.copy_(total)
        grads.append(eq.split_backward(rxd, lo, hi, q, out, col0).clone())
    gsum = torch.stack(grads).sum(0)
    for eq, (q, out, col0) in zip(ranks, bufs):
        eq._grads.copy_(gsum)
        eq.split_update(rxd, q, out, 2.5e-3, 2.5e-3, col0)
    torch.cuda.synchronize()
    assert torch.equal(ot, ot_ref) and torch.equal(oc, oc_ref)
    parts = kept_owner_ranges(B, world, keep_lo, keep_n)
    assert sum(b - a for a, b in parts) == keep_n
    for eq in ranks:
        assert rel(eq.W, ref.W) < 1e-5 and rel(eq.h, ref.h) < 1e-5 and rel(eq.loss, loss_ref) < 1e-6
    with pytest.raises(Exception):
        lo, hi = ranges[-1]
        ranks[0].split_forward(rxd, lo, hi, bufs[1][0], bufs[1][1], bufs[1][2] if world > 2 else 4)   # buffers that do not cover the rank's columns: rejected


ARGS = dict(mod="64-QAM", sps=2, SNR=23, nu=0.0270955, M_est=25, theta_diff=0.0, theta=np.pi / 10, lr_optim=2.5e-3, batch_len=496 * 64,
            N_train_max=496 * 64 * 3, num_frames=3, flex_step=496 * 32, channel="h0", symb_rate=90e9, tau_cd=-26e-24,
            tau_pmd=0.1e-12 * np.sqrt(1000), phiIQ=np.array([0.0314, 0.0314], dtype=np.complex64), N_lrhalf=2)


def _worker(rank, world, port, transport, resq):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from vae_equalizer_b200.processing import processing_vaeflex_dp
        a = ARGS
        ser, var_est, var = processing_vaeflex_dp(a["mod"], a["sps"], a["SNR"], a["nu"], a["M_est"], a["theta_diff"], a["theta"], a["lr_optim"],
                                                  a["batch_len"], a["N_train_max"], a["num_frames"], a["flex_step"], a["channel"], a["symb_rate"],
                                                  a["tau_cd"], a["tau_pmd"], a["phiIQ"], a["N_lrhalf"], verbose=False, datagen="gpu", seed=5,
                                                  group=True, split_transport=transport)
        resq.put((rank, ser.cpu().numpy(), var_est.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_vaeflex_driver_batch_split_over_two_gpus(transport):
    """func_VAEflex_DP's frame loop with every window split over 2 GPUs (NVLink peer reductions / NCCL all-reduces) against the same
    driver on one GPU: same frames (seeded device generator), Var_est within 1e-4, SER within two decisions per polarisation."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from vae_equalizer_b200.processing import processing_vaeflex_dp
    a = ARGS
    ser1, ve1, _ = processing_vaeflex_dp(a["mod"], a["sps"], a["SNR"], a["nu"], a["M_est"], a["theta_diff"], a["theta"], a["lr_optim"], a["batch_len"],
                                         a["N_train_max"], a["num_frames"], a["flex_step"], a["channel"], a["symb_rate"], a["tau_cd"], a["tau_pmd"],
                                         a["phiIQ"], a["N_lrhalf"], verbose=False, datagen="gpu", seed=5)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    resq = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, transport, resq)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([resq.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n_eval = (ARGS["N_train_max"] - ARGS["batch_len"]) - 50
    for rank, ser, ve in res:
        assert np.array_equal(ser, res[0][1])                                   # broadcast: every rank returns rank 0's SER
        assert np.abs(ve - ve1.cpu().numpy()).max() / np.abs(ve1.cpu().numpy()).max() < 1e-4
        assert np.abs(ser - ser1.cpu().numpy()).max() <= 2.5 / n_eval
