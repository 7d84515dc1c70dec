"""The AWGN single-polarisation CMA module (second half) against outputs of the UNMODIFIED reference
`AWGN_channel/func_CMA_MQAM_shaping.py` (tests/golden/make_golden_awgn_cma.py): CMA :142-168, CPE without unwrap :170-196,
SER_CMA :63-93, find_shift_symb :127-140 and the whole `processing()` :201-256 on replayed frames."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import vaeq_oracle as O

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.mark.parametrize("name", ["awgncma_ops_16qam_M25", "awgncma_ops_64qam_M9"])
def test_awgn_cma_operators_against_the_reference(name):
    from vae_equalizer_b200 import awgn_cma as cm
    g = load(name)
    amp, lr = T(g["amp"]).cuda(), float(g["lr"])
    for i in range(2):                                        # teacher forcing: each frame starts from the reference's incoming taps
        h = T(g[f"h_in{i}"]).cuda()
        out, h2, e = cm.CMA(T(g[f"rx{i}"]).cuda(), 1, h, lr, 2, True)
        assert h2 is h                                        # updated in place AND returned (cm:163-168)
        assert float((out.cpu() - T(g[f"out{i}"])).abs().max()) < 2e-5 and float((e.cpu() - T(g[f"e{i}"])).abs().max()) < 5e-5
        assert float((h.cpu() - T(g[f"h_out{i}"])).abs().max()) < 2e-5
    h = T(g["h_out1"]).cuda()
    h_before = h.clone()
    out_v, _, e_v = cm.CMA(T(g["rx_v"]).cuda(), 1, h, lr, 2, False)
    assert torch.equal(h, h_before)                           # eval == False: no update
    assert float((out_v.cpu() - T(g["out_v"])).abs().max()) < 2e-5
    cpe = cm.CPE(T(g["out_v"]).cuda())
    assert float((cpe.cpu() - T(g["cpe"])).abs().max()) < 5e-6
    shift, corr = cm.find_shift_symb(T(g["cpe"]).cuda(), T(g["tx_v"]).cuda(), 21, return_corr=True)
    assert int(shift) == int(g["shift"]) and shift.dtype == torch.int64
    a = T(g["ser_in"]).cuda()
    ser, counts = cm.SER_CMA(a, T(g["ser_tx"]).cuda(), 2, amp, amp.numel(), "cuda", return_counts=True)
    assert np.array_equal(ser.cpu().numpy(), g["ser"])        # identical float
    assert float((a.cpu() - T(g["ser_in_scaled"])).abs().max()) < 1e-6        # rescaled in place (cm:73)
    oc, n = O.awgn_ser_cma_counts(T(g["ser_in"]).clone(), T(g["ser_tx"]), T(g["amp"]))
    assert counts.cpu().tolist() == oc.tolist()               # all four rotations' integer counts against the oracle


def test_awgn_cma_batched_streams_equal_single_calls():
    from vae_equalizer_b200 import awgn_cma as cm
    g = load("awgncma_ops_16qam_M25")
    rx = torch.stack((T(g["rx0"]), T(g["rx1"]))).cuda()
    h = torch.stack((T(g["h_in0"]), T(g["h_in1"]))).cuda()
    out, h2, e = cm.CMA(rx, 1, h, float(g["lr"]), 2, True)
    for i in range(2):
        hi = T(g[f"h_in{i}"]).cuda()
        oi, _, ei = cm.CMA(rx[i].contiguous(), 1, hi, float(g["lr"]), 2, True)
        assert torch.equal(oi, out[i]) and torch.equal(hi, h[i]) and torch.equal(ei, e[i])
    y = torch.stack((T(g["out_v"]), T(g["out_v"]).flip(0))).cuda()
    c = cm.CPE(y)
    assert torch.equal(c[0], cm.CPE(y[0])) and torch.equal(c[1], cm.CPE(y[1]))


@pytest.mark.parametrize("dropin", [False, True])
def test_awgn_cma_driver_on_the_references_frames(dropin):
    """processing() on the frames the reference's processing() was given: same shifts, SER within a few decisions of the reference's
    (0.675 -> 0.198 -> 0.017 over three evaluations: the equalizer converges)."""
    g = load("awgncma_drv_16qam")
    SNR, nu, M, lr, N_valid, N_train, num_epochs, epe = [float(v) for v in g["args"]]
    frames = iter([(g[f"rx{i}"], g[f"tx{i}"]) for i in range(int(g["n_frames"]))])
    if dropin:
        sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "vae_equalizer_b200", "dropin"))
        import func_CMA_MQAM_shaping as process
        assert process.__file__.endswith(os.path.join("dropin", "func_CMA_MQAM_shaping.py"))
        fn = process.processing
        assert all(hasattr(process, n) for n in ("CMA", "CPE", "SER_CMA", "find_shift_symb"))
    else:
        from vae_equalizer_b200.processing import processing_cma_awgn as fn
    SER = fn(str(g["mod"]), 2, SNR, nu, int(M), lr, int(N_valid), int(N_train), int(num_epochs), int(epe), str(g["channel"]), verbose=False,
             datagen=frames)
    assert SER.shape == (int(num_epochs) // int(epe),) and SER.dtype == torch.float32
    assert np.abs(SER.cpu().numpy() - g["SER"]).max() <= 4.0 / (int(N_valid) - 30), (SER, g["SER"])
