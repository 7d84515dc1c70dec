"""Host-side run constants: the product-side counterpart of shared_funcs.init (reference
optical_DP_channel/shared_funcs.py:544-588).  Pure numpy on the host, float64 like the reference,
cast to float32 only where the reference does."""
from __future__ import annotations

import numpy as np
import torch

CHANNEL_TAPS = {
    "h0": (1.0 + 0.0j,),                                                            # optical channel only (sf:549)
    "h1": (0.0545 + 0.05j, 0.2823 - 0.11971j, -0.7676 + 0.2788j, -0.0641 - 0.0576j, 0.0466 - 0.02275j),   # sf:546
    "h2": (0.0545 + 0.0165j, -1.3449 - 0.4523j, 1.0067 + 1.1524j, 0.3476 + 0.3153j),                       # sf:548
}
QAM_SIDE = {"4-QAM": 2, "16-QAM": 4, "64-QAM": 8}


def upsampled_channel(channel: str, sps: int) -> np.ndarray:
    if channel not in CHANNEL_TAPS:
        raise KeyError(f"unknown channel {channel!r}; expected one of {sorted(CHANNEL_TAPS)}")
    taps = np.array(CHANNEL_TAPS[channel]).astype(np.complex64)
    ir = np.zeros(sps * (len(taps) - 1) + 1, dtype=np.complex64)
    ir[::sps] = taps
    return ir / np.linalg.norm(ir)


def qam_points(mod: str) -> np.ndarray:
    if mod not in QAM_SIDE:
        raise KeyError(f"unknown modulation {mod!r}; expected one of {sorted(QAM_SIDE)}")
    side = QAM_SIDE[mod]
    axis = np.arange(1 - side, side, 2, dtype=np.float64)
    pts = axis[:, None] + 1j * axis[None, :]                 # I-major ordering, like the table at sf:556-559
    return pts.reshape(-1)


def pcs_constants(mod: str, nu: float):
    """Normalised ASK levels, PCS pmf over them, re-scaled shaping factor, mean constellation power."""
    pts = qam_points(mod)
    pts = pts / np.sqrt(np.mean(np.abs(pts) ** 2))
    side = QAM_SIDE[mod]
    amps = pts.real[::side].copy()
    unit = np.min(np.abs(amps))
    nu_sc = nu / unit ** 2                                    # sf:570
    pmf = np.exp(-nu * np.abs(amps / unit) ** 2)
    pmf = pmf / pmf.sum()                                     # sf:572
    joint = np.outer(pmf, pmf)
    joint = joint / joint.sum()
    pow_mean = float(np.sum(joint.reshape(-1) * np.abs(pts) ** 2))   # sf:579
    return amps, pmf, nu_sc, pow_mean, pts, joint


def init(channel, mod, device, nu, sps, M_est, SNR):
    """Same 9-tuple as the reference's init: h_est, h_channel, P, amp_levels, amps, pol, nu_sc, var, pow_mean."""
    h_channel = upsampled_channel(channel, sps)
    amps, pmf, nu_sc, pow_mean, _, _ = pcs_constants(mod, nu)
    amp_levels = torch.tensor(amps, device=device, dtype=torch.float32)
    var = torch.full((2,), pow_mean / 10 ** (SNR / 10) / 2, device=device, dtype=torch.float32)      # sf:581
    h0 = np.zeros((2, 2, 2, M_est))
    h0[0, 0, 0, M_est // 2] = 1.0
    h0[1, 1, 0, M_est // 2] = 1.0
    h_est = torch.tensor(h0, requires_grad=True, dtype=torch.float32, device=device)
    return h_est, h_channel, pmf, amp_levels, amps, 2, nu_sc, var, pow_mean


def awgn_constants(mod: str, nu: float, SNR: float):
    """amps, P, amp_mean, var of the AWGN driver (AWGN_channel/func_VAELE_MQAM_shaping.py:252-272)."""
    amps, pmf, _, _, pts, _ = pcs_constants(mod, nu)
    w = (np.outer(pmf, pmf)).reshape(-1) * pts
    amp_mean = float(np.sum(np.abs(w.real) + np.abs(w.imag)) / 2)
    return amps, pmf, amp_mean, 10 ** (-SNR / 10)
