"""Host-side owner of one AWGN single-polarisation VAE-LE run (reference:
AWGN_channel/func_VAELE_MQAM_shaping.py, twoFIR :206-231, loss_function :63-95, Adam(amsgrad) :283-286)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .datagen import PULSE_SPAN, ROLLOFF, rrcfir

_F32 = torch.float32


class AWGNEqualizer:
    def __init__(self, M_est, sps, amp_levels, P, amp_mean, var, device="cuda", W0=None, h0=None):
        self.lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.VaeqError("AWGNEqualizer needs a CUDA device: there is no CPU path")
        self.device, self.M, self.sps = dev, int(M_est), int(sps)
        self.amp = torch.as_tensor(amp_levels, dtype=_F32).to(dev).contiguous()
        self.P = torch.as_tensor(P, dtype=_F32).to(dev).contiguous()
        self.n_lev = int(self.amp.numel())
        self.amp_mean, self.var = float(amp_mean), float(var)
        M = self.M
        if W0 is None:
            W0 = torch.zeros(1, 2, M, dtype=_F32)
            W0[0, 0, M // 2] = 1.0                         # nn.init.dirac_ on Conv1d(2,1,M)  (:210)
        if h0 is None:
            h0 = torch.zeros(2, M, dtype=_F32)
            h0[0, M // 2] = 1.0                            # :279
        self.W = torch.as_tensor(W0, dtype=_F32).detach().clone().to(dev).contiguous()
        self.h = torch.as_tensor(h0, dtype=_F32).detach().clone().to(dev).contiguous()
        self.adam = torch.zeros(int(self.lib.vaeq_adam_state_floats_awgn(M)), dtype=_F32, device=dev)
        self.gW = torch.zeros(1, 2, M, dtype=_F32, device=dev)
        self.gh = torch.zeros(2, M, dtype=_F32, device=dev)
        self.loss = torch.zeros(1, dtype=_F32, device=dev)
        self._ws, self._ws_B = None, -1

    def _desc(self, rx, q, out, B):
        if not rx.is_cuda or rx.dtype != _F32 or not rx.is_contiguous() or rx.dim() != 2 or rx.shape[0] != 2:
            raise _lib.VaeqError("rx must be a contiguous float32 CUDA tensor of shape (2, L)")
        if self._ws is None or self._ws_B < B:
            self._ws = torch.empty(int(self.lib.vaeq_awgn_workspace_bytes(B, self.M, self.n_lev)), dtype=torch.uint8, device=self.device)
            self._ws_B = B
        d = _lib.AwgnDesc()
        d.B, d.sps, d.M, d.n_lev = B, self.sps, self.M, self.n_lev
        d.amp_mean, d.var = self.amp_mean, self.var
        d.rx, d.amp, d.P = rx.data_ptr(), self.amp.data_ptr(), self.P.data_ptr()
        d.W, d.h, d.adam = self.W.data_ptr(), self.h.data_ptr(), self.adam.data_ptr()
        d.q, d.out, d.loss = q.data_ptr(), out.data_ptr(), self.loss.data_ptr()
        d.gW, d.gh = self.gW.data_ptr(), self.gh.data_ptr()
        d.workspace, d.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        return d

    @_lib.device_guard
    def _run(self, fn, name, rx, *extra):
        B = rx.shape[-1] // self.sps
        q = torch.empty(2 * self.n_lev, B, dtype=_F32, device=self.device)
        out = torch.empty(2, B, dtype=_F32, device=self.device)
        d = self._desc(rx, q, out, B)
        _lib.check(fn(C.byref(d), *extra, _lib.current_stream()), name)
        return q, out

    def forward(self, rx):
        q, out = self._run(self.lib.vaeq_awgn_forward, "vaeq_awgn_forward", rx)
        return q, out, self.loss

    def forward_backward(self, rx):
        q, out = self._run(self.lib.vaeq_awgn_forward_backward, "vaeq_awgn_forward_backward", rx)
        return q, out, self.loss, self.gW, self.gh

    def train_step(self, rx, lr_w, lr_h=None):
        lr_h = lr_w if lr_h is None else lr_h
        q, out = self._run(self.lib.vaeq_awgn_train_step, "vaeq_awgn_train_step", rx, C.c_float(lr_w), C.c_float(lr_h))
        return q, out, self.loss


def generate_data(N, M, amps, SNR, h_channel, sps, device, P, rng=None):
    """Single-polarisation test signal (reference generate_data, func_VAELE_MQAM_shaping.py:39-61)."""
    rng = np.random.default_rng() if rng is None else rng
    n_conv = N + len(h_channel) + 4 * PULSE_SPAN
    data = rng.choice(amps, (2, n_conv), p=P)
    up = np.zeros(sps * (n_conv - 1) + 1, dtype=np.complex64)
    up[::sps] = data[0] + 1j * data[1]
    sig = np.convolve(np.convolve(up, rrcfir(PULSE_SPAN, sps, ROLLOFF), mode="valid"), h_channel, mode="valid")
    sigma_n = np.sqrt(sps * np.mean(np.abs(sig) ** 2) / 2 / 10 ** (SNR / 10))
    sig = sig + sigma_n * (rng.standard_normal(sig.shape) + 1j * rng.standard_normal(sig.shape))
    rx = torch.from_numpy(np.stack((sig[:sps * N].real, sig[:sps * N].imag)).astype(np.float32)).to(device)
    sl = slice(PULSE_SPAN + M - 1, N + PULSE_SPAN + M - 1)
    tx = torch.from_numpy(np.stack((data[0, sl], data[1, sl]))).to(device, torch.float16)
    return rx.contiguous(), tx.contiguous()
