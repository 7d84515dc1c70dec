"""Host-side owner of one DP VAE-LE / VAE-flex run: taps, channel estimate, Adam state, scratch.

Mirrors what func_VAELE_DP_MQAM_shaping.processing builds at lines 21-36 of the reference
(net = twoXtwoFIR, optimizer = Adam(net.parameters()) + add_param_group(h_est)) and steps it with
the fused CUDA path (vaeq_dp_* in include/vaeq.h).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_F32 = torch.float32


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise _lib.VaeqError(f"{name} must be a CUDA tensor: vae_equalizer_b200 has no CPU path")
    if t.dtype != _F32:
        raise _lib.VaeqError(f"{name} must be float32, got {t.dtype}")
    _lib.require_current_device(t, name)


class DPEqualizer:
    """Trainable state of one run + the fused step.

    W (2,4,M): twoXtwoFIR.conv_w.weight, Dirac-initialised (sf:494-495);
    h (2,2,2,M): h_est, Dirac-initialised (sf:583-585).
    """

    def __init__(self, M_est: int, sps: int, amp_levels, P, var, nu_sc: float, device="cuda", W0=None, h0=None,
                 amsgrad=False):
        self.lib = _lib.load()
        self.M, self.sps, self.nu_sc = int(M_est), int(sps), float(nu_sc)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.VaeqError("DPEqualizer needs a CUDA device: there is no CPU path")
        self.device = dev
        self.amp = torch.as_tensor(amp_levels, dtype=_F32).to(dev).contiguous()
        self.P = torch.as_tensor(P, dtype=_F32).to(dev).contiguous()
        self.var = torch.as_tensor(var, dtype=_F32).to(dev).contiguous()
        self.n_lev = int(self.amp.numel())
        M = self.M
        if W0 is None:
            W0 = torch.zeros(2, 4, M, dtype=_F32)
            W0[0, 0, M // 2] = 1.0
            W0[1, 1, M // 2] = 1.0
        if h0 is None:
            h0 = torch.zeros(2, 2, 2, M, dtype=_F32)
            h0[0, 0, 0, M // 2] = 1.0
            h0[1, 1, 0, M // 2] = 1.0
        self.W = torch.as_tensor(W0, dtype=_F32).detach().clone().to(dev).contiguous()
        self.h = torch.as_tensor(h0, dtype=_F32).detach().clone().to(dev).contiguous()
        self.adam = torch.zeros(int(self.lib.vaeq_adam_state_floats(M)), dtype=_F32, device=dev)
        self.flags = 1 if amsgrad else 0
        self.gW = torch.zeros(2, 4, M, dtype=_F32, device=dev)
        self.gh = torch.zeros(2, 2, 2, M, dtype=_F32, device=dev)
        self.loss = torch.zeros(1, dtype=_F32, device=dev)
        self.var_est = torch.zeros(2, dtype=_F32, device=dev)
        self._ws = None
        self._ws_B = -1

    # -- scratch ---------------------------------------------------------------------------------
    def _workspace(self, B: int):
        if self._ws is None or self._ws_B < B:
            nbytes = int(self.lib.vaeq_dp_workspace_bytes(B, self.M, self.n_lev))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_B = B
        return self._ws

    def step_count(self) -> int:
        return int(self.adam[48 * self.M:48 * self.M + 1].view(torch.int32).item())

    def _desc(self, rx, q, out, B, ld_rx=None, q_keep=None, out_keep=None, keep_lo=0, keep_n=0, col0=0):
        """col0 != 0: q / out hold the columns [col0, col0 + q.shape[-1]) of the minibatch only (batch-split ranks); the descriptor
        then carries base pointers shifted back by col0 columns -- the library touches no column outside the rank's range."""
        _require_cuda(rx, "rx")
        ws = self._workspace(B)
        d = _lib.DpDesc()
        d.B, d.sps, d.M, d.n_lev = B, self.sps, self.M, self.n_lev
        d.nu_sc, d.flags = self.nu_sc, self.flags
        d.rx, d.ld_rx = rx.data_ptr(), int(rx.stride(1) if ld_rx is None else ld_rx)
        d.amp, d.P, d.var = self.amp.data_ptr(), self.P.data_ptr(), self.var.data_ptr()
        d.W, d.h, d.adam = self.W.data_ptr(), self.h.data_ptr(), self.adam.data_ptr()
        if q is not None:
            d.q, d.ld_q = q.data_ptr() - 4 * int(col0), int(q.stride(1))
            d.out, d.ld_out = out.data_ptr() - 4 * int(col0), int(out.stride(1))
        if q_keep is not None:
            d.q_keep, d.ld_q_keep = q_keep.data_ptr(), int(q_keep.stride(1))
            d.out_keep, d.ld_out_keep = out_keep.data_ptr(), int(out_keep.stride(1))
            d.keep_lo, d.keep_n = int(keep_lo), int(keep_n)
        d.loss, d.var_est = self.loss.data_ptr(), self.var_est.data_ptr()
        d.gW, d.gh = self.gW.data_ptr(), self.gh.data_ptr()
        d.workspace, d.workspace_bytes = ws.data_ptr(), ws.numel()
        return d

    def _check_rx(self, rx):
        if rx.dim() != 3 or rx.shape[0] != 2 or rx.shape[1] != 2 or rx.stride(2) != 1 or rx.stride(0) != 2 * rx.stride(1):
            raise _lib.VaeqError(f"rx must be (2,2,L) with unit time stride, got {tuple(rx.shape)} strides {rx.stride()}")
        return rx.shape[2] // self.sps

    def _alloc_out(self, B):
        q = torch.empty(2, 2 * self.n_lev, B, dtype=_F32, device=self.device)
        out = torch.empty(2, 2, B, dtype=_F32, device=self.device)
        return q, out

    # -- the three entry points --------------------------------------------------------------------
    @_lib.device_guard
    def forward(self, rx, q=None, out=None):
        """net(minibatch) + loss_function_shaping without gradients -> (q, out, loss(1), var_est(2))."""
        B = self._check_rx(rx)
        if q is None:
            q, out = self._alloc_out(B)
        d = self._desc(rx, q, out, B)
        _lib.check(self.lib.vaeq_dp_forward(C.byref(d), _lib.current_stream()), "vaeq_dp_forward")
        return q, out, self.loss, self.var_est

    @_lib.device_guard
    def forward_backward(self, rx, q=None, out=None):
        """As forward, plus gW (2,4,M) and gh (2,2,2,M); parameters untouched."""
        B = self._check_rx(rx)
        if q is None:
            q, out = self._alloc_out(B)
        d = self._desc(rx, q, out, B)
        _lib.check(self.lib.vaeq_dp_forward_backward(C.byref(d), _lib.current_stream()), "vaeq_dp_forward_backward")
        return q, out, self.loss, self.var_est, self.gW, self.gh

    @_lib.device_guard
    def train_step(self, rx, lr_w, lr_h=None, q=None, out=None):
        """One optimizer step (VAELE_DP:59-66).  lr_h defaults to lr_w."""
        B = self._check_rx(rx)
        if q is None:
            q, out = self._alloc_out(B)
        d = self._desc(rx, q, out, B)
        lr_h = lr_w if lr_h is None else lr_h
        _lib.check(self.lib.vaeq_dp_train_step(C.byref(d), float(lr_w), float(lr_h), _lib.current_stream()),
                   "vaeq_dp_train_step")
        return q, out, self.loss, self.var_est

    @_lib.device_guard
    def capture_steps(self, rx_list, lr_w, lr_h=None, q=None, out=None):
        """CUDA graph of len(rx_list) consecutive train_step calls (one per rx tensor, in order): `g.replay()` re-issues the
        6 kernels of every step without the per-launch host work (profiles/r01d: 559 -> 547 us per step at batch_len 2^22).
        The step counter of Adam lives on the device, so a replayed step is a new optimizer step; lr_w / lr_h are baked into
        the graph.  Call train_step at least once with the same shapes before capturing (workspaces, kernel attributes)."""
        for rx in rx_list:
            B = self._check_rx(rx)
        if q is None:
            q, out = self._alloc_out(B)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for rx in rx_list:
                self.train_step(rx, lr_w, lr_h, q=q, out=out)
        g.outputs = (q, out, self.loss, self.var_est)
        return g

    @_lib.device_guard
    def train_frame(self, rx_frame, batch_len, stride_sym, n_steps, lr_w, lr_h, out_train, out_const, keep_lo, keep_n,
                    keep_lo_in_dst=False):
        """All minibatches of one frame (VAELE_DP:57-66 / VAEflex_DP:59-70).

        Returns (loss_steps (n_steps,), var_est_steps (2,n_steps)) device tensors."""
        self._check_rx(rx_frame)
        B = int(batch_len)
        if not hasattr(self, "_q_scratch") or self._q_scratch.shape[-1] != B:
            self._q_scratch, self._out_scratch = self._alloc_out(B)
        d = self._desc(rx_frame, self._q_scratch, self._out_scratch, B, ld_rx=rx_frame.stride(1), q_keep=out_train,
                       out_keep=out_const, keep_lo=keep_lo, keep_n=keep_n)
        loss_steps = torch.empty(n_steps, dtype=_F32, device=self.device)
        var_steps = torch.empty(2, n_steps, dtype=_F32, device=self.device)
        _lib.check(self.lib.vaeq_dp_train_frame(C.byref(d), int(n_steps), int(stride_sym), 1 if keep_lo_in_dst else 0,
                                                float(lr_w), float(lr_h), loss_steps.data_ptr(), var_steps.data_ptr(),
                                                _lib.current_stream()), "vaeq_dp_train_frame")
        return loss_steps, var_steps

    # -- batch-split phases (vaeq_dp_split_*; the caller all-reduces `stats` and `grads` between them) ---------------
    def _check_split_cols(self, B, sym_lo, sym_hi, q, out, col0):
        if q is None and out is None:           # only the kept columns are wanted (frame loops): no per-window q / out at all
            return
        clo, chi = max(0, sym_lo - 16), min(B, sym_hi + 16)
        if col0 % 4 or col0 > clo or col0 + q.shape[-1] < chi or col0 + out.shape[-1] < chi:
            raise _lib.VaeqError(f"q / out hold columns [{col0}, {col0 + q.shape[-1]}) but the rank writes [{clo}, {chi})")

    @_lib.device_guard
    def split_forward(self, rx, sym_lo, sym_hi, q, out, col0=0, q_keep=None, out_keep=None, keep_lo=0, keep_n=0):
        B = self._check_rx(rx)
        self._check_split_cols(B, sym_lo, sym_hi, q, out, col0)
        d = self._desc(rx, q, out, B, q_keep=q_keep, out_keep=out_keep, keep_lo=keep_lo, keep_n=keep_n, col0=col0)
        if not hasattr(self, "_stats"):
            self._stats = torch.zeros(int(self.lib.vaeq_dp_split_stats_doubles(self.M)), dtype=torch.float64, device=self.device)
            self._grads = torch.zeros(16 * self.M, dtype=_F32, device=self.device)
        _lib.check(self.lib.vaeq_dp_split_forward(C.byref(d), int(sym_lo), int(sym_hi), self._stats.data_ptr(),
                                                  _lib.current_stream()), "vaeq_dp_split_forward")
        return self._stats

    @_lib.device_guard
    def split_backward(self, rx, sym_lo, sym_hi, q, out, col0=0):
        B = self._check_rx(rx)
        d = self._desc(rx, q, out, B, col0=col0)
        _lib.check(self.lib.vaeq_dp_split_backward(C.byref(d), int(sym_lo), int(sym_hi), self._stats.data_ptr(),
                                                   self._grads.data_ptr(), _lib.current_stream()), "vaeq_dp_split_backward")
        return self._grads

    @_lib.device_guard
    def split_update(self, rx, q, out, lr_w, lr_h, col0=0):
        B = self._check_rx(rx)
        d = self._desc(rx, q, out, B, col0=col0)
        _lib.check(self.lib.vaeq_dp_split_update(C.byref(d), self._grads.data_ptr(), float(lr_w), float(lr_h),
                                                 _lib.current_stream()), "vaeq_dp_split_update")

    @_lib.device_guard
    def split_step_peer(self, rx, sym_lo, sym_hi, q, out, comm, lr_w, lr_h, col0=0, q_keep=None, out_keep=None, keep_lo=0, keep_n=0):
        """The whole batch-split step in one call, both reductions over NVLink peer memory (vaeq_dp_split_step_peer)."""
        B = self._check_rx(rx)
        self._check_split_cols(B, sym_lo, sym_hi, q, out, col0)
        d = self._desc(rx, q, out, B, q_keep=q_keep, out_keep=out_keep, keep_lo=keep_lo, keep_n=keep_n, col0=col0)
        _lib.check(self.lib.vaeq_dp_split_step_peer(C.byref(d), int(sym_lo), int(sym_hi), C.byref(comm), float(lr_w), float(lr_h),
                                                    _lib.current_stream()), "vaeq_dp_split_step_peer")


class DPEqualizerRuns:
    """R independent DP VAE runs (sweep cells of Eval_run_DP.py:68-95 that share batch_len, M_est and the modulation
    order) stepped together: one persistent launch per frame, one CTA per run (vaeq_dp_train_frame_runs).

    Per-run state has a leading run dimension: W (R,2,4,M), h (R,2,2,2,M), var (R,2), P (R,n), nu_sc (R), lr_w / lr_h (R).
    """

    def __init__(self, n_runs, M_est, sps, amp_levels, P, var, nu_sc, device="cuda", amsgrad=False):
        self.lib = _lib.load()
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.VaeqError("DPEqualizerRuns needs a CUDA device: there is no CPU path")
        self.device, self.R, self.M, self.sps = dev, int(n_runs), int(M_est), int(sps)
        R, M = self.R, self.M
        self.amp = torch.as_tensor(amp_levels, dtype=_F32).to(dev).contiguous()
        self.n_lev = int(self.amp.numel())

        def per_run(x, tail):
            t = torch.as_tensor(x, dtype=_F32).to(dev)
            return t.expand(R, *tail).contiguous() if t.dim() == len(tail) else t.reshape(R, *tail).contiguous()

        self.P = per_run(P, (self.n_lev,))
        self.var = per_run(var, (2,))
        self.nu_sc = per_run(nu_sc, ())
        self.W = torch.zeros(R, 2, 4, M, dtype=_F32, device=dev)
        self.W[:, 0, 0, M // 2] = 1.0
        self.W[:, 1, 1, M // 2] = 1.0
        self.h = torch.zeros(R, 2, 2, 2, M, dtype=_F32, device=dev)
        self.h[:, 0, 0, 0, M // 2] = 1.0
        self.h[:, 1, 1, 0, M // 2] = 1.0
        self.n_adam = int(self.lib.vaeq_adam_state_floats(M))
        self.adam = torch.zeros(R, self.n_adam, dtype=_F32, device=dev)
        self.flags = 1 if amsgrad else 0
        self.gW = torch.zeros(R, 2, 4, M, dtype=_F32, device=dev)
        self.gh = torch.zeros(R, 2, 2, 2, M, dtype=_F32, device=dev)
        self.loss = torch.zeros(R, dtype=_F32, device=dev)
        self.var_est = torch.zeros(R, 2, dtype=_F32, device=dev)
        self._ws = None
        self._B = -1

    @_lib.device_guard
    def train_frame(self, rx_frames, batch_len, stride_sym, n_steps, lr_w, lr_h, out_train, out_const, keep_lo, keep_n,
                    keep_lo_in_dst=False):
        """rx_frames (R,2,2,L_frame); out_train (R,2,2n,N_keep), out_const (R,2,2,N_keep); lr_w / lr_h: floats or (R,) tensors.
        Returns (loss_steps (R,n_steps), var_est_steps (R,2,n_steps))."""
        R, B, dev = self.R, int(batch_len), self.device
        _require_cuda(rx_frames, "rx_frames")
        if rx_frames.dim() != 4 or rx_frames.shape[:3] != (R, 2, 2) or not rx_frames.is_contiguous():
            raise _lib.VaeqError(f"rx_frames must be contiguous (R,2,2,L), got {tuple(rx_frames.shape)}")
        if not (out_train.is_contiguous() and out_const.is_contiguous()):
            raise _lib.VaeqError("out_train / out_const must be contiguous")
        if self._B != B:
            self._ws = torch.empty(int(self.lib.vaeq_dp_runs_workspace_bytes(B, self.M, self.n_lev, R)), dtype=torch.uint8, device=dev)
            self._q = torch.empty(R, 2, 2 * self.n_lev, B, dtype=_F32, device=dev)
            self._out = torch.empty(R, 2, 2, B, dtype=_F32, device=dev)
            self._B = B
        lrw = torch.as_tensor(lr_w, dtype=_F32).to(dev).expand(R).contiguous()
        lrh = torch.as_tensor(lr_h, dtype=_F32).to(dev).expand(R).contiguous()
        d = _lib.DpDesc()
        d.B, d.sps, d.M, d.n_lev = B, self.sps, self.M, self.n_lev
        d.nu_sc, d.flags = 0.0, self.flags
        d.rx, d.ld_rx = rx_frames.data_ptr(), int(rx_frames.stride(2))
        d.amp, d.P, d.var = self.amp.data_ptr(), self.P.data_ptr(), self.var.data_ptr()
        d.W, d.h, d.adam = self.W.data_ptr(), self.h.data_ptr(), self.adam.data_ptr()
        d.q, d.ld_q = self._q.data_ptr(), B
        d.out, d.ld_out = self._out.data_ptr(), B
        d.q_keep, d.ld_q_keep = out_train.data_ptr(), int(out_train.stride(2))
        d.out_keep, d.ld_out_keep = out_const.data_ptr(), int(out_const.stride(2))
        d.keep_lo, d.keep_n = int(keep_lo), int(keep_n)
        d.loss, d.var_est = self.loss.data_ptr(), self.var_est.data_ptr()
        d.gW, d.gh = self.gW.data_ptr(), self.gh.data_ptr()
        d.workspace, d.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        r = _lib.DpRuns()
        r.n_runs = R
        r.rs_rx, r.rs_amp, r.rs_P, r.rs_var = int(rx_frames.stride(0)), 0, self.n_lev, 2
        r.rs_W, r.rs_h, r.rs_adam = 8 * self.M, 8 * self.M, self.n_adam
        r.rs_q, r.rs_out = int(self._q.stride(0)), int(self._out.stride(0))
        r.rs_q_keep, r.rs_out_keep = int(out_train.stride(0)), int(out_const.stride(0))
        r.nu_sc, r.lr_w, r.lr_h = self.nu_sc.data_ptr(), lrw.data_ptr(), lrh.data_ptr()
        loss_steps = torch.empty(R, n_steps, dtype=_F32, device=dev)
        var_steps = torch.empty(R, 2, n_steps, dtype=_F32, device=dev)
        _lib.check(self.lib.vaeq_dp_train_frame_runs(C.byref(d), C.byref(r), int(n_steps), int(stride_sym),
                                                     1 if keep_lo_in_dst else 0, 0.0, 0.0, loss_steps.data_ptr(),
                                                     var_steps.data_ptr(), _lib.current_stream()), "vaeq_dp_train_frame_runs")
        self._keepalive = (lrw, lrh)
        return loss_steps, var_steps
