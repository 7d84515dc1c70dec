"""Persistent frame kernels (reference-size minibatches) and batched independent runs, through the C ABI.

vaeq_dp_persistent_frames(mode): 0 = one set of launches per step (the path pinned to the reference's golden vectors in
test_dp_step_gpu.py), 1 = dp_small.cu (default: state and intermediates in shared memory, fast point-wise math),
2 = the per-step kernels' bodies inside one launch.  Mode 2 must be BITWISE identical to mode 0; mode 1 must agree with
mode 0 within the step tolerances, and its first steps are also compared with the CPU oracle directly."""
import numpy as np
import pytest
import torch

from oracle import vaeq_oracle as O

pytestmark = pytest.mark.gpu


def synth_frame(mod, M, N, nu, seed, snr=20):
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", mod, "cpu", nu, 2, M, snr)
    rx, tx, _ = O.generate_data_shaping(N, amps, snr, h_ch, P, 2, 90e9, 2, -26e-24, 0.1e-12 * np.sqrt(1000),
                                        np.array([0.0314, 0.0314], dtype=np.complex64), np.pi / 10, "cpu", rng=np.random.default_rng(seed))
    return dict(P=P, amp=amp, nu_sc=nu_sc, var=var), rx, tx


def run_frame(c, rx, M, B, stride, n_steps, keep_lo, keep_n, keep_lo_in_dst, lr_w, lr_h, persistent):
    from vae_equalizer_b200 import _lib
    from vae_equalizer_b200.dp import DPEqualizer
    lib = _lib.load()
    n = c["amp"].numel()
    eq = DPEqualizer(M, 2, c["amp"], c["P"], c["var"], c["nu_sc"])
    N_keep = (n_steps - 1) * stride + keep_n + keep_lo
    ot = torch.zeros(2, 2 * n, N_keep, device="cuda")
    oc = torch.zeros(2, 2, N_keep, device="cuda")
    _lib.check(lib.vaeq_dp_persistent_frames(int(persistent)))
    try:
        n0 = int(lib.vaeq_launch_count(-1))
        loss, ve = eq.train_frame(rx.cuda(), B, stride, n_steps, lr_w, lr_h, ot, oc, keep_lo, keep_n, keep_lo_in_dst=keep_lo_in_dst)
        torch.cuda.synchronize()
        launches = int(lib.vaeq_launch_count(-1)) - n0
    finally:
        _lib.check(lib.vaeq_dp_persistent_frames(1))
    return dict(loss=loss.cpu(), ve=ve.cpu(), ot=ot.cpu(), oc=oc.cpu(), W=eq.W.cpu(), h=eq.h.cpu(), gW=eq.gW.cpu(), gh=eq.gh.cpu(),
                last_loss=eq.loss.cpu(), last_ve=eq.var_est.cpu(), steps=eq.step_count(), launches=launches)


FRAME_CASES = [  # mod, M, B, stride, n_steps, keep_lo, keep_n, keep_lo_in_dst, nu
    ("64-QAM", 25, 100, 100, 12, 0, 100, True, 0.0270955),      # VAE-LE at the reference's own batch_len (RUN_DP:38)
    ("64-QAM", 25, 100, 10, 30, 45, 10, False, 0.0),            # VAE-flex: window 100, flex_step 10 (RUN_DP:39)
    ("16-QAM", 9, 64, 64, 7, 0, 64, True, 0.0),
    ("4-QAM", 5, 48, 16, 9, 16, 16, False, 0.05),
    ("64-QAM", 13, 500, 500, 3, 0, 500, True, 0.0270955),
]


@pytest.mark.parametrize("mod,M,B,stride,n_steps,keep_lo,keep_n,kd,nu", FRAME_CASES)
def test_persistent_frame_is_bitwise_the_stepped_frame(mod, M, B, stride, n_steps, keep_lo, keep_n, kd, nu):
    c, rx, _ = synth_frame(mod, M, (n_steps - 1) * stride + B, nu, seed=B + M)
    a = run_frame(c, rx, M, B, stride, n_steps, keep_lo, keep_n, kd, 2.5e-3, 1.5e-3, persistent=2)
    b = run_frame(c, rx, M, B, stride, n_steps, keep_lo, keep_n, kd, 2.5e-3, 1.5e-3, persistent=0)
    assert a["launches"] == 1 and b["launches"] == 4 * n_steps
    assert a["steps"] == n_steps == b["steps"]
    for k in ("loss", "ve", "ot", "oc", "W", "h", "gW", "gh"):
        assert torch.equal(a[k], b[k]), k
    assert torch.isfinite(a["loss"]).all()
    # the desc's own loss / var_est outputs hold the frame's last step
    assert torch.equal(a["last_loss"].reshape(()), a["loss"][-1]) and torch.equal(a["last_ve"], a["ve"][:, -1])


def rel(x, y):
    return float((x.double() - y.double()).abs().max() / y.double().abs().max())


@pytest.mark.parametrize("mod,M,B,stride,n_steps,keep_lo,keep_n,kd,nu", FRAME_CASES + [("64-QAM", 63, 256, 256, 4, 0, 256, True, 0.0),
                                                                                 ("16-QAM", 3, 40, 8, 11, 16, 8, False, 0.0),
                                                                                 ("64-QAM", 7, 130, 130, 5, 0, 130, True, 0.02),
                                                                                 ("4-QAM", 63, 512, 512, 2, 0, 512, True, 0.0),     # largest shared-memory plan
                                                                                 ("64-QAM", 25, 28, 28, 6, 0, 28, True, 0.0),       # shortest useful minibatch
                                                                                 ("64-QAM", 25, 900, 900, 3, 0, 900, True, 0.0270955),   # between 512 and the fast path
                                                                                 ("16-QAM", 13, 1000, 250, 4, 375, 250, False, 0.0)])
def test_fast_persistent_frame_matches_stepped_frame(mod, M, B, stride, n_steps, keep_lo, keep_n, kd, nu):
    """dp_small.cu against the launch-by-launch path on the same frame: every step's loss / var_est, the kept q / out
    columns and the taps after the last step."""
    c, rx, _ = synth_frame(mod, M, (n_steps - 1) * stride + B, nu, seed=B + M)
    a = run_frame(c, rx, M, B, stride, n_steps, keep_lo, keep_n, kd, 2.5e-3, 1.5e-3, persistent=1)
    b = run_frame(c, rx, M, B, stride, n_steps, keep_lo, keep_n, kd, 2.5e-3, 1.5e-3, persistent=0)
    assert a["launches"] == 1 and a["steps"] == n_steps
    # first step: same taps on both sides -> the single-step tolerances
    n0 = keep_n + (keep_lo if kd else 0)
    assert float((a["oc"][:, :, :n0] - b["oc"][:, :, :n0]).abs().max()) < 5e-6
    assert float((a["ot"][:, :, :n0] - b["ot"][:, :, :n0]).abs().max()) < 5e-5
    assert rel(a["loss"][:1], b["loss"][:1]) < 1e-5 and rel(a["ve"][:, :1], b["ve"][:, :1]) < 1e-5
    # whole frame: the two trajectories drift apart by the taps' 1e-4 (q amplifies out by ~1/(2 var) in the logits)
    assert float((a["oc"] - b["oc"]).abs().max()) < 1e-4
    assert float((a["ot"] - b["ot"]).abs().max()) < 3e-3
    assert rel(a["loss"], b["loss"]) < 1e-4 and rel(a["ve"], b["ve"]) < 1e-4
    assert rel(a["W"], b["W"]) < 1e-4 and rel(a["h"], b["h"]) < 1e-4
    assert rel(a["gW"], b["gW"]) < 1e-3 and rel(a["gh"], b["gh"]) < 1e-3          # gradient of the LAST step, at taps that drifted by 1e-4
    assert torch.equal(a["last_loss"].reshape(()), a["loss"][-1]) and torch.equal(a["last_ve"], a["ve"][:, -1])


def test_persistent_frame_against_cpu_oracle():
    """The first steps of a VAE-LE frame at batch_len = 100 against the torch CPU port (autograd + torch Adam)."""
    M, B, n_steps, lr = 25, 100, 6, 2.5e-3
    c, rx, _ = synth_frame("64-QAM", M, n_steps * B, 0.0270955, seed=5)
    a = run_frame(c, rx, M, B, B, n_steps, 0, B, True, lr, lr, persistent=1)
    tr = O.DPTrainer(M, 2, lr)
    Pt = torch.tensor(c["P"], dtype=torch.float32)
    for m in range(n_steps):
        qo, oo, lo, vo, gWo, gho = tr.step(rx[:, :, 2 * m * B:2 * (m + 1) * B].contiguous(), c["amp"], c["var"], c["nu_sc"], Pt)
        assert float((a["oc"][:, :, m * B:(m + 1) * B] - oo).abs().max()) < 5e-6, m
        assert float((a["ot"][:, :, m * B:(m + 1) * B] - qo).abs().max()) < 5e-5, m
        assert rel(a["loss"][m], lo.reshape(())) < 1e-4 and rel(a["ve"][:, m], vo) < 1e-4, m
    assert rel(a["W"], tr.W.detach()) < 1e-4 and rel(a["h"], tr.h.detach()) < 1e-4
    assert rel(a["gW"], gWo) < 2e-4 and rel(a["gh"], gho) < 2e-4


def test_batched_runs_equal_single_runs():
    """R runs with different SNR (var), PCS (P, nu_sc), learning rates and data in ONE launch == each run on its own."""
    from vae_equalizer_b200.dp import DPEqualizerRuns
    M, B, n_steps, R = 25, 100, 8, 5
    cs, rxs = [], []
    for r in range(R):
        c, rx, _ = synth_frame("64-QAM", M, n_steps * B, [0.0, 0.0270955, 0.05, 0.0270955, 0.01][r], seed=100 + r, snr=[17, 20, 23, 26, 29][r])
        cs.append(c)
        rxs.append(rx)
    lr_w = torch.tensor([2.5e-3, 1e-3, 2.5e-3, 5e-3, 2e-3])
    lr_h = torch.tensor([2.5e-3, 2.5e-3, 1e-3, 5e-3, 3e-3])
    eqr = DPEqualizerRuns(R, M, 2, cs[0]["amp"], torch.stack([torch.as_tensor(c["P"], dtype=torch.float32) for c in cs]),
                          torch.stack([torch.as_tensor(c["var"], dtype=torch.float32) for c in cs]),
                          torch.tensor([c["nu_sc"] for c in cs], dtype=torch.float32))
    rx_all = torch.stack(rxs).cuda()
    ot = torch.zeros(R, 2, 16, n_steps * B, device="cuda")
    oc = torch.zeros(R, 2, 2, n_steps * B, device="cuda")
    from vae_equalizer_b200 import _lib
    lib = _lib.load()
    n0 = int(lib.vaeq_launch_count(-1))
    loss, ve = eqr.train_frame(rx_all, B, B, n_steps, lr_w, lr_h, ot, oc, 0, B, keep_lo_in_dst=True)
    torch.cuda.synchronize()
    assert int(lib.vaeq_launch_count(-1)) - n0 == 1
    for r in range(R):
        a = run_frame(cs[r], rxs[r], M, B, B, n_steps, 0, B, True, float(lr_w[r]), float(lr_h[r]), persistent=1)
        assert torch.equal(a["loss"], loss[r].cpu()) and torch.equal(a["ve"], ve[r].cpu()), r
        assert torch.equal(a["ot"], ot[r].cpu()) and torch.equal(a["oc"], oc[r].cpu()), r
        assert torch.equal(a["W"], eqr.W[r].cpu()) and torch.equal(a["h"], eqr.h[r].cpu()), r
        assert torch.equal(a["gW"], eqr.gW[r].cpu()) and torch.equal(a["last_loss"].reshape(()), eqr.loss[r].cpu()), r
    # a second frame continues from the state (Adam moments, step counter) of the first
    loss2, _ = eqr.train_frame(rx_all, B, B, n_steps, lr_w, lr_h, ot, oc, 0, B, keep_lo_in_dst=True)
    torch.cuda.synchronize()
    assert (loss2[:, -1] < loss[:, 0]).all()


def test_sweep_engine_equals_single_run_drivers():
    """sweep.sweep_vae_dp (all cells in one batched run set) returns, cell by cell, exactly what the single-run drivers
    processing_vaele_dp / processing_vaeflex_dp return for the same cell and seed."""
    from vae_equalizer_b200.processing import processing_vaele_dp, processing_vaeflex_dp
    from vae_equalizer_b200.sweep import sweep_vae_dp
    cells = [dict(SNR=20, nu=0.0270955, lr_optim=2.5e-3, theta=np.pi / 10, theta_diff=0.06 * np.pi, seed=3),
             dict(SNR=24, nu=0.0, lr_optim=5e-3, theta=0.2, theta_diff=0.0, seed=4),
             dict(SNR=17, nu=0.0270955, lr_optim=1e-3, theta=0.0, theta_diff=0.01, seed=5)]
    common = dict(symb_rate=90e9, tau_cd=-26e-24, tau_pmd=0.1e-12 * np.sqrt(1000), phiIQ=np.array([0.0314, 0.0314], dtype=np.complex64))
    for kind, fn, flex in (("VAE", processing_vaele_dp, 10), ("VAEflex", processing_vaeflex_dp, 20)):
        ser, ve, var = sweep_vae_dp(cells, "64-QAM", 2, 25, 100, 2000, 3, flex_step=flex, channel="h0", N_lrhalf=2, kind=kind,
                                    eval_mode="per_cell", **common)
        torch.cuda.synchronize()
        for r, c in enumerate(cells):
            s1, v1, var1 = fn("64-QAM", 2, c["SNR"], c["nu"], 25, c["theta_diff"], c["theta"], c["lr_optim"], 100, 2000, 3, flex, "h0",
                              common["symb_rate"], common["tau_cd"], common["tau_pmd"], common["phiIQ"], 2, verbose=False, datagen="gpu",
                              seed=c["seed"])
            assert torch.equal(s1, ser[r]) and torch.equal(v1, ve[r]) and torch.equal(var1, var[r]), (kind, r)
        assert torch.isfinite(ser).all() and float(ser.max()) <= 1.0


def test_run_dp_sweep_end_to_end(tmp_path):
    """Eval_run_DP.py's sweep through the batched engine (VAE, VAEflex) and cell by cell (CMAbatch): shapes, finite SERs, .mat file."""
    from vae_equalizer_b200 import sweep
    lists = dict(nu_vec=[0.0270955], symb_rate_vec=[90e9], theta_vec=[np.pi / 10], theta_diff_vec=[0.0], SNR_vec=[20, 26], M_vec=[25],
                 batch_len_vec=[100], flex_step_vec=[20], lr_optim_vec=[2.5e-3, 2e-3])
    for loss_type in ("VAE", "VAEflex", "CMAbatch"):
        SER, Var_est, var_real = sweep.run_dp_sweep(loss_type=loss_type, iter=2, num_frames=2, N_frame_max=1500, N_lrhalf=170, **lists)
        assert SER.shape == (4, 2, 1, 1, 1, 1, 2, 1, 1, 1, 2, 2)
        assert torch.isfinite(SER).all() and float(SER.min()) >= 0.0 and float(SER.max()) <= 1.0
        assert torch.isfinite(Var_est).all() and (var_real > 0).all()
        if loss_type != "CMAbatch":
            assert (Var_est > 0).all()
    sweep.save_mat(str(tmp_path / "s.mat"), SER, Var_est, var_real, SNR_vec=lists["SNR_vec"], nu_vec=lists["nu_vec"],
                   theta_diff_vec=lists["theta_diff_vec"], theta_vec=lists["theta_vec"], M_vec=lists["M_vec"], lr_optim_vec=lists["lr_optim_vec"],
                   batch_len_vec=lists["batch_len_vec"], symb_rate_vec=lists["symb_rate_vec"], flex_step_vec=lists["flex_step_vec"])


@pytest.mark.parametrize("seg_len", [0, 100])
def test_batched_frame_evaluation_equals_per_run_calls(seg_len):
    """vaeq_frame_eval_runs (roll / cut / slice as index arithmetic, shifts on the device, all runs at once) against the
    single-run sequence find_shift -> roll -> cut -> SER_IQflip / SER_constell_shaping, on runs with different time shifts,
    a polarisation swap, different SNR and PCS."""
    import vae_equalizer_b200.shared_funcs as sfun
    from vae_equalizer_b200.processing import _align
    from vae_equalizer_b200.sweep import _score
    R, N, n = 5, 3000, 8
    gen = torch.Generator().manual_seed(3)
    consts, outs, txs = [], [], []
    for r in range(R):
        k = O.init("h0", "64-QAM", "cpu", [0.0, 0.0270955, 0.05, 0.0, 0.0270955][r], 2, 25, [30, 24, 27, 21, 33][r])
        amps, P = torch.tensor(k[4], dtype=torch.float32), torch.tensor(k[2], dtype=torch.float32)
        tx = amps[torch.multinomial(P, 4 * N, True, generator=gen)].reshape(2, 2, N)
        sh, swap = [(1, 0), (-3, -3), (0, 0), (4, 4), (2, -5)][r], [0, 1, 0, 1, 0][r]
        y = tx + float(k[7][0].sqrt()) * torch.randn(2, 2, N, generator=gen)
        y = torch.stack((y[0].roll(sh[0], -1), y[1].roll(sh[1], -1)))          # delay per pol ...
        if swap:
            y = y.roll(1, 0)                                                    # ... and crossed polarisations
        consts.append(k)
        outs.append(y)
        txs.append(tx.to(torch.float16))
    amp = consts[0][3].cuda()
    out_const = torch.stack(outs).cuda().contiguous()
    tx = torch.stack(txs).cuda().contiguous()
    var = torch.stack([k[7] for k in consts]).cuda()
    nu_sc = torch.tensor([k[6] for k in consts], dtype=torch.float32).cuda()
    out_train = torch.stack([sfun.soft_dec(out_const[r], var[r], amp, float(nu_sc[r])) for r in range(R)])
    ser, align, counts = sfun.frame_eval_runs(out_train, out_const, tx, amp, var, nu_sc, seg_len, return_counts=True)
    torch.cuda.synchronize()
    kind, B, m_max = ("VAE", seg_len, N // seg_len) if seg_len else ("VAEflex", 100, N)
    for r in range(R):
        sq, rq = sfun.find_shift(out_train[r], tx[r], 21, amp, 2)
        so, ro = sfun.find_shift_symb_full(out_const[r], tx[r], 21)
        sq, so = [int(v) for v in sq.tolist()], [int(v) for v in so.tolist()]
        assert align[r, 0, :3].tolist() == [sq[0], sq[1], rq] and align[r, 1, :3].tolist() == [so[0], so[1], ro], r
        s_q = _score(kind, "q", _align(out_train[r].clone(), sq, rq), tx[r], sq, B, m_max, 2, n, amp, consts[r])
        s_c = _score(kind, "c", _align(out_const[r].clone(), so, ro), tx[r], so, B, m_max, 2, n, amp, (None,) * 6 + (float(nu_sc[r]), var[r]))
        n_q, n_c = int(align[r, 0, 3]), int(align[r, 1, 3])
        assert torch.equal(s_q.cpu(), ser[r, 2:].cpu()), (r, s_q, ser[r])                     # integer decisions: exact
        assert torch.equal(s_c.cpu(), ser[r, :2].cpu()), (r, s_c, ser[r])                     # fixed-order norm sums: exact as well
        assert 0 < n_q <= N and 0 < n_c <= N
    assert float(ser.max()) < 0.2                                                             # aligned correctly: a misaligned run scores ~0.98


def test_cuda_data_generator_against_the_torch_formulation():
    """csrc/datagen.cu (vaeq_gen_levels / _pulse / _jones / _noise): the noise-free signal equals the torch.fft formulation fed with
    the same amplitude levels; the levels follow the run's pmf; the noise is white, unit-variance Gaussian scaled by sigma_n;
    a run's data does not depend on the batch it is generated in."""
    from vae_equalizer_b200 import datagen as dg
    N, sps = 6000, 2
    c64 = O.init("h0", "64-QAM", "cpu", 0.0270955, 2, 25, 23)
    c0 = O.init("h0", "64-QAM", "cpu", 0.0, 2, 25, 23)
    amps, P = c64[4], np.stack([np.asarray(c64[2]), np.asarray(c0[2]), np.asarray(c64[2])]).astype(np.float32)
    snr, theta = [23.0, 13.0, 30.0], [0.3, 1.0, -0.7]
    args = (90e9, -26e-24, 0.1e-12 * np.sqrt(1000), np.array([0.0314, 0.0314], dtype=np.complex64))
    Pt = torch.as_tensor(P, device="cuda")
    rx, tx, sigma, lev, sig = dg._generate_frames_cuda(N, amps, snr, Pt, theta, torch.device("cuda"), 41, *args, return_parts=True)
    assert rx.shape == (3, 2, 2, sps * N) and tx.shape == (3, 2, 2, N) and tx.dtype == torch.float16
    # tx is the slice of the drawn levels the reference returns (sf:89)
    for p in range(2):
        for c in range(2):
            assert torch.equal(tx[:, p, c], lev[:, 2 * p + c, dg.PULSE_SPAN:dg.PULSE_SPAN + N].to(torch.float16))
    # deterministic part against the torch.fft formulation on the same levels
    sig_t = dg._frames_torch_from_levels(lev, N, sps, theta, *args)
    assert sig_t.shape == sig.shape
    assert float((sig - sig_t).abs().max()) < 3e-5 * float(sig_t.abs().max())
    # sigma_n (sf:83)
    pw = (sig_t.abs() ** 2).mean(dim=(1, 2)).cpu().numpy()
    assert np.allclose(sigma.cpu().numpy(), np.sqrt(pw * sps / 2 / 10 ** (np.asarray(snr) / 10)), rtol=1e-4)
    # the levels follow the pmf of the run
    a = np.asarray(amps, dtype=np.float32)
    for r in range(3):
        x = lev[r].cpu().numpy().ravel()
        cnt = np.array([(x == a[l]).sum() for l in range(len(a))])
        assert cnt.sum() == x.size
        exp = P[r] / P[r].sum() * x.size
        assert np.all(np.abs(cnt - exp) < 5 * np.sqrt(exp) + 1)
    # noise: zero mean, unit variance, Gaussian kurtosis, I/Q and neighbouring samples uncorrelated
    clean = torch.stack((sig.real, sig.imag), dim=2)[..., :sps * N]
    z = ((rx - clean) / sigma[:, None, None, None]).double()
    n_el = z[0].numel()
    for r in range(3):
        zr = z[r]
        assert abs(float(zr.mean())) < 5 / np.sqrt(n_el) and abs(float(zr.var()) - 1) < 5 * np.sqrt(2 / n_el)
        assert abs(float((zr ** 4).mean()) - 3) < 0.2
        assert abs(float((zr[:, 0] * zr[:, 1]).mean())) < 5 / np.sqrt(n_el / 2)
        assert abs(float((zr[..., 1:] * zr[..., :-1]).mean())) < 5 / np.sqrt(n_el)
    # same seed -> same data; run 0 alone = run 0 of the batch; another seed differs
    rx2, tx2, _ = dg.generate_frames_gpu(N, amps, snr, P, sps, theta, "cuda", 41, *args)
    rx1, tx1, _ = dg.generate_frames_gpu(N, amps, snr[:1], P[:1], sps, theta[:1], "cuda", 41, *args)
    rx3, _, _ = dg.generate_frames_gpu(N, amps, snr, P, sps, theta, "cuda", 42, *args)
    assert torch.equal(rx2, rx) and torch.equal(tx2, tx) and torch.equal(rx1[0], rx[0]) and torch.equal(tx1[0], tx[0])
    assert not torch.equal(rx3, rx)
    # and the statistics of the full signal agree with the torch generator on the GPU
    dg._FORCE_TORCH = True
    try:
        rxt, txt, sgt = dg.generate_frames_gpu(N, amps, snr, P, sps, theta, "cuda", 41, *args)
    finally:
        dg._FORCE_TORCH = False
    assert np.allclose(sgt.cpu().numpy(), sigma.cpu().numpy(), rtol=0.05)
    assert np.allclose(rxt.double().var(dim=(1, 2, 3)).cpu().numpy(), rx.double().var(dim=(1, 2, 3)).cpu().numpy(), rtol=0.05)


@pytest.mark.parametrize("kind", ["CMA", "CMAbatch", "CMAflex"])
def test_cma_sweep_engine_equals_single_run_drivers(kind):
    """sweep.sweep_cma_dp (all cells per equalizer / CPE launch, batched CMA streams) against processing_cma*_dp cell by cell: a cell's
    SER estimates are identical alone and inside a batch (datagen='gpu': its data depends on its own seed only)."""
    from vae_equalizer_b200 import processing as pr, sweep
    fn = {"CMA": pr.processing_cma_dp, "CMAbatch": pr.processing_cmabatch_dp, "CMAflex": pr.processing_cmaflex_dp}[kind]
    lr = {"CMA": 1e-3, "CMAbatch": 1e-5, "CMAflex": 1e-6}[kind]
    phiIQ = np.array([0.0314, 0.0314], dtype=np.complex64)
    cells = [dict(SNR=20, nu=0.0270955, lr_optim=lr, theta=0.3, theta_diff=0.05, seed=3),
             dict(SNR=24, nu=0.0, lr_optim=lr, theta=0.1, theta_diff=0.0, seed=4),
             dict(SNR=22, nu=0.0270955, lr_optim=2 * lr, theta=-0.2, theta_diff=0.1, seed=5)]
    M, B, N, F, step, half = 25, 100, 3000, 3, 20, 2
    common = ("h0", 90e9, -26e-24, 0.1e-12 * np.sqrt(1000), phiIQ, half)
    ser_b, ve_b, var_b = sweep.sweep_cma_dp(cells, "64-QAM", 2, M, B, N, F, step, *common, kind=kind, datagen="gpu")      # batched evaluation
    assert ser_b.shape == (3, 4, F) and ve_b.shape == (3, 2, F) and float(ve_b.abs().max()) == 0.0
    ser_p, _, _ = sweep.sweep_cma_dp(cells, "64-QAM", 2, M, B, N, F, step, *common, kind=kind, datagen="gpu", eval_mode="per_cell")
    assert torch.equal(ser_p, ser_b), (ser_p.tolist(), ser_b.tolist())
    for r, c in enumerate(cells):
        s, v, var = fn("64-QAM", 2, c["SNR"], c["nu"], M, c["theta_diff"], c["theta"], c["lr_optim"], B, N, F, step, *common,
                       verbose=False, datagen="gpu", seed=c["seed"])
        assert torch.equal(s, ser_b[r]), (kind, r, s.tolist(), ser_b[r].tolist())
        assert torch.equal(var.to(var_b.device), var_b[r])
    # one batched generation for all cells: runs, finite, same shapes
    ser_g, _, _ = sweep.sweep_cma_dp(cells, "64-QAM", 2, M, B, N, 2, step, *common, kind=kind, datagen="gpu_batched")
    assert ser_g.shape == (3, 4, 2) and bool(torch.isfinite(ser_g).all())


def test_frame_kernel_three_and_four_runs_per_sm_are_bitwise_identical():
    """k_dp_frame_fast is built for 3 and for 4 resident runs per SM (80 / 64 registers); the library picks by the number of CTA waves.
    Same arithmetic: the results must not depend on the choice."""
    from vae_equalizer_b200 import _lib
    from vae_equalizer_b200.datagen import generate_frames_gpu
    from vae_equalizer_b200.dp import DPEqualizerRuns
    lib = _lib.load()
    R, B, n_steps, M = 5, 100, 12, 25
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", "64-QAM", "cpu", 0.0270955, 2, M, 23)
    rxs, _, _ = generate_frames_gpu(B * n_steps, amps, [23.0] * R, np.tile(np.asarray(P, dtype=np.float32)[None], (R, 1)), 2, [0.3] * R, "cuda", 9)
    res = {}
    try:
        for per_sm in (3, 4):
            _lib.check(lib.vaeq_dp_frame_runs_per_sm(per_sm), "vaeq_dp_frame_runs_per_sm")
            eq = DPEqualizerRuns(R, M, 2, amp, torch.tensor(P, dtype=torch.float32), var, nu_sc, device="cuda")
            ot, oc = torch.empty(R, 2, 16, B * n_steps, device="cuda"), torch.empty(R, 2, 2, B * n_steps, device="cuda")
            loss, ve = eq.train_frame(rxs, B, B, n_steps, 2.5e-3, 2.5e-3, ot, oc, 0, B, keep_lo_in_dst=True)
            torch.cuda.synchronize()
            res[per_sm] = (eq.W.clone(), eq.h.clone(), ot, oc, loss.clone(), ve.clone())
    finally:
        _lib.check(lib.vaeq_dp_frame_runs_per_sm(0), "vaeq_dp_frame_runs_per_sm")
    for a, b in zip(res[3], res[4]):
        assert torch.equal(a, b)
    assert lib.vaeq_dp_frame_runs_per_sm(5) != 0


@pytest.mark.parametrize("kind", ["VAE", "VAEflex"])
def test_single_run_drivers_fused_evaluation_equals_per_op(kind):
    """processing_vaele_dp / processing_vaeflex_dp with eval_mode="fused" (one batched evaluation call per frame, no host sync) against the
    reference's call-by-call evaluation: identical SER estimates."""
    from vae_equalizer_b200 import processing as pr
    phiIQ = np.array([0.0314, 0.0314], dtype=np.complex64)
    fn = pr.processing_vaele_dp if kind == "VAE" else pr.processing_vaeflex_dp
    args = ("64-QAM", 2, 22, 0.0270955, 25, 0.05, 0.3, 2.5e-3, 100, 3000 if kind == "VAE" else 1000, 3, 20, "h0", 90e9, -26e-24,
            0.1e-12 * np.sqrt(1000), phiIQ, 2)
    a = fn(*args, verbose=False, datagen="gpu", seed=6)
    b = fn(*args, verbose=False, datagen="gpu", seed=6, eval_mode="fused")
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
