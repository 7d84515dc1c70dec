"""GPU parity (through the C ABI / the shared_funcs mirror) of the evaluation kernels, the CMA family,
CPE and the AWGN variant against the golden vectors generated from the reference and the CPU oracle.
Integer work (decisions, error counts, shifts) must be bit-exact; floats within the stated tolerances."""
import os

import numpy as np
import pytest
import torch

from oracle import vaeq_oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
T = torch.from_numpy


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", ["eval_64qam_N2000", "eval_16qam_N1500_swap"])
def test_eval_against_reference_golden(name):
    import vae_equalizer_b200.shared_funcs as sfun
    g = load(name)
    amp, var, nu_sc = T(g["amp"]).cuda(), T(g["var"]).cuda(), float(g["nu_sc"])
    q = sfun.soft_dec(T(g["out"]).cuda(), var, amp, nu_sc)
    assert np.abs(q.cpu().numpy() - g["q"]).max() < 2e-5
    tx = T(g["tx"]).cuda()
    s, r = sfun.find_shift(T(g["q"]).cuda(), tx, 21, amp, 2)
    assert s.dtype == torch.int16 and s.tolist() == g["shift_q"].tolist() and r == int(g["r_q"])
    s, r, corr = sfun.find_shift_symb_full(T(g["out"]).cuda(), tx, 21, return_corr=True)
    assert s.tolist() == g["shift_c"].tolist() and r == int(g["r_c"])
    _, _, corr_o = O.find_shift_symb_full(T(g["out"]), T(g["tx"]), 21, return_corr=True)
    assert rel(corr, corr_o) < 1e-5
    # SER from posteriors: counts bit-exact against the oracle, SER equal to the reference's float
    ser, counts = sfun.SER_IQflip(T(g["q_cut"]).cuda(), T(g["tx_cut_q"]).cuda(), return_counts=True)
    counts_o, N = O.ser_iqflip_counts(T(g["q_cut"]), T(g["tx_cut_q"]))
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), counts_o.numpy())
    assert np.array_equal(ser.cpu().numpy(), g["ser_q"])
    # SER from the constellation: in-place rescale reproduced, counts bit-exact
    rx = T(g["o_cut_in"].copy()).cuda()
    ser, counts = sfun.SER_constell_shaping(rx, T(g["tx_cut_c"]).cuda(), amp, nu_sc, var, return_counts=True)
    counts_o, N = O.ser_constell_counts(T(g["o_cut_in"].copy()), T(g["tx_cut_c"]), T(g["amp"]), nu_sc, T(g["var"]))
    assert np.array_equal(counts.cpu().numpy().astype(np.int64), counts_o.numpy())
    assert np.array_equal(ser.cpu().numpy(), g["ser_c"])
    assert np.abs(rx.cpu().numpy() - g["o_cut_scaled"]).max() < 1e-6
    # strided views (what the drivers pass): same answer as the contiguous call
    big = torch.zeros(2, 16 if name.startswith("eval_64") else 8, g["q_cut"].shape[-1] + 30, device="cuda")
    big[:, :, 7:7 + g["q_cut"].shape[-1]] = T(g["q_cut"]).cuda()
    ser2 = sfun.SER_IQflip(big[:, :, 7:7 + g["q_cut"].shape[-1]], T(g["tx_cut_q"]).cuda())
    assert np.array_equal(ser2.cpu().numpy(), g["ser_q"])
    # GMI extension against its CPU definition
    gmi = sfun.GMI(T(g["q_cut"]).cuda(), T(g["tx_cut_q"]).cuda(), g["P"])
    gmi_o = O.gmi_from_posteriors(T(g["q_cut"]), T(g["tx_cut_q"]), T(g["P"]))
    assert np.abs(gmi.cpu().numpy() - gmi_o.numpy()).max() < 1e-4


def test_ser_large_random_counts_bit_exact():
    import vae_equalizer_b200.shared_funcs as sfun
    gen = torch.Generator().manual_seed(3)
    N, n = 300001, 8
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", "64-QAM", "cpu", 0.0270955, 2, 25, 16)
    idx = torch.randint(0, n, (2, 2, N), generator=gen)
    tx = amp[idx].to(torch.float16)
    out = (amp[idx] + 0.05 * torch.randn(2, 2, N, generator=gen)) * 1.1
    q = O.soft_demap(out, var, amp, nu_sc)
    _, c = sfun.SER_IQflip(q.cuda(), tx.cuda(), return_counts=True)
    co, _ = O.ser_iqflip_counts(q, tx)
    assert np.array_equal(c.cpu().numpy().astype(np.int64), co.numpy())
    rx = out.clone().cuda()
    _, c = sfun.SER_constell_shaping(rx, tx.cuda(), amp.cuda(), nu_sc, var.cuda(), return_counts=True)
    co, _ = O.ser_constell_counts(out.clone(), tx, amp, nu_sc, var)
    # the global rescale factor is a float reduction; a symbol within 1 ulp of a threshold may flip
    assert np.abs(c.cpu().numpy().astype(np.int64) - co.numpy()).max() <= 2


@pytest.mark.parametrize("name", ["cma_16qam_M25_N600", "cma_4qam_M7_N300"])
def test_cma_family_against_reference_golden(name):
    import vae_equalizer_b200.shared_funcs as sfun
    g = load(name)
    rx = T(g["rx"]).cuda()
    bl, st = int(g["batchlen"]), int(g["symb_step"])
    runs = {
        "cma": lambda h: sfun.CMA(rx, 1, h, float(g["lr"]), 2, True),
        "batch": lambda h: sfun.CMAbatch(rx, 1, h, float(g["lr_batch"]), bl, 2, True),
        "flex": lambda h: sfun.CMAflex(rx, 1, h, float(g["lr_flex"]), bl, st, 2, True),
        "eval": lambda h: sfun.CMA(rx, 1, h, float(g["lr"]), 2, False),
    }
    for tag, fn in runs.items():
        h = T(g["h0"].copy()).cuda()
        out, h_new, e = fn(h)
        assert h_new is h
        assert np.abs(out.cpu().numpy() - g[tag + "_out"]).max() < 5e-5, tag
        assert np.abs(e.cpu().numpy() - g[tag + "_e"]).max() < 1e-4, tag
        assert rel(h_new, g[tag + "_h"]) < 1e-4, tag
    y = sfun.CPE(T(g["cpe_in"]).cuda())
    assert np.abs(y.cpu().numpy() - g["cpe_out"]).max() < 5e-5


def test_cpe_unwrap_against_reference_golden():
    import vae_equalizer_b200.shared_funcs as sfun
    g = load("cpe_N3000")
    y = sfun.CPE(T(g["y"]).cuda())
    assert np.abs(y.cpu().numpy() - g["out"]).max() < 5e-5


@pytest.mark.parametrize("name", ["awgn_16qam_M25_B350", "awgn_64qam_M9_B200"])
def test_awgn_trajectory_against_reference_golden(name):
    from vae_equalizer_b200.awgn import AWGNEqualizer
    from vae_equalizer_b200.processing import awgn_ser_q
    g = load(name)
    M = g["W0"].shape[-1]
    eq = AWGNEqualizer(M, 2, g["amp"], g["P"], float(g["amp_mean"]), float(g["var"]), W0=T(g["W0"]), h0=T(g["h0"]))
    lr = float(g["lr"])
    for s in range(g["rx"].shape[0]):
        rx = T(g["rx"][s]).cuda()
        q, out, loss, gW, gh = eq.forward_backward(rx)
        assert np.abs(out.cpu().numpy() - g["out"][s]).max() < 5e-6, s
        assert np.abs(q.cpu().numpy() - g["q"][s]).max() < 5e-5, s
        assert rel(loss, g["loss"][s]) < 1e-4, s
        assert rel(gW, g["gW"][s]) < 2e-4 and rel(gh, g["gh"][s]) < 2e-4, s
        eq.train_step(rx, lr, lr)
        assert rel(eq.W, g["W"][s]) < 1e-4 and rel(eq.h, g["h"][s]) < 1e-4, s
    ser, counts = awgn_ser_q(T(g["q"][-1]).cuda()[:, 11:-11].contiguous(), T(g["tx_last"]).cuda()[:, 11:-11].contiguous())
    assert abs(float(ser) - float(g["ser_last"])) < 1e-7


def test_operator_level_autograd_matches_oracle():
    """The reference's call sequence: net(x) -> loss_function_shaping -> backward -> torch Adam."""
    import vae_equalizer_b200.shared_funcs as sfun
    g = load("dp_step_16qam_M9_B64")
    M = 9
    dev = "cuda"
    amp, var, P, nu_sc = T(g["amp"]).to(dev), T(g["var"]).to(dev), T(g["P"]).to(dev), float(g["nu_sc"])
    net = sfun.twoXtwoFIR(M, 2).to(dev)
    with torch.no_grad():
        net.conv_w.weight.copy_(T(g["W0"]))
    h_est = T(g["h0"]).to(dev).requires_grad_(True)
    opt = torch.optim.Adam(net.parameters(), lr=float(g["lr"]))
    opt.add_param_group({"params": h_est})
    for s in range(g["rx"].shape[0]):
        x = T(g["rx"][s]).to(dev)
        opt.zero_grad()
        q, out = net(x, amp, var, nu_sc)
        loss, ve = sfun.loss_function_shaping(q.squeeze(), x.squeeze(), h_est, amp, P)
        loss.backward()
        assert rel(net.conv_w.weight.grad, g["gW"][s]) < 2e-4 and rel(h_est.grad, g["gh"][s]) < 2e-4
        opt.step()
        assert rel(loss, g["loss"][s]) < 1e-4 and rel(ve, g["var_est"][s]) < 1e-4
        assert rel(net.conv_w.weight, g["W"][s]) < 1e-4 and rel(h_est, g["h"][s]) < 1e-4


@pytest.mark.parametrize("mod,M,B", [("64-QAM", 25, 100), ("16-QAM", 9, 1300), ("4-QAM", 5, 64)])
def test_equalizer_autograd_through_a_derived_q(mod, M, B):
    """A q DERIVED from net(x) (here a clamp + mask, and a second loss on `out`) must differentiate into conv_w.weight like the
    reference's nn.Module (sf:500-527): the fused shortcut does not apply, autograd goes through vaeq_dp_loss_from_q (dL/dq) and
    vaeq_eq_backward (softmin + FIR backward for arbitrary upstream gradients).  Checked against the CPU oracle's autograd."""
    import vae_equalizer_b200.shared_funcs as sfun
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", mod, "cpu", 0.0270955, 2, M, 12)
    gen = torch.Generator().manual_seed(3 * B + M)
    rx = 0.6 * torch.randn(2, 2, 2 * B, generator=gen)
    W0 = torch.zeros(2, 4, M)
    W0[0, 0, M // 2] = W0[1, 1, M // 2] = 1.0
    W0 += 0.03 * torch.randn(2, 4, M, generator=gen)
    h0 = h_est.detach() + 0.03 * torch.randn(2, 2, 2, M, generator=gen)
    Pt = torch.tensor(P, dtype=torch.float32)
    mask = (torch.rand(1, 1, B, generator=gen) > 0.2).float()

    def objective(q, out, rx_, h, amp_, P_, lossfn):
        qd = q.clamp(min=1e-6) * mask.to(q.device) + (1.0 - mask.to(q.device)) / q.shape[1] * 2
        loss, ve = lossfn(qd, rx_, h, amp_, P_)
        return loss + 3.0 * (out ** 2).sum(), ve

    Wo, ho = W0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
    qo, oo = O.equalizer_forward(rx, Wo, amp, var, nu_sc, 2)
    lo, vo = objective(qo, oo, rx, ho, amp, Pt, O.elbo_loss)
    lo.backward()
    net = sfun.twoXtwoFIR(M, 2).cuda()
    with torch.no_grad():
        net.conv_w.weight.copy_(W0)
    hd = h0.cuda().requires_grad_(True)
    x = rx.cuda()
    q, out = net(x, amp.cuda(), var.cuda(), nu_sc)
    assert q.grad_fn is not None and out.grad_fn is not None
    ld, vd = objective(q, out, x, hd, amp.cuda(), Pt.cuda(), sfun.loss_function_shaping)
    ld.backward()
    assert rel(ld.detach(), lo.detach()) < 1e-4 and rel(vd, vo) < 1e-4
    assert rel(net.conv_w.weight.grad, Wo.grad) < 2e-4 and rel(hd.grad, ho.grad) < 2e-4
    # and the unmodified q (through the .squeeze() view the reference drivers pass) still takes the fused path with the same result
    net.zero_grad()
    q2, _ = net(x, amp.cuda(), var.cuda(), nu_sc)
    n_launch = sfun._lib.load().vaeq_launch_count(7)
    l2, _ = sfun.loss_function_shaping(q2.squeeze(), x.squeeze(), hd, amp.cuda(), Pt.cuda())
    l2.backward()
    assert sfun._lib.load().vaeq_launch_count(7) == n_launch       # no vaeq_eq_backward launch: fused
    Wf = W0.clone().requires_grad_(True)
    qf, _ = O.equalizer_forward(rx, Wf, amp, var, nu_sc, 2)
    lf, _ = O.elbo_loss(qf, rx, h0, amp, Pt)
    lf.backward()
    assert rel(l2.detach(), lf.detach()) < 1e-4 and rel(net.conv_w.weight.grad, Wf.grad) < 2e-4


def test_calls_follow_the_tensors_device():
    """Every Python entry runs on the device of its tensors, not on the caller's current device (two GPUs needed)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import vae_equalizer_b200.shared_funcs as sfun
    from vae_equalizer_b200.dp import DPEqualizer
    g = load("dp_step_16qam_M9_B64")
    res = []
    for dev in ("cuda:0", "cuda:1"):
        with torch.cuda.device(0):                               # current device stays 0 on purpose
            eq = DPEqualizer(9, 2, T(g["amp"]), T(g["P"]), T(g["var"]), float(g["nu_sc"]), device=dev, W0=T(g["W0"]), h0=T(g["h0"]))
            q, out, loss, ve, gW, gh = eq.forward_backward(T(g["rx"][0]).to(dev))
            q2 = sfun.soft_dec(out, T(g["var"]).to(dev), T(g["amp"]).to(dev), float(g["nu_sc"]))
            res.append((loss.cpu(), gW.cpu(), q2.cpu()))
            assert q.device == torch.device(dev) and q2.device == torch.device(dev)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    with pytest.raises(sfun._lib.VaeqError):
        sfun.soft_dec(out, T(g["var"]).to("cuda:0"), T(g["amp"]).to(dev), float(g["nu_sc"]))


@pytest.mark.parametrize("mod,M,B", [("64-QAM", 25, 100), ("16-QAM", 9, 700), ("4-QAM", 5, 64), ("64-QAM", 13, 1200)])
def test_loss_function_shaping_on_arbitrary_q(mod, M, B):
    """loss_function_shaping(q, rx, h_est, amp, P) as a plain operator on a q that did NOT come from this package's equalizer
    (a softmax of random logits): loss, var_est, dL/dq and dL/dh_est against the CPU oracle's autograd (sf:92-137)."""
    import vae_equalizer_b200.shared_funcs as sfun
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", mod, "cpu", 0.0270955, 2, M, 20)
    n = amp.numel()
    gen = torch.Generator().manual_seed(B + M)
    q0 = torch.softmax(3.0 * torch.randn(2, 2, n, B, generator=gen), dim=2).reshape(2, 2 * n, B)
    rx = 0.7 * torch.randn(2, 2, 2 * B, generator=gen)
    h0 = h_est.detach() + 0.05 * torch.randn(2, 2, 2, M, generator=gen)
    Pt = torch.tensor(P, dtype=torch.float32)
    qo, ho = q0.clone().requires_grad_(True), h0.clone().requires_grad_(True)
    lo, vo = O.elbo_loss(qo, rx, ho, amp, Pt)
    lo.backward()
    qd, hd = q0.cuda().requires_grad_(True), h0.cuda().requires_grad_(True)
    loss, ve = sfun.loss_function_shaping(qd, rx.cuda(), hd, amp.cuda(), Pt.cuda())
    (2.0 * loss).backward()                                  # also checks that the upstream gradient is applied
    assert rel(loss, lo.detach()) < 1e-5 and rel(ve, vo) < 1e-5
    assert rel(qd.grad / 2.0, qo.grad) < 1e-4 and rel(hd.grad / 2.0, ho.grad) < 1e-4
    assert not ve.requires_grad


def test_dropin_processing_runs_and_converges():
    """func_VAELE_DP_MQAM_shaping.processing with the Eval_run_DP defaults (shortened): the SER must fall the
    way SURVEY.md §8c reports for the reference (converged SER of a few 1e-2 at SNR 23 dB)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "vae_equalizer_b200", "dropin"))
    import func_VAELE_DP_MQAM_shaping as process
    rng = np.random.default_rng(7)
    phiIQ = np.array([0.0314, 0.0314], dtype=np.complex64)
    SER, Var_est, var = process.processing("64-QAM", 2, 23, 0, 25, 0.0, np.pi / 10, 2.5e-3, 100, 10000, 24, 10, "h0", 90e9,
                                           -26e-24, 0.1e-12 * np.sqrt(1000), phiIQ, 170, rng=rng, verbose=False)
    SER = SER.cpu().numpy()
    assert SER.shape == (4, 24) and Var_est.shape == (2, 24)
    assert SER[:, 0].min() > 0.5                      # untrained
    assert SER[:, -4:].max() < 0.08                   # converged
    assert np.isfinite(Var_est.cpu().numpy()).all()


def test_dropin_cma_variants_run():
    from vae_equalizer_b200 import processing as pr
    rng = np.random.default_rng(9)
    phiIQ = np.array([0.0314, 0.0314], dtype=np.complex64)
    for fn, lr in ((pr.processing_cma_dp, 1e-3), (pr.processing_cmabatch_dp, 1e-5), (pr.processing_cmaflex_dp, 1e-6)):
        SER, Var_est, var = fn("16-QAM", 2, 23, 0, 25, 0.0, np.pi / 10, lr, 100, 3000, 3, 10, "h0", 90e9, -26e-24,
                               0.1e-12 * np.sqrt(1000), phiIQ, 170, rng=rng, verbose=False)
        assert SER.shape == (4, 3) and torch.isfinite(SER).all()


def test_dropin_awgn_runs():
    from vae_equalizer_b200 import processing as pr
    SER = pr.processing_vaele_awgn("16-QAM", 2, 20, 0, 25, 5e-3, 350, 5000, 1200, 40, 10, "h1", rng=np.random.default_rng(1), verbose=False)
    assert SER.shape == (4,) and torch.isfinite(SER).all() and float(SER[-1]) < float(SER[0]) + 1e-6


@pytest.mark.parametrize("N,n_shift", [(64, 21), (516, 21), (1003, 21), (20000, 41), (4100, 64), (300000, 21), (516, 5), (1003, 24), (2000, 1), (3000, 25)])
def test_find_shift_one_pass_kernel_against_oracle(N, n_shift):
    """csrc/shift_corr.cuh: tiles, circular wrap at both ends, vector and scalar staging, one and two shift passes."""
    import vae_equalizer_b200.shared_funcs as sfun
    gen = torch.Generator().manual_seed(N + n_shift)
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", "64-QAM", "cpu", 0.0270955, 2, 25, 18)
    idx = torch.randint(0, 8, (2, 2, N), generator=gen)
    tx = amp[idx].to(torch.float16)
    true_shift = (3, -2)
    out = amp[idx] + 0.08 * torch.randn(2, 2, N, generator=gen)
    out = torch.stack((out[1].roll(true_shift[0], -1), out[0].roll(true_shift[1], -1)))      # crossed polarisations, shifted
    q = O.soft_demap(out, var, amp, nu_sc)
    so, ro, co = O.find_shift(q, tx, n_shift, amp, 2, return_corr=True)
    s, r, c_q = sfun.find_shift(q.cuda(), tx.cuda(), n_shift, amp.cuda(), 2, return_corr=True)
    assert s.tolist() == so.tolist() and r == ro and rel(c_q, co) < 2e-5
    sq, rq = so.tolist(), ro
    so, ro, co = O.find_shift_symb_full(out, tx, n_shift, return_corr=True)
    s, r, c = sfun.find_shift_symb_full(out.cuda(), tx.cuda(), n_shift, return_corr=True)
    assert s.tolist() == so.tolist() and r == ro and rel(c, co) < 2e-5
    # a view with a misaligned base takes the scalar staging path: same decisions, same correlations to rounding
    big_q, big_o, big_t = torch.zeros(2, 16, N + 5).cuda(), torch.zeros(2, 2, N + 5).cuda(), torch.zeros(2, 2, N + 5, dtype=torch.float16).cuda()
    big_q[:, :, 1:N + 1], big_o[:, :, 1:N + 1], big_t[:, :, 1:N + 1] = q.cuda(), out.cuda(), tx.cuda()
    s2, r2, c2 = sfun.find_shift(big_q[:, :, 1:N + 1], big_t[:, :, 1:N + 1], n_shift, amp.cuda(), 2, return_corr=True)
    assert s2.tolist() == sq and r2 == rq and rel(c2, c_q) < 1e-6
    s3, r3 = sfun.find_shift_symb_full(big_o[:, :, 1:N + 1], big_t[:, :, 1:N + 1], n_shift)
    assert s3.tolist() == so.tolist() and r3 == ro


def test_vector_and_scalar_scan_paths_agree():
    """4-symbols-per-thread kernels (aligned rows) against the one-symbol kernels (misaligned views of the same data)."""
    import vae_equalizer_b200.shared_funcs as sfun
    gen = torch.Generator().manual_seed(11)
    N = 200000
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", "64-QAM", "cpu", 0.0270955, 2, 25, 17)
    idx = torch.randint(0, 8, (2, 2, N), generator=gen)
    tx = amp[idx].to(torch.float16).cuda()
    out = ((amp[idx] + 0.06 * torch.randn(2, 2, N, generator=gen)) * 0.93).cuda()
    ampc, varc = amp.cuda(), var.cuda()
    big_o, big_t = torch.zeros(2, 2, N + 8, device="cuda"), torch.zeros(2, 2, N + 8, dtype=torch.float16, device="cuda")
    big_o[:, :, 3:N + 3], big_t[:, :, 3:N + 3] = out, tx
    vo, vt = big_o[:, :, 3:N + 3], big_t[:, :, 3:N + 3]
    q_vec = sfun.soft_dec(out, varc, ampc, nu_sc)                       # vector kernel
    big_q = torch.zeros(2, 16, N + 8, device="cuda")
    q_sca = big_q[:, :, 3:N + 3]
    q_sca.copy_(q_vec)                                                  # same values behind a misaligned view
    q_prec = sfun.soft_dec(out[:, :, :N - 1].contiguous(), varc, ampc, nu_sc)     # N - 1 is odd: the precise one-symbol kernel
    assert float((q_prec - q_vec[:, :, :N - 1]).abs().max()) < 1e-6
    assert torch.equal(q_prec.reshape(2, 2, 8, -1).argmax(2), q_vec[:, :, :N - 1].reshape(2, 2, 8, -1).argmax(2))
    qo = O.soft_demap(out.cpu(), var, amp, nu_sc)
    assert float((qo - q_vec.cpu()).abs().max()) < 2e-5
    _, c_vec = sfun.SER_IQflip(q_vec, tx, return_counts=True)
    _, c_sca = sfun.SER_IQflip(q_sca, vt, return_counts=True)
    assert torch.equal(c_vec, c_sca)
    g_vec, g_sca = sfun.GMI(q_vec, tx, P), sfun.GMI(q_sca, vt, P)
    assert float((g_vec - g_sca).abs().max()) < 1e-6
    r1, r2 = out.clone(), vo                                              # r2 rescaled in place inside big_o
    s1, c1 = sfun.SER_constell_shaping(r1, tx, ampc, nu_sc, varc, return_counts=True)
    s2, c2 = sfun.SER_constell_shaping(r2, vt, ampc, nu_sc, varc, return_counts=True)
    assert torch.equal(c1, c2) and torch.equal(s1, s2) and torch.equal(r1, big_o[:, :, 3:N + 3])


def test_full_size_evaluation_properties():
    """Evaluation kernels at the bench size (N = 2^22, where no CPU oracle finishes in seconds): size-independent properties.
    * the shift search finds a planted circular shift / polarisation swap, and rolling the input moves the answer with it;
    * SER_IQflip is invariant under the rotations and the IQ flip it searches over (its counts permute), and counts the planted errors;
    * soft_dec posteriors sum to one and decide like a nearest-threshold slicer away from the thresholds;
    * SER_constell_shaping agrees with SER_IQflip on the same decisions when there is no shaping (nu = 0)."""
    import vae_equalizer_b200.shared_funcs as sfun
    N, n = 1 << 22, 8
    gen = torch.Generator(device="cuda").manual_seed(5)
    h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = O.init("h0", "64-QAM", "cpu", 0.0, 2, 25, 25)
    amp, var = amp.cuda(), var.cuda()
    idx = torch.randint(0, n, (2, 2, N), device="cuda", generator=gen)
    tx = amp[idx].to(torch.float16)
    clean = amp[idx]
    out = (clean + 0.03 * torch.randn(2, 2, N, device="cuda", generator=gen)).contiguous()
    # planted errors: every 1000th symbol of pol 0 gets its I level replaced by the neighbouring one
    bad = torch.arange(0, N, 1000, device="cuda")
    out[0, 0, bad] = amp[(idx[0, 0, bad] + 1) % n]
    q = sfun.soft_dec(out, var, amp, float(nu_sc))
    assert float((q.reshape(2, 2, n, N).sum(2) - 1).abs().max()) < 1e-5
    dec = q.reshape(2, 2, n, N).argmax(2)
    slicer = (out.unsqueeze(2) - amp.view(1, 1, n, 1)).abs().argmin(2)
    assert torch.equal(dec, slicer)
    ser, counts = sfun.SER_IQflip(q, tx, return_counts=True)
    n_err0 = int(((dec[0] != idx[0]).any(0)).sum())
    n_err1 = int(((dec[1] != idx[1]).any(0)).sum())
    assert counts[0, 0, 0].item() == n_err0 and counts[0, 1, 0].item() == n_err1 and n_err0 >= bad.numel() - 8
    assert ser.tolist() == [n_err0 / N, n_err1 / N] or np.allclose(ser.cpu().numpy(), [n_err0 / N, n_err1 / N], rtol=1e-6)
    # a 90 degree rotation of the constellation (I, Q) -> (-Q, I) is one of the searched hypotheses: same minimum, permuted counts
    q4 = q.reshape(2, 2, n, N)
    q_rot = torch.stack((q4[:, 1].flip(1), q4[:, 0]), dim=1).reshape(2, 2 * n, N).contiguous()
    ser_r, counts_r = sfun.SER_IQflip(q_rot, tx, return_counts=True)
    assert torch.equal(ser_r, ser) and sorted(counts_r[0, 0].tolist()) == sorted(counts[0, 0].tolist())
    # constellation SER without shaping = the same decisions
    ser_c, counts_c = sfun.SER_constell_shaping(out.clone(), tx, amp, float(nu_sc), var, return_counts=True)
    assert abs(counts_c[0, 0, 0].item() - n_err0) <= 16 and abs(counts_c[0, 1, 0].item() - n_err1) <= 16   # global rescale: threshold-edge symbols may flip
    # shift search: plant (shift, swap), then roll by 3 more
    out_s = torch.stack((out[1].roll(4, -1), out[0].roll(-2, -1))).contiguous()
    s, r = sfun.find_shift_symb_full(out_s, tx, 21)
    assert r == 1 and s.tolist() == [4, -2]            # as the CPU oracle answers on a 20 000-symbol prefix of the same construction
    s2, r2 = sfun.find_shift_symb_full(out_s.roll(3, -1), tx, 21)
    assert r2 == 1 and s2.tolist() == [7, 1]
    qs = sfun.soft_dec(out_s, var, amp, float(nu_sc))
    s3, r3 = sfun.find_shift(qs, tx, 21, amp, 2)
    assert r3 == r and s3.tolist() == s.tolist()
