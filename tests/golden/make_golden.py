"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference has no tests/fixtures of its own, so these outputs of the reference
itself are what pins the oracle (oracle/vaeq_oracle.py) and, through it, the CUDA path.
All inputs are deterministic (torch.Generator seeds below); every array is float32 unless noted.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _load_reference as ref_loader  # noqa: E402

sf = ref_loader.load("shared_funcs")
awgn = ref_loader.load("func_VAELE_MQAM_shaping", subdir="AWGN_channel")


def npy(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def pcs_symbols(gen, amps, P, shape):
    idx = torch.multinomial(torch.tensor(P, dtype=torch.float64), int(np.prod(shape)), replacement=True, generator=gen)
    return torch.tensor(amps, dtype=torch.float64)[idx].reshape(shape)


def synth_rx(gen, amps, P, B, sps, snr_db, theta=0.3):
    """A small deterministic DP 'channel': PCS symbols, 3-tap ISI, rotation theta, AWGN.  Returns rx, tx."""
    sym = pcs_symbols(gen, amps, P, (2, 2, B + 8)).to(torch.float32)
    s = torch.complex(sym[:, 0], sym[:, 1])
    up = torch.zeros(2, sps * (B + 8), dtype=torch.complex64)
    up[:, ::sps] = s
    g = torch.tensor([0.05, 0.25, 0.9, 0.25, 0.05], dtype=torch.complex64)
    sh = torch.stack([torch.from_numpy(np.convolve(up[i].numpy(), g.numpy(), mode="same")) for i in range(2)])
    c, sn = np.cos(theta), np.sin(theta)
    mix = torch.stack((c * sh[0] + sn * sh[1], -sn * sh[0] + c * sh[1]))
    sig = 10 ** (-snr_db / 20) * 0.7
    noise = sig * torch.complex(torch.randn(mix.shape, generator=gen), torch.randn(mix.shape, generator=gen))
    r = (mix + noise)[:, sps * 4: sps * (B + 4)]
    rx = torch.stack((r.real, r.imag), dim=1).to(torch.float32).contiguous()
    tx = sym[:, :, 4:B + 4].to(torch.float16).contiguous()
    return rx, tx


def case_dp_step(name, mod, M, B, nu, snr, seed, steps=3, lr=2.5e-3, perturb=0.05):
    gen = torch.Generator().manual_seed(seed)
    h_est, h_channel, P, amp, amps, pol, nu_sc, var, pow_mean = sf.init("h0", mod, "cpu", nu, 2, M, snr)
    rx_all = [synth_rx(gen, amps, P, B, 2, snr)[0] for _ in range(steps)]
    net = sf.twoXtwoFIR(M, 2)
    with torch.no_grad():
        net.conv_w.weight += perturb * torch.randn(2, 4, M, generator=gen)
        h_est += perturb * torch.randn(2, 2, 2, M, generator=gen)
    W0, h0 = npy(net.conv_w.weight).copy(), npy(h_est).copy()
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    opt.add_param_group({"params": h_est})
    Pt = torch.tensor(P, dtype=torch.float32)
    rec = dict(W0=W0, h0=h0, amp=npy(amp), P=npy(Pt), var=npy(var), nu_sc=np.float64(nu_sc), lr=np.float64(lr),
               sps=np.int64(2), rx=np.stack([npy(r) for r in rx_all]))
    qs, outs, losses, ves, gWs, ghs, Ws, hs = [], [], [], [], [], [], [], []
    for s in range(steps):
        opt.zero_grad()
        q, out = net(rx_all[s], amp, var, nu_sc)
        loss, ve = sf.loss_function_shaping(q, rx_all[s], h_est, amp, Pt)
        loss.backward()
        gWs.append(npy(net.conv_w.weight.grad).copy())
        ghs.append(npy(h_est.grad).copy())
        opt.step()
        qs.append(npy(q).copy()); outs.append(npy(out).copy()); losses.append(loss.item()); ves.append(npy(ve).copy())
        Ws.append(npy(net.conv_w.weight).copy()); hs.append(npy(h_est).copy())
    rec.update(q=np.stack(qs), out=np.stack(outs), loss=np.asarray(losses, np.float32), var_est=np.stack(ves),
               gW=np.stack(gWs), gh=np.stack(ghs), W=np.stack(Ws), h=np.stack(hs))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "loss", losses)


def case_eval(name, mod, N, nu, snr, seed, shift=(2, -1), swap=False):
    """find_shift / find_shift_symb_full / SER_IQflip / SER_constell_shaping / soft_dec on one synthetic frame."""
    gen = torch.Generator().manual_seed(seed)
    h_est, h_channel, P, amp, amps, pol, nu_sc, var, pow_mean = sf.init("h0", mod, "cpu", nu, 2, 9, snr)
    tx64 = pcs_symbols(gen, amps, P, (2, 2, N))
    tx = tx64.to(torch.float16)
    out = tx64.to(torch.float32) + np.sqrt(var[0].item()) * 1.4 * torch.randn(2, 2, N, generator=gen)
    out = out * 0.83                                            # a gain the constellation SER must undo
    out[0] = torch.roll(out[0], shift[0], -1)
    out[1] = torch.roll(out[1], shift[1], -1)
    if swap:
        out = out.roll(1, 0)
    out = out.contiguous()
    q = sf.soft_dec(out, var, amp, nu_sc)
    s1, r1 = sf.find_shift(q, tx, 21, amp, 2)
    s2, r2 = sf.find_shift_symb_full(out, tx, 21)
    # aligned copies, cut like func_VAEflex_DP_MQAM_shaping.py:74-84
    qa = q.roll(r1, 0)
    qa[0], qa[1] = qa[0].roll(int(-s1[0]), -1).clone(), qa[1].roll(int(-s1[1]), -1).clone()
    cut1 = int(torch.max(torch.abs(s1)))
    q_cut, tx_cut1 = qa[:, :, 11:-11 - cut1].contiguous(), tx[:, :, 11:-11 - cut1].contiguous()
    ser_q = sf.SER_IQflip(q_cut, tx_cut1)
    oa = out.roll(r2, 0)
    oa[0], oa[1] = oa[0].roll(int(-s2[0]), -1).clone(), oa[1].roll(int(-s2[1]), -1).clone()
    cut2 = int(torch.max(torch.abs(s2)))
    o_cut, tx_cut2 = oa[:, :, 11:-11 - cut2].contiguous(), tx[:, :, 11:-11 - cut2].contiguous()
    o_in = o_cut.clone()
    ser_c = sf.SER_constell_shaping(o_cut, tx_cut2, amp, nu_sc, var)      # mutates o_cut (sf:242)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), amp=npy(amp), var=npy(var), nu_sc=np.float64(nu_sc),
                        P=np.asarray(P, np.float32), tx=npy(tx), out=npy(out), q=npy(q),
                        shift_q=npy(s1), r_q=np.int64(r1), shift_c=npy(s2), r_c=np.int64(r2),
                        q_cut=npy(q_cut), tx_cut_q=npy(tx_cut1), ser_q=npy(ser_q),
                        o_cut_in=npy(o_in), o_cut_scaled=npy(o_cut), tx_cut_c=npy(tx_cut2), ser_c=npy(ser_c))
    print(name, "shift", s1.tolist(), r1, s2.tolist(), r2, "SER", ser_q.tolist(), ser_c.tolist())


def case_cma(name, mod, M, N, lr, batchlen, step, seed):
    gen = torch.Generator().manual_seed(seed)
    h_est, h_channel, P, amp, amps, pol, nu_sc, var, pow_mean = sf.init("h0", mod, "cpu", 0.0, 2, M, 20)
    rx, tx = synth_rx(gen, amps, P, N, 2, 20)
    lr_b, lr_f = lr / 10, lr / 100          # the batched variants sum batchlen increments per update
    rec = dict(rx=npy(rx), lr=np.float64(lr), lr_batch=np.float64(lr_b), lr_flex=np.float64(lr_f),
               batchlen=np.int64(batchlen), symb_step=np.int64(step), M=np.int64(M))
    with torch.no_grad():
        for tag, fn in (("cma", lambda h: sf.CMA(rx, 1, h, lr, 2, True)),
                        ("batch", lambda h: sf.CMAbatch(rx, 1, h, lr_b, batchlen, 2, True)),
                        ("flex", lambda h: sf.CMAflex(rx, 1, h, lr_f, batchlen, step, 2, True)),
                        ("eval", lambda h: sf.CMA(rx, 1, h, lr, 2, False))):
            h = h_est.detach().clone()
            h[0, 1, 1, M // 2] = 0.1
            h[1, 0, 0, M // 2 - 1] = -0.07
            rec["h0"] = npy(h).copy()
            out, h_new, e = fn(h)
            rec[tag + "_out"], rec[tag + "_h"], rec[tag + "_e"] = npy(out), npy(h_new).copy(), npy(e)
        y = sf.CPE(torch.from_numpy(rec["cma_out"])[:, :, 10:-10])
        rec["cpe_in"], rec["cpe_out"] = rec["cma_out"][:, :, 10:-10].copy(), npy(y)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "e sums", rec["cma_e"].sum(), rec["batch_e"].sum(), rec["flex_e"].sum())


def case_cpe(name, N, seed):
    """CPE on a 16-QAM-like signal with a phase ramp that forces unwrap jumps (sf:140-186)."""
    gen = torch.Generator().manual_seed(seed)
    lev = torch.tensor([-3., -1., 1., 3.]) / np.sqrt(10)
    s = lev[torch.randint(0, 4, (2, 2, N), generator=gen)]
    t = torch.arange(N, dtype=torch.float32)
    ph = torch.stack((2.5e-3 * t + 0.2, -1.8e-3 * t - 0.4))
    c, sn = torch.cos(ph), torch.sin(ph)
    y = torch.stack((s[:, 0] * c - s[:, 1] * sn, s[:, 1] * c + s[:, 0] * sn), dim=1)
    y = (y + 0.03 * torch.randn(y.shape, generator=gen)).contiguous()
    out = sf.CPE(y)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), y=npy(y), out=npy(out))
    print(name, "done")


def case_awgn(name, mod, M, B, nu, snr, seed, steps=3, lr=5e-3):
    gen = torch.Generator().manual_seed(seed)
    from oracle.vaeq_oracle import awgn_constants
    amps, P, amp_mean, var = awgn_constants(mod, nu, snr)
    amp = torch.tensor(amps, dtype=torch.float32)
    Pt = torch.tensor(P, dtype=torch.float32)
    net = awgn.twoFIR(M, 2)
    h = torch.zeros(2, M)
    h[0, M // 2] = 1
    with torch.no_grad():
        net.conv_w.weight += 0.05 * torch.randn(1, 2, M, generator=gen)
        h += 0.05 * torch.randn(2, M, generator=gen)
    h.requires_grad_(True)
    W0, h0 = npy(net.conv_w.weight).copy(), npy(h).copy()
    opt = torch.optim.Adam(net.parameters(), lr=lr, amsgrad=True)
    opt.add_param_group({"params": h})
    rxs, qs, outs, losses, gWs, ghs, Ws, hs = [], [], [], [], [], [], [], []
    for s in range(steps):
        rx2, tx2 = synth_rx(gen, amps, P, B, 2, snr, theta=0.0)
        x = rx2[0].contiguous()
        opt.zero_grad()
        q, out = net(x, amp, amp_mean, var)
        loss = awgn.loss_function(q, x, h, "cpu", amp, Pt)
        loss.backward()
        gWs.append(npy(net.conv_w.weight.grad).copy()); ghs.append(npy(h.grad).copy())
        opt.step()
        rxs.append(npy(x)); qs.append(npy(q).copy()); outs.append(npy(out).copy()); losses.append(loss.item())
        Ws.append(npy(net.conv_w.weight).copy()); hs.append(npy(h).copy())
    tx = tx2[0]
    ser = awgn.SER_q(q.detach()[:, 11:-11], tx[:, 11:-11], 2, len(amps), "cpu")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), W0=W0, h0=h0, amp=npy(amp), P=npy(Pt), amp_mean=np.float64(amp_mean),
                        var=np.float64(var), lr=np.float64(lr), rx=np.stack(rxs), q=np.stack(qs), out=np.stack(outs),
                        loss=np.asarray(losses, np.float32), gW=np.stack(gWs), gh=np.stack(ghs), W=np.stack(Ws), h=np.stack(hs),
                        tx_last=npy(tx), ser_last=npy(ser))
    print(name, "loss", losses, "ser", ser.item())


def case_kat(name):
    """The survey's known-answer setup, regenerated."""
    h_est, h_channel, P, amp, amps, pol, nu_sc, var, pow_mean = sf.init("h0", "64-QAM", "cpu", 0.0270955, 2, 25, 23)
    t = torch.arange(200, dtype=torch.float64)
    rx = torch.stack([torch.stack([0.7 * torch.sin(0.37 * t + 1.3 * p + 0.7 * c) + 0.2 * torch.cos(0.11 * t * (1 + p) + c)
                                   for c in range(2)]) for p in range(2)]).float()
    k = torch.arange(25, dtype=torch.float64)
    net = sf.twoXtwoFIR(25, 2)
    with torch.no_grad():
        for o in range(2):
            for c in range(4):
                net.conv_w.weight[o, c] += (0.02 * torch.cos(0.9 * k + o + 2 * c)).float()
        h_est += (0.03 * torch.sin(0.5 * k + torch.arange(8, dtype=torch.float64).reshape(2, 2, 2, 1))).float()
    Pt = torch.tensor(P, dtype=torch.float32)
    q, out = net(rx, amp, var, nu_sc)
    loss, ve = sf.loss_function_shaping(q, rx, h_est, amp, Pt)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), rx=npy(rx), W0=npy(net.conv_w.weight), h0=npy(h_est), amp=npy(amp),
                        P=npy(Pt), var=npy(var), nu_sc=np.float64(nu_sc), q=npy(q), out=npy(out), loss=np.float32(loss.item()),
                        var_est=npy(ve), gW=npy(net.conv_w.weight.grad), gh=npy(h_est.grad))
    print(name, "loss", loss.item(), "var_est", ve.tolist())


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    torch.set_num_threads(1)          # summation order independent of the host's core count
    case_kat("kat_64qam_M25_B100")
    case_dp_step("dp_step_64qam_M25_B100", "64-QAM", 25, 100, 0.0270955, 23, seed=11)
    case_dp_step("dp_step_16qam_M9_B64", "16-QAM", 9, 64, 0.0, 18, seed=12)
    case_dp_step("dp_step_4qam_M5_B48", "4-QAM", 5, 48, 0.1, 12, seed=13)
    case_dp_step("dp_step_64qam_M25_B1000", "64-QAM", 25, 1000, 0.0, 23, seed=14, steps=2)
    case_eval("eval_64qam_N2000", "64-QAM", 2000, 0.0270955, 23, seed=21, shift=(2, -1))
    case_eval("eval_16qam_N1500_swap", "16-QAM", 1500, 0.0, 17, seed=22, shift=(3, 3), swap=True)
    case_cma("cma_16qam_M25_N600", "16-QAM", 25, 600, 1e-3, 100, 10, seed=31)
    case_cma("cma_4qam_M7_N300", "4-QAM", 7, 300, 2e-3, 50, 25, seed=32)
    case_cpe("cpe_N3000", 3000, seed=41)
    case_awgn("awgn_16qam_M25_B350", "16-QAM", 25, 350, 0.0, 20, seed=51)
    case_awgn("awgn_64qam_M9_B200", "64-QAM", 9, 200, 0.0270955, 24, seed=52)
