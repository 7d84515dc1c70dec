"""Driver-level golden vectors: the UNMODIFIED reference `processing()` functions run for a few short frames.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_drivers.py
What is pinned (SURVEY.md §8 row a16 and the §8d trajectory gate):
  * func_VAELE_DP / func_VAEflex_DP / func_CMA_DP / func_CMAbatch_DP / func_CMAflex_DP `processing()`: per frame the received frame and
    the transmitted symbols the driver was given, what its training loop left in out_train / out_const (or what CMA / CPE returned), the
    two (shift, r) results, the tensors it handed to the SER functions and the SER_valid / Var_est columns it returned.
  * a 100-step VAE-LE trajectory (one 10 000-symbol frame at batch_len 100, the Eval_run_DP.py defaults): loss and var_est of every
    step, taps after 10 / 50 / 100 steps.
The reference's generator is unseeded and breaks on numpy >= 1.24 (sf:49), so `sfun.generate_data_shaping` is replaced by a function
that hands out frames made by the oracle's seeded restatement of it; the frames are part of the fixture, so the tests feed the
identical tensors to the CUDA drivers.  Some frames carry a deliberate time shift / polarisation swap of the transmitted symbols so
that the alignment stage (roll, per-minibatch cut, slice) is exercised with non-zero shifts.
The reference's own functions are recorded by wrapping the attributes of its `shared_funcs` module (no edits to /root/reference).
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _load_reference as ref_loader  # noqa: E402
from oracle import vaeq_oracle as O  # noqa: E402

sf = ref_loader.load("shared_funcs")
PHI_IQ = np.array([0.0314, 0.0314], dtype=np.complex64)
CHAN = dict(symb_rate=90e9, tau_cd=-26e-24, tau_pmd=0.1e-12 * np.sqrt(1000))


def npy(t):
    return t.detach().cpu().numpy().copy() if torch.is_tensor(t) else np.asarray(t)


def make_frames(mod, nu, SNR, N, num_frames, theta, theta_diff, seed, shifts):
    """Seeded frames (oracle generator) with the transmitted symbols rolled by shifts[f] = (roll_x, roll_y, swap)."""
    _, h_channel, P, _, amps, pol, _, _, _ = sf.init("h0", mod, "cpu", nu, 2, 25, SNR)
    rng = np.random.default_rng(seed)
    frames = []
    for f in range(num_frames):
        rx, tx, _ = O.generate_data_shaping(N, amps, SNR, h_channel, P, pol, CHAN["symb_rate"], 2, CHAN["tau_cd"], CHAN["tau_pmd"], PHI_IQ,
                                            theta + f * theta_diff, "cpu", rng=rng)
        sx, sy, swap = shifts[f % len(shifts)]
        tx = torch.stack((tx[0].roll(sx, -1), tx[1].roll(sy, -1)))
        if swap:
            tx = tx.roll(1, 0)
        frames.append((rx.contiguous(), tx.contiguous()))
    return frames


class Recorder:
    """Wraps attributes of the reference's shared_funcs module; restores them on exit."""

    def __init__(self, frames):
        self.frames, self.k, self.log, self.saved = frames, 0, [], {}

    def __enter__(self):
        def wrap(name, fn):
            orig = getattr(sf, name)
            self.saved[name] = orig
            setattr(sf, name, lambda *a, **k: fn(orig, *a, **k))

        def gen(orig, N, *a, **k):
            rx, tx = self.frames[self.k]
            assert tx.shape[-1] == N, (tx.shape, N)
            self.k += 1
            self.log.append(dict(rx=npy(rx), tx=npy(tx)))
            return rx.clone(), tx.clone(), 0.0

        def rec_inout(key, clone_in=(0,)):
            def f(orig, *a, **k):
                ins = [npy(a[i]) for i in clone_in]
                res = orig(*a, **k)
                snap = tuple(r.detach().clone() if torch.is_tensor(r) else r for r in res) if isinstance(res, tuple) else res.detach().clone()
                self.log[-1].setdefault(key, []).append((ins, snap))        # a snapshot: CMA* return the h they mutate in place
                return res
            return f

        wrap("generate_data_shaping", gen)
        wrap("find_shift", rec_inout("find_shift", (0, 1)))
        wrap("find_shift_symb_full", rec_inout("find_shift_symb_full", (0, 1)))
        wrap("SER_IQflip", rec_inout("SER_IQflip", (0, 1)))
        wrap("SER_constell_shaping", rec_inout("SER_constell_shaping", (0, 1)))
        wrap("CPE", rec_inout("CPE", (0,)))
        wrap("soft_dec", rec_inout("soft_dec", (0,)))
        for n in ("CMA", "CMAbatch", "CMAflex"):
            wrap(n, rec_inout("cma", (0, 2)))
        return self

    def __exit__(self, *exc):
        for n, f in self.saved.items():
            setattr(sf, n, f)


def run_driver(name, module, mod, nu, SNR, M, lr, batch_len, N_max, num_frames, flex_step, theta, theta_diff, seed, shifts, N_lrhalf=2):
    proc = ref_loader.load(module)
    if "VAEflex" in module or "VAELE" in module:
        N = (N_max // batch_len) * batch_len
    else:
        N = N_max
    frames = make_frames(mod, nu, SNR, N, num_frames, theta, theta_diff, seed, shifts)
    torch.manual_seed(0)
    with Recorder(frames) as rec, contextlib.redirect_stdout(io.StringIO()) as so:
        SER, Var_est, var = proc.processing(mod, 2, SNR, nu, M, theta_diff, theta, lr, batch_len, N_max, num_frames, flex_step, "h0",
                                            CHAN["symb_rate"], CHAN["tau_cd"], CHAN["tau_pmd"], PHI_IQ, N_lrhalf)
    out = dict(SER=npy(SER), Var_est=npy(Var_est), var=npy(var), args=np.array([SNR, nu, M, lr, batch_len, N_max, num_frames, flex_step,
                                                                                  theta, theta_diff, N_lrhalf], dtype=np.float64),
               mod=np.array(mod), stdout=np.array(so.getvalue()))
    for f, L in enumerate(rec.log):
        out[f"f{f}_rx"], out[f"f{f}_tx"] = L["rx"], L["tx"]
        (ins, (sh, r)), = L["find_shift"]
        out[f"f{f}_fs_q"], out[f"f{f}_fs_tx"], out[f"f{f}_fs_shift"], out[f"f{f}_fs_r"] = ins[0], ins[1], npy(sh), np.int64(r)
        (ins, (sh, r)), = L["find_shift_symb_full"]
        out[f"f{f}_fss_out"], out[f"f{f}_fss_tx"], out[f"f{f}_fss_shift"], out[f"f{f}_fss_r"] = ins[0], ins[1], npy(sh), np.int64(r)
        (ins, res), = L["SER_IQflip"]
        out[f"f{f}_iq_n"], out[f"f{f}_iq_ser"] = np.int64(ins[0].shape[-1]), npy(res)
        (ins, res), = L["SER_constell_shaping"]
        out[f"f{f}_cs_n"], out[f"f{f}_cs_ser"] = np.int64(ins[0].shape[-1]), npy(res)
        if "cma" in L:
            (ins, (o, h, e)), = L["cma"]
            out[f"f{f}_cma_h_in"], out[f"f{f}_cma_out"], out[f"f{f}_cma_h"], out[f"f{f}_cma_esum"] = ins[1], npy(o), npy(h), np.float32(torch.sum(e).item())
            (ins, res), = L["CPE"]
            out[f"f{f}_cpe_out"] = npy(res)
            (ins, res), = L["soft_dec"]
            out[f"f{f}_sd_in"] = ins[0]                 # out_const as soft_dec saw it: aligned, evaluated slice rescaled in place (CMA_DP:44,48)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "SER", np.round(out["SER"], 4).tolist(), "shifts", [(out[f"f{f}_fs_shift"].tolist(), int(out[f"f{f}_fs_r"]),
                                                                     out[f"f{f}_fss_shift"].tolist(), int(out[f"f{f}_fss_r"])) for f in range(num_frames)])


SNAPS = (10, 25, 50, 75, 100)


def trajectory(name, mod="64-QAM", nu=0.0270955, SNR=23, M=25, B=100, steps=100, lr=2.5e-3, seed=77):
    """One VAE-LE frame of the Eval_run_DP.py defaults stepped by the reference's own modules (sf.twoXtwoFIR, sf.loss_function_shaping,
    torch.optim.Adam with the two parameter groups of VAELE_DP:28-31).  Recorded: loss / var_est of every step, the taps AND the Adam
    moments after every step (teacher-forced step-by-step checks), and the taps of two more runs of the SAME reference code at a few
    steps: one with 8 host threads instead of 1, one whose input samples were moved by one float32 ulp at random.  The trajectory
    is chaotic (Adam's m / sqrt(v) normalisation amplifies the rounding noise of near-zero gradients), so the reference's own spread
    under such last-bit perturbations is the yardstick for a free-running comparison."""
    rx = None

    def run(threads, ulp_noise=False):
        nonlocal rx
        torch.set_num_threads(threads)
        h_est, h_channel, P, amp, amps, pol, nu_sc, var, pow_mean = sf.init("h0", mod, "cpu", nu, 2, M, SNR)
        if rx is None:
            rng = np.random.default_rng(seed)
            rx = O.generate_data_shaping(B * steps, amps, SNR, h_channel, P, pol, CHAN["symb_rate"], 2, CHAN["tau_cd"], CHAN["tau_pmd"], PHI_IQ,
                                         np.pi / 10, "cpu", rng=rng)[0]
        x = rx
        if ulp_noise:                                            # every second sample (at random) moved to the next float32
            xn = npy(rx)
            x = torch.from_numpy(np.where(np.random.default_rng(0).random(xn.shape) < 0.5, np.nextafter(xn, np.float32(np.inf)), xn).astype(np.float32))
        net = sf.twoXtwoFIR(M, 2)
        opt = torch.optim.Adam(net.parameters(), lr=lr)
        opt.add_param_group({"params": h_est})
        Pt = torch.tensor(P, dtype=torch.float32)
        rec = dict(loss=[], ve=[], W=[], h=[], mW=[], vW=[], mh=[], vh=[])
        for m in range(steps):
            mb = x[:, :, m * B * 2:(m + 1) * B * 2].contiguous()
            opt.zero_grad()
            q, out = net(mb, amp, var, nu_sc)
            loss, ve = sf.loss_function_shaping(q.squeeze(), mb.squeeze(), h_est, amp, Pt)
            loss.backward()
            opt.step()
            sW, sh = opt.state[net.conv_w.weight], opt.state[h_est]
            for k, v in (("loss", loss.item()), ("ve", npy(ve)), ("W", npy(net.conv_w.weight)), ("h", npy(h_est)), ("mW", npy(sW["exp_avg"])),
                         ("vW", npy(sW["exp_avg_sq"])), ("mh", npy(sh["exp_avg"])), ("vh", npy(sh["exp_avg_sq"]))):
                rec[k].append(v)
        consts = dict(amp=npy(amp), P=npy(Pt), var=npy(var), nu_sc=np.float64(nu_sc))
        return rec, consts, npy(q), npy(out)

    a, consts, q_last, out_last = run(1)
    b = run(8)[0]
    c = run(1, ulp_noise=True)[0]
    torch.set_num_threads(1)
    out = dict(rx=npy(rx), lr=np.float64(lr), B=np.int64(B), M=np.int64(M), loss=np.asarray(a["loss"], np.float32), var_est=np.stack(a["ve"]),
               W_steps=np.stack(a["W"]), h_steps=np.stack(a["h"]), mW_steps=np.stack(a["mW"]), vW_steps=np.stack(a["vW"]),
               mh_steps=np.stack(a["mh"]), vh_steps=np.stack(a["vh"]), q_last=q_last, out_last=out_last, snaps=np.asarray(SNAPS),
               loss_8threads=np.asarray(b["loss"], np.float32), **consts)
    for k in SNAPS:
        out[f"W_{k}"], out[f"h_{k}"] = a["W"][k - 1], a["h"][k - 1]
        out[f"W8_{k}"], out[f"h8_{k}"] = b["W"][k - 1], b["h"][k - 1]
        out[f"Wn_{k}"], out[f"hn_{k}"] = c["W"][k - 1], c["h"][k - 1]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)

    def rel(x, y):
        return float(np.abs(np.asarray(x, np.float64) - np.asarray(y, np.float64)).max() / np.abs(y).max())
    print(name, "loss[0], loss[-1]", a["loss"][0], a["loss"][-1], "reference 1 vs 8 threads, taps W/h:",
          [(k, f"{rel(b['W'][k - 1], a['W'][k - 1]):.1e}", f"{rel(b['h'][k - 1], a['h'][k - 1]):.1e}") for k in SNAPS],
          "reference with 1-ulp input noise:", [(k, f"{rel(c['W'][k - 1], a['W'][k - 1]):.1e}", f"{rel(c['h'][k - 1], a['h'][k - 1]):.1e}") for k in SNAPS])


if __name__ == "__main__":
    torch.set_num_threads(1)
    only = set(sys.argv[1:])                                 # optional: regenerate only the named fixtures
    S = [(0, 0, 0), (2, 1, 0), (-3, -1, 1)]                  # frame 0 aligned, frame 1 shifted, frame 2 shifted the other way + pol swap
    jobs = [
        ("drv_vaele_64qam", lambda n: run_driver(n, "func_VAELE_DP_MQAM_shaping", "64-QAM", 0.0270955, 23, 25, 2.5e-3, 100, 1200, 3, 10, np.pi / 10, 0.06 * np.pi, 101, S)),
        ("drv_vaeflex_16qam", lambda n: run_driver(n, "func_VAEflex_DP_MQAM_shaping", "16-QAM", 0.0, 18, 9, 2.5e-3, 100, 1000, 3, 20, np.pi / 10, 0.0, 102, S)),
        ("drv_cma_16qam", lambda n: run_driver(n, "func_CMA_DP_MQAM_shaping", "16-QAM", 0.0, 20, 25, 1e-3, 100, 1200, 3, 10, np.pi / 10, 0.0, 103, S)),
        ("drv_cmabatch_64qam", lambda n: run_driver(n, "func_CMAbatch_DP_MQAM_shaping", "64-QAM", 0.0270955, 23, 25, 1e-4, 100, 1200, 3, 10, np.pi / 10, 0.0, 104, S)),
        ("drv_cmaflex_16qam", lambda n: run_driver(n, "func_CMAflex_DP_MQAM_shaping", "16-QAM", 0.0, 20, 13, 1e-5, 100, 1200, 3, 20, np.pi / 10, 0.0, 105, S)),
        ("traj_vaele_64qam_M25_B100_100steps", trajectory),
    ]
    for name, job in jobs:
        if not only or name in only:
            job(name)
