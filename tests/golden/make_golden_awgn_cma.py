"""Golden vectors for the AWGN single-polarisation CMA module, from the UNMODIFIED reference
(`AWGN_channel/func_CMA_MQAM_shaping.py`: CMA :142-168, CPE :170-196, SER_CMA :63-93, find_shift_symb :127-140, processing :201-256).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_awgn_cma.py
A separate script because the reference module switches autograd off globally at import (`torch.set_grad_enabled(False)`, :14).
The reference's generate_data is unseeded; the frames come from this package's seeded host-side restatement of it
(vae_equalizer_b200.awgn.generate_data) and are replayed into `processing()` by replacing the module's `generate_data` attribute."""
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
from vae_equalizer_b200.awgn import generate_data  # noqa: E402
from vae_equalizer_b200.constants import awgn_constants, upsampled_channel  # noqa: E402

cm = ref_loader.load("func_CMA_MQAM_shaping", subdir="AWGN_channel")


def npy(t):
    return t.detach().cpu().numpy().copy() if torch.is_tensor(t) else np.asarray(t)


def frames_for(mod, nu, SNR, channel, sizes, seed):
    h_channel = upsampled_channel(channel, 2)
    M = (len(h_channel) - 1) // 2 + 1
    amps, P, _, _ = awgn_constants(mod, nu, SNR)
    rng = np.random.default_rng(seed)
    return [generate_data(N, M, amps, SNR, h_channel, 2, "cpu", P, rng=rng) for N in sizes], amps


def case_ops(name, mod, M_est, N_train, N_valid, lr, SNR, channel, seed, phase=0.3):
    """One training pass, one evaluation pass, CPE, find_shift_symb, SER_CMA -- each function's inputs and outputs."""
    (fr, amps) = frames_for(mod, 0.0, SNR, channel, [N_train, N_train, N_valid], seed)
    amp = torch.tensor(amps, dtype=torch.float32)
    h = torch.zeros(2, M_est)
    h[0, M_est // 2] = 1
    rec = dict(amp=npy(amp), lr=np.float64(lr), M=np.int64(M_est))
    for i in range(2):                                            # two sequential training frames: taps carried over
        rec[f"rx{i}"], rec[f"h_in{i}"] = npy(fr[i][0]), npy(h)
        out, h, e = cm.CMA(fr[i][0].clone(), 1, h, lr, 2, True)
        rec[f"out{i}"], rec[f"h_out{i}"], rec[f"e{i}"] = npy(out), npy(h), npy(e)
    rx_v, tx_v = fr[2]
    # a slow phase ramp on the validation frame so that CPE has something to remove (and would need unwrapping, which this CPE lacks)
    t = torch.arange(rx_v.shape[1], dtype=torch.float32)
    ph = phase + 2e-5 * t
    rx_rot = torch.stack((rx_v[0] * torch.cos(ph) - rx_v[1] * torch.sin(ph), rx_v[1] * torch.cos(ph) + rx_v[0] * torch.sin(ph)))
    out_v, _, e_v = cm.CMA(rx_rot.clone(), 1, h, lr, 2, False)
    rec["rx_v"], rec["tx_v"], rec["out_v"], rec["e_v"] = npy(rx_rot), npy(tx_v), npy(out_v), npy(e_v)
    cpe = cm.CPE(out_v)
    rec["cpe"] = npy(cpe)
    shift = cm.find_shift_symb(cpe, tx_v, 21)
    rec["shift"] = np.int64(int(shift))
    a, b = cpe[:, 11 + shift:-11].clone(), tx_v[:, 11:-11 - shift]
    rec["ser_in"], rec["ser_tx"] = npy(a), npy(b)
    ser = cm.SER_CMA(a, b, 2, amp, len(amps), "cpu")
    rec["ser_in_scaled"], rec["ser"] = npy(a), npy(ser)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "shift", int(shift), "ser", float(ser), "e mean", float(e.abs().mean()))


def case_driver(name, mod, M_est, lr, SNR, nu, channel, N_valid, N_train, num_epochs, epe, seed):
    sizes = []
    for ep in range(num_epochs):
        sizes.append(N_train)
        if ep % epe == 0:
            sizes.append(N_valid)
    fr, amps = frames_for(mod, nu, SNR, channel, sizes, seed)
    it = iter(fr)
    shifts = []
    orig_gen, orig_fs = cm.generate_data, cm.find_shift_symb

    def fs(*a, **k):
        s = orig_fs(*a, **k)
        shifts.append(int(s))
        return s
    cm.generate_data = lambda N, *a, **k: tuple(t.clone() for t in next(it))
    cm.find_shift_symb = fs
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            SER = cm.processing(mod, 2, SNR, nu, M_est, lr, N_valid, N_train, num_epochs, epe, channel)
    finally:
        cm.generate_data, cm.find_shift_symb = orig_gen, orig_fs
    rec = dict(SER=npy(SER), shifts=np.asarray(shifts), args=np.array([SNR, nu, M_est, lr, N_valid, N_train, num_epochs, epe], dtype=np.float64),
               mod=np.array(mod), channel=np.array(channel), n_frames=np.int64(len(fr)))
    for i, (rx, tx) in enumerate(fr):
        rec[f"rx{i}"], rec[f"tx{i}"] = npy(rx), npy(tx)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, "SER", npy(SER).tolist(), "shifts", shifts)


if __name__ == "__main__":
    torch.set_num_threads(1)
    case_ops("awgncma_ops_16qam_M25", "16-QAM", 25, 800, 3000, 5e-4, 22, "h1", seed=61)
    case_ops("awgncma_ops_64qam_M9", "64-QAM", 9, 500, 2000, 2e-4, 26, "h2", seed=62, phase=-0.5)
    case_driver("awgncma_drv_16qam", "16-QAM", 25, 5e-4, 22, 0.0, "h1", 3000, 1500, 6, 2, seed=63)
