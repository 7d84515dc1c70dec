"""SURVEY.md §8(d), last rows: the evaluation kernels as HBM scans (achieved GB/s against MEASURED_PEAKS.json) and the CMA
family as symbols/s for one stream and for S independent streams batched in one launch (vaeq_cma n_runs).  CUDA-event timing,
3 warm-ups, inputs larger than L2 at the large N (q alone is 512 MiB at N = 2^22)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_equalizer_b200 import _lib, shared_funcs as sfun

dev = "cuda:0"
peak = 6537.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = sfun.init("h0", "64-QAM", dev, 0.0270955, 2, 25, 23)
n = int(amp.numel())
print(f"HBM peak used: {peak:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs or the 6537 fallback); 64-QAM, n_lev {n}")
for logN in (13, 17, 22, 24):
    N = 1 << logN
    g = torch.Generator(device=dev).manual_seed(1)
    idx = torch.randint(0, n, (2, 2, N), device=dev, generator=g)
    tx = amp.to(dev)[idx].to(torch.float16)
    out = (amp.to(dev)[idx] + 0.05 * torch.randn(2, 2, N, device=dev, generator=g)).float().contiguous()
    q = sfun.soft_dec(out, var, amp, nu_sc)
    reps = 20 if logN <= 22 else 5
    rows = [
        ("soft_dec (out -> q)", lambda: sfun.soft_dec(out, var, amp, nu_sc), 16 + 16 * n),
        ("SER_IQflip (q, tx)", lambda: sfun.SER_IQflip(q, tx), 16 * n + 8),
        ("SER_constell_shaping (out, tx)", lambda: sfun.SER_constell_shaping(out, tx, amp, nu_sc, var), 2 * 16 + 8 + 16),
        ("find_shift (q, tx), 21 shifts", lambda: sfun._find_shift(q, None, tx, 21, amp, False, sync=False), 8 * n + 8),
        ("find_shift_symb_full (out, tx)", lambda: sfun._find_shift(None, out, tx, 21, None, False, sync=False), 8 + 8),
        ("GMI (q, tx) [extension]", lambda: sfun.GMI(q, tx, P), 16 * n + 8),
    ]
    for name, fn, bps in rows:
        ms = timed(fn, reps)
        gbs = bps * N / ms / 1e6
        print(f"N=2^{logN:<2d} {name:34s} {ms * 1e3:9.1f} us  {N / ms / 1e6:8.2f} G symbols/s  {gbs:7.0f} GB/s algorithmic ({bps} B/symbol) = {gbs / peak:5.3f} of HBM peak")
    del q, out, tx, idx

# CMA family: S independent streams per launch (one warp / one CTA per run)
lib = _lib.load()
Ns, M = 20000, 25
for mode, name, lr, bl, st in ((0, "CMA", 1e-3, 0, 0), (1, "CMAbatch", 1e-5, 100, 0), (2, "CMAflex", 1e-6, 200, 20)):
    for S in (1, 148, 592, 2368):
        g = torch.Generator(device=dev).manual_seed(2)
        Rx = torch.randn(S, 2, 2, Ns, device=dev, generator=g) * 0.7
        h0 = torch.zeros(S, 2, 2, 2, M, device=dev)
        h0[:, 0, 0, 0, M // 2] = 1
        h0[:, 1, 1, 0, M // 2] = 1
        h = h0.clone()
        out = torch.zeros(S, 2, 2, Ns // 2, device=dev)
        e = torch.empty(S, Ns // 2, 2, device=dev)
        scr = torch.empty(int(lib.vaeq_cma_scratch_bytes(Ns, M, S)), dtype=torch.uint8, device=dev)

        def run():
            h.copy_(h0)
            _lib.check(lib.vaeq_cma(mode, Rx.data_ptr(), Ns, 1.0, h.data_ptr(), M, lr, bl, st, 2, 1, out.data_ptr(), e.data_ptr(), S,
                                    scr.data_ptr(), _lib.current_stream()), "vaeq_cma")
        ms = timed(run, 5)
        sym = S * Ns // 2
        print(f"{name:9s} S={S:5d} streams x {Ns // 2} symbols: {ms:8.3f} ms  {sym / ms / 1e3:10.1f} M symbols/s  ({56 * sym / ms / 1e6:7.1f} GB/s at 56 B/symbol)")
