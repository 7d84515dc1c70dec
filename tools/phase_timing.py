"""Debug: per-phase cycles of one CTA of k_dp_fwd_fast (library built with -DVAEQ_PHASE_TIMING)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_equalizer_b200 import _lib
from vae_equalizer_b200.constants import init
from vae_equalizer_b200.datagen import generate_data_gpu
from vae_equalizer_b200.dp import DPEqualizer
lib = _lib.load()
h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = init("h0", "64-QAM", "cpu", 0.0270955, 2, 25, 23)
B = 1 << 22
rx = generate_data_gpu(B, amps, 23, P, 2, np.pi / 10, "cuda", 1)[0]
eq = DPEqualizer(25, 2, amp, P, var, nu_sc)
q = torch.empty(2, 16, B, device="cuda"); out = torch.empty(2, 2, B, device="cuda")
for _ in range(3):
    eq.train_step(rx, 2.5e-3, 2.5e-3, q=q, out=out)
torch.cuda.synchronize()
raw = C.CDLL(_lib.LIB_PATH)
NC = 296
buf = (C.c_ulonglong * (12 + 3 * NC))()
raw.vaeq_debug_phase_cycles(buf)
names = ["load_x", "sync1", "FIR", "pointwise+q stores", "m1s+sync2", "D+e stores", "sync3", "loop/prefetch"]
tot = sum(list(buf)[:8])
for n, v in zip(names, list(buf)[:8]):
    print(f"{n:22s} {v:10d} cycles  {v / tot:6.3f}")
print("total", tot, "tiles per CTA ~", (B + 2031) // 2032 / 148)
print("loop cycles", buf[8], "loop wall ns", buf[9], "=> SM clock GHz", buf[8] / max(buf[9], 1))
t0 = np.array(list(buf)[12:12 + NC], dtype=np.float64); t1 = np.array(list(buf)[12 + NC:12 + 2 * NC], dtype=np.float64); sm = np.array(list(buf)[12 + 2 * NC:12 + 3 * NC])
base = t0.min()
dur = (t1 - t0) / 1e3
print("CTA loop start offset us: min %.1f med %.1f max %.1f" % (((t0 - base) / 1e3).min(), np.median((t0 - base) / 1e3), ((t0 - base) / 1e3).max()))
print("CTA loop duration us: min %.1f p10 %.1f med %.1f p90 %.1f max %.1f" % (dur.min(), np.percentile(dur, 10), np.median(dur), np.percentile(dur, 90), dur.max()))
print("last CTA end offset us %.1f" % ((t1.max() - base) / 1e3))
order = np.argsort(dur)
print("slowest CTAs (id, sm, dur):", [(int(i), int(sm[i]), round(float(dur[i]), 1)) for i in order[-8:]])
print("fastest CTAs (id, sm, dur):", [(int(i), int(sm[i]), round(float(dur[i]), 1)) for i in order[:8]])
persm = {}
for i in range(NC): persm.setdefault(int(sm[i]), []).append(float(dur[i]))
print("CTAs per SM histogram:", np.bincount([len(v) for v in persm.values()]))
np.save("gpurun_out/cta_dur.npy", np.stack([np.arange(NC), sm, dur, (t0 - base) / 1e3]))
