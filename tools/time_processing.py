"""Wall-clock of the drop-in processing() at the reference's own sizes (Eval_run_DP.py defaults)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_equalizer_b200 import processing as pr
phiIQ = np.array([0.0314, 0.0314], dtype=np.complex64)
args = ("64-QAM", 2, 23, 0, 25, 0.06 * np.pi, np.pi / 10, 2.5e-3, 100, 10000, 30, 10, "h0", 90e9, -26e-24, 0.1e-12 * np.sqrt(1000), phiIQ, 170)
for dg in ("numpy", "gpu"):
    pr.processing_vaele_dp(*args[:10], 2, *args[11:], rng=np.random.default_rng(0), verbose=False, datagen=dg)   # warm-up
    torch.cuda.synchronize(); t0 = time.perf_counter()
    SER, Var_est, var = pr.processing_vaele_dp(*args, rng=np.random.default_rng(1), verbose=False, datagen=dg)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"VAE-LE DP processing, datagen={dg}: {dt / 30 * 1e3:.1f} ms per frame of 10000 symbols ({10000 * 30 / dt:.0f} symbols/s); final SER {SER[:, -1].tolist()}")
torch.cuda.synchronize(); t0 = time.perf_counter()
SER, Var_est, var = pr.processing_vaele_dp(*args, rng=np.random.default_rng(1), verbose=False, datagen="gpu", eval_mode="fused")
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"VAE-LE DP processing, datagen=gpu, eval_mode=fused: {dt / 30 * 1e3:.1f} ms per frame of 10000 symbols ({10000 * 30 / dt:.0f} symbols/s); final SER {SER[:, -1].tolist()}")
flex = list(args); flex[9] = 2000; flex[10] = 5
for dg in ("numpy", "gpu"):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pr.processing_vaeflex_dp(*flex, rng=np.random.default_rng(1), verbose=False, datagen=dg)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"VAE-flex DP processing (2000-symbol frames, 190 steps/frame), datagen={dg}: {dt / 5 * 1e3:.1f} ms per frame")
import cProfile, pstats
prof = cProfile.Profile(); prof.enable()
pr.processing_vaele_dp(*args[:10], 10, *args[11:], rng=np.random.default_rng(2), verbose=False, datagen="gpu")
torch.cuda.synchronize(); prof.disable()
pstats.Stats(prof).sort_stats("cumulative").print_stats(14)
