"""Wall time of the sweep engine (sweep.sweep_vae_dp) at the reference's settings: batch_len 100, 10 000-symbol frames, 64-QAM,
M_est 25; R cells x F frames.  Prints the split between data generation, training and evaluation."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_equalizer_b200 import sweep
R, F = int(os.environ.get("R", 64)), int(os.environ.get("F", 6))
cells = [dict(SNR=15 + 2 * (i % 8), nu=0.0270955, lr_optim=2.5e-3, theta=np.pi / 10, theta_diff=0.06 * np.pi, seed=i) for i in range(R)]
for ev, dg in ((1, "gpu"), (F + 1, "gpu"), (1, "gpu_batched"), (F + 1, "gpu_batched")):
    sweep.sweep_vae_dp(cells[:2], "64-QAM", 2, 25, 100, 10000, 1, kind="VAE", eval_every=1, datagen=dg)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ser, ve, var = sweep.sweep_vae_dp(cells, "64-QAM", 2, 25, 100, 10000, F, kind="VAE", eval_every=ev, datagen=dg)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"R={R} cells x F={F} frames, datagen={dg}, eval_every={ev}: {dt:.2f} s wall = {dt / F * 1e3:.1f} ms/frame, {R * F * 10000 / dt / 1e6:.2f} M symbols/s; last-frame SER[0] {ser[0, :, -1].tolist()}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
sweep.sweep_vae_dp(cells, "64-QAM", 2, 25, 100, 10000, 3, kind="VAE", eval_every=1, datagen="gpu_batched"); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
