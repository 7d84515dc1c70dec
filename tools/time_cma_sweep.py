"""Wall time of the batched CMA sweep engine (sweep.sweep_cma_dp) at the reference's settings: 10 000-symbol frames, 64-QAM, M_est 25;
R cells x F frames per kind; next to the single-run drivers' per-cell rate (tools/time_cma.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_equalizer_b200 import sweep
R, F = int(os.environ.get("R", 148)), int(os.environ.get("F", 4))
for kind, lr in (("CMA", 1e-3), ("CMAbatch", 1e-5), ("CMAflex", 1e-6)):
    cells = [dict(SNR=15 + 2 * (i % 8), nu=0.0270955, lr_optim=lr, theta=np.pi / 10, theta_diff=0.06 * np.pi, seed=i) for i in range(R)]
    sweep.sweep_cma_dp(cells, "64-QAM", 2, 25, 100, 10000, 1, 20, kind=kind, datagen="gpu_batched")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ser, _, _ = sweep.sweep_cma_dp(cells, "64-QAM", 2, 25, 100, 10000, F, 20, kind=kind, datagen="gpu_batched")
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{kind:9s} R={R} cells x F={F} frames: {dt / F * 1e3:8.1f} ms/frame = {dt / F / R * 1e3:6.3f} ms per cell and frame, {R * F * 10000 / dt / 1e6:7.2f} M symbols/s; "
          f"last-frame SER[0] {[round(v, 4) for v in ser[0, :, -1].tolist()]}")
