"""GPU time split of one frame of the batched CMA sweep engine (sweep.sweep_cma_dp, 592 cells; KIND = CMA | CMAbatch | CMAflex)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from vae_equalizer_b200 import sweep
R = 592
kind = os.environ.get("KIND", "CMAbatch"); lr = {"CMA": 1e-3, "CMAbatch": 1e-5, "CMAflex": 1e-6}[kind]
cells = [dict(SNR=15 + 2 * (i % 8), nu=0.0270955, lr_optim=lr, theta=np.pi / 10, theta_diff=0.06 * np.pi, seed=i) for i in range(R)]
sweep.sweep_cma_dp(cells, "64-QAM", 2, 25, 100, 10000, 2, 20, kind=kind, datagen="gpu_batched"); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    sweep.sweep_cma_dp(cells, "64-QAM", 2, 25, 100, 10000, 2, 20, kind=kind, datagen="gpu_batched"); torch.cuda.synchronize()
ev = sorted((e for e in prof.key_averages() if e.device_time_total > 0), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ev)
print(kind, "GPU time per frame", tot / 2 / 1e3, "ms")
for e in ev[:14]: print(f"{e.key[:80]:80s} {e.device_time_total / 2:9.1f} us/frame x{e.count // 2}")
