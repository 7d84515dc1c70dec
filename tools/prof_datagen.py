"""Device data generator (datagen.generate_frames_gpu), R runs x 10 000 symbols: wall time per call and the kernel list (torch profiler)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from vae_equalizer_b200 import shared_funcs as sfun
from vae_equalizer_b200.datagen import generate_frames_gpu
R = int(os.environ.get("R", 592))
c = sfun.init("h0", "64-QAM", "cuda", 0.0270955, 2, 25, 23)
P = np.tile(np.asarray(c[2])[None], (R, 1)); amps = c[4]
snr = [15 + 2 * (i % 8) for i in range(R)]; th = [0.3] * R
for i in range(3): generate_frames_gpu(10000, amps, snr, P, 2, th, "cuda", i)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(5): generate_frames_gpu(10000, amps, snr, P, 2, th, "cuda", i)
torch.cuda.synchronize(); print("ms per call", (time.perf_counter() - t0) / 5 * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    generate_frames_gpu(10000, amps, snr, P, 2, th, "cuda", 7); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
