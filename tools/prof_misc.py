"""One pass over the kernels OUTSIDE the large-batch DP step, for `ncu --set full -k regex:...` (profiles/r02_ncu_misc_*): the sweep
engine's frame (batched data generation k_gen_*, persistent training k_dp_frame_fast, batched evaluation k_er_*), the CMA family
(k_cma_sample, k_cma_block, k_cpe_*) and the AWGN step (k_awgn_step).  Each is run twice (the first call allocates / sets attributes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_equalizer_b200 import sweep, processing as pr
R = int(os.environ.get("R", 592))
PART = os.environ.get("PART", "vae,cma,awgn").split(",")
cells = [dict(SNR=15 + 2 * (i % 8), nu=0.0270955, lr_optim=2.5e-3, theta=np.pi / 10, theta_diff=0.06 * np.pi, seed=i) for i in range(R)]
for _ in range(2 if "vae" in PART else 0):
    sweep.sweep_vae_dp(cells, "64-QAM", 2, 25, 100, 10000, 1, kind="VAE", datagen="gpu_batched", device="cuda")
torch.cuda.synchronize()
for kind, lr in ((("CMA", 1e-3), ("CMAbatch", 1e-5), ("CMAflex", 1e-6)) if "cma" in PART else ()):
    cc = [dict(c, lr_optim=lr) for c in cells]
    for _ in range(2):
        sweep.sweep_cma_dp(cc, "64-QAM", 2, 25, 100, 10000, 1, 20, kind=kind, datagen="gpu_batched")
torch.cuda.synchronize()
args = ("16-QAM", 2, 12, 0.0, 25, 2e-3, 350, 15000, 1200, 4, 2, "h1")
for _ in range(2 if "awgn" in PART else 0):
    pr.processing_vaele_awgn(*args, rng=np.random.default_rng(0), verbose=False)
torch.cuda.synchronize()
print("done")
