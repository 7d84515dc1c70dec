"""Wall-clock of the CMA-family drop-in processing() at the Eval_run_DP.py defaults (10 000-symbol frames, M_est 25, batch_len 100,
flex_step 10).  SURVEY.md §6, the reference on 8 host cores: CMA ~1.0 k, CMAbatch ~1.2 k, CMAflex ~1.1 k symbols/s."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_equalizer_b200 import processing as pr
phiIQ = np.array([0.0314, 0.0314], dtype=np.complex64)
args = ["64-QAM", 2, 23, 0, 25, 0.06 * np.pi, np.pi / 10, 1e-3, 100, 10000, 20, 10, "h0", 90e9, -26e-24, 0.1e-12 * np.sqrt(1000), phiIQ, 170]
for name, fn, lr in (("CMA", pr.processing_cma_dp, 1e-3), ("CMAbatch", pr.processing_cmabatch_dp, 2e-5), ("CMAflex", pr.processing_cmaflex_dp, 2e-5)):
    args[7] = lr                                         # the batch variants SUM the increments of batch_len symbols (sf:424-433)
    for dg in ("numpy", "gpu"):
        a = list(args); a[10] = 2
        fn(*a, rng=np.random.default_rng(0), verbose=False, datagen=dg)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        SER, _, _ = fn(*args, rng=np.random.default_rng(1), verbose=False, datagen=dg)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name:9s} datagen={dg:5s}: {dt / 20 * 1e3:7.2f} ms per frame of 10000 symbols ({10000 * 20 / dt / 1e3:8.1f} k symbols/s); last-frame SER {[round(v, 4) for v in SER[:, -1].tolist()]}")
