"""Wall-clock of the AWGN drop-in processing() (AWGN_channel/Eval_run_shaping_vaele.py defaults, shortened): 16-QAM, M_est 25."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_equalizer_b200 import processing as pr
args = ("16-QAM", 2, 12, 0.0, 25, 2e-3, 350, 15000, 1200, 60, 20, "h1")      # mod, sps, SNR, nu, M_est, lr, batch_len, N_valid, N_train, epochs, epe, channel
pr.processing_vaele_awgn(*args[:9], 4, 2, args[11], rng=np.random.default_rng(0), verbose=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
ser = pr.processing_vaele_awgn(*args, rng=np.random.default_rng(1), verbose=False)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"AWGN VAE-LE processing: {dt / 60 * 1e3:.2f} ms per epoch (3 steps of 350 symbols + validation every 20 epochs); SER {ser.tolist()}")
import cProfile, pstats
prof = cProfile.Profile(); prof.enable()
pr.processing_vaele_awgn(*args[:9], 20, 10, args[11], rng=np.random.default_rng(2), verbose=False)
torch.cuda.synchronize(); prof.disable()
pstats.Stats(prof).sort_stats("cumulative").print_stats(12)
