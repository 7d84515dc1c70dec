"""Launch-side overhead of the large-batch DP step (batch_len 2^22): plain loop, loop with the library's per-kernel event pairs
(what bench.py's timed region does), and the same steps replayed from a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vae_equalizer_b200 import _lib, shared_funcs as sfun
from vae_equalizer_b200.datagen import generate_data_gpu
from vae_equalizer_b200.dp import DPEqualizer
dev = torch.device("cuda", 0)
lib = _lib.load()
h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = sfun.init("h0", "64-QAM", dev, 0.0270955, 2, 25, 23)
B, NB, K = 1 << 22, 3, 48
rx = [generate_data_gpu(B, amps, 23, P, 2, np.pi / 10, dev, 1234 + i)[0] for i in range(NB)]
eq = DPEqualizer(25, 2, amp, P, var, nu_sc, device=dev)
q = torch.empty(2, 16, B, device=dev); out = torch.empty(2, 2, B, device=dev)
def loop(n):
    for i in range(n): eq.train_step(rx[i % NB], 2.5e-3, 2.5e-3, q=q, out=out)
def timed(fn):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K
print(f"plain loop              {timed(lambda: loop(K)) * 1e3:8.1f} us/step")
_lib.check(lib.vaeq_kernel_timing(1))
print(f"with per-kernel events  {timed(lambda: loop(K)) * 1e3:8.1f} us/step")
_lib.check(lib.vaeq_kernel_timing(0))
s = torch.cuda.Stream(device=dev)
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    loop(3)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loop(NB)
print(f"CUDA graph of {NB} steps   {timed(lambda: [g.replay() for _ in range(K // NB)]) * 1e3:8.1f} us/step")
print("loss", float(eq.loss))
