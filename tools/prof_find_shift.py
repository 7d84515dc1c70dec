"""Per-kernel GPU time of find_shift / find_shift_symb_full at N = 2^22 (torch profiler): correlation kernels vs the decide kernel."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from vae_equalizer_b200 import shared_funcs as sfun
dev="cuda:0"
h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = sfun.init("h0", "64-QAM", dev, 0.0270955, 2, 25, 23)
n=8; N=1<<22
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, n, (2, 2, N), device=dev, generator=g)
tx = amp.to(dev)[idx].to(torch.float16)
out = (amp.to(dev)[idx] + 0.05 * torch.randn(2, 2, N, device=dev, generator=g)).float().contiguous()
q = sfun.soft_dec(out, var, amp, nu_sc)
for _ in range(3):
    sfun._find_shift(q, None, tx, 21, amp, False, sync=False); sfun._find_shift(None, out, tx, 21, None, False, sync=False)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    sfun._find_shift(q, None, tx, 21, amp, False, sync=False); sfun._find_shift(None, out, tx, 21, None, False, sync=False); torch.cuda.synchronize()
for e in prof.key_averages():
    if e.device_time_total > 0: print(f"{e.key[:70]:70s} {e.device_time_total:10.1f} us  x{e.count}")
