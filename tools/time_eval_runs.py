"""Batched per-frame evaluation (vaeq_frame_eval_runs) alone: R runs x N symbols, CUDA-event time and the kernel list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_equalizer_b200 import shared_funcs as sfun
R, N = int(os.environ.get("R", 592)), int(os.environ.get("N", 10000))
dev = "cuda:0"
h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = sfun.init("h0", "64-QAM", dev, 0.0270955, 2, 25, 23)
n = int(amp.numel())
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, n, (R, 2, 2, N), device=dev, generator=g)
tx = amp.to(dev)[idx].to(torch.float16)
out = (amp.to(dev)[idx] + 0.05 * torch.randn(R, 2, 2, N, device=dev, generator=g)).float().roll(3, -1).contiguous()
q = torch.stack([sfun.soft_dec(out[r], var, amp, nu_sc) for r in range(R)])
var_all = var.reshape(1, 2).repeat(R, 1)
nu_all = torch.full((R,), float(nu_sc), device=dev)
f = lambda: sfun.frame_eval_runs(q, out, tx, amp, var_all, nu_all, 100)
for _ in range(3): ser, al = f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): f()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
byt = R * N * (16 * n + 16 + 8)
print(f"frame_eval_runs R={R} N={N}: {ms * 1e3:.1f} us per call, {byt / ms / 1e6:.0f} GB/s over one read of q, out, tx; align[0] {al[0].tolist()} ser[0] {ser[0].tolist()}")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    f(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))
