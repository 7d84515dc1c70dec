"""One call of every evaluation operator at N = 2^22 (64-QAM) for an ncu capture (profiles/r01d_*): soft_dec, SER_IQflip,
SER_constell_shaping, find_shift, find_shift_symb_full, GMI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_equalizer_b200 import shared_funcs as sfun
dev = "cuda:0"
h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = sfun.init("h0", "64-QAM", dev, 0.0270955, 2, 25, 23)
n, N = int(amp.numel()), 1 << int(os.environ.get("LOGN", 22))
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, n, (2, 2, N), device=dev, generator=g)
tx = amp.to(dev)[idx].to(torch.float16)
out = (amp.to(dev)[idx] + 0.05 * torch.randn(2, 2, N, device=dev, generator=g)).float().contiguous()
for rep in range(int(os.environ.get("REPS", 2))):
    q = sfun.soft_dec(out, var, amp, nu_sc)
    sfun.SER_IQflip(q, tx)
    sfun.SER_constell_shaping(out.clone(), tx, amp, nu_sc, var)
    sfun._find_shift(q, None, tx, 21, amp, False, sync=False)
    sfun._find_shift(None, out, tx, 21, None, False, sync=False)
    sfun.GMI(q, tx, P)
torch.cuda.synchronize()
print("ok")
