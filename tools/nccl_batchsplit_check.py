"""torchrun --nproc-per-node N tools/nccl_batchsplit_check.py : BatchSplitDP over NCCL vs the single-GPU step (rank 0 checks)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_equalizer_b200.constants import init  # noqa: E402
from vae_equalizer_b200.datagen import generate_data_gpu  # noqa: E402
from vae_equalizer_b200.dp import DPEqualizer  # noqa: E402
from vae_equalizer_b200.parallel import BatchSplitDP  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
M, B = 25, 496 * 128
h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = init("h0", "64-QAM", "cpu", 0.0270955, 2, M, 23)
rx, tx, _ = generate_data_gpu(B, amps, 23, P, 2, np.pi / 10, dev, 77)          # same seed on every rank = replicated window
eq = DPEqualizer(M, 2, amp, P, var, nu_sc, device=dev)
bs = BatchSplitDP(eq)
q = torch.zeros(2, 16, B, device=dev)
out = torch.zeros(2, 2, B, device=dev)
losses = []
for step in range(5):
    loss, ve, (lo, hi) = bs.train_step(rx, 2.5e-3, 2.5e-3, q, out)
    losses.append(float(loss))
# every rank must hold bit-identical parameters
Wl = [torch.zeros_like(eq.W) for _ in range(world)]
dist.all_gather(Wl, eq.W)
same = all(torch.equal(Wl[0], w) for w in Wl)
if rank == 0:
    ref = DPEqualizer(M, 2, amp, P, var, nu_sc, device=dev)
    rl = []
    for step in range(5):
        _, _, l, _ = ref.train_step(rx, 2.5e-3, 2.5e-3)
        rl.append(float(l))
    relW = float((eq.W - ref.W).abs().max() / ref.W.abs().max())
    relh = float((eq.h - ref.h).abs().max() / ref.h.abs().max())
    rell = max(abs(a - b) / abs(b) for a, b in zip(losses, rl))
    ok = same and relW < 1e-4 and relh < 1e-4 and rell < 1e-5
    print(f"NCCL batch-split world={world}: params identical on all ranks={same}, W rel {relW:.2e}, h rel {relh:.2e}, loss rel {rell:.2e} -> {'OK' if ok else 'FAIL'}")
    if not ok:
        sys.exit(1)
dist.destroy_process_group()
