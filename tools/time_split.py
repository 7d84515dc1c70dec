"""torchrun --nproc-per-node N tools/time_split.py : per-kernel CUDA-event times of the batch-split step (vaeq_kernel_timing) for both
reduction transports, next to the plain single-GPU step on the same per-rank symbol count -- where do the extra microseconds go?"""
import ctypes as C, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vae_equalizer_b200 import _lib
from vae_equalizer_b200.constants import init
from vae_equalizer_b200.datagen import generate_data_gpu
from vae_equalizer_b200.dp import DPEqualizer
from vae_equalizer_b200.parallel import BatchSplitDP
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
M, B1 = 25, 1 << 22
Bt = B1 * world
h_est, h_ch, P, amp, amps, pol, nu_sc, var, pow_mean = init("h0", "64-QAM", "cpu", 0.0270955, 2, M, 23)
rx = torch.cat([generate_data_gpu(B1, amps, 23, P, 2, np.pi / 10, dev, 11 + c)[0] for c in range(world)], dim=-1).contiguous()
names = {0: "fwd", 1: "fin(+stats exchange)", 2: "bwd1", 3: "adam(+grad exchange)", 8: "taps"}
n = 30

def timed(step, label):
    for _ in range(5):
        step()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        step()
    b.record(); torch.cuda.synchronize()
    plain = a.elapsed_time(b) / n
    lib.vaeq_kernel_timing(1)
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    ms = (C.c_float * 16)(); cnt = (C.c_int32 * 16)()
    lib.vaeq_kernel_timing_read(ms, cnt); lib.vaeq_kernel_timing(0)
    per = ", ".join(f"{names.get(k, k)}: {ms[k] / n * 1e3:.1f}" for k in range(16) if cnt[k])
    print(f"[rank {rank}] {label}: {plain * 1e3:.1f} us per step (plain launches); per step by kind (us): {per}", flush=True)
    dist.barrier()

eq = DPEqualizer(M, 2, amp, P, var, nu_sc, device=dev)
q1, o1 = torch.empty(2, 16, B1, device=dev), torch.empty(2, 2, B1, device=dev)
timed(lambda: eq.train_step(rx[:, :, :2 * B1], 2.5e-3, 2.5e-3, q=q1, out=o1), "single-GPU step, 2^22 symbols")
del q1, o1
keep_lo, keep_n = Bt // 4, Bt // 2
ot, oc = torch.zeros(2, 16, keep_n, device=dev), torch.zeros(2, 2, keep_n, device=dev)
for transport in ("peer", "nccl"):
    eq2 = DPEqualizer(M, 2, amp, P, var, nu_sc, device=dev)
    bs = BatchSplitDP(eq2, None, transport)
    timed(lambda: bs.train_step(rx, 2.5e-3, 2.5e-3, None, None, 0, ot, oc, keep_lo, keep_n), f"batch-split step, {transport}, 2^22 symbols per rank, kept columns only")
    q2, o2, col0 = bs.alloc_local(Bt)
    timed(lambda: bs.train_step(rx, 2.5e-3, 2.5e-3, q2, o2, col0), f"batch-split step, {transport}, q / out of the rank's columns written")
dist.destroy_process_group()
