"""Why is bench.py's e2e leg at half the PCIe rate?  Times every H2D copy and every step of the double-buffered loop."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench as Bn
from vae_equalizer_b200.datagen import generate_data_gpu
from vae_equalizer_b200.dp import DPEqualizer

dev = torch.device("cuda", 0)
cst = Bn.run_constants()
B = 1 << 22
rx_dev = [generate_data_gpu(B, cst["amps"], Bn.SNR, cst["P"], 2, np.pi / 10, dev, 1234 + i)[0] for i in range(2)]
print("rx_dev", rx_dev[0].shape, rx_dev[0].stride(), rx_dev[0].is_contiguous())
eq = DPEqualizer(25, 2, cst["amp"], cst["P"], cst["var"], cst["nu_sc"], device=dev)
q = torch.empty(2, 16, B, dtype=torch.float32, device=dev)
out = torch.empty(2, 2, B, dtype=torch.float32, device=dev)
rx_host = [r.cpu().pin_memory() for r in rx_dev]
print("rx_host", rx_host[0].stride(), rx_host[0].is_pinned(), rx_host[0].is_contiguous())
NBUF = int(sys.argv[1]) if len(sys.argv) > 1 else 2
stage = [torch.empty_like(rx_dev[0]) for _ in range(NBUF)]
res_host = torch.empty(256, 3, dtype=torch.float32).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)
main = torch.cuda.current_stream()


def run(n, compute=True, d2h=True, label=""):
    ready = [torch.cuda.Event() for _ in range(NBUF)]
    free = [torch.cuda.Event() for _ in range(NBUF)]
    for s in range(NBUF):
        free[s].record(main)
    ce = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    se = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(main)
    for i in range(n):
        s = i % NBUF
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(free[s])
            ce[i][0].record(copy_stream)
            stage[s].copy_(rx_host[i % 2], non_blocking=True)
            ce[i][1].record(copy_stream)
            ready[s].record(copy_stream)
        main.wait_event(ready[s])
        se[i][0].record(main)
        if compute:
            eq.train_step(stage[s], 2.5e-3, 2.5e-3, q=q, out=out)
        se[i][1].record(main)
        free[s].record(main)
        if d2h:
            res_host[i, 0:1].copy_(eq.loss, non_blocking=True)
            res_host[i, 1:3].copy_(eq.var_est, non_blocking=True)
    t1.record(main)
    torch.cuda.synchronize()
    tot = t0.elapsed_time(t1)
    cms = [a.elapsed_time(b) for a, b in ce]
    sms = [a.elapsed_time(b) for a, b in se]
    starts_c = [t0.elapsed_time(a) for a, _ in ce]
    starts_s = [t0.elapsed_time(a) for a, _ in se]
    print(f"{label}: {tot / n:.3f} ms/step; copy avg {np.mean(cms[2:]):.3f} ms ({64 * 1.048576 / np.mean(cms[2:]):.1f} GB/s), step avg {np.mean(sms[2:]):.3f} ms")
    print("   copy starts", np.round(starts_c[:6], 2), " step starts", np.round(starts_s[:6], 2))


for _ in range(3):
    eq.train_step(rx_dev[0], 2.5e-3, 2.5e-3, q=q, out=out)
run(20, compute=False, d2h=False, label="copies only")
run(20, compute=True, d2h=False, label="copies + steps, no D2H")
run(20, compute=True, d2h=True, label="copies + steps + D2H (bench.py e2e)")
