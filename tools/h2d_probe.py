"""Host->device copy bandwidth of this box for the e2e leg of bench.py: one 64 MiB pinned buffer per step, copied as
1/2/4/8 chunks on as many streams, plus a write-combined cudaHostAlloc buffer through the runtime directly."""
import ctypes as C
import subprocess
import time

import torch

print(subprocess.run(["nvidia-smi", "--query-gpu=name,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max",
                      "--format=csv"], capture_output=True, text=True).stdout)
print(subprocess.run("nproc; lscpu | grep -E 'Model name|Socket|NUMA' ; nvidia-smi topo -m | head -12", shell=True, capture_output=True, text=True).stdout)

dev = torch.device("cuda", 0)
N = 64 << 20
host = torch.empty(N, dtype=torch.uint8).pin_memory()
host.fill_(3)
dst = torch.empty(N, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()


def bench(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    return N * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, N * reps / wall / 1e9


for nchunk in (1, 2, 4, 8):
    streams = [torch.cuda.Stream(device=dev) for _ in range(nchunk)]
    ch = N // nchunk

    def fn():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        for i, s in enumerate(streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                dst[i * ch:(i + 1) * ch].copy_(host[i * ch:(i + 1) * ch], non_blocking=True)
            e = torch.cuda.Event()
            e.record(s)
            cur.wait_event(e)

    g, w = bench(fn)
    print(f"torch pinned, {nchunk} chunk(s)/stream(s): {g:.1f} GB/s (events), {w:.1f} GB/s (wall)")

rt = C.CDLL("libcudart.so.12")
for flags, name in ((0, "cudaHostAllocDefault"), (4, "cudaHostAllocWriteCombined")):
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(N), C.c_uint(flags)) == 0
    C.memset(p, 5, N)
    st = torch.cuda.current_stream().cuda_stream

    def fn():
        assert rt.cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), p, C.c_size_t(N), C.c_int(1), C.c_void_p(st)) == 0

    g, w = bench(fn)
    print(f"{name}: {g:.1f} GB/s (events), {w:.1f} GB/s (wall)")
    rt.cudaFreeHost(p)

# device -> host for completeness
back = torch.empty(N, dtype=torch.uint8).pin_memory()
g, w = bench(lambda: back.copy_(dst, non_blocking=True))
print(f"D2H torch pinned: {g:.1f} GB/s")
