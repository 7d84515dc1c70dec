"""torchrun --nproc-per-node N tools/h2d_numa_probe.py : why do the host->device links not add up at N >= 4 (bench.py e2e)?
Every rank copies 128 MiB pinned buffers to its GPU, all ranks AT THE SAME TIME, (a) with the pinned buffer wherever the process
happened to run, (b) with the process bound to the CPUs of its GPU's NUMA node before the buffer is allocated and touched; next to
that the box's topology (NUMA nodes, CPUs, GPU <-> node affinity) and a host memcpy rate per rank as the memory-bandwidth yardstick."""
import os, subprocess, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N = 128 << 20


def gpu_numa_node():
    bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True).stdout.strip()
    path = f"/sys/bus/pci/devices/{bus.lower()[4:] if bus.lower().startswith('0000') and len(bus) > 12 else bus.lower()}/numa_node"
    for cand in (path, f"/sys/bus/pci/devices/{bus.lower()}/numa_node", f"/sys/bus/pci/devices/0000:{bus.lower()[-7:]}/numa_node"):
        try:
            return bus, int(open(cand).read())
        except Exception:
            continue
    return bus, -1


def node_cpus(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
    except Exception:
        return None
    cpus = []
    for part in txt.split(","):
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def h2d_rate(host, reps=12):
    dst = torch.empty(N, dtype=torch.uint8, device=dev)
    for _ in range(2):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(host, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return N * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


bus, node = gpu_numa_node()
if rank == 0:
    print(subprocess.run("nproc; lscpu | grep -E 'Model name|Socket|NUMA|Thread|Core'; numactl -H 2>/dev/null | head -20; nvidia-smi topo -m | head -16", shell=True,
                         capture_output=True, text=True).stdout, flush=True)
aff0 = sorted(os.sched_getaffinity(0))
a = torch.empty(N, dtype=torch.uint8).pin_memory(); a.fill_(1)
t0 = time.perf_counter(); b = a.clone(); host_gbs = 2 * N / (time.perf_counter() - t0) / 1e9        # read + write
r_default = h2d_rate(a)
del a, b
cpus = node_cpus(node) if node >= 0 else None
bound = False
if cpus:
    try:
        os.sched_setaffinity(0, set(cpus) & set(aff0) or set(cpus))
        bound = True
    except Exception:
        pass
c = torch.empty(N, dtype=torch.uint8).pin_memory(); c.fill_(2)          # first touch on the GPU's node when the binding worked
r_bound = h2d_rate(c)                                                   # every rank runs the same sequence of collectives
vals = torch.tensor([r_default, r_bound], dtype=torch.float64, device=dev)
allv = [torch.zeros_like(vals) for _ in range(world)]
dist.all_gather(allv, vals)
print(f"[rank {rank}] GPU {local} bus {bus} NUMA node {node}; process affinity {aff0[0]}..{aff0[-1]} ({len(aff0)} CPUs); host clone {host_gbs:.1f} GB/s; "
      f"H2D with all {world} ranks copying: {r_default:.1f} GB/s default placement, {r_bound:.1f} GB/s {'bound to the GPU node' if bound else '(no NUMA information: second run, same placement)'}", flush=True)
if rank == 0:
    print(f"aggregate H2D: {sum(float(v[0]) for v in allv):.1f} GB/s default, {sum(float(v[1]) for v in allv):.1f} GB/s NUMA-bound ({world} ranks at once; one link alone: ~55 GB/s)", flush=True)
dist.destroy_process_group()
