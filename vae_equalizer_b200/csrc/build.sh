#!/usr/bin/env bash
# Build libvaeq.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.  One nvcc process per source file, in parallel.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libvaeq.so"
obj="${here}/build"
mkdir -p "${obj}"
flags=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3,-Wall ${VAEQ_NVCC_EXTRA:-})
pids=()
objs=()
for src in "${here}"/*.cu; do
    o="${obj}/$(basename "${src}" .cu).o"
    objs+=("${o}")
    nvcc "${flags[@]}" -c -o "${o}" "${src}" &
    pids+=($!)
done
for p in "${pids[@]}"; do wait "${p}"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "${out}" "${objs[@]}"
echo "built ${out}"
