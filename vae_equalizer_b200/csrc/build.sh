#!/usr/bin/env bash
# Build libvaeq.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="${here}/../libvaeq.so"
srcs=("${here}"/*.cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC,-O3,-Wall -shared \
     ${VAEQ_NVCC_EXTRA:-} -o "${out}" "${srcs[@]}"
echo "built ${out}"
