#!/usr/bin/env bash
# profiles/r02c_tc_forward.txt, table of section 2: the tensor-core forward kernel with 0 / 1 / 3 centre taps kept on the CUDA cores
# (build-time FWDTC_CENTER).  Builds the three libraries (nvcc must be on the box), then times the step and runs the two accuracy tests
# for each; restores the default build at the end.  usage (repo root, one B200): bash tools/run_tc_variants.sh > gpurun_out/tc_variants.log
set -u
for v in 0 1 3; do
  VAEQ_NVCC_EXTRA="-DFWDTC_CENTER=$v" bash vae_equalizer_b200/csrc/build.sh > /dev/null 2>&1 || { echo "build failed for FWDTC_CENTER=$v"; exit 1; }
  echo "== FWDTC_CENTER $v"
  TC_FWD=1 python tools/time_step.py 2 2>&1 | tail -1 | cut -c1-200
  python -m pytest tests/test_dp_step_gpu.py -q -s -k "tensor_core_forward_against or out_error" 2>&1 | grep -E "^.?.?forward kernel|^.?.?tcgen05 vs|passed|failed" | cut -c1-220
done
bash vae_equalizer_b200/csrc/build.sh > /dev/null 2>&1
