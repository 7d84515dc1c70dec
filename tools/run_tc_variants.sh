for v in 0 1 3; do
  if [ $v != 0 ]; then cp libvaeq_c$v.so vae_equalizer_b200/libvaeq.so; fi
  echo "== CENTER $v"
  TC_FWD=1 python tools/time_step.py 2 2>&1 | tail -1 | cut -c1-200
  python -m pytest tests/test_dp_step_gpu.py -q -s -k "tensor_core_forward or out_error" 2>&1 | grep -E "^.?.?forward kernel|^.?.?tcgen05 vs|passed|failed" | cut -c1-220
done
