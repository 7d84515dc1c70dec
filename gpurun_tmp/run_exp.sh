cp vae_equalizer_b200/libvaeq.so /tmp/base.so
for v in "$@"; do
  if [ $v != base ]; then cp gpurun_tmp/libvaeq_$v.so vae_equalizer_b200/libvaeq.so; else cp /tmp/base.so vae_equalizer_b200/libvaeq.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu --no-small 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],4), {k[:12]:round(v['avg_ms'],4) for k,v in d['roofline']['kernels'].items()})"
done
cp /tmp/base.so vae_equalizer_b200/libvaeq.so
