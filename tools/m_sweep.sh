#!/usr/bin/env bash
# M_est / batch_len sweep of the DP step: shows where the step leaves the FP32-bound regime.
for m in 5 9 13 25; do
  python bench.py --m-est $m --steps 20 --no-cpu --no-small 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('M_est=%2d  batch 2^22: %.4f ms/step  %6.2f G symbols/s  whole-step %.3f of HBM roofline (176 B/symbol)  dominant %s %.3f   kernels us:' % ($m, d['ms_per_step'], d['value']/1e9, r['whole_step_frac'], r['kernel'], r['frac']), {k[:12]:round(v['avg_ms']*1e3,1) for k,v in r['kernels'].items()})"
done
for b in 20 24; do
  python bench.py --batch-log2 $b --steps 20 --no-cpu --no-small 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('M_est=25  batch 2^$b: %.4f ms/step  %6.2f G symbols/s  whole-step %.3f of HBM roofline' % (d['ms_per_step'], d['value']/1e9, r['whole_step_frac']))"
done
