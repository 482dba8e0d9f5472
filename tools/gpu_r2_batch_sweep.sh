#!/bin/bash
# Batch-size sweep (band groups per launch) with the psi(r) cache on: smaller batches keep the
# column work space in L2 between the z passes and the plane kernels, larger ones balance the
# persistent CTAs better.
mkdir -p gpurun_out
for cfg in ${CFGS:-C2 C3b}; do
for bg in ${BGS:-16 32 48 64 96 128 0}; do
  JRB_BATCH_GROUPS=$bg python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bg_${cfg}_$bg.json 2> gpurun_out/r02_bg_${cfg}_$bg.err
  python - <<PY
import json
try:
  d=json.loads(open('gpurun_out/r02_bg_${cfg}_$bg.json').read().strip().splitlines()[-1])
  print('$cfg', 'batch_groups', '$bg', round(d['value'],2), 'eval/s', round(d['ms_per_step'],3), 'ms', {k:round(v,3) for k,v in d.get('phases_ms',{}).items()}, 'launches', d['gpu_launches']//d['steps'])
except Exception as e:
  print('$cfg', '$bg', 'failed', e)
PY
done
done
