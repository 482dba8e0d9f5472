#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in ${CONFIGS:-C2 C3a C1 C5 C4}; do
python bench.py --config $cfg --steps 5 ${NOCPU:---no-cpu} > gpurun_out/bench_r1b_$cfg.json 2> gpurun_out/bench_r1b_$cfg.err
tail -2 gpurun_out/bench_r1b_$cfg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1b_$cfg.json'))
print('$cfg', round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3), d['config'].get('orbital_grid'), {k:round(v,2) for k,v in d.get('phases_ms',{}).items()}, 'launches', d['gpu_launches']//d['steps'], d.get('band_step'))
PY
done
