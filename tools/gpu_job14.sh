#!/bin/bash
mkdir -p gpurun_out
for er in 2 4 8; do
python bench.py --config C2 --steps 10 --no-cpu --emulate-ranks $er > gpurun_out/bench_er$er.json 2> gpurun_out/bench_er$er.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_er$er.json'))
print('er$er', round(d['ms_per_step'],3),'ms (ideal', round(13.14/$er,3),')', {k:round(v,3) for k,v in d.get('phases_ms',{}).items()}, 'launches', d['gpu_launches']//d['steps'], d['driver_step'])
PY
done
