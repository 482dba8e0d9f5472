#!/bin/bash
# orbital grid: parity tests + sweep of the z length
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_orbital_grid_gpu.py -m gpu -x -q 2>&1 | tail -8
run() { # cfg og
  python bench.py --config $1 --steps 5 --no-cpu --orbital-grid $2 > gpurun_out/bench_og_$1_$2.json 2> gpurun_out/bench_og_$1_$2.err || tail -3 gpurun_out/bench_og_$1_$2.err
  python - <<PY
import json
try:
  d=json.load(open('gpurun_out/bench_og_$1_$2.json'))
  print('$1 $2', round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],2), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'E', sum(d['energies_ha']))
except Exception as e:
  print('$1 $2 FAILED', e)
PY
}
for og in full 64,64,49 64,64,50 64,64,54 64,64,56 64,64,60; do run C2 $og; done
for og in full 128,128,80 128,128,81 128,128,90 128,128,96; do run C3a $og; done
for og in full 64,64,45 64,64,48 48,48,48 48,48,45; do run C4 $og; done
