#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_orbital_grid_gpu.py -m gpu -x -q 2>&1 | tail -4
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
for a in "$@"; do run ${a%%:*} ${a##*:}; done
