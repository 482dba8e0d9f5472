#!/bin/bash
# psi(r) cache follow-up: parity of the paths it touches, then the bench (cache on).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_full_size_oracle_gpu.py tests/test_orbital_grid_gpu.py -m gpu -x -q ) > gpurun_out/r02_psi2_pytest.log 2>&1
tail -4 gpurun_out/r02_psi2_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_psi2_bench_on.json 2> gpurun_out/r02_psi2_bench_on.err
tail -3 gpurun_out/r02_psi2_bench_on.err | cut -c1-300
python - <<'PY'
import json
def load(f):
  return json.loads(open(f).read().strip().splitlines()[-1])
def show(n,d):
  r=d['roofline']
  print(n, round(d['value'],2), d['unit'], round(d['ms_per_step'],3),'ms', 'e2e',round(d['e2e']['value'],2), 'fp64', round(r.get('fp64',{}).get('frac',0),3), {k:round(v,3) for k,v in d.get('phases_ms',{}).items()}, 'launches', d['gpu_launches']//d['steps'])
d=load('gpurun_out/r02_psi2_bench_on.json'); show('C2',d)
for k,v in d.get('diamond64',{}).items(): show(k,v)
PY
