#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite (with the BASELINE-shape oracle tests), the
# default bench line (C2 + diamond64) and the reference arm.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
free -g | head -2; nproc
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu.log
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -5 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
try:
  d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
  def show(n,d):
    r=d['roofline']
    print(n, round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],2),'hbm frac',round(r['frac'],3),'fp64 frac',round(r['fp64']['frac'],3), {k:round(v,3) for k,v in d.get('phases_ms',{}).items()}, 'launches', d['gpu_launches']//d['steps'])
  show('C2',d)
  for k,v in d.get('diamond64',{}).items(): show(k,v)
  print('cpu', d.get('cpu_baseline'))
except Exception as e: print('bench parse failed', e)
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
cat gpurun_out/r2_bench_reference.json | cut -c1-600
tail -3 gpurun_out/r2_bench_reference.err
