#!/bin/bash
# Round 2: psi(r) cache A/B on one GPU.  Parity first (the whole -m gpu suite runs with the cache
# on by default; test_energy_and_grad also covers JRB_PSI_CACHE_MB=0), then the bench with and
# without the cache on the same box.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_psi_pytest.log 2>&1
tail -6 gpurun_out/r02_psi_pytest.log
( time python bench.py --steps 20 --warmup 5 --no-cpu ) > gpurun_out/r02_psi_bench_on.json 2> gpurun_out/r02_psi_bench_on.err
tail -3 gpurun_out/r02_psi_bench_on.err | cut -c1-300
JRB_PSI_CACHE_MB=0 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_psi_bench_off.json 2> gpurun_out/r02_psi_bench_off.err
for cfg in C1 C4; do
  python bench.py --config $cfg --steps 10 --no-cpu > gpurun_out/r02_psi_bench_$cfg.json 2> gpurun_out/r02_psi_bench_$cfg.err
done
python - <<'PY'
import json
def load(f):
  return json.loads(open(f).read().strip().splitlines()[-1])
def show(n,d):
  r=d['roofline']
  print(n, round(d['value'],2), d['unit'], round(d['ms_per_step'],3),'ms', 'e2e',round(d['e2e']['value'],2), 'fp64', round(r.get('fp64',{}).get('frac',0),3), {k:round(v,3) for k,v in d.get('phases_ms',{}).items()}, 'launches', d['gpu_launches']//d['steps'])
for tag in ('on','off'):
  try:
    d=load(f'gpurun_out/r02_psi_bench_{tag}.json'); show('C2 cache '+tag,d)
    for k,v in d.get('diamond64',{}).items(): show(k+' cache '+tag,v)
  except Exception as e: print(tag,'failed',e)
for cfg in ('C1','C4'):
  try: show(cfg, load(f'gpurun_out/r02_psi_bench_{cfg}.json'))
  except Exception as e: print(cfg, 'failed', e)
PY
