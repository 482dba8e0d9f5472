#!/bin/bash
# Round 2, final single-GPU record: the -m gpu suite, the default bench line (C2 + diamond64 + CPU
# sample), every other configuration, the reference arm, and the stamped traffic capture of C2.
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log
# traffic capture first: profiles/traffic.json then carries the stamp of these very sources
n=$(python - <<'PY'
print(55)
PY
)
timeout 900 ncu --set full --clock-control none --launch-skip 56 --launch-count 55 -f -o gpurun_out/r02_final_C2 \
    python tools/profile_eval.py --config C2 --evals 2 > gpurun_out/r02_final_C2.log 2>&1
ncu -i gpurun_out/r02_final_C2.ncu-rep --page raw --csv > gpurun_out/r02_final_ncu_full_C2_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_final_ncu_full_C2_raw.csv > gpurun_out/r02_final_ncu_full_C2_summary.txt
python tools/capture_traffic.py C2 gpurun_out/r02_final_ncu_full_C2_raw.csv "profiles/r02_final_ncu_full_C2_raw.csv (ncu --set full of one whole evaluation, 55 launches)"
cp profiles/traffic.json gpurun_out/r02_final_traffic.json
rm -f gpurun_out/r02_final_C2.ncu-rep
grep -c . gpurun_out/r02_final_ncu_full_C2_summary.txt
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -3 gpurun_out/r02_bench_default.err
for cfg in C1 C4 C5; do
  python bench.py --config $cfg --steps 10 > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err
done
python bench.py --config C2 --orbital-grid full --steps 10 --no-cpu > gpurun_out/r02_bench_C2_fullgrid.json 2>/dev/null
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python - <<'PY'
import json
def load(f):
  return json.loads(open(f).read().strip().splitlines()[-1])
def show(n,d):
  r=d['roofline']
  print(n, round(d['value'],2), d['unit'], round(d['ms_per_step'],3),'ms', d['config'].get('launch','')[:10], 'e2e',round(d['e2e']['value'],2), 'copies', round(d['e2e'].get('copies_alone_ms',0),2), 'hbm frac',round(r['frac'],3),'fp64', round(r.get('fp64',{}).get('frac',0),3), 'traffic', r.get('traffic'), {k:round(v,3) for k,v in d.get('phases_ms',{}).items()}, 'launches', d['gpu_launches']//d['steps'])
d=load('gpurun_out/r02_bench_default.json'); show('C2',d)
for k,v in d.get('diamond64',{}).items(): show(k,v)
print('cpu', d.get('cpu_baseline'))
for cfg in ('C1','C4','C5','C2_fullgrid'):
  try: show(cfg, load(f'gpurun_out/r02_bench_{cfg}.json'))
  except Exception as e: print(cfg, 'failed', e)
r=load('gpurun_out/r02_bench_reference.json'); print('reference', r['value'], r['cpu_baseline']['sample'][:120])
PY
