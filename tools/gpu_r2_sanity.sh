#!/bin/bash
# Last check of the round on one GPU: smoke() and the default bench line as the driver runs them.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_default_last.json 2> gpurun_out/r02_bench_default_last.err
tail -2 gpurun_out/r02_bench_default_last.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_default_last.json').read().strip().splitlines()[-1])
r=d['roofline']
print('C2', round(d['value'],2), d['unit'], round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value'],2), '; bound', r['bound'], 'frac', round(r['frac'],3), 'measured', r.get('measured') and round(r['measured']['frac'],3), 'design', round(r['design_bytes']['frac'],3), 'fp64', round(r['fp64']['frac'],3), 'launches', d['gpu_launches'], 'clocks', d.get('clocks'))
for k,v in d['diamond64'].items(): print(k, round(v['value'],2), round(v['ms_per_step'],3), v['roofline']['bound'], round(v['roofline']['design_bytes']['frac'],3))
print('cpu', d['cpu_baseline']['value'])
PY
