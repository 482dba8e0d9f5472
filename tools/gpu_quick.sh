#!/bin/bash
# quick check: GPU parity tests + C2/C3a/C1 bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in ${CONFIGS:-C2 C3a C1}; do
python bench.py --config $cfg --steps 5 --no-cpu > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
tail -3 gpurun_out/bench_$cfg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$cfg.json'))
print('$cfg', round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'fft',round(d['roofline']['fft_density_path']['frac'],3), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches']//d['steps'])
PY
done
