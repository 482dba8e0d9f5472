#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for cfg in C2 C1 C3a; do
python bench.py --config $cfg --steps 5 --no-cpu > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
tail -3 gpurun_out/bench_$cfg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$cfg.json'))
print('$cfg', round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'fft',round(d['roofline']['fft_density_path']['frac'],3), {k:round(v,2) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches']//d['steps'])
PY
done
for ch in 1 4 16; do
JRB_HOST_CHUNKS=$ch python bench.py --config C2 --steps 3 --no-cpu > gpurun_out/bench_C2_ch$ch.json 2> gpurun_out/bench_C2_ch$ch.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_C2_ch$ch.json'))
print('C2 host chunks $ch', round(d['value'],2),'eval/s e2e',round(d['e2e']['value'],2))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches_c2.csv python tools/profile_eval.py --config C2 --evals 3 > gpurun_out/launches.log 2>&1
ls -la gpurun_out | head -30
