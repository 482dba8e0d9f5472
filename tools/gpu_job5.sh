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
for bg in 32 96; do
JRB_BATCH_GROUPS=$bg python bench.py --config C2 --steps 5 --no-cpu > gpurun_out/bench_C2_bg$bg.json 2> gpurun_out/bench_C2_bg$bg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_C2_bg$bg.json'))
print('C2 bg$bg', round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches']//d['steps'])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_yx|k_z_|k_gram|k_apply|k_rho" \
    --launch-skip 24 --launch-count 12 -f -o gpurun_out/prof_fused2 \
    python tools/profile_eval.py --config C2 --evals 2 > gpurun_out/prof_fused2.log 2>&1
ncu -i gpurun_out/prof_fused2.ncu-rep --page raw --csv > gpurun_out/prof_fused2_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
