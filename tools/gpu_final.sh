#!/bin/bash
# round-end check: the whole GPU suite, smoke(), the default bench line (with the CPU baseline leg)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final_default.json 2> gpurun_out/bench_final_default.err
tail -2 gpurun_out/bench_final_default.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_final_default.json'))
print(d['config']['workload'], round(d['value'],2),'eval/s e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3),'cpu',d.get('cpu_baseline'),'clocks',d['clocks'])
PY
[ -n "$SKIP_NCU" ] && exit 0
# ncu --set full of the fused 81 x 81 plane kernels (C3a), two launches
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_yx" --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_C3a_f81 \
    python tools/profile_eval.py --config C3a --evals 2 > gpurun_out/prof_C3a_f81.log 2>&1
ncu -i gpurun_out/prof_C3a_f81.ncu-rep --page raw --csv > gpurun_out/prof_C3a_f81_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/prof_C3a_f81_raw.csv > gpurun_out/prof_C3a_f81_summary.txt
rm -f gpurun_out/prof_C3a_f81.ncu-rep
cut -c1-220 gpurun_out/prof_C3a_f81_summary.txt
