#!/bin/bash
# GPU job: FP64 ceilings, parity tests, C2 bench, launch list, ncu --set full of each hot kernel.
mkdir -p gpurun_out
./tools/fp64_peak.bin > gpurun_out/fp64_peak.txt 2>&1
cat gpurun_out/fp64_peak.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --config C2 --steps 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 1500 gpurun_out/bench_c2.json
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"k_x_inv_density|k_x_vmul|k_y_inv|k_z_inv_scatter|k_y_fwd|k_z_fwd_gather|k_gram|k_apply|k_chol_inv" \
  --launch-skip ${SKIP:-48} --launch-count ${COUNT:-48} -f -o gpurun_out/prof_C2 \
  python tools/profile_eval.py --config C2 --evals 2 > gpurun_out/prof_C2.log 2>&1
tail -3 gpurun_out/prof_C2.log
ls -la gpurun_out/
