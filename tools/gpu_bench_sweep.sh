#!/bin/bash
# GPU box: first measurements (bench lines, batch-size sweep, ncu launch list).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
python bench.py --config C1 --steps 20 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
python bench.py --config C2 --steps 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for bg in 1 8 32 128; do
  JRB_BATCH_GROUPS=$bg python bench.py --config C2 --steps 3 --no-cpu > gpurun_out/bench_c2_bg$bg.json 2> gpurun_out/bench_c2_bg$bg.err
done
python bench.py --config C3a --steps 5 > gpurun_out/bench_c3a.json 2> gpurun_out/bench_c3a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_c2.csv python bench.py --config C2 --steps 1 --warmup 1 --no-cpu \
  > gpurun_out/ncu_c2.log 2>&1
tail -c 600 gpurun_out/bench_c1.json gpurun_out/bench_c2.json gpurun_out/bench_c3a.json
for f in gpurun_out/*.err; do echo "== $f"; tail -3 $f; done
