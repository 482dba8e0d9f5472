#!/bin/bash
# 2-GPU check of the k-sharded and the row/band-sharded layouts with the orbital grid
mkdir -p gpurun_out
for cfg in C2 C3a; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --config $cfg --no-cpu > gpurun_out/bench_r1b_${cfg}_n2.json 2> gpurun_out/bench_r1b_${cfg}_n2.err
tail -2 gpurun_out/bench_r1b_${cfg}_n2.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r1b_${cfg}_n2.json'))
print('$cfg n2', round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],2), d['config'].get('orbital_grid'), d['config']['sharding'], 'E', sum(d['energies_ha']))
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/check_row_sharded.py si8_64_k2 2>&1 | tail -3
