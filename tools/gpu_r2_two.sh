#!/bin/bash
# Round 2, N GPUs (gpurun --gpus N): multi-GPU parity over NVLink peer memory, then the bench at N.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12
export JRB_COMM_TIMEOUT_S=10
( time timeout 600 python -m pytest tests/test_full_size_oracle_gpu.py -k two_gpus -x -q ) > gpurun_out/r2_pytest_n2.log 2>&1
tail -25 gpurun_out/r2_pytest_n2.log
for layout in k rows; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_parity.py $layout 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -8
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
( time timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -5 gpurun_out/r2_bench_n$N.err | cut -c1-400
( time JRB_NO_PEER=1 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --config C2 ) > gpurun_out/r2_bench_n${N}_nccl.json 2> gpurun_out/r2_bench_n${N}_nccl.err
python - <<PY
import json
def show(n,d):
  r=d['roofline']
  print(n, d['config'].get('sharding'), d['config'].get('reduce_path'), round(d['value'],2),'eval/s', round(d['ms_per_step'],3),'ms e2e',round(d['e2e']['value'],2),'fp64 frac',round(r['fp64']['frac'],3), {k:round(v,3) for k,v in d.get('phases_ms',{}).items()})
for f in ['gpurun_out/r2_bench_n$N.json','gpurun_out/r2_bench_n${N}_nccl.json']:
  try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    show(f,d)
    for k,v in d.get('diamond64',{}).items(): show(k,v)
  except Exception as e: print(f,'parse failed',e)
PY
