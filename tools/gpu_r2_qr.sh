#!/bin/bash
# QR small-chain kernel: parity tests that exercise it, then the emulated 8-rank share and N=1 bench
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py tests/test_full_size_gpu.py tests/test_full_size_oracle_gpu.py -m gpu -x -q ) > gpurun_out/r2_pytest_qr.log 2>&1
tail -6 gpurun_out/r2_pytest_qr.log
for er in 8 1; do
  for fused in 0 1; do
    JRB_NO_FUSED_SMALL=$((1-fused)) python bench.py --config C2 --steps 10 --no-cpu --emulate-ranks $er > gpurun_out/r2_qr_er${er}_f$fused.json 2> gpurun_out/r2_qr_er${er}_f$fused.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/r2_qr_er${er}_f$fused.json').read().strip().splitlines()[-1])
print('emulate $er fused $fused:', round(d['value'],2),'eval/s', round(d['ms_per_step'],4),'ms', {k:round(v,3) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches']//d['steps'])
PY
  done
done
