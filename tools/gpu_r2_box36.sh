#!/bin/bash
# 36 x 36 fused plane kernels for 48^3 grids (C5): parity, A/B against the 48 x 48 x 36 box, and the
# traffic re-stamp for the sources that now instantiate them.
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_orbital_grid_gpu.py -m gpu -x -q -k "si_48 or 36_box" ) > gpurun_out/r02_box36_pytest.log 2>&1
tail -3 gpurun_out/r02_box36_pytest.log
python bench.py --config C5 --steps 10 --no-cpu > gpurun_out/r02_box36_C5_auto.json 2> gpurun_out/r02_box36_C5_auto.err
python bench.py --config C5 --steps 10 --no-cpu --orbital-grid 36,36,36 > gpurun_out/r02_box36_C5_36.json 2> gpurun_out/r02_box36_C5_36.err
python - <<'PY'
import json
for t in ('auto','36'):
  try:
    d=json.loads(open(f'gpurun_out/r02_box36_C5_{t}.json').read().strip().splitlines()[-1])
    print('C5', t, d['config'].get('orbital_grid'), round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'ms')
  except Exception as e: print(t, 'failed', e)
PY
tail -2 gpurun_out/r02_box36_C5_36.err | cut -c1-300
for cfg in C2 C3a; do
timeout 240 ncu --set full --clock-control none -k regex:"k_x_vmul_cached|k_yx_vmul|k_z_fwd_gather" --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_happly_${cfg} \
    python tools/profile_eval.py --config $cfg --evals 2 > gpurun_out/r02_happly_${cfg}.log 2>&1
ncu -i gpurun_out/r02_happly_${cfg}.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${cfg}_happly_raw.csv 2>/dev/null
python tools/capture_traffic.py $cfg gpurun_out/r02_ncu_full_${cfg}_happly_raw.csv "profiles/r02_ncu_full_${cfg}_happly_raw.csv (ncu --set full of the H-apply sweep kernels of one evaluation)"
cp profiles/traffic.json gpurun_out/r02_traffic_box36.json
rm -f gpurun_out/r02_happly_${cfg}.ncu-rep
done
