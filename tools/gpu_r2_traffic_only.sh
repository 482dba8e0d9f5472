#!/bin/bash
# Re-stamp profiles/traffic.json after a source edit: ncu --set full of the H-apply sweep kernels
# of the second C2 evaluation only (k_x_vmul_cached + k_z_fwd_gather), then a C5 launch list.
mkdir -p gpurun_out
for cfg in C2 C3a; do
timeout 240 ncu --set full --clock-control none -k regex:"k_x_vmul_cached|k_yx_vmul|k_z_fwd_gather" --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_happly_${cfg} \
    python tools/profile_eval.py --config $cfg --evals 2 > gpurun_out/r02_happly_${cfg}.log 2>&1
tail -1 gpurun_out/r02_happly_${cfg}.log
ncu -i gpurun_out/r02_happly_${cfg}.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${cfg}_happly_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_ncu_full_${cfg}_happly_raw.csv | cut -c1-200
python tools/capture_traffic.py $cfg gpurun_out/r02_ncu_full_${cfg}_happly_raw.csv "profiles/r02_ncu_full_${cfg}_happly_raw.csv (ncu --set full of the H-apply sweep kernels of one evaluation)"
cp profiles/traffic.json gpurun_out/r02_traffic.json
rm -f gpurun_out/r02_happly_${cfg}.ncu-rep
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_C5.csv \
    python bench.py --config C5 --steps 2 --warmup 1 --no-cpu --no-graph > gpurun_out/r02_launches_C5.log 2>&1
grep -c jrb gpurun_out/r02_launches_C5.csv
