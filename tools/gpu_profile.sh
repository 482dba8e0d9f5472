#!/bin/bash
# ncu --set full capture of the hot kernels (C2), one eval after one warm-up eval.
mkdir -p gpurun_out
CFG=${1:-C2}
# kernels per eval at default batch (C2): ~63 -> skip the first eval
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:"k_x_inv_density|k_x_vmul|k_y_inv|k_z_inv_scatter|k_y_fwd|k_z_fwd_gather|k_gram|k_apply|k_chol_inv|k_sphere_reduce" \
  --launch-skip ${SKIP:-45} --launch-count ${COUNT:-24} -f -o gpurun_out/prof_${CFG} \
  python tools/profile_eval.py --config $CFG --evals 2 > gpurun_out/prof_${CFG}.log 2>&1
tail -5 gpurun_out/prof_${CFG}.log
ls -la gpurun_out/
