#!/bin/bash
# ncu --set full capture of one launch of each hot kernel (C2), in the second evaluation.
# Launch order of the regex-matching kernels per evaluation at the default batch (5 batches):
#   0-5 gram chol apply gram chol apply | 6-20 (z y xdens) x5 | 21-45 (z y xvmul yfwd zfwd) x5 | 46 gram 47 apply
mkdir -p gpurun_out
CFG=${1:-C2}
PER_EVAL=${PER_EVAL:-48}
RE='regex:k_x_inv_density|k_x_vmul|k_y_inv|k_z_inv_scatter|k_y_fwd|k_z_fwd_gather|k_gram|k_apply|k_chol_inv'
i=0
for spec in "0 9" "$((PER_EVAL-27)) 5" "$((PER_EVAL-2)) 2"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k "$RE" \
    --launch-skip $((PER_EVAL + $1)) --launch-count $2 -f -o gpurun_out/prof_${CFG}_$i \
    python tools/profile_eval.py --config $CFG --evals 2 > gpurun_out/prof_${CFG}_$i.log 2>&1
  ncu -i gpurun_out/prof_${CFG}_$i.ncu-rep --page raw --csv > gpurun_out/prof_${CFG}_${i}_raw.csv 2>/dev/null
  i=$((i+1))
done
ls -la gpurun_out/
