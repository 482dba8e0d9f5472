#!/bin/bash
# per-kernel times of one evaluation with an orbital grid (ncu launch list)
mkdir -p gpurun_out
for og in ${OGS:-64,64,49 64,64,50}; do
export JRB_ORBITAL_GRID=$og
bash tools/gpu_launch_list.sh ${CFG:-C2} 300 | grep -E "k_z_|k_yx|kernel"
cp gpurun_out/launches_${CFG:-C2}.csv gpurun_out/launches_${CFG:-C2}_og_$og.csv
done
