#!/bin/bash
# FIRST GPU call of the next session (usage: tools/gpu.sh 900 bash tools/gpu_first_call_next_round.sh).
# The tests below were written after round 1's GPU budget was spent; their bodies pass on the
# CPU stand-in (tests/emulated_plan.py) but have not run on hardware yet.  Run them first, WITHOUT
# -x, then the whole suite; remove a file from tests/conftest.py:NOT_YET_RUN_ON_HARDWARE once green.
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu \
  tests/test_reference_golden.py tests/test_autograd.py tests/test_pseudopotential.py \
  tests/test_drivers_host.py tests/test_utils_api.py \
  2>&1 | tail -25 | tee gpurun_out/new_gpu_tests.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/all_gpu_tests.log
python bench.py --steps 5 --no-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 600 gpurun_out/bench_default.json
# Then (separate calls; see DESIGN.md section 9 for what each one decides):
#   tools/gpu.sh 900 bash tools/gpu_profile_orbital_grid.sh      # ncu --set full of one C2 evaluation: baseline of the round
#   gpurun --gpus 2 -- bash tools/gpu_check_2gpu.sh              # row-sharded evaluation bit-for-bit, C2 / C3a at N = 2
#   gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
#       --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json'   # the 80 % target
#   JRB_HOST_CHUNKS=... / --emulate-ranks 8 are the single-GPU tuning aids for the e2e path and the per-rank share.
