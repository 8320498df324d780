#!/bin/bash
# Source-level ncu captures of the FP64-issue-bound kernels (cfg 3: fp16_kernel, cfg 4: env_real_kernel).
# Usage (under gpurun): bash tools/ncu_src.sh <tag>
TAG=${1:-r01}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fixed_point_kernel' -c 1 -f -o $OUT/prof_fp16_$TAG \
  python tools/profile_driver.py --what fp4 --reps 1 > $OUT/ncu_fp16.log 2>&1; tail -2 $OUT/ncu_fp16.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'env_real_kernel' -c 1 -f -o $OUT/prof_er8_$TAG \
  python tools/profile_driver.py --what en8 --reps 1 > $OUT/ncu_er8.log 2>&1; tail -2 $OUT/ncu_er8.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gauge_kernel|expect_kernel' -c 4 -f -o $OUT/prof_canon_$TAG \
  python tools/profile_driver.py --what canon --reps 1 > $OUT/ncu_canon.log 2>&1; tail -2 $OUT/ncu_canon.log
ls -la $OUT | tail
