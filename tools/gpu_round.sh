#!/bin/bash
# One GPU-box visit: peaks, parity tests, benches, ncu launch list + full captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.txt 2>&1
nproc > $OUT/nproc.txt
[ -x tools/peaks ] && timeout 120 tools/peaks > $OUT/peaks_$TAG.json 2> $OUT/peaks.err; cat $OUT/peaks_$TAG.json
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 600 python tools/bench_configs.py --cfg 2,3,4,5 --c64 > $OUT/configs_$TAG.jsonl 2> $OUT/configs.err; echo "configs rc=$?"
cat $OUT/configs_$TAG.jsonl; tail -3 $OUT/configs.err
timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench.err
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/ncu_launch.log 2>&1
# full captures of the hot kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:env_d2_stream -s 2 -c 2 -f -o $OUT/prof_env_d2_$TAG \
    python tools/profile_driver.py --what d2 > $OUT/ncu_d2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fixed_point_kernel|env_generic_kernel|zgemm_dmma' -c 6 -f -o $OUT/prof_generic_$TAG \
    python tools/profile_driver.py --what fp4,en8,pw64 --reps 1 > $OUT/ncu_generic.log 2>&1
ls -la $OUT
