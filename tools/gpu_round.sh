#!/bin/bash
# One GPU-box visit: parity tests, benches, D=2 sweep, ncu captures.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh [tag] [what...]
#   what: peaks test configs bench sweep launches ncu ncutc smoke cpu dist memcheck racecheck ncuthread
TAG=${1:-r01}; shift
WHAT=${@:-"test configs bench sweep ncu"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $OUT/smi_$TAG.txt 2>&1
nproc > $OUT/nproc.txt
for w in $WHAT; do
case $w in
peaks) [ -x tools/peaks ] && timeout 120 tools/peaks > $OUT/peaks_$TAG.json 2> $OUT/peaks.err; cat $OUT/peaks_$TAG.json;;
test) timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu_$TAG.log
      tail -25 $OUT/pytest_gpu_$TAG.log;;
configs) timeout 600 python tools/bench_configs.py --cfg 2,3,4,5,6,7,8 --c64 > $OUT/configs_$TAG.jsonl 2> $OUT/configs.err; echo "configs rc=$?"
      cat $OUT/configs_$TAG.jsonl; tail -3 $OUT/configs.err;;
bench) timeout 600 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json; tail -3 $OUT/bench.err;;
sweep) timeout 300 python tools/sweep_d2.py > $OUT/sweep_d2_$TAG.jsonl 2> $OUT/sweep.err; cat $OUT/sweep_d2_$TAG.jsonl; tail -3 $OUT/sweep.err;;
launches) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
      python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/ncu_launch.log 2>&1;;
ncu) timeout 300 ncu --set full --clock-control none --import-source on -k regex:'env_d2_stream' -s 2 -c 1 -f -o $OUT/prof_env_d2_$TAG \
      python tools/profile_driver.py --what d2 > $OUT/ncu_d2.log 2>&1
     timeout 300 ncu --set full --clock-control none --import-source on -k regex:'fixed_point_kernel|env_real_kernel|env_generic_kernel|zgemm_dmma' -c 4 -f -o $OUT/prof_generic_$TAG \
      python tools/profile_driver.py --what fp4,en8,pw64 --reps 1 > $OUT/ncu_generic.log 2>&1;;
ncutc) timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cgemm_tc_kernel|gauge_kernel|expect_kernel|fp16_kernel' -c 6 -f -o $OUT/prof_tc_$TAG \
      python tools/profile_driver.py --what tc64,tc256,fp4,canon --reps 1 > $OUT/ncu_tc.log 2>&1; tail -3 $OUT/ncu_tc.log;;
smoke) timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; tail -2 $OUT/smoke_$TAG.log;;
cpu) timeout 200 python tools/cpu_baselines.py --seconds 3 > $OUT/cpu_baselines_$TAG.jsonl 2> $OUT/cpu_baselines.err; cut -c1-160 $OUT/cpu_baselines_$TAG.jsonl;;
dist) timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > $OUT/dist_check_$TAG.json 2> $OUT/dist_check.err; cat $OUT/dist_check_$TAG.json;;
memcheck) timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_driver.py > $OUT/sanitize_memcheck_$TAG.log 2>&1; tail -3 $OUT/sanitize_memcheck_$TAG.log;;
racecheck) timeout 330 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_driver.py reduced > $OUT/sanitize_racecheck_$TAG.log 2>&1; tail -3 $OUT/sanitize_racecheck_$TAG.log;;
ncuthread) timeout 200 ncu --set full --clock-control none --import-source on -k regex:'fp_d2_kernel|bw_cost_thread_kernel' -c 2 -f -o $OUT/prof_thread_$TAG \
      python tools/profile_driver.py --what fpd2,bw --reps 1 > $OUT/ncu_thread.log 2>&1; tail -2 $OUT/ncu_thread.log;;
esac
done
ls -la $OUT | tail -30
