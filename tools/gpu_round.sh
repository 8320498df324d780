mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
lscpu | head -20 > gpurun_out/lscpu.txt; nproc >> gpurun_out/lscpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:env_d2 -s 4 -c 2 -o gpurun_out/prof_env_d2 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
