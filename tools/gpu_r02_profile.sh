# round 2: GPU tier, default bench, launch list and ncu --set full captures of the hot kernels (one B200)
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python tools/bench_summary.py gpurun_out/bench_n1.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "launch list rc=$?"
for K in k_search_keys k2_partition k_partition k_count_smem; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$K -c 1 -f -o gpurun_out/r02_$K python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/ncu_$K.log 2>&1; echo "ncu $K rc=$?"
  python tools/ncu_summary.py gpurun_out/r02_$K.ncu-rep gpurun_out/r02_ncu_full_$K.csv
done
ls -la gpurun_out/*.ncu-rep
