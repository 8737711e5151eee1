mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
# launch list of one bench run (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches_r01.csv
# full capture of the two dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_count15|k_search15" -s 2 -c 2 -o gpurun_out/prof_r01_count_search python bench.py --steps 1 --warmup 3 --no-cpu-baseline --reads 200000 > gpurun_out/bench_under_ncu2.log 2>&1
ls -la gpurun_out
