mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/traffic_step.csv python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/traffic_step.log 2>&1; echo "traffic rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_search_keys -c 1 -f -o gpurun_out/r01v9_k_search_keys python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/ncu_search.log 2>&1; echo "ncu rc=$?"
