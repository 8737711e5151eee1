mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_partition|k_count_keys|k_search_keys|k_bucket_hist|k_composition" -s 20 -c 12 -o gpurun_out/prof_r01_partitioned python bench.py --steps 1 --warmup 3 --no-cpu-baseline --reads 200000 > gpurun_out/bench_under_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
