mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --reads 200000"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_partition -s 3 -c 1 -o gpurun_out/prof_r01_k_partition $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search_keys -s 200 -c 2 -o gpurun_out/prof_r01_k_search_keys $B > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bucket_hist|k_composition" -s 6 -c 2 -o gpurun_out/prof_r01_k_hist_comp $B > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
