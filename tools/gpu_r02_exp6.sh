# round 2, batch 6 (one B200): composition beside the search — device-resident step, host pipeline, parity
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B --composition beside > gpurun_out/exp6_comp_beside.json 2> gpurun_out/exp6.err; echo "beside rc=$?"
LRB_COMP_BESIDE=1 $B > gpurun_out/exp6_host_beside.json 2>> gpurun_out/exp6.err; echo "host beside rc=$?"
python tools/bench_summary.py gpurun_out/exp6_comp_beside.json gpurun_out/exp6_host_beside.json
LRB_COMP_BESIDE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dropin or synthetic or contigs or config1_full" > gpurun_out/pytest_comp_beside.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_comp_beside.log
