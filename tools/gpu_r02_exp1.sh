# round 2 experiments on one B200: writing count (no table memset), mirror beside/after the search, evict_last gathers,
# bulk-copy run copy-out; then the changed GPU tests and the default bench
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/exp_default.json 2> gpurun_out/exp_default.err; echo "default rc=$?"
$B --mirror after > gpurun_out/exp_mirror_after.json 2>/dev/null; echo "mirror-after rc=$?"
LRB_SEARCH_HINT=1 $B > gpurun_out/exp_search_hint.json 2>/dev/null; echo "hint rc=$?"
LRB_PART_BULK=1 $B > gpurun_out/exp_part_bulk.json 2>/dev/null; echo "bulk rc=$?"
python tools/bench_summary.py gpurun_out/exp_default.json gpurun_out/exp_mirror_after.json gpurun_out/exp_search_hint.json gpurun_out/exp_part_bulk.json
LRB_PART_BULK=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partitioned or skew or synthetic" > gpurun_out/pytest_bulk.log 2>&1; echo "pytest bulk rc=$?"; tail -3 gpurun_out/pytest_bulk.log
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python tools/bench_summary.py gpurun_out/bench_n1.json
