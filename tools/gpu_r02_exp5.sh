# round 2, batch 5 (one B200): vectorised mirror A/B, full GPU tier incl. the config-#1 byte compare
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/exp5_mirror_v4.json 2> gpurun_out/exp5.err; echo "v4 rc=$?"
LRB_MIRROR_SCALAR=1 $B > gpurun_out/exp5_mirror_scalar.json 2>/dev/null; echo "scalar rc=$?"
python tools/bench_summary.py gpurun_out/exp5_mirror_v4.json gpurun_out/exp5_mirror_scalar.json
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
