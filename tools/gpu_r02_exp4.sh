# round 2, batch 4 (one B200): adaptive segment capacities — tests, skew + GC experiment, default bench
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/exp_skew.py --gc 0.4,0.3,0.2 > gpurun_out/r02_exp_skew.jsonl 2> gpurun_out/exp_skew.err; echo "skew rc=$?"; cut -c1-420 gpurun_out/r02_exp_skew.jsonl; tail -3 gpurun_out/exp_skew.err
LRB_K2_UNIFORM=1 timeout 900 python tools/exp_skew.py --fractions "" --gc 0.4,0.3,0.2 > gpurun_out/r02_exp_skew_uniform_caps.jsonl 2>> gpurun_out/exp_skew.err; echo "skew uniform rc=$?"; cut -c1-420 gpurun_out/r02_exp_skew_uniform_caps.jsonl
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python tools/bench_summary.py gpurun_out/bench_n1.json
