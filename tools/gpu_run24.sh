mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms']); print({k:round(v['ms'],2) for k,v in d['kernels'].items()}); print(d['cpu_baseline'])"
tail -3 gpurun_out/bench_n1.err
