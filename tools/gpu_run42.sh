mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partitioned or skew or config1 or building" > gpurun_out/pytest_sel.log 2>&1; tail -2 gpurun_out/pytest_sel.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()}); print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'])"
