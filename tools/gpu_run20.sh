mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv,noheader > gpurun_out/box.txt; nproc >> gpurun_out/box.txt; free -g | head -2 >> gpurun_out/box.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --durations=5 -k "partitioned or skew or config2 or synthetic_reads" > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sel.log
tail -12 gpurun_out/pytest_sel.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms']); print({k:round(v['ms'],2) for k,v in d['kernels'].items()}); print(d['cpu_baseline']); print(d['clocks']); print(d['roofline']['frac'], d['roofline']['peak_source'])"
tail -3 gpurun_out/bench_n1.err
LRB_SEARCH_UNROLL=8 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_u8.json 2> gpurun_out/bench_n1_u8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_u8.json')); print('U8 value',d['value'],'ms',d['ms_per_step']); print({k:round(v['ms'],2) for k,v in d['kernels'].items()})"
