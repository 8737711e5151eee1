mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partitioned or skew" 2>&1 | tail -2
LRB_K2_LAYOUT=cta timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partitioned or skew" 2>&1 | tail -2
for m in "sub strided" "cta strided" "cta block"; do
set -- $m
LRB_K2_LAYOUT=$1 LRB_K2_ROWS=$2 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2.json 2> gpurun_out/bench_$1_$2.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$1_$2.json')); print('$m value',d['value'],'ms',d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()})"
done
