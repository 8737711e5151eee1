mkdir -p gpurun_out
for m in strided block strided block; do
LRB_K2_ROWS=$m timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$m.json')); print('$m value',d['value'],'ms',d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['kernels'].items()})"
done
