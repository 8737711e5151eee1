mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partitioned or skew" > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sel.log
tail -4 gpurun_out/pytest_sel.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms']); print({k:round(v['ms'],2) for k,v in d['kernels'].items()})"
tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_partition -s 1 -c 1 -f -o gpurun_out/r01o_k2_partition python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/ncu_k2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_k2.log
