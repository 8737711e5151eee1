mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partitioned or skew" > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sel.log
tail -4 gpurun_out/pytest_sel.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms']); print({k:round(v['ms'],2) for k,v in d['kernels'].items()})"
tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_step.csv python tools/prof_step.py --reads 1000000 --steps 2 > gpurun_out/prof_step.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_step.csv | head -12
