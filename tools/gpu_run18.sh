mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?" >> gpurun_out/bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms']); print({k:round(v['ms'],2) for k,v in d['kernels'].items()}); print(d['cpu_baseline']); print(d['clocks'])"
tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "launch list rc=$?"
for spec in k_search_keys:10 k_count_smem:10 k_composition:0; do
  kn=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -f -o gpurun_out/r01j_$kn python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/ncu_$kn.log 2>&1
  echo "ncu $kn rc=$?"
done
