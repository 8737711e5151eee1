mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for cp in smem; do
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --count-path $cp > gpurun_out/bench_n1_$cp.json 2> gpurun_out/bench_n1_$cp.err; echo "bench rc=$?" >> gpurun_out/bench_n1_$cp.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n1_$cp.json')); print('$cp value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['phases_ms']); print({k:round(v['ms'],2) for k,v in d['kernels'].items()})"
tail -3 gpurun_out/bench_n1_$cp.err
done

for spec in k2_partition:6 k_count_smem:6; do
  kn=${spec%%:*}; skip=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kn -s $skip -c 1 -f -o gpurun_out/r01m_$kn python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/ncu_$kn.log 2>&1
  echo "ncu $kn rc=$?"
done
