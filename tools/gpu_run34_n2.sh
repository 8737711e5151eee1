mkdir -p gpurun_out
for ctas in 16 6; do
LRB_XCHG_CTAS=$ctas timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_$ctas.json 2> gpurun_out/bench_n2_$ctas.err; echo "bench n2 rc=$?"
tail -2 gpurun_out/bench_n2_$ctas.err | cut -c1-300
python -c "
import json
for ln in open('gpurun_out/bench_n2_$ctas.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('ctas $ctas value',d['value'],'ms',d['ms_per_step'],'plan',d['config']['plan'],d['config']['plan_ms']); print(d['phases_ms_rank0'])"
done
