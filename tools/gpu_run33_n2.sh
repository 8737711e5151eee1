mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -5 gpurun_out/bench_n2.err
python -c "
import json
for ln in open('gpurun_out/bench_n2.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('value',d['value'],'ms',d['ms_per_step'],'plan',d['config']['plan'],d['config']['plan_ms']); print('e2e',d['e2e']); print(d['phases_ms_rank0'])"
