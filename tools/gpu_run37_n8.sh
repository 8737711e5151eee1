mkdir -p gpurun_out
N=${1:-8}
free -g | head -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_n$N.err | tail -5 | cut -c1-400
python -c "
import json
for ln in open('gpurun_out/bench_n$N.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('value',d['value'],'ms',d['ms_per_step'],'plan',d['config']['plan'],d['config']['plan_ms']); print('e2e',d['e2e']); print(d['phases_ms_rank0'])"
