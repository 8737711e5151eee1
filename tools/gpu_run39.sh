mkdir -p gpurun_out
N=${1:-4}
nvidia-smi topo -m | head -12 | cut -c1-160; numactl -H 2>/dev/null | head -4; nproc
LRB_FEED_DEBUG=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
grep "feed timeline" gpurun_out/bench_n$N.err | tail -1 | cut -c1-700
python -c "
import json
for ln in open('gpurun_out/bench_n$N.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('value',d['value'],'ms',d['ms_per_step'],'plan',d['config']['plan']); print('e2e',d['e2e'])"
