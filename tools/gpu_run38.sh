mkdir -p gpurun_out
N=${1:-4}
LRB_FEED_DEBUG=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
grep "feed timeline" gpurun_out/bench_n$N.err | tail -3 | cut -c1-1500
