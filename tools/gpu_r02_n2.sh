# round 2, 2 GPUs: every multi-GPU plan row-for-row vs the oracle, the multi-GPU context through the boundary, bench --gpus 2
# (gpurun --gpus 2 -- 'bash tools/gpu_r02_n2.sh')
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box_n2.txt; nvidia-smi topo -m >> gpurun_out/box_n2.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_gpu_dist_n2.log 2>&1; echo "pytest dist rc=$?" | tee -a gpurun_out/pytest_gpu_dist_n2.log
tail -5 gpurun_out/pytest_gpu_dist_n2.log; tail -3 gpurun_out/gpu_dist_world2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_gpu_context or streamed or contigs" > gpurun_out/pytest_multi_ctx_n2.log 2>&1; echo "pytest ctx rc=$?" | tee -a gpurun_out/pytest_multi_ctx_n2.log
tail -8 gpurun_out/pytest_multi_ctx_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
