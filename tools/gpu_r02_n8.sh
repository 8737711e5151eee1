# round 2, 8 GPUs: plans row-for-row vs the oracle at world 8, the 8-GPU context through the boundary, bench --gpus 8 with the
# strong-scaled north-star configs (gpurun --gpus 8 -- 'bash tools/gpu_r02_n8.sh')
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box_n8.txt; nvidia-smi topo -m >> gpurun_out/box_n8.txt 2>&1; nproc >> gpurun_out/box_n8.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k "8" > gpurun_out/pytest_gpu_dist_n8.log 2>&1; echo "pytest dist rc=$?" | tee -a gpurun_out/pytest_gpu_dist_n8.log
tail -4 gpurun_out/pytest_gpu_dist_n8.log; tail -2 gpurun_out/gpu_dist_world8.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_gpu_context and 8" > gpurun_out/pytest_multi_ctx_n8.log 2>&1; echo "pytest ctx rc=$?" | tee -a gpurun_out/pytest_multi_ctx_n8.log
tail -4 gpurun_out/pytest_multi_ctx_n8.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
python tools/bench_summary.py gpurun_out/bench_n8.json; tail -5 gpurun_out/bench_n8.err
