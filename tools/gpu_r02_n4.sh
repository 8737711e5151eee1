# round 2, 4 GPUs: plans vs the oracle at world 4, 3- and 4-GPU contexts, bench --gpus 4 (gpurun --gpus 4 -- 'bash tools/gpu_r02_n4.sh')
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -k "4" > gpurun_out/pytest_gpu_dist_n4.log 2>&1; echo "pytest dist rc=$?" | tee -a gpurun_out/pytest_gpu_dist_n4.log
tail -3 gpurun_out/pytest_gpu_dist_n4.log; tail -1 gpurun_out/gpu_dist_world4.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_gpu_context and (3 or 4)" > gpurun_out/pytest_multi_ctx_n4.log 2>&1; echo "pytest ctx rc=$?" | tee -a gpurun_out/pytest_multi_ctx_n4.log
tail -3 gpurun_out/pytest_multi_ctx_n4.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench n4 rc=$?"
python tools/bench_summary.py gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
