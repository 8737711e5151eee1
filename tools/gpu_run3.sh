mkdir -p gpurun_out
./build/ubench_l2fetch > gpurun_out/ubench_l2fetch.jsonl 2>&1
cat gpurun_out/ubench_l2fetch.jsonl
timeout 600 python tools/exp_partition.py 1000000 > gpurun_out/exp_partition.jsonl 2> gpurun_out/exp_partition.err; tail -3 gpurun_out/exp_partition.err
cat gpurun_out/exp_partition.jsonl
LRB_L2_FETCH=32 timeout 600 python tools/exp_partition.py 1000000 > gpurun_out/exp_partition_l2f32.jsonl 2>> gpurun_out/exp_partition.err
cat gpurun_out/exp_partition_l2f32.jsonl
timeout 600 python tools/exp_multipass.py 400000 > gpurun_out/exp_multipass.jsonl 2> gpurun_out/exp_multipass.err; tail -3 gpurun_out/exp_multipass.err
cat gpurun_out/exp_multipass.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "partitioned" > gpurun_out/pytest_part.log 2>&1; tail -15 gpurun_out/pytest_part.log
