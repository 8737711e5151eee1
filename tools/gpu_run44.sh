mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config3" --durations=2 > gpurun_out/pytest_cfg3.log 2>&1; tail -5 gpurun_out/pytest_cfg3.log
nvidia-smi --query-gpu=memory.used --format=csv,noheader
