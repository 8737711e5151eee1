timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "building_blocks or partitioned" 2>&1 | tail -5
