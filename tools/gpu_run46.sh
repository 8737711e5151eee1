mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "not full_size and not synthetic_reads_bit_exact" > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
