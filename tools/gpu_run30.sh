mkdir -p gpurun_out
timeout 600 python tools/exp_chunks.py > gpurun_out/exp_chunks.jsonl 2> gpurun_out/exp_chunks.err; cat gpurun_out/exp_chunks.jsonl; tail -3 gpurun_out/exp_chunks.err
