# multi-GPU e2e experiment: one vs two device buffers for the packed stream (gpurun --gpus N -- 'bash tools/gpu_r02_e2e.sh N')
N=${1:-2}
mkdir -p gpurun_out
for B in 2 1; do
  LRB_E2E_BUFFERS=$B timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-north-star > gpurun_out/bench_n${N}_buf$B.json 2> gpurun_out/bench_n${N}_buf$B.err; echo "buffers=$B rc=$?"
  python tools/bench_summary.py gpurun_out/bench_n${N}_buf$B.json
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n${N}_buf$B.json").read().splitlines() if l.startswith("{")][-1])
e=d["e2e"]; print("  e2e", round(e["value"],1), "Gb/s", round(e["ms_per_step"],2), "ms; h2d copy", round(e["rank0_ms_since_step_start"]["h2d_copy_ms"],1), "ms; ceiling", round(e["box_h2d_ceiling_GBps_all_ranks_copying"],1), "GB/s; verify", d["verify"]["ok"])
PY
done
