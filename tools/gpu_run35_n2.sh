mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/probe_symm.py > gpurun_out/probe_symm.log 2>&1; echo "rc=$?"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/probe_symm.log | tail -25
nvidia-smi topo -m | head -8
