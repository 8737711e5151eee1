mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_partition -s 1 -c 1 -f -o gpurun_out/r01n_k2_partition python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/ncu_k2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_k2.log
