mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2_partition -s 5 -c 1 -f -o gpurun_out/r01q_k2_partition_chunk python tools/prof_chunks.py 16 > gpurun_out/ncu_k2c.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_k2c.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k2_partition --csv --log-file gpurun_out/launches_k2c.csv python tools/prof_chunks.py 16 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/launches_k2c.csv; grep k2_part gpurun_out/launches_k2c.csv | awk -F'","' '{print $NF}' | tr '\n' ' '
