# round 2 experiments, batch 2 (one B200)
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/exp2_default.json 2> gpurun_out/exp2_default.err; echo "default rc=$?"
LRB_SEARCH_HINT=1 $B > gpurun_out/exp2_search_hint.json 2>/dev/null; echo "hint rc=$?"
LRB_H2D_CHUNKS=8 $B > gpurun_out/exp2_chunks8.json 2>/dev/null; echo "chunks8 rc=$?"
LRB_H2D_CHUNKS=32 $B > gpurun_out/exp2_chunks32.json 2>/dev/null; echo "chunks32 rc=$?"
python tools/bench_summary.py gpurun_out/exp2_default.json gpurun_out/exp2_search_hint.json gpurun_out/exp2_chunks8.json gpurun_out/exp2_chunks32.json
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_search_keys -c 2 --csv --log-file gpurun_out/ncu_search_nohint.csv python tools/prof_step.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu nohint rc=$?"
LRB_SEARCH_HINT=1 timeout 600 ncu --metrics $M --clock-control none -k regex:k_search_keys -c 2 --csv --log-file gpurun_out/ncu_search_hint.csv python tools/prof_step.py --reads 1000000 --steps 1 > /dev/null 2>&1; echo "ncu hint rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:k_partition -c 2 --csv --log-file gpurun_out/ncu_part_nobulk.csv python tools/prof_step.py --reads 200000 --steps 1 > /dev/null 2>&1; echo "ncu nobulk rc=$?"
LRB_PART_BULK=1 timeout 600 ncu --metrics $M --clock-control none -k regex:k_partition -c 2 --csv --log-file gpurun_out/ncu_part_bulk.csv python tools/prof_step.py --reads 200000 --steps 1 > /dev/null 2>&1; echo "ncu bulk rc=$?"
grep -h "k_search_keys\|k_partition" gpurun_out/ncu_search_nohint.csv gpurun_out/ncu_search_hint.csv gpurun_out/ncu_part_nobulk.csv gpurun_out/ncu_part_bulk.csv | awk -F'","' '{print FILENAME, $5, $(NF-2), $NF}' | cut -c1-220
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_traffic_step.csv python tools/prof_step.py --reads 1000000 --steps 1 > gpurun_out/traffic_step.log 2>&1; echo "traffic rc=$?"
python tools/traffic_from_ncu.py gpurun_out/r02_traffic_step.csv cfg2_1M_5kb_ont_k4 1000000 gpurun_out/traffic.json
timeout 900 python tools/exp_skew.py > gpurun_out/r02_exp_skew.jsonl 2> gpurun_out/exp_skew.err; echo "skew rc=$?"; cat gpurun_out/r02_exp_skew.jsonl; tail -3 gpurun_out/exp_skew.err
