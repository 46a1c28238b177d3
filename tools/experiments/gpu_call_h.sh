#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02h}
python bench.py --no-cpu-baseline --no-cfg4 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
python tools/dist_bench.py > $out/${tag}_dist_plain.log 2>&1 &&
REPS=2 ncu --metrics $M --clock-control none -k regex:'pair_tile2_kernel|core_planes_kernel|pair_planes_grouped_kernel|pair_acc_kernel' -c 12 --csv --log-file $out/${tag}_l2_pair.csv python tools/dist_bench.py > $out/${tag}_ncu_pair.log 2>&1
cat $out/${tag}_dist_plain.log
