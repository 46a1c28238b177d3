#!/bin/bash
# Run on the GPU box (gpurun): final bench line, ncu launch list and ncu --set full summaries -> gpurun_out/
set -u
out=gpurun_out
tag=${1:-r01}
python bench.py > $out/${tag}_bench_final.json 2> $out/${tag}_bench_final.err
python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps5.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py $out/${tag}_launches_bench_steps5.csv > $out/${tag}_launches_summary.txt
: > $out/${tag}_ncu_summary.txt
# launch 7 of core_mut_kernel (0-based) is the first with recombination events pending after a materialised state
for spec in "core_mut_kernel 7" "hr_collect_kernel 1" "hr_apply_kernel 1" "acc_inter_mma_kernel 5" "fitness_lane_kernel 3" "fitness_kernel 1" \
            "avg_distance_kernel 5" "select_parents_small_kernel 5" "acc_gather_flip_kernel 5" "acc_gain_threshold_kernel 5" \
            "acc_hgt_apply_kernel 5" "pair_core_grouped_kernel 2"; do
    set -- $spec
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 -c 1 -f \
        -o $out/${tag}_prof_$1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/ncu_$1.log 2>&1
    python profiles/ncu_summary.py $out/${tag}_prof_$1.ncu-rep >> $out/${tag}_ncu_summary.txt 2>&1
done
