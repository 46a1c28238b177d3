#!/bin/bash
# Last collection of round 2 (one B200): GPU tests, smoke, the bench line and the reference arm, the ncu launch list of the
# bench command, ncu --set full of the generation-step kernels, DRAM counters of core_mut_kernel and the pair kernels.
set -u
out=gpurun_out
tag=r02
python -m pytest tests -m gpu -q > $out/${tag}_gpu_tests.txt 2>&1; echo "pytest rc=$?" >> $out/${tag}_gpu_tests.txt; tail -3 $out/${tag}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.txt; tail -2 $out/${tag}_smoke.txt
python bench.py > $out/${tag}_bench_final.json 2> $out/${tag}_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench_final.err
B="python bench.py --steps 8 --warmup 3 --repeats 2 --no-cpu-baseline --no-cfg4"
$B > $out/${tag}_plain_short.json 2> $out/${tag}_plain_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench_steps8.csv $B > /dev/null 2>&1
python tools/launch_summary.py $out/${tag}_launches_bench_steps8.csv > $out/${tag}_launches_summary.txt
: > $out/${tag}_ncu_summary.txt
for spec in "core_mut_kernel 7" "acc_inter_umma 5" "fitness_lane_kernel 3" "avg_distance_kernel 5" "select_parents_small_kernel 5" \
            "acc_gather_flip_kernel 5" "acc_gain_threshold_kernel 5" "acc_hgt_apply_kernel 5" "pair_tile2_kernel 1"; do
    set -- $spec
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 -c 1 -f \
        -o $out/${tag}_prof_$1 $B > $out/ncu_$1.log 2>&1
    python profiles/ncu_summary.py $out/${tag}_prof_$1.ncu-rep >> $out/${tag}_ncu_summary.txt 2>&1
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_ltcfabric.sum,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:core_mut_kernel -s 4 -c 3 --csv --log-file $out/${tag}_l2_core_mut_cfg2_selection.csv $B > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:'pair_tile2_kernel|core_planes_kernel|pair_acc_kernel' -s 3 -c 6 --csv --log-file $out/${tag}_l2_pair.csv $B > /dev/null 2>&1
echo collected
