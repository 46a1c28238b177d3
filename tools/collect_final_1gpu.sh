#!/bin/bash
# final single-GPU collection of the round: tests, bench lines, the other BASELINE configs, DRAM counters of the pair kernels
set -u
out=gpurun_out
tag=r02
python -m pytest tests -m gpu -q > $out/${tag}_gpu_tests.txt 2>&1; echo "pytest rc=$?" >> $out/${tag}_gpu_tests.txt; tail -3 $out/${tag}_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.txt; tail -2 $out/${tag}_smoke.txt
python bench.py > $out/${tag}_bench_final.json 2> $out/${tag}_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 0 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench_final.err
python tools/run_config.py --gens 50 --warm 5 --pairs 100000 > $out/${tag}_cfg1.json 2> $out/${tag}_cfg1.err; cat $out/${tag}_cfg1.json
python tools/run_config.py --HR_rate 1.0 --HGT_rate 1.0 --rate_genes2 1000 --gens 30 --warm 5 --pairs 100000 > $out/${tag}_cfg3.json 2> $out/${tag}_cfg3.err; cat $out/${tag}_cfg3.json
python tools/all_pairs_bench.py --pop_size 20000 --gpus 1 > $out/${tag}_cfg5_1gpu.json 2> $out/${tag}_cfg5_1gpu.err; cat $out/${tag}_cfg5_1gpu.json
B="python bench.py --steps 8 --warmup 3 --repeats 2 --no-cpu-baseline --no-cfg4"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum
$B > /dev/null 2>&1 &&
ncu --metrics $M --clock-control none -k regex:'pair_tile2_kernel|core_planes_kernel|pair_acc_kernel' -s 3 -c 6 --csv --log-file $out/${tag}_l2_pair.csv $B > /dev/null 2>&1
echo collected
