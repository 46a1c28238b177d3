#!/bin/bash
# round-2 call A: full GPU test suite, default bench, L2/DRAM counters of the core kernel + copy control
set -u
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/r02a_gpu_tests.txt 2>&1
echo "pytest rc=$?" >> $out/r02a_gpu_tests.txt
tail -5 $out/r02a_gpu_tests.txt
python bench.py --no-cpu-baseline > $out/r02a_bench.json 2> $out/r02a_bench.err
echo "bench rc=$?"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__sectors_read.sum,dram__sectors_write.sum,lts__t_bytes.sum,lts__t_sectors.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_lookup_hit.sum,lts__t_sectors_lookup_miss.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcnode_gpc.sum,lts__t_sectors_srcnode_fbp.sum,lts__t_sectors_srcnode_hub.sum,smsp__inst_executed.sum
python tools/run_config.py --gens 3 --warm 2 --pairs 1000 > $out/r02a_plain_rng.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:core_mut_kernel -s 2 -c 2 --csv --log-file $out/r02a_l2_core_mut_rng.csv \
    python tools/run_config.py --gens 3 --warm 2 --pairs 1000 > $out/r02a_ncu_rng.log 2>&1
python tools/run_config.py --gens 3 --warm 2 --pairs 1000 --core_mu 0 --HR_rate 0 > $out/r02a_plain_copy.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:core_mut_kernel -s 2 -c 2 --csv --log-file $out/r02a_l2_core_mut_copy.csv \
    python tools/run_config.py --gens 3 --warm 2 --pairs 1000 --core_mu 0 --HR_rate 0 > $out/r02a_ncu_copy.log 2>&1
ncu --metrics $M --clock-control none --cache-control none -k regex:core_mut_kernel -s 2 -c 2 --csv --log-file $out/r02a_l2_core_mut_rng_nocachectl.csv \
    python tools/run_config.py --gens 3 --warm 2 --pairs 1000 > $out/r02a_ncu_rng2.log 2>&1
echo done
