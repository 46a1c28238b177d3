#!/bin/bash
set -u
out=gpurun_out
tag=r02z3
run() { name=$1; shift
env BENCH_DIAG=1 BENCH_DIAG_NAME=$name PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=0 "$@" timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --repeats 7 > $out/${tag}_diag_$name.json 2> $out/${tag}_diag_$name.err || tail -2 $out/${tag}_diag_$name.err
cat $out/${tag}_diag_$name.json | cut -c1-330
}
L=$PWD/pansim_b200/variants/lib_s2.so
run s3_ipb5 PANSIM_CORE_ITEMS_BATCH=5
for ipb in 4 5 6 7; do run s2_ipb$ipb PANSIM_B200_LIB=$L PANSIM_CORE_ITEMS_BATCH=$ipb; done
run s3_ipb5_b PANSIM_CORE_ITEMS_BATCH=5
run s3_ipb5_nograph PANSIM_CORE_ITEMS_BATCH=5 PANSIM_GRAPH=0
