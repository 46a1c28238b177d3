#!/bin/bash
set -u
out=gpurun_out
tag=r02z
run() { name=$1; shift
env BENCH_DIAG=1 BENCH_DIAG_NAME=$name "$@" timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_diag_$name.json 2> $out/${tag}_diag_$name.err || tail -2 $out/${tag}_diag_$name.err
cat $out/${tag}_diag_$name.json | cut -c1-330
}
run u2_rcp0 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=0
run u2_rcp1 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1
run u2_rcp0_ipb3 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=0 PANSIM_CORE_ITEMS_BATCH=3
run u2_rcp0_ipb5 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=0 PANSIM_CORE_ITEMS_BATCH=5
run u2_rcp0_b PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=0
