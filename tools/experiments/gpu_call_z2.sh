#!/bin/bash
set -u
out=gpurun_out
tag=r02z2
run() { name=$1; shift
env BENCH_DIAG=1 BENCH_DIAG_NAME=$name "$@" timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --repeats 7 > $out/${tag}_diag_$name.json 2> $out/${tag}_diag_$name.err || tail -2 $out/${tag}_diag_$name.err
cat $out/${tag}_diag_$name.json | cut -c1-330
}
for ipb in 4 5 6 7 8 10 12; do run rcp1_ipb$ipb PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_CORE_ITEMS_BATCH=$ipb; done
for ipb in 5 6 8; do run rcp0_ipb$ipb PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=0 PANSIM_CORE_ITEMS_BATCH=$ipb; done
run rcp1_ipb5_fit1 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_CORE_ITEMS_BATCH=5 PANSIM_FITNESS_MODE=1
run rcp1_ipb5_u3 PANSIM_INTER_UMMA=3 PANSIM_AVG_RCP=1 PANSIM_CORE_ITEMS_BATCH=5
run rcp1_ipb5_b PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_CORE_ITEMS_BATCH=5
