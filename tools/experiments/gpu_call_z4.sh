#!/bin/bash
set -u
out=gpurun_out
tag=r02z4
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "variants" 2>&1 | tail -2
run() { name=$1; shift
env BENCH_DIAG=1 BENCH_DIAG_NAME=$name "$@" timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --repeats 7 > $out/${tag}_diag_$name.json 2> $out/${tag}_diag_$name.err || tail -2 $out/${tag}_diag_$name.err
cat $out/${tag}_diag_$name.json | cut -c1-330
}
run rcp0 PANSIM_AVG_RCP=0
run rcp2 PANSIM_AVG_RCP=2
run rcp2_ipb4 PANSIM_AVG_RCP=2 PANSIM_CORE_ITEMS_BATCH=4
run rcp2_ipb6 PANSIM_AVG_RCP=2 PANSIM_CORE_ITEMS_BATCH=6
run rcp2_b PANSIM_AVG_RCP=2
run rcp0_b PANSIM_AVG_RCP=0
