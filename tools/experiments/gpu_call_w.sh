#!/bin/bash
# second pass over the selection-chain kernels: parity, FP64 latency micro-benchmark, step timing per variant, isolated kernel times
set -u
out=gpurun_out
tag=r02w
mkdir -p $out
./tools/experiments/fp64_latency > $out/${tag}_fp64_latency.txt 2>&1; cat $out/${tag}_fp64_latency.txt
timeout -k 10 400 python tests/manual/check_chain_variants.py 012 > $out/${tag}_check.txt 2>&1
echo "check rc=$?"; tail -2 $out/${tag}_check.txt; grep -c "^ok" $out/${tag}_check.txt; grep "MISMATCH\|Error\|error" $out/${tag}_check.txt | head -20
run() { name=$1; shift
env BENCH_DIAG=1 BENCH_DIAG_NAME=$name "$@" timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_diag_$name.json 2> $out/${tag}_diag_$name.err || tail -2 $out/${tag}_diag_$name.err
cat $out/${tag}_diag_$name.json | cut -c1-330
}
run base PANSIM_INTER_UMMA=0 PANSIM_AVG_RCP=0
run new PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3
run new_fit2 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1
run umma1 PANSIM_INTER_UMMA=1 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3
run mma_rcp PANSIM_INTER_UMMA=0 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3
run new_ipb5 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3 PANSIM_CORE_ITEMS_BATCH=5
run new_ipb6 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3 PANSIM_CORE_ITEMS_BATCH=6
run new_ipb8 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3 PANSIM_CORE_ITEMS_BATCH=8
run new_ipb3 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3 PANSIM_CORE_ITEMS_BATCH=3
env BENCH_DIAG=1 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=3 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
  -k regex:"umma|expand|avg_distance|fitness|select_parents|acc_" -c 60 --csv --log-file $out/${tag}_chain_kernels.csv \
  python bench.py --no-cpu-baseline --no-cfg4 --repeats 1 --steps 8 --warmup 3 > $out/${tag}_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02w_chain_kernels.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
acc=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    acc[r[ik][:60]][r[im]].append(float(r[iv].replace(',','')))
for k,v in acc.items():
    print(k.ljust(62), ' '.join(f"{m.split('.')[0][-14:]}={sum(x)/len(x):.1f}" for m,x in v.items()), 'n=%d'%len(list(v.values())[0]))
PY
