#!/bin/bash
# third pass over the selection chain: parity of every variant, the GPU test suite, step timing per variant, isolated kernel times
set -u
out=gpurun_out
tag=r02x
mkdir -p $out
timeout -k 10 400 python tests/manual/check_chain_variants.py 012 > $out/${tag}_check.txt 2>&1
echo "check rc=$?"; tail -2 $out/${tag}_check.txt; grep -c "^ok" $out/${tag}_check.txt; grep "MISMATCH\|Error\|error" $out/${tag}_check.txt | head -20
timeout -k 10 900 python -m pytest tests -x -q -m gpu > $out/${tag}_gpu_tests.txt 2>&1; echo "pytest rc=$?"; tail -5 $out/${tag}_gpu_tests.txt
run() { name=$1; shift
env BENCH_DIAG=1 BENCH_DIAG_NAME=$name "$@" timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_diag_$name.json 2> $out/${tag}_diag_$name.err || tail -2 $out/${tag}_diag_$name.err
cat $out/${tag}_diag_$name.json | cut -c1-330
}
run old PANSIM_INTER_UMMA=0 PANSIM_AVG_RCP=0
run new2 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1
run new1 PANSIM_INTER_UMMA=1 PANSIM_AVG_RCP=1
run new0 PANSIM_INTER_UMMA=0 PANSIM_AVG_RCP=1
run umma2_rcp0 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=0
run new2_fit1 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FITNESS_MODE=1
run new2_ipb3 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_CORE_ITEMS_BATCH=3
run new2_ipb5 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_CORE_ITEMS_BATCH=5
run new2_ipb6 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_CORE_ITEMS_BATCH=6
run old_fine PANSIM_INTER_UMMA=0 PANSIM_AVG_RCP=0 PANSIM_FINE_TIMING=1
grep "fine timing" $out/${tag}_diag_old_fine.err | tail -2
run new2_fine PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 PANSIM_FINE_TIMING=1
grep "fine timing" $out/${tag}_diag_new2_fine.err | tail -2
run new1_fine PANSIM_INTER_UMMA=1 PANSIM_AVG_RCP=1 PANSIM_FINE_TIMING=1
grep "fine timing" $out/${tag}_diag_new1_fine.err | tail -2
env BENCH_DIAG=1 PANSIM_INTER_UMMA=2 PANSIM_AVG_RCP=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
  -k regex:"umma|expand|avg_distance|fitness|select_parents|acc_" -c 60 --csv --log-file $out/${tag}_chain_kernels.csv \
  python bench.py --no-cpu-baseline --no-cfg4 --repeats 1 --steps 8 --warmup 3 > $out/${tag}_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02x_chain_kernels.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
acc=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    acc[r[ik][:60]][r[im]].append(float(r[iv].replace(',','')))
for k,v in acc.items():
    print(k.ljust(62), ' '.join(f"{m.split('.')[0][-14:]}={sum(x)/len(x):.1f}" for m,x in v.items()), 'n=%d'%len(list(v.values())[0]))
PY
