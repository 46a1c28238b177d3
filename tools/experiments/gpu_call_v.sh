#!/bin/bash
# chain-kernel variants: parity against the oracle, then generation-step timing per variant, then isolated kernel times
set -u
out=gpurun_out
tag=r02v
mkdir -p $out
timeout -k 10 300 python tests/manual/check_chain_variants.py 0 > $out/${tag}_check0.txt 2>&1
echo "check0 rc=$?"; tail -3 $out/${tag}_check0.txt; grep -c "^ok" $out/${tag}_check0.txt; grep "MISMATCH" $out/${tag}_check0.txt | head -20
timeout -k 10 300 python tests/manual/check_chain_variants.py 1 > $out/${tag}_check1.txt 2>&1
rc1=$?
echo "check1 rc=$rc1"; tail -3 $out/${tag}_check1.txt; grep -c "^ok" $out/${tag}_check1.txt; grep "MISMATCH" $out/${tag}_check1.txt | head -20
run() { name=$1; shift
env BENCH_DIAG=1 BENCH_DIAG_NAME=$name "$@" timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_diag_$name.json 2> $out/${tag}_diag_$name.err || tail -2 $out/${tag}_diag_$name.err
cat $out/${tag}_diag_$name.json | cut -c1-400
}
run base PANSIM_INTER_UMMA=0 PANSIM_AVG_LANE=0
if [ $rc1 -ne 0 ] && ! grep -q "umma=1" $out/${tag}_check1.txt; then echo "tcgen05 kernel did not run: skipping its timings"; SKIP_UMMA=1; else SKIP_UMMA=0; fi
[ $SKIP_UMMA = 1 ] || run umma PANSIM_INTER_UMMA=1 PANSIM_AVG_LANE=0
run lane PANSIM_INTER_UMMA=0 PANSIM_AVG_LANE=1
[ $SKIP_UMMA = 1 ] || run umma_lane PANSIM_INTER_UMMA=1 PANSIM_AVG_LANE=1
[ $SKIP_UMMA = 1 ] || run umma_lane_walk PANSIM_INTER_UMMA=1 PANSIM_AVG_LANE=1 PANSIM_FITNESS_MODE=3
run lane_walk PANSIM_INTER_UMMA=0 PANSIM_AVG_LANE=1 PANSIM_FITNESS_MODE=3
run base2 PANSIM_INTER_UMMA=0 PANSIM_AVG_LANE=0
U=1; [ $SKIP_UMMA = 1 ] && U=0
env BENCH_DIAG=1 PANSIM_INTER_UMMA=$U PANSIM_AVG_LANE=1 PANSIM_FITNESS_MODE=3 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
  -k regex:"umma|expand|avg_distance|fitness|select_parents|acc_" -c 60 --csv --log-file $out/${tag}_chain_kernels.csv \
  python bench.py --no-cpu-baseline --no-cfg4 --repeats 1 --steps 8 --warmup 3 > $out/${tag}_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02v_chain_kernels.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
acc=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    acc[r[ik][:60]][r[im]].append(float(r[iv].replace(',','')))
for k,v in acc.items():
    print(k.ljust(62), ' '.join(f"{m.split('.')[0][-14:]}={sum(x)/len(x):.1f}" for m,x in v.items()), 'n=%d'%len(list(v.values())[0]))
PY
