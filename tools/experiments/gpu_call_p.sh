#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02p}
python -m pytest tests/test_gpu_parity.py -x -q -k "competition or fitness or average or hgt or generate or graph or run_generations or shards" > $out/${tag}_tests.txt 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_tests.txt
run() { name=$1; shift
env "$@" python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err
python - $out/${tag}_bench_$name.json $name <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d['roofline']['breakdown_ms']
print(sys.argv[2].ljust(14), 'gen/s %.0f step %.1f us wall %.1f core %.1f select %.1f acc %.1f frac %.3f step_frac %.3f e2e %.0f' % (d['value'], 1e3*d['ms_per_step'], 1e3*d['wall_ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e']['value']))
PY
}
run base A=1
run hole PANSIM_B200_LIB=$PWD/pansim_b200/variants/lib_hole.so
run hole_ipb8 PANSIM_B200_LIB=$PWD/pansim_b200/variants/lib_hole.so PANSIM_CORE_ITEMS_BATCH=8
run hole_ipb12 PANSIM_B200_LIB=$PWD/pansim_b200/variants/lib_hole.so PANSIM_CORE_ITEMS_BATCH=12
run hole_ipb3 PANSIM_B200_LIB=$PWD/pansim_b200/variants/lib_hole.so PANSIM_CORE_ITEMS_BATCH=3
