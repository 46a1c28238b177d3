#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02r}
run() { name=$1; shift
env "$@" python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err || tail -2 $out/${tag}_bench_$name.err
python - $out/${tag}_bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); b=d['roofline']['breakdown_ms']
    print(sys.argv[2].ljust(14), 'gen/s %.0f step %.1f us wall %.1f core %.1f select %.1f acc %.1f frac %.3f step_frac %.3f e2e %.0f' % (d['value'], 1e3*d['ms_per_step'], 1e3*d['wall_ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e']['value']))
except Exception as e: print(sys.argv[2], 'ERR', e)
PY
}
run base A=1
run pdl2 PANSIM_PDL=2
run pdl2_nograph PANSIM_PDL=2 PANSIM_GRAPH=0
run fit1 PANSIM_FITNESS_MODE=1
run popc PANSIM_INTER_POPC=1
run ipb5 PANSIM_CORE_ITEMS_BATCH=5
