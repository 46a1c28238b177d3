#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02l}
run() { name=$1; shift
env "$@" python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_bench_$name.json 2> $out/${tag}_bench.err
python - $out/${tag}_bench_$name.json $name <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d['roofline']['breakdown_ms']
print(sys.argv[2].ljust(14), 'gen/s %.0f step %.1f us wall %.1f core %.1f select %.1f acc %.1f' % (d['value'], 1e3*d['ms_per_step'], 1e3*d['wall_ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step']))
PY
}
run nospans PANSIM_GRAPH_SPANS=0
run nospans_ipb3 PANSIM_GRAPH_SPANS=0 PANSIM_CORE_ITEMS_BATCH=3
run nospans_ipb8 PANSIM_GRAPH_SPANS=0 PANSIM_CORE_ITEMS_BATCH=8
run nograph PANSIM_GRAPH=0
