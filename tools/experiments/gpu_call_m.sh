#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02m}
run() { name=$1; shift
env "$@" python bench.py --no-cpu-baseline --no-cfg4 --repeats 3 > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err
python - $out/${tag}_bench_$name.json $name <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d['roofline']['breakdown_ms']
print(sys.argv[2].ljust(14), 'gen/s %.0f step %.1f us wall %.1f core %.1f select %.1f acc %.1f' % (d['value'], 1e3*d['ms_per_step'], 1e3*d['wall_ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step']))
PY
}
run prio0 PANSIM_GRAPH_PRIO=0 PANSIM_GRAPH_DEBUG=1
grep "\[graph\]" $out/${tag}_bench_prio0.err | sort | uniq -c | sort -rn | head -20
run prio1 PANSIM_GRAPH_PRIO=1
run prio2 PANSIM_GRAPH_PRIO=2
run conn32_prio1 CUDA_DEVICE_MAX_CONNECTIONS=32 PANSIM_GRAPH_PRIO=1
run conn32_nograph CUDA_DEVICE_MAX_CONNECTIONS=32 PANSIM_GRAPH=0
