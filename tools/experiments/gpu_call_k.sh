#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02k}
python -m pytest tests/test_gpu_parity.py -x -q -k "graph or run_generations" > $out/${tag}_tests.txt 2>&1
echo "pytest rc=$?"; tail -3 $out/${tag}_tests.txt
for ipb in 4 3 6 8; do
PANSIM_CORE_ITEMS_BATCH=$ipb python bench.py --no-cpu-baseline --no-cfg4 --repeats 5 > $out/${tag}_bench_ipb$ipb.json 2> $out/${tag}_bench.err
python - $out/${tag}_bench_ipb$ipb.json $ipb <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); b=d['roofline']['breakdown_ms']
print('ipb', sys.argv[2], 'gen/s %.0f step %.1f us wall %.1f core %.1f select %.1f acc %.1f frac %.3f step_frac %.3f' % (d['value'], 1e3*d['ms_per_step'], 1e3*d['wall_ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step'], d['roofline']['frac'], d['roofline']['step_frac']))
PY
done
