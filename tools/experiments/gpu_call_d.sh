#!/bin/bash
set -u
out=gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --no-cpu-baseline --steps 200 > $out/r02d_$name.json 2>> $out/r02d.err
  python - $out/r02d_$name.json $name <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
b=d['roofline']['breakdown_ms']
print(sys.argv[2].ljust(12), 'gen/s %.0f  step %.1f us  core %.1f  select %.1f  acc %.1f  e2e %.0f' % (d['value'], 1e3*d['ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step'], d['e2e']['value']))
PY
}
run base A=1
run ipb3 PANSIM_CORE_ITEMS_BATCH=3
run ipb4 PANSIM_CORE_ITEMS_BATCH=4
run ipb6 PANSIM_CORE_ITEMS_BATCH=6
run ipb12 PANSIM_CORE_ITEMS_BATCH=12
for v in s2 r56s2 r56s3 w7s2; do
  run $v PANSIM_B200_LIB=$PWD/pansim_b200/variants/lib_$v.so
  run ${v}_ipb4 PANSIM_B200_LIB=$PWD/pansim_b200/variants/lib_$v.so PANSIM_CORE_ITEMS_BATCH=4
done
