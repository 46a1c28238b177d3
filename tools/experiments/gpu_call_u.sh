#!/bin/bash
set -u
out=gpurun_out
tag=r02u
run() { name=$1; shift
env "$@" python bench.py --no-cpu-baseline --no-cfg4 --repeats 3 > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err || tail -2 $out/${tag}_bench_$name.err
python - $out/${tag}_bench_$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); b=d['roofline']['breakdown_ms']
    print(sys.argv[2].ljust(16), 'gen/s %.0f step %.1f us core %.1f select %.1f acc %.1f frac %.3f step_frac %.3f e2e %.0f' % (d['value'], 1e3*d['ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e']['value']))
except Exception as e: print(sys.argv[2], 'ERR', e)
PY
}
L=$PWD/pansim_b200/variants/lib_s2.so
run base A=1
run s2 PANSIM_B200_LIB=$L
run s2_3cta PANSIM_B200_LIB=$L PANSIM_CORE_SMEM_PAD_KB=20
run s2_3cta_ipb8 PANSIM_B200_LIB=$L PANSIM_CORE_SMEM_PAD_KB=20 PANSIM_CORE_ITEMS_BATCH=8
run s2_3cta_ipb12 PANSIM_B200_LIB=$L PANSIM_CORE_SMEM_PAD_KB=20 PANSIM_CORE_ITEMS_BATCH=12
run s3_3cta PANSIM_CORE_SMEM_PAD_KB=16
