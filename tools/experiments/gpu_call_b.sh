#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02b}
python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"
python tools/run_config.py --gens 3 --warm 2 --pairs 1000 > $out/${tag}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:core_mut_kernel -s 3 -c 1 -f -o $out/${tag}_prof_core_mut \
    python tools/run_config.py --gens 3 --warm 2 --pairs 1000 > $out/${tag}_ncu.log 2>&1
cat $out/${tag}_plain.log
