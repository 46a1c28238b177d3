#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02c}
python -m pytest tests/test_gpu_parity.py -x -q -k "deferred or dump or benchmark or recomb or generate or shards or run_generations or entry_point" > $out/${tag}_tests.txt 2>&1
tail -3 $out/${tag}_tests.txt
python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"
python bench.py --no-cpu-baseline > $out/${tag}_bench2.json 2>> $out/${tag}_bench.err
