#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02e}
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.txt 2>&1
echo "pytest rc=$?"; tail -15 $out/${tag}_tests.txt
python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc=$?"; tail -3 $out/${tag}_bench.err
