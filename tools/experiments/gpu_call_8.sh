#!/bin/bash
# 8-GPU call: group tests, bench --gpus 8, cfg5 all pairs on 8 GPUs
set -u
out=gpurun_out
tag=${1:-r02n8}
python -m pytest tests/test_gpu_multi.py -x -q > $out/${tag}_multi_tests.txt 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_multi_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 > $out/${tag}_bench_8gpu.json 2> $out/${tag}_bench_8gpu.err
echo "bench8 rc=$?"; tail -3 $out/${tag}_bench_8gpu.err
python tools/all_pairs_bench.py --pop_size 20000 --gpus 8 > $out/${tag}_cfg5_8gpu.json 2> $out/${tag}_cfg5_8gpu.err
echo "cfg5 rc=$?"; cat $out/${tag}_cfg5_8gpu.json; tail -3 $out/${tag}_cfg5_8gpu.err
