#!/bin/bash
# 2-GPU call: multi-GPU tests (group API, torchrun worker), bench --gpus 2, cfg4 chain breakdown
set -u
out=gpurun_out
tag=${1:-r02f}
python -m pytest tests/test_gpu_multi.py -x -q > $out/${tag}_multi_tests.txt 2>&1
echo "pytest rc=$?"; tail -15 $out/${tag}_multi_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $out/${tag}_bench_2gpu.json 2> $out/${tag}_bench_2gpu.err
echo "bench2 rc=$?"; tail -5 $out/${tag}_bench_2gpu.err
PANSIM_FINE_TIMING=1 python tools/run_config.py --pop_size 10000 --core_size 5000000 --pan_genes 20000 --gens 5 --warm 2 --pairs 1000 > $out/${tag}_cfg4_fine.log 2>&1
tail -3 $out/${tag}_cfg4_fine.log
