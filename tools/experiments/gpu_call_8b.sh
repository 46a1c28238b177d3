#!/bin/bash
set -u
out=gpurun_out
tag=${1:-r02n8b}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 100 --warmup 5 > $out/${tag}_bench_8gpu.json 2> $out/${tag}_bench_8gpu.err
echo "bench8 rc=$?"; tail -3 $out/${tag}_bench_8gpu.err
PANSIM_GRAPH=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 5 --no-cfg4 > $out/${tag}_bench_8gpu_nograph.json 2> $out/${tag}_bench_8gpu_nograph.err
echo "bench8 nograph rc=$?"
nproc; cat /proc/cpuinfo | grep "model name" | head -1
