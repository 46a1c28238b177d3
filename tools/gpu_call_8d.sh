#!/bin/bash
# which ingredient of the 8-rank run makes every rank's kernels ~17 % slower?
set -u
out=gpurun_out
tag=${1:-r02n8d}
export BENCH_DIAG=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
A="bench.py --gpus 8 --steps 50 --warmup 5 --repeats 3 --no-cfg4 --no-cpu-baseline"
: > $out/${tag}_diag.jsonl
BENCH_DIAG_NAME=torchrun_nccl_shard_comm $TR --master-port 29521 $A >> $out/${tag}_diag.jsonl 2> $out/${tag}_a.err
BENCH_DIAG_NAME=torchrun_nccl_noshard BENCH_NO_SHARD=1 $TR --master-port 29522 $A >> $out/${tag}_diag.jsonl 2> $out/${tag}_b.err
BENCH_DIAG_NAME=torchrun_gloo_shard_nocomm BENCH_DIST_BACKEND=gloo BENCH_NO_COMM=1 $TR --master-port 29523 $A >> $out/${tag}_diag.jsonl 2> $out/${tag}_c.err
BENCH_DIAG_NAME=torchrun_gloo_noshard BENCH_DIST_BACKEND=gloo BENCH_NO_SHARD=1 $TR --master-port 29524 $A >> $out/${tag}_diag.jsonl 2> $out/${tag}_d.err
# independent processes, shard configuration of rank i of 8, no process group at all
for i in 0 1 2 3 4 5 6 7; do
  BENCH_DIAG_NAME=independent_shardcfg_$i BENCH_FAKE_WORLD=8 BENCH_FAKE_RANK=$i CUDA_VISIBLE_DEVICES=$i python bench.py --steps 50 --warmup 5 --repeats 3 --no-cfg4 --no-cpu-baseline >> $out/${tag}_diag_ind$i.jsonl 2> $out/${tag}_e$i.err &
done
wait
cat $out/${tag}_diag_ind*.jsonl >> $out/${tag}_diag.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02n8d_diag.jsonl'):
    try:
        d=json.loads(l)
        print(d['diag'].ljust(30), 'step %.1f core %.1f select %.1f acc %.1f' % (d['us_per_step'], d['core_us'], d['select_us'], d['acc_us']))
    except Exception as e:
        print('ERR', l[:100])
PY
