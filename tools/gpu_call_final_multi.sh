#!/bin/bash
# final multi-GPU collection: N ranks under torchrun (bench line incl. cfg4_strong, shard_parity) [+ cfg5 and the group tests at 8]
set -u
out=gpurun_out
N=$1
tag=r02_n$N
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 100 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench$N rc=$?"; tail -2 $out/${tag}_bench.err
if [ "$N" = "8" ]; then
  python -m pytest tests/test_gpu_multi.py -q > $out/${tag}_multi_tests.txt 2>&1; echo "pytest rc=$?" >> $out/${tag}_multi_tests.txt; tail -3 $out/${tag}_multi_tests.txt
  python tools/all_pairs_bench.py --pop_size 20000 --gpus 8 > $out/${tag}_cfg5.json 2> $out/${tag}_cfg5.err; echo "cfg5 rc=$?"; cat $out/${tag}_cfg5.json
  pansim_b200/pansim --pop_size 2000 --core_size 300000 --n_gen 30 --max_distances 20000 --gpus 8 --print_dist --outpref $out/${tag}_cpp8 > $out/${tag}_cpp8.log 2>&1; echo "cpp host --gpus 8 rc=$?"
  pansim_b200/pansim --pop_size 2000 --core_size 300000 --n_gen 30 --max_distances 20000 --gpus 1 --print_dist --outpref $out/${tag}_cpp1 > $out/${tag}_cpp1.log 2>&1; echo "cpp host --gpus 1 rc=$?"
  cmp $out/${tag}_cpp8.tsv $out/${tag}_cpp1.tsv && cmp $out/${tag}_cpp8_per_gen.tsv $out/${tag}_cpp1_per_gen.tsv && cmp $out/${tag}_cpp8_freqs.txt $out/${tag}_cpp1_freqs.txt && echo "CPP_HOST_8GPU_EQUALS_1GPU"
  rm -f $out/${tag}_cpp8.tsv $out/${tag}_cpp1.tsv
fi
