#!/bin/bash
# are 8 INDEPENDENT single-GPU runs (no NCCL, no torchrun) slower than one alone on the same box?
set -u
out=gpurun_out
tag=${1:-r02n8c}
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader -lms 250 > $out/${tag}_smi.csv &
SMI=$!
CUDA_VISIBLE_DEVICES=0 python bench.py --no-cpu-baseline --no-cfg4 --steps 100 --repeats 5 > $out/${tag}_alone.json 2> $out/${tag}_alone.err
echo "MARK concurrent start" >> $out/${tag}_smi.csv
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i python bench.py --no-cpu-baseline --no-cfg4 --steps 100 --repeats 20 > $out/${tag}_conc$i.json 2> $out/${tag}_conc$i.err &
done
wait %2 %3 %4 %5 %6 %7 %8 %9 2>/dev/null
sleep 1
kill $SMI
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02n8c_*.json')):
    try:
        d=json.load(open(f)); b=d['roofline']['breakdown_ms']
        print(f.split('_')[-1], 'step %.1f us core %.1f select %.1f acc %.1f clocks %s' % (1e3*d['ms_per_step'], 1e3*b['core_mut'], 1e3*b['select'], 1e3*b['acc_step'], d['clocks']))
    except Exception as e:
        print(f, 'ERR', e)
PY
