#!/bin/bash
set -u
out=gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct
for cs in 4 8 16 32 64; do
  echo "== col_split $cs"
  PANSIM_TILE_COLSPLIT=$cs python tools/dist_bench.py
  PANSIM_TILE_COLSPLIT=$cs REPS=1 ncu --metrics $M --clock-control none -k regex:pair_tile2_kernel -c 2 --csv --log-file $out/r02t_cs$cs.csv python tools/dist_bench.py > /dev/null 2>&1
  python - $out/r02t_cs$cs.csv <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r))
    if d['ID']=='1': print('   ', d['Metric Name'], d['Metric Value'])
PY
done
