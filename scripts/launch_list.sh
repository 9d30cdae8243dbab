#!/usr/bin/env bash
# per-kernel gpu time of one resident step (300 frames) under ncu; shares only, never bench values
tag=${1:-x}
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches.stdout 2>&1
python - <<PY
import csv, collections
lines=[l for l in open('gpurun_out/${tag}_launches.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
# first 2 hot_path passes are resident (warmup + step); print the launches of kernels in order with times for the second pass
names=[r['Kernel Name'].split('(')[0][:40] for r in rows]; t=[float(r['Metric Value'].replace(',','')) for r in rows]
for n,v in list(zip(names,t))[:80]: print(f'{n:42s} {v/1e3:9.1f} us')
PY
