#!/usr/bin/env bash
# ncu evidence for round 1 (run under gpurun, one GPU).  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${1:-r01}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_${1:-r01}.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_kernel -s 1 -c 1 -o gpurun_out/warp_${1:-r01} -f \
    python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'vertex_median|crop_resize_kernel' -s 2 -c 2 -o gpurun_out/other_${1:-r01} -f \
    python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
