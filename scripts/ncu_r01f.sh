#!/usr/bin/env bash
# ncu evidence r01f (warp fast path + rolling-row resize); run under gpurun, one GPU.  Numbers printed under ncu are never bench values.
tag=${1:-r01f}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:'warp_fast_kernel|crop_resize|row_segments|cell_setup|cell_spans' -s 5 -c 5 \
    -o gpurun_out/${tag}_kernels -f python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -3
