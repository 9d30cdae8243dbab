#!/usr/bin/env bash
# ncu evidence, round 2 (prepare + fused pixel pass); run under gpurun, one GPU.  Numbers printed under ncu are never bench values.
#   usage: scripts/ncu_r02.sh <tag> [kernel regex] [extra bench args]
tag=${1:-r02a}
regex=${2:-'warp_fused|cell_setup|row_segments|crop_edges|cell_spans'}
shift; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --quick "$@" > gpurun_out/${tag}_launches.stdout 2>&1
ncu --set full --clock-control none --import-source on -k regex:"${regex}" -s 5 -c 5 \
    -o gpurun_out/${tag}_kernels -f python bench.py --steps 1 --warmup 1 --frames 60 --quick "$@" > /dev/null 2>&1
ls -la gpurun_out | tail -3
