#!/usr/bin/env bash
# quick loop for the warp fast path: parity tests, bench line, one ncu capture (60 frames)
tag=${1:-x}
python -m pytest tests -m gpu -q -x -k "warp or crop or streamed or stabilize" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_${tag}.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${tag}.json')); print(d['value'], d['e2e']['value'], d['stages_ms'], d['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:'warp_fast_kernel|crop_resize' -s 2 -c 2 -o gpurun_out/fast_${tag} -f python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu-baseline > /dev/null 2>&1
