#!/usr/bin/env bash
# quick loop for the warp kernel: parity tests, bench line, one ncu capture (60 frames)
tag=${1:-x}
python -m pytest tests -m gpu -q -k "warp or crop" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['stages_ms'], d['roofline']['frac'])"
ncu --set full --clock-control none --import-source on -k regex:warp_kernel -s 1 -c 1 -o gpurun_out/warp_${tag} -f python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu-baseline > /dev/null 2>&1
