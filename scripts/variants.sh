#!/usr/bin/env bash
python -m pytest tests -m gpu -q -k "warp or crop or smoke" 2>&1 | tail -3
for v in "" variants/lib_mb4.so variants/lib_mb5.so variants/lib_mb8.so; do
  echo "== lib: ${v:-default(mb6)}"
  MESHFLOW_B200_LIB=${v:+$PWD/$v} python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k:round(v,3) for k,v in d['stages_ms'].items()}, round(d['roofline']['frac'],4))"
done
ncu --set full --clock-control none --import-source on -k regex:'warp_kernel|crop_resize_kernel' -s 2 -c 2 -o gpurun_out/warp_v7 -f python bench.py --steps 1 --warmup 1 --frames 60 --no-cpu-baseline > /dev/null 2>&1
