#!/usr/bin/env bash
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "warp or reciprocal or stabilize" 2>&1 | tail -3
for v in "" variants/lib_mb6.so; do
  echo "== lib: ${v:-default(mb8)}"
  MESHFLOW_B200_LIB=${v:+$PWD/$v} python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tracks synthetic 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k[:12]:round(v,3) for k,v in d['stages_ms'].items()}, round(d['roofline']['frac'],4))"
done
