#!/usr/bin/env bash
for v in "" variants/lib_nogather.so variants/lib_nomap.so; do
  echo "== lib: ${v:-default}"
  MESHFLOW_B200_LIB=${v:+$PWD/$v} python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tracks synthetic 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k[:12]:round(v,3) for k,v in d['stages_ms'].items()}, round(d['roofline']['frac'],4))"
done
