#!/usr/bin/env bash
# A/B of library variants (variants/lib_<name>.so, see build_variant.sh): short bench line per variant
for v in "$@"; do
  lib=""; [ "$v" != "main" ] && lib="$PWD/variants/lib_${v}.so"
  MESHFLOW_B200_LIB=$lib python bench.py --steps 5 --warmup 3 --quick ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['e2e']['value']), {k[:12]:round(v,3) for k,v in d['stages_ms'].items()}, round(d['roofline']['frac'],4))"
done
