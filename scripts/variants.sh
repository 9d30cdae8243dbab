#!/usr/bin/env bash
python -m pytest tests -m gpu -q -k "crop or resize or warp_matches or stabilize or streamed" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tracks synthetic 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['e2e']['value']), {k[:12]:round(v,3) for k,v in d['stages_ms'].items()}, round(d['roofline']['frac'],4), round(d['crop_resize_gbs']))"
