#!/usr/bin/env bash
# Multi-GPU evidence of round 2, run under `gpurun --gpus N`:  scripts/multi_gpu_r02.sh N [c2 c3 c4 c5 test]
# Keeps one JSON line per configuration under gpurun_out/ (copied to profiles/ for the record).
n=${1:-2}; shift
what=${@:-"test c2 c3 c4 c5 pcie"}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
for w in $what; do
  case $w in
    test) python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3; cp gpurun_out/multi_gpu_check.log gpurun_out/r02_multi_gpu_check_N${n}.log 2>/dev/null;;
    c2) run bench.py --gpus $n --steps 10 --warmup 3 --quick > gpurun_out/r02_c2_N${n}.json 2> gpurun_out/r02_c2_N${n}.err;;
    c3) run bench.py --gpus $n --config c3 --steps 2 --warmup 1 > gpurun_out/r02_c3_N${n}.json 2> gpurun_out/r02_c3_N${n}.err;;
    c4) run bench.py --gpus $n --config c4 --steps 3 --warmup 1 > gpurun_out/r02_c4_N${n}.json 2> gpurun_out/r02_c4_N${n}.err;;
    c5) run bench.py --gpus $n --config c5 --steps 3 --warmup 2 > gpurun_out/r02_c5_N${n}.json 2> gpurun_out/r02_c5_N${n}.err;;
    pcie) run scripts/pcie_probe.py --bind > gpurun_out/r02_pcie_N${n}.json 2> gpurun_out/r02_pcie_N${n}.err; tail -1 gpurun_out/r02_pcie_N${n}.json;;
  esac
  for f in gpurun_out/r02_${w}_N${n}.json; do [ -f "$f" ] && python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["metric"], round(d["value"], 2), d["unit"], "ms/step", round(d["ms_per_step"], 3), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.get("stages_ms", {}).items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
  done
done
tail -n 3 gpurun_out/r02_*_N${n}.err 2>/dev/null | tail -20
