#!/usr/bin/env bash
# scaling diagnostics: bench c2 --diag on N ranks (per-rank solo step time, host enqueue time)
n=${1:-4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 --quick --diag 2>/dev/null | tail -1 | tee -a gpurun_out/r02_diag_N${n}.jsonl | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), {k[:8]:round(v,3) for k,v in d['stages_ms'].items()}, d.get('diag'), round(d['e2e']['value']))"; }
nproc
run
