"""In-situ kernel durations of the bench step (CUPTI through torch.profiler: warm caches, no serialisation -- unlike
the ncu launch list).  usage: python scripts/kernel_times.py [bench.py args]  (run under gpurun)"""
import collections
import runpy
import sys

from torch.profiler import ProfilerActivity, profile

args = sys.argv[1:] or ["--steps", "4", "--warmup", "3", "--quick"]
sys.argv = ["bench.py"] + args
steps = int(args[args.index("--steps") + 1]) if "--steps" in args else 20
warm = int(args[args.index("--warmup") + 1]) if "--warmup" in args else 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    try:
        runpy.run_path("bench.py", run_name="__main__")
    except SystemExit:
        pass
tot = collections.Counter()
cnt = collections.Counter()
for e in prof.events():
    if e.device_type.name == "CUDA":
        tot[e.name[:60]] += e.device_time
        cnt[e.name[:60]] += 1
print(f"{'kernel':60s} {'launches':>8s} {'total us':>10s} {'avg us':>9s}")
for k, v in tot.most_common(40):
    print(f"{k:60s} {cnt[k]:8d} {v:10.1f} {v / cnt[k]:9.2f}")
