"""Pinned host <-> device copy bandwidth per rank and in aggregate, all ranks copying at once (run under torchrun):
why the host-to-host number (e2e) stops scaling with the number of GPUs on one box.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_probe.py

Each rank times, with CUDA events, 20 rounds of (a) H2D only, (b) D2H only, (c) both directions at once on two
streams, 256 MiB per copy from / to pinned memory allocated by the rank itself (first touched on the NUMA node the
rank runs on; with --bind each rank first adopts its GPU's NVML CPU affinity, as bench.py does)."""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if "--bind" in sys.argv:
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out.fill_(2)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(fn, rounds=20):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(rounds):
        fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return n * rounds / (e0.elapsed_time(e1) / 1e3) / 1e9


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


res = torch.tensor([timed(h2d), timed(d2h), timed(both)], dtype=torch.float64, device="cuda")
allr = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    per = [[round(v, 1) for v in r.tolist()] for r in allr]
    print(json.dumps({"n_gpus": world, "bind_numa": "--bind" in sys.argv, "cores": os.cpu_count(),
                      "per_rank_gbs [h2d, d2h, each direction while both run]": per,
                      "aggregate_h2d_gbs": round(sum(p[0] for p in per), 1), "aggregate_d2h_gbs": round(sum(p[1] for p in per), 1),
                      "aggregate_bidirectional_gbs_per_direction": round(sum(p[2] for p in per), 1)}))
if world > 1:
    dist.destroy_process_group()
