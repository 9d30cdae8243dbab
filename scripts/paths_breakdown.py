"""Where the 'paths' stage goes on N GPUs (run under torchrun): CUDA-event time of every sub-step of
distributed.sharded_paths, max over ranks.  Synthetic tracks, 300 frames per rank, c2 geometry."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshflow_b200 import DeviceCore, MeshSpec, distributed as mfd, workloads  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
F, W, H, R = 300, 1920, 1080, 16
core = DeviceCore(MeshSpec(W, H, R, R), device=dev)
V = core.mesh.vertices
tr = workloads.synthetic_tracks(np.random.default_rng(rank), F, 3700, W, H)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t = dict(early=d(tr["early"]), late=d(tr["late"]), offset=d(tr["offset"]), keep=d(tr["keep"]), pair_start=d(tr["pair_start"]),
         homographies=d(tr["homographies"].reshape(-1, 9)))
plan = mfd.ShardPlan(rank, world, [F] * world) if world > 1 else None
ident = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)
names = ["vertex_motion", "gather_pairs", "prefix", "jacobi(vertex shard)", "gather_paths", "whole sharded_paths()"]
acc = np.zeros(len(names))
reps = 10
for it in range(reps + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 2)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    vel = core.vertex_velocities(t["early"], t["late"], t["offset"], t["keep"], t["pair_start"], t["homographies"], pair_start_host=tr["pair_start"])
    ev[1].record()
    if world > 1:
        counts = [plan.pairs_needed(r) for r in range(world)]
        vel_all, hp = mfd.gather_pairs(vel, t["homographies"], counts, plan.group)
    else:
        vel_all, hp = vel[:F - 1], t["homographies"][:F - 1]
    homs = torch.cat([hp, ident])
    ev[2].record()
    u = core.prefix_displacements(vel_all)
    ev[3].record()
    v0, v1, _ = mfd.vertex_shard(V, world, rank)
    s = torch.empty_like(u)
    core.stabilized_displacements(u, homs, 0, vertex_range=(v0, v1), out=s)
    ev[4].record()
    s = mfd.gather_paths(s.view(u.shape[0], V, 2), V, plan.group if plan else None)
    ev[5].record()
    mfd.sharded_paths(core, t, F, 0, pair_start_host=tr["pair_start"], plan=plan)
    ev[6].record()
    torch.cuda.synchronize()
    if it >= 3:
        acc += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))]) / reps
tt = torch.tensor(acc, device=dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"n_gpus": world, "frames_total": F * world, "ms (max over ranks)": {n: round(v, 3) for n, v in zip(names, tt.tolist())}}))
if world > 1:
    dist.destroy_process_group()
