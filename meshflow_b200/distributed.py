"""Sharding plan and the three small exchanges of the multi-GPU path.

One process per GPU (``torch.distributed``, backend ``nccl``; the CPU test-suite drives the same code
with ``gloo``).  The path shards without any data-path collective inside a stage:

* vertex motion, warp, crop/resize : contiguous FRAME shards (frame pairs / frames are independent);
* Jacobi                            : VERTEX shards (each vertex's trajectory is an independent system).

Between the stages three exchanges remain, all latency- not bandwidth-bound at these sizes:

1. ``gather_velocities``  all-gather of the per-pair vertex velocities ``[P_local, V, 2] f32`` and the
   pair homographies, after which the float64 prefix sum is replicated on every rank (it must stay a
   sequential scan over ALL frames, mfs.py:281);
2. ``gather_paths``       all-gather of the vertex-sharded solved paths back to ``[F, V, 2] f64``;
3. ``reduce_crop``        one ``all_reduce(MAX)`` on ``[left, top, -right, -bottom]`` (mfs.py:1103-1106).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def frame_shard(num_frames: int, world: int, rank: int):
    """Contiguous [begin, end) frame range of ``rank``; the first ``num_frames % world`` ranks get one
    extra frame."""
    base, extra = divmod(num_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def vertex_shard(num_vertices: int, world: int, rank: int):
    """Equal (padded) vertex shards: returns (begin, end, shard_size); ``end - begin`` may be smaller
    than ``shard_size`` on the last ranks (and zero when there are more ranks than vertices)."""
    size = -(-num_vertices // world)
    begin = min(rank * size, num_vertices)
    return begin, min(begin + size, num_vertices), size


def gather_velocities(vel_local: torch.Tensor, counts):
    """All-gather frame-sharded pair velocities.  ``vel_local`` is ``[P_local, ...]``; ``counts`` lists
    every rank's number of pairs (shards may be ragged, so each is padded to the maximum)."""
    rank, world = world_info()
    if world == 1:
        return vel_local
    pmax = max(counts)
    pad = torch.zeros((pmax,) + tuple(vel_local.shape[1:]), dtype=vel_local.dtype, device=vel_local.device)
    pad[:vel_local.shape[0]] = vel_local
    out = torch.empty((world * pmax,) + tuple(pad.shape[1:]), dtype=pad.dtype, device=pad.device)
    dist.all_gather_into_tensor(out, pad)          # concatenated along dim 0 (the layout gloo also takes)
    out = out.view((world, pmax) + tuple(pad.shape[1:]))
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def gather_paths(s_full: torch.Tensor, num_vertices: int):
    """``s_full`` is ``[F, V, 2]`` with only this rank's vertex shard solved; returns the complete
    ``[F, V, 2]`` on every rank."""
    rank, world = world_info()
    if world == 1:
        return s_full
    F = s_full.shape[0]
    v0, v1, size = vertex_shard(num_vertices, world, rank)
    mine = torch.zeros((F, size, 2), dtype=s_full.dtype, device=s_full.device)
    mine[:, :v1 - v0] = s_full[:, v0:v1]
    out = torch.empty((world * F, size, 2), dtype=s_full.dtype, device=s_full.device)
    dist.all_gather_into_tensor(out, mine)
    out = out.view(world, F, size, 2)
    return out.permute(1, 0, 2, 3).reshape(F, world * size, 2)[:, :num_vertices].contiguous()


def reduce_crop(crop_enc: torch.Tensor):
    """In-place MAX all-reduce of the encoded crop ``[left, top, -right, -bottom]``."""
    _, world = world_info()
    if world > 1:
        dist.all_reduce(crop_enc, op=dist.ReduceOp.MAX)
    return crop_enc


def sharded_paths(core, tracks_dev, frames_local, definition, pair_start_host=None):
    """Unstabilized and stabilized vertex paths of the WHOLE video on every rank.

    Each rank holds the tracks of its own ``frames_local`` frame pairs (pair t joins frames t and
    t+1; the last rank's last pair is dropped because the video ends there).  Velocities and pair
    homographies are all-gathered, the float64 prefix sum is replicated, the Jacobi solve is sharded
    by vertex and the solved paths all-gathered.  Returns (u, s, homographies) with
    ``world * frames_local`` frames each.
    """
    rank, world = world_info()
    dev = core.device
    tr = tracks_dev
    vel = core.vertex_velocities(tr["early"], tr["late"], tr["offset"], tr["keep"], tr["pair_start"],
                                 tr["homographies"], pair_start_host=pair_start_host)
    total = world * frames_local
    ident = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)
    if world > 1:
        counts = [int(vel.shape[0])] * world
        vel_all = gather_velocities(vel, counts)[:total - 1]
        homs = torch.cat([gather_velocities(tr["homographies"], counts)[:total - 1], ident])
    else:
        vel_all = vel[:total - 1]
        homs = torch.cat([tr["homographies"][:total - 1], ident])
    u = core.prefix_displacements(vel_all)
    V = core.mesh.vertices
    if world > 1:
        v0, v1, _ = vertex_shard(V, world, rank)
        s = torch.empty_like(u)
        core.stabilized_displacements(u, homs, definition, vertex_range=(v0, v1), out=s)
        s = gather_paths(s.view(total, V, 2), V).view(u.shape)
    else:
        s = core.stabilized_displacements(u, homs, definition)
    return u, s, homs
