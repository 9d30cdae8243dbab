"""Sharding plan and the small exchanges of the multi-GPU path.

One process per GPU (``torch.distributed``, backend ``nccl``; the CPU test-suite drives the same code
with ``gloo``).  The path shards without any data-path collective inside a stage:

* vertex motion, warp, crop/resize : contiguous FRAME shards (frame pairs / frames are independent);
* Jacobi                            : VERTEX shards (each vertex's trajectory is an independent system).

Between the stages three exchanges remain, all latency- not bandwidth-bound at these sizes:

1. ``gather_pairs``   ONE all-gather of the per-pair vertex velocities ``[P_local, V, 2] f32`` packed
   together with the pair homographies ``[P_local, 9] f64``, after which the float64 prefix sum is
   replicated on every rank (it must stay a sequential scan over ALL frames, mfs.py:281);
2. ``gather_paths``   all-gather of the vertex-sharded solved paths back to ``[F, V, 2] f64``;
3. ``reduce_crop``    one ``all_reduce(MAX)`` on ``[left, top, -right, -bottom]`` (mfs.py:1103-1106).

Sharding is an explicit opt-in: every entry point takes a ``ShardPlan`` (or ``None`` for the
single-GPU path).  A process that merely runs under ``torchrun`` is NOT sharded.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
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


@dataclass
class ShardPlan:
    """Who owns which frames of ONE video.  ``frames[r]`` = number of consecutive frames of rank r
    (ragged shards are fine; a rank may own zero frames).  Rank r supplies the tracks of the pairs that
    START at its frames: ``frames[r]`` pairs, except the last non-empty rank, which supplies one less
    (the video ends there; extra pairs are ignored)."""
    rank: int
    world: int
    frames: List[int]
    group: Optional[object] = field(default=None, repr=False)

    def __post_init__(self):
        if len(self.frames) != self.world or not (0 <= self.rank < self.world):
            raise ValueError(f"ShardPlan: {len(self.frames)} shard sizes for world {self.world}, rank {self.rank}")
        if any(f < 0 for f in self.frames) or sum(self.frames) < 2:
            raise ValueError("ShardPlan: shard sizes must be >= 0 and the video needs at least two frames")

    @classmethod
    def even(cls, num_frames: int, group=None):
        """The default plan of ``frame_shard`` for the calling rank of ``group``."""
        rank, world = world_info(group)
        spans = [frame_shard(num_frames, world, r) for r in range(world)]
        return cls(rank, world, [e - b for b, e in spans], group)

    @property
    def total_frames(self) -> int:
        return sum(self.frames)

    @property
    def first_frame(self) -> int:
        return sum(self.frames[:self.rank])

    @property
    def local_frames(self) -> int:
        return self.frames[self.rank]

    def pairs_needed(self, r: Optional[int] = None) -> int:
        """Pairs rank r must supply: one per owned frame, minus the video's last frame."""
        r = self.rank if r is None else r
        begin = sum(self.frames[:r])
        return max(0, min(begin + self.frames[r], self.total_frames - 1) - begin)

    def validate(self, frames_local: int, pairs_local: int):
        """Local checks plus one (CPU-synchronising) all-gather so that ranks that disagree about the
        plan fail loudly instead of issuing mismatched collectives."""
        if frames_local != self.local_frames:
            raise ValueError(f"rank {self.rank} holds {frames_local} frames, the plan says {self.local_frames}")
        if pairs_local < self.pairs_needed():
            raise ValueError(f"rank {self.rank} supplies {pairs_local} frame pairs, needs {self.pairs_needed()}")
        if self.world > 1:
            seen = [None] * self.world
            dist.all_gather_object(seen, (self.rank, list(self.frames)), group=self.group)
            for r, (rr, fr) in enumerate(seen):
                if rr != r or list(fr) != list(self.frames):
                    raise ValueError(f"rank {self.rank}: rank {r} runs a different shard plan {fr} vs {self.frames}")


def _all_gather_rows(local: torch.Tensor, counts: List[int], group=None):
    """All-gather ``local[:counts[rank]]`` (rows along dim 0, ragged) -> concatenation over ranks."""
    rank, world = world_info(group)
    pmax = max(max(counts), 1)
    pad = torch.zeros((pmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    n = counts[rank]
    if n:
        pad[:n] = local[:n]
    out = torch.empty((world * pmax,) + tuple(pad.shape[1:]), dtype=pad.dtype, device=pad.device)
    dist.all_gather_into_tensor(out, pad, group=group)      # concatenated along dim 0 (the layout gloo also takes)
    if all(c == pmax for c in counts):
        return out
    out = out.view((world, pmax) + tuple(pad.shape[1:]))
    return torch.cat([out[r, :counts[r]] for r in range(world)], dim=0)


def gather_velocities(vel_local: torch.Tensor, counts, group=None):
    """All-gather frame-sharded rows.  ``vel_local`` is ``[>= counts[rank], ...]``; ``counts`` lists every
    rank's number of rows (shards may be ragged, so each is padded to the maximum)."""
    _, world = world_info(group)
    if world == 1:
        return vel_local[:counts[0]]
    return _all_gather_rows(vel_local, list(counts), group)


def gather_pairs(vel_local: torch.Tensor, homs_local: torch.Tensor, counts, group=None):
    """Velocities ``[P, V, 2] f32`` and pair homographies ``[P, 9] f64`` of every rank in ONE all-gather:
    both are packed into one byte row per pair (4*2V + 72 bytes)."""
    _, world = world_info(group)
    if world == 1:
        return vel_local[:counts[0]], homs_local[:counts[0]]
    P = vel_local.shape[0]
    vshape = tuple(vel_local.shape[1:])
    per_pair = 1
    for d in vshape:
        per_pair *= int(d)
    vb = vel_local.contiguous().view(P, per_pair).view(torch.uint8)     # [P, 8V]
    hb = homs_local.contiguous().view(P, 9).view(torch.uint8)           # [P, 72]
    nv = vb.shape[1]
    packed = _all_gather_rows(torch.cat([vb, hb], dim=1), list(counts), group)
    vel = packed[:, :nv].contiguous().view(torch.float32).view((-1,) + vshape)
    homs = packed[:, nv:].contiguous().view(torch.float64).view(-1, 9)
    return vel, homs


def gather_paths(s_full: torch.Tensor, num_vertices: int, group=None):
    """``s_full`` is ``[F, V, 2]`` with only this rank's vertex shard solved; returns the complete
    ``[F, V, 2]`` on every rank."""
    rank, world = world_info(group)
    if world == 1:
        return s_full
    F = s_full.shape[0]
    v0, v1, size = vertex_shard(num_vertices, world, rank)
    mine = torch.zeros((F, size, 2), dtype=s_full.dtype, device=s_full.device)
    mine[:, :v1 - v0] = s_full[:, v0:v1]
    out = torch.empty((world * F, size, 2), dtype=s_full.dtype, device=s_full.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.view(world, F, size, 2)
    return out.permute(1, 0, 2, 3).reshape(F, world * size, 2)[:, :num_vertices].contiguous()


def reduce_crop(crop_enc: torch.Tensor, plan: Optional[ShardPlan] = None):
    """In-place MAX all-reduce of the encoded crop ``[left, top, -right, -bottom]``."""
    if plan is not None and plan.world > 1:
        dist.all_reduce(crop_enc, op=dist.ReduceOp.MAX, group=plan.group)
    return crop_enc


def sharded_paths(core, tracks_dev, frames_local, definition, pair_start_host=None, plan: Optional[ShardPlan] = None):
    """Unstabilized and stabilized vertex paths of the WHOLE video on every rank.

    Each rank holds the tracks of the frame pairs that start at its own frames (pair t joins frames t
    and t+1).  Velocities and pair homographies are all-gathered in one exchange, the float64 prefix
    sum is replicated, the Jacobi solve is sharded by vertex and the solved paths all-gathered.
    Returns (u, s, homographies) of ``plan.total_frames`` frames (``frames_local`` without a plan).
    """
    dev = core.device
    tr = tracks_dev
    vel = core.vertex_velocities(tr["early"], tr["late"], tr["offset"], tr["keep"], tr["pair_start"],
                                 tr["homographies"], pair_start_host=pair_start_host)
    ident = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)
    V = core.mesh.vertices
    if plan is None or plan.world == 1:
        total = int(frames_local)
        if vel.shape[0] < total - 1:
            raise ValueError(f"{vel.shape[0]} frame pairs for {total} frames")
        homs = torch.cat([tr["homographies"][:total - 1], ident])
        u = core.prefix_displacements(vel[:total - 1])
        return u, core.stabilized_displacements(u, homs, definition), homs
    if frames_local != plan.local_frames or vel.shape[0] < plan.pairs_needed():
        raise ValueError(f"rank {plan.rank}: {frames_local} frames / {vel.shape[0]} pairs do not match the shard plan "
                         f"({plan.local_frames} frames, {plan.pairs_needed()} pairs)")
    total = plan.total_frames
    counts = [plan.pairs_needed(r) for r in range(plan.world)]
    vel_all, homs_pairs = gather_pairs(vel, tr["homographies"], counts, plan.group)
    homs = torch.cat([homs_pairs, ident])
    u = core.prefix_displacements(vel_all)
    v0, v1, _ = vertex_shard(V, plan.world, plan.rank)
    s = torch.empty_like(u)
    core.stabilized_displacements(u, homs, definition, vertex_range=(v0, v1), out=s)
    s = gather_paths(s.view(total, V, 2), V, plan.group).view(u.shape)
    return u, s, homs
