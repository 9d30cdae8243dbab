"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (``tests/test_gpu_multi.py`` spawns it
under pytest; it can also be run by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

1. Core: every rank computes the WHOLE video alone (1-GPU stage sequence) and its own RAGGED frame shard
   through the sharded, streamed schedule (one packed all-gather of velocities + homographies,
   vertex-sharded Jacobi, all-gathered paths, local warp, ncclMax crop); the shard must equal the
   corresponding slice bit for bit.
2. Drop-in API: ``MeshFlowStabilizer.stabilize_frames`` with a shard plan (host tracking of the rank's own
   pairs included) against the single-rank call on the whole clip: same paths, crop, cropped frames and
   metric tuple on every rank.
3. A process that merely runs under torchrun but passes no plan must not exchange anything."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshflow_b200 import DeviceCore, MeshFlowStabilizer, MeshSpec, StreamedCore  # noqa: E402
from meshflow_b200 import distributed as mfd  # noqa: E402
from tests import synth  # noqa: E402


def core_check(rank, world):
    W, H, R, C = 320, 180, 8, 8
    shard = [11 + (r % 3) for r in range(world)]             # ragged: 11, 12, 13, 11, ...
    total = sum(shard)
    rng = np.random.default_rng(2025)                        # same data on every rank
    frames = rng.integers(0, 256, (total, H, W, 3), dtype=np.uint8)
    tr = synth.synthetic_tracks(rng, total, 500, W, H)       # pair t joins frames t, t+1 (last one unused)
    core = DeviceCore(MeshSpec(W, H, R, C))
    dev = core.device
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    plan = mfd.ShardPlan(rank, world, shard)
    plan.validate(frames_local=shard[rank], pairs_local=shard[rank])
    lo, F = plan.first_frame, plan.local_frames
    streamed = StreamedCore(core, chunk_frames=4)
    crop = None
    for definition in (0, 2):
        # --- whole video on this GPU alone
        vel = core.vertex_velocities(d(tr["early"]), d(tr["late"]), d(tr["offset"]), d(tr["keep"]), d(tr["pair_start"]),
                                     d(tr["homographies"].reshape(-1, 9)), pair_start_host=tr["pair_start"])
        u = core.prefix_displacements(vel[:total - 1])
        homs = torch.cat([d(tr["homographies"].reshape(-1, 9))[:total - 1],
                          torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)])
        s = core.stabilized_displacements(u, homs, definition)
        stab, crop_pf = core.warp_frames(d(frames), u, s)
        enc = core.combine_crop(crop_pf)
        ref = core.crop_resize_device(stab, enc).cpu().numpy()
        # --- this rank's shard through the sharded + streamed schedule
        a, b = tr["pair_start"][lo], tr["pair_start"][lo + F]
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        tracks = dict(early=pin(tr["early"][a:b]), late=pin(tr["late"][a:b]), offset=pin(tr["offset"][a:b]),
                      keep=pin(tr["keep"][a:b]),
                      pair_start=pin((tr["pair_start"][lo:lo + F + 1] - a).astype(np.int32)),
                      homographies=pin(tr["homographies"][lo:lo + F].reshape(-1, 9)))
        h_out = torch.zeros((F, H, W, 3), dtype=torch.uint8).pin_memory()
        enc2, u2, s2 = streamed.run(pin(frames[lo:lo + F]), tracks, h_out, definition, plan=plan)
        torch.cuda.synchronize()
        assert torch.equal(u2, u), "gathered unstabilized paths differ"
        assert torch.equal(s2, s), "vertex-sharded Jacobi + gather differs from the single-GPU solve"
        assert core.decode_crop(enc2) == core.decode_crop(enc), "all-reduced crop differs"
        assert np.array_equal(h_out.numpy(), ref[lo:lo + F]), "sharded frames differ"
        # --- no plan: this rank's frames are a whole video of their own, nothing is exchanged
        if F >= 2:
            solo = torch.zeros((F, H, W, 3), dtype=torch.uint8).pin_memory()
            enc3, u3, s3 = streamed.run(pin(frames[lo:lo + F]), tracks, solo, definition)
            torch.cuda.synchronize()
            assert u3.shape[0] == F and torch.equal(u3, core.prefix_displacements(vel[lo:lo + F - 1]))
        crop = core.decode_crop(enc)
    return crop


def api_check(rank, world):
    W, H, F = 320, 180, 7 * world + 3
    frames = synth.textured_video(np.random.default_rng(12), F, W, H, jitter=2.0)
    whole = MeshFlowStabilizer(mesh_row_count=8, mesh_col_count=8, chunk_frames=4).stabilize_frames(frames, 0)
    plan = mfd.ShardPlan.even(F)
    b, e = plan.first_frame, plan.first_frame + plan.local_frames
    part = MeshFlowStabilizer(mesh_row_count=8, mesh_col_count=8, chunk_frames=4).stabilize_frames(
        frames[b:e], 0, plan=plan, lookahead_frame=frames[e] if e < F else None)
    assert np.array_equal(part["u"], whole["u"]) and np.array_equal(part["s"], whole["s"])
    assert np.array_equal(part["homographies"], whole["homographies"])
    assert tuple(part["crop_boundaries"]) == tuple(whole["crop_boundaries"])
    assert all(np.array_equal(x, y) for x, y in zip(part["cropped_frames"], whole["cropped_frames"][b:e]))
    for k in ("cropping_ratio", "distortion_score", "stability_score"):
        assert part[k] == whole[k] and type(part[k]) is type(whole[k]), k
    return tuple(int(c) for c in whole["crop_boundaries"])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    crop = core_check(rank, world)
    crop_api = api_check(rank, world)
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity OK on {world} GPUs: paths, crop {crop} and frames bit-identical (ragged shards); "
              f"stabilize_frames sharded == single rank (crop {crop_api}, frames, metric tuple)")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
