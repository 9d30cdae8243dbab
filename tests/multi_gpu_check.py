"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank computes the WHOLE video alone (1-GPU stage sequence) and its own frame shard through the
sharded, streamed schedule (all-gathered velocities, vertex-sharded Jacobi, all-gathered paths, local
warp, ncclMax crop); the shard must equal the corresponding slice bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshflow_b200 import DeviceCore, MeshSpec, StreamedCore  # noqa: E402
from tests import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, R, C, F = 320, 180, 8, 8, 11                      # F frames per rank
    total = world * F
    rng = np.random.default_rng(2025)                        # same data on every rank
    frames = rng.integers(0, 256, (total, H, W, 3), dtype=np.uint8)
    tr = synth.synthetic_tracks(rng, total, 500, W, H)       # pair t joins frames t, t+1 (last one unused)
    core = DeviceCore(MeshSpec(W, H, R, C))
    dev = core.device
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
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
        a, b = tr["pair_start"][rank * F], tr["pair_start"][(rank + 1) * F]
        pin = lambda x: torch.from_numpy(np.ascontiguousarray(x)).pin_memory()
        tracks = dict(early=pin(tr["early"][a:b]), late=pin(tr["late"][a:b]), offset=pin(tr["offset"][a:b]),
                      keep=pin(tr["keep"][a:b]),
                      pair_start=pin((tr["pair_start"][rank * F:(rank + 1) * F + 1] - a).astype(np.int32)),
                      homographies=pin(tr["homographies"][rank * F:(rank + 1) * F].reshape(-1, 9)))
        h_out = torch.zeros((F, H, W, 3), dtype=torch.uint8).pin_memory()
        enc2, u2, s2 = StreamedCore(core, chunk_frames=4).run(pin(frames[rank * F:(rank + 1) * F]), tracks, h_out, definition)
        torch.cuda.synchronize()
        assert torch.equal(u2, u), "gathered unstabilized paths differ"
        assert torch.equal(s2, s), "vertex-sharded Jacobi + gather differs from the single-GPU solve"
        assert core.decode_crop(enc2) == core.decode_crop(enc), "all-reduced crop differs"
        assert np.array_equal(h_out.numpy(), ref[rank * F:(rank + 1) * F]), "sharded frames differ"
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU parity OK on {world} GPUs: paths, crop {core.decode_crop(enc)} and frames bit-identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
