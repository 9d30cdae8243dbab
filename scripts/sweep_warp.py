"""BASELINE.json configs[4]: warp-only throughput sweep 720p -> 8K with 16x16 and 64x64 meshes.
Prints one line per point: ms per frame, algorithmic GB/s (6*H*W per frame) and fraction of the
measured HBM copy peak; also crop/resize on the same frames."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshflow_b200 import DeviceCore, MeshSpec
peak = 6439.5
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
rows = []
for (W, H, nf) in [(1280, 720, 64), (1920, 1080, 64), (2560, 1440, 64), (3840, 2160, 32), (7680, 4320, 8)]:
    for R in (16, 64):
        rng = np.random.default_rng(99)
        core = DeviceCore(MeshSpec(W, H, R, R))
        frames = torch.randint(0, 256, (nf, H, W, 3), dtype=torch.uint8, device=core.device)
        u = np.cumsum(rng.normal(0, 2.0, (nf, R + 1, R + 1, 2)), axis=0)
        amp = 2.5 * min(1.0, (W / R) / 120.0)            # keep the mesh un-folded on small cells
        s = u + rng.normal(0, amp, u.shape) + rng.normal(0, 3.0, (nf, 1, 1, 2))
        ud, sd = torch.from_numpy(u).to(core.device), torch.from_numpy(s).to(core.device)
        out = torch.empty_like(frames); out2 = torch.empty_like(frames)
        for _ in range(2):
            _, crop = core.warp_frames(frames, ud, sd, out=out)
        enc = core.combine_crop(crop)
        core.crop_resize_device(out, enc, out=out2)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 5
        ev[0].record()
        for _ in range(reps):
            core.warp_frames(frames, ud, sd, out=out)
        ev[1].record()
        for _ in range(reps):
            core.crop_resize_device(out, enc, out=out2)
        ev[2].record(); torch.cuda.synchronize()
        tw = ev[0].elapsed_time(ev[1]) / reps / nf; tr = ev[1].elapsed_time(ev[2]) / reps / nf
        gb = 6.0 * H * W / 1e9
        rows.append((W, H, R, tw * 1e3, gb / (tw / 1e3), gb / (tw / 1e3) / peak, tr * 1e3, gb / (tr / 1e3), core.decode_crop(enc)))
        print(f"{W}x{H} mesh {R}x{R}: warp {tw*1e3:8.1f} us/frame {gb/(tw/1e3):7.0f} GB/s ({gb/(tw/1e3)/peak*100:4.1f}% of {peak:.0f})  "
              f"resize {tr*1e3:8.1f} us/frame {gb/(tr/1e3):7.0f} GB/s  crop {core.decode_crop(enc)}", flush=True)
        del frames, out, out2
        torch.cuda.empty_cache()
