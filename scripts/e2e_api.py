"""Wall time of the public API on the c2 clip: MeshFlowStabilizer.stabilize_frames() (host tracking, staging,
GPU passes, metric tracking) -- the number bench.py reports as e2e_api.

    python scripts/e2e_api.py [--impl new|r01] [--frames 300] [--calls 3]

--impl r01 imports round 1's host side (variants/r01_api, git-ignored copy of commit 35fdf87's .py files) on top
of the current C-ABI library: the "first measurement" the round-2 target (>= 3x) is quoted against."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="new", choices=["new", "r01"])
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--calls", type=int, default=3)
ap.add_argument("--workers", type=int, default=None)
args = ap.parse_args()

if args.impl == "r01":
    os.environ["MESHFLOW_B200_LIB"] = os.path.join(ROOT, "meshflow_b200", "libmeshflow_b200.so")
    sys.path.insert(0, os.path.join(ROOT, "variants", "r01_api"))
sys.path.insert(1, ROOT)

import torch  # noqa: E402
from meshflow_b200 import MeshFlowStabilizer  # noqa: E402
from tests import synth  # noqa: E402

torch.cuda.set_device(0)
frames = synth.textured_video(np.random.default_rng(1234), args.frames, args.width, args.height)
st = MeshFlowStabilizer(host_workers=args.workers)
walls, details = [], []
for i in range(args.calls):
    t0 = time.perf_counter()
    kw = {} if args.impl == "r01" else {"reuse_output": True}
    r = st.stabilize_frames(frames, 0, **kw)
    torch.cuda.synchronize()
    walls.append(time.perf_counter() - t0)
    details.append({k: round(v, 4) for k, v in r.get("timings", {}).items()})
print(json.dumps({"impl": args.impl, "frames": args.frames, "size": [args.width, args.height], "cores": os.cpu_count(),
                  "wall_s": [round(w, 3) for w in walls], "fps_first_call": args.frames / walls[0],
                  "fps_best": args.frames / min(walls), "crop": [int(c) for c in r["crop_boundaries"]],
                  "tuple": [float(r["cropping_ratio"]), float(r["distortion_score"]), float(r["stability_score"])],
                  "timings": details}))
