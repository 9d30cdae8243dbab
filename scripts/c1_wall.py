"""BASELINE.json configs[0]: wall time of stabilize() file -> file on videos/video-1 (494 frames, 640x360, defaults,
ORIGINAL), this implementation and -- with --reference -- the UNMODIFIED reference (baseline/_ref) on the same box.

    python scripts/c1_wall.py [--reference] [--calls 3]"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--reference", action="store_true")
ap.add_argument("--calls", type=int, default=3)
args = ap.parse_args()
video = os.path.join(ROOT, "baseline", "_ref", "video-1.m4v")
out = {"video": "videos/video-1/video-1.m4v (494 frames, 640x360)", "cores": os.cpu_count()}
tmp = tempfile.mkdtemp()

import torch  # noqa: E402
from meshflow_b200 import MeshFlowStabilizer  # noqa: E402
st = MeshFlowStabilizer()
walls = []
for i in range(args.calls):
    t0 = time.perf_counter()
    tup = st.stabilize(video, os.path.join(tmp, "ours.m4v"), 0)
    torch.cuda.synchronize()
    walls.append(time.perf_counter() - t0)
out["meshflow_b200"] = {"wall_s": [round(w, 3) for w in walls], "frames_per_s_best": 494 / min(walls),
                        "tuple": [float(v) for v in tup], "stage_wall_s": {k: round(v, 4) for k, v in st.last_timings.items()}}
if args.reference:
    from oracle import make_ref
    ref = make_ref.reference_module()
    R = ref.MeshFlowStabilizer()
    t0 = time.perf_counter()
    tup_ref = R.stabilize(video, os.path.join(tmp, "ref.m4v"), 0)
    w = time.perf_counter() - t0
    out["reference"] = {"wall_s": round(w, 2), "frames_per_s": 494 / w, "tuple": [float(v) for v in tup_ref]}
    out["speedup_stabilize_file_to_file"] = w / min(walls)
    out["tuple_rel_diff"] = [abs(a - b) / abs(b) for a, b in zip(out["meshflow_b200"]["tuple"], out["reference"]["tuple"])]
print(json.dumps(out))
