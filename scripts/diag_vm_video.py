"""Diagnosis: vertex velocities of a reference clip, pair by pair, GPU (fast and generic path) vs the CPU port.
usage (under gpurun): python scripts/diag_vm_video.py N [max_pairs]"""
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meshflow_b200 import DeviceCore, MeshFlowStabilizer, MeshSpec, _cabi  # noqa: E402
from meshflow_b200 import host_features as hf  # noqa: E402
from oracle import reference_port as port  # noqa: E402

n = int(sys.argv[1])
limit = int(sys.argv[2]) if len(sys.argv) > 2 else None
cap = cv2.VideoCapture(os.path.join(ROOT, "baseline", "_ref", f"video-{n}.m4v"))
frames = []
while True:
    ok, fr = cap.read()
    if not ok:
        break
    frames.append(fr)
if limit:
    frames = frames[:limit + 1]
H, W = frames[0].shape[:2]
tracks = hf.track_all_pairs(frames[:-1], frames[1:])
p = MeshFlowStabilizer.pack_tracks(tracks)
core = DeviceCore(MeshSpec(W, H, 16, 16))
to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(core.device)
args = (to(p["early"]), to(p["late"]), to(p["offset"]), to(p["keep"]), to(p["pair_start"]), to(p["homographies"].reshape(-1, 9)))
fast, cf = core.vertex_velocities(*args, pair_start_host=p["pair_start"], return_counts=True)
gen, cg = core.vertex_velocities(*args, pair_start_host=None, return_counts=True)
fast, gen, cf, cg = fast.cpu().numpy(), gen.cpu().numpy(), cf.cpu().numpy(), cg.cpu().numpy()
print("pairs", len(tracks), "max features per pair", p["max_pair"], "fast==generic:", np.array_equal(fast.view(np.uint32), gen.view(np.uint32)),
      "counts equal:", np.array_equal(cf, cg))
P = port.Params()
bad = 0
for t, tr in enumerate(tracks):
    e, l = tr.compacted()
    ref = port.vertex_velocities_from_matches(P, W, H, e.reshape(-1, 1, 2), l.reshape(-1, 1, 2), tr.homography)
    for name, got in (("fast", fast[t]), ("generic", gen[t])):
        if not np.array_equal(got.view(np.uint32), ref.view(np.uint32)):
            d = np.argwhere(got.view(np.uint32) != ref.view(np.uint32))
            bad += 1
            if bad <= 12:
                r, c, k = d[0]
                print(f"pair {t} ({len(tr.keep)} candidates, {int(tr.keep.sum())} kept) {name}: {len(d)} values differ; first at vertex ({r},{c}) comp {k}: "
                      f"got {got[r, c, k]!r} ref {ref[r, c, k]!r} members {cf[t].reshape(17, 17)[r, c]}")
print("mismatching (pair, path) combinations:", bad)
