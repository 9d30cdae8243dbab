"""Golden record of BASELINE.json configs[0]: the UNMODIFIED reference on the full ``videos/video-1``
(494 frames, 640x360, constructor defaults) for all four ADAPTIVE_WEIGHTS_DEFINITION_* variants.

Run in the build container only (``/root/reference`` does not exist on the GPU box; takes ~35 min on 8
cores -- the reference's warp stage costs ~0.8 s per frame):

    python tests/golden/make_golden_c1.py

The stages are called through the reference's own private methods in the order ``stabilize()`` calls
them (mfs.py:148-162); the feature matching + vertex motion, which do not depend on the definition,
run once.  Writes ``tests/golden/video1_full.npz``:

  u_sha, homographies_sha                      sha256[:16] of the C-contiguous float64 bytes
  homographies (494,3,3), u_sample, s_sample_d arrays at ``sample_vertices`` (17 of the 289 vertices)
  crop_d, tuple_d, s_sha_d, stabilized_sha_d, cropped_sha_d, lambda_d

``tests/test_gpu_video1.py`` runs ``meshflow_b200.MeshFlowStabilizer.stabilize()`` on the same file on
the B200 and compares against this record (bit-exact hashes for u, homographies, crop, cropped frames;
<= 1e-9 relative for s; <= 1e-4 relative for the metric tuple).
"""
import contextlib
import hashlib
import os
import sys
import time
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

import cv2  # noqa: E402
import tqdm  # noqa: E402

tqdm.trange = lambda n: contextlib.nullcontext(
    type("T", (), {"set_description": lambda s, d: None, "__iter__": lambda s: iter(range(n))})())

import meshflowstabilizer as ref  # noqa: E402

VIDEO = "/root/reference/videos/video-1/video-1.m4v"
# python tests/golden/make_golden_c1.py --video N [--definitions 0,2]  ->  tests/golden/videoN_full.npz


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("limit", nargs="?", type=int, default=None, help="frame limit for a quick dry run")
    ap.add_argument("--video", type=int, default=1)
    ap.add_argument("--definitions", default="0,1,2,3")
    a = ap.parse_args()
    limit = a.limit
    video = f"/root/reference/videos/video-{a.video}/video-{a.video}.m4v"
    definitions = [int(d) for d in a.definitions.split(",")]
    R = ref.MeshFlowStabilizer()
    frames, num_frames, fps, codec = R._get_unstabilized_frames_and_video_features(video)
    if limit:
        frames, num_frames = frames[:limit], limit
    h, w = frames[0].shape[:2]
    t0 = time.time()
    u, homs = R._get_unstabilized_vertex_displacements_and_homographies(num_frames, frames)
    print(f"vertex motion: {time.time() - t0:.1f} s", flush=True)
    V = u.shape[1] * u.shape[2]
    sample = np.arange(0, V, 18)[:17]
    out = dict(cv2_version=np.array(cv2.__version__), numpy_version=np.array(np.__version__),
               num_frames=np.array(num_frames), frame_size=np.array([w, h]), sample_vertices=sample,
               u_sha=np.array(sha(u)), homographies_sha=np.array(sha(homs)), homographies=homs,
               u_sample=u.reshape(num_frames, V, 2)[:, sample], frames_sha=np.array(sha(np.stack(frames))))
    for d in definitions:
        t0 = time.time()
        lam = R._get_adaptive_weights(num_frames, w, h, d, homs)
        s = R._get_stabilized_vertex_displacements(num_frames, frames, d, u, homs)
        stab, crop = R._get_stabilized_frames_and_crop_boundaries(num_frames, frames, u, s)
        cropped = R._crop_frames(stab, crop)
        cr, ds = R._compute_cropping_ratio_and_distortion_score(num_frames, frames, cropped)
        st = R._compute_stability_score(num_frames, s)
        out[f"lambda_{d}"] = np.asarray(lam, dtype=np.float64)
        out[f"s_sample_{d}"] = s.reshape(num_frames, V, 2)[:, sample]
        out[f"s_sha_{d}"] = np.array(sha(s))
        out[f"crop_{d}"] = np.array([int(c) for c in crop], dtype=np.int64)
        out[f"tuple_{d}"] = np.array([float(cr), float(ds), float(st)], dtype=np.float64)
        out[f"tuple_types_{d}"] = np.array([type(cr).__name__, type(ds).__name__, type(st).__name__])
        out[f"stabilized_sha_{d}"] = np.array(sha(np.stack(stab)))
        out[f"cropped_sha_{d}"] = np.array(sha(np.stack(cropped)))
        print(f"definition {d}: {time.time() - t0:.1f} s  crop {tuple(int(c) for c in crop)} tuple "
              f"({float(cr)!r}, {float(ds)!r}, {float(st)!r}) cropped sha {out[f'cropped_sha_{d}']}", flush=True)
        name = f"video{a.video}_full.npz" if not limit else f"video{a.video}_first{limit}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)       # after every definition: a partial record survives


if __name__ == "__main__":
    main()
