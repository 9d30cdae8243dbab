"""Generates the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports how4rd/meshflow's ``meshflowstabilizer.py`` from /root/reference, calls its private stage
methods on small inputs, asserts that ``oracle/reference_port.py`` and ``oracle/spec.py`` reproduce
every output (bit-exact for integer/pixel work, 1e-12 relative for the float64 paths) and writes
``tests/golden/*.npz``.  The fixtures pin the oracle and are what the GPU tests compare against.
Values depend on the OpenCV build (recorded in the fixture); regenerate after changing the wheel.
"""
import contextlib
import hashlib
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import cv2  # noqa: E402
import tqdm  # noqa: E402

# silence the reference's progress bars
tqdm.trange = lambda n: contextlib.nullcontext(
    type("T", (), {"set_description": lambda s, d: None, "__iter__": lambda s: iter(range(n))})())

import meshflowstabilizer as ref  # noqa: E402
from oracle import reference_port as port, spec  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def video_frames(n):
    cap = cv2.VideoCapture("/root/reference/videos/video-1/video-1.m4v")
    frames = [cap.read()[1] for _ in range(n)]
    cap.release()
    return frames


def golden_vertex_motion_and_paths():
    """video-1, first 13 frames, reference defaults: matched features of 4 pairs (hot-path input),
    velocities of all 12 pairs, u, homographies, lambda and s for the four weight definitions."""
    frames = video_frames(13)
    R = ref.MeshFlowStabilizer()
    h, w = frames[0].shape[:2]
    feats, vels, homs = {}, [], []
    spy = R._get_matched_features_and_homography
    captured = []
    R._get_matched_features_and_homography = lambda a, b: captured.append(spy(a, b)) or captured[-1]
    for t in range(12):
        v, hm = R._get_unstabilized_vertex_velocities(frames[t], frames[t + 1])
        vels.append(v); homs.append(hm)
        e, l, _ = captured[-1]
        # oracle checks
        assert np.array_equal(port.vertex_velocities_from_matches(port.Params(), w, h, e, l, hm), v)
        assert np.array_equal(spec.vertex_velocities(e.reshape(-1, 2), l.reshape(-1, 2), hm, w, h, 16, 16, 10, 10), v)
        if t < 4:
            feats[f"early_{t}"] = e.reshape(-1, 2)
            feats[f"late_{t}"] = l.reshape(-1, 2)
    R._get_matched_features_and_homography = spy
    u, H = R._get_unstabilized_vertex_displacements_and_homographies(13, frames)
    assert np.array_equal(u, spec.prefix_displacements(np.stack(vels)))
    out = dict(velocities=np.stack(vels), pair_homographies=np.stack(homs), u=u, homographies=H, **feats)
    for d in range(4):
        s = R._get_stabilized_vertex_displacements(13, frames, d, u, H)
        lam = R._get_adaptive_weights(13, w, h, d, H)
        assert np.abs(spec.jacobi_banded(u, H, w, h, 10, 100, d) - s).max() <= 1e-12 * np.abs(s).max()
        assert np.abs(spec.adaptive_lambda(H, w, h, d) - lam).max() <= 1e-12
        out[f"s_{d}"] = s
        out[f"lambda_{d}"] = np.asarray(lam, dtype=np.float64)
        out[f"stability_{d}"] = np.float64(R._compute_stability_score(13, s))
        assert abs(spec.stability_score(s) - out[f"stability_{d}"]) <= 1e-12
    out["frame_size"] = np.array([w, h])
    np.savez_compressed(os.path.join(HERE, "video1_motion_paths.npz"), **out)
    print("video1_motion_paths.npz", sha(u), sha(out["s_0"]))


def golden_jacobi_synthetic():
    """Synthetic 6x6 mesh, 48 frames, radius 7, 30 iterations: the four definitions."""
    rng = np.random.default_rng(17)
    F, Rm = 48, 6
    u = np.cumsum(rng.normal(0, 2.0, (F, Rm + 1, Rm + 1, 2)), axis=0)
    homs = np.tile(np.eye(3), (F, 1, 1))
    homs[:, :2, :2] += rng.normal(0, 0.02, (F, 2, 2))
    homs[:, :2, 2] = rng.normal(0, 12.0, (F, 2))
    homs[5, :2, :2] = [[np.cos(0.4), -np.sin(0.4)], [np.sin(0.4), np.cos(0.4)]]     # complex eigenvalue pair
    homs[-1] = np.eye(3)
    R = ref.MeshFlowStabilizer(mesh_row_count=Rm, mesh_col_count=Rm, temporal_smoothing_radius=7,
                               optimization_num_iterations=30)
    frames = [np.zeros((270, 480, 3), np.uint8)]
    out = dict(u=u, homographies=homs, size=np.array([480, 270]), radius=np.int64(7), iterations=np.int64(30))
    for d in range(4):
        s = R._get_stabilized_vertex_displacements(F, frames, d, u, homs)
        assert np.abs(spec.jacobi_banded(u, homs, 480, 270, 7, 30, d) - s).max() <= 1e-12 * np.abs(s).max()
        out[f"s_{d}"] = s
        out[f"lambda_{d}"] = np.asarray(R._get_adaptive_weights(F, 480, 270, d, homs), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "jacobi_synthetic.npz"), **out)
    print("jacobi_synthetic.npz", sha(out["s_0"]))


def golden_warp_small():
    """3 frames of video-1 shrunk to 160x90, 8x8 mesh, synthetic displacements (a mild case and one
    with folds): stabilized frames, per-frame crop edges are implied by the crop, cropped frames."""
    frames = [cv2.resize(f, (160, 90), interpolation=cv2.INTER_AREA) for f in video_frames(3)]
    out = dict(frames=np.stack(frames), opencv=np.array(cv2.__version__))
    for name, amp in (("mild", 2.0), ("wild", 9.0)):
        rng = np.random.default_rng(int(amp * 10))
        u = np.cumsum(rng.normal(0, 1.5, (3, 9, 9, 2)), axis=0)
        s = u + rng.normal(0, amp, (3, 9, 9, 2)) + rng.normal(0, amp, (3, 1, 1, 2))
        R = ref.MeshFlowStabilizer(mesh_row_count=8, mesh_col_count=8, color_outside_image_area_bgr=(10, 200, 30))
        stab, crop = R._get_stabilized_frames_and_crop_boundaries(3, frames, u, s)
        p = port.Params(mesh_row_count=8, mesh_col_count=8, color_outside_image_area_bgr=(10, 200, 30))
        pstab, pcrop = port.warp_frames_and_crop(p, frames, u, s)
        sstab, scrop = spec.warp_stage(frames, u, s, 8, 8, (10, 200, 30))
        assert all(np.array_equal(a, b) for a, b in zip(stab, pstab)) and tuple(crop) == tuple(pcrop)
        assert all(np.array_equal(a, b) for a, b in zip(stab, sstab)) and tuple(int(c) for c in crop) == tuple(scrop)
        out[f"{name}_u"] = u; out[f"{name}_s"] = s
        out[f"{name}_stabilized"] = np.stack(stab)
        out[f"{name}_crop"] = np.array([int(c) for c in crop])
        if crop[0] <= crop[2] and crop[1] <= crop[3]:
            cropped = R._crop_frames(stab, crop)
            assert all(np.array_equal(a, b) for a, b in zip(cropped, spec.crop_stage(stab, crop)))
            out[f"{name}_cropped"] = np.stack(cropped)
        print("warp_small", name, tuple(int(c) for c in crop), sha(out[f"{name}_stabilized"]))
    np.savez_compressed(os.path.join(HERE, "warp_small.npz"), **out)


def golden_end_to_end():
    """video-1, first 12 frames, defaults, ORIGINAL: the tuple stabilize() returns, the crop and
    hashes of the frames (the frames themselves are too big to commit)."""
    frames = video_frames(12)
    R = ref.MeshFlowStabilizer()
    u, H = R._get_unstabilized_vertex_displacements_and_homographies(12, frames)
    s = R._get_stabilized_vertex_displacements(12, frames, 0, u, H)
    stab, crop = R._get_stabilized_frames_and_crop_boundaries(12, frames, u, s)
    cropped = R._crop_frames(stab, crop)
    cr, ds = R._compute_cropping_ratio_and_distortion_score(12, frames, cropped)
    ss = R._compute_stability_score(12, s)
    o = port.stabilize_frames(port.Params(), frames, 0)
    assert (cr, ds, ss) == (o["cropping_ratio"], o["distortion_score"], o["stability_score"])
    np.savez_compressed(os.path.join(HERE, "video1_end_to_end.npz"),
                        crop=np.array([int(c) for c in crop]), cropping_ratio=np.float32(cr),
                        distortion_score=np.float32(ds), stability_score=np.float64(ss),
                        sha_u=np.array(sha(u)), sha_s=np.array(sha(s)), sha_stabilized=np.array(sha(np.stack(stab))),
                        sha_cropped=np.array(sha(np.stack(cropped))), opencv=np.array(cv2.__version__),
                        first_cropped_row=cropped[0][0], frames=np.int64(12))
    print("video1_end_to_end.npz", tuple(int(c) for c in crop), cr, ds, ss)


if __name__ == "__main__":
    golden_vertex_motion_and_paths()
    golden_jacobi_synthetic()
    golden_warp_small()
    golden_end_to_end()
