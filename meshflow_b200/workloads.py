"""Seeded synthetic inputs shared by the CPU and GPU tests and by bench.py.  Test and benchmark
data only: nothing on the product path imports this module."""
from __future__ import annotations

import numpy as np


def random_homography(rng, W, H, rot=0.004, scale=0.003, trans=4.0, persp=2e-6):
    a = rng.normal(0, rot)
    s = np.exp(rng.normal(0, scale))
    Hm = np.array([[s * np.cos(a), -s * np.sin(a), rng.normal(0, trans)],
                   [s * np.sin(a), s * np.cos(a), rng.normal(0, trans)],
                   [rng.normal(0, persp), rng.normal(0, persp), 1.0]])
    return Hm


def synthetic_tracks(rng, P, n_per_pair, W, H, sub_rows=4, sub_cols=4, keep_prob=0.85, local_motion=1.0, homography=None):
    """Un-compacted tracks of P frame pairs in the layout mf_vertex_motion takes.
    Returns a dict of flat arrays plus per-pair homographies.  ``local_motion`` scales the per-feature motion on top
    of the homography (1e-6: an almost static pair, residuals at the float32 resolution of the coordinates);
    ``homography``: keyword arguments for ``random_homography``."""
    sw = -(-W // sub_cols)
    sh = -(-H // sub_rows)
    early, late, offset, keep, start, homs = [], [], [], [], [0], []
    for p in range(P):
        n = int(n_per_pair * rng.uniform(0.7, 1.3))
        Hm = random_homography(rng, W, H, **(homography or {}))
        sx = rng.integers(0, sub_cols, n)
        sy = rng.integers(0, sub_rows, n)
        off = np.stack([sx * sw, sy * sh], axis=1).astype(np.int32)
        wmax = np.minimum(sw, W - off[:, 0]).astype(np.float64)
        hmax = np.minimum(sh, H - off[:, 1]).astype(np.float64)
        e = np.stack([rng.uniform(0, 1, n) * (wmax - 1), rng.uniform(0, 1, n) * (hmax - 1)], axis=1)
        e[: n // 8] = np.round(e[: n // 8])                     # FAST corners are integer valued
        ef = e.astype(np.float32)
        full = ef.astype(np.float64) + off
        w = full[:, 0] * Hm[2, 0] + full[:, 1] * Hm[2, 1] + Hm[2, 2]
        lx = (full[:, 0] * Hm[0, 0] + full[:, 1] * Hm[0, 1] + Hm[0, 2]) / w
        ly = (full[:, 0] * Hm[1, 0] + full[:, 1] * Hm[1, 1] + Hm[1, 2]) / w
        local = (rng.normal(0, 0.8, (n, 2)) + rng.normal(0, 1.5, (1, 2)) * (full[:, :1] / W)) * local_motion
        lf = (np.stack([lx, ly], axis=1) + local - off).astype(np.float32)
        early.append(ef); late.append(lf); offset.append(off)
        keep.append((rng.uniform(0, 1, n) < keep_prob).astype(np.uint8))
        start.append(start[-1] + n)
        homs.append(Hm)
    return dict(early=np.concatenate(early), late=np.concatenate(late), offset=np.concatenate(offset),
                keep=np.concatenate(keep), pair_start=np.asarray(start, dtype=np.int32),
                homographies=np.stack(homs), max_pair=int(np.diff(start).max()))


def synthetic_paths(rng, F, R, C, step=3.0, drift=1.0):
    """Unstabilized vertex displacements (F,R+1,C+1,2) float64 and homographies (F,3,3) (c4 recipe)."""
    cam = np.cumsum(rng.normal(0, drift, (F, 1, 1, 2)), axis=0)
    u = np.cumsum(rng.normal(0, step, (F, R + 1, C + 1, 2)) * 0.15, axis=0) + cam
    u[0] = 0
    homs = np.tile(np.eye(3), (F, 1, 1))
    homs[:, :2, :2] += rng.normal(0, 0.01, (F, 2, 2))
    homs[:, :2, 2] = rng.normal(0, 8.0, (F, 2))
    homs[-1] = np.eye(3)
    return u, homs


def synthetic_warp_inputs(rng, F, W, H, R, C, walk=2.0, per_vertex=2.5, per_frame=3.0):
    """c5 recipe: random frames + (u, s) with s - u = per-vertex noise + per-frame shift."""
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    u = np.cumsum(rng.normal(0, walk, (F, R + 1, C + 1, 2)), axis=0)
    s = u + rng.normal(0, per_vertex, (F, R + 1, C + 1, 2)) + rng.normal(0, per_frame, (F, 1, 1, 2))
    return frames, u, s


def textured_video(rng, F, W, H, jitter=3.0):
    """Corner-rich synthetic video (c2 recipe, scaled): a fixed canvas seen through a jittered camera."""
    import cv2
    cw, ch = W + 160, H + 120
    canvas = rng.integers(0, 256, (ch, cw, 3), dtype=np.uint8)
    canvas = cv2.GaussianBlur(canvas, (0, 0), 3.0)
    n_shapes = max(40, (cw * ch) // 2500)      # ~3.7k FAST candidates per 1080p pair, like real footage
    for _ in range(n_shapes):
        color = tuple(int(v) for v in rng.integers(0, 256, 3))
        x, y = int(rng.integers(0, cw)), int(rng.integers(0, ch))
        if rng.uniform() < 0.5:
            cv2.rectangle(canvas, (x, y), (x + int(rng.integers(4, 40)), y + int(rng.integers(4, 40))), color, -1)
        else:
            cv2.circle(canvas, (x, y), int(rng.integers(3, 20)), color, -1)
    frames = []
    for t in range(F):
        ang = rng.normal(0, 0.004)
        sc = float(np.exp(rng.normal(0, 0.003)))
        tx = 80 + 30 * np.sin(2 * np.pi * t / max(F, 2)) + rng.normal(0, jitter)
        ty = 60 + 10 * np.sin(4 * np.pi * t / max(F, 2)) + rng.normal(0, jitter)
        M = np.array([[sc * np.cos(ang), -sc * np.sin(ang), -tx], [sc * np.sin(ang), sc * np.cos(ang), -ty]])
        frames.append(cv2.warpAffine(canvas, M, (W, H), flags=cv2.INTER_LINEAR))
    return frames
