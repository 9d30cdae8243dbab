"""CPU restatement ("port") of the MeshFlow stabilization pipeline -- TEST INFRASTRUCTURE ONLY.

This module is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped package
``meshflow_b200`` never imports anything from ``oracle/``.

It restates how4rd/meshflow's ``meshflowstabilizer.py`` stage by stage with the *same library calls in
the same order* (OpenCV / NumPy / ``statistics``), so that (a) its outputs equal the reference's
bit for bit and (b) its cost structure is the reference's (R*C full-frame passes per warped frame,
dense F x F Jacobi matrices, Python-list medians) -- which is what makes it a fair CPU baseline.

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by the reference
itself).  This port is pinned against the *unmodified reference imported from /root/reference* in the
build container by ``tests/golden/make_golden.py`` (which asserts equality stage by stage and writes
the committed fixtures under ``tests/golden/``).

Every function cites the reference lines it follows as ``mfs.py:N`` (= meshflowstabilizer.py).
"""
from __future__ import annotations

import math
import statistics
from dataclasses import dataclass

import cv2
import numpy as np

ORIGINAL, FLIPPED, CONSTANT_HIGH, CONSTANT_LOW = 0, 1, 2, 3          # mfs.py:32-35
CONSTANT_HIGH_VALUE, CONSTANT_LOW_VALUE = 100, 1                     # mfs.py:39-40


@dataclass
class Params:
    """Constructor arguments of the reference class, same names and defaults (mfs.py:43-49)."""
    mesh_row_count: int = 16
    mesh_col_count: int = 16
    mesh_outlier_subframe_row_count: int = 4
    mesh_outlier_subframe_col_count: int = 4
    feature_ellipse_row_count: int = 10
    feature_ellipse_col_count: int = 10
    homography_min_number_corresponding_features: int = 4
    temporal_smoothing_radius: int = 10
    optimization_num_iterations: int = 100
    color_outside_image_area_bgr: tuple = (0, 0, 255)


# --------------------------------------------------------------------------------------------
# mesh geometry
# --------------------------------------------------------------------------------------------
def vertex_xy(p: Params, width: int, height: int) -> np.ndarray:
    """Rest position of every mesh vertex, (V,1,2) float32, row-major.  mfs.py:881-906."""
    pts = []
    for r in range(p.mesh_row_count + 1):
        for c in range(p.mesh_col_count + 1):
            pts.append([[math.ceil((width - 1) * (c / p.mesh_col_count)),
                         math.ceil((height - 1) * (r / p.mesh_row_count))]])
    return np.array(pts, dtype=np.float32)


# --------------------------------------------------------------------------------------------
# host feature stage (FAST -> PyrLK -> per-subframe RANSAC -> global homography)
# --------------------------------------------------------------------------------------------
def _tracked_in_subframe(p: Params, detector, early_sub, late_sub):
    """FAST corners tracked by pyramidal LK, status-filtered.  mfs.py:581-629."""
    kps = detector.detect(early_sub)
    if len(kps) < p.homography_min_number_corresponding_features:
        return None, None
    pts0 = np.float32(cv2.KeyPoint_convert(kps)[:, np.newaxis, :])
    pts1, status, _ = cv2.calcOpticalFlowPyrLK(early_sub, late_sub, pts0, None)
    keep = status.flatten().astype(bool)
    pts0, pts1 = pts0[keep], pts1[keep]
    if len(pts0) < p.homography_min_number_corresponding_features:
        return None, None
    return pts0, pts1


def _inliers_in_subframe(p: Params, detector, early_sub, late_sub, offset):
    """RANSAC inliers of one subframe, shifted to frame coordinates (float64).  mfs.py:531-578."""
    pts0, pts1 = _tracked_in_subframe(p, detector, early_sub, late_sub)
    if pts0 is None:
        return None, None
    _, inlier = cv2.findHomography(pts0, pts1, method=cv2.RANSAC)
    keep = inlier.flatten().astype(bool)
    return pts0[keep] + offset, pts1[keep] + offset


def matched_features_and_homography(p: Params, detector, early, late):
    """Inlier correspondences of a frame pair and their least-squares homography.  mfs.py:455-528."""
    h, w = early.shape[:2]
    sw = math.ceil(w / p.mesh_outlier_subframe_col_count)
    sh = math.ceil(h / p.mesh_outlier_subframe_row_count)
    e_parts, l_parts = [], []
    for x0 in range(0, w, sw):
        for y0 in range(0, h, sh):
            e, l = _inliers_in_subframe(p, detector, early[y0:y0 + sh, x0:x0 + sw],
                                        late[y0:y0 + sh, x0:x0 + sw], [x0, y0])
            if e is not None:
                e_parts.append(e)
            if l is not None:
                l_parts.append(l)
    e_all = np.concatenate(e_parts)
    l_all = np.concatenate(l_parts)
    if len(e_all) < p.homography_min_number_corresponding_features:
        return None, None, None
    hom, _ = cv2.findHomography(e_all, l_all)
    return e_all, l_all, hom


# --------------------------------------------------------------------------------------------
# vertex-motion estimation
# --------------------------------------------------------------------------------------------
def vertex_feature_lists(p: Params, width, height, early_pts, late_pts, hom):
    """Per-vertex lists of residual feature velocities inside each feature's ellipse.  mfs.py:365-452."""
    R, C = p.mesh_row_count, p.mesh_col_count
    xs = [[[] for _ in range(C + 1)] for _ in range(R + 1)]
    ys = [[[] for _ in range(C + 1)] for _ in range(R + 1)]
    if early_pts is None:
        return xs, ys
    resid = late_pts - cv2.perspectiveTransform(early_pts, hom)             # mfs.py:420
    for row4 in np.c_[early_pts, resid]:
        fx, fy, rvx, rvy = row4[0]
        frow = (fy / height) * R                                              # mfs.py:426
        fcol = (fx / width) * C                                               # mfs.py:427
        top = max(0, math.ceil(frow - p.feature_ellipse_row_count / 2))      # mfs.py:438
        bot = min(R, math.floor(frow + p.feature_ellipse_row_count / 2))     # mfs.py:439
        for vr in range(top, bot + 1):
            half = p.feature_ellipse_col_count * math.sqrt(
                (1 / 4) - ((vr - frow) / p.feature_ellipse_row_count) ** 2)  # mfs.py:444
            left = max(0, math.ceil(fcol - half))
            right = min(C, math.floor(fcol + half))
            for vc in range(left, right + 1):
                xs[vr][vc].append(rvx)
                ys[vr][vc].append(rvy)
    return xs, ys


def vertex_velocities_from_matches(p: Params, width, height, early_pts, late_pts, hom):
    """Vertex velocities of one frame pair from given correspondences.  mfs.py:323-362."""
    R, C = p.mesh_row_count, p.mesh_col_count
    vxy = vertex_xy(p, width, height)
    glob = (cv2.perspectiveTransform(vxy, hom) - vxy).reshape(R + 1, C + 1, 2)   # mfs.py:325
    xs, ys = vertex_feature_lists(p, width, height, early_pts, late_pts, hom)
    med_x = np.array([[statistics.median(v) if v else 0 for v in row] for row in xs])
    med_y = np.array([[statistics.median(v) if v else 0 for v in row] for row in ys])
    vel_x = (glob[:, :, 0] + med_x).astype(np.float32)                        # mfs.py:354
    vel_y = (glob[:, :, 1] + med_y).astype(np.float32)                        # mfs.py:355
    return np.dstack((cv2.medianBlur(vel_x, 3), cv2.medianBlur(vel_y, 3)))    # mfs.py:359-361


def vertex_velocities(p: Params, detector, early, late):
    """mfs.py:287-362."""
    e, l, hom = matched_features_and_homography(p, detector, early, late)
    h, w = early.shape[:2]
    return vertex_velocities_from_matches(p, w, h, e, l, hom), hom


def accumulate_displacements(p: Params, velocities, homs):
    """Sequential float64 prefix sum of the float32 pair velocities.  mfs.py:268-284."""
    F = len(velocities) + 1
    disp = np.empty((F, p.mesh_row_count + 1, p.mesh_col_count + 1, 2))
    disp[0].fill(0)
    H = np.empty((F, 3, 3))
    H[-1] = np.identity(3)
    for t in range(F - 1):
        disp[t + 1] = disp[t] + velocities[t]
        H[t] = homs[t]
    return disp, H


def unstabilized_displacements(p: Params, detector, frames):
    """mfs.py:236-284."""
    vels, homs = [], []
    for t in range(len(frames) - 1):
        v, h = vertex_velocities(p, detector, frames[t], frames[t + 1])
        vels.append(v)
        homs.append(h)
    return accumulate_displacements(p, vels, homs)


# --------------------------------------------------------------------------------------------
# Jacobi path optimisation
# --------------------------------------------------------------------------------------------
def adaptive_weights(F, width, height, definition, homs):
    """lambda_t per frame.  mfs.py:786-841."""
    if definition in (ORIGINAL, FLIPPED):
        aff = homs.copy()
        aff[:, 2, :] = [0, 0, 1]
        lam = np.empty((F,))
        for t in range(F):
            m = aff[t]
            mags = np.sort(np.abs(np.linalg.eigvals(m)))
            trans = math.sqrt((m[0, 2] / width) ** 2 + (m[1, 2] / height) ** 2)
            ratio = mags[-2] / mags[-1]
            c1 = -1.93 * trans + 0.95
            c2 = 5.83 * ratio + 4.88 if definition == ORIGINAL else 5.83 * ratio - 4.88
            lam[t] = max(min(c1, c2), 0)
        return lam
    if definition == CONSTANT_HIGH:
        return np.full((F,), CONSTANT_HIGH_VALUE)
    if definition == CONSTANT_LOW:
        return np.full((F,), CONSTANT_LOW_VALUE)
    raise ValueError("bad adaptive_weights_definition")


def jacobi_system(p: Params, F, width, height, definition, homs):
    """Dense (F,F) off-diagonal matrix and (F,) diagonal.  mfs.py:713-783 (quirks kept: w[t,t]=1,
    diagonal sums over ALL frames, the masked band keeps its k=0 entry)."""
    rows, cols = np.indices((F, F))
    w = np.exp(-np.square((3 / p.temporal_smoothing_radius) * (rows - cols)))
    lam = adaptive_weights(F, width, height, definition, homs)
    lw = np.matmul(np.diag(lam), w)
    off = -2 * lw
    diag = 1 + 2 * np.sum(lw, axis=1)
    band = np.zeros(off.shape)
    for k in range(-p.temporal_smoothing_radius, p.temporal_smoothing_radius + 1):
        band += np.diag(np.ones(F - abs(k)), k)
    return np.where(band, off, 0), diag


def jacobi_iterate(p: Params, off, diag, x0, b):
    """mfs.py:844-878."""
    x = x0.copy()
    rinv = np.diag(np.reciprocal(diag))
    for _ in range(p.optimization_num_iterations):
        x = np.matmul(rinv, b - np.matmul(off, x))
    return x


def stabilized_displacements(p: Params, width, height, definition, u, homs):
    """Per-vertex Jacobi solve.  mfs.py:632-710 (square meshes only: mfs.py:696-697 indexes
    ``k // (R+1)``, ``k % (C+1)``)."""
    F = u.shape[0]
    off, diag = jacobi_system(p, F, width, height, definition, homs)
    by_coord = np.moveaxis(u, 0, 2)
    out = np.empty(by_coord.shape)
    for k in range((p.mesh_row_count + 1) * (p.mesh_col_count + 1)):
        r = k // (p.mesh_row_count + 1)
        c = k % (p.mesh_col_count + 1)
        out[r][c] = jacobi_iterate(p, off, diag, by_coord[r][c], by_coord[r][c])
    return np.moveaxis(out, 2, 0)


# --------------------------------------------------------------------------------------------
# mesh warp + crop
# --------------------------------------------------------------------------------------------
def warp_frames_and_crop(p: Params, frames, u, s, return_maps=False):
    """Per-cell homography warp of every frame + the global crop rectangle.  mfs.py:909-1108."""
    F = len(frames)
    H_, W_ = frames[0].shape[:2]
    R, C = p.mesh_row_count, p.mesh_col_count
    rest = vertex_xy(p, W_, H_)
    rest_rc = rest.reshape(R + 1, C + 1, 2)
    motion = np.reshape(s - u, (F, -1, 1, 2))                                 # mfs.py:964-967
    map_x0 = np.full((H_, W_), W_ + 1)                                        # mfs.py:983
    map_y0 = np.full((H_, W_), H_ + 1)                                        # mfs.py:984
    grid = np.swapaxes(np.indices((W_, H_), dtype=np.float32), 0, 2).reshape((-1, 1, 2))
    left = np.full(F, 0)
    right = np.full(F, W_ - 1)
    top = np.full(F, 0)
    bottom = np.full(F, H_ - 1)
    out, maps = [], []
    for f in range(F):
        mx, my = np.copy(map_x0), np.copy(map_y0)
        stab_rc = (rest + motion[f]).reshape(R + 1, C + 1, 2)                 # mfs.py:1025
        for r in range(R):
            for c in range(C):
                src_q = rest_rc[r:r + 2, c:c + 2].reshape(-1, 2)
                dst_q = stab_rc[r:r + 2, c:c + 2].reshape(-1, 2)
                h_us, _ = cv2.findHomography(src_q, dst_q)                    # mfs.py:1041
                h_su, _ = cv2.findHomography(dst_q, src_q)                    # mfs.py:1042
                qx, qy = np.transpose(src_q)
                x0, x1 = math.floor(np.min(qx)), math.ceil(np.max(qx))
                y0, y1 = math.floor(np.min(qy)), math.ceil(np.max(qy))
                rect = np.zeros((H_, W_))
                rect[y0:y1 + 1, x0:x1 + 1] = 255
                mask = cv2.warpPerspective(rect, h_us, (W_, H_))              # mfs.py:1052
                cell = cv2.perspectiveTransform(grid, h_su).reshape((H_, W_, 2))
                cx, cy = np.moveaxis(cell, 2, 0)
                mx = np.where(mask, cx, mx)                                   # mfs.py:1060
                my = np.where(mask, cy, my)                                   # mfs.py:1061
        mx32 = mx.reshape((H_, W_, 1)).astype(np.float32)
        my32 = my.reshape((H_, W_, 1)).astype(np.float32)
        out.append(cv2.remap(frames[f], mx32, my32, cv2.INTER_LINEAR,
                             borderValue=p.color_outside_image_area_bgr))     # mfs.py:1063-1069
        if return_maps:
            maps.append((mx32[:, :, 0], my32[:, :, 0]))
        hit = np.where(np.abs(mx - 0) < 1)[1]                                 # mfs.py:1075
        if hit.size > 0:
            left[f] = np.max(hit)
        hit = np.where(np.abs(mx - (W_ - 1)) < 1)[1]                          # mfs.py:1082
        if hit.size > 0:
            right[f] = np.min(hit)
        hit = np.where(np.abs(my - 0) < 1)[0]                                 # mfs.py:1089
        if hit.size > 0:
            top[f] = np.max(hit)
        hit = np.where(np.abs(my - (H_ - 1)) < 1)[0]                          # mfs.py:1096
        if hit.size > 0:
            bottom[f] = np.min(hit)
    crop = (np.max(left), np.max(top), np.min(right), np.min(bottom))        # mfs.py:1103-1108
    if return_maps:
        return out, crop, maps, (left, top, right, bottom)
    return out, crop


def crop_frames(frames, crop):
    """Slice to the crop rectangle and stretch back to W x H.  mfs.py:1111-1157 (``fx``/``fy`` are
    passed like the reference does; OpenCV ignores them because ``dsize`` is given)."""
    H_, W_ = frames[0].shape[:2]
    l, t, r, b = crop
    if (r + 1 - l) / (b + 1 - t) >= W_ / H_:
        scale = H_ / (b + 1 - t)
    else:
        scale = W_ / (r + 1 - l)
    return [cv2.resize(f[t:b + 1, l:r + 1], (W_, H_), fx=scale, fy=scale) for f in frames]


# --------------------------------------------------------------------------------------------
# metrics
# --------------------------------------------------------------------------------------------
def crop_and_distortion(p: Params, detector, frames, cropped):
    """mfs.py:1160-1212 (np.mean / np.min of float32 arrays)."""
    F = len(frames)
    ratios = np.empty((F), dtype=np.float32)
    dist = np.empty((F), dtype=np.float32)
    for f in range(F):
        _, _, hom = matched_features_and_homography(p, detector, frames[f], cropped[f])
        ratios[f] = 1 / (hom[0][0] * hom[1][1])
        aff = np.copy(hom)
        aff[2] = [0, 0, 1]
        mags = np.sort(np.abs(np.linalg.eigvals(aff)))
        dist[f] = mags[-2] / mags[-1]
    return np.mean(ratios), np.min(dist)


def stability_score(s):
    """mfs.py:1216-1259."""
    sx, sy = np.swapaxes(s, 0, 3)
    ex = np.square(np.abs(np.fft.fft(np.diff(sx))))
    ey = np.square(np.abs(np.fft.fft(np.diff(sy))))
    rx = np.sum(ex[:, :, 1:6], axis=2) / np.sum(ex, axis=2)
    ry = np.sum(ey[:, :, 1:6], axis=2) / np.sum(ey, axis=2)
    return (np.mean(rx) + np.mean(ry)) / 2.0


# --------------------------------------------------------------------------------------------
# whole pipeline on in-memory frames (decode / encode stay outside)
# --------------------------------------------------------------------------------------------
def stabilize_frames(p: Params, frames, definition=ORIGINAL, with_metrics=True):
    """mfs.py:102-169 minus decode/encode; returns every intermediate for stage-level parity."""
    if definition not in (ORIGINAL, FLIPPED, CONSTANT_HIGH, CONSTANT_LOW):
        raise ValueError("Invalid value for `adaptive_weights_definition`.")
    det = cv2.FastFeatureDetector_create()
    H_, W_ = frames[0].shape[:2]
    u, homs = unstabilized_displacements(p, det, frames)
    s = stabilized_displacements(p, W_, H_, definition, u, homs)
    stab, crop = warp_frames_and_crop(p, frames, u, s)
    cropped = crop_frames(stab, crop)
    res = dict(u=u, homographies=homs, s=s, stabilized=stab, crop=crop, cropped=cropped)
    if with_metrics:
        cr, ds = crop_and_distortion(p, det, frames, cropped)
        res.update(cropping_ratio=cr, distortion_score=ds, stability_score=stability_score(s))
    return res
