"""Host-side OpenCV front end, unchanged in behaviour from the reference.

The north star keeps FAST detection, pyramidal LK tracking, the per-subframe RANSAC outlier test and
the global least-squares homography on the host (how4rd/meshflow ``meshflowstabilizer.py`` lines
455-629, cited as mfs.py:N) so that the reference and this implementation consume identical
correspondences.  The only difference is the hand-off: instead of compacting the tracked features on
the host, ``track_pair`` returns them UN-compacted together with the two masks folded into one
``keep`` byte per feature; the device applies the mask (``mf_vertex_motion``).  The global homography
still needs the compacted inliers on the host because ``cv2.findHomography`` runs there.
"""
from __future__ import annotations

import contextlib
import math
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import cv2
import numpy as np


@dataclass
class PairTracks:
    """Tracked candidates of one frame pair, before the masks are applied."""
    early_xy: np.ndarray      # (n,2) float32, subframe-relative   (mfs.py:617)
    late_xy: np.ndarray       # (n,2) float32, subframe-relative   (mfs.py:618)
    offset_xy: np.ndarray     # (n,2) int32, subframe top-left     (mfs.py:509)
    keep: np.ndarray          # (n,) uint8: LK status (mfs.py:622) AND RANSAC inlier (mfs.py:569-574)
    homography: np.ndarray    # (3,3) float64 early->late           (mfs.py:524)

    def compacted(self):
        """Inlier correspondences in frame coordinates, float64 -- what the reference passes on."""
        k = self.keep.astype(bool)
        off = self.offset_xy[k].astype(np.float64)
        return self.early_xy[k].astype(np.float64) + off, self.late_xy[k].astype(np.float64) + off


_tls = threading.local()


def _detector():
    # one FAST detector per thread (the reference shares a single one: mfs.py:99)
    det = getattr(_tls, "fast", None)
    if det is None:
        det = cv2.FastFeatureDetector_create()
        _tls.fast = det
    return det


def track_pair(early, late, subframe_rows=4, subframe_cols=4, min_features=4) -> PairTracks:
    """mfs.py:455-629 for one pair of BGR frames.  Raises ValueError where the reference fails
    (no subframe with enough features -> ``np.concatenate`` of an empty list, mfs.py:518; fewer than
    ``min_features`` inliers overall -> the reference returns None and crashes downstream)."""
    det = _detector()
    h, w = early.shape[:2]
    sw = math.ceil(w / subframe_cols)
    sh = math.ceil(h / subframe_rows)
    e_raw, l_raw, off_raw, keep_raw = [], [], [], []
    e_in, l_in = [], []
    for x0 in range(0, w, sw):                       # x outer, y inner (mfs.py:503-504)
        for y0 in range(0, h, sh):
            e_sub = early[y0:y0 + sh, x0:x0 + sw]
            l_sub = late[y0:y0 + sh, x0:x0 + sw]
            kps = det.detect(e_sub)
            if len(kps) < min_features:              # mfs.py:614
                continue
            p0 = np.float32(cv2.KeyPoint_convert(kps)[:, np.newaxis, :])
            p1, status, _ = cv2.calcOpticalFlowPyrLK(e_sub, l_sub, p0, None)
            st = status.flatten().astype(bool)
            if int(st.sum()) < min_features:         # mfs.py:626
                continue
            _, inl = cv2.findHomography(p0[st], p1[st], method=cv2.RANSAC)   # mfs.py:569
            inl = inl.flatten().astype(bool)
            keep = np.zeros(len(p0), dtype=np.uint8)
            keep[np.flatnonzero(st)[inl]] = 1
            e_raw.append(p0[:, 0, :]); l_raw.append(p1[:, 0, :])
            off_raw.append(np.tile(np.array([[x0, y0]], dtype=np.int32), (len(p0), 1)))
            keep_raw.append(keep)
            e_in.append(p0[st][inl] + [x0, y0])      # float64 from here on (mfs.py:578)
            l_in.append(p1[st][inl] + [x0, y0])
    if not e_in:
        raise ValueError("need at least one array to concatenate")      # what mfs.py:518 raises
    e_all = np.concatenate(e_in)
    l_all = np.concatenate(l_in)
    if len(e_all) < min_features:
        raise ValueError(f"fewer than {min_features} corresponding features between two frames")
    hom, _ = cv2.findHomography(e_all, l_all)                            # mfs.py:524
    return PairTracks(np.ascontiguousarray(np.concatenate(e_raw), dtype=np.float32),
                      np.ascontiguousarray(np.concatenate(l_raw), dtype=np.float32),
                      np.ascontiguousarray(np.concatenate(off_raw), dtype=np.int32),
                      np.ascontiguousarray(np.concatenate(keep_raw), dtype=np.uint8), hom)


def default_workers() -> int:
    """One worker per host core this process may use (the pairs are independent; OpenCV releases the GIL)."""
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    return max(1, min(64, n))


_cv_threads_lock = threading.Lock()
_cv_threads_depth = 0
_cv_threads_saved = None


@contextlib.contextmanager
def single_threaded_opencv():
    """While many pairs are tracked concurrently each OpenCV call runs on its caller's thread only:
    the pool already occupies every core, and OpenCV's own parallel_for on quarter-frame subimages
    would only oversubscribe them.  Results do not depend on OpenCV's thread count."""
    global _cv_threads_depth, _cv_threads_saved
    with _cv_threads_lock:
        if _cv_threads_depth == 0:
            _cv_threads_saved = cv2.getNumThreads()
            cv2.setNumThreads(1)
        _cv_threads_depth += 1
    try:
        yield
    finally:
        with _cv_threads_lock:
            _cv_threads_depth -= 1
            if _cv_threads_depth == 0:
                cv2.setNumThreads(_cv_threads_saved)


def track_all_pairs(frames_a, frames_b, subframe_rows=4, subframe_cols=4, min_features=4, workers=None, pool=None):
    """``track_pair`` over many independent pairs on a thread pool (OpenCV releases the GIL).
    Results are returned in pair order and do not depend on the number of workers.  ``pool``: an
    existing ``ThreadPoolExecutor`` to run on (else a temporary one with ``workers`` threads)."""
    n = len(frames_a)
    if workers is None:
        workers = default_workers()
    job = lambda i: track_pair(frames_a[i], frames_b[i], subframe_rows, subframe_cols, min_features)
    if n <= 1 or (pool is None and workers <= 1):
        return [job(i) for i in range(n)]
    with single_threaded_opencv():
        if pool is not None:
            return list(pool.map(job, range(n)))
        with ThreadPoolExecutor(max_workers=workers) as own:
            return list(own.map(job, range(n)))
