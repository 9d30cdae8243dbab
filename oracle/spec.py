"""Per-element CPU specification of the three hot-path subsystems -- TEST INFRASTRUCTURE ONLY.

Where ``oracle/reference_port.py`` follows the reference call by call (and therefore costs
O(cells x pixels) per frame), this module states what every *output element* is, in NumPy, with the
fixed-point models of the OpenCV primitives the reference leans on (``cv2.remap``,
``cv2.warpPerspective``, ``cv2.resize``, ``cv2.medianBlur``, 4-point ``cv2.findHomography``).  It is
what the CUDA kernels are compared with at sizes the port cannot reach (1080p+, F = 10 000).

Pinning: every function here is asserted equal to ``reference_port`` (itself pinned bit-for-bit to
the unmodified reference) by ``tests/test_oracle.py`` (seeded inputs up to 1920x1080 and the committed goldens) and by
``tests/golden/make_golden.py`` (against the reference itself).
The OpenCV models were checked against opencv-python-headless 4.13.0; a different wheel that changes
them makes those tests fail loudly.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs (cpu_baseline, --impl reference, the
parity self-check outside the timed region) may import this.
References are to meshflowstabilizer.py (``mfs.py:N``).
"""
from __future__ import annotations

import math

import numpy as np

INT_MIN, INT_MAX = -2147483648, 2147483647


# --------------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------------
def _two_sum(a, b):
    s = a + b
    t = s - a
    return s, (a - (s - t)) + (b - t)


def fma_f64(a, b, c):
    """Correctly rounded a*b + c in float64 without a hardware fma (Boldo & Melquiond: exact product by Veltkamp
    splitting, the low parts added with rounding to odd, one final rounding).  Valid while nothing over- or
    underflows, which holds for pixel coordinates and homography entries."""
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64))
    split = 134217729.0                                       # 2^27 + 1
    t = split * a; ah = t - (t - a); al = a - ah
    t = split * b; bh = t - (t - b); bl = b - bh
    uh = a * b
    ul = ((ah * bh - uh) + ah * bl + al * bh) + al * bl       # uh + ul == a * b exactly
    th, tl = _two_sum(c, uh)
    v, e = _two_sum(tl, ul)                                   # round tl + ul to odd
    bits = np.ascontiguousarray(v).view(np.int64).copy()
    fix = (e != 0) & ((bits & 1) == 0)
    bits = np.where(fix, bits + np.where((e > 0) == (v > 0), 1, -1), bits)
    return th + bits.view(np.float64).reshape(np.shape(v))


def persp_f64(x, y, M):
    """cv2.perspectiveTransform in float64: w = 1/w (0 if |w| <= eps), then multiply.  The sums are evaluated as
    OpenCV's AVX2 / AVX-512 build evaluates them (GCC contracts x*m0 + y*m1 + m2 to fma(x, m0, y*m1) + m2); checked
    bit for bit against cv2 in tests/test_oracle.py."""
    M = np.asarray(M, dtype=np.float64).reshape(3, 3)
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    lin = lambda r: fma_f64(x, M[r, 0], y * M[r, 1]) + M[r, 2]
    w = lin(2)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(np.abs(w) > np.finfo(np.float64).eps, 1.0 / w, 0.0)
    return lin(0) * w, lin(1) * w


def vertex_xy(W, H, R, C):
    """mfs.py:881-906 -> (V,2) float32, integer valued."""
    out = np.empty(((R + 1) * (C + 1), 2), dtype=np.float32)
    k = 0
    for r in range(R + 1):
        for c in range(C + 1):
            out[k, 0] = math.ceil((W - 1) * (c / C))
            out[k, 1] = math.ceil((H - 1) * (r / R))
            k += 1
    return out


def solve8_partial_pivot(A, b):
    """Batched 8x8 Gaussian elimination with partial pivoting, explicit operation order, no FMA.
    A: (n,8,8), b: (n,8).  The CUDA ``cell_setup`` kernel performs exactly this sequence."""
    A = np.array(A, dtype=np.float64)
    b = np.array(b, dtype=np.float64)
    n = A.shape[0]
    idx = np.arange(n)
    for k in range(8):
        piv = k + np.argmax(np.abs(A[:, k:, k]), axis=1)          # first maximum wins
        rk = A[idx, k].copy(); rp = A[idx, piv].copy()
        A[idx, k] = rp; A[idx, piv] = rk
        bk = b[idx, k].copy(); bp = b[idx, piv].copy()
        b[idx, k] = bp; b[idx, piv] = bk
        for i in range(k + 1, 8):
            f = A[:, i, k] / A[:, k, k]
            A[:, i, k:] = A[:, i, k:] - f[:, None] * A[:, k, k:]
            b[:, i] = b[:, i] - f * b[:, k]
    x = np.zeros((n, 8))
    for i in range(7, -1, -1):
        acc = b[:, i].copy()
        for j in range(i + 1, 8):
            acc = acc - A[:, i, j] * x[:, j]
        x[:, i] = acc / A[:, i, i]
    return x


def homography_4pt(src, dst):
    """Exact homography (h22 = 1) through 4 correspondences, batched: src,dst (n,4,2) float64.
    Stands in for 4-point cv2.findHomography (mfs.py:1041-1042), which rounds its input to float32
    first (callers pass float32-valued coordinates)."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    n = src.shape[0]
    A = np.zeros((n, 8, 8))
    b = np.zeros((n, 8))
    for i in range(4):
        x, y = src[:, i, 0], src[:, i, 1]
        X, Y = dst[:, i, 0], dst[:, i, 1]
        A[:, 2 * i, 0] = x; A[:, 2 * i, 1] = y; A[:, 2 * i, 2] = 1
        A[:, 2 * i, 6] = -(x * X); A[:, 2 * i, 7] = -(y * X); b[:, 2 * i] = X
        A[:, 2 * i + 1, 3] = x; A[:, 2 * i + 1, 4] = y; A[:, 2 * i + 1, 5] = 1
        A[:, 2 * i + 1, 6] = -(x * Y); A[:, 2 * i + 1, 7] = -(y * Y); b[:, 2 * i + 1] = Y
    h = solve8_partial_pivot(A, b)
    return np.concatenate([h, np.ones((n, 1))], axis=1)           # (n,9)


def inverse3x3_adjugate(M):
    """Closed-form inverse, batched (n,9) -> (n,9): d = 1/det, every cofactor times d
    (the order cv2.warpPerspective's cv::invert uses for 3x3)."""
    m = np.asarray(M, dtype=np.float64).reshape(-1, 3, 3)
    a, b, c = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    d, e, f = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    g, h, i = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g)
    with np.errstate(divide="ignore"):
        r = np.where(det != 0, 1.0 / det, 0.0)
    out = np.empty_like(m)
    out[:, 0, 0] = (e * i - f * h) * r
    out[:, 0, 1] = (c * h - b * i) * r
    out[:, 0, 2] = (b * f - c * e) * r
    out[:, 1, 0] = (f * g - d * i) * r
    out[:, 1, 1] = (a * i - c * g) * r
    out[:, 1, 2] = (c * d - a * f) * r
    out[:, 2, 0] = (d * h - e * g) * r
    out[:, 2, 1] = (b * g - a * h) * r
    out[:, 2, 2] = (a * e - b * d) * r
    return out.reshape(-1, 9)


# --------------------------------------------------------------------------------------------
# A.1 vertex motion
# --------------------------------------------------------------------------------------------
def feature_vertex_ranges(fx, fy, W, H, R, C, er, ec):
    """For every feature: first row ``top``, last row ``bot`` and per-row inclusive column range.
    Returns (top, bot, left[n,R+1], right[n,R+1]); rows outside top..bot get left > right.
    mfs.py:426-446 (float64, divide-then-multiply, same operation order)."""
    fx = np.asarray(fx, dtype=np.float64)
    fy = np.asarray(fy, dtype=np.float64)
    frow = (fy / H) * R
    fcol = (fx / W) * C
    top = np.maximum(0, np.ceil(frow - er / 2)).astype(np.int64)
    bot = np.minimum(R, np.floor(frow + er / 2)).astype(np.int64)
    n = fx.shape[0]
    left = np.full((n, R + 1), 1, dtype=np.int64)
    right = np.full((n, R + 1), 0, dtype=np.int64)
    for vr in range(R + 1):
        on = (vr >= top) & (vr <= bot)
        q = (vr - frow) / er
        arg = 0.25 - q * q
        half = ec * np.sqrt(np.where(on, np.maximum(arg, 0.0), 0.0))
        l = np.maximum(0, np.ceil(fcol - half)).astype(np.int64)
        r = np.minimum(C, np.floor(fcol + half)).astype(np.int64)
        left[:, vr] = np.where(on, l, 1)
        right[:, vr] = np.where(on, r, 0)
    return top, bot, left, right


def median_stat(values):
    """statistics.median on float64 values (empty -> 0).  mfs.py:338-353."""
    n = len(values)
    if n == 0:
        return 0.0
    s = np.sort(np.asarray(values, dtype=np.float64))
    if n % 2 == 1:
        return float(s[n // 2])
    return float((s[n // 2 - 1] + s[n // 2]) / 2)


def median3x3_replicate(g):
    """cv2.medianBlur(float32, 3): 3x3 median, replicated border.  mfs.py:359-360."""
    p = np.pad(g, 1, mode="edge")
    h, w = g.shape
    stack = np.stack([p[i:i + h, j:j + w] for i in range(3) for j in range(3)], axis=0)
    return np.sort(stack, axis=0)[4].astype(np.float32)


def vertex_velocities(early, late, Hm, W, H, R, C, er, ec, return_assignment=False):
    """Vertex velocities of one frame pair from inlier correspondences (A.1).
    early, late: (N,2) float64 frame coordinates.  Returns (R+1,C+1,2) float32."""
    vxy = vertex_xy(W, H, R, C)
    gx, gy = persp_f64(vxy[:, 0].astype(np.float64), vxy[:, 1].astype(np.float64), Hm)
    glob_x = gx.astype(np.float32) - vxy[:, 0]                                # mfs.py:325 (f32)
    glob_y = gy.astype(np.float32) - vxy[:, 1]
    V = (R + 1) * (C + 1)
    med = np.zeros((V, 2))
    member = None
    if early is not None and len(early):
        early = np.asarray(early, dtype=np.float64).reshape(-1, 2)
        late = np.asarray(late, dtype=np.float64).reshape(-1, 2)
        px, py = persp_f64(early[:, 0], early[:, 1], Hm)
        rvx, rvy = late[:, 0] - px, late[:, 1] - py                           # mfs.py:420
        top, bot, left, right = feature_vertex_ranges(early[:, 0], early[:, 1], W, H, R, C, er, ec)
        cols = np.arange(C + 1)
        member = (cols[None, None, :] >= left[:, :, None]) & (cols[None, None, :] <= right[:, :, None])
        member = member.reshape(len(early), V)                                # [feature, vertex]
        for v in range(V):
            sel = member[:, v]
            med[v, 0] = median_stat(rvx[sel])
            med[v, 1] = median_stat(rvy[sel])
    vel_x = (glob_x.astype(np.float64) + med[:, 0]).astype(np.float32).reshape(R + 1, C + 1)
    vel_y = (glob_y.astype(np.float64) + med[:, 1]).astype(np.float32).reshape(R + 1, C + 1)
    out = np.dstack((median3x3_replicate(vel_x), median3x3_replicate(vel_y)))
    if return_assignment:
        return out, member
    return out


def prefix_displacements(vel):
    """disp[0] = 0, disp[t+1] = disp[t] + f64(vel[t]) sequentially.  mfs.py:271, 281."""
    vel = np.asarray(vel)
    out = np.zeros((vel.shape[0] + 1,) + vel.shape[1:], dtype=np.float64)
    for t in range(vel.shape[0]):
        out[t + 1] = out[t] + vel[t].astype(np.float64)
    return out


# --------------------------------------------------------------------------------------------
# A.5 Jacobi (banded restatement of the dense reference system)
# --------------------------------------------------------------------------------------------
def adaptive_lambda(homs, W, H, definition):
    """Closed-form lambda_t (A.5).  mfs.py:786-841."""
    homs = np.asarray(homs, dtype=np.float64).reshape(-1, 3, 3)
    F = homs.shape[0]
    if definition == 2:
        return np.full(F, 100.0)
    if definition == 3:
        return np.full(F, 1.0)
    a, b, tx = homs[:, 0, 0], homs[:, 0, 1], homs[:, 0, 2]
    c, d, ty = homs[:, 1, 0], homs[:, 1, 1], homs[:, 1, 2]
    tr = a + d
    det = a * d - b * c
    disc = tr * tr - 4.0 * det
    sq = np.sqrt(np.abs(disc))
    m1 = np.where(disc >= 0, np.abs((tr + sq) / 2.0), np.sqrt(np.abs(det)))
    m2 = np.where(disc >= 0, np.abs((tr - sq) / 2.0), np.sqrt(np.abs(det)))
    mags = np.sort(np.stack([np.ones(F), m1, m2], axis=1), axis=1)
    ratio = mags[:, 1] / mags[:, 2]
    trans = np.sqrt((tx / W) ** 2 + (ty / H) ** 2)
    c1 = -1.93 * trans + 0.95
    c2 = 5.83 * ratio + 4.88 if definition == 0 else 5.83 * ratio - 4.88
    return np.maximum(np.minimum(c1, c2), 0.0)


def jacobi_banded(u, homs, W, H, radius, iters, definition):
    """x^{n+1}_t = (b_t + 2 lam_t sum_{|k|<=radius} w_k x^n_{t+k}) / diag_t, k = 0 included,
    diag_t = 1 + 2 lam_t sum_{r=0}^{F-1} w_{t-r}.  u: (F, ..., 2) float64."""
    u = np.asarray(u, dtype=np.float64)
    F = u.shape[0]
    lam = adaptive_lambda(homs, W, H, definition)
    k = np.arange(-radius, radius + 1)
    w = np.exp(-np.square((3.0 / radius) * k))
    t = np.arange(F)
    full = np.exp(-np.square((3.0 / radius) * (t[:, None] - t[None, :])))
    diag = 1.0 + 2.0 * lam * full.sum(axis=1)
    b = u.reshape(F, -1)
    x = b.copy()
    shape = (F,) + (1,) * (b.ndim - 1)
    for _ in range(iters):
        acc = np.zeros_like(x)
        for kk, wk in zip(k, w):
            lo, hi = max(0, -kk), min(F, F - kk)
            if hi > lo:                                   # F <= |k|: no frame has this neighbour
                acc[lo:hi] += wk * x[lo + kk:hi + kk]
        x = (b + (2.0 * lam).reshape(shape) * acc) / diag.reshape(shape)
    return x.reshape(u.shape)


# --------------------------------------------------------------------------------------------
# A.2 warp maps + crop edges
# --------------------------------------------------------------------------------------------
def cell_setup(rest_xy, delta, R, C):
    """Per-cell quantities of one frame.  rest_xy: (V,2) float32, delta = s - u: (V,2) float64.
    Returns dict with Hsu (n,9), Mi (n,9), lo_x, hi_x, lo_y, hi_y (1/32-px integer bounds) and the
    stabilized corners.  mfs.py:1025-1048."""
    rest64 = rest_xy.astype(np.float64)
    stab = (rest64 + delta).astype(np.float32).astype(np.float64)        # findHomography rounds to f32
    rest_rc = rest64.reshape(R + 1, C + 1, 2)
    stab_rc = stab.reshape(R + 1, C + 1, 2)
    src = np.empty((R * C, 4, 2)); dst = np.empty((R * C, 4, 2))
    n = 0
    for r in range(R):
        for c in range(C):
            src[n] = rest_rc[r:r + 2, c:c + 2].reshape(4, 2)
            dst[n] = stab_rc[r:r + 2, c:c + 2].reshape(4, 2)
            n += 1
    Hus = homography_4pt(src, dst)
    Hsu = homography_4pt(dst, src)
    Mi = inverse3x3_adjugate(Hus)
    L = np.floor(src[:, :, 0].min(axis=1)).astype(np.int64)
    Rr = np.ceil(src[:, :, 0].max(axis=1)).astype(np.int64)
    T = np.floor(src[:, :, 1].min(axis=1)).astype(np.int64)
    B = np.ceil(src[:, :, 1].max(axis=1)).astype(np.int64)
    return dict(Hus=Hus, Hsu=Hsu, Mi=Mi, lo_x=32 * L - 31, hi_x=32 * Rr + 31,
                lo_y=32 * T - 31, hi_y=32 * B + 31, dst=dst, src=src)


def _rint_sat(v):
    # cvRound on x86 (cvtsd2si) turns NaN into INT_MIN, which no membership bound contains
    nan = np.isnan(v)
    out = np.rint(np.clip(np.where(nan, 0.0, v), INT_MIN, INT_MAX)).astype(np.int64)
    return np.where(nan, INT_MIN, out)


def cell_inside(cells, n, xs, ys):
    """cv2.warpPerspective(rect mask) != 0 at output pixels (xs, ys) for cell n (A.2)."""
    Mi = cells["Mi"][n]
    Wd = Mi[6] * xs + Mi[7] * ys + Mi[8]
    with np.errstate(divide="ignore", invalid="ignore"):
        Wd = np.where(Wd != 0, 32.0 / Wd, 0.0)
        X = _rint_sat((Mi[0] * xs + Mi[1] * ys + Mi[2]) * Wd)
        Y = _rint_sat((Mi[3] * xs + Mi[4] * ys + Mi[5]) * Wd)
    return ((X >= cells["lo_x"][n]) & (X <= cells["hi_x"][n]) &
            (Y >= cells["lo_y"][n]) & (Y <= cells["hi_y"][n]))


def warp_maps(W, H, cells, prune=True):
    """float32 (map_x, map_y, cell_id) of one frame: the inside cell with the largest id wins,
    uncovered pixels keep (W+1, H+1).  mfs.py:1050-1061."""
    map_x = np.full((H, W), W + 1, dtype=np.float32)
    map_y = np.full((H, W), H + 1, dtype=np.float32)
    cell_id = np.full((H, W), -1, dtype=np.int32)
    ncell = cells["Hsu"].shape[0]
    for n in range(ncell):
        x0, x1, y0, y1 = 0, W - 1, 0, H - 1
        mild = False
        if prune:
            # the quad's bounding box only bounds the support while the cell's projective
            # denominator keeps its sign and hardly varies over the rest rectangle (no fold, no
            # horizon crossing)
            hus = cells["Hus"][n]
            wq = cells["src"][n][:, 0] * hus[6] + cells["src"][n][:, 1] * hus[7] + hus[8]
            aw = np.abs(wq)
            mild = bool(np.all(np.isfinite(wq)) and (np.all(wq > 0) or np.all(wq < 0)) and aw.max() < 1.5 * aw.min())
        if mild:
            q = cells["dst"][n]
            x0 = max(0, int(math.floor(q[:, 0].min())) - 3); x1 = min(W - 1, int(math.ceil(q[:, 0].max())) + 3)
            y0 = max(0, int(math.floor(q[:, 1].min())) - 3); y1 = min(H - 1, int(math.ceil(q[:, 1].max())) + 3)
            if x0 > x1 or y0 > y1:
                continue
        ys, xs = np.mgrid[y0:y1 + 1, x0:x1 + 1].astype(np.float64)
        ins = cell_inside(cells, n, xs, ys)
        mx, my = persp_f64(xs, ys, cells["Hsu"][n])
        sub = (slice(y0, y1 + 1), slice(x0, x1 + 1))
        map_x[sub] = np.where(ins, mx.astype(np.float32), map_x[sub])
        map_y[sub] = np.where(ins, my.astype(np.float32), map_y[sub])
        cell_id[sub] = np.where(ins, n, cell_id[sub])
    return map_x, map_y, cell_id


def crop_edges(map_x, map_y):
    """Per-frame (left, top, right, bottom).  mfs.py:1075-1098 (defaults 0, 0, W-1, H-1)."""
    H, W = map_x.shape
    mx = map_x.astype(np.float64); my = map_y.astype(np.float64)
    left, top, right, bottom = 0, 0, W - 1, H - 1
    hit = np.where(np.abs(mx) < 1)[1]
    if hit.size: left = int(hit.max())
    hit = np.where(np.abs(mx - (W - 1)) < 1)[1]
    if hit.size: right = int(hit.min())
    hit = np.where(np.abs(my) < 1)[0]
    if hit.size: top = int(hit.max())
    hit = np.where(np.abs(my - (H - 1)) < 1)[0]
    if hit.size: bottom = int(hit.min())
    return left, top, right, bottom


# --------------------------------------------------------------------------------------------
# A.3 cv2.remap 8UC3 INTER_LINEAR BORDER_CONSTANT (fixed point, 1/32 px)
# --------------------------------------------------------------------------------------------
def remap_fixed(src, map_x, map_y, border):
    H, W = src.shape[:2]
    SX = _rint_sat(map_x.astype(np.float32).astype(np.float64) * 32.0)
    SY = _rint_sat(map_y.astype(np.float32).astype(np.float64) * 32.0)
    ix, iy, ax, ay = SX >> 5, SY >> 5, SX & 31, SY & 31
    bord = np.asarray(border, dtype=np.int64).reshape(1, 1, 3)

    def tap(i, j):
        ok = (i >= 0) & (i < W) & (j >= 0) & (j < H)
        v = src[np.clip(j, 0, H - 1), np.clip(i, 0, W - 1)].astype(np.int64)
        return np.where(ok[..., None], v, bord)

    ax = ax[..., None]; ay = ay[..., None]
    acc = (tap(ix, iy) * (32 - ax) * (32 - ay) + tap(ix + 1, iy) * ax * (32 - ay) +
           tap(ix, iy + 1) * (32 - ax) * ay + tap(ix + 1, iy + 1) * ax * ay + 512) >> 10
    return acc.astype(np.uint8)


# --------------------------------------------------------------------------------------------
# A.4 cv2.resize 8UC3 INTER_LINEAR (fixed point, 11 bits)
# --------------------------------------------------------------------------------------------
def resize_tables(src_len, dst_len):
    """(index0, index1, coef0, coef1) of cv2.resize's linear pass along one axis, clamped like the x
    axis (the y axis is *not* clamped: see ``resize_fixed``)."""
    scale = np.float64(src_len) / np.float64(dst_len)
    d = np.arange(dst_len)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_fixed(src, W, H):
    sh, sw = src.shape[:2]
    sx, fx = resize_tables(sw, W)
    lo = sx < 0
    fx = np.where(lo, np.float32(0), fx); sx = np.where(lo, 0, sx)
    hi = sx >= sw - 1
    fx = np.where(hi, np.float32(0), fx); sx = np.where(hi, sw - 1, sx)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int64)
    x0 = sx; x1 = np.minimum(sx + 1, sw - 1)
    sy, fy = resize_tables(sh, H)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int64)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int64)
    r0 = np.clip(sy, 0, sh - 1); r1 = np.clip(sy + 1, 0, sh - 1)
    s = src.astype(np.int64)
    S0 = a0[None, :, None] * s[r0][:, x0] + a1[None, :, None] * s[r0][:, x1]
    S1 = a0[None, :, None] * s[r1][:, x0] + a1[None, :, None] * s[r1][:, x1]
    out = (((b0[:, None, None] * (S0 >> 4)) >> 16) + ((b1[:, None, None] * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# whole warp stage from the per-element pieces
# --------------------------------------------------------------------------------------------
def warp_stage(frames, u, s, R, C, border, return_maps=False):
    F = len(frames)
    H, W = frames[0].shape[:2]
    rest = vertex_xy(W, H, R, C)
    delta = (np.asarray(s) - np.asarray(u)).reshape(F, -1, 2)
    out, per_frame, maps = [], [], []
    for f in range(F):
        cells = cell_setup(rest, delta[f], R, C)
        mx, my, cid = warp_maps(W, H, cells)
        out.append(remap_fixed(frames[f], mx, my, border))
        per_frame.append(crop_edges(mx, my))
        if return_maps:
            maps.append((mx, my, cid))
    pf = np.asarray(per_frame, dtype=np.int64).reshape(F, 4)
    crop = (int(pf[:, 0].max()), int(pf[:, 1].max()), int(pf[:, 2].min()), int(pf[:, 3].min()))
    if return_maps:
        return out, crop, maps, pf
    return out, crop


def crop_stage(frames, crop):
    H, W = frames[0].shape[:2]
    l, t, r, b = [int(v) for v in crop]
    return [resize_fixed(f[t:b + 1, l:r + 1], W, H) for f in frames]


# --------------------------------------------------------------------------------------------
# A.6 stability score
# --------------------------------------------------------------------------------------------
def stability_score(s):
    """Direct DFT bins 1..5 + Parseval total.  mfs.py:1216-1259."""
    s = np.asarray(s, dtype=np.float64)
    F = s.shape[0]
    p = np.diff(s.reshape(F, -1, 2), axis=0)                      # (F-1, V, 2)
    n = F - 1
    t = np.arange(n)
    low = np.zeros(p.shape[1:])
    for k in range(1, 6):
        ang = -2.0 * np.pi * k * t / n
        re = np.tensordot(np.cos(ang), p, axes=(0, 0))
        im = np.tensordot(np.sin(ang), p, axes=(0, 0))
        low += re * re + im * im
    total = n * np.sum(p * p, axis=0)
    ratio = low / total
    return (ratio[:, 0].mean() + ratio[:, 1].mean()) / 2.0
