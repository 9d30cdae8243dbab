"""CPU tests of the kernels' per-element arithmetic (meshflow_b200/csrc/mf_math.cuh) driven through the
test-only host-emulation harness tests/hostemu (same source, g++ -ffp-contract=off) against the oracle.
The GPU tests cover the kernels' parallel plumbing; these catch arithmetic drift without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import spec
from tests import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "hostemu", "libmf_hostemu.so")


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "hostemu", "hostemu.cpp")
    hdr = os.path.join(HERE, "..", "meshflow_b200", "csrc", "mf_math.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", "-o", SO, src], check=True)
    lib = ctypes.CDLL(SO)
    lib.emu_median9.restype = ctypes.c_float
    lib.emu_lambda.restype = ctypes.c_double
    return lib


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("W,H,R,C,er,ec", [(640, 360, 16, 16, 10, 10), (1920, 1080, 16, 16, 10, 10),
                                           (1280, 720, 64, 64, 10, 10), (640, 360, 8, 12, 6, 9)])
def test_feature_to_vertex_membership_is_bit_exact(emu, W, H, R, C, er, ec):
    rng = np.random.default_rng(W + R)
    n = 4000
    fx = rng.uniform(-5, W + 5, n); fy = rng.uniform(-5, H + 5, n)
    fx[:60] = np.arange(60) * W / C / 2; fy[:60] = np.arange(60) * H / R / 2      # exactly on vertex lines
    top = np.zeros(n, np.int32); bot = np.zeros(n, np.int32)
    l = np.zeros((n, R + 1), np.int32); r = np.zeros((n, R + 1), np.int32)
    emu.emu_feature_ranges(P(fx), P(fy), n, W, H, R, C, er, ec, P(top), P(bot), P(l), P(r))
    _, _, l2, r2 = spec.feature_vertex_ranges(fx, fy, W, H, R, C, er, ec)
    cols = np.arange(C + 1)
    m1 = (cols[None, None, :] >= l[:, :, None]) & (cols[None, None, :] <= r[:, :, None])
    m2 = (cols[None, None, :] >= l2[:, :, None]) & (cols[None, None, :] <= r2[:, :, None])
    assert np.array_equal(m1, m2)


def test_kernel_perspective_transform_equals_opencv_bit_for_bit(emu):
    """mf_math.cuh persp() -- the arithmetic both vertex-motion kernels run -- against cv2.perspectiveTransform on
    float64 points: every bit (OpenCV's FMA contraction included)."""
    import cv2
    rng = np.random.default_rng(5)
    for k in range(4):
        M = synth.random_homography(rng, 1920, 1080) if k < 2 else np.eye(3) + rng.normal(0, 0.3, (3, 3))
        n = 20000
        x = rng.uniform(-50, 2000, n); y = rng.uniform(-50, 1200, n)
        if k == 1:
            x, y = np.round(x), np.round(y)
        ox = np.zeros(n); oy = np.zeros(n)
        emu.emu_persp(P(np.ascontiguousarray(M.reshape(-1))), P(x), P(y), n, P(ox), P(oy))
        ref = cv2.perspectiveTransform(np.stack([x, y], axis=1).reshape(-1, 1, 2), M).reshape(-1, 2)
        assert np.array_equal(ox, ref[:, 0]) and np.array_equal(oy, ref[:, 1])


def test_key_transform_preserves_order_and_roundtrips(emu):
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.normal(0, 1, 500) * 10.0 ** rng.integers(-300, 300, 500), [0.0, 1e-320, -1e-320]])   # (-0.0 sorts below +0.0: harmless, both are zero)
    keys = np.zeros(v.size, np.uint64); back = np.zeros(v.size)
    emu.emu_key_roundtrip(P(v), v.size, P(keys), P(back))
    assert np.array_equal(back.view(np.uint64), v.view(np.uint64))
    order = np.argsort(v, kind="stable")
    assert np.all(np.diff(keys[order].astype(object)) >= 0)


def test_median9_matches_numpy(emu):
    rng = np.random.default_rng(2)
    for _ in range(3000):
        v = rng.normal(size=9).astype(np.float32)
        if rng.random() < 0.3:
            v[rng.integers(0, 9, 3)] = v[0]
        assert emu.emu_median9(P(v)) == np.float32(np.median(v))


def test_adaptive_lambda_closed_form(emu):
    rng = np.random.default_rng(3)
    for _ in range(500):
        Hm = synth.random_homography(rng, 640, 360, rot=0.2, scale=0.1, trans=40.0)
        for d in range(4):
            assert abs(emu.emu_lambda(P(Hm), 640, 360, d) - spec.adaptive_lambda(Hm[None], 640, 360, d)[0]) <= 1e-15


@pytest.mark.parametrize("W,H,R,C,amp,seed", [(320, 180, 8, 8, 2.5, 1), (200, 120, 4, 6, 15.0, 3), (256, 144, 16, 16, 1.0, 5),
                                              (640, 360, 16, 16, 3.0, 7)])
def test_warp_arithmetic_matches_spec(emu, W, H, R, C, amp, seed):
    csz = emu.emu_sizeof_cell()
    assert csz == 240
    rng = np.random.default_rng(seed)
    frames, u, s = synth.synthetic_warp_inputs(rng, 1, W, H, R, C, per_vertex=amp, per_frame=1.2 * amp)
    rest = spec.vertex_xy(W, H, R, C)
    uu = np.ascontiguousarray(u[0].reshape(-1, 2)); ss = np.ascontiguousarray(s[0].reshape(-1, 2))
    cells = np.zeros(R * C * csz, np.uint8)
    emu.emu_cell_setup(P(rest), P(uu), P(ss), W, H, R, C, P(cells))
    dst = np.zeros((H, W, 3), np.uint8); maps = np.zeros((H, W, 2), np.float32); crop = np.zeros(4, np.int32)
    src = np.ascontiguousarray(frames[0])
    emu.emu_warp_frame(P(src), P(cells), R * C, W, H, 9, 8, 7, P(dst), P(maps), P(crop), 1)
    sc = spec.cell_setup(rest, ss - uu, R, C)
    mx, my, _ = spec.warp_maps(W, H, sc, prune=False)
    assert np.array_equal(maps[..., 0], mx) and np.array_equal(maps[..., 1], my)
    assert np.array_equal(dst, spec.remap_fixed(src, mx, my, (9, 8, 7)))
    assert tuple(crop.tolist()) == spec.crop_edges(mx, my)
    audit = np.zeros(4, np.int64)
    emu.emu_screen_audit(P(cells), R * C, W, H, P(audit))
    assert audit[2] == 0, "float32 screen contradicted the float64 membership test"
    if amp < 5:
        assert audit[3] >= 0.6 * R * C and audit[1] < 0.4 * audit[0]     # mild meshes: mostly screened
    if crop[0] <= crop[2] and crop[1] <= crop[3]:
        out = np.zeros_like(dst)
        l, t, r, b = (int(v) for v in crop)
        emu.emu_crop_resize(P(dst), W, H, l, t, r, b, P(out))
        assert np.array_equal(out, spec.resize_fixed(dst[t:b + 1, l:r + 1], W, H))


# ----------------------------------------------------------------------------------------------
# fast path of the warp (csrc/warp_fast.cuh): exact row spans + float32 coordinates outside the band
# ----------------------------------------------------------------------------------------------
FAST_CASES = [(320, 180, 8, 8, 2.5, 1), (200, 120, 4, 6, 15.0, 3), (256, 144, 16, 16, 1.0, 5), (640, 360, 16, 16, 3.0, 7),
              (333, 217, 6, 9, 2.0, 9), (640, 360, 40, 40, 0.8, 11), (160, 96, 48, 80, 0.15, 13)]


@pytest.mark.parametrize("W,H,R,C,amp,seed", FAST_CASES)
def test_row_spans_equal_the_membership_test_pixel_for_pixel(emu, W, H, R, C, amp, seed):
    rng = np.random.default_rng(seed)
    _, u, s = synth.synthetic_warp_inputs(rng, 1, W, H, R, C, per_vertex=amp, per_frame=1.2 * amp)
    rest = spec.vertex_xy(W, H, R, C)
    uu = np.ascontiguousarray(u[0].reshape(-1, 2)); ss = np.ascontiguousarray(s[0].reshape(-1, 2))
    audit = np.zeros(4, np.int64)
    emu.emu_span_audit(P(rest), P(uu), P(ss), W, H, R, C, P(audit))
    assert audit[2] == 0, "a row span disagrees with cell_inside"
    assert audit[1] == 0 or amp > 5
    if amp < 5 and W // C >= 8:
        assert audit[3] >= 0.95 * R * C


@pytest.mark.parametrize("W,H,R,C,amp,seed", FAST_CASES)
def test_fast_warp_path_matches_spec(emu, W, H, R, C, amp, seed):
    rng = np.random.default_rng(seed)
    frames, u, s = synth.synthetic_warp_inputs(rng, 1, W, H, R, C, per_vertex=amp, per_frame=1.2 * amp)
    rest = spec.vertex_xy(W, H, R, C)
    uu = np.ascontiguousarray(u[0].reshape(-1, 2)); ss = np.ascontiguousarray(s[0].reshape(-1, 2))
    src = np.ascontiguousarray(frames[0])
    dst = np.zeros((H, W, 3), np.uint8); crop = np.zeros(4, np.int32); stats = np.zeros(6, np.int64)
    emu.emu_warp_frame_fast(P(src), P(rest), P(uu), P(ss), W, H, R, C, 9, 8, 7, P(dst), P(crop), 0, P(stats))
    sc = spec.cell_setup(rest, ss - uu, R, C)
    mx, my, _ = spec.warp_maps(W, H, sc, prune=False)
    ref = spec.remap_fixed(src, mx, my, (9, 8, 7))
    assert np.array_equal(dst, ref), f"{(dst != ref).any(axis=2).sum()} pixels differ; stats {stats.tolist()}"
    assert tuple(crop.tolist()) == spec.crop_edges(mx, my)
    crop2 = np.zeros(4, np.int32); stats2 = np.zeros(6, np.int64)
    emu.emu_warp_frame_fast(P(src), P(rest), P(uu), P(ss), W, H, R, C, 9, 8, 7, None, P(crop2), 1, P(stats2))
    assert crop2.tolist() == crop.tolist()
    if amp < 3 and W // C >= 8:
        assert stats[3] <= 0.01 * stats[4]           # (almost) no irregular row segments on mild meshes
    if W // C < 4:
        assert stats[3] == stats[4]                  # 2 x 2 px cells: every candidate list overflows, all rows exact


def test_fast_warp_path_takes_most_pixels_of_a_smooth_warp(emu):
    """Camera-like warp (small rotation / scale / translation, as in configs[1]): nearly every group of
    four pixels has adjacent footprints and takes the shared-window gather."""
    W, H, R, C = 640, 360, 16, 16
    rng = np.random.default_rng(33)
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    rest = spec.vertex_xy(W, H, R, C).astype(np.float64)
    Hm = synth.random_homography(rng, W, H, rot=0.004, scale=0.003, trans=4.0)
    w = rest[:, 0] * Hm[2, 0] + rest[:, 1] * Hm[2, 1] + 1.0
    moved = np.stack([(rest[:, 0] * Hm[0, 0] + rest[:, 1] * Hm[0, 1] + Hm[0, 2]) / w,
                      (rest[:, 0] * Hm[1, 0] + rest[:, 1] * Hm[1, 1] + Hm[1, 2]) / w], axis=1)
    d = np.ascontiguousarray(moved - rest + rng.normal(0, 0.05, rest.shape))
    zero = np.zeros_like(d)
    dst = np.zeros_like(src); crop = np.zeros(4, np.int32); stats = np.zeros(6, np.int64)
    emu.emu_warp_frame_fast(P(src), P(spec.vertex_xy(W, H, R, C)), P(zero), P(d), W, H, R, C, 0, 0, 255, P(dst), P(crop), 0, P(stats))
    sc = spec.cell_setup(spec.vertex_xy(W, H, R, C), d, R, C)
    mx, my, _ = spec.warp_maps(W, H, sc, prune=False)
    assert np.array_equal(dst, spec.remap_fixed(src, mx, my, (0, 0, 255)))
    assert tuple(crop.tolist()) == spec.crop_edges(mx, my)
    assert stats[0] > 0.85 * W * H, f"fast path took only {stats[0]} of {W * H} pixels: {stats.tolist()}"


def test_fast_warp_path_identity_and_folded_mesh(emu):
    W, H, R, C = 256, 144, 8, 8
    rng = np.random.default_rng(21)
    src = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    rest = spec.vertex_xy(W, H, R, C)
    zero = np.zeros(((R + 1) * (C + 1), 2))
    dst = np.zeros_like(src); crop = np.zeros(4, np.int32); stats = np.zeros(6, np.int64)
    emu.emu_warp_frame_fast(P(src), P(rest), P(zero), P(zero), W, H, R, C, 0, 0, 255, P(dst), P(crop), 0, P(stats))
    assert np.array_equal(dst, src) and crop.tolist() == [0, 0, W - 1, H - 1]
    # a folded mesh: one vertex pulled across its neighbours (cells overlap, some are not convex)
    d = rng.normal(0, 2.0, ((R + 1) * (C + 1), 2))
    d[4 * (C + 1) + 4] += (70.0, 45.0)
    emu.emu_warp_frame_fast(P(src), P(rest), P(zero), P(np.ascontiguousarray(d)), W, H, R, C, 0, 0, 255, P(dst), P(crop), 0, P(stats))
    sc = spec.cell_setup(rest, d, R, C)
    mx, my, _ = spec.warp_maps(W, H, sc, prune=False)
    assert np.array_equal(dst, spec.remap_fixed(src, mx, my, (0, 0, 255)))
    assert tuple(crop.tolist()) == spec.crop_edges(mx, my)


def _crop_by_segments(emu, W, H, R, C, d):
    rest = spec.vertex_xy(W, H, R, C)
    zero = np.zeros_like(d)
    crop = np.zeros(4, np.int32); stats = np.zeros(6, np.int64)
    emu.emu_warp_frame_fast(None, P(rest), P(zero), P(np.ascontiguousarray(d)), W, H, R, C, 0, 0, 255, None, P(crop), 1, P(stats))
    return crop.tolist()


def _crop_by_scan(W, H, R, C, d):
    sc = spec.cell_setup(spec.vertex_xy(W, H, R, C), d, R, C)
    # every cell over the whole frame on the small geometries; the 640x360 mesh (256 cells x 230 k pixels, 11 s per
    # case that way) goes through the spec's bounding-box pruning, itself checked against the cv2 port in test_oracle
    mx, my, _ = spec.warp_maps(W, H, sc, prune=W * H > 100000)
    return list(spec.crop_edges(mx, my))


@pytest.mark.parametrize("W,H,R,C", [(320, 180, 8, 8), (640, 360, 16, 16), (333, 217, 6, 9)])
def test_crop_edges_from_row_segments_equal_the_pixel_scan(emu, W, H, R, C):
    """The closed-form band search on row segments (crop_edges_kernel) against the reference's scan of the
    float32 maps (mfs.py:1075-1098): pure translations (the map is constant along a row: zero slope), integer
    and half-pixel shifts that put map values exactly ON a band limit, large shifts in all four directions,
    rotation / scale / perspective, rough per-vertex noise."""
    rng = np.random.default_rng(W + R)
    V = (R + 1) * (C + 1)
    rest = spec.vertex_xy(W, H, R, C).astype(np.float64)
    cases = []
    for shift in [(0.0, 0.0), (7.0, -5.0), (-3.0, 4.0), (0.5, 0.5), (-0.5, 1.0), (1.0, -1.0), (12.25, 9.75), (-17.5, -6.125),
                  (0.96875, -0.96875), (1.03125, 0.984375)]:
        cases.append(np.tile(np.array(shift), (V, 1)))
    for k in range(6):
        Hm = synth.random_homography(rng, W, H, rot=0.01 * (k + 1), scale=0.01 * (k + 1), trans=3.0 * (k + 1), persp=3e-6 * k)
        w = rest[:, 0] * Hm[2, 0] + rest[:, 1] * Hm[2, 1] + 1.0
        moved = np.stack([(rest[:, 0] * Hm[0, 0] + rest[:, 1] * Hm[0, 1] + Hm[0, 2]) / w,
                          (rest[:, 0] * Hm[1, 0] + rest[:, 1] * Hm[1, 1] + Hm[1, 2]) / w], axis=1)
        cases.append(moved - rest + rng.normal(0, 0.05 * k, rest.shape))
    for amp in (1.0, 4.0, 12.0):
        cases.append(rng.normal(0, amp, (V, 2)) + rng.normal(0, 2 * amp, (1, 2)))
    for d in cases:
        d = np.ascontiguousarray(d, dtype=np.float64)
        assert _crop_by_segments(emu, W, H, R, C, d) == _crop_by_scan(W, H, R, C, d)


def test_segment_resolution_matches_brute_force(emu):
    """'The last cell written wins': random overlapping intervals in priority order against a per-pixel loop."""
    rng = np.random.default_rng(5)
    overflowed = 0
    for trial in range(1200):
        # whole tiles and the narrower last tile of a row (1 .. 128 pixels), word boundaries of the 128-bit mask included
        x0 = 128 * int(rng.integers(0, 30))
        x1 = x0 + (127 if trial % 3 else int(rng.choice([0, 1, 30, 31, 32, 33, 63, 64, 65, 95, 96, 97, 126, int(rng.integers(0, 128))])))
        n = int(rng.integers(0, 12))
        a = rng.integers(x0 - 30, x1 + 10, n).astype(np.int32)
        b = (a + rng.integers(-3, 90, n)).astype(np.int32)
        ids = np.sort(rng.choice(5000, n, replace=False))[::-1].astype(np.int32)      # descending id = priority
        cap = 8 if trial % 2 else 16
        seg = np.zeros(16, np.uint32); opx = np.zeros(128, np.uint32); ogr = np.zeros(32, np.uint32)
        ns = emu.emu_resolve_segments(x0, x1, n, P(a), P(b), P(ids), cap, P(seg), P(opx), P(ogr))
        npx = x1 - x0 + 1
        ref = np.full(npx, 0xFFFF, np.uint32)
        for k in range(n - 1, -1, -1):                       # lowest priority first, higher ones overwrite
            lo, hi = max(int(a[k]), x0), min(int(b[k]), x1)
            if lo <= hi:
                ref[lo - x0:hi - x0 + 1] = ids[k]
        changes = 1 + int((np.diff(ref.astype(np.int64)) != 0).sum())
        if ns < 0:
            overflowed += 1
            assert changes > cap // 2                        # only gives up when there really are many segments
            continue
        assert ns == changes                                 # one segment per run of equal owners
        assert np.array_equal(opx[:npx], ref)
        for g in range((npx + 3) // 4):
            grp = ref[4 * g:4 * g + 4]
            want = 0xFFFD if (grp != grp[0]).any() else grp[0]
            assert ogr[g] == want
        assert (np.diff((seg[:ns] >> 16).astype(np.int64)) > 0).all() and (seg[ns:] == 0xFFFFFFFF).all()
        assert seg[0] >> 16 == x0
    assert overflowed < 600


@pytest.mark.parametrize("W,H,R,C", [(1920, 1080, 16, 16), (3840, 2160, 16, 16), (1280, 720, 64, 64), (1920, 1080, 32, 32)])
def test_float32_coordinates_never_disagree_outside_the_band(emu, W, H, R, C):
    """Every member pixel of every cell, camera-like and rough meshes: a pixel the float32 path calls safe has the
    reference's 1/32-px coordinate, and the observed float32 error stays well inside the band the cell reserves."""
    rng = np.random.default_rng(W + R)
    rest = spec.vertex_xy(W, H, R, C)
    rest64 = rest.astype(np.float64)
    total = 0
    for trial in range(3):
        Hm = synth.random_homography(rng, W, H, rot=0.004 * (1 + 4 * trial), scale=0.003 * (1 + 4 * trial), trans=6.0 * (1 + trial),
                                     persp=2e-6 * (1 + trial))
        w = rest64[:, 0] * Hm[2, 0] + rest64[:, 1] * Hm[2, 1] + 1.0
        d = np.stack([(rest64[:, 0] * Hm[0, 0] + rest64[:, 1] * Hm[0, 1] + Hm[0, 2]) / w - rest64[:, 0],
                      (rest64[:, 0] * Hm[1, 0] + rest64[:, 1] * Hm[1, 1] + Hm[1, 2]) / w - rest64[:, 1]], axis=1)
        d = np.ascontiguousarray(d + rng.normal(0, 0.3 * trial, d.shape))
        zero = np.zeros_like(d)
        out = np.zeros(3, np.int64); ratio = np.zeros(1)
        emu.emu_fast_coord_audit(P(rest), P(zero), P(d), W, H, R, C, P(out), P(ratio))
        assert out[2] == 0, f"{out[2]} safe pixels with a wrong coordinate"
        assert ratio[0] < 1.0
        assert out[1] < 0.06 * out[0]                          # a few per cent of the pixels take the float64 path
        total += int(out[0])
    assert total > 2 * W * H
