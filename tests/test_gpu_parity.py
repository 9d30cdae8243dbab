"""GPU parity tests proper: every stage through the C ABI against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): bit-exact for feature->vertex assignment, masks, crop indices and
(because the fixed-point OpenCV models are reproduced) pixels and float32 maps; <= 1e-4 relative on
optimised vertex paths and metrics -- asserted here at 1e-9 since the kernels are float64.
"""
import numpy as np
import pytest

from oracle import spec
from tests import synth

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _core(W, H, R=16, C=16, **kw):
    from meshflow_b200 import DeviceCore, MeshSpec
    return DeviceCore(MeshSpec(W, H, R, C), **kw)


def _dev(a, core):
    return torch.from_numpy(np.ascontiguousarray(a)).to(core.device)


# ----------------------------------------------------------------------------------------------
# (1) vertex motion
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", ["fast", "generic"])
@pytest.mark.parametrize("W,H,R,C,n,P", [(640, 360, 16, 16, 1800, 5), (1920, 1080, 16, 16, 3000, 3),
                                         (320, 180, 8, 12, 400, 4), (640, 360, 16, 16, 3, 3),
                                         (1280, 720, 64, 64, 2500, 2), (1920, 1080, 16, 16, 13000, 2),
                                         (640, 360, 40, 40, 700, 2)])
def test_vertex_motion_matches_oracle(W, H, R, C, n, P, path):
    """fast = sort + bit-matrix selection (needs the host copy of pair_start); generic = per-vertex
    radix select.  Both must reproduce the oracle's float32 velocities and assignment counts exactly."""
    rng = np.random.default_rng(100 + n)
    tr = synth.synthetic_tracks(rng, P, n, W, H)
    core = _core(W, H, R, C)
    vel, counts = core.vertex_velocities(_dev(tr["early"], core), _dev(tr["late"], core), _dev(tr["offset"], core),
                                         _dev(tr["keep"], core), _dev(tr["pair_start"], core),
                                         _dev(tr["homographies"].reshape(-1, 9), core),
                                         pair_start_host=tr["pair_start"] if path == "fast" else None,
                                         return_counts=True)
    vel = vel.cpu().numpy(); counts = counts.cpu().numpy()
    for p in range(P):
        a, b = tr["pair_start"][p], tr["pair_start"][p + 1]
        k = tr["keep"][a:b].astype(bool)
        off = tr["offset"][a:b][k].astype(np.float64)
        e = tr["early"][a:b][k].astype(np.float64) + off
        l = tr["late"][a:b][k].astype(np.float64) + off
        ref, member = spec.vertex_velocities(e, l, tr["homographies"][p], W, H, R, C, 10, 10, return_assignment=True)
        assert np.array_equal(counts[p], member.sum(axis=0)), "feature->vertex assignment differs"
        assert np.array_equal(vel[p], ref), f"pair {p}: max |d| = {np.abs(vel[p] - ref).max()}"


def test_vertex_motion_with_duplicate_values_and_ties():
    """Many identical residuals (sort ties, equal 48-bit key prefixes) must still give the exact median."""
    W, H, R, C = 640, 360, 8, 8
    rng = np.random.default_rng(21)
    tr = synth.synthetic_tracks(rng, 2, 900, W, H, keep_prob=1.0)
    tr["late"] = (tr["early"] + np.round(rng.normal(0, 1.0, tr["early"].shape) * 4) / 4).astype(np.float32)
    tr["homographies"][:] = np.eye(3)
    tr["late"][::7] = tr["late"][::7] + np.float32(2.0 ** -18)          # differ only far below the 48-bit prefix
    core = _core(W, H, R, C)
    outs = []
    for host in (tr["pair_start"], None):
        outs.append(core.vertex_velocities(_dev(tr["early"], core), _dev(tr["late"], core), _dev(tr["offset"], core),
                                           _dev(tr["keep"], core), _dev(tr["pair_start"], core),
                                           _dev(tr["homographies"].reshape(-1, 9), core), pair_start_host=host).cpu().numpy())
    for p in range(2):
        a, b = tr["pair_start"][p], tr["pair_start"][p + 1]
        off = tr["offset"][a:b].astype(np.float64)
        ref = spec.vertex_velocities(tr["early"][a:b].astype(np.float64) + off, tr["late"][a:b].astype(np.float64) + off,
                                     tr["homographies"][p], W, H, R, C, 10, 10)
        assert np.array_equal(outs[0][p], ref) and np.array_equal(outs[1][p], ref)


def test_vertex_motion_of_an_almost_static_pair():
    """Residuals late - H(early) of ~1e-6 px (videos/video-2, pair 0): the float32 velocity shows the last bit of the
    float64 perspective transform, which OpenCV evaluates with one fused multiply-add per sum (mf_math.cuh persp)."""
    W, H, R, C = 640, 360, 16, 16
    rng = np.random.default_rng(32)
    tr = synth.synthetic_tracks(rng, 6, 5000, W, H, keep_prob=0.95, local_motion=1e-6,
                                homography=dict(rot=1e-7, scale=1e-7, trans=1e-4, persp=1e-10))
    core = _core(W, H, R, C)
    for host in (tr["pair_start"], None):
        vel = core.vertex_velocities(_dev(tr["early"], core), _dev(tr["late"], core), _dev(tr["offset"], core),
                                     _dev(tr["keep"], core), _dev(tr["pair_start"], core),
                                     _dev(tr["homographies"].reshape(-1, 9), core), pair_start_host=host).cpu().numpy()
        for p in range(6):
            a, b = tr["pair_start"][p], tr["pair_start"][p + 1]
            k = tr["keep"][a:b].astype(bool)
            off = tr["offset"][a:b][k].astype(np.float64)
            ref = spec.vertex_velocities(tr["early"][a:b][k].astype(np.float64) + off, tr["late"][a:b][k].astype(np.float64) + off,
                                         tr["homographies"][p], W, H, R, C, 10, 10)
            assert np.abs(ref).max() < 1e-3
            assert np.array_equal(vel[p].view(np.uint32), ref.view(np.uint32))


def test_vertex_motion_overflowing_candidate_list():
    """More candidates per vertex than the generic path's shared-memory list holds -> re-scan path."""
    W, H, R, C = 640, 360, 4, 4
    rng = np.random.default_rng(7)
    tr = synth.synthetic_tracks(rng, 2, 6000, W, H, keep_prob=1.0)
    core = _core(W, H, R, C)
    vel = core.vertex_velocities(_dev(tr["early"], core), _dev(tr["late"], core), _dev(tr["offset"], core),
                                 _dev(tr["keep"], core), _dev(tr["pair_start"], core),
                                 _dev(tr["homographies"].reshape(-1, 9), core), pair_start_host=None).cpu().numpy()
    for p in range(2):
        a, b = tr["pair_start"][p], tr["pair_start"][p + 1]
        off = tr["offset"][a:b].astype(np.float64)
        ref = spec.vertex_velocities(tr["early"][a:b].astype(np.float64) + off, tr["late"][a:b].astype(np.float64) + off,
                                     tr["homographies"][p], W, H, R, C, 10, 10)
        assert np.array_equal(vel[p], ref)


def test_prefix_sum_is_sequential_float64():
    rng = np.random.default_rng(3)
    vel = (rng.normal(0, 3, (257, 17, 17, 2)) * 10 ** rng.uniform(-3, 3, (257, 1, 1, 1))).astype(np.float32)
    core = _core(640, 360)
    disp = core.prefix_displacements(_dev(vel, core)).cpu().numpy()
    assert np.array_equal(disp, spec.prefix_displacements(vel))


# ----------------------------------------------------------------------------------------------
# (2) Jacobi
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,R,radius,iters", [(12, 4, 10, 100), (300, 16, 10, 100), (700, 6, 10, 60),
                                              (1500, 4, 7, 40), (2600, 2, 30, 25), (300, 4, 30, 50),
                                              (3100, 2, 10, 30), (2300, 2, 12, 20)])
@pytest.mark.parametrize("definition", [0, 1, 2, 3])
def test_jacobi_matches_oracle(F, R, radius, iters, definition):
    rng = np.random.default_rng(F + definition)
    u, homs = synth.synthetic_paths(rng, F, R, R)
    core = _core(640, 360, R, R, radius=radius, iterations=iters)
    s, lam = core.stabilized_displacements(_dev(u, core), _dev(homs, core), definition, return_lambda=True)
    s = s.cpu().numpy(); lam = lam.cpu().numpy()
    ref = spec.jacobi_banded(u, homs, 640, 360, radius, iters, definition)
    assert np.allclose(lam, spec.adaptive_lambda(homs, 640, 360, definition), rtol=1e-10, atol=1e-14)
    scale = np.abs(ref).max()
    assert np.abs(s - ref).max() <= 1e-9 * scale            # north-star tolerance: 1e-4 relative


def test_jacobi_vertex_shard_only_touches_its_range():
    rng = np.random.default_rng(11)
    u, homs = synth.synthetic_paths(rng, 100, 8, 8)
    core = _core(640, 360, 8, 8)
    full = core.stabilized_displacements(_dev(u, core), _dev(homs, core), 0).cpu().numpy()
    out = torch.full((100, 9, 9, 2), -7.0, dtype=torch.float64, device=core.device)
    core.stabilized_displacements(_dev(u, core), _dev(homs, core), 0, vertex_range=(20, 50), out=out)
    out = out.cpu().numpy().reshape(100, 81, 2)
    assert np.array_equal(out[:, 20:50], full.reshape(100, 81, 2)[:, 20:50])
    assert np.all(out[:, :20] == -7.0) and np.all(out[:, 50:] == -7.0)


@pytest.mark.parametrize("F,radius", [(2400, 10), (2400, 30), (4000, 10), (5100, 30)])
def test_jacobi_small_vertex_shard_of_a_long_video(F, radius):
    """A multi-GPU vertex shard of a long video: few systems, many frames.  The register-window kernel spreads the
    frames over more threads then (5 or 10 frames per thread instead of 20); the operations and their order are
    the same, so the shard equals the single-GPU solve of all vertices bit for bit -- and the oracle to 1e-9."""
    rng = np.random.default_rng(F + radius)
    R = 16
    u, homs = synth.synthetic_paths(rng, F, R, R)
    core = _core(640, 360, R, R, radius=radius, iterations=30)
    full = core.stabilized_displacements(_dev(u, core), _dev(homs, core), 1).cpu().numpy().reshape(F, -1, 2)
    out = torch.full((F, R + 1, R + 1, 2), -7.0, dtype=torch.float64, device=core.device)
    core.stabilized_displacements(_dev(u, core), _dev(homs, core), 1, vertex_range=(36, 73), out=out)
    out = out.cpu().numpy().reshape(F, -1, 2)
    assert np.array_equal(out[:, 36:73], full[:, 36:73])
    assert np.all(out[:, :36] == -7.0) and np.all(out[:, 73:] == -7.0)
    ref = spec.jacobi_banded(u[:, 2:3, 2:5], homs, 640, 360, radius, 30, 1)          # vertices 36..38
    assert np.abs(out[:, 36:39].reshape(ref.shape) - ref).max() <= 1e-9 * np.abs(ref).max()


def test_jacobi_rejects_bad_definition():
    from meshflow_b200._cabi import MeshflowNativeError
    core = _core(64, 64, 2, 2)
    u = torch.zeros((4, 3, 3, 2), dtype=torch.float64, device=core.device)
    homs = torch.eye(3, dtype=torch.float64, device=core.device).repeat(4, 1, 1)
    with pytest.raises(MeshflowNativeError):
        core.stabilized_displacements(u, homs, 4)


# ----------------------------------------------------------------------------------------------
# (3) warp + crop
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H,R,C,F,amp", [(320, 180, 8, 8, 3, 2.5), (640, 360, 16, 16, 2, 3.0),
                                           (333, 217, 6, 9, 2, 2.0), (200, 120, 4, 6, 2, 15.0),
                                           (1920, 1080, 16, 16, 1, 3.0), (1280, 720, 64, 64, 1, 1.0),
                                           (160, 96, 48, 80, 2, 0.15)])     # 2 x 2 px cells: candidate lists overflow
def test_warp_matches_oracle(W, H, R, C, F, amp):
    rng = np.random.default_rng(W + R)
    frames, u, s = synth.synthetic_warp_inputs(rng, F, W, H, R, C, per_vertex=amp, per_frame=1.2 * amp)
    core = _core(W, H, R, C, border_bgr=(7, 99, 250))
    out, crop, maps = core.warp_frames(_dev(frames, core), _dev(u, core), _dev(s, core), return_maps=True)
    out = out.cpu().numpy(); crop = crop.cpu().numpy(); maps = maps.cpu().numpy()
    ref_frames, ref_crop, ref_maps, ref_pf = spec.warp_stage(list(frames), u, s, R, C, (7, 99, 250), return_maps=True)
    for f in range(F):
        assert np.array_equal(maps[f, :, :, 0], ref_maps[f][0]), f"map_x differs in {(maps[f, :, :, 0] != ref_maps[f][0]).sum()} px"
        assert np.array_equal(maps[f, :, :, 1], ref_maps[f][1])
        assert np.array_equal(out[f], ref_frames[f]), f"{(out[f] != ref_frames[f]).sum()} samples differ"
        assert crop[f].tolist() == ref_pf[f].tolist()
    # production path (no maps requested): row segments + float32 coordinates outside the rounding band
    out_fast, crop_fast = core.warp_frames(_dev(frames, core), _dev(u, core), _dev(s, core))
    out_fast = out_fast.cpu().numpy(); crop_fast = crop_fast.cpu().numpy()
    for f in range(F):
        assert np.array_equal(out_fast[f], ref_frames[f]), f"fast path: {(out_fast[f] != ref_frames[f]).any(axis=2).sum()} px differ"
        assert crop_fast[f].tolist() == ref_pf[f].tolist()
    enc = core.combine_crop(torch.from_numpy(crop).to(core.device))
    assert core.decode_crop(enc) == tuple(ref_crop)
    bounds = core.warp_crop_bounds(_dev(u, core), _dev(s, core)).cpu().numpy()       # pass A: no pixels read
    assert np.array_equal(bounds, crop)
    if ref_crop[0] <= ref_crop[2] and ref_crop[1] <= ref_crop[3]:
        a = core.crop_resize_device(torch.from_numpy(out).to(core.device), enc).cpu().numpy()
        b = core.crop_resize(torch.from_numpy(out).to(core.device), ref_crop).cpu().numpy()
        assert np.array_equal(a, b)
        assert np.array_equal(a[0], spec.resize_fixed(out[0][ref_crop[1]:ref_crop[3] + 1, ref_crop[0]:ref_crop[2] + 1], W, H))


@pytest.mark.parametrize("W,H,R,C", [(1920, 1080, 16, 16), (1280, 720, 64, 64), (1000, 562, 16, 16), (3840, 2160, 32, 32)])
def test_warp_fast_path_equals_generic_kernel_on_camera_like_warps(W, H, R, C):
    """Smooth (rotation / scale / perspective) warps as a stabilizer produces them: the production kernel
    (shared-window gather for almost every group) against the generic kernel, whole frames, bit for bit."""
    rng = np.random.default_rng(W + C)
    F = 3
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    rest = spec.vertex_xy(W, H, R, C).astype(np.float64)
    d = np.zeros((F, (R + 1) * (C + 1), 2))
    for f in range(F):
        Hm = synth.random_homography(rng, W, H, rot=0.004 * (1 + 3 * f), scale=0.003 * (1 + 3 * f), trans=4.0 * (1 + f))
        w = rest[:, 0] * Hm[2, 0] + rest[:, 1] * Hm[2, 1] + 1.0
        d[f, :, 0] = (rest[:, 0] * Hm[0, 0] + rest[:, 1] * Hm[0, 1] + Hm[0, 2]) / w - rest[:, 0]
        d[f, :, 1] = (rest[:, 0] * Hm[1, 0] + rest[:, 1] * Hm[1, 1] + Hm[1, 2]) / w - rest[:, 1]
    d += rng.normal(0, 0.1, d.shape)
    u = np.zeros((F, R + 1, C + 1, 2)); s = d.reshape(F, R + 1, C + 1, 2)
    core = _core(W, H, R, C, border_bgr=(1, 2, 3))
    fd = _dev(frames, core)
    gen, crop_gen, _ = core.warp_frames(fd, _dev(u, core), _dev(s, core), return_maps=True)
    fast, crop_fast = core.warp_frames(fd, _dev(u, core), _dev(s, core))
    assert torch.equal(gen, fast), f"{(gen != fast).any(dim=3).sum().item()} px differ"
    assert torch.equal(crop_gen, crop_fast)
    assert torch.equal(core.warp_crop_bounds(_dev(u, core), _dev(s, core)), crop_gen)


def _camera_like(rng, W, H, R, C, F, rough=0.1):
    rest = spec.vertex_xy(W, H, R, C).astype(np.float64)
    d = np.zeros((F, (R + 1) * (C + 1), 2))
    for f in range(F):
        Hm = synth.random_homography(rng, W, H, rot=0.004 * (1 + 3 * f), scale=0.003 * (1 + 3 * f), trans=4.0 * (1 + f))
        w = rest[:, 0] * Hm[2, 0] + rest[:, 1] * Hm[2, 1] + 1.0
        d[f, :, 0] = (rest[:, 0] * Hm[0, 0] + rest[:, 1] * Hm[0, 1] + Hm[0, 2]) / w - rest[:, 0]
        d[f, :, 1] = (rest[:, 0] * Hm[1, 0] + rest[:, 1] * Hm[1, 1] + Hm[1, 2]) / w - rest[:, 1]
    d += rng.normal(0, rough, d.shape)
    return np.zeros((F, R + 1, C + 1, 2)), d.reshape(F, R + 1, C + 1, 2)


@pytest.mark.parametrize("W,H,R,C,rough", [(1920, 1080, 16, 16, 0.1), (1280, 720, 64, 64, 0.05), (333, 217, 6, 9, 1.5),
                                           (3840, 2160, 32, 32, 0.1), (640, 360, 16, 16, 4.0), (200, 120, 4, 6, 8.0),
                                           (130, 70, 4, 4, 0.3)])
def test_fused_pass_equals_warp_then_resize(W, H, R, C, rough):
    """Pass B in one kernel (stabilized tile in shared memory, resized from there) against the two-kernel
    sequence warp_frames -> crop_resize, bit for bit, for the video's own crop rectangle and for hand-picked
    ones: the whole frame (scale 1: every tile needs its one extra row and column), a tight crop, and crops
    that put tile borders at odd phases."""
    rng = np.random.default_rng(W + 7 * C)
    F = 3
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    u, s = _camera_like(rng, W, H, R, C, F, rough)
    core = _core(W, H, R, C, border_bgr=(5, 120, 250))
    fd, ud, sd = _dev(frames, core), _dev(u, core), _dev(s, core)
    stab, crop_pf = core.warp_frames(fd, ud, sd)
    crop_a, tables = core.warp_prepare(ud, sd)
    assert torch.equal(crop_a, crop_pf)
    own = core.decode_crop(core.combine_crop(crop_pf))
    crops = [(0, 0, W - 1, H - 1), (W // 5, H // 7, W - 1 - W // 6, H - 1 - H // 5), (1, 2, W - 4, H - 3),
             (W // 3, H // 3, W // 3 + max(8, W // 4), H // 3 + max(8, H // 4))]
    if own[0] <= own[2] and own[1] <= own[3]:
        crops.append(own)
    for (l, t, r, b) in crops:
        enc = torch.tensor([l, t, -r, -b], dtype=torch.int32, device=core.device)
        ref = core.crop_resize_device(stab, enc)
        got = core.warp_resize_frames(fd, enc, tables)
        assert torch.equal(got, ref), f"crop {(l, t, r, b)}: {(got != ref).any(dim=3).sum().item()} px differ"
        # a chunk in the middle of the prepared video
        part = core.warp_resize_frames(fd[1:3], enc, tables, first_frame=1)
        assert torch.equal(part, ref[1:3])
    # tables rebuilt per chunk (what long videos do) give the same frames
    enc = torch.tensor([crops[1][0], crops[1][1], -crops[1][2], -crops[1][3]], dtype=torch.int32, device=core.device)
    assert torch.equal(core.warp_crop_resize(fd[2:3], ud[2:3], sd[2:3], enc), core.crop_resize_device(stab, enc)[2:3])


def test_warp_many_frames_back_to_back_on_one_workspace():
    """70 frames per call, three calls in a row on the same workspace (different displacements in between): same
    frames and crop edges as the generic kernel every time."""
    W, H, R, C, F = 320, 200, 8, 8, 70
    rng = np.random.default_rng(77)
    frames, u, s = synth.synthetic_warp_inputs(rng, F, W, H, R, C, per_vertex=1.0, per_frame=2.0)
    core = _core(W, H, R, C, border_bgr=(3, 2, 1))
    fd, ud, sd = _dev(frames, core), _dev(u, core), _dev(s, core)
    gen, crop_gen, _ = core.warp_frames(fd, ud, sd, return_maps=True)
    a, crop_a = core.warp_frames(fd, ud, sd)
    b, crop_b = core.warp_frames(fd, ud, _dev(u, core))          # identity right behind it, same workspace
    c, crop_c = core.warp_frames(fd, ud, sd)
    assert torch.equal(a, gen) and torch.equal(crop_a, crop_gen)
    assert torch.equal(b, fd)
    assert torch.equal(c, gen) and torch.equal(crop_c, crop_gen)
    assert torch.equal(core.warp_crop_bounds(ud, sd), crop_gen)


def test_warp_identity_is_a_copy():
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (2, 180, 320, 3), dtype=np.uint8)
    u = rng.normal(0, 5, (2, 9, 9, 2))
    core = _core(320, 180, 8, 8)
    out, crop = core.warp_frames(_dev(frames, core), _dev(u, core), _dev(u, core))
    assert np.array_equal(out.cpu().numpy(), frames)
    assert crop.cpu().numpy().tolist() == [[0, 0, 319, 179]] * 2


@pytest.mark.parametrize("W,H,crop", [(640, 360, (13, 9, 620, 344)), (1920, 1080, (79, 55, 1850, 1000)),
                                      (333, 217, (0, 0, 332, 216)), (640, 360, (100, 10, 300, 350))])
def test_crop_resize_matches_oracle(W, H, crop):
    rng = np.random.default_rng(W)
    frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    core = _core(W, H)
    out = core.crop_resize(_dev(frames, core), crop).cpu().numpy()
    l, t, r, b = crop
    for f in range(2):
        assert np.array_equal(out[f], spec.resize_fixed(frames[f][t:b + 1, l:r + 1], W, H))


def test_stability_score_matches_oracle():
    rng = np.random.default_rng(5)
    u, _ = synth.synthetic_paths(rng, 240, 16, 16)
    core = _core(640, 360)
    got = float(core.stability_score(_dev(u, core)).item())
    from oracle import reference_port as port
    assert abs(got - port.stability_score(u)) <= 1e-10 * abs(got)


def test_streamed_schedule_equals_stage_sequence():
    """Host buffers in / out with chunked, overlapped copies must give the very same bytes."""
    from meshflow_b200 import StreamedCore
    W, H, R, C, F = 320, 180, 8, 8, 23
    rng = np.random.default_rng(77)
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    tr = synth.synthetic_tracks(rng, F - 1, 500, W, H)
    core = _core(W, H, R, C)
    vel = core.vertex_velocities(_dev(tr["early"], core), _dev(tr["late"], core), _dev(tr["offset"], core),
                                 _dev(tr["keep"], core), _dev(tr["pair_start"], core),
                                 _dev(tr["homographies"].reshape(-1, 9), core), pair_start_host=tr["pair_start"])
    u = core.prefix_displacements(vel)
    homs = torch.cat([_dev(tr["homographies"].reshape(-1, 9), core), torch.eye(3, dtype=torch.float64, device=core.device).reshape(1, 9)])
    s = core.stabilized_displacements(u, homs, 2)
    stab, crop_pf = core.warp_frames(_dev(frames, core), u, s)
    enc = core.combine_crop(crop_pf)
    ref = core.crop_resize_device(stab, enc).cpu().numpy()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    tracks = {k: pin(tr[k]) for k in ("early", "late", "offset", "keep", "pair_start")}
    tracks["homographies"] = pin(tr["homographies"].reshape(-1, 9))
    h_out = torch.zeros((F, H, W, 3), dtype=torch.uint8).pin_memory()
    sc = StreamedCore(core, chunk_frames=5)
    for _ in range(2):                                        # second run reuses every buffer
        enc2, u2, s2 = sc.run(pin(frames), tracks, h_out, 2)
        torch.cuda.synchronize()
        assert core.decode_crop(enc2) == core.decode_crop(enc)
        assert torch.equal(u2, u) and torch.equal(s2, s)
        assert np.array_equal(h_out.numpy(), ref)
    # long videos: no table budget -> crop rectangle chunk by chunk, tables rebuilt per chunk in pass B;
    # frames already resident on the device; every landed chunk reported to the host in order
    h_out.zero_()
    landed = []
    sc0 = StreamedCore(core, chunk_frames=4, table_budget_bytes=0)
    enc3, u3, s3 = sc0.run(None, tracks, h_out, 2, d_frames=_dev(frames, core), on_chunk_landed=lambda f0, n: landed.append((f0, n)))
    assert core.decode_crop(enc3) == core.decode_crop(enc) and torch.equal(s3, s)
    assert np.array_equal(h_out.numpy(), ref)
    assert [f0 for f0, _ in landed] == sorted(f0 for f0, _ in landed) and sum(n for _, n in landed) == F


def test_stabilize_frames_matches_the_reference_port_end_to_end():
    """Whole pipeline on a small synthetic video through the drop-in class (host OpenCV front end,
    streamed GPU core, host metrics) against the CPU port of the reference."""
    from meshflow_b200 import MeshFlowStabilizer
    from oracle import reference_port as port
    frames = synth.textured_video(np.random.default_rng(12), 14, 320, 180, jitter=2.0)
    for definition in (0, 2):
        ref = port.stabilize_frames(port.Params(mesh_row_count=8, mesh_col_count=8), frames, definition)
        got = MeshFlowStabilizer(mesh_row_count=8, mesh_col_count=8, chunk_frames=4).stabilize_frames(frames, definition)
        assert np.array_equal(got["u"], ref["u"]) and np.array_equal(got["homographies"], ref["homographies"])
        assert np.abs(got["s"] - ref["s"]).max() <= 1e-9 * max(1.0, np.abs(ref["s"]).max())
        assert tuple(int(c) for c in got["crop_boundaries"]) == tuple(int(c) for c in ref["crop"])
        assert all(np.array_equal(a, b) for a, b in zip(got["cropped_frames"], ref["cropped"]))
        # the three returned metrics: within 1e-4 relative (north star); types as the reference's
        assert abs(got["cropping_ratio"] - ref["cropping_ratio"]) <= 1e-4 * abs(ref["cropping_ratio"])
        assert abs(got["distortion_score"] - ref["distortion_score"]) <= 1e-4 * abs(ref["distortion_score"])
        assert abs(got["stability_score"] - ref["stability_score"]) <= 1e-4 * abs(ref["stability_score"])
        assert isinstance(got["cropping_ratio"], np.float32) and isinstance(got["stability_score"], np.float64)


def test_reference_named_stage_methods_against_goldens():
    """The reference's private stage methods, same names and signatures, fed with the committed golden
    inputs produced by the unmodified reference (tests/golden/make_golden.py)."""
    import os
    from meshflow_b200 import MeshFlowStabilizer
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    g = np.load(os.path.join(G, "video1_motion_paths.npz"))
    w, h = (int(v) for v in g["frame_size"])
    st = MeshFlowStabilizer()
    dummy = [np.zeros((h, w, 3), np.uint8)]
    for d in range(4):
        s = st._get_stabilized_vertex_displacements(13, dummy, d, g["u"], g["homographies"])
        assert np.abs(s - g[f"s_{d}"]).max() <= 1e-9 * np.abs(g[f"s_{d}"]).max()
        lam = st._get_adaptive_weights(13, w, h, d, g["homographies"])
        assert np.allclose(lam, g[f"lambda_{d}"], rtol=1e-12, atol=0)
        assert abs(st._compute_stability_score(13, g[f"s_{d}"]) - float(g[f"stability_{d}"])) <= 1e-10
    gw = np.load(os.path.join(G, "warp_small.npz"))
    frames = list(gw["frames"])
    st8 = MeshFlowStabilizer(mesh_row_count=8, mesh_col_count=8, color_outside_image_area_bgr=(10, 200, 30))
    for case in ("mild", "wild"):
        stab, crop = st8._get_stabilized_frames_and_crop_boundaries(3, frames, gw[f"{case}_u"], gw[f"{case}_s"])
        assert np.array_equal(np.stack(stab), gw[f"{case}_stabilized"])
        assert [int(c) for c in crop] == gw[f"{case}_crop"].tolist() and isinstance(crop[0], np.int64)
        if f"{case}_cropped" in gw:
            assert np.array_equal(np.stack(st8._crop_frames(stab, crop)), gw[f"{case}_cropped"])
    # vertex motion from the golden matched features (device masks = all kept)
    from meshflow_b200 import DeviceCore, MeshSpec
    core = DeviceCore(MeshSpec(w, h, 16, 16))
    for t in range(4):
        e, l = g[f"early_{t}"], g[f"late_{t}"]
        n = len(e)
        # the reference's features are float32 subframe coordinates + the integer subframe origin (same
        # origin for the early and the late point); the golden holds the float64 sums
        off = np.stack([np.floor(e[:, 0] / 160.0) * 160, np.floor(e[:, 1] / 90.0) * 90], axis=1)
        es, ls = (e - off).astype(np.float32), (l - off).astype(np.float32)
        assert np.array_equal(es.astype(np.float64) + off, e) and np.array_equal(ls.astype(np.float64) + off, l)
        vel = core.vertex_velocities(_dev(es, core), _dev(ls, core), _dev(off.astype(np.int32), core),
                                     _dev(np.ones(n, np.uint8), core), _dev(np.array([0, n], np.int32), core),
                                     _dev(g["pair_homographies"][t].reshape(1, 9), core),
                                     pair_start_host=np.array([0, n], np.int32)).cpu().numpy()[0]
        assert np.array_equal(vel, g["velocities"][t])


def test_fast_reciprocal_is_correctly_rounded():
    """The warp kernel's reciprocal (hardware seed + Newton + Markstein correction) must equal the
    IEEE correctly rounded 1/w bit for bit: 200 M random doubles, 1/8 of them in the w ~ 1 regime."""
    import ctypes
    from meshflow_b200 import _cabi
    lib = _cabi.load()
    lib.mf_debug_rcp_mismatches.restype = ctypes.c_longlong
    lib.mf_debug_rcp_mismatches.argtypes = [ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
    scratch = torch.zeros(1, dtype=torch.int64, device="cuda")
    bad = lib.mf_debug_rcp_mismatches(200_000_000, 12345, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert bad == 0
