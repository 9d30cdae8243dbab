"""GPU tests at BASELINE.json's full sizes (configs 3-5), where the CPU oracle is too slow to run in
full: a sampled comparison with the oracle plus size-independent properties (linearity of the Jacobi
solve, identity / integer-translation warps, identity resize), and the ragged / empty edge cases."""
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


def test_jacobi_config4_full_size_sampled_oracle_and_linearity():
    """c4: 64x64 mesh x 10 000 frames, radius 30, 500 iterations (the reference would need ~45 h)."""
    F, R, radius, iters = 10000, 64, 30, 500
    rng = np.random.default_rng(7)
    core = _core(1920, 1080, R, R, radius=radius, iterations=iters)
    g = torch.Generator(device=core.device).manual_seed(7)
    V = (R + 1) * (R + 1)
    steps = torch.randn((F, V, 2), generator=g, device=core.device, dtype=torch.float64) * 3.0
    u1 = torch.cumsum(steps, dim=0).view(F, R + 1, R + 1, 2).contiguous()
    del steps
    u2 = torch.randn((F, R + 1, R + 1, 2), generator=g, device=core.device, dtype=torch.float64) * 10.0
    homs = np.tile(np.eye(3), (F, 1, 1))
    homs[:, :2, :2] += rng.normal(0, 0.01, (F, 2, 2))
    homs[:, :2, 2] = rng.normal(0, 8.0, (F, 2))
    homs[-1] = np.eye(3)
    hd = _dev(homs, core)
    for definition in (0, 2):
        s1 = core.stabilized_displacements(u1, hd, definition)
        # sampled oracle: 3 vertices, every frame, every sweep
        idx = [0, 2077, V - 1]
        sub = u1.view(F, V, 2)[:, idx].cpu().numpy()
        ref = spec.jacobi_banded(sub, homs, 1920, 1080, radius, iters, definition)
        got = s1.view(F, V, 2)[:, idx].cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
        # linearity over the full tensor: solve(2 u1 - 0.5 u2) == 2 solve(u1) - 0.5 solve(u2)
        s2 = core.stabilized_displacements(u2, hd, definition)
        mix = core.stabilized_displacements(2.0 * u1 - 0.5 * u2, hd, definition)
        err = (mix - (2.0 * s1 - 0.5 * s2)).abs().max().item()
        assert err <= 1e-10 * s1.abs().max().item()
        del s1, s2, mix


@pytest.mark.parametrize("W,H,R", [(3840, 2160, 32), (7680, 4320, 16), (1280, 720, 64)])
def test_warp_identity_and_integer_translation_at_full_resolution(W, H, R):
    """u == s must copy the frame; s - u = (7, -5) for every vertex must shift it by whole pixels."""
    rng = np.random.default_rng(W)
    frame = rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
    core = _core(W, H, R, R, border_bgr=(1, 2, 3))
    u = rng.normal(0, 4, (1, R + 1, R + 1, 2))
    fd = _dev(frame, core)
    out, crop = core.warp_frames(fd, _dev(u, core), _dev(u, core))
    assert torch.equal(out, fd)
    assert crop.cpu().numpy().tolist() == [[0, 0, W - 1, H - 1]]
    s = u + np.array([7.0, -5.0])
    out, crop = core.warp_frames(fd, _dev(u, core), _dev(s, core))
    out = out.cpu().numpy()[0]
    expect = np.empty_like(frame[0]); expect[:] = (1, 2, 3)
    expect[:H - 5, 7:] = frame[0][5:, :W - 7]          # output (x, y) <- source (x - 7, y + 5)
    assert np.array_equal(out, expect)
    assert crop.cpu().numpy().tolist() == [[7, 0, W - 1, H - 6]]
    # the bounds-only pass agrees without reading a pixel
    assert core.warp_crop_bounds(_dev(u, core), _dev(s, core)).cpu().numpy().tolist() == [[7, 0, W - 1, H - 6]]


@pytest.mark.parametrize("W,H", [(3840, 2160), (7680, 4320)])
def test_resize_identity_crop_is_a_copy(W, H):
    rng = np.random.default_rng(H)
    frame = rng.integers(0, 256, (1, H, W, 3), dtype=np.uint8)
    core = _core(W, H)
    fd = _dev(frame, core)
    assert torch.equal(core.crop_resize(fd, (0, 0, W - 1, H - 1)), fd)


def test_warp_4k_sampled_against_the_oracle():
    """c3 geometry (4K, 32x32 mesh): one frame against the per-element oracle."""
    W, H, R = 3840, 2160, 32
    rng = np.random.default_rng(4321)
    frames, u, s = synth.synthetic_warp_inputs(rng, 1, W, H, R, R, per_vertex=2.5, per_frame=3.0)
    core = _core(W, H, R, R)
    out, crop, maps = core.warp_frames(_dev(frames, core), _dev(u, core), _dev(s, core), return_maps=True)
    ref_frames, ref_crop, ref_maps, ref_pf = spec.warp_stage(list(frames), u, s, R, R, (0, 0, 255), return_maps=True)
    maps = maps.cpu().numpy()
    assert np.array_equal(maps[0, :, :, 0], ref_maps[0][0]) and np.array_equal(maps[0, :, :, 1], ref_maps[0][1])
    assert np.array_equal(out.cpu().numpy()[0], ref_frames[0])
    assert crop.cpu().numpy()[0].tolist() == ref_pf[0].tolist()


def test_vertex_motion_ragged_and_empty_pairs():
    """Pairs without a single kept feature (and a pair with no features at all) reduce to the global
    homography's velocity; neighbours are unaffected."""
    W, H, R, C = 640, 360, 16, 16
    rng = np.random.default_rng(9)
    tr = synth.synthetic_tracks(rng, 4, 800, W, H)
    a, b = tr["pair_start"][1], tr["pair_start"][2]
    tr["keep"][a:b] = 0                                             # pair 1: everything masked out
    # pair 2: no features at all
    c, d = tr["pair_start"][2], tr["pair_start"][3]
    for k in ("early", "late", "offset", "keep"):
        tr[k] = np.concatenate([tr[k][:c], tr[k][d:]])
    tr["pair_start"] = np.array([tr["pair_start"][0], tr["pair_start"][1], tr["pair_start"][2], tr["pair_start"][2],
                                 tr["pair_start"][4] - (d - c)], dtype=np.int32)
    core = _core(W, H, R, C)
    for host in (tr["pair_start"], None):
        vel = core.vertex_velocities(_dev(tr["early"], core), _dev(tr["late"], core), _dev(tr["offset"], core),
                                     _dev(tr["keep"], core), _dev(tr["pair_start"], core),
                                     _dev(tr["homographies"].reshape(-1, 9), core), pair_start_host=host).cpu().numpy()
        for p in range(4):
            i, j = tr["pair_start"][p], tr["pair_start"][p + 1]
            k = tr["keep"][i:j].astype(bool)
            off = tr["offset"][i:j][k].astype(np.float64)
            ref = spec.vertex_velocities(tr["early"][i:j][k].astype(np.float64) + off, tr["late"][i:j][k].astype(np.float64) + off,
                                         tr["homographies"][p], W, H, R, C, 10, 10)
            assert np.array_equal(vel[p], ref), f"pair {p}"


def test_two_frame_video_runs_through_every_stage():
    """Smallest possible input: F = 2 (the reference itself needs F > radius, mfs.py:780)."""
    W, H, R, C = 320, 180, 8, 8
    rng = np.random.default_rng(2)
    frames = rng.integers(0, 256, (2, H, W, 3), dtype=np.uint8)
    tr = synth.synthetic_tracks(rng, 1, 300, W, H)
    core = _core(W, H, R, C)
    vel = core.vertex_velocities(_dev(tr["early"], core), _dev(tr["late"], core), _dev(tr["offset"], core),
                                 _dev(tr["keep"], core), _dev(tr["pair_start"], core),
                                 _dev(tr["homographies"].reshape(-1, 9), core), pair_start_host=tr["pair_start"])
    u = core.prefix_displacements(vel)
    homs = np.concatenate([tr["homographies"], np.eye(3)[None]])
    s = core.stabilized_displacements(u, _dev(homs, core), 0)
    ref = spec.jacobi_banded(u.cpu().numpy(), homs, W, H, 10, 100, 0)
    assert np.abs(s.cpu().numpy() - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    out, crop = core.warp_frames(_dev(frames, core), u, s)
    ref_frames, ref_crop = spec.warp_stage(list(frames), u.cpu().numpy(), s.cpu().numpy(), R, C, (0, 0, 255))
    assert np.array_equal(out.cpu().numpy(), np.stack(ref_frames))


def test_config3_geometry_streamed_equals_stage_sequence():
    """c3 geometry (4K, 32x32 mesh, CONSTANT_HIGH) on a short clip: host-in/host-out streamed schedule
    == resident stage sequence, and the bounds-only pass == the pixel pass."""
    from meshflow_b200 import StreamedCore
    W, H, R, F = 3840, 2160, 32, 20
    rng = np.random.default_rng(4321)
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    tr = synth.synthetic_tracks(rng, F - 1, 2500, W, H)
    core = _core(W, H, R, R)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    tracks = {k: pin(tr[k]) for k in ("early", "late", "offset", "keep", "pair_start")}
    tracks["homographies"] = pin(tr["homographies"].reshape(-1, 9))
    h_out = torch.zeros((F, H, W, 3), dtype=torch.uint8).pin_memory()
    enc, u, s = StreamedCore(core, chunk_frames=8).run(pin(frames), tracks, h_out, 2)
    torch.cuda.synchronize()
    stab, crop_pf = core.warp_frames(_dev(frames, core), u, s)
    assert torch.equal(core.combine_crop(crop_pf), enc)
    assert torch.equal(core.warp_crop_bounds(u, s), crop_pf)
    ref = core.crop_resize_device(stab, enc).cpu().numpy()
    assert np.array_equal(h_out.numpy(), ref)


def test_jacobi_beyond_the_on_chip_frame_limit():
    """Videos longer than the on-chip kernels hold (F > 10 240) take the global-memory sweeps: same results
    (sampled oracle), odd and even sweep counts (the ping-pong must end in the output), vertex shard respected."""
    F, R, radius = 10400, 2, 10
    rng = np.random.default_rng(31)
    u, homs = synth.synthetic_paths(rng, F, R, R)
    for iters in (7, 8):
        core = _core(640, 360, R, R, radius=radius, iterations=iters)
        ud, hd = _dev(u, core), _dev(homs, core)
        s = core.stabilized_displacements(ud, hd, 0).cpu().numpy()
        ref = spec.jacobi_banded(u, homs, 640, 360, radius, iters, 0)
        assert np.abs(s - ref).max() <= 1e-9 * np.abs(ref).max()
        out = torch.full((F, R + 1, R + 1, 2), -7.0, dtype=torch.float64, device=core.device)
        core.stabilized_displacements(ud, hd, 0, vertex_range=(2, 5), out=out)
        o = out.cpu().numpy().reshape(F, -1, 2)
        assert np.array_equal(o[:, 2:5], s.reshape(F, -1, 2)[:, 2:5]) and np.all(o[:, :2] == -7.0) and np.all(o[:, 5:] == -7.0)
