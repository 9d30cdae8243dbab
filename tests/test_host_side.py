"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the Python
boundary mirrors the reference's interface and error behaviour, the host feature stage hands the
device exactly what the reference would have compacted, and the multi-GPU plumbing (gloo, 2 ranks)."""
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_the_header_declares():
    from meshflow_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "meshflow_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_cabi.exported_symbols()) == declared, "ctypes table and header disagree"
    assert lib.mf_built_for_sm() == 100 and lib.mf_version() >= 100


def test_native_errors_are_reported_without_a_gpu():
    from meshflow_b200 import _cabi
    lib = _cabi.load()
    rc = lib.mf_jacobi_solve(None, None, None, 1, 2, 0, 2, 1, 1, 1, 1, 0, None, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.mf_last_error()
    with pytest.raises(_cabi.MeshflowNativeError):
        _cabi.check(rc)
    assert lib.mf_jacobi_workspace_bytes(300, 578) > 0 and lib.mf_warp_workspace_bytes(2, 1920, 1080, 16, 16) > 0


def test_constructor_and_constants_mirror_the_reference():
    from meshflow_b200 import MeshFlowStabilizer as M
    assert (M.ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL, M.ADAPTIVE_WEIGHTS_DEFINITION_FLIPPED,
            M.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_HIGH, M.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_LOW) == (0, 1, 2, 3)
    assert (M.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_HIGH_VALUE, M.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_LOW_VALUE) == (100, 1)
    sig = inspect.signature(M.__init__)
    expected = dict(mesh_row_count=16, mesh_col_count=16, mesh_outlier_subframe_row_count=4,
                    mesh_outlier_subframe_col_count=4, feature_ellipse_row_count=10, feature_ellipse_col_count=10,
                    homography_min_number_corresponding_features=4, temporal_smoothing_radius=10,
                    optimization_num_iterations=100, color_outside_image_area_bgr=(0, 0, 255), visualize=False)
    positional = [p for p in sig.parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD and p.name != "self"]
    assert [p.name for p in positional] == list(expected)            # same order: positional calls still work
    assert {p.name: p.default for p in positional} == expected
    st = inspect.signature(M.stabilize)
    assert list(st.parameters) == ["self", "input_path", "output_path", "adaptive_weights_definition"]
    assert st.parameters["adaptive_weights_definition"].default == 0
    for name in ("_get_unstabilized_vertex_displacements_and_homographies", "_get_unstabilized_vertex_velocities",
                 "_get_stabilized_vertex_displacements", "_get_stabilized_frames_and_crop_boundaries", "_crop_frames",
                 "_compute_cropping_ratio_and_distortion_score", "_compute_stability_score", "_get_vertex_x_y",
                 "_get_matched_features_and_homography", "_get_adaptive_weights", "_write_stabilized_video"):
        assert callable(getattr(M, name))


def test_invalid_definition_raises_value_error_before_any_work():
    from meshflow_b200 import MeshFlowStabilizer
    with pytest.raises(ValueError, match="adaptive_weights_definition"):
        MeshFlowStabilizer().stabilize("/nonexistent.m4v", "/tmp/out.m4v", adaptive_weights_definition=7)


def test_short_video_raises_ioerror(tmp_path):
    from meshflow_b200 import MeshFlowStabilizer
    import cv2

    class FakeCapture:
        def __init__(self, path): self.n = 0
        def get(self, prop): return {cv2.CAP_PROP_FRAME_COUNT: 3, cv2.CAP_PROP_FPS: 30.0, cv2.CAP_PROP_FOURCC: 0}[prop]
        def read(self):
            self.n += 1
            return (self.n <= 2, np.zeros((4, 4, 3), np.uint8) if self.n <= 2 else None)
        def release(self): pass

    orig = cv2.VideoCapture
    cv2.VideoCapture = FakeCapture
    try:
        with pytest.raises(IOError, match="did not have frame 2 of 3"):
            MeshFlowStabilizer()._get_unstabilized_frames_and_video_features("x")
    finally:
        cv2.VideoCapture = orig


def test_vertex_xy_matches_reference_expression():
    from meshflow_b200 import MeshSpec, vertex_xy
    from oracle import reference_port as port
    for (w, h, r, c) in [(640, 360, 16, 16), (1920, 1080, 16, 16), (3840, 2160, 32, 32), (333, 217, 5, 9)]:
        a = vertex_xy(MeshSpec(w, h, r, c))
        b = port.vertex_xy(port.Params(mesh_row_count=r, mesh_col_count=c), w, h).reshape(-1, 2)
        assert a.dtype == np.float32 and np.array_equal(a, b)


def test_host_tracks_apply_masks_like_the_reference():
    """keep-mask compaction of the un-compacted tracks == the reference's compacted features."""
    import cv2
    from meshflow_b200 import host_features, MeshFlowStabilizer
    from oracle import reference_port as port
    from tests import synth
    frames = synth.textured_video(np.random.default_rng(0), 3, 320, 180)
    det = cv2.FastFeatureDetector_create()
    tracks = host_features.track_all_pairs(frames[:-1], frames[1:], workers=2)
    for t, tr in enumerate(tracks):
        e, l, hm = port.matched_features_and_homography(port.Params(), det, frames[t], frames[t + 1])
        ce, cl = tr.compacted()
        assert np.array_equal(ce, e.reshape(-1, 2)) and np.array_equal(cl, l.reshape(-1, 2))
        assert np.array_equal(tr.homography, hm)
        assert tr.early_xy.dtype == np.float32 and tr.keep.dtype == np.uint8 and tr.offset_xy.dtype == np.int32
    packed = MeshFlowStabilizer.pack_tracks(tracks)
    assert packed["pair_start"].tolist() == [0, len(tracks[0].keep), len(tracks[0].keep) + len(tracks[1].keep)]
    assert packed["early"].shape == (packed["pair_start"][-1], 2) and packed["homographies"].shape == (2, 3, 3)
    serial = host_features.track_all_pairs(frames[:-1], frames[1:], workers=1)
    assert all(np.array_equal(a.keep, b.keep) and np.array_equal(a.late_xy, b.late_xy) for a, b in zip(tracks, serial))


def test_too_few_features_raises_like_the_reference():
    from meshflow_b200 import host_features
    blank = np.zeros((90, 160, 3), np.uint8)
    with pytest.raises(ValueError):
        host_features.track_pair(blank, blank)


def test_shard_plans_cover_everything_exactly_once():
    from meshflow_b200 import distributed as d
    for n, world in [(300, 1), (300, 8), (7, 4), (3, 8), (289, 8), (4225, 8)]:
        spans = [d.frame_shard(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        vs = [d.vertex_shard(n, world, r) for r in range(world)]
        covered = sum(e - b for b, e, _ in vs)
        assert covered == n and all(e - b <= size for b, e, size in vs)


def _gloo_worker(rank, world, port_no, tmp):
    import torch
    import torch.distributed as dist
    from meshflow_b200 import distributed as d
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        V, F = 13, 9
        vel_all = torch.randn((F - 1, V, 2), generator=g)
        homs_all = torch.randn((F - 1, 9), generator=g, dtype=torch.float64)
        # ragged plan: rank 0 owns 5 frames, rank 1 owns 4 (the last of them ends the video: 3 pairs)
        plan = d.ShardPlan(rank, world, [5, 4])
        assert plan.total_frames == F and [plan.pairs_needed(r) for r in range(world)] == [5, 3]
        assert plan.first_frame == (0 if rank == 0 else 5)
        plan.validate(frames_local=plan.local_frames, pairs_local=plan.pairs_needed())
        counts = [plan.pairs_needed(r) for r in range(world)]
        b = plan.first_frame
        mine_v = vel_all[b:b + counts[rank]].clone()
        mine_h = homs_all[b:b + counts[rank]].clone()
        assert torch.equal(d.gather_velocities(mine_v, counts), vel_all)
        gv, gh = d.gather_pairs(mine_v, mine_h, counts)                  # one packed exchange
        assert torch.equal(gv, vel_all) and torch.equal(gh, homs_all) and gv.dtype == torch.float32
        # a rank that supplies too few pairs, or disagrees about the plan, fails loudly
        with pytest.raises(ValueError):
            plan.validate(frames_local=plan.local_frames, pairs_local=plan.pairs_needed() - 1)
        bad = d.ShardPlan(rank, world, [5, 4] if rank == 0 else [4, 5])
        with pytest.raises(ValueError):
            bad.validate(frames_local=bad.local_frames, pairs_local=9)
        s_true = torch.randn((F, V, 2), generator=g, dtype=torch.float64)
        v0, v1, _ = d.vertex_shard(V, world, rank)
        mine = torch.full((F, V, 2), float("nan"), dtype=torch.float64)
        mine[:, v0:v1] = s_true[:, v0:v1]
        assert torch.equal(d.gather_paths(mine, V), s_true)
        per_rank = torch.tensor([[5, 2, -600, -340], [9, 1, -610, -330]], dtype=torch.int32)
        enc = per_rank[rank % 2].clone()
        d.reduce_crop(enc, plan)
        assert enc.tolist() == [9, 2, -600, -330]
        alone = per_rank[rank % 2].clone()
        d.reduce_crop(alone, None)                                       # no plan: not sharded, nothing exchanged
        assert alone.tolist() == per_rank[rank % 2].tolist()
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_multi_gpu_exchanges_on_gloo_world_size_2(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    mp.spawn(_gloo_worker, args=(2, port_no, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")
