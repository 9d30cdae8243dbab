"""BASELINE.json configs[0] on the GPU: the full ``videos/video-1`` (494 frames, 640x360, constructor
defaults) through ``MeshFlowStabilizer.stabilize()`` -- file in, file out -- for all four
ADAPTIVE_WEIGHTS_DEFINITION_* variants, against the record the UNMODIFIED reference produced
(``tests/golden/video1_full.npz``, written by ``tests/golden/make_golden_c1.py``; the same values are in
BASELINE.md "Golden values").

Bars: u, homographies, crop rectangle and every cropped pixel bit-exact (sha256 of the bytes); stabilized
paths <= 1e-9 relative (north-star tolerance 1e-4); the returned tuple <= 1e-4 relative with the
reference's types.  The video travels to the GPU box in the git-ignored ``baseline/_ref/``
(``oracle/make_ref.py``)."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VIDEO = os.path.join(ROOT, "baseline", "_ref", "video-1.m4v")
GOLDEN = os.path.join(ROOT, "tests", "golden", "video1_full.npz")
OTHER_VIDEOS = (2, 3, 5, 8, 9, 10)          # the reference's other input clips (videos/credits.txt)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def golden():
    if not os.path.exists(VIDEO):
        pytest.skip("baseline/_ref/video-1.m4v not staged (python oracle/make_ref.py in the build container)")
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/video1_full.npz missing")
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def stabilizer():
    from meshflow_b200 import MeshFlowStabilizer

    class Spy(MeshFlowStabilizer):
        """stabilize() only returns the tuple; keep what it computed on the way."""
        def stabilize_frames(self, frames, *a, **k):
            self.seen_frames_sha = sha(np.stack(frames))
            self.seen = super().stabilize_frames(frames, *a, **k)
            return self.seen

        def _write_stabilized_video(self, output_path, num_frames, fps, codec, frames):
            self.written = (num_frames, sha(np.stack(list(frames))))
            return super()._write_stabilized_video(output_path, num_frames, fps, codec, frames)

    return Spy()


@pytest.mark.parametrize("definition", [0, 1, 2, 3])
def test_video1_full_stabilize_matches_the_reference(golden, stabilizer, definition, tmp_path):
    g = golden
    if f"tuple_{definition}" not in g:
        pytest.skip(f"golden record has no definition {definition}")
    got = stabilizer.stabilize(VIDEO, str(tmp_path / "out.m4v"), definition)
    if stabilizer.seen_frames_sha != str(g["frames_sha"]):
        pytest.skip("this box decodes video-1 to different pixels than the build container (other FFmpeg build)")
    r = stabilizer.seen
    F = int(g["num_frames"])
    assert len(r["cropped_frames"]) == F == 494
    assert sha(r["u"]) == str(g["u_sha"]), "unstabilized vertex displacements differ from the reference"
    assert sha(r["homographies"]) == str(g["homographies_sha"])
    V = r["s"].shape[1] * r["s"].shape[2]
    s_got = r["s"].reshape(F, V, 2)[:, g["sample_vertices"]]
    s_ref = g[f"s_sample_{definition}"]
    assert np.abs(s_got - s_ref).max() <= 1e-9 * np.abs(s_ref).max()
    assert [int(c) for c in r["crop_boundaries"]] == g[f"crop_{definition}"].tolist()
    assert sha(np.stack(r["cropped_frames"])) == str(g[f"cropped_sha_{definition}"]), "cropped pixels differ"
    assert stabilizer.written == (F, str(g[f"cropped_sha_{definition}"]))
    ref = g[f"tuple_{definition}"]
    for k in range(3):
        assert abs(float(got[k]) - ref[k]) <= 1e-4 * abs(ref[k]), (k, got, ref)
    assert [type(v).__name__ for v in got] == [str(t) for t in g[f"tuple_types_{definition}"]]


@pytest.mark.parametrize("n", OTHER_VIDEOS)
def test_other_reference_videos_match_the_reference(stabilizer, n, tmp_path):
    """The reference's other six input clips (246-572 frames of 640x360), file -> file: same bars as video-1 against
    ``tests/golden/videoN_full.npz`` (the unmodified reference run in the build container) for every
    ADAPTIVE_WEIGHTS_DEFINITION the record holds (all four for video-2, -5 and -10, ORIGINAL for the rest)."""
    video = os.path.join(ROOT, "baseline", "_ref", f"video-{n}.m4v")
    golden = os.path.join(ROOT, "tests", "golden", f"video{n}_full.npz")
    if not os.path.exists(video) or not os.path.exists(golden):
        pytest.skip(f"video-{n} or its golden record not staged")
    g = np.load(golden)
    definitions = [d for d in range(4) if f"tuple_{d}" in g]
    if not definitions:
        pytest.skip("golden record incomplete")
    for d in definitions:
        got = stabilizer.stabilize(video, str(tmp_path / "out.m4v"), d)
        if stabilizer.seen_frames_sha != str(g["frames_sha"]):
            pytest.skip("this box decodes the clip to different pixels than the build container (other FFmpeg build)")
        r = stabilizer.seen
        F = int(g["num_frames"])
        assert len(r["cropped_frames"]) == F
        assert sha(r["u"]) == str(g["u_sha"]), "unstabilized vertex displacements differ from the reference"
        assert sha(r["homographies"]) == str(g["homographies_sha"])
        V = r["s"].shape[1] * r["s"].shape[2]
        s_ref = g[f"s_sample_{d}"]
        assert np.abs(r["s"].reshape(F, V, 2)[:, g["sample_vertices"]] - s_ref).max() <= 1e-9 * np.abs(s_ref).max()
        assert [int(c) for c in r["crop_boundaries"]] == g[f"crop_{d}"].tolist()
        assert sha(np.stack(r["cropped_frames"])) == str(g[f"cropped_sha_{d}"]), f"definition {d}: cropped pixels differ"
        ref = g[f"tuple_{d}"]
        for k in range(3):
            assert abs(float(got[k]) - ref[k]) <= 1e-4 * abs(ref[k]), (d, k, got, ref)
