"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference for runs on the GPU box.

The reference (how4rd/meshflow) is one pure-Python file with no build system (no setup.py / pyproject:
``pip install /root/reference`` has nothing to install), so "building" it means copying, byte for byte,

    /root/reference/meshflowstabilizer.py               -> baseline/_ref/meshflowstabilizer.py
    /root/reference/videos/video-N/video-N.m4v          -> baseline/_ref/video-N.m4v    (the seven input clips, 14 MB)

``baseline/_ref/`` is git-ignored (the reference's sources never enter this repository's history) but NOT
gpurun-ignored, so the copy travels to the GPU box, where ``/root/reference`` does not exist.  It is used
by ``bench.py --impl reference`` (kind "reference": the reference's own stage methods timed on the box's
host cores) and as the input of ``tests/test_gpu_video1.py`` (BASELINE.json configs[0]).  Nothing under
``meshflow_b200/`` imports it.

    python oracle/make_ref.py        # no-op when /root/reference is absent (e.g. on the GPU box)
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
VIDEOS = (1, 2, 3, 5, 8, 9, 10)
FILES = {"meshflowstabilizer.py": "meshflowstabilizer.py"}
FILES.update({f"videos/video-{n}/video-{n}.m4v": f"video-{n}.m4v" for n in VIDEOS})


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: keeping whatever is already under {DST}")
        return False
    os.makedirs(DST, exist_ok=True)
    for rel, name in FILES.items():
        src, dst = os.path.join(SRC, rel), os.path.join(DST, name)
        if not os.path.exists(dst) or os.path.getsize(dst) != os.path.getsize(src):
            shutil.copyfile(src, dst)
            os.chmod(dst, 0o644)
    if verbose:
        print(f"staged the unmodified reference under {DST}: {sorted(os.listdir(DST))}")
    return True


def reference_module():
    """Import ``baseline/_ref/meshflowstabilizer.py`` (None when it was never staged)."""
    path = os.path.join(DST, "meshflowstabilizer.py")
    if not os.path.exists(path):
        return None
    import contextlib
    import importlib.util
    import warnings
    import tqdm
    # silence the reference's progress bars (its loops only iterate over them)
    tqdm.trange = lambda n, *a, **k: contextlib.nullcontext(
        type("T", (), {"set_description": lambda s, d: None, "__iter__": lambda s: iter(range(n))})())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = importlib.util.spec_from_file_location("meshflowstabilizer_ref", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return mod


def video_path(n=1):
    path = os.path.join(DST, f"video-{n}.m4v")
    return path if os.path.exists(path) else None


if __name__ == "__main__":
    stage()
    sys.exit(0)
