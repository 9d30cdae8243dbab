"""Multi-GPU bit-identity, collected by ``pytest -m gpu``: spawns ``tests/multi_gpu_check.py`` under
``torch.distributed.run`` on every GPU of the box (at most 8); skipped on a single-GPU box.  The log of the
last run is kept in ``gpurun_out/multi_gpu_check.log`` (copied to ``profiles/`` for the record)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_sharded_schedule_is_bit_identical_to_one_gpu():
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs at least two GPUs on the box")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    proc = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    log = proc.stdout + "\n--- stderr ---\n" + proc.stderr
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "multi_gpu_check.log"), "w") as fo:
        fo.write(f"$ {' '.join(cmd)}\nexit code {proc.returncode}\n{log}")
    assert proc.returncode == 0, log[-4000:]
    assert f"multi-GPU parity OK on {n} GPUs" in proc.stdout


def test_core_on_a_device_that_is_not_the_current_one():
    """DeviceCore / StreamedCore / MeshFlowStabilizer(device=...) bound to cuda:1 while cuda:0 is the current device:
    every native call must run on the core's own GPU and stream (round-1 advisor finding)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs on the box")
    import numpy as np
    from meshflow_b200 import DeviceCore, MeshSpec, StreamedCore
    from tests import synth
    torch.cuda.set_device(0)
    W, H, R, C, F = 320, 180, 8, 8, 9
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
    tr = synth.synthetic_tracks(rng, F - 1, 400, W, H)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    tracks = {k: pin(tr[k]) for k in ("early", "late", "offset", "keep", "pair_start")}
    tracks["homographies"] = pin(tr["homographies"].reshape(-1, 9))
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        core = DeviceCore(MeshSpec(W, H, R, C), device=dev)
        assert torch.cuda.current_device() == 0
        h_out = torch.zeros((F, H, W, 3), dtype=torch.uint8).pin_memory()
        enc, u, s = StreamedCore(core, chunk_frames=4).run(pin(frames), tracks, h_out, 0)
        torch.cuda.synchronize(core.device)
        assert u.device == core.device and enc.device == core.device and torch.cuda.current_device() == 0
        outs.append((h_out.numpy().copy(), u.cpu(), s.cpu(), core.decode_crop(enc)))
    assert np.array_equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    assert outs[0][3] == outs[1][3]
