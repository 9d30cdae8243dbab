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
