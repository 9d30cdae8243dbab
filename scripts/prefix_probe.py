"""Time mf_prefix_displacements alone for a long video (the multi-GPU path scans world x F frames on every rank)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meshflow_b200 import DeviceCore, MeshSpec
core = DeviceCore(MeshSpec(1920, 1080, 16, 16))
a = torch.randn(8192, 8192, device="cuda")
for _ in range(40):
    a @ a                                                   # clocks up
torch.cuda.synchronize()
out = []
for P in (299, 2399, 9999):
    vel = torch.from_numpy(np.random.default_rng(0).normal(0, 1, (P, 17, 17, 2)).astype(np.float32)).cuda()
    u = core.prefix_displacements(vel)
    ref = np.concatenate([np.zeros((1, 17, 17, 2)), np.cumsum(vel.cpu().numpy().astype(np.float64), axis=0)])
    assert np.array_equal(u.cpu().numpy(), ref), "prefix differs from the sequential float64 sum"
    ts = []
    for r in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(30): core.prefix_displacements(vel)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 30 * 1000)
    out.append(f"P={P}: min {min(ts):.1f} med {sorted(ts)[3]:.1f} us")
print(" | ".join(out))
