"""Times mf_jacobi_solve at the c2 and c4 sizes (CUDA events, 3 repetitions) and prints f64 GFLOP/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshflow_b200 import DeviceCore, MeshSpec
for (F, R, radius, iters) in [(300, 16, 10, 100), (1000, 32, 10, 100), (10000, 64, 30, 500)]:
    core = DeviceCore(MeshSpec(1920, 1080, R, R), radius=radius, iterations=iters)
    V = (R + 1) ** 2
    u = torch.cumsum(torch.randn((F, R + 1, R + 1, 2), device=core.device, dtype=torch.float64), dim=0)
    homs = torch.eye(3, dtype=torch.float64, device=core.device).reshape(1, 9).repeat(F, 1)
    s = torch.empty_like(u)
    core.stabilized_displacements(u, homs, 0, out=s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        core.stabilized_displacements(u, homs, 0, out=s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flop = iters * V * F * 2 * (2 * (2 * radius + 1) + 3)
    print(f"F={F} mesh={R}x{R} radius={radius} iters={iters}: {ms:.3f} ms  {flop / ms / 1e6:.1f} GFLOP/s f64  "
          f"{32.0 * V * F / ms / 1e6:.2f} GB/s algorithmic HBM")
