"""Debug helper: fast warp path vs generic kernel on one synthetic case; prints the differing pixels."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tests import synth
from meshflow_b200 import DeviceCore, MeshSpec

W, H, R, C, amp = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]))
rng = np.random.default_rng(W + R)
frames, u, s = synth.synthetic_warp_inputs(rng, 1, W, H, R, C, per_vertex=amp, per_frame=1.2 * amp)
core = DeviceCore(MeshSpec(W, H, R, C), border_bgr=(7, 99, 250))
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(core.device)
gen, crop_gen, maps = core.warp_frames(d(frames), d(u), d(s), return_maps=True)
fast, crop_fast = core.warp_frames(d(frames), d(u), d(s))
bad = (gen != fast).any(dim=3)[0].nonzero().cpu().numpy()
print("differ:", len(bad), crop_gen.tolist(), crop_fast.tolist())
maps = maps.cpu().numpy()[0]; g = gen.cpu().numpy()[0]; f = fast.cpu().numpy()[0]
for y, x in bad[:40]:
    print(x, y, x % 4, "gen", g[y, x], "fast", f[y, x], "map", maps[y, x], maps[y, x] * 32)
