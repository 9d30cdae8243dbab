import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from meshflow_b200 import DeviceCore, MeshSpec
from oracle import spec
from tests import synth
W,H,R,C=640,360,16,16
core = DeviceCore(MeshSpec(W,H,R,C))
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(core.device)
for n, keepp in [(3,1.0),(40,1.0),(1800,0.85),(1800,0.0)]:
    rng = np.random.default_rng(1)
    tr = synth.synthetic_tracks(rng, 2, n, W, H, keep_prob=keepp)
    vel, counts = core.vertex_velocities(dev(tr["early"]), dev(tr["late"]), dev(tr["offset"]), dev(tr["keep"]), dev(tr["pair_start"]), dev(tr["homographies"].reshape(-1,9)), pair_start_host=tr["pair_start"], return_counts=True)
    vel=vel.cpu().numpy(); counts=counts.cpu().numpy()
    for p in range(2):
        a,b = tr["pair_start"][p], tr["pair_start"][p+1]
        k = tr["keep"][a:b].astype(bool)
        off = tr["offset"][a:b][k].astype(np.float64)
        e = tr["early"][a:b][k].astype(np.float64)+off; l = tr["late"][a:b][k].astype(np.float64)+off
        ref, member = spec.vertex_velocities(e,l,tr["homographies"][p],W,H,R,C,10,10,return_assignment=True)
        cnt = member.sum(axis=0) if member is not None else np.zeros(289,int)
        d = np.abs(vel[p]-ref)
        print(n, keepp, p, 'counts equal', np.array_equal(counts[p],cnt), 'n feats', k.sum(), 'max diff', d.max(), 'n diff', (d>0).sum(), 'argmax', np.unravel_index(d.argmax(), d.shape), 'cnt at argmax', cnt.reshape(17,17)[np.unravel_index(d.argmax(), d.shape)[:2]])
