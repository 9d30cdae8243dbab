"""Device-resident stages of the stabilization core (torch tensors in, torch tensors out).

Each method is one C-ABI call (``include/meshflow_b200.h``) on the current CUDA stream; PyTorch is
only the owner of device memory and streams.  Nothing here falls back to the CPU: the module cannot
be used without ``libmeshflow_b200.so`` and a CUDA device.

Shapes follow the reference (how4rd/meshflow ``meshflowstabilizer.py``, cited as mfs.py:N):
``u``/``s`` are ``(F, R+1, C+1, 2)`` float64 displacements, frames are ``(F, H, W, 3)`` uint8 BGR.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi


@dataclass(frozen=True)
class MeshSpec:
    width: int
    height: int
    rows: int = 16
    cols: int = 16

    @property
    def vertices(self) -> int:
        return (self.rows + 1) * (self.cols + 1)


def vertex_xy(mesh: MeshSpec) -> np.ndarray:
    """Rest positions of the mesh vertices, (V,2) float32 -- same expression as mfs.py:901-906
    (Python float: divide, multiply, ceil)."""
    out = np.empty((mesh.vertices, 2), dtype=np.float32)
    k = 0
    for r in range(mesh.rows + 1):
        for c in range(mesh.cols + 1):
            out[k, 0] = math.ceil((mesh.width - 1) * (c / mesh.cols))
            out[k, 1] = math.ceil((mesh.height - 1) * (r / mesh.rows))
            k += 1
    return out


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class DeviceCore:
    """The three hot-path subsystems on one GPU."""

    def __init__(self, mesh: MeshSpec, device=None, ellipse_rows=10, ellipse_cols=10,
                 radius=10, iterations=100, border_bgr=(0, 0, 255)):
        if not torch.cuda.is_available():
            raise RuntimeError("meshflow_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        self.lib = _cabi.load()
        self.mesh = mesh
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ellipse_rows, self.ellipse_cols = int(ellipse_rows), int(ellipse_cols)
        self.radius, self.iterations = int(radius), int(iterations)
        self.border_bgr = tuple(int(v) for v in border_bgr)
        self.vertex_xy_host = vertex_xy(mesh)
        self.vertex_xy = torch.from_numpy(self.vertex_xy_host).to(self.device)
        self._ws = {}

    # -- scratch --------------------------------------------------------------------------------
    def _workspace(self, key, nbytes):
        nbytes = max(int(nbytes), 256)
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    # -- (1) vertex motion ------------------------------------------------------------------------
    def vertex_velocities(self, early_xy, late_xy, offset_xy, keep, pair_start, homographies,
                          pair_start_host=None, return_counts=False):
        """Batched over P pairs.  Feature tensors are the concatenation over pairs; ``pair_start`` is a
        (P+1,) int32 device tensor and ``pair_start_host`` the same array on the host (NumPy int32 or a
        CPU tensor; it lets the library pick the fast path).  Returns (P, R+1, C+1, 2) float32
        (mfs.py:287-362)."""
        m = self.mesh
        P = int(pair_start.numel()) - 1
        N = int(early_xy.shape[0])
        vel = torch.empty((P, m.rows + 1, m.cols + 1, 2), dtype=torch.float32, device=self.device)
        counts = torch.empty((P, m.vertices), dtype=torch.int32, device=self.device) if return_counts else None
        ws = self._workspace("vm", self.lib.mf_vertex_motion_workspace_bytes(N, P, m.rows, m.cols))
        host_ptr = None
        if pair_start_host is not None:
            if isinstance(pair_start_host, torch.Tensor):
                pair_start_host = pair_start_host.numpy()
            pair_start_host = np.ascontiguousarray(pair_start_host, dtype=np.int32)
            host_ptr = pair_start_host.ctypes.data
        _cabi.check(self.lib.mf_vertex_motion(
            _ptr(early_xy), _ptr(late_xy), _ptr(offset_xy), _ptr(keep), _ptr(pair_start), host_ptr, N, P,
            _ptr(homographies), _ptr(self.vertex_xy), m.width, m.height, m.rows,
            m.cols, self.ellipse_rows, self.ellipse_cols, _ptr(vel), _ptr(counts), _ptr(ws), ws.numel(),
            _stream()))
        return (vel, counts) if return_counts else vel

    def prefix_displacements(self, vel, disp0=None):
        """(P, ...) float32 velocities -> (P+1, ...) float64 displacements (mfs.py:271, 281)."""
        P = int(vel.shape[0])
        n = int(vel[0].numel()) if P else int(disp0.numel())
        disp = torch.empty((P + 1,) + tuple(vel.shape[1:]), dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.mf_prefix_displacements(_ptr(vel), _ptr(disp0), _ptr(disp), P, n, _stream()))
        return disp

    # -- (2) Jacobi -------------------------------------------------------------------------------
    def stabilized_displacements(self, u, homographies, definition, vertex_range=None, out=None,
                                 return_lambda=False):
        """u: (F, R+1, C+1, 2) float64 -> s of the same shape (mfs.py:632-878).  ``vertex_range``
        (v0, v1) restricts the solve to a vertex shard; the rest of ``out`` is left untouched."""
        m = self.mesh
        F = int(u.shape[0])
        n_sys = int(u[0].numel())
        s = torch.empty_like(u) if out is None else out
        v0, v1 = (0, n_sys // 2) if vertex_range is None else vertex_range
        lam = torch.empty(F, dtype=torch.float64, device=self.device) if return_lambda else None
        ws = self._workspace("jac", self.lib.mf_jacobi_workspace_bytes(F, n_sys))
        _cabi.check(self.lib.mf_jacobi_solve(
            _ptr(u), _ptr(homographies), _ptr(s), F, n_sys, 2 * int(v0), 2 * int(v1), m.width, m.height,
            self.radius, self.iterations, int(definition), _ptr(lam), _ptr(ws), ws.numel(), _stream()))
        return (s, lam) if return_lambda else s

    # -- (3) warp + crop --------------------------------------------------------------------------
    def warp_frames(self, frames, u, s, out=None, return_maps=False):
        """frames (nf,H,W,3) uint8 + displacements of the same nf frames -> stabilized frames and the
        per-frame crop edges (nf,4) int32 [left, top, right, bottom] (mfs.py:909-1100)."""
        m = self.mesh
        nf = int(frames.shape[0])
        dst = torch.empty_like(frames) if out is None else out
        crop = torch.empty((nf, 4), dtype=torch.int32, device=self.device)
        maps = torch.empty((nf, m.height, m.width, 2), dtype=torch.float32, device=self.device) if return_maps else None
        ws = self._workspace("warp", self.lib.mf_warp_workspace_bytes(nf, m.width, m.height, m.rows, m.cols))
        b, g, r = self.border_bgr
        _cabi.check(self.lib.mf_warp_frames(
            _ptr(frames), _ptr(u), _ptr(s), _ptr(self.vertex_xy), nf, m.width, m.height, m.rows, m.cols,
            b, g, r, _ptr(dst), _ptr(crop), _ptr(maps), _ptr(ws), ws.numel(), _stream()))
        return (dst, crop, maps) if return_maps else (dst, crop)

    def warp_crop_bounds(self, u, s):
        """Per-frame crop edges (nf,4) int32 exactly as ``warp_frames`` returns them, computed from the
        vertex paths alone (no pixel is read): pass A of the streamed schedule."""
        m = self.mesh
        nf = int(u.shape[0])
        crop = torch.empty((nf, 4), dtype=torch.int32, device=self.device)
        ws = self._workspace("warp", self.lib.mf_warp_workspace_bytes(nf, m.width, m.height, m.rows, m.cols))
        _cabi.check(self.lib.mf_warp_crop_bounds(_ptr(u), _ptr(s), _ptr(self.vertex_xy), nf, m.width, m.height,
                                                 m.rows, m.cols, _ptr(crop), _ptr(ws), ws.numel(), _stream()))
        return crop

    def combine_crop(self, per_frame_crop):
        """(nf,4) per-frame edges -> device int32[4] = [max left, max top, -min right, -min bottom]
        (mfs.py:1103-1106): encoded so that ONE max-reduction -- and one all_reduce(MAX) across GPUs --
        combines it."""
        enc = torch.empty(4, dtype=torch.int32, device=self.device)
        _cabi.check(self.lib.mf_crop_combine(_ptr(per_frame_crop), int(per_frame_crop.shape[0]), _ptr(enc), _stream()))
        return enc

    @staticmethod
    def decode_crop(enc):
        """Encoded device crop -> (left, top, right, bottom) Python ints (synchronises)."""
        l, t, nr, nb = (int(v) for v in enc.tolist())
        return (l, t, -nr, -nb)

    def crop_resize_device(self, frames, crop_enc, out=None):
        """``crop_resize`` with the rectangle taken from device memory (no host round trip)."""
        m = self.mesh
        nf = int(frames.shape[0])
        dst = torch.empty_like(frames) if out is None else out
        ws = self._workspace("resize", self.lib.mf_crop_resize_workspace_bytes(m.width, m.height))
        _cabi.check(self.lib.mf_crop_resize_device(_ptr(frames), nf, m.width, m.height, _ptr(crop_enc), _ptr(dst),
                                                   _ptr(ws), ws.numel(), _stream()))
        return dst

    def crop_resize(self, frames, crop, out=None):
        """crop = (left, top, right, bottom) inclusive Python ints (mfs.py:1111-1157)."""
        m = self.mesh
        nf = int(frames.shape[0])
        dst = torch.empty_like(frames) if out is None else out
        ws = self._workspace("resize", self.lib.mf_crop_resize_workspace_bytes(m.width, m.height))
        l, t, r, b = (int(v) for v in crop)
        _cabi.check(self.lib.mf_crop_resize(_ptr(frames), nf, m.width, m.height, l, t, r, b, _ptr(dst),
                                            _ptr(ws), ws.numel(), _stream()))
        return dst

    # -- stability score ----------------------------------------------------------------------------
    def stability_score(self, s):
        """mfs.py:1216-1259: mean over vertices of the x ratio and of the y ratio, averaged."""
        F = int(s.shape[0])
        n_sys = int(s[0].numel())
        ratio = torch.empty(n_sys, dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.mf_stability_ratios(_ptr(s), F, n_sys, _ptr(ratio), _stream()))
        r = ratio.view(-1, 2)
        return (r[:, 0].mean() + r[:, 1].mean()) / 2.0


class StreamedCore:
    """Host buffers in, host buffers out, with the copies hidden behind the kernels.

    Schedule for one video (F frames in pinned host memory):

    1. tracks H2D -> vertex motion -> prefix sum -> Jacobi            (needs no pixels, ~1 ms)
    2. ``warp_crop_bounds`` over all frames -> crop rectangle          (needs no pixels either)
    3. per chunk of frames, on three streams: H2D chunk | warp + crop/resize chunk | D2H chunk

    so the PCIe link runs in both directions at once and the stabilized (uncropped) frames never
    exist beyond one chunk.  Results are identical to the unstreamed stage sequence.
    """

    def __init__(self, core: DeviceCore, chunk_frames=16):
        self.core = core
        self.chunk = int(chunk_frames)
        dev = core.device
        self.copy_in = torch.cuda.Stream(device=dev)
        self.copy_out = torch.cuda.Stream(device=dev)
        self._bufs = None

    def _buffers(self, n_slots, shape):
        key = (n_slots, tuple(shape))
        if self._bufs is None or self._bufs[0] != key:
            dev = self.core.device
            mk = lambda: [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(n_slots)]
            self._bufs = (key, mk(), mk(), torch.empty(shape, dtype=torch.uint8, device=dev))
        return self._bufs[1], self._bufs[2], self._bufs[3]

    def run(self, h_frames, tracks, h_out, definition):
        """h_frames / h_out: pinned (F,H,W,3) uint8 CPU tensors holding THIS RANK's frames.  tracks:
        dict of pinned CPU tensors (early, late, offset, keep, pair_start, homographies[P,9]) of this
        rank's frame pairs (at least F-1 of them; F when another rank's frames follow).
        Returns (crop_enc device tensor, u, s of the whole video) -- all work is enqueued, nothing is
        synchronised."""
        from . import distributed as mfd
        core = self.core
        dev = core.device
        F = int(h_frames.shape[0])
        main = torch.cuda.current_stream(dev)
        rank, _ = mfd.world_info()
        # the chunk buffers are reused from call to call: uploads may not start before the work already queued on
        # the caller's stream (e.g. the previous video's last warp) is done with them
        self.copy_in.wait_stream(main)
        # 1. paths of the whole video (all-gathers / vertex-sharded solve when there are several ranks)
        tr = {k: v.to(dev, non_blocking=True) for k, v in tracks.items()}
        u_all, s_all, _ = mfd.sharded_paths(core, tr, F, definition, pair_start_host=tracks["pair_start"])
        u, s = u_all[rank * F:(rank + 1) * F], s_all[rank * F:(rank + 1) * F]
        # 2. crop rectangle of the whole video: local bounds, then one MAX all-reduce
        enc = mfd.reduce_crop(core.combine_crop(core.warp_crop_bounds(u, s)))
        # 3. chunked, triple-buffered pixel pass
        n_slots = 3
        cshape = (self.chunk,) + tuple(h_frames.shape[1:])
        ins, outs, stab = self._buffers(n_slots, cshape)
        in_ready = [torch.cuda.Event() for _ in range(n_slots)]
        in_free = [None] * n_slots
        out_ready = [torch.cuda.Event() for _ in range(n_slots)]
        out_free = [None] * n_slots
        starts = list(range(0, F, self.chunk))
        for ci, f0 in enumerate(starts):
            n = min(self.chunk, F - f0)
            slot = ci % n_slots
            with torch.cuda.stream(self.copy_in):
                if in_free[slot] is not None:
                    self.copy_in.wait_event(in_free[slot])      # the warp that last read this slot is done
                ins[slot][:n].copy_(h_frames[f0:f0 + n], non_blocking=True)
                in_ready[slot].record(self.copy_in)
            main.wait_event(in_ready[slot])
            if out_free[slot] is not None:
                main.wait_event(out_free[slot])                 # the D2H that last read this slot is done
            core.warp_frames(ins[slot][:n], u[f0:f0 + n], s[f0:f0 + n], out=stab[:n])
            e = torch.cuda.Event(); e.record(main); in_free[slot] = e
            core.crop_resize_device(stab[:n], enc, out=outs[slot][:n])
            out_ready[slot].record(main)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(out_ready[slot])
                h_out[f0:f0 + n].copy_(outs[slot][:n], non_blocking=True)
                e2 = torch.cuda.Event(); e2.record(self.copy_out); out_free[slot] = e2
        main.wait_stream(self.copy_out)
        return enc, u_all, s_all
