"""Device-resident stages of the stabilization core (torch tensors in, torch tensors out).

Each method is one C-ABI call (``include/meshflow_b200.h``) on the current CUDA stream; PyTorch is
only the owner of device memory and streams.  Nothing here falls back to the CPU: the module cannot
be used without ``libmeshflow_b200.so`` and a CUDA device.

Shapes follow the reference (how4rd/meshflow ``meshflowstabilizer.py``, cited as mfs.py:N):
``u``/``s`` are ``(F, R+1, C+1, 2)`` float64 displacements, frames are ``(F, H, W, 3)`` uint8 BGR.
"""
from __future__ import annotations

import functools
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


def _on_device(method):
    """Run a ``DeviceCore`` method with the core's GPU as the current CUDA device: the kernels are launched
    on THAT device's current stream whatever device the caller happens to have selected."""
    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return method(self, *args, **kwargs)
    return wrapper


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

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- scratch --------------------------------------------------------------------------------
    def _workspace(self, key, nbytes):
        nbytes = max(int(nbytes), 256)
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    # -- (1) vertex motion ------------------------------------------------------------------------
    @_on_device
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
            self._stream()))
        return (vel, counts) if return_counts else vel

    @_on_device
    def prefix_displacements(self, vel, disp0=None):
        """(P, ...) float32 velocities -> (P+1, ...) float64 displacements (mfs.py:271, 281)."""
        P = int(vel.shape[0])
        n = int(vel[0].numel()) if P else int(disp0.numel())
        disp = torch.empty((P + 1,) + tuple(vel.shape[1:]), dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.mf_prefix_displacements(_ptr(vel), _ptr(disp0), _ptr(disp), P, n, self._stream()))
        return disp

    # -- (2) Jacobi -------------------------------------------------------------------------------
    @_on_device
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
            self.radius, self.iterations, int(definition), _ptr(lam), _ptr(ws), ws.numel(), self._stream()))
        return (s, lam) if return_lambda else s

    # -- (3) warp + crop --------------------------------------------------------------------------
    @_on_device
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
            b, g, r, _ptr(dst), _ptr(crop), _ptr(maps), _ptr(ws), ws.numel(), self._stream()))
        return (dst, crop, maps) if return_maps else (dst, crop)

    @_on_device
    def warp_prepare(self, u, s, key="warp"):
        """Pass A: per-frame crop edges (nf,4) int32 exactly as ``warp_frames`` returns them, computed from the
        vertex paths alone (no pixel is read), plus the row-segment tables of these nf frames, which stay in
        the core's workspace ``key`` for ``warp_resize_frames``.  Returns (crop, tables handle)."""
        m = self.mesh
        nf = int(u.shape[0])
        crop = torch.empty((nf, 4), dtype=torch.int32, device=self.device)
        ws = self._workspace(key, self.lib.mf_warp_workspace_bytes(nf, m.width, m.height, m.rows, m.cols))
        _cabi.check(self.lib.mf_warp_prepare(_ptr(u), _ptr(s), _ptr(self.vertex_xy), nf, m.width, m.height,
                                             m.rows, m.cols, _ptr(crop), _ptr(ws), ws.numel(), self._stream()))
        return crop, (ws, nf)

    def warp_crop_bounds(self, u, s):
        """Round-1 name of pass A: only the per-frame crop edges."""
        return self.warp_prepare(u, s)[0]

    @_on_device
    def warp_resize_frames(self, frames, crop_enc, tables, first_frame=0, out=None):
        """Pass B, fused: frames ``first_frame ..`` of the video ``tables`` was prepared for -> cropped frames
        stretched back to the frame size, one kernel, no stabilized intermediate in DRAM."""
        m = self.mesh
        nf = int(frames.shape[0])
        dst = torch.empty_like(frames) if out is None else out
        ws, table_frames = tables
        rws = self._workspace("resize", self.lib.mf_crop_resize_workspace_bytes(m.width, m.height))
        b, g, r = self.border_bgr
        _cabi.check(self.lib.mf_warp_resize_frames(
            _ptr(frames), nf, int(first_frame), int(table_frames), m.width, m.height, m.rows, m.cols, b, g, r,
            _ptr(crop_enc), _ptr(dst), _ptr(ws), ws.numel(), _ptr(rws), rws.numel(), self._stream()))
        return dst

    @property
    def fused_pass_available(self):
        """The fused kernel needs the row-segment tables (frames at least 16 pixels wide)."""
        m = self.mesh
        return m.width >= 16 and m.height >= 2 and m.width <= 32767 and m.height <= 32767 and m.rows * m.cols < 0xfffd

    @_on_device
    def combine_crop(self, per_frame_crop):
        """(nf,4) per-frame edges -> device int32[4] = [max left, max top, -min right, -min bottom]
        (mfs.py:1103-1106): encoded so that ONE max-reduction -- and one all_reduce(MAX) across GPUs --
        combines it."""
        enc = torch.empty(4, dtype=torch.int32, device=self.device)
        _cabi.check(self.lib.mf_crop_combine(_ptr(per_frame_crop), int(per_frame_crop.shape[0]), _ptr(enc), self._stream()))
        return enc

    @staticmethod
    def decode_crop(enc):
        """Encoded device crop -> (left, top, right, bottom) Python ints (synchronises)."""
        l, t, nr, nb = (int(v) for v in enc.tolist())
        return (l, t, -nr, -nb)

    @_on_device
    def crop_resize_device(self, frames, crop_enc, out=None):
        """``crop_resize`` with the rectangle taken from device memory (no host round trip)."""
        m = self.mesh
        nf = int(frames.shape[0])
        dst = torch.empty_like(frames) if out is None else out
        ws = self._workspace("resize", self.lib.mf_crop_resize_workspace_bytes(m.width, m.height))
        _cabi.check(self.lib.mf_crop_resize_device(_ptr(frames), nf, m.width, m.height, _ptr(crop_enc), _ptr(dst),
                                                   _ptr(ws), ws.numel(), self._stream()))
        return dst

    @_on_device
    def crop_resize(self, frames, crop, out=None):
        """crop = (left, top, right, bottom) inclusive Python ints (mfs.py:1111-1157)."""
        m = self.mesh
        nf = int(frames.shape[0])
        dst = torch.empty_like(frames) if out is None else out
        ws = self._workspace("resize", self.lib.mf_crop_resize_workspace_bytes(m.width, m.height))
        l, t, r, b = (int(v) for v in crop)
        _cabi.check(self.lib.mf_crop_resize(_ptr(frames), nf, m.width, m.height, l, t, r, b, _ptr(dst),
                                            _ptr(ws), ws.numel(), self._stream()))
        return dst

    def warp_crop_resize(self, frames, u, s, crop_enc, out=None, tables=None, first_frame=0):
        """Pass B of the streamed schedule: frames + displacements + the (already known) crop rectangle ->
        cropped frames stretched back to the frame size (mfs.py:909-1100 followed by 1111-1157).  ``tables``:
        handle of a ``warp_prepare`` call that covered these frames (they are rebuilt for the chunk otherwise)."""
        nf = int(frames.shape[0])
        if self.fused_pass_available:
            if tables is None:
                _, tables = self.warp_prepare(u, s, key="warp_chunk")
                first_frame = 0
            return self.warp_resize_frames(frames, crop_enc, tables, first_frame, out=out)
        key = ("stab", tuple(frames.shape[1:]))
        buf = self._ws.get(key)
        if buf is None or buf.shape[0] < nf:
            buf = torch.empty((nf,) + tuple(frames.shape[1:]), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        self.warp_frames(frames, u, s, out=buf[:nf])
        return self.crop_resize_device(buf[:nf], crop_enc, out=out)

    # -- stability score ----------------------------------------------------------------------------
    @_on_device
    def stability_score(self, s):
        """mfs.py:1216-1259: mean over vertices of the x ratio and of the y ratio, averaged."""
        F = int(s.shape[0])
        n_sys = int(s[0].numel())
        ratio = torch.empty(n_sys, dtype=torch.float64, device=self.device)
        _cabi.check(self.lib.mf_stability_ratios(_ptr(s), F, n_sys, _ptr(ratio), self._stream()))
        r = ratio.view(-1, 2)
        return (r[:, 0].mean() + r[:, 1].mean()) / 2.0


class StreamedCore:
    """Host buffers in, host buffers out, with the copies hidden behind the kernels.

    Schedule for one video (F frames in pinned host memory):

    1. tracks H2D -> vertex motion -> prefix sum -> Jacobi            (needs no pixels, ~1 ms)
    2. ``warp_crop_bounds`` chunk by chunk -> crop rectangle           (needs no pixels either)
    3. per chunk of frames, on three streams: H2D chunk | warp + crop/resize chunk | D2H chunk

    so the PCIe link runs in both directions at once and the stabilized (uncropped) frames never
    exist beyond one chunk; device scratch is O(chunk), not O(video).  Results are identical to the
    unstreamed stage sequence.  One ``StreamedCore`` is meant to live as long as its ``DeviceCore``:
    the chunk buffers are allocated once and reused from call to call.
    """

    N_SLOTS = 3

    def __init__(self, core: DeviceCore, chunk_frames=16, table_budget_bytes=4 << 30):
        self.core = core
        self.chunk = int(chunk_frames)
        self.table_budget_bytes = int(table_budget_bytes)
        dev = core.device
        self.copy_in = torch.cuda.Stream(device=dev)
        self.copy_out = torch.cuda.Stream(device=dev)
        self._bufs = None

    def _buffers(self, shape):
        key = tuple(shape)
        if self._bufs is None or self._bufs[0] != key:
            dev = self.core.device
            mk = lambda: [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(self.N_SLOTS)]
            self._bufs = (key, mk(), mk())
        return self._bufs[1], self._bufs[2]

    def chunk_schedule(self, F):
        """(first frame, count) of every chunk: full chunks in the middle, short ones at both ends -- the first
        upload and the last download are the only copies nothing overlaps with, so they are kept small."""
        c = self.chunk
        ramp = [max(1, c // 4), max(1, c // 4), max(1, c // 2)]
        if F <= 4 * c:
            sizes = [min(c, F - f0) for f0 in range(0, F, c)]
        else:
            body = F - 2 * sum(ramp)
            sizes = ramp + [c] * (body // c) + ([body % c] if body % c else []) + ramp[::-1]
        out, f0 = [], 0
        for n in sizes:
            out.append((f0, n))
            f0 += n
        assert f0 == F
        return out

    def crop_of_video(self, u, s, plan=None):
        """Pass A: encoded crop rectangle of the whole video from the vertex paths of THIS rank's frames, then
        one MAX all-reduce across the plan.  When the row-segment tables of all frames fit ``table_budget_bytes``
        they are built once and kept for pass B (returned handle); longer videos are evaluated chunk by
        chunk (scratch stays O(chunk)) and pass B rebuilds the tables of each chunk."""
        from . import distributed as mfd
        core = self.core
        m = core.mesh
        enc, tables = None, None
        F = int(u.shape[0])
        if F and core.lib.mf_warp_workspace_bytes(F, m.width, m.height, m.rows, m.cols) <= self.table_budget_bytes:
            crop, tables = core.warp_prepare(u, s)
            enc = core.combine_crop(crop)
        else:
            step = max(self.chunk, 64)
            for f0 in range(0, F, step):
                part = core.combine_crop(core.warp_prepare(u[f0:f0 + step], s[f0:f0 + step], key="warp_chunk")[0])
                enc = part if enc is None else torch.maximum(enc, part)
        if enc is None:       # a rank without frames contributes the identity of the max-reduction
            m = core.mesh
            enc = torch.tensor([0, 0, -(m.width - 1), -(m.height - 1)], dtype=torch.int32, device=core.device)
        return mfd.reduce_crop(enc, plan), tables

    def run(self, h_frames, tracks, h_out, definition, plan=None, d_frames=None, on_chunk_landed=None,
            return_homographies=False):
        """h_frames / h_out: pinned (F,H,W,3) uint8 CPU tensors holding THIS RANK's frames.  tracks:
        dict of pinned CPU tensors (early, late, offset, keep, pair_start, homographies[P,9]) of the
        frame pairs that start at this rank's frames (F-1 of them when the video ends here, F when another
        rank's frames follow).  ``plan``: a ``distributed.ShardPlan`` when the video is sharded over
        several ranks -- an explicit opt-in; ``None`` treats the frames as a whole video even when
        ``torch.distributed`` happens to be initialised.  ``d_frames``: the same frames already resident
        on the device (then ``h_frames`` is not read and nothing is uploaded).
        Returns (crop_enc device tensor, u, s of the whole video[, homographies]).  Without
        ``on_chunk_landed`` all work is only enqueued, nothing is synchronised; with it the call waits for
        each chunk's device-to-host copy in turn and calls ``on_chunk_landed(first_frame, count)`` as soon
        as those frames are in ``h_out`` (the remaining chunks are still in flight)."""
        from . import distributed as mfd
        core = self.core
        dev = core.device
        F = int(h_out.shape[0])
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            # chunk buffers first: they are reused from call to call, and copies may not start before the
            # work already queued on the caller's stream (e.g. the previous video's last warp) is done with them
            cshape = (self.chunk,) + tuple(h_out.shape[1:])
            ins, outs = self._buffers(cshape)
            self.copy_in.wait_stream(main)
            self.copy_out.wait_stream(main)
            # 1. paths of the whole video (one all-gather + vertex-sharded solve when there is a plan)
            tr = {k: v.to(dev, non_blocking=True) for k, v in tracks.items()}
            u_all, s_all, homs = mfd.sharded_paths(core, tr, F, definition, pair_start_host=tracks["pair_start"],
                                                   plan=plan)
            first = plan.first_frame if plan is not None else 0
            u, s = u_all[first:first + F], s_all[first:first + F]
            if int(u.shape[0]) != F or int(s.shape[0]) != F:
                raise ValueError(f"paths cover {int(u.shape[0])} of this rank's {F} frames")
            # 2. crop rectangle of the whole video
            enc, tables = self.crop_of_video(u, s, plan)
            # 3. chunked, triple-buffered pixel pass
            n_slots = self.N_SLOTS
            in_ready = [torch.cuda.Event() for _ in range(n_slots)]
            in_free = [None] * n_slots
            out_ready = [torch.cuda.Event() for _ in range(n_slots)]
            out_free = [None] * n_slots
            landed = []
            for ci, (f0, n) in enumerate(self.chunk_schedule(F)):
                slot = ci % n_slots
                if d_frames is None:
                    with torch.cuda.stream(self.copy_in):
                        if in_free[slot] is not None:
                            self.copy_in.wait_event(in_free[slot])      # the warp that last read this slot is done
                        ins[slot][:n].copy_(h_frames[f0:f0 + n], non_blocking=True)
                        in_ready[slot].record(self.copy_in)
                    main.wait_event(in_ready[slot])
                    src = ins[slot][:n]
                else:
                    src = d_frames[f0:f0 + n]
                if out_free[slot] is not None:
                    main.wait_event(out_free[slot])                 # the D2H that last read this slot is done
                core.warp_crop_resize(src, u[f0:f0 + n], s[f0:f0 + n], enc, out=outs[slot][:n], tables=tables,
                                      first_frame=f0)
                if d_frames is None:
                    e = torch.cuda.Event(); e.record(main); in_free[slot] = e
                out_ready[slot].record(main)
                with torch.cuda.stream(self.copy_out):
                    self.copy_out.wait_event(out_ready[slot])
                    h_out[f0:f0 + n].copy_(outs[slot][:n], non_blocking=True)
                    e2 = torch.cuda.Event(); e2.record(self.copy_out); out_free[slot] = e2
                landed.append((e2, f0, n))
            main.wait_stream(self.copy_out)
            if on_chunk_landed is not None:
                for e2, f0, n in landed:
                    e2.synchronize()
                    on_chunk_landed(f0, n)
        return (enc, u_all, s_all, homs) if return_homographies else (enc, u_all, s_all)
