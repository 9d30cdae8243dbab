"""``MeshFlowStabilizer`` -- drop-in for how4rd/meshflow's class of the same name.

Same constructor arguments and defaults (``meshflowstabilizer.py`` lines 43-49, cited as mfs.py:N),
same four ``ADAPTIVE_WEIGHTS_DEFINITION_*`` constants (mfs.py:32-35), same
``stabilize(input_path, output_path, adaptive_weights_definition)`` entry point returning
``(cropping_ratio, distortion_score, stability_score)`` (mfs.py:102-169), and the same private stage
methods with the same signatures and return types, so stage-level tests read like tests of the
reference.  What changed is where the three data-parallel stages run:

* vertex-motion estimation  (mfs.py:236-452)  -> ``mf_vertex_motion`` + ``mf_prefix_displacements``
* Jacobi path optimisation  (mfs.py:632-878)  -> ``mf_jacobi_solve``
* mesh warp + crop          (mfs.py:909-1157) -> ``mf_warp_prepare`` + ``mf_warp_resize_frames``

all hand-written sm_100a kernels behind the C ABI in ``include/meshflow_b200.h``.  Video decode,
FAST/LK/RANSAC feature matching, the two OpenCV-derived metrics and video encode stay host OpenCV
code, as in the reference, so both implementations see identical correspondences.  There is no CPU
fallback for the three stages.

Host / device overlap: the OpenCV front end runs on a persistent thread pool, one
job per frame pair; while it runs, the frames are staged into pinned memory and uploaded; once the
paths and the crop rectangle are known the pixel pass runs chunk by chunk, and the metric tracking
(mfs.py:1160-1212) of a chunk's frames is handed to the pool the moment the chunk's device-to-host
copy has landed, so it overlaps the rest of the pixel pass.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np
import torch

from . import host_features
from .pipeline import DeviceCore, MeshSpec, StreamedCore


class MeshFlowStabilizer:
    ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL = 0
    ADAPTIVE_WEIGHTS_DEFINITION_FLIPPED = 1
    ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_HIGH = 2
    ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_LOW = 3

    ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_HIGH_VALUE = 100
    ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_LOW_VALUE = 1

    def __init__(self, mesh_row_count=16, mesh_col_count=16,
                 mesh_outlier_subframe_row_count=4, mesh_outlier_subframe_col_count=4,
                 feature_ellipse_row_count=10, feature_ellipse_col_count=10,
                 homography_min_number_corresponding_features=4,
                 temporal_smoothing_radius=10, optimization_num_iterations=100,
                 color_outside_image_area_bgr=(0, 0, 255),
                 visualize=False, *, device=None, host_workers=None, chunk_frames=16, distributed=False,
                 resident_limit_bytes=16 << 30):
        self.mesh_col_count = mesh_col_count
        self.mesh_row_count = mesh_row_count
        self.mesh_outlier_subframe_row_count = mesh_outlier_subframe_row_count
        self.mesh_outlier_subframe_col_count = mesh_outlier_subframe_col_count
        self.feature_ellipse_row_count = feature_ellipse_row_count
        self.feature_ellipse_col_count = feature_ellipse_col_count
        self.homography_min_number_corresponding_features = homography_min_number_corresponding_features
        self.temporal_smoothing_radius = temporal_smoothing_radius
        self.optimization_num_iterations = optimization_num_iterations
        self.color_outside_image_area_bgr = color_outside_image_area_bgr
        self.visualize = visualize
        self.feature_detector = cv2.FastFeatureDetector_create()
        # extensions (keyword only, not in the reference)
        self.device = device
        self.host_workers = host_workers
        self.chunk_frames = chunk_frames
        # distributed=True: under torchrun (one process per GPU, torch.distributed initialised) stabilize() shards
        # the video's frames over the ranks.  Off by default: a process that merely runs under torchrun
        # stabilizes its own whole video.
        self.distributed = bool(distributed)
        self.resident_limit_bytes = int(resident_limit_bytes)
        self._cores = {}
        self._streamed = {}
        self._pool = None
        self._pinned = {}
        self.last_timings = {}

    # ------------------------------------------------------------------------------------------
    # public entry point (mfs.py:102-169)
    # ------------------------------------------------------------------------------------------
    def stabilize(self, input_path, output_path,
                  adaptive_weights_definition=ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL):
        self._validate_definition(adaptive_weights_definition)
        plan = self._plan_for_video(input_path) if self.distributed else None
        if plan is not None and plan.world > 1:
            return self._stabilize_sharded(input_path, output_path, adaptive_weights_definition, plan)
        t0 = time.perf_counter()
        # decode (mfs.py:172-213) with the front end of every decoded pair already running on the pool
        pair_futures = []
        with host_features.single_threaded_opencv():
            frames, num_frames, fps, codec = self._get_unstabilized_frames_and_video_features(
                input_path, on_frame=lambda fr: self._submit_pair(fr, pair_futures))
        t_decode = time.perf_counter() - t0
        result = self.stabilize_frames(frames, adaptive_weights_definition, reuse_output=True, pair_futures=pair_futures)
        t0 = time.perf_counter()
        self._write_stabilized_video(output_path, num_frames, fps, codec, result["cropped_frames"])
        self.last_timings = dict(result["timings"], decode=t_decode, encode=time.perf_counter() - t0)
        if self.visualize:
            self._display_unstablilized_and_cropped_video_loop(num_frames, fps, frames, result["cropped_frames"])
        return (result["cropping_ratio"], result["distortion_score"], result["stability_score"])

    def stabilize_frames(self, unstabilized_frames, adaptive_weights_definition=ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL,
                         with_metrics=True, plan=None, lookahead_frame=None, reuse_output=False, pair_futures=None):
        """``stabilize()`` on in-memory frames: everything between decode and encode, device resident
        between the stages.  Returns a dict with the cropped frames, the crop rectangle, ``u``,
        ``homographies``, ``s`` (NumPy, whole video), the three metrics and per-stage wall times.

        ``plan`` (a ``distributed.ShardPlan``) + ``lookahead_frame`` (the next rank's first frame; None on
        the rank that owns the video's last frame): the frames are this rank's contiguous shard of a longer
        video; paths, crop rectangle and metrics are those of the whole video on every rank.

        ``reuse_output=True`` writes the cropped frames into a pinned buffer that this object keeps and
        reuses: the returned frames are then only valid until the next call (``stabilize()`` works this
        way -- it encodes them right away).  By default the frames live in a buffer of their own."""
        from . import distributed as mfd
        self._validate_definition(adaptive_weights_definition)
        t_start = time.perf_counter()
        timings = {}
        frames = unstabilized_frames
        F = len(frames)
        sharded = plan is not None and plan.world > 1
        if not sharded:
            plan = None
            if F < 2:
                raise ValueError("need at least two frames to stabilize")
        elif F == 0:
            raise ValueError("a rank without frames cannot take part (use fewer ranks than frames)")
        height, width = frames[0].shape[:2]
        core = self._core(width, height)
        streamed = self._streamed_for(core)
        dev = core.device
        pool = self._thread_pool()
        total_frames = plan.total_frames if sharded else F
        # -- 1. host front end on the pool: one job per frame pair that starts at one of this rank's frames
        late = list(frames[1:]) + ([lookahead_frame] if (sharded and lookahead_frame is not None) else [])
        n_pairs = len(late)
        if sharded and n_pairs < plan.pairs_needed():
            raise ValueError("this rank's last frame needs the next rank's first frame (lookahead_frame)")
        track = lambda a, b: host_features.track_pair(
            a, b, self.mesh_outlier_subframe_row_count, self.mesh_outlier_subframe_col_count,
            self.homography_min_number_corresponding_features)
        with host_features.single_threaded_opencv():
            # ``pair_futures``: the caller already started the front end of the consecutive pairs (stabilize() does,
            # while it decodes); only the look-ahead pair of a shard is still missing then
            futures = list(pair_futures) if pair_futures is not None else []
            futures += [pool.submit(track, frames[i], late[i]) for i in range(len(futures), n_pairs)]
            # -- 2. meanwhile: frames -> pinned memory -> device (resident when they fit, else streamed later)
            t0 = time.perf_counter()
            frame_bytes = height * width * 3
            h_in = self._pinned_buffer("in", (F, height, width, 3))
            h_out = (self._pinned_buffer("out", (F, height, width, 3)) if reuse_output
                     else torch.empty((F, height, width, 3), dtype=torch.uint8, pin_memory=True))
            resident = F * frame_bytes <= self.resident_limit_bytes
            view = h_in.numpy()
            with torch.cuda.device(dev):
                main = torch.cuda.current_stream(dev)
                d_frames = torch.empty((F, height, width, 3), dtype=torch.uint8, device=dev) if resident else None
                for f0 in range(0, F, self.chunk_frames):
                    n = min(self.chunk_frames, F - f0)
                    for i in range(f0, f0 + n):
                        view[i] = frames[i]
                    if resident:
                        with torch.cuda.stream(streamed.copy_in):
                            d_frames[f0:f0 + n].copy_(h_in[f0:f0 + n], non_blocking=True)
                if resident:
                    main.wait_stream(streamed.copy_in)
            timings["stage_frames"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            tracks = [f.result() for f in futures]
            timings["track_wait"] = time.perf_counter() - t0
            # -- 3. paths + crop rectangle, then the pixel pass chunk by chunk; metric tracking per landed chunk
            t0 = time.perf_counter()
            packed = self.pack_tracks(tracks)
            pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            h_tracks = {k: pin(packed[k]) for k in ("early", "late", "offset", "keep", "pair_start")}
            h_tracks["homographies"] = pin(packed["homographies"].reshape(-1, 9))
            cropped = h_out.numpy()
            metric_futures = []
            if with_metrics:
                def on_chunk(f0, n):
                    metric_futures.extend(pool.submit(track, frames[i], cropped[i]) for i in range(f0, f0 + n))
            else:
                on_chunk = None
            crop_enc, u_d, s_d, homs_d = streamed.run(h_in, h_tracks, h_out, adaptive_weights_definition, plan=plan,
                                                      d_frames=d_frames, on_chunk_landed=on_chunk, return_homographies=True)
            stability = core.stability_score(s_d) if with_metrics else None
            crop = self._crop_tuple(crop_enc)                    # synchronises
            torch.cuda.synchronize(dev)
            self._check_crop(crop, width, height)
            timings["device"] = time.perf_counter() - t0
            cropped_frames = [cropped[i] for i in range(F)]
            out = dict(cropped_frames=cropped_frames, crop_boundaries=crop, u=u_d.cpu().numpy(),
                       homographies=homs_d.cpu().numpy().reshape(total_frames, 3, 3), s=s_d.cpu().numpy())
            if with_metrics:
                t0 = time.perf_counter()
                ratios, scores = self._ratios_and_scores([f.result() for f in metric_futures])
                if sharded:
                    ratios, scores = self._gather_metric_arrays(ratios, scores, plan, dev)
                timings["metrics_wait"] = time.perf_counter() - t0
                out.update(cropping_ratio=np.mean(ratios), distortion_score=np.min(scores),
                           stability_score=np.float64(stability.item()))
        timings["total"] = time.perf_counter() - t_start
        out["timings"] = timings
        self.last_timings = timings
        return out

    # ------------------------------------------------------------------------------------------
    # multi-GPU: frames sharded over the ranks of torch.distributed (opt-in: distributed=True)
    # ------------------------------------------------------------------------------------------
    def _plan_for_video(self, input_path):
        import torch.distributed as dist
        from . import distributed as mfd
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        video = cv2.VideoCapture(input_path)
        num_frames = int(video.get(cv2.CAP_PROP_FRAME_COUNT))
        video.release()
        return mfd.ShardPlan.even(num_frames)

    def _stabilize_sharded(self, input_path, output_path, definition, plan):
        """Every rank decodes the file but keeps only its contiguous frame range (+ one look-ahead frame),
        tracks its own pairs, and runs the sharded core; cropped frames meet in one host shared-memory
        buffer (the ranks of one box share the host), rank 0 encodes.  Every rank returns the tuple."""
        import torch.distributed as dist
        from multiprocessing import shared_memory
        begin = plan.first_frame
        end = begin + plan.local_frames
        video = cv2.VideoCapture(input_path)
        num_frames = int(video.get(cv2.CAP_PROP_FRAME_COUNT))
        fps = video.get(cv2.CAP_PROP_FPS)
        codec = int(video.get(cv2.CAP_PROP_FOURCC))
        frames, lookahead = [], None
        for frame_index in range(min(end + 1, num_frames)):
            frame = self._get_next_frame(video)
            if frame is None:
                raise IOError(f'Video at <{input_path}> did not have frame {frame_index} of '
                              f'{num_frames} (indexed from 0).')
            if begin <= frame_index < end:
                frames.append(frame)
            elif frame_index == end:
                lookahead = frame
        video.release()
        result = self.stabilize_frames(frames, definition, plan=plan, lookahead_frame=lookahead, reuse_output=True)
        h, w = frames[0].shape[:2]
        # cropped frames of all ranks -> one shared host buffer, in frame order
        names = [None]
        shm = None
        if plan.rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=num_frames * h * w * 3)
            names[0] = shm.name
        dist.broadcast_object_list(names, src=0, group=plan.group)
        if plan.rank != 0:
            shm = shared_memory.SharedMemory(name=names[0])
        whole = np.ndarray((num_frames, h, w, 3), dtype=np.uint8, buffer=shm.buf)
        for i, f in enumerate(result["cropped_frames"]):
            whole[begin + i] = f
        dist.barrier(group=plan.group)
        if plan.rank == 0:
            self._write_stabilized_video(output_path, num_frames, fps, codec, whole)
        dist.barrier(group=plan.group)
        del whole
        shm.close()
        if plan.rank == 0:
            shm.unlink()
        return (result["cropping_ratio"], result["distortion_score"], result["stability_score"])

    @staticmethod
    def _gather_metric_arrays(ratios, scores, plan, dev):
        """Per-frame metric values of every rank, in frame order, on every rank."""
        from . import distributed as mfd
        local = torch.from_numpy(np.stack([ratios, scores], axis=1)).to(dev)       # [F_local, 2] float32
        allv = mfd.gather_velocities(local, list(plan.frames), plan.group).cpu().numpy()
        return np.ascontiguousarray(allv[:, 0]), np.ascontiguousarray(allv[:, 1])

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    @classmethod
    def _validate_definition(cls, definition):
        if not (definition == cls.ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL or
                definition == cls.ADAPTIVE_WEIGHTS_DEFINITION_FLIPPED or
                definition == cls.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_HIGH or
                definition == cls.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_LOW):
            raise ValueError(
                'Invalid value for `adaptive_weights_definition`. Expecting value of '
                '`MeshFlowStabilizer.ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL`, '
                '`MeshFlowStabilizer.ADAPTIVE_WEIGHTS_DEFINITION_FLIPPED`, '
                '`MeshFlowStabilizer.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_HIGH`, or'
                '`MeshFlowStabilizer.ADAPTIVE_WEIGHTS_DEFINITION_CONSTANT_LOW`.')

    def _core(self, width, height) -> DeviceCore:
        key = (int(width), int(height))
        core = self._cores.get(key)
        if core is None:
            core = DeviceCore(MeshSpec(int(width), int(height), self.mesh_row_count, self.mesh_col_count),
                              device=self.device, ellipse_rows=self.feature_ellipse_row_count,
                              ellipse_cols=self.feature_ellipse_col_count,
                              radius=self.temporal_smoothing_radius,
                              iterations=self.optimization_num_iterations,
                              border_bgr=self.color_outside_image_area_bgr)
            self._cores[key] = core
        return core

    def _streamed_for(self, core) -> StreamedCore:
        sc = self._streamed.get(id(core))
        if sc is None or sc.chunk != int(self.chunk_frames):
            sc = StreamedCore(core, self.chunk_frames)
            self._streamed[id(core)] = sc
        return sc

    def _track_job(self, early, late):
        return host_features.track_pair(early, late, self.mesh_outlier_subframe_row_count,
                                        self.mesh_outlier_subframe_col_count,
                                        self.homography_min_number_corresponding_features)

    def _submit_pair(self, frames_so_far, futures):
        """Decode-time hook: the pair (previous frame, newest frame) goes to the pool at once."""
        if len(frames_so_far) >= 2:
            futures.append(self._thread_pool().submit(self._track_job, frames_so_far[-2], frames_so_far[-1]))

    def _thread_pool(self) -> ThreadPoolExecutor:
        if self._pool is None:
            workers = self.host_workers
            if workers is None:
                workers = host_features.default_workers()
            self._pool = ThreadPoolExecutor(max_workers=max(1, int(workers)), thread_name_prefix="meshflow-host")
        return self._pool

    def _pinned_buffer(self, name, shape):
        """Pinned staging buffers live as long as the stabilizer and are reused by later calls of the same
        size (page-locking ~2 GB costs about as much as the whole GPU pass)."""
        key = (name, tuple(shape))
        buf = self._pinned.get(key)
        if buf is None:
            for k in [k for k in self._pinned if k[0] == name]:
                del self._pinned[k]
            buf = torch.empty(shape, dtype=torch.uint8, pin_memory=True)
            self._pinned[key] = buf
        return buf

    def _track(self, early_frames, late_frames):
        return host_features.track_all_pairs(
            early_frames, late_frames, self.mesh_outlier_subframe_row_count,
            self.mesh_outlier_subframe_col_count, self.homography_min_number_corresponding_features,
            workers=self.host_workers, pool=self._thread_pool())

    @staticmethod
    def _upload_frames(core, frames):
        n = len(frames)
        h, w = frames[0].shape[:2]
        host = torch.empty((n, h, w, 3), dtype=torch.uint8, pin_memory=True)
        view = host.numpy()
        for i, f in enumerate(frames):
            view[i] = f
        return host.to(core.device, non_blocking=True)

    @staticmethod
    def pack_tracks(tracks):
        """Concatenate per-pair tracks into the flat arrays ``mf_vertex_motion`` takes."""
        counts = np.array([len(t.keep) for t in tracks], dtype=np.int64)
        start = np.zeros(len(tracks) + 1, dtype=np.int32)
        np.cumsum(counts, out=start[1:])
        cat = lambda name, dt, tail: (np.concatenate([getattr(t, name) for t in tracks]).astype(dt, copy=False)
                                      if len(tracks) else np.zeros((0,) + tail, dt))
        homs = (np.stack([t.homography for t in tracks]).astype(np.float64) if len(tracks)
                else np.zeros((0, 3, 3), np.float64))
        return dict(early=cat("early_xy", np.float32, (2,)), late=cat("late_xy", np.float32, (2,)),
                    offset=cat("offset_xy", np.int32, (2,)), keep=cat("keep", np.uint8, ()),
                    pair_start=start, homographies=homs, max_pair=int(counts.max()) if len(counts) else 0)

    def _device_displacements(self, core, tracks, return_velocities=False):
        """Vertex velocities of all pairs + float64 prefix sum, on the device (mfs.py:236-362)."""
        m = core.mesh
        num_frames = len(tracks) + 1
        homs = np.empty((num_frames, 3, 3))
        homs[-1] = np.identity(3)                                            # mfs.py:273-274
        if not tracks:
            u = torch.zeros((1, m.rows + 1, m.cols + 1, 2), dtype=torch.float64, device=core.device)
            return (u, homs, None) if return_velocities else (u, homs)
        p = self.pack_tracks(tracks)
        homs[:-1] = p["homographies"]
        dev = core.device
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        vel = core.vertex_velocities(to(p["early"]), to(p["late"]), to(p["offset"]), to(p["keep"]),
                                     to(p["pair_start"]), to(p["homographies"].reshape(-1, 9)),
                                     pair_start_host=p["pair_start"])
        u = core.prefix_displacements(vel)
        return (u, homs, vel) if return_velocities else (u, homs)

    @staticmethod
    def _check_crop(crop, width, height):
        l, t, r, b = (int(v) for v in crop)
        if not (0 <= l <= r < width and 0 <= t <= b < height):
            # the reference slices an empty array here and cv2.resize raises (mfs.py:1150-1151)
            raise ValueError(f"empty crop rectangle {(l, t, r, b)}: the stabilized frames share no common area")

    @staticmethod
    def _crop_tuple(combined):
        l, t, nr, nb = (int(v) for v in combined.tolist())
        return (np.int64(l), np.int64(t), np.int64(-nr), np.int64(-nb))

    @staticmethod
    def _ratios_and_scores(tracks):
        """mfs.py:1199-1210 on the homographies of already tracked (unstabilized, cropped) pairs."""
        cropping_ratios = np.empty((len(tracks)), dtype=np.float32)
        distortion_scores = np.empty((len(tracks)), dtype=np.float32)
        for i, t in enumerate(tracks):
            hom = t.homography
            cropping_ratios[i] = 1 / (hom[0][0] * hom[1][1])
            affine = np.copy(hom)
            affine[2] = [0, 0, 1]
            mags = np.sort(np.abs(np.linalg.eigvals(affine)))
            distortion_scores[i] = mags[-2] / mags[-1]
        return cropping_ratios, distortion_scores

    # ------------------------------------------------------------------------------------------
    # reference-named stage methods (NumPy in / NumPy out)
    # ------------------------------------------------------------------------------------------
    def _get_unstabilized_frames_and_video_features(self, input_path, on_frame=None):
        """mfs.py:172-213 (host code, unchanged behaviour).  ``on_frame(frames_so_far)`` is called after every
        decoded frame (stabilize() starts the feature tracking of the newest pair there)."""
        video = cv2.VideoCapture(input_path)
        num_frames = int(video.get(cv2.CAP_PROP_FRAME_COUNT))
        fps = video.get(cv2.CAP_PROP_FPS)
        codec = int(video.get(cv2.CAP_PROP_FOURCC))
        frames = []
        for frame_index in range(num_frames):
            frame = self._get_next_frame(video)
            if frame is None:
                raise IOError(f'Video at <{input_path}> did not have frame {frame_index} of '
                              f'{num_frames} (indexed from 0).')
            frames.append(frame)
            if on_frame is not None:
                on_frame(frames)
        video.release()
        return (frames, num_frames, fps, codec)

    def _get_next_frame(self, video):
        ok, pixels = video.read()
        return pixels if ok else None

    def _get_vertex_x_y(self, frame_width, frame_height):
        """mfs.py:881-906: (V,1,2) float32."""
        return self._core(frame_width, frame_height).vertex_xy_host.reshape(-1, 1, 2).copy()

    def _get_matched_features_and_homography(self, early_frame, late_frame):
        """mfs.py:455-528: compacted float64 inliers (N,1,2) and the global homography."""
        t = host_features.track_pair(early_frame, late_frame, self.mesh_outlier_subframe_row_count,
                                     self.mesh_outlier_subframe_col_count,
                                     self.homography_min_number_corresponding_features)
        e, l = t.compacted()
        return (e[:, np.newaxis, :], l[:, np.newaxis, :], t.homography)

    def _get_unstabilized_vertex_velocities(self, early_frame, late_frame):
        """mfs.py:287-362 for one pair -> ((R+1,C+1,2) float32, (3,3) float64)."""
        h, w = early_frame.shape[:2]
        core = self._core(w, h)
        tracks = self._track([early_frame], [late_frame])
        _, homs, vel = self._device_displacements(core, tracks, return_velocities=True)
        return (vel[0].cpu().numpy(), homs[0])

    def _get_unstabilized_vertex_displacements_and_homographies(self, num_frames, unstabilized_frames):
        """mfs.py:236-284 -> (u (F,R+1,C+1,2) float64, homographies (F,3,3) float64)."""
        h, w = unstabilized_frames[0].shape[:2]
        core = self._core(w, h)
        tracks = self._track(unstabilized_frames[:num_frames - 1], unstabilized_frames[1:num_frames])
        u, homs = self._device_displacements(core, tracks)
        return (u.cpu().numpy(), homs)

    def _get_adaptive_weights(self, num_frames, frame_width, frame_height, adaptive_weights_definition, homographies):
        """mfs.py:786-841 -> (F,) float64 (computed by the Jacobi coefficient kernel)."""
        core = self._core(frame_width, frame_height)
        homs_d = torch.from_numpy(np.ascontiguousarray(homographies, dtype=np.float64)).to(core.device)
        u = torch.zeros((num_frames, core.mesh.rows + 1, core.mesh.cols + 1, 2), dtype=torch.float64, device=core.device)
        _, lam = core.stabilized_displacements(u, homs_d, adaptive_weights_definition, vertex_range=(0, 0),
                                               return_lambda=True)
        return lam.cpu().numpy()

    def _get_stabilized_vertex_displacements(self, num_frames, unstabilized_frames, adaptive_weights_definition,
                                             vertex_unstabilized_displacements_by_frame_index, homographies):
        """mfs.py:632-710 -> s, same shape as u."""
        self._validate_definition(adaptive_weights_definition)
        h, w = unstabilized_frames[0].shape[:2]
        core = self._core(w, h)
        u_d = torch.from_numpy(np.ascontiguousarray(vertex_unstabilized_displacements_by_frame_index,
                                                    dtype=np.float64)).to(core.device)
        homs_d = torch.from_numpy(np.ascontiguousarray(homographies, dtype=np.float64)).to(core.device)
        return core.stabilized_displacements(u_d, homs_d, adaptive_weights_definition).cpu().numpy()

    def _get_stabilized_frames_and_crop_boundaries(self, num_frames, unstabilized_frames,
                                                   vertex_unstabilized_displacements_by_frame_index,
                                                   vertex_stabilized_displacements_by_frame_index):
        """mfs.py:909-1108 -> (list of stabilized frames, (left, top, right, bottom) np.int64)."""
        h, w = unstabilized_frames[0].shape[:2]
        core = self._core(w, h)
        frames_d = self._upload_frames(core, unstabilized_frames[:num_frames])
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(core.device)
        stab_d, crop_pf = core.warp_frames(frames_d, to(vertex_unstabilized_displacements_by_frame_index),
                                           to(vertex_stabilized_displacements_by_frame_index))
        stab = stab_d.cpu().numpy()
        return ([stab[i] for i in range(num_frames)], self._crop_tuple(core.combine_crop(crop_pf)))

    def _crop_frames(self, uncropped_frames, crop_boundaries):
        """mfs.py:1111-1157 -> list of cropped frames stretched back to the frame size."""
        h, w = uncropped_frames[0].shape[:2]
        core = self._core(w, h)
        frames_d = self._upload_frames(core, uncropped_frames)
        out = core.crop_resize(frames_d, crop_boundaries).cpu().numpy()
        return [out[i] for i in range(len(uncropped_frames))]

    def _compute_cropping_ratio_and_distortion_score(self, num_frames, unstabilized_frames, cropped_frames):
        """mfs.py:1160-1212 (host OpenCV; the per-frame matching runs on the thread pool)."""
        tracks = self._track(unstabilized_frames[:num_frames], cropped_frames[:num_frames])
        cropping_ratios, distortion_scores = self._ratios_and_scores(tracks)
        return (np.mean(cropping_ratios), np.min(distortion_scores))

    def _compute_stability_score(self, num_frames, vertex_stabilized_displacements_by_frame_index):
        """mfs.py:1216-1259 on the device (direct DFT of bins 1..5 + Parseval)."""
        s = np.ascontiguousarray(vertex_stabilized_displacements_by_frame_index, dtype=np.float64)
        core = self._core(*self._any_core_size())
        return np.float64(core.stability_score(torch.from_numpy(s).to(core.device)).item())

    def _any_core_size(self):
        if self._cores:
            return next(iter(self._cores))
        return (64, 64)   # the score does not depend on the frame size

    def _write_stabilized_video(self, output_path, num_frames, frames_per_second, codec, stabilized_frames):
        """mfs.py:1290-1322 (host code, unchanged behaviour)."""
        h, w = stabilized_frames[0].shape[:2]
        video = cv2.VideoWriter(output_path, codec, frames_per_second, (w, h))
        for i in range(num_frames):
            video.write(stabilized_frames[i])
        video.release()

    def _display_unstablilized_and_cropped_video_loop(self, num_frames, frames_per_second, unstabilized_frames, cropped_frames):
        """mfs.py:1262-1287."""
        ms = int(1000 / frames_per_second)
        while True:
            for i in range(num_frames):
                cv2.imshow('unstabilized and stabilized video', np.vstack((unstabilized_frames[i], cropped_frames[i])))
                if cv2.waitKey(ms) & 0xFF == ord('q'):
                    return
