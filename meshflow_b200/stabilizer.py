"""``MeshFlowStabilizer`` -- drop-in for how4rd/meshflow's class of the same name.

Same constructor arguments and defaults (``meshflowstabilizer.py`` lines 43-49, cited as mfs.py:N),
same four ``ADAPTIVE_WEIGHTS_DEFINITION_*`` constants (mfs.py:32-35), same
``stabilize(input_path, output_path, adaptive_weights_definition)`` entry point returning
``(cropping_ratio, distortion_score, stability_score)`` (mfs.py:102-169), and the same private stage
methods with the same signatures and return types, so stage-level tests read like tests of the
reference.  What changed is where the three data-parallel stages run:

* vertex-motion estimation  (mfs.py:236-452)  -> ``mf_vertex_motion`` + ``mf_prefix_displacements``
* Jacobi path optimisation  (mfs.py:632-878)  -> ``mf_jacobi_solve``
* mesh warp + crop          (mfs.py:909-1157) -> ``mf_warp_frames`` + ``mf_crop_resize``

all hand-written sm_100a kernels behind the C ABI in ``include/meshflow_b200.h``.  Video decode,
FAST/LK/RANSAC feature matching, the two OpenCV-derived metrics and video encode stay host OpenCV
code, as in the reference, so both implementations see identical correspondences.  There is no CPU
fallback for the three stages.
"""
from __future__ import annotations


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
                 visualize=False, *, device=None, host_workers=None, chunk_frames=16):
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
        self._cores = {}

    # ------------------------------------------------------------------------------------------
    # public entry point (mfs.py:102-169)
    # ------------------------------------------------------------------------------------------
    def stabilize(self, input_path, output_path,
                  adaptive_weights_definition=ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL):
        self._validate_definition(adaptive_weights_definition)
        frames, num_frames, fps, codec = self._get_unstabilized_frames_and_video_features(input_path)
        result = self.stabilize_frames(frames, adaptive_weights_definition)
        self._write_stabilized_video(output_path, num_frames, fps, codec, result["cropped_frames"])
        if self.visualize:
            self._display_unstablilized_and_cropped_video_loop(num_frames, fps, frames, result["cropped_frames"])
        return (result["cropping_ratio"], result["distortion_score"], result["stability_score"])

    def stabilize_frames(self, unstabilized_frames, adaptive_weights_definition=ADAPTIVE_WEIGHTS_DEFINITION_ORIGINAL,
                         with_metrics=True):
        """``stabilize()`` on in-memory frames: everything between decode and encode, device resident
        between the stages.  Returns a dict with the cropped frames, the crop rectangle, ``u``,
        ``homographies``, ``s`` (NumPy) and the three metrics."""
        self._validate_definition(adaptive_weights_definition)
        num_frames = len(unstabilized_frames)
        if num_frames < 2:
            raise ValueError("need at least two frames to stabilize")
        height, width = unstabilized_frames[0].shape[:2]
        core = self._core(width, height)
        tracks = self._track(unstabilized_frames[:-1], unstabilized_frames[1:])
        packed = self.pack_tracks(tracks)
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        h_tracks = {k: pin(packed[k]) for k in ("early", "late", "offset", "keep", "pair_start")}
        h_tracks["homographies"] = pin(packed["homographies"].reshape(-1, 9))
        h_frames = torch.empty((num_frames, height, width, 3), dtype=torch.uint8, pin_memory=True)
        view = h_frames.numpy()
        for i, f in enumerate(unstabilized_frames):
            view[i] = f
        h_out = torch.empty_like(h_frames, pin_memory=True)
        # host buffers in, host buffers out: copies overlap the kernels (pipeline.StreamedCore)
        crop_enc, u_d, s_d = StreamedCore(core, self.chunk_frames).run(h_frames, h_tracks, h_out,
                                                                       adaptive_weights_definition)
        stability = core.stability_score(s_d) if with_metrics else None
        crop = self._crop_tuple(crop_enc)                    # synchronises
        torch.cuda.synchronize(core.device)
        self._check_crop(crop, width, height)
        cropped = h_out.numpy()
        cropped_frames = [cropped[i] for i in range(num_frames)]
        homs = np.empty((num_frames, 3, 3))
        homs[:-1] = packed["homographies"]
        homs[-1] = np.identity(3)                            # mfs.py:273-274
        out = dict(cropped_frames=cropped_frames, crop_boundaries=crop, u=u_d.cpu().numpy(), homographies=homs,
                   s=s_d.cpu().numpy())
        if with_metrics:
            cr, ds = self._compute_cropping_ratio_and_distortion_score(num_frames, unstabilized_frames, cropped_frames)
            out.update(cropping_ratio=cr, distortion_score=ds, stability_score=np.float64(stability.item()))
        return out

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

    def _track(self, early_frames, late_frames):
        return host_features.track_all_pairs(
            early_frames, late_frames, self.mesh_outlier_subframe_row_count,
            self.mesh_outlier_subframe_col_count, self.homography_min_number_corresponding_features,
            workers=self.host_workers)

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
        homs = np.stack([t.homography for t in tracks]).astype(np.float64)
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

    # ------------------------------------------------------------------------------------------
    # reference-named stage methods (NumPy in / NumPy out)
    # ------------------------------------------------------------------------------------------
    def _get_unstabilized_frames_and_video_features(self, input_path):
        """mfs.py:172-213 (host code, unchanged behaviour)."""
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
        cropping_ratios = np.empty((num_frames), dtype=np.float32)
        distortion_scores = np.empty((num_frames), dtype=np.float32)
        for i, t in enumerate(tracks):
            hom = t.homography
            cropping_ratios[i] = 1 / (hom[0][0] * hom[1][1])
            affine = np.copy(hom)
            affine[2] = [0, 0, 1]
            mags = np.sort(np.abs(np.linalg.eigvals(affine)))
            distortion_scores[i] = mags[-2] / mags[-1]
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
