#!/usr/bin/env python
"""Benchmark of the MeshFlow stabilization hot path on B200 (and the CPU reference arm).

A *step* is one pass of the hot path (vertex-motion estimation -> float64 prefix sum -> Jacobi path
optimisation -> per-cell homography warp -> crop combine -> crop/resize -> stability score) over one
batch: BASELINE.json configs[1], a synthetic 1080p 300-frame jittered video, 16x16 mesh, radius 10,
100 Jacobi iterations.  The OpenCV front end (FAST / LK / RANSAC / global homography) is the
reference's own host code on both arms and runs once, outside the timed region, to produce the tracks.

  value : whole-job frames/s with frames + tracks already resident in HBM (CUDA events, max over ranks)
  e2e   : frames/s through the host-facing call: pinned HOST frames + tracks -> H2D -> hot path ->
          D2H of the cropped frames, all inside the timed region
  roofline : the warp kernel (dominant), algorithmic bytes 6*H*W per frame over its CUDA-event time,
          against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline : the oracle port (oracle/reference_port.py: same OpenCV/NumPy calls as the reference)
          timed on this box's host cores on a bounded sample (rank 0, N=1 only)

`--impl reference` times only that CPU port (the reference arm).  Under torchrun (N>1) every rank owns
`--frames` frames of an N*frames-frame video (weak scaling): velocities are all-gathered, the Jacobi
solve is vertex-sharded, the solved paths all-gathered, frames warped locally and the crop rectangle
combined with one ncclMax all-reduce.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--mesh", type=int, default=16)
    ap.add_argument("--radius", type=int, default=10)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--definition", type=int, default=0)
    ap.add_argument("--tracks", default="real", choices=["real", "synthetic"],
                    help="real = run the host OpenCV front end on the synthetic video (outside the timed region)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--chunk", type=int, default=16, help="frames per chunk of the streamed (e2e) schedule")
    ap.add_argument("--pixel-path", default="fused", choices=["fused", "two-kernel"],
                    help="fused = prepare (tables + crop) -> one warp+crop+resize kernel; two-kernel = round-1 sequence")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_workload(args, rank):
    """Frames of this rank's segment (+1 look-ahead frame for its last pair) and their tracks."""
    from tests import synth
    from meshflow_b200 import host_features
    from meshflow_b200.stabilizer import MeshFlowStabilizer
    rng = np.random.default_rng(args.seed + 7919 * rank)
    n = args.frames
    frames = synth.textured_video(rng, n + 1, args.width, args.height)
    if args.tracks == "real":
        tracks = host_features.track_all_pairs(frames[:-1], frames[1:])
        packed = MeshFlowStabilizer.pack_tracks(tracks)
    else:
        packed = synth.synthetic_tracks(rng, n, 3000, args.width, args.height)
    return frames[:n], packed


class ClockSampler(threading.Thread):
    """Polls SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port = the reference's algorithm and library calls)
# ------------------------------------------------------------------------------------------------
def cpu_hot_path_fps(args, frames, packed, warp_frames=2, vm_pairs=16):
    """Times the port on a bounded sample of the same workload and extrapolates per-frame costs
    (every stage's cost per frame / per pair is constant; Jacobi is timed in full)."""
    import cv2
    from oracle import reference_port as port
    p = port.Params(mesh_row_count=args.mesh, mesh_col_count=args.mesh,
                    temporal_smoothing_radius=args.radius, optimization_num_iterations=args.iters)
    W, H, F = args.width, args.height, len(frames)
    starts = packed["pair_start"]
    t0 = time.perf_counter()
    vels = []
    npairs = min(vm_pairs, len(starts) - 1)
    for i in range(npairs):
        a, b = starts[i], starts[i + 1]
        k = packed["keep"][a:b].astype(bool)
        off = packed["offset"][a:b][k].astype(np.float64)
        e = (packed["early"][a:b][k].astype(np.float64) + off)[:, None, :]
        l = (packed["late"][a:b][k].astype(np.float64) + off)[:, None, :]
        vels.append(port.vertex_velocities_from_matches(p, W, H, e, l, packed["homographies"][i]))
    t_vm = (time.perf_counter() - t0) / max(npairs, 1)
    rng = np.random.default_rng(5)
    u = np.cumsum(rng.normal(0, 1.0, (F, args.mesh + 1, args.mesh + 1, 2)), axis=0)
    homs = np.concatenate([packed["homographies"][:F - 1], np.eye(3)[None]])
    t0 = time.perf_counter()
    s = port.stabilized_displacements(p, W, H, args.definition, u, homs)
    t_jac = time.perf_counter() - t0
    t0 = time.perf_counter()
    stab, crop = port.warp_frames_and_crop(p, frames[:warp_frames], u[:warp_frames], s[:warp_frames])
    t_warp = (time.perf_counter() - t0) / warp_frames
    t0 = time.perf_counter()
    port.crop_frames(stab, (8, 8, W - 9, H - 9))
    t_crop = (time.perf_counter() - t0) / warp_frames
    per_frame = t_vm + t_jac / F + t_warp + t_crop
    return 1.0 / per_frame, {
        "vertex_motion_s_per_pair": t_vm, "jacobi_s_total": t_jac, "warp_s_per_frame": t_warp,
        "crop_s_per_frame": t_crop, "cores": os.cpu_count(), "cv2_threads": cv2.getNumThreads(),
        "sample": f"vertex motion on {npairs} pairs, Jacobi in full ({F} frames), warp+crop on "
                  f"{warp_frames} frames of {W}x{H}; per-frame costs summed and inverted"}


def workload_name(args):
    return (f"synthetic {args.height}p {args.frames}-frame jittered video, {args.mesh}x{args.mesh} mesh, "
            f"radius {args.radius}, {args.iters} Jacobi iterations (BASELINE.json configs[1])")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames, packed = make_workload(args, 0)
    values = []
    detail = None
    for i in range(args.warmup + args.steps):
        fps, detail = cpu_hot_path_fps(args, frames, packed, warp_frames=1, vm_pairs=4 if i < args.warmup else 16)
        if i >= args.warmup:
            values.append(fps)
    v = float(np.mean(values))
    print(json.dumps({
        "impl": "reference", "metric": "stabilized frames/sec", "value": v, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.frames / v,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/u8", "data": "synthetic",
        "config": {"workload": workload_name(args)},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": detail["cores"], "kind": "port",
                         "sample": detail["sample"]},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "detail": detail}))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from meshflow_b200 import DeviceCore, MeshSpec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # keep this rank's pinned host buffers (and the threads that fill them) on the GPU's own NUMA node:
    # with several ranks per box the host side of the copies is otherwise the bottleneck
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    W, H, F = args.width, args.height, args.frames
    mesh = MeshSpec(W, H, args.mesh, args.mesh)
    core = DeviceCore(mesh, device=dev, radius=args.radius, iterations=args.iters)
    V = mesh.vertices

    frames, packed = make_workload(args, rank)
    # host buffers (pinned) for the e2e arm
    h_frames = torch.empty((F, H, W, 3), dtype=torch.uint8, pin_memory=True)
    for i, f in enumerate(frames):
        h_frames[i] = torch.from_numpy(f)
    h_out = torch.empty((F, H, W, 3), dtype=torch.uint8, pin_memory=True)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    homs_pairs = packed["homographies"].reshape(-1, 9)
    h_tracks = {k: pin(packed[k]) for k in ("early", "late", "offset", "keep", "pair_start")}
    h_tracks["homographies"] = pin(homs_pairs)
    P = len(packed["pair_start"]) - 1          # = F pairs per rank (the look-ahead frame closes the last one)
    last_rank = rank == world - 1
    F_total = F * world

    d_frames = h_frames.to(dev)
    d_tracks = {k: v.to(dev) for k, v in h_tracks.items()}
    d_out = torch.empty_like(d_frames)
    d_stab = torch.empty_like(d_frames)
    ident = torch.eye(3, dtype=torch.float64, device=dev).reshape(1, 9)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    fused = args.pixel_path == "fused"
    stage_names = (["paths (vertex motion + prefix + Jacobi + exchanges)", "prepare (cells, spans, segments, crop edges)",
                    "warp", "stability"] if fused else
                   ["paths (vertex motion + prefix + Jacobi + exchanges)", "warp", "crop_resize", "stability"])

    from meshflow_b200 import StreamedCore, distributed as mfd
    host_pair_start = packed["pair_start"]
    plan = mfd.ShardPlan(rank, world, [F] * world) if world > 1 else None

    def hot_path(tr, frames_d, out_d, marks=None):
        def mark():
            if marks is not None:
                e = ev(); e.record(); marks.append(e)
        mark()
        u, s, _ = mfd.sharded_paths(core, tr, F, args.definition, pair_start_host=host_pair_start, plan=plan)
        mark()
        lo = rank * F
        if fused:
            crop_pf, tables = core.warp_prepare(u[lo:lo + F], s[lo:lo + F])
            mark()
            enc = mfd.reduce_crop(core.combine_crop(crop_pf), plan)
            core.warp_resize_frames(frames_d, enc, tables, 0, out=out_d)
            mark()
        else:
            _, crop_pf = core.warp_frames(frames_d, u[lo:lo + F], s[lo:lo + F], out=d_stab)
            mark()
            enc = mfd.reduce_crop(core.combine_crop(crop_pf), plan)
            core.crop_resize_device(d_stab, enc, out=out_d)
            mark()
        score = core.stability_score(s)
        mark()
        return enc, score

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing --------------------------------------------------------------------
    for _ in range(args.warmup):
        hot_path(d_tracks, d_frames, d_out)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    t_begin, t_end = ev(), ev()
    marks_all = []
    barrier()
    t_begin.record()
    for _ in range(args.steps):
        marks = []
        enc, score = hot_path(d_tracks, d_frames, d_out, marks)
        marks_all.append(marks)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    ms_total = t_begin.elapsed_time(t_end)
    stage_ms = {n: 0.0 for n in stage_names}
    for marks in marks_all:
        for i, n in enumerate(stage_names):
            stage_ms[n] += marks[i].elapsed_time(marks[i + 1]) / args.steps
    crop = core.decode_crop(enc)

    # ---- end-to-end timing: host buffers in, host buffers out ------------------------------------------
    streamed = StreamedCore(core, chunk_frames=args.chunk)

    def e2e_step():
        # pinned host frames + tracks in, pinned host frames out; copies overlap the kernels
        enc, u, s = streamed.run(h_frames, h_tracks, h_out, args.definition, plan=plan)
        score = core.stability_score(s)
        return enc, score.item()           # device -> host read of the step's result (synchronises)

    for _ in range(max(1, args.warmup - 1)):
        e2e_step()
    barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames_per_step = F_total
    value = frames_per_step * args.steps / (ms_total / 1e3)
    e2e_value = frames_per_step * args.steps / (ms_e2e / 1e3)
    h2d = int(h_frames.numel() + sum(v.numel() * v.element_size() for v in h_tracks.values()))
    d2h = int(h_out.numel() + 8)

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    warp_bytes = 6.0 * H * W * F                       # read source once + write stabilized once, per launch
    achieved = warp_bytes / (stage_ms["warp"] / 1e3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "warp_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_per_1080p_frame"] * F * (W * H) / (1920.0 * 1080.0)
        traffic_src = tj["source"]
    resize_gbs = warp_bytes / (stage_ms["crop_resize"] / 1e3) / 1e9 if "crop_resize" in stage_ms else None
    out = {
        "metric": "stabilized frames/sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/u8",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "frames_per_gpu": F, "tracks": args.tracks,
                   "features_per_pair": int(np.diff(packed["pair_start"]).mean()),
                   "l2": "inputs larger than L2 (1.87 GB of frames per step per GPU)",
                   "crop": list(crop), "parallelism": f"frames x{world}, Jacobi vertices x{world}"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        # kernels of this library per step: feature_prepare_masks, pair_sort, row_select, median3x3, prefix,
        # jacobi_coeff, jacobi_solve, cell_setup, tile_sort, cell_spans, row_segments, warp_fast, crop_combine,
        # resize_table, crop_resize_rows, stability
        "gpu_launches": 16 * args.steps,
        "stages_ms": stage_ms,
        "roofline": {"kernel": "warp_fast_kernel (timed: the whole mf_warp_frames stage = cell_setup + tile_sort + "
                               "row_segments + warp_fast)", "bound": "hbm",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read + dram__bytes_write of the stage's kernels from one ncu --set full capture
                     # (profiles/warp_traffic.json names it), per 1080p frame, scaled to this launch
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": warp_bytes},
        "crop_resize_gbs": resize_gbs,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        fps, detail = cpu_hot_path_fps(args, frames, packed)
        out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": detail["cores"], "kind": "port",
                               "sample": detail["sample"], "detail": detail}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
