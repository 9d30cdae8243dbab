/*
 * meshflow_b200 -- C ABI of the B200-native MeshFlow stabilization core.
 *
 * The reference (how4rd/meshflow, meshflowstabilizer.py, cited below as mfs.py:N) has no FFI of its
 * own: its boundary is the Python class MeshFlowStabilizer.  Each entry point below replaces the body
 * of one of the private stage methods that stabilize() (mfs.py:102-169) calls; the Python host side
 * (meshflow_b200/stabilizer.py) keeps the reference's method names and signatures and binds these
 * symbols with ctypes (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a contiguous array owned by the caller;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing and
 *     never synchronises; scratch memory is passed in as `workspace` (size from the *_workspace_bytes
 *     query of the same stage);
 *   - return value: 0 on success, a negative MF_E_* code otherwise; mf_last_error() then returns a
 *     thread-local, human-readable message;
 *   - V = (R+1)*(C+1) mesh vertices in row-major order (row outer, column inner), exactly the order
 *     of _get_vertex_x_y (mfs.py:881-906);
 *   - "definition" is the reference's ADAPTIVE_WEIGHTS_DEFINITION_* value 0..3 (mfs.py:32-35).
 */
#ifndef MESHFLOW_B200_H
#define MESHFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Only the entry points below are exported: the library is built with -fvisibility=hidden. */
#if defined(__GNUC__)
#define MF_API __attribute__((visibility("default")))
#else
#define MF_API
#endif

#define MF_OK 0
#define MF_E_INVALID (-1)   /* bad argument (null pointer, non-positive size, unknown definition) */
#define MF_E_WORKSPACE (-2) /* workspace too small */
#define MF_E_LAUNCH (-3)    /* CUDA launch / runtime error (message holds cudaGetErrorString) */
#define MF_E_UNSUPPORTED (-4) /* size outside what the kernels are built for */

/* Library / build identification.  mf_version() = 10000*major + 100*minor + patch. */
MF_API int mf_version(void);
/* Compute capability the embedded cubin was built for (100 for sm_100a). */
MF_API int mf_built_for_sm(void);
MF_API const char* mf_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * (1) Vertex-motion estimation -- replaces _get_unstabilized_vertex_velocities (mfs.py:287-362) minus
 *     the host OpenCV matching, _get_vertex_nearby_feature_residual_velocities (mfs.py:365-452) and
 *     the prefix sum of _get_unstabilized_vertex_displacements_and_homographies (mfs.py:268-284),
 *     batched over P frame pairs.
 *
 * Features arrive UN-compacted, as the host tracker produced them; `keep` applies the LK status mask
 * (mfs.py:622-624) and the per-subframe RANSAC inlier mask (mfs.py:569-574) on the device.
 *   early_xy, late_xy : [N,2] float32, subframe-relative coordinates of all tracked candidates
 *   offset_xy         : [N,2] int32, top-left corner of the feature's subframe (mfs.py:509, 578)
 *   keep              : [N] uint8, 1 = survives both masks
 *   pair_start        : [P+1] int32, features of pair p are [pair_start[p], pair_start[p+1])
 *   pair_start_host   : the same [P+1] array in HOST memory, or NULL.  With it the library knows the
 *                       largest pair and takes the fast path (one shared-memory sort per pair and
 *                       component + bit-matrix rank selection per mesh row; needs <= 16384 features
 *                       per pair and <= 96 mesh columns); without it, or beyond those sizes, the
 *                       generic per-vertex radix-select path runs.  Both paths give identical results.
 *   homographies      : [P,9] float64, global early->late homography of each pair (mfs.py:524)
 *   vertex_xy         : [V,2] float32 rest positions (mfs.py:881-906)
 *   vel_out           : [P,V,2] float32 vertex velocities after both median filters
 *   assign_count_out  : optional [P,V] int32, number of features assigned to each vertex (parity of
 *                       the feature->vertex assignment); may be NULL
 * ---------------------------------------------------------------------------------------------- */
MF_API size_t mf_vertex_motion_workspace_bytes(int64_t N, int P, int R, int C);
MF_API int mf_vertex_motion(const float* early_xy, const float* late_xy, const int32_t* offset_xy,
                     const uint8_t* keep, const int32_t* pair_start, const int32_t* pair_start_host,
                     int64_t N, int P,
                     const double* homographies, const float* vertex_xy,
                     int W, int H, int R, int C, int ellipse_rows, int ellipse_cols,
                     float* vel_out, int32_t* assign_count_out,
                     void* workspace, size_t workspace_bytes, void* stream);

/* disp[0] = 0, disp[t+1] = disp[t] + (double)vel[t]  -- sequential float64 scan (mfs.py:271, 281).
 *   vel: [P, n] float32, disp: [P+1, n] float64, n = 2V.  An optional `disp0` [n] float64 seeds
 *   disp[0] (frame-sharded callers chain shards with it); NULL means zeros. */
MF_API int mf_prefix_displacements(const float* vel, const double* disp0, double* disp, int P, int64_t n,
                            void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Jacobi path optimisation -- replaces _get_stabilized_vertex_displacements (mfs.py:632-710),
 *     _get_jacobi_method_input (713-783), _get_adaptive_weights (786-841) and
 *     _get_jacobi_method_output (844-878).  Band-limited restatement of the dense system; the
 *     adaptive weight lambda_t is computed on the device from `homographies`.
 *   u, s          : [F, n_sys] float64, n_sys = 2V systems (vertex-major, x/y interleaved); systems
 *                   [sys_begin, sys_end) are solved (vertex sharding), the rest of `s` is untouched
 *   homographies  : [F,9] float64 (last one identity, mfs.py:273-274)
 *   lambda_out    : optional [F] float64 copy of the adaptive weights; may be NULL
 * ---------------------------------------------------------------------------------------------- */
MF_API size_t mf_jacobi_workspace_bytes(int F, int64_t n_sys);
MF_API int mf_jacobi_solve(const double* u, const double* homographies, double* s, int F, int64_t n_sys,
                    int64_t sys_begin, int64_t sys_end, int W, int H, int radius, int iterations,
                    int definition, double* lambda_out, void* workspace, size_t workspace_bytes,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Mesh warp + crop -- replaces _get_stabilized_frames_and_crop_boundaries (mfs.py:909-1108).
 *   frames_in   : [nf, H, W, 3] uint8 BGR
 *   u, s        : [nf, V, 2] float64 displacements OF THESE nf FRAMES (caller passes its shard)
 *   vertex_xy   : [V,2] float32
 *   frames_out  : [nf, H, W, 3] uint8 stabilized frames (cv2.remap fixed-point bilinear, constant
 *                 border b,g,r)
 *   crop_out    : [nf,4] int32 per-frame (left, top, right, bottom) (mfs.py:1075-1098); the caller
 *                 combines them with max/max/min/min (mfs.py:1103-1106; an all-reduce when sharded)
 *   map_out     : optional [nf, H, W, 2] float32 (map_x, map_y) for parity checks; may be NULL.  The float32
 *                 maps only exist in the generic kernel, so asking for them selects it; without map_out the
 *                 production path runs (row segments + float32 coordinates outside a proven rounding band,
 *                 csrc/warp_fast.cuh) -- same frames_out and crop_out, bit for bit.  Environment switches
 *                 for A/B runs: MF_WARP_GENERIC=1, MF_RESIZE_GENERIC=1.
 * ---------------------------------------------------------------------------------------------- */
MF_API size_t mf_warp_workspace_bytes(int nf, int W, int H, int R, int C);
MF_API int mf_warp_frames(const uint8_t* frames_in, const double* u, const double* s, const float* vertex_xy,
                   int nf, int W, int H, int R, int C, int border_b, int border_g, int border_r,
                   uint8_t* frames_out, int32_t* crop_out, float* map_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Pass A of the streamed schedule: everything about the warp that depends on the vertex paths only -- the cell
 * homographies, the exact member interval of every cell on every output row, the row segments ("the last
 * cell written wins", mfs.py:1060-1061) and, from the segments of the border cells, the per-frame crop edges
 * (mfs.py:1075-1098) by a closed-form band search: no pixel is read.  With the crop rectangle of the whole
 * video known up front, pass B (mf_warp_resize_frames) runs chunk by chunk while frames are still being
 * uploaded and results downloaded.  Same crop_out as mf_warp_frames, same workspace size; the tables stay in
 * `workspace` for mf_warp_resize_frames.  mf_warp_crop_bounds is the round-1 name of the same call. */
MF_API int mf_warp_prepare(const double* u, const double* s, const float* vertex_xy, int nf, int W, int H,
                           int R, int C, int32_t* crop_out, void* workspace, size_t workspace_bytes,
                           void* stream);
MF_API int mf_warp_crop_bounds(const double* u, const double* s, const float* vertex_xy, int nf, int W, int H,
                        int R, int C, int32_t* crop_out, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Pass B, fused: _get_stabilized_frames_and_crop_boundaries' remap (mfs.py:1063-1069) followed by
 * _crop_frames (mfs.py:1111-1157) in ONE kernel.  A thread block produces the stabilized pixels one tile of
 * the final frame needs into shared memory and resizes from there: the stabilized frame never exists in
 * DRAM and pixels outside the crop rectangle are never computed.  Output identical to
 * mf_warp_frames + mf_crop_resize_device.
 *   frames_in     : [nf, H, W, 3] uint8, frames first_frame .. first_frame + nf - 1 of the prepared video
 *   workspace     : the workspace mf_warp_prepare filled for `table_frames` frames (first_frame + nf <= table_frames)
 *   crop_enc      : device int32[4] = [left, top, -right, -bottom] (mf_crop_combine, all-reduced when sharded)
 *   resize_workspace : mf_crop_resize_workspace_bytes(W, H) bytes
 * Returns MF_E_UNSUPPORTED for frames too small for the row-segment tables (W < 16): use the two calls. */
MF_API int mf_warp_resize_frames(const uint8_t* frames_in, int nf, int first_frame, int table_frames, int W, int H,
                                 int R, int C, int border_b, int border_g, int border_r, const int32_t* crop_enc,
                                 uint8_t* frames_out, void* workspace, size_t workspace_bytes,
                                 void* resize_workspace, size_t resize_workspace_bytes, void* stream);

/* Crop rectangle (inclusive) stretched back to W x H -- replaces _crop_frames (mfs.py:1111-1157);
 * cv2.resize INTER_LINEAR 11-bit fixed point. */
MF_API size_t mf_crop_resize_workspace_bytes(int W, int H);
MF_API int mf_crop_resize(const uint8_t* frames_in, int nf, int W, int H, int left, int top, int right,
                   int bottom, uint8_t* frames_out, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Device-side crop plumbing, so that a step never waits for the host (and so that the multi-GPU
 * combine is one ncclMax all-reduce on the same 4 ints):
 *   mf_crop_combine       : per-frame edges [nf,4] -> crop_enc_out[4] = [max left, max top, -min right,
 *                           -min bottom]  (mfs.py:1103-1106 as a single max-reduction)
 *   mf_crop_resize_device : mf_crop_resize with the rectangle read from device memory in that
 *                           encoding; an empty / out-of-frame rectangle leaves frames_out untouched
 *                           (the host raises when it reads the rectangle back). */
MF_API int mf_crop_combine(const int32_t* per_frame_crop, int nf, int32_t* crop_enc_out, void* stream);
MF_API int mf_crop_resize_device(const uint8_t* frames_in, int nf, int W, int H, const int32_t* crop_enc,
                          uint8_t* frames_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stability score -- replaces _compute_stability_score (mfs.py:1216-1259): per system the energy of
 * DFT bins 1..5 of the frame-to-frame differences over the total energy (Parseval).
 *   s : [F, n_sys] float64;  ratio_out : [n_sys] float64 (caller averages x and y systems).
 * ---------------------------------------------------------------------------------------------- */
MF_API int mf_stability_ratios(const double* s, int F, int64_t n_sys, double* ratio_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Diagnostics (used by the test-suite; not needed by a caller of the path).
 *   mf_debug_rcp_mismatches : runs the warp kernels' correctly rounded float64 reciprocal on n seeded inputs
 *       and returns how many differ from IEEE 1/w (expected 0); synchronises `stream`.
 *   mf_debug_force_generic_vertex_motion : 1 = route mf_vertex_motion through the generic per-vertex path even
 *       when the fast path applies (A/B comparisons), 0 = automatic.
 * ---------------------------------------------------------------------------------------------- */
MF_API long long mf_debug_rcp_mismatches(long long n, unsigned long long seed, void* scratch_u64, void* stream);
MF_API void mf_debug_force_generic_vertex_motion(int on);

#ifdef __cplusplus
}
#endif
#endif /* MESHFLOW_B200_H */
