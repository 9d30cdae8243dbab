// Jacobi path optimisation on the device (hot-path subsystem 2).
//
// The reference builds two dense F x F matrices and runs K sweeps of two dense mat-muls per vertex
// (mfs.py:713-783, 844-878).  The system is banded: element form
//     x'[t] = (1/diag[t]) * ( b[t] + 2*lambda[t] * sum_{|k| <= radius, 0 <= t+k < F} w[k] * x[t+k] )
// with w[k] = exp(-((3/radius) k)^2) INCLUDING k = 0, diag[t] = 1 + 2 lambda[t] sum_{r=0}^{F-1} w[t-r]
// over ALL frames (both quirks of the reference are kept).
//
//   jacobi_coeff_kernel : lambda[t] from the global homography (mfs.py:786-841), 1/diag[t].
//   jacobi_solve_kernel : persistent -- one CTA owns the x and y trajectories of one vertex
//                         (double2 per frame), keeps them in shared memory with `radius` zero cells
//                         of padding on either side (so the boundary needs no branches), and runs
//                         every sweep without leaving the SM: HBM is touched twice (load b, store x).
#include "mf_common.cuh"
#include "mf_math.cuh"
#include <cmath>

namespace mf {

__global__ void __launch_bounds__(128) jacobi_coeff_kernel(const double* __restrict__ homographies, int F,
                                                           int W, int H, int radius, int definition, int table_entries,
                                                           double* __restrict__ inv_diag,
                                                           double* __restrict__ two_lambda,
                                                           double* __restrict__ lambda_out) {
  // exp(-(3k/radius)^2) underflows to exactly 0 beyond |k| ~ 9.1 radius: summing 10 radius either
  // side equals the reference's sum over all frames.  The CTA evaluates the reach + 1 distinct weights once
  // (table_entries > 0: they fit the shared-memory table); a frame then adds them in frame order.
  extern __shared__ double wtab[];
  const double c = 3.0 / (double)radius;
  const int reach = 10 * radius + 1;
  for (int k = threadIdx.x; k < table_entries; k += blockDim.x) {
    const double a = c * (double)k;
    wtab[k] = exp(-(a * a));
  }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= F) return;
  const double lam = adaptive_lambda(homographies + (size_t)t * 9, W, H, definition);
  const int lo = max(0, t - reach), hi = min(F - 1, t + reach);
  double sum = 0.0;
  if (table_entries > 0) {
    for (int r = lo; r <= hi; ++r) sum += wtab[abs(t - r)];
  } else {
    for (int r = lo; r <= hi; ++r) {
      const double a = c * (double)(t - r);
      sum += exp(-(a * a));
    }
  }
  const double diag = 1.0 + 2.0 * (lam * sum);
  inv_diag[t] = 1.0 / diag;
  two_lambda[t] = 2.0 * lam;
  if (lambda_out) lambda_out[t] = lam;
}

// One CTA per vertex.  PT = frames per thread (strided by blockDim so that shared-memory reads of
// neighbouring threads are neighbouring double2 cells).  RADIUS > 0: compile-time radius, weights in
// registers; RADIUS == 0: run-time radius, weights in shared memory.
template <int RADIUS, int PT>
__global__ void __launch_bounds__(512) jacobi_solve_kernel(
    const double* __restrict__ u, double* __restrict__ s, int F, int64_t n_sys, int64_t sys_begin,
    int radius_rt, int iterations, const double* __restrict__ inv_diag,
    const double* __restrict__ two_lambda) {
  extern __shared__ double2 xs_raw[];  // [F + 2 radius] trajectory, then (generic) [2 radius + 1] weights
  const int radius = RADIUS > 0 ? RADIUS : radius_rt;
  double2* xs = xs_raw + radius;  // xs[t], t in [-radius, F + radius)
  double* wsm = reinterpret_cast<double*>(xs_raw + F + 2 * radius);
  const int64_t q = sys_begin + 2 * (int64_t)blockIdx.x;  // first of the two systems (x, y)
  const int nt = blockDim.x, tid = threadIdx.x;

  double wreg[RADIUS > 0 ? 2 * RADIUS + 1 : 1];
  const double c = 3.0 / (double)radius;
  if (RADIUS > 0) {
#pragma unroll
    for (int k = 0; k < 2 * RADIUS + 1; ++k) {
      const double a = c * (double)(k - RADIUS);
      wreg[k] = exp(-(a * a));
    }
  } else {
    for (int k = tid; k < 2 * radius + 1; k += nt) {
      const double a = c * (double)(k - radius);
      wsm[k] = exp(-(a * a));
    }
  }
  for (int k = tid; k < radius; k += nt) {
    xs[-1 - k] = make_double2(0.0, 0.0);
    xs[F + k] = make_double2(0.0, 0.0);
  }
  double2 b[PT];
  double invd[PT], twol[PT];
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    const int t = tid + j * nt;
    if (t < F) {
      b[j] = *reinterpret_cast<const double2*>(u + (size_t)t * n_sys + q);
      invd[j] = inv_diag[t];
      twol[j] = two_lambda[t];
      xs[t] = b[j];
    } else {
      b[j] = make_double2(0.0, 0.0); invd[j] = 0.0; twol[j] = 0.0;
    }
  }
  __syncthreads();
  for (int it = 0; it < iterations; ++it) {
    double2 nv[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const int t = tid + j * nt;
      double ax = 0.0, ay = 0.0;
      if (t < F) {
        const double2* win = xs + t - radius;
        if (RADIUS > 0) {
#pragma unroll
          for (int k = 0; k < 2 * RADIUS + 1; ++k) {
            const double2 v = win[k];
            ax = fma(wreg[k], v.x, ax);
            ay = fma(wreg[k], v.y, ay);
          }
        } else {
          for (int k = 0; k < 2 * radius + 1; ++k) {
            const double2 v = win[k];
            const double w = wsm[k];
            ax = fma(w, v.x, ax);
            ay = fma(w, v.y, ay);
          }
        }
      }
      nv[j].x = invd[j] * fma(twol[j], ax, b[j].x);
      nv[j].y = invd[j] * fma(twol[j], ay, b[j].y);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const int t = tid + j * nt;
      if (t < F) xs[t] = nv[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < PT; ++j) {
    const int t = tid + j * nt;
    if (t < F) *reinterpret_cast<double2*>(s + (size_t)t * n_sys + q) = xs[t];
  }
}

// Large-F variant: the per-frame constants do not fit in registers next to the staged sweep, so
// b / inv_diag / two_lambda are re-read (L1/L2 resident) every sweep.
template <int RADIUS>
__global__ void __launch_bounds__(512) jacobi_solve_long_kernel(
    const double* __restrict__ u, double* __restrict__ s, int F, int64_t n_sys, int64_t sys_begin,
    int radius_rt, int iterations, const double* __restrict__ inv_diag,
    const double* __restrict__ two_lambda) {
  constexpr int PT = 20;  // 512 threads x 20 frames = 10240 frames
  extern __shared__ double2 xs_raw[];
  const int radius = RADIUS > 0 ? RADIUS : radius_rt;
  double2* xs = xs_raw + radius;
  double* wsm = reinterpret_cast<double*>(xs_raw + F + 2 * radius);
  const int64_t q = sys_begin + 2 * (int64_t)blockIdx.x;
  const int nt = blockDim.x, tid = threadIdx.x;
  const double c = 3.0 / (double)radius;
  for (int k = tid; k < 2 * radius + 1; k += nt) {
    const double a = c * (double)(k - radius);
    wsm[k] = exp(-(a * a));
  }
  for (int k = tid; k < radius; k += nt) {
    xs[-1 - k] = make_double2(0.0, 0.0);
    xs[F + k] = make_double2(0.0, 0.0);
  }
  for (int t = tid; t < F; t += nt) xs[t] = *reinterpret_cast<const double2*>(u + (size_t)t * n_sys + q);
  __syncthreads();
  for (int it = 0; it < iterations; ++it) {
    double2 nv[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const int t = tid + j * nt;
      if (t < F) {
        const double2* win = xs + t - radius;
        double ax = 0.0, ay = 0.0;
#pragma unroll 4
        for (int k = 0; k < 2 * radius + 1; ++k) {
          const double2 v = win[k];
          const double w = wsm[k];
          ax = fma(w, v.x, ax);
          ay = fma(w, v.y, ay);
        }
        const double2 bb = __ldg(reinterpret_cast<const double2*>(u + (size_t)t * n_sys + q));
        const double id = __ldg(inv_diag + t), tl = __ldg(two_lambda + t);
        nv[j].x = id * fma(tl, ax, bb.x);
        nv[j].y = id * fma(tl, ay, bb.y);
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PT; ++j) {
      const int t = tid + j * nt;
      if (t < F) xs[t] = nv[j];
    }
    __syncthreads();
  }
  for (int t = tid; t < F; t += nt) *reinterpret_cast<double2*>(s + (size_t)t * n_sys + q) = xs[t];
}

// Register-window variant for long videos (F up to T*512 frames).  One CTA per SYSTEM (vertex
// component).  Thread t owns the T consecutive frames [T t, T t + T): per sweep it streams the
// T + 2 RADIUS inputs it needs through registers once and accumulates every (input, output) pair with
// compile-time offsets -- T (2 RADIUS + 1) FMAs for T + 2 RADIUS shared-memory reads, instead of one
// read per FMA.  The trajectory lives in shared memory in a [T][threads] layout (frame T t + i at
// [i][t]) so that every access of a warp is to consecutive words; the right-hand side is kept there
// too as B = b / diag, the per-frame gain g = 2 lambda / diag in registers, the weights in the
// kernel-parameter constant bank.  x' = B + g * sum_k w_k x_{t+k}  ==  (b + 2 lambda sum) / diag.
struct JacobiWeights { double w[32]; };   // w[|k|], k <= 31

template <int T, int RADIUS>
__global__ void __launch_bounds__(512) jacobi_window_kernel(
    const double* __restrict__ u, double* __restrict__ s, int F, int64_t n_sys, int64_t sys_begin, int iterations,
    const double* __restrict__ inv_diag, const double* __restrict__ two_lambda, const JacobiWeights wts) {
  constexpr int HP = (RADIUS + T - 1) / T;          // pad columns either side (frames of absent neighbours)
  extern __shared__ double sm[];
  const int nt = blockDim.x, t = threadIdx.x;
  const int ntp = nt + 2 * HP;
  double* xs = sm;                                   // [T][ntp]
  double* bs = sm + T * ntp;                         // [T][nt]
  const int64_t q = sys_begin + blockIdx.x;
  for (int i = t; i < T * ntp; i += nt) xs[i] = 0.0;
  __syncthreads();
  double g[T];
#pragma unroll
  for (int i = 0; i < T; ++i) {
    const int f = T * t + i;
    double b = 0.0, id = 0.0, tl = 0.0;
    if (f < F) { b = u[(size_t)f * n_sys + q]; id = inv_diag[f]; tl = two_lambda[f]; }
    xs[i * ntp + t + HP] = b;
    bs[i * nt + t] = id * b;
    g[i] = id * tl;
  }
  __syncthreads();
  for (int it = 0; it < iterations; ++it) {
    double acc[T];
#pragma unroll
    for (int i = 0; i < T; ++i) acc[i] = 0.0;
#pragma unroll
    for (int j = 0; j < T + 2 * RADIUS; ++j) {
      constexpr int kBias = 64 * T;                  // keeps the division below on non-negative numbers
      const int jj = j - RADIUS;                      // input frame relative to this thread's first frame
      const int dt = (jj + kBias) / T - 64, ii = (jj + kBias) % T;
      const double v = xs[ii * ntp + t + HP + dt];
#pragma unroll
      for (int i = 0; i < T; ++i) {
        const int d = jj - i;
        if (d >= -RADIUS && d <= RADIUS) acc[i] = fma(wts.w[d < 0 ? -d : d], v, acc[i]);
      }
    }
    double nx[T];
#pragma unroll
    for (int i = 0; i < T; ++i) nx[i] = fma(g[i], acc[i], bs[i * nt + t]);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < T; ++i)
      if (T * t + i < F) xs[i * ntp + t + HP] = nx[i];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < T; ++i) {
    const int f = T * t + i;
    if (f < F) s[(size_t)f * n_sys + q] = xs[i * ntp + t + HP];
  }
}

// Fallback without a frame limit (videos longer than the on-chip kernels hold: F > 10 240, or a radius whose
// padded trajectory does not fit shared memory): one launch per sweep, trajectories ping-pong between two global
// buffers ([F][n] layout, adjacent threads = adjacent systems, so every access is coalesced and the 2 radius + 1
// neighbours of a frame come from L2).  Same element formula and summation order as the on-chip kernels.
__global__ void __launch_bounds__(256) jacobi_weights_kernel(int radius, double* __restrict__ w) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > radius) return;
  const double a = (3.0 / (double)radius) * (double)k;
  w[k] = exp(-(a * a));
}

__global__ void __launch_bounds__(256) jacobi_sweep_global_kernel(
    const double* __restrict__ x_in, int64_t in_stride, int64_t in_col0, const double* __restrict__ b, int64_t n_sys,
    int64_t sys_begin, double* __restrict__ x_out, int64_t out_stride, int64_t out_col0, int F, int64_t n_cols, int radius,
    const double* __restrict__ inv_diag, const double* __restrict__ two_lambda, const double* __restrict__ w) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y;
  if (c >= n_cols) return;
  double acc = 0.0;
  const int lo = max(0, t - radius), hi = min(F - 1, t + radius);
  for (int r = lo; r <= hi; ++r) {
    const int k = r - t;
    acc = fma(__ldg(w + (k < 0 ? -k : k)), x_in[(int64_t)r * in_stride + in_col0 + c], acc);
  }
  x_out[(int64_t)t * out_stride + out_col0 + c] = inv_diag[t] * fma(two_lambda[t], acc, b[(int64_t)t * n_sys + sys_begin + c]);
}

__global__ void __launch_bounds__(256) jacobi_copy_kernel(const double* __restrict__ b, int64_t n_sys, int64_t sys_begin,
                                                          double* __restrict__ out, int F, int64_t n_cols) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_cols) out[(int64_t)blockIdx.y * n_sys + sys_begin + c] = b[(int64_t)blockIdx.y * n_sys + sys_begin + c];
}

static constexpr int kGlobalMaxRadius = 1023;
static size_t global_scratch_bytes(int F, int64_t n_sys) {
  return align_up((size_t)(kGlobalMaxRadius + 1) * sizeof(double), 256) + align_up((size_t)F * (size_t)n_sys * sizeof(double), 256);
}

// Frames the on-chip kernels hold per solve (jacobi_window_kernel: 20 frames x 512 threads).
static constexpr int kOnChipMaxFrames = 20 * 512;
static bool needs_global_path(int F, int radius) {
  const size_t smem = (size_t)(F + 2 * radius) * sizeof(double2) + (size_t)(2 * radius + 1) * sizeof(double);
  return F > kOnChipMaxFrames || smem > 227 * 1024;
}

static int launch_global(const double* u, double* s, int F, int64_t n_sys, int64_t sys_begin, int64_t n_cols, int radius,
                         int iterations, const double* inv_diag, const double* two_lambda, void* scratch, cudaStream_t st) {
  if (radius > kGlobalMaxRadius) return fail(MF_E_UNSUPPORTED, "jacobi: radius %d exceeds %d", radius, kGlobalMaxRadius);
  if (F > 65535) return fail(MF_E_UNSUPPORTED, "jacobi: F=%d exceeds 65535 frames per solve", F);
  double* w = (double*)scratch;
  double* tmp = (double*)((char*)scratch + align_up((size_t)(kGlobalMaxRadius + 1) * sizeof(double), 256));
  jacobi_weights_kernel<<<(radius + 256) / 256, 256, 0, st>>>(radius, w);
  const dim3 grid((unsigned)((n_cols + 255) / 256), (unsigned)F);
  if (iterations == 0) {
    jacobi_copy_kernel<<<grid, 256, 0, st>>>(u, n_sys, sys_begin, s, F, n_cols);
    return check_launch("jacobi_copy");
  }
  // sweep i (1-based) writes `s` when (iterations - i) is even, so that the last one lands in `s`
  const double* in = u; int64_t in_stride = n_sys, in_col0 = sys_begin;
  for (int i = 1; i <= iterations; ++i) {
    const bool to_s = ((iterations - i) & 1) == 0;
    double* out = to_s ? s : tmp;
    const int64_t out_stride = to_s ? n_sys : n_cols, out_col0 = to_s ? sys_begin : 0;
    jacobi_sweep_global_kernel<<<grid, 256, 0, st>>>(in, in_stride, in_col0, u, n_sys, sys_begin, out, out_stride, out_col0, F,
                                                     n_cols, radius, inv_diag, two_lambda, w);
    in = out; in_stride = out_stride; in_col0 = out_col0;
  }
  return check_launch("jacobi_sweep_global");
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(MF_E_LAUNCH, "jacobi: shared memory opt-in (%zu B): %s", bytes,
                                      cudaGetErrorString(e));
  }
  return MF_OK;
}

template <int T, int RADIUS>
static int launch_window(const double* u, double* s, int F, int64_t n_sys, int64_t sys_begin, int64_t n_vert, int iterations,
                         const double* inv_diag, const double* two_lambda, cudaStream_t st) {
  const int nt = ((F + T - 1) / T + 31) / 32 * 32;
  constexpr int hp = (RADIUS + T - 1) / T;
  const size_t wsmem = ((size_t)T * (nt + 2 * hp) + (size_t)T * nt) * sizeof(double);
  JacobiWeights wts;
  for (int k = 0; k < 32; ++k) {
    const double a = (3.0 / (double)RADIUS) * (double)k;
    wts.w[k] = k <= RADIUS ? exp(-(a * a)) : 0.0;
  }
  auto k = jacobi_window_kernel<T, RADIUS>;
  if (int e = set_smem(k, wsmem)) return e;
  k<<<dim3((unsigned)(2 * n_vert)), nt, wsmem, st>>>(u, s, F, n_sys, sys_begin, iterations, inv_diag, two_lambda, wts);
  return check_launch("jacobi_window");
}

// The two radii the reference's users run (10: constructor default, 30: BASELINE config 4) take the register-window
// kernel for every video length; T frames per thread:
//   20  reads the fewest shared-memory words per FMA: best when the systems outnumber the SMs and the video is long (c4);
//   5   spreads a system over the most threads: short videos (c2: 300 frames -> 64 threads per system) and the vertex
//       shard of a multi-GPU run, which has few systems (74 at N = 8 on c2) but world x F frames.
// Every T performs the same operations in the same order, so the choice never changes a result bit.
template <int RADIUS>
static int launch_windowed(const double* u, double* s, int F, int64_t n_sys, int64_t sys_begin, int64_t n_vert, int iterations,
                           const double* inv_diag, const double* two_lambda, cudaStream_t st) {
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool few = 2 * n_vert < 2 * (int64_t)sms;
  int T = F >= 2560 ? 20 : F >= 640 ? 10 : 5;
  if (few) T = F <= 5 * 512 ? 5 : F <= 10 * 512 ? 10 : 20;
  return T == 5 ? launch_window<5, RADIUS>(u, s, F, n_sys, sys_begin, n_vert, iterations, inv_diag, two_lambda, st)
       : T == 10 ? launch_window<10, RADIUS>(u, s, F, n_sys, sys_begin, n_vert, iterations, inv_diag, two_lambda, st)
                 : launch_window<20, RADIUS>(u, s, F, n_sys, sys_begin, n_vert, iterations, inv_diag, two_lambda, st);
}

template <int RADIUS>
static int launch_solve(const double* u, double* s, int F, int64_t n_sys, int64_t sys_begin, int64_t n_vert,
                        int radius, int iterations, const double* inv_diag, const double* two_lambda,
                        cudaStream_t st) {
  if (F > kOnChipMaxFrames) return fail(MF_E_UNSUPPORTED, "jacobi: F=%d exceeds %d frames per solve", F, kOnChipMaxFrames);
  if (radius == 10) return launch_windowed<10>(u, s, F, n_sys, sys_begin, n_vert, iterations, inv_diag, two_lambda, st);
  if (radius == 30) return launch_windowed<30>(u, s, F, n_sys, sys_begin, n_vert, iterations, inv_diag, two_lambda, st);
  const size_t smem = (size_t)(F + 2 * radius) * sizeof(double2) + (size_t)(2 * radius + 1) * sizeof(double);
  if (smem > 227 * 1024)
    return fail(MF_E_UNSUPPORTED, "jacobi: F=%d radius=%d needs %zu B of shared memory per vertex (max 232448)",
                F, radius, smem);
  const dim3 grid((unsigned)n_vert);
  if (F <= 512) {
    const int nt = (F + 31) / 32 * 32;
    auto k = jacobi_solve_kernel<RADIUS, 1>;
    if (int e = set_smem(k, smem)) return e;
    k<<<grid, nt, smem, st>>>(u, s, F, n_sys, sys_begin, radius, iterations, inv_diag, two_lambda);
  } else if (F <= 1024) {
    const int nt = ((F + 1) / 2 + 31) / 32 * 32;
    auto k = jacobi_solve_kernel<RADIUS, 2>;
    if (int e = set_smem(k, smem)) return e;
    k<<<grid, nt, smem, st>>>(u, s, F, n_sys, sys_begin, radius, iterations, inv_diag, two_lambda);
  } else if (F <= 2048) {
    const int nt = ((F + 3) / 4 + 31) / 32 * 32;
    auto k = jacobi_solve_kernel<RADIUS, 4>;
    if (int e = set_smem(k, smem)) return e;
    k<<<grid, nt, smem, st>>>(u, s, F, n_sys, sys_begin, radius, iterations, inv_diag, two_lambda);
  } else {
    auto k = jacobi_solve_long_kernel<0>;
    if (int e = set_smem(k, smem)) return e;
    k<<<grid, 512, smem, st>>>(u, s, F, n_sys, sys_begin, radius, iterations, inv_diag, two_lambda);
  }
  return check_launch("jacobi_solve");
}

}  // namespace mf

extern "C" size_t mf_jacobi_workspace_bytes(int F, int64_t n_sys) {
  if (F <= 0 || n_sys <= 0) return 0;
  size_t bytes = 2 * mf::align_up((size_t)F * sizeof(double), 256);
  // beyond the on-chip kernels' frame limit the global-memory sweeps need a second trajectory buffer
  if (F > mf::kOnChipMaxFrames) bytes += mf::global_scratch_bytes(F, n_sys);
  return bytes;
}

extern "C" int mf_jacobi_solve(const double* u, const double* homographies, double* s, int F, int64_t n_sys,
                               int64_t sys_begin, int64_t sys_end, int W, int H, int radius, int iterations,
                               int definition, double* lambda_out, void* workspace, size_t workspace_bytes,
                               void* stream) {
  MF_REQUIRE(u && homographies && s && workspace, "mf_jacobi_solve: null pointer");
  MF_REQUIRE(F > 0 && n_sys > 0 && W > 0 && H > 0, "mf_jacobi_solve: bad sizes");
  MF_REQUIRE(radius > 0 && iterations >= 0, "mf_jacobi_solve: radius must be positive, iterations >= 0");
  MF_REQUIRE(definition >= 0 && definition <= 3,
             "mf_jacobi_solve: invalid adaptive_weights_definition %d (expected 0..3)", definition);
  MF_REQUIRE((n_sys % 2) == 0 && (sys_begin % 2) == 0 && (sys_end % 2) == 0,
             "mf_jacobi_solve: systems come in (x, y) pairs; n_sys, sys_begin, sys_end must be even");
  MF_REQUIRE(0 <= sys_begin && sys_begin <= sys_end && sys_end <= n_sys, "mf_jacobi_solve: bad system range");
  MF_REQUIRE((((uintptr_t)u | (uintptr_t)s) & 15u) == 0, "mf_jacobi_solve: u and s must be 16-byte aligned (the kernels read a vertex's (x, y) as one double2)");
  if (workspace_bytes < mf_jacobi_workspace_bytes(F, n_sys))
    return mf::fail(MF_E_WORKSPACE, "mf_jacobi_solve: workspace %zu < %zu bytes", workspace_bytes,
                    mf_jacobi_workspace_bytes(F, n_sys));
  cudaStream_t st = (cudaStream_t)stream;
  double* inv_diag = (double*)workspace;
  double* two_lambda = (double*)((char*)workspace + mf::align_up((size_t)F * sizeof(double), 256));
  const int table_entries = radius <= 500 ? 10 * radius + 2 : 0;      // <= 40 KB of shared memory, else on the fly
  mf::jacobi_coeff_kernel<<<(F + 127) / 128, 128, (size_t)table_entries * sizeof(double), st>>>(
      homographies, F, W, H, radius, definition, table_entries, inv_diag, two_lambda, lambda_out);
  if (int e = mf::check_launch("jacobi_coeff")) return e;
  const int64_t n_vert = (sys_end - sys_begin) / 2;
  if (n_vert == 0) return MF_OK;
  if (mf::needs_global_path(F, radius)) {
    if (F <= mf::kOnChipMaxFrames)
      return mf::fail(MF_E_UNSUPPORTED, "mf_jacobi_solve: radius %d with %d frames does not fit shared memory", radius, F);
    void* scratch = (char*)workspace + 2 * mf::align_up((size_t)F * sizeof(double), 256);
    return mf::launch_global(u, s, F, n_sys, sys_begin, sys_end - sys_begin, radius, iterations, inv_diag, two_lambda, scratch, st);
  }
  if (n_vert > 2147483647LL) return mf::fail(MF_E_UNSUPPORTED, "mf_jacobi_solve: too many vertices");
  return mf::launch_solve<0>(u, s, F, n_sys, sys_begin, n_vert, radius, iterations, inv_diag, two_lambda, st);
}
