// Vertex-motion estimation on the device (hot-path subsystem 1).
//
//   feature_prepare : one thread per tracked candidate.  Applies the keep mask, moves the feature to
//                     frame coordinates, takes its residual velocity against the pair's global
//                     homography and writes, for every mesh row, the inclusive column range of the
//                     vertices inside the feature's ellipse (mfs.py:420-446).
//   vertex_median   : one CTA per (vertex, frame pair).  Gathers the residuals of the features whose
//                     range covers the vertex, selects the median of x and of y independently with an
//                     8-pass radix select on order-preserving 64-bit keys (statistics.median,
//                     mfs.py:338-353), adds the vertex's global velocity (mfs.py:325, 354-355).
//   median3x3       : cv2.medianBlur(.,3) on the (R+1)x(C+1) grid, replicated border (mfs.py:359-360).
//   prefix          : sequential float64 scan over frame pairs (mfs.py:271, 281).
//
// Layout in HBM: residuals [N] double2; ranges [(R+1)][N] uint32 (left | right<<16) so that the
// vertex CTAs of one mesh row stream one contiguous plane; velocities [P][V][2] float32.
#include "mf_common.cuh"
#include "mf_math.cuh"

namespace mf {

static constexpr int kMedianThreads = 128;
static constexpr uint32_t kEmptyRange = 1u;  // left = 1, right = 0

__global__ void __launch_bounds__(256) feature_prepare_kernel(
    const float* __restrict__ early_xy, const float* __restrict__ late_xy,
    const int32_t* __restrict__ offset_xy, const uint8_t* __restrict__ keep,
    const int32_t* __restrict__ pair_start, int64_t N, int P, const double* __restrict__ homographies,
    int W, int H, int R, int C, int er, int ec, double2* __restrict__ resid,
    uint32_t* __restrict__ ranges) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double2 rv = make_double2(0.0, 0.0);
  bool on = keep[i] != 0;
  FeatureCell fc;
  fc.top = 1; fc.bot = 0; fc.frow = 0.0; fc.fcol = 0.0;
  if (on) {
    // pair of this feature: last p with pair_start[p] <= i
    int lo = 0, hi = P;  // invariant: pair_start[lo] <= i < pair_start[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)pair_start[mid] <= i) lo = mid; else hi = mid;
    }
    const double* Hm = homographies + (size_t)lo * 9;
    // float32 subframe coordinate + integer subframe offset, in float64 (mfs.py:578)
    const double ex = (double)early_xy[2 * i] + (double)offset_xy[2 * i];
    const double ey = (double)early_xy[2 * i + 1] + (double)offset_xy[2 * i + 1];
    const double lx = (double)late_xy[2 * i] + (double)offset_xy[2 * i];
    const double ly = (double)late_xy[2 * i + 1] + (double)offset_xy[2 * i + 1];
    double px, py;
    persp(Hm, ex, ey, px, py);
    rv.x = MF_SUB(lx, px);
    rv.y = MF_SUB(ly, py);
    fc = feature_cell(ex, ey, W, H, R, C, er);
  }
  resid[i] = rv;
  for (int vr = 0; vr <= R; ++vr) {
    uint32_t packed = kEmptyRange;
    if (on && vr >= fc.top && vr <= fc.bot) {
      int l, r;
      col_range(fc, vr, C, er, ec, l, r);
      if (l <= r) packed = (uint32_t)l | ((uint32_t)r << 16);
    }
    ranges[(size_t)vr * N + i] = packed;
  }
}

// ---- radix select ---------------------------------------------------------------------------
struct SelectState {
  unsigned int hist[2][256];
  unsigned long long prefix[2];
  unsigned int rank[2];
  unsigned long long max_less[2];
  unsigned int count_less[2];
  unsigned int count;
};

// Item sources: the compacted shared-memory list, or (when a vertex has more candidates than the
// list holds) a re-scan of the pair's features.
struct ListSource {
  const unsigned long long* kx;
  const unsigned long long* ky;
  int n;
  template <typename F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int i = threadIdx.x; i < n; i += blockDim.x) f(kx[i], ky[i]);
  }
};

struct ScanSource {
  const uint32_t* row_ranges;  // plane of this vertex row
  const double2* resid;
  int beg, end, vc;
  template <typename F>
  __device__ __forceinline__ void for_each(F f) const {
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const uint32_t w = row_ranges[i];
      const int l = (int)(w & 0xffffu), r = (int)(w >> 16);
      if (l <= vc && vc <= r) {
        const double2 v = resid[i];
        f(key_of(v.x), key_of(v.y));
      }
    }
  }
};

// Finds the keys of rank k (0-based) among n items for both components; for even n also the key of
// rank k-1.  All threads of the CTA must call it.  Results in st.prefix (rank k) and lo_key.
template <typename Source>
__device__ void select_medians(const Source& src, SelectState& st, int n, double& med_x, double& med_y) {
  const int k = n >> 1;
  if (threadIdx.x < 2) { st.prefix[threadIdx.x] = 0ull; st.rank[threadIdx.x] = (unsigned)k; }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&st.hist[0][0])[i] = 0u;
    __syncthreads();
    const unsigned long long px = st.prefix[0], py = st.prefix[1];
    src.for_each([&](unsigned long long kx, unsigned long long ky) {
      const bool mx = (pass == 0) || ((kx >> (shift + 8)) == (px >> (shift + 8)));
      const bool my = (pass == 0) || ((ky >> (shift + 8)) == (py >> (shift + 8)));
      if (mx) atomicAdd(&st.hist[0][(unsigned)(kx >> shift) & 255u], 1u);
      if (my) atomicAdd(&st.hist[1][(unsigned)(ky >> shift) & 255u], 1u);
    });
    __syncthreads();
    // warp 0 resolves component 0, warp 1 component 1: each lane owns 8 consecutive buckets
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
      const unsigned int* h = st.hist[warp];
      unsigned int local[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { local[j] = h[lane * 8 + j]; sum += local[j]; }
      unsigned int incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      const unsigned int excl = incl - sum;
      const unsigned int want = st.rank[warp];
      if (want >= excl && want < incl) {
        // b = first bucket of this lane whose cumulative count exceeds `want`
        unsigned int run = excl;
        int b = 7;
        bool found = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (!found) {
            if (want < run + local[j]) { b = j; found = true; }
            else run += local[j];
          }
        }
        st.prefix[warp] |= ((unsigned long long)(lane * 8 + b)) << shift;
        st.rank[warp] = want - run;
      }
    }
    __syncthreads();
  }
  const unsigned long long hx = st.prefix[0], hy = st.prefix[1];
  double lo_x = value_of(hx), lo_y = value_of(hy);
  if ((n & 1) == 0) {
    if (threadIdx.x < 2) { st.max_less[threadIdx.x] = 0ull; st.count_less[threadIdx.x] = 0u; }
    __syncthreads();
    unsigned int cx = 0, cy = 0;
    unsigned long long bx = 0ull, by = 0ull;
    src.for_each([&](unsigned long long kx, unsigned long long ky) {
      if (kx < hx) { ++cx; bx = kx > bx ? kx : bx; }
      if (ky < hy) { ++cy; by = ky > by ? ky : by; }
    });
    if (cx) { atomicAdd(&st.count_less[0], cx); atomicMax(&st.max_less[0], bx); }
    if (cy) { atomicAdd(&st.count_less[1], cy); atomicMax(&st.max_less[1], by); }
    __syncthreads();
    if (st.count_less[0] == (unsigned)k) lo_x = value_of(st.max_less[0]);
    if (st.count_less[1] == (unsigned)k) lo_y = value_of(st.max_less[1]);
    med_x = MF_DIV(MF_ADD(lo_x, value_of(hx)), 2.0);
    med_y = MF_DIV(MF_ADD(lo_y, value_of(hy)), 2.0);
  } else {
    med_x = lo_x;
    med_y = lo_y;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kMedianThreads) vertex_median_kernel(
    const double2* __restrict__ resid, const uint32_t* __restrict__ ranges,
    const int32_t* __restrict__ pair_start, int64_t N, const double* __restrict__ homographies,
    const float* __restrict__ vertex_xy, int R, int C, int cap, float* __restrict__ vel_raw,
    int32_t* __restrict__ assign_count) {
  extern __shared__ unsigned long long s_keys[];  // [2][cap]
  __shared__ SelectState st;
  const int v = blockIdx.x, pair = blockIdx.y;
  const int V = (R + 1) * (C + 1);
  const int vr = v / (C + 1), vc = v - vr * (C + 1);
  const int beg = pair_start[pair], end = pair_start[pair + 1];
  const uint32_t* row_ranges = ranges + (size_t)vr * N;
  unsigned long long* kx = s_keys;
  unsigned long long* ky = s_keys + cap;
  if (threadIdx.x == 0) st.count = 0u;
  __syncthreads();
  for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
    const uint32_t w = row_ranges[i];
    const int l = (int)(w & 0xffffu), r = (int)(w >> 16);
    if (l <= vc && vc <= r) {
      const unsigned int slot = atomicAdd(&st.count, 1u);
      if (slot < (unsigned)cap) {
        const double2 rv = resid[i];
        kx[slot] = key_of(rv.x);
        ky[slot] = key_of(rv.y);
      }
    }
  }
  __syncthreads();
  const int n = (int)st.count;
  double med_x = 0.0, med_y = 0.0;
  if (n > 0) {
    if (n <= cap) {
      ListSource src{kx, ky, n};
      select_medians(src, st, n, med_x, med_y);
    } else {
      ScanSource src{row_ranges, resid, beg, end, vc};
      select_medians(src, st, n, med_x, med_y);
    }
  }
  if (threadIdx.x == 0) {
    const float vx = vertex_xy[2 * v], vy = vertex_xy[2 * v + 1];
    double gx, gy;
    persp(homographies + (size_t)pair * 9, (double)vx, (double)vy, gx, gy);
    const float glob_x = MF_FSUB((float)gx, vx);      // float32 subtraction (mfs.py:325)
    const float glob_y = MF_FSUB((float)gy, vy);
    float* out = vel_raw + ((size_t)pair * V + v) * 2;
    out[0] = (float)MF_ADD((double)glob_x, med_x);    // mfs.py:354
    out[1] = (float)MF_ADD((double)glob_y, med_y);    // mfs.py:355
    if (assign_count) assign_count[(size_t)pair * V + v] = n;
  }
}

// =================================================================================================
// Fast path: one 64-bit sort per (pair, component) + a bit-matrix selection per (pair, mesh row).
//
// The per-vertex candidate lists overlap heavily (every feature belongs to ~70 vertices), so instead
// of selecting a median per vertex from its own list, the pair's features are sorted ONCE per
// component; a vertex's median is then the member of rank n/2 in that order.  Membership of a feature
// in the vertices of one mesh row is a contiguous column range, kept as a bit mask (1 bit per
// column).  For 32 consecutive features of the sorted order a warp transposes the 32 x 32 bit matrix
// with five shuffle steps; lane c then holds the membership bits of column c and a popcount advances
// that column's running rank.
// =================================================================================================
static constexpr int kSortMax = 16384;   // features per pair the shared-memory sort holds
#ifndef MF_SELECT_WARPS
#define MF_SELECT_WARPS 16
#endif
#ifndef MF_SORT_THREADS
#define MF_SORT_THREADS 512
#endif
static constexpr int kSelectWarps = MF_SELECT_WARPS;
static constexpr int kSortThreads = MF_SORT_THREADS;       // compare-exchanges per thread and stage: npad / 2 / threads
static constexpr int kChunks = kSelectWarps / 2;

__global__ void __launch_bounds__(256) feature_prepare_masks_kernel(
    const float* __restrict__ early_xy, const float* __restrict__ late_xy,
    const int32_t* __restrict__ offset_xy, const uint8_t* __restrict__ keep,
    const int32_t* __restrict__ pair_start, int64_t N, int P, const double* __restrict__ homographies,
    int W, int H, int R, int C, int er, int ec, int nw, unsigned long long* __restrict__ keys,
    uint32_t* __restrict__ masks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  double2 rv = make_double2(0.0, 0.0);
  const bool on = keep[i] != 0;
  FeatureCell fc;
  fc.top = 1; fc.bot = 0; fc.frow = 0.0; fc.fcol = 0.0;
  if (on) {
    int lo = 0, hi = P;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((int64_t)pair_start[mid] <= i) lo = mid; else hi = mid;
    }
    const double* Hm = homographies + (size_t)lo * 9;
    const double ex = (double)early_xy[2 * i] + (double)offset_xy[2 * i];
    const double ey = (double)early_xy[2 * i + 1] + (double)offset_xy[2 * i + 1];
    const double lx = (double)late_xy[2 * i] + (double)offset_xy[2 * i];
    const double ly = (double)late_xy[2 * i + 1] + (double)offset_xy[2 * i + 1];
    double px, py;
    persp(Hm, ex, ey, px, py);
    rv.x = MF_SUB(lx, px);
    rv.y = MF_SUB(ly, py);
    fc = feature_cell(ex, ey, W, H, R, C, er);
  }
  keys[i] = key_of(rv.x);
  keys[(size_t)N + i] = key_of(rv.y);
  for (int vr = 0; vr <= R; ++vr) {
    int l = 1, r = 0;
    if (on && vr >= fc.top && vr <= fc.bot) col_range(fc, vr, C, er, ec, l, r);
    for (int w = 0; w < nw; ++w) {
      const int a = max(l, 32 * w), b = min(r, 32 * w + 31);
      uint32_t m = 0u;
      if (a <= b) m = (b - a == 31) ? 0xffffffffu : (((1u << (b - a + 1)) - 1u) << (a - 32 * w));
      masks[((size_t)vr * nw + w) * N + i] = m;
    }
  }
}

// Sort of one pair's features by one component.  Elements are single 64-bit words
// (top 48 key bits | 16-bit local index) so a compare-exchange moves one word; the rare elements whose
// keys agree in the top 48 bits are put in exact order by a final odd-even pass on the full keys.
__global__ void __launch_bounds__(kSortThreads) pair_sort_kernel(const unsigned long long* __restrict__ keys,
                                                         const int32_t* __restrict__ pair_start, int64_t N,
                                                         unsigned long long* __restrict__ sorted_keys,
                                                         uint16_t* __restrict__ perm) {
  extern __shared__ unsigned long long s_w[];
  const int pair = blockIdx.x, comp = blockIdx.y;
  const int beg = pair_start[pair], n = pair_start[pair + 1] - beg;
  if (n <= 0) return;
  const unsigned long long* k = keys + (size_t)comp * N + beg;
  int npad = 64;
  while (npad < n) npad <<= 1;
  const int nt = blockDim.x, tid = threadIdx.x;
  for (int i = tid; i < npad; i += nt)
    s_w[i] = (i < n) ? ((k[i] & ~0xffffull) | (unsigned long long)i) : ~0ull;
  __syncthreads();
  for (int kk = 2; kk <= npad; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (npad >> 1); t += nt) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int q = i | j;
        const unsigned long long a = s_w[i], b = s_w[q];
        const bool up = (i & kk) == 0;
        if ((a > b) == up) { s_w[i] = b; s_w[q] = a; }
      }
      __syncthreads();
    }
  }
  // exact order inside runs of equal 48-bit prefixes (almost never needed)
  for (int round = 0; round < n; ++round) {
    int swapped = 0;
    for (int phase = 0; phase < 2; ++phase) {
      for (int t = tid; 2 * t + phase + 1 < n; t += nt) {
        const int i = 2 * t + phase;
        const unsigned long long a = s_w[i], b = s_w[i + 1];
        if ((a >> 16) == (b >> 16)) {
          if (k[a & 0xffffull] > k[b & 0xffffull]) { s_w[i] = b; s_w[i + 1] = a; swapped = 1; }
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(swapped)) break;
  }
  for (int i = tid; i < n; i += nt) {
    const unsigned int idx = (unsigned int)(s_w[i] & 0xffffull);
    perm[(size_t)comp * N + beg + i] = (uint16_t)idx;
    sorted_keys[(size_t)comp * N + beg + i] = k[idx];
  }
}

__device__ __forceinline__ int nth_set_bit(uint32_t bits, int r) {   // position of the r-th (0-based) set bit
  for (int q = 0; q < r; ++q) bits &= bits - 1u;
  return __ffs((int)bits) - 1;
}

__device__ __forceinline__ uint32_t transpose32(uint32_t v, int lane) {
  // 32 x 32 bit-matrix transpose across the warp: on return bit i of lane c = bit c of lane i
#pragma unroll
  for (int sft = 16; sft >= 1; sft >>= 1) {
    const uint32_t m = sft == 16 ? 0x0000ffffu : sft == 8 ? 0x00ff00ffu : sft == 4 ? 0x0f0f0f0fu
                     : sft == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t o = __shfl_xor_sync(0xffffffffu, v, sft);
    v = (lane & sft) ? ((v & ~m) | ((o >> sft) & m)) : ((v & m) | ((o & m) << sft));
  }
  return v;
}

template <int NW>
__global__ void __launch_bounds__(kSelectWarps * 32) row_select_kernel(
    const uint32_t* __restrict__ masks, const unsigned long long* __restrict__ sorted_keys,
    const uint16_t* __restrict__ perm, const int32_t* __restrict__ pair_start, int64_t N,
    const double* __restrict__ homographies, const float* __restrict__ vertex_xy, int R, int C, int ncap,
    float* __restrict__ vel_raw, int32_t* __restrict__ assign_count) {
  extern __shared__ uint32_t s_mask[];              // [NW][ncap] membership words of this mesh row
  __shared__ int s_cnt[NW * 32];                    // members per column
  __shared__ int s_chunk[2][kChunks][NW * 32];      // members per column inside each chunk of the sorted order
  __shared__ int s_pos[2][2][NW * 32];              // sorted positions of the lower / upper median
  const int vr = blockIdx.x, pair = blockIdx.y;
  const int beg = pair_start[pair], n = pair_start[pair + 1] - beg;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = (R + 1) * (C + 1);
  for (int i = tid; i < NW * 32; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const int groups = (n + 31) >> 5;
  {  // phase 0: stage the row's masks, count members per column
    int c[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) c[w] = 0;
    for (int g = warp; g < groups; g += kSelectWarps) {
      const int i = g * 32 + lane;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const uint32_t m = (i < n) ? masks[((size_t)vr * NW + w) * N + beg + i] : 0u;
        if (i < n) s_mask[w * ncap + i] = m;
        c[w] += __popc(transpose32(m, lane));
      }
    }
#pragma unroll
    for (int w = 0; w < NW; ++w)
      if (c[w]) atomicAdd(&s_cnt[w * 32 + lane], c[w]);
  }
  __syncthreads();
  const int comp = warp & 1, chunk = warp >> 1;
  const int gpc = (groups + kChunks - 1) / kChunks;
  const int g0 = min(groups, chunk * gpc), g1 = min(groups, g0 + gpc);
  const uint16_t* pm = perm + (size_t)comp * N + beg;
  {  // phase 1: members per column inside this warp's chunk of the sorted order
    int c[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) c[w] = 0;
    for (int g = g0; g < g1; ++g) {
      const int j = g * 32 + lane;
      const int idx = (j < n) ? (int)pm[j] : -1;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        const uint32_t m = (idx >= 0) ? s_mask[w * ncap + idx] : 0u;
        c[w] += __popc(transpose32(m, lane));
      }
    }
#pragma unroll
    for (int w = 0; w < NW; ++w) s_chunk[comp][chunk][w * 32 + lane] = c[w];
  }
  __syncthreads();
  {  // phase 2: the chunk that contains a column's median rank(s) locates them
    int run[NW], klo[NW], khi[NW];
    bool any = false;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const int col = w * 32 + lane;
      const int ncol = s_cnt[col];
      int start = 0;
      for (int q = 0; q < chunk; ++q) start += s_chunk[comp][q][col];
      const int mine = s_chunk[comp][chunk][col];
      run[w] = start;
      khi[w] = ncol >> 1;
      klo[w] = (ncol & 1) ? khi[w] : khi[w] - 1;
      if (ncol == 0) { khi[w] = -1; klo[w] = -1; }
      any = any || (klo[w] >= start && klo[w] < start + mine) || (khi[w] >= start && khi[w] < start + mine);
    }
    if (__any_sync(0xffffffffu, any)) {
      for (int g = g0; g < g1; ++g) {
        const int j = g * 32 + lane;
        const int idx = (j < n) ? (int)pm[j] : -1;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          const uint32_t m = (idx >= 0) ? s_mask[w * ncap + idx] : 0u;
          const uint32_t bits = transpose32(m, lane);
          const int p = __popc(bits);
          if (klo[w] >= run[w] && klo[w] < run[w] + p)
            s_pos[comp][0][w * 32 + lane] = g * 32 + nth_set_bit(bits, klo[w] - run[w]);
          if (khi[w] >= run[w] && khi[w] < run[w] + p)
            s_pos[comp][1][w * 32 + lane] = g * 32 + nth_set_bit(bits, khi[w] - run[w]);
          run[w] += p;
        }
      }
    }
  }
  __syncthreads();
  // phase 3: medians -> velocities (one thread per column)
  if (tid <= C) {
    const int col = tid;
    const int ncol = s_cnt[col];
    double med[2] = {0.0, 0.0};
    if (ncol > 0) {
#pragma unroll
      for (int cp = 0; cp < 2; ++cp) {
        const unsigned long long* sk = sorted_keys + (size_t)cp * N + beg;
        const double hi = value_of(sk[s_pos[cp][1][col]]);
        if (ncol & 1) med[cp] = hi;
        else med[cp] = MF_DIV(MF_ADD(value_of(sk[s_pos[cp][0][col]]), hi), 2.0);
      }
    }
    const int v = vr * (C + 1) + col;
    const float vx = vertex_xy[2 * v], vy = vertex_xy[2 * v + 1];
    double gx, gy;
    persp(homographies + (size_t)pair * 9, (double)vx, (double)vy, gx, gy);
    const float glob_x = MF_FSUB((float)gx, vx);      // float32 subtraction (mfs.py:325)
    const float glob_y = MF_FSUB((float)gy, vy);
    float* out = vel_raw + ((size_t)pair * V + v) * 2;
    out[0] = (float)MF_ADD((double)glob_x, med[0]);   // mfs.py:354
    out[1] = (float)MF_ADD((double)glob_y, med[1]);   // mfs.py:355
    if (assign_count) assign_count[(size_t)pair * V + v] = ncol;
  }
}

__global__ void __launch_bounds__(256) median3x3_kernel(const float* __restrict__ vel_raw, int P, int R,
                                                        int C, float* __restrict__ vel_out) {
  const int V = (R + 1) * (C + 1);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)P * V) return;
  const int pair = (int)(idx / V), v = (int)(idx - (int64_t)pair * V);
  const int vr = v / (C + 1), vc = v - vr * (C + 1);
  const float* g = vel_raw + (size_t)pair * V * 2;
  float ax[9], ay[9];
  int q = 0;
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr) {
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc) {
      const int rr = min(max(vr + dr, 0), R), cc = min(max(vc + dc, 0), C);
      const float2 t = reinterpret_cast<const float2*>(g)[rr * (C + 1) + cc];
      ax[q] = t.x; ay[q] = t.y; ++q;
    }
  }
  reinterpret_cast<float2*>(vel_out)[idx] = make_float2(median9(ax), median9(ay));
}

__global__ void __launch_bounds__(128) prefix_kernel(const float* __restrict__ vel,
                                                     const double* __restrict__ disp0,
                                                     double* __restrict__ disp, int P, int64_t n) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double acc = disp0 ? disp0[j] : 0.0;
  disp[j] = acc;
  // The scan itself must stay sequential (float64 additions in frame order, mfs.py:281): one dependent DADD per
  // frame is the floor.  Everything else is taken off that chain by a three-stage software pipeline over batches
  // of kDepth frames: batch b + 2 is requested from L2, batch b + 1 (requested one iteration ago) is converted to
  // float64, batch b is added and stored.  (Measured on a 2400-frame scan -- what every rank of an 8-GPU run does:
  // 8 loads in flight, converted inside the chain: 111 us; this pipeline: see profiles/r02_prefix_probe.txt.)
  constexpr int kDepth = 16;
  const float* src = vel + j;
  float f1[kDepth], f2[kDepth];
  double d0[kDepth];
  const int nb = P / kDepth;                                 // whole batches
  auto request = [&](float (&f)[kDepth], int b) {
#pragma unroll
    for (int k = 0; k < kDepth; ++k) f[k] = src[(size_t)(b * kDepth + k) * n];
  };
  if (nb > 0) request(f1, 0);
  if (nb > 1) request(f2, 1);
  if (nb > 0) {
#pragma unroll
    for (int k = 0; k < kDepth; ++k) d0[k] = (double)f1[k];
  }
  for (int b = 0; b < nb; ++b) {
    // f2 holds batch b + 1 (if any); move it on and request batch b + 2
    if (b + 1 < nb) {
#pragma unroll
      for (int k = 0; k < kDepth; ++k) f1[k] = f2[k];
    }
    if (b + 2 < nb) request(f2, b + 2);
    double* out = disp + (size_t)(b * kDepth + 1) * n + j;
#pragma unroll
    for (int k = 0; k < kDepth; ++k) {
      acc = MF_ADD(acc, d0[k]);
      out[(size_t)k * n] = acc;
    }
    if (b + 1 < nb) {
#pragma unroll
      for (int k = 0; k < kDepth; ++k) d0[k] = (double)f1[k];
    }
  }
  for (int t = nb * kDepth; t < P; ++t) {
    acc = MF_ADD(acc, (double)src[(size_t)t * n]);
    disp[(size_t)(t + 1) * n + j] = acc;
  }
}

struct VertexMotionWorkspace {
  double2* resid;                 // fallback path
  uint32_t* ranges;               // fallback path
  float* vel_raw;
  unsigned long long* keys;       // fast path: [2][N]
  unsigned long long* sorted_keys;// fast path: [2][N]
  uint16_t* perm;                 // fast path: [2][N]
  uint32_t* masks;                // fast path: [(R+1)*nw][N]
};

static int mask_words(int C) { return (C + 1 + 31) / 32; }

// Both paths are carved (the caller sizes the workspace before it knows which one runs).
static bool carve(Carver& cv, int64_t N, int P, int R, int C, VertexMotionWorkspace& w) {
  const size_t V = (size_t)(R + 1) * (C + 1);
  const size_t n = (size_t)(N > 0 ? N : 1);
  w.vel_raw = cv.take<float>((size_t)P * V * 2);
  w.keys = cv.take<unsigned long long>(2 * n);
  w.sorted_keys = cv.take<unsigned long long>(2 * n);
  w.perm = cv.take<uint16_t>(2 * n);
  const size_t fast = n * (size_t)(R + 1) * mask_words(C);
  const size_t slow = n * (size_t)(R + 1);
  w.masks = cv.take<uint32_t>(fast > slow ? fast : slow);
  w.ranges = w.masks;             // the two paths never run together
  w.resid = reinterpret_cast<double2*>(w.keys);
  return cv.ok();
}

template <int NW>
static int launch_row_select(const VertexMotionWorkspace& w, const int32_t* pair_start, int64_t N, int P,
                             const double* homographies, const float* vertex_xy, int R, int C, int ncap,
                             int32_t* assign_count, cudaStream_t st) {
  const size_t smem = (size_t)NW * ncap * sizeof(uint32_t);
  auto k = row_select_kernel<NW>;
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(MF_E_LAUNCH, "row_select: shared memory opt-in: %s", cudaGetErrorString(e));
  }
  k<<<dim3((unsigned)(R + 1), (unsigned)P), kSelectWarps * 32, smem, st>>>(
      w.masks, w.sorted_keys, w.perm, pair_start, N, homographies, vertex_xy, R, C, ncap, w.vel_raw, assign_count);
  return check_launch("row_select");
}

}  // namespace mf

// Test hook (not part of the public header): 1 forces the generic per-vertex radix-select path.
static int mf_force_generic_vertex_motion = 0;
extern "C" void mf_debug_force_generic_vertex_motion(int on) { mf_force_generic_vertex_motion = on; }

extern "C" size_t mf_vertex_motion_workspace_bytes(int64_t N, int P, int R, int C) {
  if (N < 0 || P <= 0 || R <= 0 || C <= 0) return 0;
  mf::Carver cv(nullptr, 0);
  mf::VertexMotionWorkspace w;
  mf::carve(cv, N, P, R, C, w);
  return mf::align_up(cv.used, 256);
}

extern "C" int mf_vertex_motion(const float* early_xy, const float* late_xy, const int32_t* offset_xy,
                                const uint8_t* keep, const int32_t* pair_start,
                                const int32_t* pair_start_host, int64_t N, int P,
                                const double* homographies,
                                const float* vertex_xy, int W, int H, int R, int C, int ellipse_rows,
                                int ellipse_cols, float* vel_out, int32_t* assign_count_out,
                                void* workspace, size_t workspace_bytes, void* stream) {
  MF_REQUIRE(P > 0 && N >= 0, "mf_vertex_motion: need P > 0 and N >= 0 (P=%d, N=%lld)", P, (long long)N);
  MF_REQUIRE(W > 0 && H > 0 && R > 0 && C > 0, "mf_vertex_motion: bad frame or mesh size");
  MF_REQUIRE(R < 65535 && C < 65535, "mf_vertex_motion: mesh too large for 16-bit column ranges");
  MF_REQUIRE(ellipse_rows > 0 && ellipse_cols > 0, "mf_vertex_motion: ellipse sizes must be positive");
  MF_REQUIRE(N < 2147483647LL, "mf_vertex_motion: at most 2^31-1 features per call");
  MF_REQUIRE(pair_start && homographies && vertex_xy && vel_out && workspace,
             "mf_vertex_motion: null pointer");
  MF_REQUIRE(N == 0 || (early_xy && late_xy && offset_xy && keep), "mf_vertex_motion: null feature array");
  // the sort / select kernels size their shared memory from the largest pair, so that number must be
  // exact: it is taken from the caller's HOST copy of pair_start (absent -> generic path)
  int max_pair_features = -1;
  if (pair_start_host) {
    MF_REQUIRE(pair_start_host[0] == 0 && (int64_t)pair_start_host[P] == N,
               "mf_vertex_motion: pair_start_host must run from 0 to N");
    max_pair_features = 0;
    for (int p = 0; p < P; ++p) {
      const int c = pair_start_host[p + 1] - pair_start_host[p];
      MF_REQUIRE(c >= 0, "mf_vertex_motion: pair_start_host must be non-decreasing");
      max_pair_features = c > max_pair_features ? c : max_pair_features;
    }
  }
  mf::Carver cv(workspace, workspace_bytes);
  mf::VertexMotionWorkspace w;
  if (!mf::carve(cv, N, P, R, C, w))
    return mf::fail(MF_E_WORKSPACE, "mf_vertex_motion: workspace %zu < %zu bytes", workspace_bytes, cv.used);
  cudaStream_t st = (cudaStream_t)stream;
  const int V = (R + 1) * (C + 1);
  if (P > 65535) return mf::fail(MF_E_UNSUPPORTED, "mf_vertex_motion: at most 65535 frame pairs per call");
  const int nw = mf::mask_words(C);
  const bool fast = N > 0 && max_pair_features >= 0 && max_pair_features <= mf::kSortMax && nw <= 3 &&
                    !mf_force_generic_vertex_motion;
  if (fast) {
    const unsigned blocks = (unsigned)((N + 255) / 256);
    mf::feature_prepare_masks_kernel<<<blocks, 256, 0, st>>>(early_xy, late_xy, offset_xy, keep, pair_start, N, P,
                                                             homographies, W, H, R, C, ellipse_rows, ellipse_cols,
                                                             nw, w.keys, w.masks);
    if (int e = mf::check_launch("feature_prepare_masks")) return e;
    int npad = 64;
    while (npad < max_pair_features) npad <<= 1;
    const size_t sort_smem = (size_t)npad * sizeof(unsigned long long);
    if (sort_smem > 40 * 1024) {
      cudaError_t ce = cudaFuncSetAttribute(mf::pair_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sort_smem);
      if (ce != cudaSuccess) return mf::fail(MF_E_LAUNCH, "pair_sort: shared memory opt-in: %s", cudaGetErrorString(ce));
    }
    const int sort_threads = npad / 2 < mf::kSortThreads ? npad / 2 : mf::kSortThreads;
    mf::pair_sort_kernel<<<dim3((unsigned)P, 2), sort_threads, sort_smem, st>>>(w.keys, pair_start, N, w.sorted_keys,
                                                                                w.perm);
    if (int e = mf::check_launch("pair_sort")) return e;
    const int ncap = (max_pair_features + 31) / 32 * 32;
    int e = nw == 1 ? mf::launch_row_select<1>(w, pair_start, N, P, homographies, vertex_xy, R, C, ncap, assign_count_out, st)
          : nw == 2 ? mf::launch_row_select<2>(w, pair_start, N, P, homographies, vertex_xy, R, C, ncap, assign_count_out, st)
                    : mf::launch_row_select<3>(w, pair_start, N, P, homographies, vertex_xy, R, C, ncap, assign_count_out, st);
    if (e) return e;
  } else {
    if (N > 0) {
      const unsigned blocks = (unsigned)((N + 255) / 256);
      mf::feature_prepare_kernel<<<blocks, 256, 0, st>>>(early_xy, late_xy, offset_xy, keep, pair_start, N,
                                                         P, homographies, W, H, R, C, ellipse_rows,
                                                         ellipse_cols, w.resid, w.ranges);
      if (int e = mf::check_launch("feature_prepare")) return e;
    }
    int cap = max_pair_features < 0 ? 2048 : (max_pair_features < 64 ? 64 : max_pair_features);
    if (cap > 2816) cap = 2816;  // 2 x 2816 x 8 B = 44 KB of dynamic shared memory
    cap = (cap + 63) / 64 * 64;
    mf::vertex_median_kernel<<<dim3((unsigned)V, (unsigned)P), mf::kMedianThreads,
                               (size_t)cap * 2 * sizeof(unsigned long long), st>>>(
        w.resid, w.ranges, pair_start, N, homographies, vertex_xy, R, C, cap, w.vel_raw, assign_count_out);
    if (int e = mf::check_launch("vertex_median")) return e;
  }
  const int64_t total = (int64_t)P * V;
  mf::median3x3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w.vel_raw, P, R, C, vel_out);
  return mf::check_launch("median3x3");
}

extern "C" int mf_prefix_displacements(const float* vel, const double* disp0, double* disp, int P,
                                       int64_t n, void* stream) {
  MF_REQUIRE(P >= 0 && n > 0, "mf_prefix_displacements: bad sizes");
  MF_REQUIRE(disp && (P == 0 || vel), "mf_prefix_displacements: null pointer");
  mf::prefix_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(vel, disp0, disp, P, n);
  return mf::check_launch("prefix");
}
