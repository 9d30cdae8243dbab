// Stability score on the device (mfs.py:1216-1259).
//
// For every system (vertex component) the frame-to-frame differences p[t] = s[t+1] - s[t] form the
// "vertex profile"; the score is the energy of DFT bins 1..5 over the total energy.  The total
// comes from Parseval (sum_k |P_k|^2 = n sum_t p_t^2), the five bins from a direct DFT, so s is read
// exactly once and nothing but one ratio per system goes back.
#include "mf_common.cuh"

namespace mf {

static constexpr int kStabThreads = 128;

__global__ void __launch_bounds__(kStabThreads) stability_kernel(const double* __restrict__ s, int F,
                                                                int64_t n_sys, double* __restrict__ ratio) {
  const int64_t q = blockIdx.x;
  const int n = F - 1;
  double acc[11];  // [0] = sum p^2, [1+2k] = re_k, [2+2k] = im_k
#pragma unroll
  for (int i = 0; i < 11; ++i) acc[i] = 0.0;
  for (int t = threadIdx.x; t < n; t += kStabThreads) {
    const double p = s[(size_t)(t + 1) * n_sys + q] - s[(size_t)t * n_sys + q];
    acc[0] += p * p;
#pragma unroll
    for (int k = 1; k <= 5; ++k) {
      // angle = -2 pi k t / n; reduce k*t mod n in integers so that sincospi sees |arg| <= 2
      const long long kt = ((long long)k * t) % n;
      double sn, cs;
      sincospi(-2.0 * (double)kt / (double)n, &sn, &cs);
      acc[2 * k - 1] += p * cs;
      acc[2 * k] += p * sn;
    }
  }
  __shared__ double red[11][kStabThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 11; ++i) {
    double v = acc[i];
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) red[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[11];
    for (int i = 0; i < 11; ++i) {
      tot[i] = 0.0;
      for (int w = 0; w < kStabThreads / 32; ++w) tot[i] += red[i][w];
    }
    double low = 0.0;
    const int kmax = n - 1 < 5 ? n - 1 : 5;  // np slicing [1:6] stops at the array end
    for (int k = 1; k <= kmax; ++k) low += tot[2 * k - 1] * tot[2 * k - 1] + tot[2 * k] * tot[2 * k];
    ratio[q] = low / ((double)n * tot[0]);
  }
}

}  // namespace mf

extern "C" int mf_stability_ratios(const double* s, int F, int64_t n_sys, double* ratio_out, void* stream) {
  MF_REQUIRE(s && ratio_out, "mf_stability_ratios: null pointer");
  MF_REQUIRE(F >= 2 && n_sys > 0 && n_sys <= 2147483647LL, "mf_stability_ratios: bad sizes");
  mf::stability_kernel<<<(unsigned)n_sys, mf::kStabThreads, 0, (cudaStream_t)stream>>>(s, F, n_sys, ratio_out);
  return mf::check_launch("stability");
}
