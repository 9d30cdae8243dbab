// Library identification and error reporting of the C ABI (include/meshflow_b200.h).
#include "mf_common.cuh"
#include "mf_math.cuh"

namespace mf {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace mf

extern "C" int mf_version(void) { return 100; }  // 0.1.0

extern "C" int mf_built_for_sm(void) { return 100; }

extern "C" const char* mf_last_error(void) { return mf::error_buffer(); }

// ---- debug hooks (not in the public header) -----------------------------------------------------
namespace mf {
__global__ void rcp_check_kernel(unsigned long long seed, long long n, unsigned long long* mismatches) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // splitmix64 -> doubles spread over [2^-40, 2^40] with random significands, both signs
  unsigned long long z = seed + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; z ^= z >> 31;
  const unsigned long long mant = z & 0x000fffffffffffffull;
  const long long ex = 1023 + (long long)((z >> 52) % 81) - 40;
  const unsigned long long sign = (z >> 63) << 63;
  double w = __longlong_as_double((long long)(sign | ((unsigned long long)ex << 52) | mant));
  if ((i & 7) == 0) w = 1.0 + (double)(long long)(z >> 40) * 1e-9;      // the remap regime: w ~ 1
  if (__double_as_longlong(rcp_rn_normal(w)) != __double_as_longlong(__drcp_rn(w))) atomicAdd(mismatches, 1ull);
}
}  // namespace mf

// Counts inputs on which the fast correctly-rounded reciprocal differs from __drcp_rn (expected: 0).
extern "C" long long mf_debug_rcp_mismatches(long long n, unsigned long long seed, void* scratch_u64, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(scratch_u64, 0, 8, st);
  mf::rcp_check_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seed, n, (unsigned long long*)scratch_u64);
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, scratch_u64, 8, cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  return (long long)h;
}
