// Shared plumbing of the C-ABI library: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include "../../include/meshflow_b200.h"

namespace mf {

char* error_buffer();                       // thread-local, 512 bytes (cabi.cu)
int fail(int code, const char* fmt, ...);   // formats into error_buffer(), returns code

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MF_E_LAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return MF_OK;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace; every block is 256-byte aligned.
struct Carver {
  char* base;
  size_t used, cap;
  Carver(void* p, size_t bytes) : base((char*)p), used(0), cap(bytes) {}
  template <typename T>
  T* take(size_t count) {
    used = align_up(used, 256);
    T* p = (T*)(base + used);
    used += count * sizeof(T);
    return p;
  }
  bool ok() const { return base != nullptr ? used <= cap : used == 0; }
};

}  // namespace mf

#define MF_REQUIRE(cond, ...) \
  do { if (!(cond)) return mf::fail(MF_E_INVALID, __VA_ARGS__); } while (0)
