// Library identification and error reporting of the C ABI (include/meshflow_b200.h).
#include "mf_common.cuh"

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
