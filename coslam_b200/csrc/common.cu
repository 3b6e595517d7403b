// common.cu -- error plumbing, version and launch accounting of libcoslam_b200.
#include "common.cuh"

namespace coslam {

std::atomic<uint64_t> g_launches{0};

std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

}  // namespace coslam

extern "C" {
const char* cosl_last_error(void) { return coslam::last_error_ref().c_str(); }
const char* cosl_version(void) { return "coslam_b200 0.1 (sm_100a)"; }
uint64_t cosl_kernel_launch_count(void) { return coslam::g_launches.load(); }
}
