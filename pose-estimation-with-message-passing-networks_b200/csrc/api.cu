// Library-wide entry points of libpgmp.so (version, error text, launch counter).
#include "common.cuh"

#include <cstring>

namespace pgmp {

std::atomic<uint64_t> g_kernel_launches{0};

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace pgmp

extern "C" int pgmp_version(void) { return PGMP_VERSION; }
extern "C" const char* pgmp_last_error(void) { return pgmp::last_error_buffer(); }
extern "C" uint64_t pgmp_kernel_launches(void) { return pgmp::g_kernel_launches.load(); }
