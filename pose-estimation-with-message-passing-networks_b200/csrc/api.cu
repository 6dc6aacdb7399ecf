// Library-wide entry points of libpgmp.so (version, error text, launch counter).
#include "common.cuh"

#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// ---- per-kernel CUDA-event timing -------------------------------------------------------------
bool g_profiling = false;
namespace {
struct ProfRecord { const char* name; cudaEvent_t beg, end; };
std::vector<ProfRecord> g_records;
std::vector<cudaEvent_t> g_event_pool;
std::mutex g_prof_mutex;
cudaEvent_t get_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

void profile_before(const char* name, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  ProfRecord r{name, get_event(), get_event()};
  cudaEventRecord(r.beg, st);
  g_records.push_back(r);
}
void profile_after(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  cudaEventRecord(g_records.back().end, st);
}

}  // namespace pgmp

extern "C" void pgmp_profile_enable(int on) { pgmp::g_profiling = on != 0; }

// Synchronises the device, then writes one line per kernel name: "<name> <launches> <total_ms>\n".
// Returns the number of bytes written (truncated to `size`); clears the records.
extern "C" int pgmp_profile_collect(char* out, int size) {
  using namespace pgmp;
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  std::map<std::string, std::pair<int, double>> agg;
  for (auto& r : g_records) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.beg, r.end) == cudaSuccess) {
      auto& a = agg[r.name];
      a.first += 1;
      a.second += ms;
    }
    g_event_pool.push_back(r.beg);
    g_event_pool.push_back(r.end);
  }
  g_records.clear();
  std::string text;
  char line[256];
  for (auto& kv : agg) {
    snprintf(line, sizeof line, "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    text += line;
  }
  int n = (int)text.size() < size - 1 ? (int)text.size() : size - 1;
  if (n < 0) n = 0;
  if (out && size > 0) { memcpy(out, text.data(), n); out[n] = 0; }
  return n;
}

extern "C" int pgmp_version(void) { return PGMP_VERSION; }
extern "C" const char* pgmp_last_error(void) { return pgmp::last_error_buffer(); }
extern "C" uint64_t pgmp_kernel_launches(void) { return pgmp::g_kernel_launches.load(); }
