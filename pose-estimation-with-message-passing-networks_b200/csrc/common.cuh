// Shared host/device helpers for libpgmp.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/pgmp.h"

namespace pgmp {

// thread-local error text returned by pgmp_last_error()
char* last_error_buffer();
int set_error(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_kernel_launches;

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return PGMP_OK;
  return set_error(PGMP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define PGMP_CUDA(call)                                        \
  do {                                                         \
    int _rc = ::pgmp::check_cuda((call), #call);               \
    if (_rc != PGMP_OK) return _rc;                            \
  } while (0)

// Optional per-kernel timing (pgmp_profile_*): CUDA events recorded on the launching stream around
// every launch while profiling is enabled.
extern bool g_profiling;
void profile_before(const char* name, cudaStream_t st);
void profile_after(cudaStream_t st);

// every kernel launch goes through this so that pgmp_kernel_launches() is exact
#define PGMP_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
  do {                                                                              \
    if (::pgmp::g_profiling) ::pgmp::profile_before(#kernel, (stream));             \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                     \
    if (::pgmp::g_profiling) ::pgmp::profile_after((stream));                       \
    ::pgmp::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);              \
    int _rc = ::pgmp::check_cuda(cudaGetLastError(), #kernel);                      \
    if (_rc != PGMP_OK) return _rc;                                                 \
  } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) {
  return ceil_div(a, b) * b;
}

// bump allocator over the caller's workspace; with base == nullptr it only measures
struct Carver {
  char* base;
  uint64_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(uint64_t count) {
    off = round_up<uint64_t>(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  uint64_t bytes() const { return round_up<uint64_t>(off, 256); }
};

}  // namespace pgmp
