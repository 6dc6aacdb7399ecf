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

// ATen's bilinear source index / weights, align_corners = False (area_pixel_compute_source_index + the linear
// weights of upsample_bilinear2d): src = scale * (dst + 0.5) - 0.5 clamped at 0, i1 = i0 + (i0 < in - 1), weights
// (1 - l, l), every operation rounded separately.  Shared by assemble.cu and the fused NMS loader (gc.cu).
__device__ __forceinline__ void bilinear_source_index(float scale, int dst, int in, int& i0, int& i1, float& l0, float& l1) {
  float src = __fadd_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), -0.5f);
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = __fadd_rn(src, -(float)i0);
  l0 = __fadd_rn(1.f, -l1);
}
// out = wy0 * (wx0 p00 + wx1 p01) + wy1 * (wx0 p10 + wx1 p11), products and sums rounded separately
__device__ __forceinline__ float bilinear_combine(float wx0, float wx1, float wy0, float wy1, float p00, float p01, float p10, float p11) {
  const float top = __fadd_rn(__fmul_rn(wx0, p00), __fmul_rn(wx1, p01));
  const float bot = __fadd_rn(__fmul_rn(wx0, p10), __fmul_rn(wx1, p11));
  return __fadd_rn(__fmul_rn(wy0, top), __fmul_rn(wy1, bot));
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
