// SIMT fp32 building blocks shared by mpn_simt.cu and mpn_tc.cu: thread-per-row MLP layers with
// activations transposed in shared memory ([feature][row], stride kTileP) and weights streamed through a
// shared staging buffer; merge of per-tile aggregation parts.
#pragma once

#include "mpn_common.cuh"

namespace pgmp {

constexpr int kWs = 32 * kD;   // floats of the weight staging buffer: 32 input rows x 64 outputs

__device__ __forceinline__ void stage_w(const float* __restrict__ Wt, int K, int O, int k0, int o0,
                                        float* __restrict__ ws) {
  for (int idx = threadIdx.x; idx < kWs; idx += blockDim.x) {
    const int k = idx >> 6, o = idx & 63;
    ws[idx] = (k0 + k < K && o0 + o < O) ? __ldg(Wt + (size_t)(k0 + k) * O + o0 + o) : 0.f;
  }
}

// acc[0..63] += Wt[0:K, o0:o0+64]^T . in[:, row]   (in: shared, [k][kTileP]; all threads of the CTA call this)
__device__ __forceinline__ void matvec64(float (&acc)[kD], const float* __restrict__ in, int K,
                                         const float* __restrict__ Wt, int O, int o0, float* __restrict__ ws) {
  for (int k0 = 0; k0 < K; k0 += 32) {
    __syncthreads();
    stage_w(Wt, K, O, k0, o0, ws);
    __syncthreads();
    const int kc = min(32, K - k0);
    const float* __restrict__ col = in + (size_t)k0 * kTileP + threadIdx.x;
    for (int k = 0; k < kc; ++k) {
      const float a = col[k * kTileP];
      const float4* __restrict__ w4 = reinterpret_cast<const float4*>(ws + k * kD);
#pragma unroll
      for (int q = 0; q < kD / 4; ++q) {
        const float4 w = w4[q];
        acc[4 * q + 0] = fmaf(a, w.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(a, w.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(a, w.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(a, w.w, acc[4 * q + 3]);
      }
    }
  }
}

__device__ __forceinline__ void init_bias(float (&acc)[kD], const float* __restrict__ bias, int O, int o0) {
#pragma unroll
  for (int o = 0; o < kD; ++o) acc[o] = (bias != nullptr && o0 + o < O) ? __ldg(bias + o0 + o) : 0.f;
}

// out[(o0+o)][row] = act(acc[o]) for o0+o < O
__device__ __forceinline__ void put_col(float* __restrict__ out, const float (&acc)[kD], int O, int o0, bool relu) {
#pragma unroll
  for (int o = 0; o < kD; ++o)
    if (o0 + o < O) out[(size_t)(o0 + o) * kTileP + threadIdx.x] = relu ? fmaxf(acc[o], 0.f) : acc[o];
}

// [rows][ld] global (row-major) <- / -> transposed shared tile, coalesced on the global side
__device__ __forceinline__ void load_tile_rowmajor(float* __restrict__ buf, const float* __restrict__ src, int64_t row0,
                                                   int64_t rows, int width) {
  for (int idx = threadIdx.x; idx < kTile * width; idx += blockDim.x) {
    const int r = idx / width, c = idx - r * width;
    buf[(size_t)c * kTileP + r] = (row0 + r < rows) ? src[(row0 + r) * width + c] : 0.f;
  }
}
__device__ __forceinline__ void store_tile_rowmajor(const float* __restrict__ buf, float* __restrict__ dst, int64_t row0,
                                                    int64_t rows, int width) {
  for (int idx = threadIdx.x; idx < kTile * width; idx += blockDim.x) {
    const int r = idx / width, c = idx - r * width;
    if (row0 + r < rows) dst[(row0 + r) * width + c] = buf[(size_t)c * kTileP + r];
  }
}

// A _make_mlp chain whose layer widths are all <= 64, evaluated on the CTA's tile.  The first layer
// reads `in` and writes `work`; later layers run in place on `work` (safe: a thread reads only its
// own row and has consumed all inputs before it writes).  Returns with the result in `work`.
static __device__ void run_small_chain(const pgmp_mlp& m, const float* __restrict__ in, float* __restrict__ work,
                                float* __restrict__ ws) {
  float acc[kD];
  const float* cur = in;
  for (int l = 0; l < m.n_layers; ++l) {
    init_bias(acc, m.bias[l], m.dims[l + 1], 0);
    matvec64(acc, cur, m.dims[l], m.wt[l], m.dims[l + 1], 0, ws);
    put_col(work, acc, m.dims[l + 1], 0, m.relu[l] != 0);
    cur = work;
  }
  const int O = m.dims[m.n_layers];
  if (m.post_relu || m.post_scale) {
    for (int o = 0; o < O; ++o) {
      float v = work[(size_t)o * kTileP + threadIdx.x];
      if (m.post_relu) v = fmaxf(v, 0.f);
      if (m.post_scale) v = fmaf(v, __ldg(m.post_scale + o), __ldg(m.post_shift + o));
      work[(size_t)o * kTileP + threadIdx.x] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Merge the parts of one (type, node) bin into the aggregated message U[node, type, :]
// (layers.py:234-251): softmax-weighted sum (attention), sum, max or mean.  Empty bins give 0.
// ------------------------------------------------------------------------------------------------
struct AggrView {
  const int32_t* bin_count;
  const int32_t* bin_lstart;
  const int32_t* bin_lpart;
  const int32_t* group_pstart;
  const float* part_val;
  const float* part_mx;
  const float* part_se;
  int aggr, attn;
};

__device__ __forceinline__ void merge_parts(const AggrView& a, int t, int64_t node, int64_t N, float (&u)[kD]) {
#pragma unroll
  for (int o = 0; o < kD; ++o) u[o] = 0.f;
  const int64_t bin = (int64_t)t * N + node;
  const int cnt = a.bin_count[bin];
  if (cnt == 0) return;
  const int ls = a.bin_lstart[bin];
  const int np = ((ls + cnt - 1) >> 7) - (ls >> 7) + 1;
  const int64_t p0 = (int64_t)a.group_pstart[t] + a.bin_lpart[bin];
  if (a.attn) {
    float M = -INFINITY;
    for (int i = 0; i < np; ++i) M = fmaxf(M, a.part_mx[p0 + i]);
    float den = 0.f;
    for (int i = 0; i < np; ++i) {
      const float sc = np == 1 ? 1.f : __expf(a.part_mx[p0 + i] - M);
      den = fmaf(a.part_se[p0 + i], sc, den);
      const float4* __restrict__ v4 = reinterpret_cast<const float4*>(a.part_val + (p0 + i) * kD);
#pragma unroll
      for (int q = 0; q < kD / 4; ++q) {
        const float4 v = v4[q];
        u[4 * q + 0] = fmaf(v.x, sc, u[4 * q + 0]);
        u[4 * q + 1] = fmaf(v.y, sc, u[4 * q + 1]);
        u[4 * q + 2] = fmaf(v.z, sc, u[4 * q + 2]);
        u[4 * q + 3] = fmaf(v.w, sc, u[4 * q + 3]);
      }
    }
    const float inv = 1.f / (den + 1e-12f);   // torch_scatter softmax eps
#pragma unroll
    for (int o = 0; o < kD; ++o) u[o] *= inv;
    return;
  }
  for (int i = 0; i < np; ++i) {
    const float4* __restrict__ v4 = reinterpret_cast<const float4*>(a.part_val + (p0 + i) * kD);
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
      const float4 v = v4[q];
      if (a.aggr == PGMP_AGGR_MAX) {
        u[4 * q + 0] = i == 0 ? v.x : fmaxf(u[4 * q + 0], v.x);
        u[4 * q + 1] = i == 0 ? v.y : fmaxf(u[4 * q + 1], v.y);
        u[4 * q + 2] = i == 0 ? v.z : fmaxf(u[4 * q + 2], v.z);
        u[4 * q + 3] = i == 0 ? v.w : fmaxf(u[4 * q + 3], v.w);
      } else {
        u[4 * q + 0] += v.x; u[4 * q + 1] += v.y; u[4 * q + 2] += v.z; u[4 * q + 3] += v.w;
      }
    }
  }
  if (a.aggr == PGMP_AGGR_MEAN) {
    const float inv = 1.f / (float)cnt;
#pragma unroll
    for (int o = 0; o < kD; ++o) u[o] *= inv;
  }
}


}  // namespace pgmp
