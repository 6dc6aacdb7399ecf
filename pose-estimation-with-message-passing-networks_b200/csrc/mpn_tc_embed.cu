// tcgen05 / TMEM edge embedding (PGMP_PRECISION_TC): the _make_mlp chain of NodeClassificationMPNSimple.py:66
// (layers.py:8-29, eval BatchNorm folded) evaluated per 128-slot tile as a chain of 128x64x64 products
// (layer widths <= 64 are zero-padded), followed by the step-invariant per-edge constant
// C = W1_e0 . g0 + b1.  Outputs: g0 as the bf16 hi/lo operand image the step kernel consumes, C as fp32 rows.
#include "mpn_common.cuh"
#include "umma.cuh"

namespace pgmp {
namespace {

using namespace umma;

constexpr int kATile = kTile * 128;
constexpr int kWTile = kD * 128;
constexpr int kWg = 128;
constexpr int kMaxL = PGMP_MAX_LAYERS;                     // chain layers; + 1 for the C product
constexpr int kOffBias = (kMaxL + 1) * 2 * kWTile;         // [(kMaxL + 1)][64] floats
constexpr int kOffWgE = kOffBias + 2048;                   // operand tiles must stay 1024-byte aligned
constexpr int kWgBytesE = 2 * kATile + 1024;               // A hi/lo (also the fp32 staging of C) + mbarrier
static_assert((kMaxL + 1) * kD * 4 <= 2048 && kOffWgE % 1024 == 0 && kWgBytesE % 1024 == 0, "alignment");
constexpr size_t kEmbSmem = kOffWgE + 2 * kWgBytesE + 64 + 1024;

struct EmbArgs {
  const float* edge_attr; int F;                            // [E][F]
  const int32_t* slot_edge; const int32_t* group_start; int T;
  int n_layers;                                             // chain layers (all widths <= 64)
  int relu[kMaxL];
  const float* bias[kMaxL]; int width[kMaxL];               // output width of layer l
  const __nv_bfloat16* w;                                   // [n_layers][2][64][64] zero-padded, BatchNorm folded
  const __nv_bfloat16* w_c; const float* b_c;               // W1_e0 hi/lo [2][64][64], b1 (NULL when !skip)
  float* g_img; float* c0;
};

__device__ __forceinline__ int stage_index(int row, int col) {
  return row * kD + ((((col >> 2) ^ (row & 15)) << 2) | (col & 3));
}

__global__ void __launch_bounds__(2 * kWg, 1) edge_embed_tc_kernel(const EmbArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the shared address space visible to the compiler: LDS / STS, not generic LD / ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(base);
  const int tid = threadIdx.x, wg = tid >> 7, wt = tid & 127, warp = tid >> 5;
  float* s_bias = reinterpret_cast<float*>(base + kOffBias);
  uint8_t* wgb = base + kOffWgE + wg * kWgBytesE;
  const uint32_t a_hi = smem_u32(wgb), a_lo = a_hi + kATile;
  uint64_t* bar = reinterpret_cast<uint64_t*>(wgb + 2 * kATile);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + kOffWgE + 2 * kWgBytesE);
  const int L = a.n_layers, LC = L + (a.w_c ? 1 : 0);

  if (warp == 0) tmem_alloc<128>(tmem_slot);
  if (wt == 0) mbar_init(bar, 1);
  if (tid == 0) fence_barrier_init();
  for (int l = 0; l < LC; ++l) {
    const __nv_bfloat16* src = l < L ? a.w + (size_t)l * 2 * kD * kD : a.w_c;
    load_weight_tile_a(sbase + l * 2 * kWTile, src, kD, kD, tid, 2 * kWg);
    load_weight_tile_a(sbase + l * 2 * kWTile + kWTile, src + kD * kD, kD, kD, tid, 2 * kWg);
    if (tid < kD) {
      const float* b = l < L ? a.bias[l] : a.b_c;
      const int wdt = l < L ? a.width[l] : kD;
      s_bias[l * kD + tid] = (b && tid < wdt) ? b[tid] : 0.f;
    }
  }
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot + (uint32_t)(wg * 64);
  const int bar_id = 1 + wg;
  uint32_t phase = 0;

  const int total_tiles = a.group_start[a.T] >> 7;
  const int units = 2 * gridDim.x;
  const int per_unit = (total_tiles + units - 1) / units;
  const int tile_begin = (blockIdx.x * 2 + wg) * per_unit;
  const int tile_end = min(tile_begin + per_unit, total_tiles);
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int64_t slot = (int64_t)tile * kTile + wt;
    const int e = a.slot_edge[slot];
    float d[kD];
#pragma unroll
    for (int o = 0; o < kD; ++o) d[o] = 0.f;
    if (e >= 0) {
      const float* __restrict__ row = a.edge_attr + (size_t)e * a.F;
#pragma unroll
      for (int o = 0; o < kD; ++o)
        if (o < a.F) d[o] = __ldg(row + o);
    }
    store_split_row_a(a_hi, a_lo, wt, d);
    for (int l = 0; l < LC; ++l) {
      fence_before_sync();
      fence_async_smem();
      named_bar_sync(bar_id, kWg);
      if (wt == 0) {
        fence_after_sync();
        issue_gemm_x3<kD>(tmem, a_hi, a_lo, 0, sbase + l * 2 * kWTile, sbase + l * 2 * kWTile + kWTile, 0, 1, false);
        mma_commit(bar);
      }
      if (l == L) {   // the operand tiles now hold g0: copy the image out while the C product runs
        uint8_t* __restrict__ gdst = reinterpret_cast<uint8_t*>(a.g_img) + (size_t)tile * (2 * kATile);
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
          const int idx = wt + k * kWg;
          *reinterpret_cast<float4*>(gdst + idx * 16) = lds128f(a_hi + idx * 16);
        }
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      fence_after_sync();
      tmem_ld64(tmem, 0, d);
      const bool relu = l < L && a.relu[l];
#pragma unroll
      for (int o = 0; o < kD; ++o) {
        const float v = d[o] + s_bias[l * kD + o];
        d[o] = relu ? fmaxf(v, 0.f) : v;
      }
      if (l < L) {
        store_split_row_a(a_hi, a_lo, wt, d);   // every MMA reading the operand tiles has completed
      } else {   // C rows: stage as fp32 in the (now free) operand region, then coalesced stores
        named_bar_sync(bar_id, kWg);
#pragma unroll
        for (int q = 0; q < kD / 4; ++q)
          sts128f(a_hi + 4 * stage_index(wt, 4 * q), make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]));
        named_bar_sync(bar_id, kWg);
        float* __restrict__ cdst = a.c0 + (size_t)tile * kTile * kD;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
          const int idx = wt + k * kWg;
          *reinterpret_cast<float4*>(cdst + idx * 4) = lds128f(a_hi + 4 * stage_index(idx >> 4, (idx & 15) * 4));
        }
      }
    }
    if (!a.w_c) {   // no skip connection: the image of g0 still has to be written
      fence_before_sync();
      named_bar_sync(bar_id, kWg);
      uint8_t* __restrict__ gdst = reinterpret_cast<uint8_t*>(a.g_img) + (size_t)tile * (2 * kATile);
#pragma unroll 4
      for (int k = 0; k < 16; ++k) {
        const int idx = wt + k * kWg;
        *reinterpret_cast<float4*>(gdst + idx * 16) = lds128f(a_hi + idx * 16);
      }
    }
    fence_before_sync();
    named_bar_sync(bar_id, kWg);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(*tmem_slot);
}

}  // namespace

// returns PGMP_OK and sets *done = true when the tensor-core embedding applies to this model
int mpn_edge_embed_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st, bool* done) {
  *done = false;
  const pgmp_mlp& m = p.edge_emb;
  if (!p.tc_wemb || m.n_layers > kMaxL || m.post_relu || m.post_scale) return PGMP_OK;
  for (int l = 0; l <= m.n_layers; ++l)
    if (m.dims[l] > kD) return PGMP_OK;
  if (p.skip && !p.tc_w1_e0) return PGMP_OK;
  EmbArgs a;
  a.edge_attr = p.edge_attr; a.F = m.dims[0];
  a.slot_edge = w.slot_edge; a.group_start = w.group_start; a.T = p.num_types;
  a.n_layers = m.n_layers;
  for (int l = 0; l < kMaxL; ++l) {
    a.relu[l] = l < m.n_layers ? m.relu[l] : 0;
    a.bias[l] = l < m.n_layers ? m.bias[l] : nullptr;
    a.width[l] = l < m.n_layers ? m.dims[l + 1] : 0;
  }
  a.w = static_cast<const __nv_bfloat16*>(p.tc_wemb);
  a.w_c = p.skip ? static_cast<const __nv_bfloat16*>(p.tc_w1_e0) : nullptr;
  a.b_c = p.skip ? p.b1 : nullptr;
  a.g_img = w.g; a.c0 = w.c0;
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(edge_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmbSmem));
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned max_units = (unsigned)ceil_div<uint64_t>(w.max_slots / kTile, 2);
  PGMP_LAUNCH(edge_embed_tc_kernel, max_units < (unsigned)sms ? max_units : (unsigned)sms, 2 * kWg, kEmbSmem, st, a);
  *done = true;
  return PGMP_OK;
}

}  // namespace pgmp
