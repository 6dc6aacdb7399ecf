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
constexpr int kEmbWgs = 3;                                   // warpgroups (tiles in flight) per CTA: 114 + 3 x 33 KB of shared memory
constexpr size_t kEmbSmem = kOffWgE + kEmbWgs * kWgBytesE + 64 + 1024;

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

__global__ void __launch_bounds__(kEmbWgs * kWg, 1) edge_embed_tc_kernel(const EmbArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // (offset arithmetic on the __shared__ array keeps the shared address space visible to the compiler: LDS / STS, not generic LD / ST)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(base);
  const int tid = threadIdx.x, wg = tid >> 7, wt = tid & 127, warp = tid >> 5;
  float* s_bias = reinterpret_cast<float*>(base + kOffBias);
  uint8_t* wgb = base + kOffWgE + wg * kWgBytesE;
  const uint32_t a_hi = smem_u32(wgb), a_lo = a_hi + kATile;
  uint64_t* bar = reinterpret_cast<uint64_t*>(wgb + 2 * kATile);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + kOffWgE + kEmbWgs * kWgBytesE);
  const int L = a.n_layers, LC = L + (a.w_c ? 1 : 0);

  if (warp == 0) tmem_alloc<256>(tmem_slot);
  if (wt == 0) mbar_init(bar, 1);
  if (tid == 0) fence_barrier_init();
  for (int l = 0; l < LC; ++l) {
    const __nv_bfloat16* src = l < L ? a.w + (size_t)l * 2 * kD * kD : a.w_c;
    load_weight_tile_a(sbase + l * 2 * kWTile, src, kD, kD, tid, kEmbWgs * kWg);
    load_weight_tile_a(sbase + l * 2 * kWTile + kWTile, src + kD * kD, kD, kD, tid, kEmbWgs * kWg);
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
  const int units = kEmbWgs * gridDim.x;
  const int per_unit = (total_tiles + units - 1) / units;
  const int tile_begin = (blockIdx.x * kEmbWgs + wg) * per_unit;
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
      if (wt < 32 && elect_one()) {
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
      } else {   // C rows: stage as fp32 in the (now free) operand region (the step kernel's swizzled staging
                 // layout), then copy that image out: the step kernel fetches it with one bulk copy per tile
        named_bar_sync(bar_id, kWg);
#pragma unroll
        for (int q = 0; q < kD / 4; ++q)
          sts128f(a_hi + 4 * stage_index(wt, 4 * q), make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]));
        named_bar_sync(bar_id, kWg);
        float* __restrict__ cdst = a.c0 + (size_t)tile * kTile * kD;
#pragma unroll 4
        for (int k = 0; k < 16; ++k) {
          const int idx = wt + k * kWg;
          *reinterpret_cast<float4*>(cdst + idx * 4) = lds128f(a_hi + idx * 16);   // the swizzled tile image itself
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
  if (warp == 0) tmem_dealloc<256>(*tmem_slot);
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
  const unsigned max_units = (unsigned)ceil_div<uint64_t>(w.max_slots / kTile, kEmbWgs);
  PGMP_LAUNCH(edge_embed_tc_kernel, max_units < (unsigned)sms ? max_units : (unsigned)sms, kEmbWgs * kWg, kEmbSmem, st, a);
  *done = true;
  return PGMP_OK;
}

}  // namespace pgmp

// ------------------------------------------------------------------------------------------------
// Node embedding on the tensor cores for the reference's shape 128 -> 128 -> 64 -> 64 (ReLU, ReLU, none;
// NodeClassificationMPNSimple.py:65, eval BatchNorm folded): one CTA per 128-node tile, four chained
// products (layer 1 as two 64-wide halves), h0 written as fp32 rows and as the bf16 hi/lo operand image.
// ------------------------------------------------------------------------------------------------
namespace pgmp {
namespace {

using namespace umma;

constexpr int kNeA = 2 * 2 * kATile;                 // A: 2 K-blocks x (hi, lo)
constexpr int kNeW1 = 4 * 2 * kWTile;                // W1: 2 output halves x 2 K-blocks x (hi, lo)
constexpr int kNeW2 = 2 * 2 * kWTile;                // W2: 2 K-blocks x (hi, lo)
constexpr int kNeW3 = 2 * kWTile;
constexpr size_t kNeSmem = kNeA + kNeW1 + kNeW2 + kNeW3 + 1024 + 64 + 1024;

__global__ void __launch_bounds__(kWg, 1) node_embed_tc_kernel(
    const float* __restrict__ x, int64_t sn, int64_t sc, int64_t N, const __nv_bfloat16* __restrict__ w1,
    const __nv_bfloat16* __restrict__ w2, const __nv_bfloat16* __restrict__ w3, const float* __restrict__ b1,
    const float* __restrict__ b2, const float* __restrict__ b3, float* __restrict__ h0, float* __restrict__ h0_img) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(base);
  const uint32_t a0 = sb;                              // K-block k: hi at a0 + k * 2 * kATile, lo at + kATile
  const uint32_t w1a = sb + kNeA;                      // output half o, K-block k: hi at w1a + (o * 2 + k) * 2 * kWTile
  const uint32_t w2a = w1a + kNeW1;                    // K-block k: hi at w2a + k * 2 * kWTile
  const uint32_t w3a = w2a + kNeW2;
  float* s_bias = reinterpret_cast<float*>(base + kNeA + kNeW1 + kNeW2 + kNeW3);   // b1[128] b2[64] b3[64]
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x;
  if (tid < 32) tmem_alloc<128>(tmem_slot);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  for (int o = 0; o < 2; ++o)
    for (int k = 0; k < 2; ++k) {
      load_weight_tile_a(w1a + (o * 2 + k) * 2 * kWTile, w1 + (size_t)o * 64 * 128 + k * 64, kD, 128, tid, kWg);
      load_weight_tile_a(w1a + (o * 2 + k) * 2 * kWTile + kWTile, w1 + 128 * 128 + (size_t)o * 64 * 128 + k * 64, kD, 128, tid, kWg);
    }
  for (int k = 0; k < 2; ++k) {
    load_weight_tile_a(w2a + k * 2 * kWTile, w2 + k * 64, kD, 128, tid, kWg);
    load_weight_tile_a(w2a + k * 2 * kWTile + kWTile, w2 + 64 * 128 + k * 64, kD, 128, tid, kWg);
  }
  load_weight_tile_a(w3a, w3, kD, kD, tid, kWg);
  load_weight_tile_a(w3a + kWTile, w3 + kD * kD, kD, kD, tid, kWg);
  s_bias[tid] = b1[tid];
  if (tid < kD) { s_bias[128 + tid] = b2[tid]; s_bias[192 + tid] = b3[tid]; }
  // x tile -> operand blocks (coalesced when the channel stride is 1)
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  for (int k = 0; k < 2; ++k) {
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int idx = tid + i * kWg;
      const int r = idx >> 4, c4 = idx & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < N) {
        const float* __restrict__ p = x + (row0 + r) * sn + (int64_t)(k * 64 + 4 * c4) * sc;
        v = sc == 1 ? *reinterpret_cast<const float4*>(p) : make_float4(p[0], p[sc], p[2 * sc], p[3 * sc]);
      }
      store_split4_a(a0 + k * 2 * kATile, a0 + k * 2 * kATile + kATile, r, c4, v);
    }
  }
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  float d[kD], d2[kD];
  // layer 1: two 64-wide output halves, K = 128
  if (tid < 32 && elect_one()) {
    issue_gemm_x3<kD>(tmem, a0, a0 + kATile, 2 * kATile, w1a, w1a + kWTile, 2 * kWTile, 2, false);
    issue_gemm_x3<kD>(tmem + 64, a0, a0 + kATile, 2 * kATile, w1a + 4 * kWTile, w1a + 5 * kWTile, 2 * kWTile, 2, false);
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  fence_after_sync();
  tmem_ld64(tmem, 0, d);
  tmem_ld64(tmem, 64, d2);
#pragma unroll
  for (int o = 0; o < kD; ++o) { d[o] = fmaxf(d[o] + s_bias[o], 0.f); d2[o] = fmaxf(d2[o] + s_bias[64 + o], 0.f); }
  store_split_row_a(a0, a0 + kATile, tid, d);
  store_split_row_a(a0 + 2 * kATile, a0 + 3 * kATile, tid, d2);
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  // layer 2: K = 128 -> 64
  if (tid < 32 && elect_one()) {
    fence_after_sync();
    issue_gemm_x3<kD>(tmem, a0, a0 + kATile, 2 * kATile, w2a, w2a + kWTile, 2 * kWTile, 2, false);
    mma_commit(bar);
  }
  mbar_wait(bar, 1);
  fence_after_sync();
  tmem_ld64(tmem, 0, d);
#pragma unroll
  for (int o = 0; o < kD; ++o) d[o] = fmaxf(d[o] + s_bias[128 + o], 0.f);
  store_split_row_a(a0, a0 + kATile, tid, d);
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  // layer 3: 64 -> 64, no activation
  if (tid < 32 && elect_one()) {
    fence_after_sync();
    issue_gemm_x3<kD>(tmem, a0, a0 + kATile, 0, w3a, w3a + kWTile, 0, 1, false);
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  fence_after_sync();
  tmem_ld64(tmem, 0, d);
#pragma unroll
  for (int o = 0; o < kD; ++o) d[o] += s_bias[192 + o];
  // h0 image (rows >= N are zero) through the operand tile, then coalesced copies; fp32 rows from registers
  if (row0 + tid >= N) {
#pragma unroll
    for (int o = 0; o < kD; ++o) d[o] = 0.f;
  }
  store_split_row_a(a0, a0 + kATile, tid, d);
  if (row0 + tid < N) {
    float4* __restrict__ o4 = reinterpret_cast<float4*>(h0 + (row0 + tid) * kD);
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) o4[q] = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
  }
  __syncthreads();
  uint8_t* __restrict__ img = reinterpret_cast<uint8_t*>(h0_img) + (size_t)blockIdx.x * (2 * kATile);
#pragma unroll 4
  for (int i = 0; i < 16; ++i) *reinterpret_cast<float4*>(img + (tid + i * kWg) * 16) = lds128f(a0 + (tid + i * kWg) * 16);
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<128>(tmem);
}

}  // namespace

int mpn_node_embed_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st, bool* done) {
  *done = false;
  const pgmp_mlp& m = p.node_emb;
  if (!p.tc_wnemb || m.n_layers != 3 || m.dims[0] != 128 || m.dims[1] != 128 || m.dims[2] != 64 || m.dims[3] != 64 ||
      !m.relu[0] || !m.relu[1] || m.relu[2] || m.post_relu || m.post_scale)
    return PGMP_OK;
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNeSmem));
    attr = true;
  }
  const __nv_bfloat16* wb = static_cast<const __nv_bfloat16*>(p.tc_wnemb);
  PGMP_LAUNCH(node_embed_tc_kernel, (unsigned)ceil_div<int64_t>(p.num_nodes, kTile), kWg, kNeSmem, st, p.x, p.x_stride_n,
              p.x_stride_c, p.num_nodes, wb, wb + 2 * 128 * 128, wb + 2 * 128 * 128 + 2 * 64 * 128, m.bias[0], m.bias[1],
              m.bias[2], w.h0, w.h0_img);
  *done = true;
  return PGMP_OK;
}

}  // namespace pgmp

// ------------------------------------------------------------------------------------------------
// Node and class heads on the tensor cores for the reference's shapes 64 -> 64 -> 32 -> {1, J}
// (NodeClassificationMPNSimple.py:81-83, 93-94; ReLU, ReLU, none): one CTA per 128-node tile reads the operand image
// of h, runs the two first layers as one 128-column product and the two second layers as two 32-column products;
// the last layers (32 -> 1, 32 -> J) are dot products per thread.
// ------------------------------------------------------------------------------------------------
namespace pgmp {
namespace {

using namespace umma;

constexpr int kNhA = 2 * 2 * kATile;                               // operand blocks 0 / 1, (hi, lo) each
constexpr int kNhW = 2 * 2 * kWTile + 2 * 2 * (kWTile / 2);        // nh_w1, ch_w1 (64 x 64), nh_w2, ch_w2 (32 x 64), (hi, lo) each
constexpr int kNhConst = (64 + 64 + 32 + 32 + 32 + 32 * 32 + 32) * 4;   // biases, node w3, class w3 [32][J <= 32], class b3
constexpr size_t kNhSmem = kNhA + kNhW + kNhConst + 64 + 1024;

__global__ void __launch_bounds__(kWg, 1) node_heads_tc_kernel(
    const float* __restrict__ h_img, int64_t N, int J, const __nv_bfloat16* __restrict__ w, const float* __restrict__ nb1,
    const float* __restrict__ cb1, const float* __restrict__ nb2, const float* __restrict__ cb2, const float* __restrict__ nw3,
    const float* __restrict__ nb3, const float* __restrict__ cw3, const float* __restrict__ cb3,
    float* __restrict__ node_logits, float* __restrict__ class_logits) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(base);
  const uint32_t a0 = sb, a1 = sb + 2 * kATile;                    // hi at a, lo at a + kATile
  const uint32_t nw1 = sb + kNhA, cw1 = nw1 + 2 * kWTile, nw2 = cw1 + 2 * kWTile, cw2 = nw2 + kWTile;   // lo = hi + tile bytes
  float* s_c = reinterpret_cast<float*>(base + kNhA + kNhW);
  float* s_nb1 = s_c; float* s_cb1 = s_c + 64; float* s_nb2 = s_c + 128; float* s_cb2 = s_c + 160;
  float* s_nw3 = s_c + 192; float* s_cw3 = s_c + 224; float* s_cb3 = s_c + 224 + 32 * 32;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_cb3 + 32);
  uint64_t* a_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x;
  if (tid < 32) tmem_alloc<128>(tmem_slot);
  if (tid == 0) { mbar_init(bar, 1); mbar_init(a_bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (tid < 32 && elect_one()) {
    mbar_expect_tx(a_bar, 2 * kATile);
    bulk_load(a0, reinterpret_cast<const uint8_t*>(h_img) + (size_t)blockIdx.x * (2 * kATile), 2 * kATile, a_bar);
  }
  load_weight_tile_a(nw1, w, kD, kD, tid, kWg);
  load_weight_tile_a(nw1 + kWTile, w + kD * kD, kD, kD, tid, kWg);
  load_weight_tile_a(cw1, w + 2 * kD * kD, kD, kD, tid, kWg);
  load_weight_tile_a(cw1 + kWTile, w + 3 * kD * kD, kD, kD, tid, kWg);
  load_weight_tile_a(nw2, w + 4 * kD * kD, 32, kD, tid, kWg);
  load_weight_tile_a(nw2 + kWTile / 2, w + 4 * kD * kD + 32 * kD, 32, kD, tid, kWg);
  load_weight_tile_a(cw2, w + 5 * kD * kD, 32, kD, tid, kWg);
  load_weight_tile_a(cw2 + kWTile / 2, w + 5 * kD * kD + 32 * kD, 32, kD, tid, kWg);
  if (tid < 64) { s_nb1[tid] = nb1[tid]; s_cb1[tid] = cb1[tid]; }
  if (tid < 32) { s_nb2[tid] = nb2[tid]; s_cb2[tid] = cb2[tid]; s_nw3[tid] = nw3[tid]; s_cb3[tid] = tid < J ? cb3[tid] : 0.f; }
  for (int i = tid; i < 32 * J; i += kWg) s_cw3[i] = cw3[i];
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid < 32 && elect_one()) {
    mbar_wait(a_bar, 0);
    issue_gemm_x3<kD>(tmem, a0, a0 + kATile, 0, nw1, nw1 + kWTile, 0, 1, false);
    issue_gemm_x3<kD>(tmem + 64, a0, a0 + kATile, 0, cw1, cw1 + kWTile, 0, 1, false);
    mma_commit(bar);
  }
  mbar_wait(bar, 0);
  fence_after_sync();
  float d[kD];
  tmem_ld64(tmem, 0, d);
#pragma unroll
  for (int o = 0; o < kD; ++o) d[o] = fmaxf(d[o] + s_nb1[o], 0.f);
  store_split_row_a(a0, a0 + kATile, tid, d);
  tmem_ld64(tmem, 64, d);
#pragma unroll
  for (int o = 0; o < kD; ++o) d[o] = fmaxf(d[o] + s_cb1[o], 0.f);
  store_split_row_a(a1, a1 + kATile, tid, d);
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  if (tid < 32 && elect_one()) {
    fence_after_sync();
    issue_gemm_x3<32>(tmem, a0, a0 + kATile, 0, nw2, nw2 + kWTile / 2, 0, 1, false);
    issue_gemm_x3<32>(tmem + 32, a1, a1 + kATile, 0, cw2, cw2 + kWTile / 2, 0, 1, false);
    mma_commit(bar);
  }
  mbar_wait(bar, 1);
  fence_after_sync();
  tmem_ld64(tmem, 0, d);                                           // [0, 32) node hidden, [32, 64) class hidden
  const int64_t row = (int64_t)blockIdx.x * kTile + tid;
  float nl = __ldg(nb3);
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    d[o] = fmaxf(d[o] + s_nb2[o], 0.f);
    nl = fmaf(d[o], s_nw3[o], nl);
    d[32 + o] = fmaxf(d[32 + o] + s_cb2[o], 0.f);
  }
  if (row < N) {
    node_logits[row] = nl;
    for (int j = 0; j < J; ++j) {
      float acc = s_cb3[j];
#pragma unroll
      for (int o = 0; o < 32; ++o) acc = fmaf(d[32 + o], s_cw3[o * J + j], acc);
      class_logits[row * J + j] = acc;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<128>(tmem);
}

bool heads_are_canonical(const pgmp_mlp& m, int out) {
  return m.n_layers == 3 && m.dims[0] == 64 && m.dims[1] == 64 && m.dims[2] == 32 && m.dims[3] == out && m.relu[0] && m.relu[1] &&
         !m.relu[2] && !m.post_relu && !m.post_scale;
}

}  // namespace

// node / class logits of the current h (operand image) when both heads have the reference's shape; *done = false otherwise
int mpn_node_heads_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, float* node_logits, float* class_logits,
                      cudaStream_t st, bool* done) {
  *done = false;
  if (!p.tc_wheads || p.num_classes > 32 || !heads_are_canonical(p.node_head, 1) || !heads_are_canonical(p.class_head, p.num_classes))
    return PGMP_OK;
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_heads_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNhSmem));
    attr = true;
  }
  PGMP_LAUNCH(node_heads_tc_kernel, (unsigned)ceil_div<int64_t>(p.num_nodes, kTile), kWg, kNhSmem, st, w.h_img, p.num_nodes,
              p.num_classes, static_cast<const __nv_bfloat16*>(p.tc_wheads), p.node_head.bias[0], p.class_head.bias[0],
              p.node_head.bias[1], p.class_head.bias[1], p.node_head.wt[2], p.node_head.bias[2], p.class_head.wt[2],
              p.class_head.bias[2], node_logits, class_logits);
  *done = true;
  return PGMP_OK;
}

}  // namespace pgmp
