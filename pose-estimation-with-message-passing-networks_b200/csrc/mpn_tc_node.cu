// tcgen05 / TMEM implementation of the per-node stages of a message-passing step
// (PGMP_PRECISION_TC): the per-node tables P, Q, R[t] and the node update.
//
//   tables:  [128 nodes x nd] . W_c^T for the 2 + T output chunks c (mlp_edge.0 target / source
//            columns, mlp_node[t] node columns) -- layers.py:171-175, 214, 264-274 restructured as
//            described in mpn_simt.cu.  The node tile [h0 ; h] is split into bf16 hi/lo operand
//            tiles once and reused for every chunk; weights stream through shared memory.
//   update:  h' = ReLU(sum_t Wu_t . U[:, t, :] + bu) (layers.py:253-258): for each type the merged
//            aggregation parts become the A operand of one K = 64 block, accumulated in TMEM.
#include "mpn_common.cuh"
#include "simt_mlp.cuh"
#include "umma.cuh"

namespace pgmp {
namespace {

using namespace umma;

constexpr int kABlock = kTile * 128;   // bytes of one [128][64] bf16 block
constexpr int kWBlock = kD * 128;      // bytes of one [64][64] bf16 block
constexpr int kTmemCols = 64;

__device__ __forceinline__ int stage_index(int row, int col) {   // float index into a [128][64] fp32 staging tile
  return row * kD + ((((col >> 2) ^ (row & 15)) << 2) | (col & 3));
}

// thread-per-row accumulator row -> staging tile -> coalesced global rows [row0, row0 + 128) x 64
__device__ __forceinline__ void store_rows_coalesced(float* stage, const float (&v)[kD], float* __restrict__ dst,
                                                     int64_t row0, int64_t rows) {
#pragma unroll
  for (int q = 0; q < kD / 4; ++q)
    *reinterpret_cast<float4*>(stage + stage_index(threadIdx.x, 4 * q)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  __syncthreads();
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int idx = threadIdx.x + k * kTile;
    const int r = idx >> 4, c4 = idx & 15;
    if (row0 + r < rows)
      *reinterpret_cast<float4*>(dst + (row0 + r) * kD + 4 * c4) = *reinterpret_cast<const float4*>(stage + stage_index(r, 4 * c4));
  }
  __syncthreads();
}

struct Setup {
  uint8_t* base;
  uint64_t* bar;
  uint32_t* tmem_slot;
  uint32_t tmem;
};

// common prologue: 1024-aligned carve, TMEM allocation, mbarrier
__device__ __forceinline__ Setup setup_cta(uint8_t* raw, size_t payload_bytes) {
  Setup s;
  s.base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  s.bar = reinterpret_cast<uint64_t*>(s.base + payload_bytes);
  s.tmem_slot = reinterpret_cast<uint32_t*>(s.bar + 1);
  if ((threadIdx.x >> 5) == 0) tmem_alloc<kTmemCols>(s.tmem_slot);
  if (threadIdx.x == 0) {
    mbar_init(s.bar, 1);
    fence_barrier_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  s.tmem = *s.tmem_slot;
  return s;
}
__device__ __forceinline__ void teardown_cta(const Setup& s) {
  fence_before_sync();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) tmem_dealloc<kTmemCols>(s.tmem);
}

// ------------------------------------------------------------------------------------------------
constexpr size_t kTabPayload = 2 * 2 * kABlock + 2 * 2 * kWBlock + kTile * kD * 4;   // A (2 blocks hi/lo), W, staging
constexpr size_t kTabSmem = kTabPayload + 64 + 1024;

__global__ void __launch_bounds__(kTile) node_tables_tc_kernel(
    const float* __restrict__ h0, const float* __restrict__ h, int64_t N, int skip, int T, int per_type,
    const __nv_bfloat16* __restrict__ wtab, const float* __restrict__ b1, const float* __restrict__ bm,
    float* __restrict__ tab_p, float* __restrict__ tab_q, float* __restrict__ tab_r, int chunks_per_cta) {
  extern __shared__ uint8_t smem_raw[];
  Setup s = setup_cta(smem_raw, kTabPayload);
  const int kb = skip ? 2 : 1, nd = kb * kD;
  uint8_t* a_hi = s.base;                       // kb blocks
  uint8_t* a_lo = a_hi + 2 * kABlock;
  uint8_t* w_hi = a_lo + 2 * kABlock;           // kb blocks
  uint8_t* w_lo = w_hi + 2 * kWBlock;
  float* stage = reinterpret_cast<float*>(w_lo + 2 * kWBlock);
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  for (int b = 0; b < kb; ++b) {                // [h0 ; h] (NodeClassificationMPNSimple.py:77) or [h]
    const float* __restrict__ src = (skip && b == 0) ? h0 : h;
#pragma unroll 4
    for (int k = 0; k < 16; ++k) {
      const int idx = tid + k * kTile;
      const int r = idx >> 4, c4 = idx & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < N) v = *reinterpret_cast<const float4*>(src + (row0 + r) * kD + 4 * c4);
      store_split4(a_hi + b * kABlock, a_lo + b * kABlock, r, c4, v);
    }
  }
  uint32_t phase = 0;
  const int n_chunks = 2 + T;
  const int c_begin = blockIdx.y * chunks_per_cta, c_end = min(c_begin + chunks_per_cta, n_chunks);
  for (int c = c_begin; c < c_end; ++c) {
    // weight chunk: rows = 64 outputs, nd inputs; [chunk][hi/lo][64][nd].  Chunk 2 + t uses mlp t (or 0).
    const int wc = c < 2 ? c : 2 + (per_type ? c - 2 : 0);
    const __nv_bfloat16* __restrict__ wsrc = wtab + (size_t)wc * 2 * kD * nd;
    for (int b = 0; b < kb; ++b) {
      load_weight_tile(w_hi + b * kWBlock, wsrc + b * kD, kD, nd);
      load_weight_tile(w_lo + b * kWBlock, wsrc + (size_t)kD * nd + b * kD, kD, nd);
    }
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_x3<kD>(s.tmem, smem_u32(a_hi), smem_u32(a_lo), kABlock, smem_u32(w_hi), smem_u32(w_lo), kWBlock, kb, false);
      mma_commit(s.bar);
    }
    const float* bias = c == 0 ? (skip ? nullptr : b1) : (c == 1 ? nullptr : bm + (size_t)(per_type ? c - 2 : 0) * kD);
    float* dst = c == 0 ? tab_p : (c == 1 ? tab_q : tab_r + (size_t)(c - 2) * N * kD);
    mbar_wait(s.bar, phase);
    phase ^= 1;
    fence_after_sync();
    float d[kD];
    tmem_ld64(s.tmem, 0, d);
    if (bias) {
#pragma unroll
      for (int o = 0; o < kD; ++o) d[o] += __ldg(bias + o);
    }
    store_rows_coalesced(stage, d, dst, row0, N);
  }
  teardown_cta(s);
}

// ------------------------------------------------------------------------------------------------
constexpr size_t kUpdChain = sizeof(float) * (2 * kD * kTileP + kWs);                    // SIMT head buffers (aliased)
constexpr size_t kUpdPayloadRaw = 2 * kABlock + 2 * kWBlock + kTile * kD * 4;            // A hi/lo, W hi/lo, staging
constexpr size_t kUpdPayload = kUpdPayloadRaw > kUpdChain ? kUpdPayloadRaw : kUpdChain;
constexpr size_t kUpdSmem = kUpdPayload + 64 + 1024;

__global__ void __launch_bounds__(kTile) node_update_tc_kernel(
    AggrView av, int64_t N, int T, const __nv_bfloat16* __restrict__ wu, const float* __restrict__ bu,
    float* __restrict__ h, int with_heads, const pgmp_mlp node_head, const pgmp_mlp class_head,
    float* __restrict__ node_logits, float* __restrict__ class_logits) {
  extern __shared__ uint8_t smem_raw[];
  Setup s = setup_cta(smem_raw, kUpdPayload);
  uint8_t* a_hi = s.base;
  uint8_t* a_lo = a_hi + kABlock;
  uint8_t* w_hi = a_lo + kABlock;
  uint8_t* w_lo = w_hi + kWBlock;
  float* stage = reinterpret_cast<float*>(w_lo + kWBlock);
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t row = row0 + tid;
  const int64_t srow = row < N ? row : N - 1;
  uint32_t phase = 0;
  float u[kD];
  for (int t = 0; t < T; ++t) {
    merge_parts(av, t, srow, N, u);               // U[node, t, :] from the per-tile parts
    store_split_row(a_hi, a_lo, tid, u);
    load_weight_tile(w_hi, wu + (size_t)t * 2 * kD * kD, kD, kD);
    load_weight_tile(w_lo, wu + (size_t)t * 2 * kD * kD + kD * kD, kD, kD);
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_x3<kD>(s.tmem, smem_u32(a_hi), smem_u32(a_lo), 0, smem_u32(w_hi), smem_u32(w_lo), 0, 1, t > 0);
      mma_commit(s.bar);
    }
    mbar_wait(s.bar, phase);                      // operand tiles are reused by the next type
    phase ^= 1;
  }
  fence_after_sync();
  tmem_ld64(s.tmem, 0, u);
#pragma unroll
  for (int o = 0; o < kD; ++o) u[o] = fmaxf(u[o] + __ldg(bu + o), 0.f);
  store_rows_coalesced(stage, u, h, row0, N);
  if (with_heads) {   // node / class heads (NodeClassificationMPNSimple.py:81-83, 93-94), SIMT on the CTA's tile
    float* bufA = reinterpret_cast<float*>(s.base);
    float* bufB = bufA + kD * kTileP;
    float* ws = bufB + kD * kTileP;
    put_col(bufA, u, kD, 0, false);
    __syncthreads();
    run_small_chain(node_head, bufA, bufB, ws);
    if (row < N) node_logits[row] = bufB[tid];
    __syncthreads();
    run_small_chain(class_head, bufA, bufB, ws);
    __syncthreads();
    store_tile_rowmajor(bufB, class_logits, row0, N, class_head.dims[class_head.n_layers]);
  }
  teardown_cta(s);
}

}  // namespace

int mpn_node_tables_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, const float* h, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_tables_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTabSmem));
    attr = true;
  }
  const int n_chunks = 2 + p.num_types;
  const int groups = n_chunks >= 4 ? 4 : 1;
  const int per = ceil_div(n_chunks, groups);
  PGMP_LAUNCH(node_tables_tc_kernel, dim3((unsigned)ceil_div<int64_t>(p.num_nodes, kTile), ceil_div(n_chunks, per)), kTile,
              kTabSmem, st, w.h0, h, p.num_nodes, p.skip, p.num_types, p.per_type,
              static_cast<const __nv_bfloat16*>(p.tc_wtab), p.b1, p.bm, w.tab_p, w.tab_q, w.tab_r, per);
  return PGMP_OK;
}

int mpn_node_update_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdSmem));
    attr = true;
  }
  AggrView av{w.bin_count, w.bin_lstart, w.bin_lpart, w.group_pstart, w.part_val, w.part_mx, w.part_se, p.aggr, p.attn};
  const int64_t N = p.num_nodes;
  float* nl = out_slot >= 0 ? p.node_logits + (size_t)out_slot * N : nullptr;
  float* cl = out_slot >= 0 ? p.class_logits + (size_t)out_slot * N * p.num_classes : nullptr;
  PGMP_LAUNCH(node_update_tc_kernel, (unsigned)ceil_div<int64_t>(N, kTile), kTile, kUpdSmem, st, av, N, p.num_types,
              static_cast<const __nv_bfloat16*>(p.tc_wu), p.bu, w.h, out_slot >= 0 ? 1 : 0, p.node_head, p.class_head, nl, cl);
  return PGMP_OK;
}

}  // namespace pgmp
