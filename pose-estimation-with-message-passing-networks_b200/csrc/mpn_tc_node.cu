// tcgen05 / TMEM implementation of the per-node stages of a message-passing step
// (PGMP_PRECISION_TC): the per-node tables P, Q, R[t] and the node update.
//
// Node features are kept as bf16 hi/lo SWIZZLE_128B operand images (one 32 KB image per 128 nodes),
// so a kernel's A operand is a plain asynchronous copy.  Parallelism comes from the grid, not from
// long per-CTA loops (16 k nodes are only 128 tiles):
//   tables  grid (node tiles, 2 + T output chunks): [h0 ; h] . W_c^T -> P, Q, R[t]
//           (mlp_edge.0 target / source columns, mlp_node[t] node columns; layers.py:171-175, 214, 264-274)
//   update  grid (node tiles, 4 type groups): sum over the group's types of Wu_t . U[:, t, :] accumulated
//           in TMEM (layers.py:253-258), written as partial sums; node_finish_kernel adds the groups in
//           fixed order (deterministic), applies bias + ReLU, writes h (fp32 + image) and runs the heads.
#include "mpn_common.cuh"
#include "simt_mlp.cuh"
#include "umma.cuh"

namespace pgmp {
namespace {

using namespace umma;

constexpr int kImage = kTile * kD * 4;   // bytes of one node-tile image: hi tile (16 KB) then lo tile (16 KB)
constexpr int kHalf = kTile * 128;       // bytes of one [128][64] bf16 tile
constexpr int kWBlock = kD * 128;        // bytes of one [64][64] bf16 tile
constexpr int kTmemCols = 64;
constexpr int kGroups = 4;               // type groups of the node update for large graphs (small graphs: one per type)

struct Setup {
  uint32_t base;      // shared address of the 1024-aligned payload
  uint8_t* base_ptr;
  uint64_t* bar;
  uint32_t* tmem_slot;
  uint32_t tmem;
};

__device__ __forceinline__ Setup setup_cta(uint8_t* raw, size_t payload_bytes) {
  Setup s;
  s.base_ptr = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  s.base = smem_u32(s.base_ptr);
  s.bar = reinterpret_cast<uint64_t*>(s.base_ptr + payload_bytes);
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

// asynchronous copy of a [rows][64] bf16 block (row stride src_ld elements) into a SWIZZLE_128B tile
__device__ __forceinline__ void cp_async_weight_tile(uint32_t tile, const __nv_bfloat16* __restrict__ src, int rows, int src_ld) {
  for (int idx = threadIdx.x; idx < rows * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    cp_async16(tile + sw128_offset(r, c), src + (size_t)r * src_ld + c * 8);
  }
}

// fp32 [N][64] rows -> operand images (rows >= N are zero)
__global__ void __launch_bounds__(kTile) node_to_image_kernel(const float* __restrict__ h, int64_t N, float* __restrict__ img) {
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  uint8_t* __restrict__ out = reinterpret_cast<uint8_t*>(img) + (size_t)blockIdx.x * kImage;
#pragma unroll 4
  for (int k = 0; k < 16; ++k) {
    const int idx = threadIdx.x + k * kTile;
    const int r = idx >> 4, c4 = idx & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) v = *reinterpret_cast<const float4*>(h + (row0 + r) * kD + 4 * c4);
    store_split4(out, out + kHalf, r, c4, v);
  }
}

// ------------------------------------------------------------------------------------------------
constexpr size_t kTabPayload = 2 * kImage + 4 * kWBlock;   // [h0 image][h image], W: 2 K-blocks x (hi, lo)
constexpr size_t kTabSmem = kTabPayload + 64 + 1024;

__global__ void __launch_bounds__(kTile) node_tables_tc_kernel(
    const float* __restrict__ h0_img, const float* __restrict__ h_img, int64_t N, int skip, int per_type, int n_chunks,
    int chunks_per_cta, const __nv_bfloat16* __restrict__ wtab, const float* __restrict__ b1,
    const float* __restrict__ bm, float* __restrict__ tab_p, float* __restrict__ tab_q, float* __restrict__ tab_r) {
  extern __shared__ uint8_t smem_raw[];
  Setup s = setup_cta(smem_raw, kTabPayload);
  const int kb = skip ? 2 : 1, nd = kb * kD;
  const uint32_t a0 = s.base;                     // K-block b: hi at a0 + b * kImage, lo at + kHalf
  const uint32_t w_hi = s.base + 2 * kImage;      // K-block b at + b * kWBlock
  const uint32_t w_lo = w_hi + 2 * kWBlock;
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  for (int b = 0; b < kb; ++b) {                  // [h0 ; h] (NodeClassificationMPNSimple.py:77) or [h]
    const uint8_t* __restrict__ src = reinterpret_cast<const uint8_t*>((skip && b == 0) ? h0_img : h_img) + (size_t)blockIdx.x * kImage;
#pragma unroll
    for (int k = 0; k < 16; ++k) cp_async16(a0 + b * kImage + (tid + k * kTile) * 16, src + (tid + k * kTile) * 16);
  }
  uint32_t phase = 0;
  const int c_begin = blockIdx.y * chunks_per_cta, c_end = min(c_begin + chunks_per_cta, n_chunks);
  for (int c = c_begin; c < c_end; ++c) {
    // weight chunk c: [chunk][hi/lo][64][nd]; chunk 2 + t uses message MLP t (or the single agnostic one).
    // The previous chunk's MMA has completed (waited below), so the weight tiles can be overwritten.
    const int wc = c < 2 ? c : 2 + (per_type ? c - 2 : 0);
    const __nv_bfloat16* __restrict__ wsrc = wtab + (size_t)wc * 2 * kD * nd;
    for (int b = 0; b < kb; ++b) {
      cp_async_weight_tile(w_hi + b * kWBlock, wsrc + b * kD, kD, nd);
      cp_async_weight_tile(w_lo + b * kWBlock, wsrc + (size_t)kD * nd + b * kD, kD, nd);
    }
    const float* bias = c == 0 ? (skip ? nullptr : b1) : (c == 1 ? nullptr : bm + (size_t)(per_type ? c - 2 : 0) * kD);
    float* dst = c == 0 ? tab_p : (c == 1 ? tab_q : tab_r + (size_t)(c - 2) * N * kD);
    cp_async_wait_all();
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_x3<kD>(s.tmem, a0, a0 + kHalf, kImage, w_hi, w_lo, kWBlock, kb, false);
      mma_commit(s.bar);
    }
    float bv[kD];
#pragma unroll
    for (int o = 0; o < kD; ++o) bv[o] = bias ? __ldg(bias + o) : 0.f;
    mbar_wait(s.bar, phase);
    phase ^= 1;
    fence_after_sync();
    float d[kD];
    tmem_ld64(s.tmem, 0, d);
    if (row0 + tid < N) {
      float4* __restrict__ o4 = reinterpret_cast<float4*>(dst + (row0 + tid) * kD);
#pragma unroll
      for (int q = 0; q < kD / 4; ++q)
        o4[q] = make_float4(d[4 * q] + bv[4 * q], d[4 * q + 1] + bv[4 * q + 1], d[4 * q + 2] + bv[4 * q + 2], d[4 * q + 3] + bv[4 * q + 3]);
    }
  }
  teardown_cta(s);
}

// ------------------------------------------------------------------------------------------------
constexpr size_t kUpdPayload = 2 * kHalf + 2 * kWBlock;   // A hi/lo, W hi/lo
constexpr size_t kUpdSmem = kUpdPayload + 64 + 1024;

__global__ void __launch_bounds__(kTile) node_update_tc_kernel(AggrView av, int64_t N, int64_t Np, int T, int groups,
                                                               const __nv_bfloat16* __restrict__ wu,
                                                               float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  Setup s = setup_cta(smem_raw, kUpdPayload);
  const uint32_t a_hi = s.base, a_lo = a_hi + kHalf, w_hi = a_lo + kHalf, w_lo = w_hi + kWBlock;
  const int tid = threadIdx.x, grp = blockIdx.y;
  const int per = (T + groups - 1) / groups;
  const int t0 = grp * per, t1 = min(t0 + per, T);
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t row = row0 + tid;
  const int64_t srow = row < N ? row : N - 1;
  uint32_t phase = 0;
  float u[kD];
  for (int t = t0; t < t1; ++t) {
    cp_async_weight_tile(w_hi, wu + (size_t)t * 2 * kD * kD, kD, kD);
    cp_async_weight_tile(w_lo, wu + (size_t)t * 2 * kD * kD + kD * kD, kD, kD);
    merge_parts(av, t, srow, N, u);               // U[node, t, :] from the per-tile parts
    store_split_row_a(a_hi, a_lo, tid, u);
    cp_async_wait_all();
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      issue_gemm_x3<kD>(s.tmem, a_hi, a_lo, 0, w_hi, w_lo, 0, 1, t > t0);
      mma_commit(s.bar);
    }
    mbar_wait(s.bar, phase);                      // operand tiles are reused by the next type
    phase ^= 1;
  }
  if (t1 > t0) {
    fence_after_sync();
    tmem_ld64(s.tmem, 0, u);
  } else {
#pragma unroll
    for (int o = 0; o < kD; ++o) u[o] = 0.f;
  }
  float4* __restrict__ o4 = reinterpret_cast<float4*>(partial + ((size_t)grp * Np + row) * kD);
#pragma unroll
  for (int q = 0; q < kD / 4; ++q) o4[q] = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
  teardown_cta(s);
}

// h' = ReLU(sum of the group partials + bias) -> fp32 rows + operand image; node / class heads when reported
constexpr int kFinThreads = 512;

// THREADS = 512 for the plain finish (4 items per thread, one round trip), 128 when the heads run (one thread per node)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) node_finish_kernel(const float* __restrict__ partial, int64_t N, int64_t Np,
                                                                  int groups, const float* __restrict__ bu,
                                                                  float* __restrict__ h, float* __restrict__ h_img,
                                                                  int with_heads, const pgmp_mlp node_head,
                                                                  const pgmp_mlp class_head, float* __restrict__ node_logits,
                                                                  float* __restrict__ class_logits) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;                   // [64][kTileP], only for the heads
  float* bufB = bufA + kD * kTileP;
  float* ws = bufB + kD * kTileP;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  uint8_t* __restrict__ img = reinterpret_cast<uint8_t*>(h_img) + (size_t)blockIdx.x * kImage;
  for (int base = 0; base < kTile * 16; base += 4 * THREADS) {
  float4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = base + threadIdx.x + k * THREADS;
    v[k] = __ldg(reinterpret_cast<const float4*>(partial + (row0 + (idx >> 4)) * kD + 4 * (idx & 15)));
  }
  for (int g = 1; g < groups; ++g) {
    float4 w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = base + threadIdx.x + k * THREADS;
      w[k] = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)g * Np + row0 + (idx >> 4)) * kD + 4 * (idx & 15)));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k].x += w[k].x; v[k].y += w[k].y; v[k].z += w[k].z; v[k].w += w[k].w; }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = base + threadIdx.x + k * THREADS;
    const int r = idx >> 4, c4 = idx & 15;
    const float4 b = __ldg(reinterpret_cast<const float4*>(bu + 4 * c4));
    float4 o = make_float4(fmaxf(v[k].x + b.x, 0.f), fmaxf(v[k].y + b.y, 0.f), fmaxf(v[k].z + b.z, 0.f), fmaxf(v[k].w + b.w, 0.f));
    if (row0 + r >= N) o = make_float4(0.f, 0.f, 0.f, 0.f);
    else *reinterpret_cast<float4*>(h + (row0 + r) * kD + 4 * c4) = o;
    store_split4(img, img + kHalf, r, c4, o);
    if (with_heads) {
      bufA[(size_t)(4 * c4 + 0) * kTileP + r] = o.x; bufA[(size_t)(4 * c4 + 1) * kTileP + r] = o.y;
      bufA[(size_t)(4 * c4 + 2) * kTileP + r] = o.z; bufA[(size_t)(4 * c4 + 3) * kTileP + r] = o.w;
    }
  }
  }
  if (!with_heads) return;
  __syncthreads();
  const int64_t row = row0 + threadIdx.x;
  run_small_chain(node_head, bufA, bufB, ws);     // NodeClassificationMPNSimple.py:81-83, 93-94
  if (row < N) node_logits[row] = bufB[threadIdx.x];
  __syncthreads();
  run_small_chain(class_head, bufA, bufB, ws);
  __syncthreads();
  store_tile_rowmajor(bufB, class_logits, row0, N, class_head.dims[class_head.n_layers]);
}

}  // namespace

int mpn_node_image(const MpnWorkspace& w, const float* h, int64_t N, float* img, cudaStream_t st) {
  PGMP_LAUNCH(node_to_image_kernel, (unsigned)ceil_div<int64_t>(N, kTile), kTile, 0, st, h, N, img);
  return PGMP_OK;
}

int mpn_node_tables_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, const float* h_img, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_tables_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTabSmem));
    attr = true;
  }
  const int n_chunks = 2 + p.num_types;
  // A tile loaded once per ~5 output chunks on large graphs; one chunk per CTA when there are few node tiles
  const int per = (n_chunks >= 8 && p.num_nodes > 4096) ? ceil_div(n_chunks, 4) : 1;
  PGMP_LAUNCH(node_tables_tc_kernel, dim3((unsigned)ceil_div<int64_t>(p.num_nodes, kTile), ceil_div(n_chunks, per)), kTile,
              kTabSmem, st, w.h0_img, h_img, p.num_nodes, p.skip, p.per_type, n_chunks, per,
              static_cast<const __nv_bfloat16*>(p.tc_wtab), p.b1, p.bm, w.tab_p, w.tab_q, w.tab_r);
  return PGMP_OK;
}

int mpn_node_update_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st) {
  static bool attr = false;
  const size_t fin_smem = sizeof(float) * (2 * kD * kTileP + kWs);
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_update_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdSmem));
    PGMP_CUDA(cudaFuncSetAttribute(node_finish_kernel<kTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    attr = true;
  }
  AggrView av{w.bin_count, w.bin_lstart, w.bin_lpart, w.group_pstart, w.part_val, w.part_mx, w.part_se, p.aggr, p.attn};
  const int64_t N = p.num_nodes, Np = round_up<int64_t>(N, kTile);
  // small graphs have too few node tiles to fill the GPU: one CTA per (tile, type); large ones group ~5 types per CTA
  const int groups = mpn_update_groups(p);
  const unsigned tiles = (unsigned)ceil_div<int64_t>(N, kTile);
  PGMP_LAUNCH(node_update_tc_kernel, dim3(tiles, groups), kTile, kUpdSmem, st, av, N, Np, p.num_types, groups,
              static_cast<const __nv_bfloat16*>(p.tc_wu), w.upd_partial);
  float* nl = out_slot >= 0 ? p.node_logits + (size_t)out_slot * N : nullptr;
  float* cl = out_slot >= 0 ? p.class_logits + (size_t)out_slot * N * p.num_classes : nullptr;
  if (out_slot >= 0) {
    PGMP_LAUNCH((node_finish_kernel<kTile>), tiles, kTile, fin_smem, st, w.upd_partial, N, Np, groups, p.bu, w.h, w.h_img, 1,
                p.node_head, p.class_head, nl, cl);
  } else {
    PGMP_LAUNCH((node_finish_kernel<kFinThreads>), tiles, kFinThreads, 0, st, w.upd_partial, N, Np, groups, p.bu, w.h, w.h_img,
                0, p.node_head, p.class_head, nl, cl);
  }
  return PGMP_OK;
}

}  // namespace pgmp
