// tcgen05 / TMEM implementation of the per-edge message-passing step (PGMP_PRECISION_TC).
//
// One persistent CTA (128 threads = 128 TMEM lanes = 128 edge slots per tile) walks a contiguous
// range of tiles.  Per tile the three per-edge 64x64 products of a step
//     hidden = ReLU(W1_e g + [C + P[dst] + Q[src]])       (layers.py:171-175, 214)
//     g'     = ReLU(W2 hidden + b2)
//     m      = ReLU(Wm_e[type(src)] g' + R[type][dst])     (layers.py:222-224, 264-274)
// run on the tensor cores: the A operand (g, hidden, g') is written by the threads into a
// SWIZZLE_128B shared-memory tile as a bf16 hi/lo pair, the weights stay resident in shared memory
// (the per-type message matrix is swapped when the tile's source type changes), the fp32
// accumulator lives in tensor memory and is read back one row per thread for the fused epilogue:
// bias / gathered per-node terms, ReLU, the attention logit, the in-tile per-(target, type)
// softmax / sum / max reduction (no atomics) and the write-back of g'.
#include "mpn_common.cuh"
#include "umma.cuh"

namespace pgmp {

int mpn_embed(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st);
int mpn_node_tables(const pgmp_mpn_params& p, const MpnWorkspace& w, const float* h, cudaStream_t st);
int mpn_node_update(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st);
int mpn_edge_head(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st);
int mpn_node_tables_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, const float* h, cudaStream_t st);
int mpn_node_update_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st);

namespace {

using namespace umma;

constexpr int kATile = kTile * 128;        // bytes of one [128][64] bf16 operand tile
constexpr int kWTile = kD * 128;           // bytes of one [64][64] bf16 weight tile
constexpr int kTmemCols = 64;

struct EdgeTcArgs {
  const int32_t* slot_edge; const int32_t* slot_src; const int32_t* slot_dst;
  const int32_t* group_start; const int32_t* group_pstart; const int32_t* bin_lstart; const int32_t* bin_lpart;
  float* g; const float* c0; const float* tab_p; const float* tab_q; const float* tab_r;
  const __nv_bfloat16* w1; const __nv_bfloat16* w2; const __nv_bfloat16* wm;   // [2][64][64], [2][64][64], [Tm][2][64][64]
  const float* b2; const float* wa; const float* ba;
  float* part_val; float* part_mx; float* part_se;
  int64_t N;
  int T, per_type, aggr, attn, attn_cols;
};

struct TcSmem {
  uint8_t* a_hi; uint8_t* a_lo;
  uint8_t* w1_hi; uint8_t* w1_lo; uint8_t* w2_hi; uint8_t* w2_lo; uint8_t* wm_hi; uint8_t* wm_lo;
  float* m;          // [128][64] fp32, 16-byte chunks XOR-swizzled by row; aliases a_hi / a_lo
  float* att; int* dst; float* b2; float* wa;
  uint64_t* bar; uint32_t* tmem;
};
constexpr size_t kTcSmemBytes = 2 * kATile + 6 * kWTile + sizeof(float) * (kTile + kD + kD) + sizeof(int) * kTile + 64 + 1024;

__device__ __forceinline__ TcSmem carve_smem(uint8_t* raw) {
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  TcSmem s;
  s.a_hi = base; s.a_lo = base + kATile;
  s.w1_hi = base + 2 * kATile; s.w1_lo = s.w1_hi + kWTile; s.w2_hi = s.w1_lo + kWTile; s.w2_lo = s.w2_hi + kWTile;
  s.wm_hi = s.w2_lo + kWTile; s.wm_lo = s.wm_hi + kWTile;
  s.m = reinterpret_cast<float*>(base);
  uint8_t* misc = s.wm_lo + kWTile;
  s.att = reinterpret_cast<float*>(misc);
  s.dst = reinterpret_cast<int*>(s.att + kTile);
  s.b2 = reinterpret_cast<float*>(s.dst + kTile);
  s.wa = s.b2 + kD;
  s.bar = reinterpret_cast<uint64_t*>(s.wa + kD);
  s.tmem = reinterpret_cast<uint32_t*>(s.bar + 1);
  return s;
}

__device__ __forceinline__ int m_index(int row, int col) {   // float index into the swizzled m tile
  return row * kD + ((((col >> 2) ^ (row & 15)) << 2) | (col & 3));
}

// all threads: publish shared-memory operand writes and TMEM reads, then one thread issues the GEMM
__device__ __forceinline__ void sync_and_issue(const TcSmem& s, uint32_t tmem, uint8_t* w_hi, uint8_t* w_lo) {
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_after_sync();
    issue_gemm_x3<kD>(tmem, smem_u32(s.a_hi), smem_u32(s.a_lo), 0, smem_u32(w_hi), smem_u32(w_lo), 0, 1, false);
    mma_commit(s.bar);
  }
}

__global__ void __launch_bounds__(kTile) edge_step_tc_kernel(const EdgeTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const TcSmem s = carve_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<kTmemCols>(s.tmem);
  if (tid == 0) {
    mbar_init(s.bar, 1);
    fence_barrier_init();
  }
  load_weight_tile(s.w1_hi, a.w1, kD, kD);
  load_weight_tile(s.w1_lo, a.w1 + kD * kD, kD, kD);
  load_weight_tile(s.w2_hi, a.w2, kD, kD);
  load_weight_tile(s.w2_lo, a.w2 + kD * kD, kD, kD);
  if (tid < kD) {
    s.b2[tid] = a.b2[tid];
    s.wa[tid] = 0.f;
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s.tmem;
  uint32_t phase = 0;
  int cur_tm = -1, cur_col = -1;

  const int total_tiles = a.group_start[a.T] >> 7;
  const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int tile_begin = blockIdx.x * per_cta;
  const int tile_end = min(tile_begin + per_cta, total_tiles);

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int64_t slot0 = (int64_t)tile * kTile;
    int t = 0;
    while (t + 1 < a.T && slot0 >= a.group_start[t + 1]) ++t;
    const int tm = a.per_type ? t : 0;
    const int col = a.attn == PGMP_ATTN_PER_TYPE ? t : 0;
    if (tm != cur_tm) {     // the previous tile's MMAs have completed (we waited on them)
      load_weight_tile(s.wm_hi, a.wm + (size_t)tm * 2 * kD * kD, kD, kD);
      load_weight_tile(s.wm_lo, a.wm + (size_t)tm * 2 * kD * kD + kD * kD, kD, kD);
      cur_tm = tm;
    }
    if (a.attn && col != cur_col) {
      if (tid < kD) s.wa[tid] = a.wa[tid * a.attn_cols + col];
      cur_col = col;
    }
    // ---- A <- split(g tile), coalesced 16-byte loads
    {
      const float4* __restrict__ g4 = reinterpret_cast<const float4*>(a.g + slot0 * kD);
#pragma unroll 4
      for (int k = 0; k < 16; ++k) {
        const int idx = tid + k * kTile;
        store_split4(s.a_hi, s.a_lo, idx >> 4, idx & 15, g4[idx]);
      }
    }
    sync_and_issue(s, tmem, s.w1_hi, s.w1_lo);

    const int64_t slot = slot0 + tid;
    const int e = a.slot_edge[slot];
    const int src = a.slot_src[slot], dst = a.slot_dst[slot];
    float add[kD], d[kD];
    if (e >= 0) {     // per-edge constant + per-node tables, fetched while the MMA runs
      const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.tab_p + (size_t)dst * kD);
      const float4* __restrict__ q4 = reinterpret_cast<const float4*>(a.tab_q + (size_t)src * kD);
      const float4* __restrict__ c4 = a.c0 ? reinterpret_cast<const float4*>(a.c0 + slot * kD) : nullptr;
#pragma unroll
      for (int q = 0; q < kD / 4; ++q) {
        const float4 p = p4[q], sq = q4[q];
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4) c = c4[q];
        add[4 * q + 0] = c.x + p.x + sq.x;
        add[4 * q + 1] = c.y + p.y + sq.y;
        add[4 * q + 2] = c.z + p.z + sq.z;
        add[4 * q + 3] = c.w + p.w + sq.w;
      }
    } else {
#pragma unroll
      for (int o = 0; o < kD; ++o) add[o] = 0.f;
    }
    mbar_wait(s.bar, phase);
    phase ^= 1;
    fence_after_sync();
    tmem_ld64(tmem, 0, d);
#pragma unroll
    for (int o = 0; o < kD; ++o) d[o] = fmaxf(d[o] + add[o], 0.f);
    store_split_row(s.a_hi, s.a_lo, tid, d);
    sync_and_issue(s, tmem, s.w2_hi, s.w2_lo);

    if (e >= 0) {     // R[type][dst] for the message, fetched while the MMA runs
      const float4* __restrict__ r4 = reinterpret_cast<const float4*>(a.tab_r + ((size_t)t * a.N + dst) * kD);
#pragma unroll
      for (int q = 0; q < kD / 4; ++q) {
        const float4 r = r4[q];
        add[4 * q + 0] = r.x; add[4 * q + 1] = r.y; add[4 * q + 2] = r.z; add[4 * q + 3] = r.w;
      }
    }
    mbar_wait(s.bar, phase);
    phase ^= 1;
    fence_after_sync();
    tmem_ld64(tmem, 0, d);
    float att = a.attn ? __ldg(a.ba + col) : 0.f;
#pragma unroll
    for (int o = 0; o < kD; ++o) {
      d[o] = fmaxf(d[o] + s.b2[o], 0.f);
      att = fmaf(d[o], s.wa[o], att);
    }
    {   // write back g' (row per thread)
      float4* __restrict__ o4 = reinterpret_cast<float4*>(a.g + slot * kD);
#pragma unroll
      for (int q = 0; q < kD / 4; ++q) o4[q] = make_float4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
    }
    store_split_row(s.a_hi, s.a_lo, tid, d);
    sync_and_issue(s, tmem, s.wm_hi, s.wm_lo);
    mbar_wait(s.bar, phase);
    phase ^= 1;
    fence_after_sync();
    tmem_ld64(tmem, 0, d);
    // ---- message -> shared (aliases the operand tiles: all MMAs reading them are complete)
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
      float4 v;
      v.x = fmaxf(d[4 * q + 0] + add[4 * q + 0], 0.f);
      v.y = fmaxf(d[4 * q + 1] + add[4 * q + 1], 0.f);
      v.z = fmaxf(d[4 * q + 2] + add[4 * q + 2], 0.f);
      v.w = fmaxf(d[4 * q + 3] + add[4 * q + 3], 0.f);
      *reinterpret_cast<float4*>(s.m + m_index(tid, 4 * q)) = v;
    }
    s.att[tid] = att;
    s.dst[tid] = e >= 0 ? dst : -1;
    __syncthreads();
    // ---- reduce every run of equal targets (a bin, or the part of it inside this tile)
    if (e >= 0 && (tid == 0 || s.dst[tid - 1] != dst)) {
      int r1 = tid;
      while (r1 + 1 < kTile && s.dst[r1 + 1] == dst) ++r1;
      const int64_t bin = (int64_t)t * a.N + dst;
      const int first_slot = a.group_start[t] + a.bin_lstart[bin];
      const int64_t prow = (int64_t)a.group_pstart[t] + a.bin_lpart[bin] + (tile - (first_slot >> 7));
      float u[kD];
      if (a.attn) {
        float mx = -INFINITY;
        for (int r = tid; r <= r1; ++r) mx = fmaxf(mx, s.att[r]);
        float se = 0.f;
#pragma unroll
        for (int o = 0; o < kD; ++o) u[o] = 0.f;
        for (int r = tid; r <= r1; ++r) {
          const float wgt = __expf(s.att[r] - mx);
          se += wgt;
#pragma unroll
          for (int q = 0; q < kD / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(s.m + m_index(r, 4 * q));
            u[4 * q + 0] = fmaf(wgt, v.x, u[4 * q + 0]);
            u[4 * q + 1] = fmaf(wgt, v.y, u[4 * q + 1]);
            u[4 * q + 2] = fmaf(wgt, v.z, u[4 * q + 2]);
            u[4 * q + 3] = fmaf(wgt, v.w, u[4 * q + 3]);
          }
        }
        a.part_mx[prow] = mx;
        a.part_se[prow] = se;
      } else {
#pragma unroll
        for (int q = 0; q < kD / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(s.m + m_index(tid, 4 * q));
          u[4 * q + 0] = v.x; u[4 * q + 1] = v.y; u[4 * q + 2] = v.z; u[4 * q + 3] = v.w;
        }
        for (int r = tid + 1; r <= r1; ++r) {
#pragma unroll
          for (int q = 0; q < kD / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(s.m + m_index(r, 4 * q));
            if (a.aggr == PGMP_AGGR_MAX) {
              u[4 * q + 0] = fmaxf(u[4 * q + 0], v.x); u[4 * q + 1] = fmaxf(u[4 * q + 1], v.y);
              u[4 * q + 2] = fmaxf(u[4 * q + 2], v.z); u[4 * q + 3] = fmaxf(u[4 * q + 3], v.w);
            } else {
              u[4 * q + 0] += v.x; u[4 * q + 1] += v.y; u[4 * q + 2] += v.z; u[4 * q + 3] += v.w;
            }
          }
        }
      }
      float4* __restrict__ o4 = reinterpret_cast<float4*>(a.part_val + prow * kD);
#pragma unroll
      for (int q = 0; q < kD / 4; ++q) o4[q] = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
    }
    __syncthreads();   // the next tile overwrites the operand / message region
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

// D = A . W^T through the same building blocks (pgmp_selftest_umma)
__global__ void __launch_bounds__(kTile) selftest_umma_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                               float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  const TcSmem s = carve_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<kTmemCols>(s.tmem);
  if (tid == 0) {
    mbar_init(s.bar, 1);
    fence_barrier_init();
  }
  const float4* __restrict__ a4 = reinterpret_cast<const float4*>(A);
  for (int k = 0; k < 16; ++k) {
    const int idx = tid + k * kTile;
    store_split4(s.a_hi, s.a_lo, idx >> 4, idx & 15, a4[idx]);
  }
  const float4* __restrict__ w4 = reinterpret_cast<const float4*>(W);
  for (int k = 0; k < 8; ++k) {
    const int idx = tid + k * kTile;
    store_split4(s.w1_hi, s.w1_lo, idx >> 4, idx & 15, w4[idx]);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *s.tmem;
  sync_and_issue(s, tmem, s.w1_hi, s.w1_lo);
  mbar_wait(s.bar, 0);
  fence_after_sync();
  float d[kD];
  tmem_ld64(tmem, 0, d);
  for (int o = 0; o < kD; ++o) D[tid * kD + o] = d[o];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kTmemCols>(tmem);
}

}  // namespace

int mpn_forward_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st) {
  if (!p.tc_w1_e || !p.tc_w2 || !p.tc_wm_e || !p.tc_wtab || (p.has_update_mlp && !p.tc_wu))
    return set_error(PGMP_ERR_INVALID, "PGMP_PRECISION_TC needs the bf16 hi/lo weights");
  // the node update runs on the tensor cores when it is a matrix product (update MLP), else it is a plain merge
  auto node_update = [&](int out_slot) {
    return p.has_update_mlp ? mpn_node_update_tc(p, w, out_slot, st) : mpn_node_update(p, w, out_slot, st);
  };
  const int64_t N = p.num_nodes, E = p.num_edges;
  int rc;
  if ((rc = mpn_embed(p, w, st)) != PGMP_OK) return rc;
  PGMP_CUDA(cudaMemcpyAsync(w.h, w.h0, sizeof(float) * N * kD, cudaMemcpyDeviceToDevice, st));
  PGMP_CUDA(cudaFuncSetAttribute(edge_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  EdgeTcArgs a;
  a.slot_edge = w.slot_edge; a.slot_src = w.slot_src; a.slot_dst = w.slot_dst;
  a.group_start = w.group_start; a.group_pstart = w.group_pstart; a.bin_lstart = w.bin_lstart; a.bin_lpart = w.bin_lpart;
  a.g = w.g; a.c0 = p.skip ? w.c0 : nullptr; a.tab_p = w.tab_p; a.tab_q = w.tab_q; a.tab_r = w.tab_r;
  a.w1 = static_cast<const __nv_bfloat16*>(p.tc_w1_e); a.w2 = static_cast<const __nv_bfloat16*>(p.tc_w2);
  a.wm = static_cast<const __nv_bfloat16*>(p.tc_wm_e);
  a.b2 = p.b2; a.wa = p.wa; a.ba = p.ba;
  a.part_val = w.part_val; a.part_mx = w.part_mx; a.part_se = w.part_se;
  a.N = N; a.T = p.num_types; a.per_type = p.per_type; a.aggr = p.aggr; a.attn = p.attn;
  a.attn_cols = p.attn == PGMP_ATTN_PER_TYPE ? 17 : 1;
  const unsigned max_tiles = (unsigned)(w.max_slots / kTile);
  const unsigned grid = max_tiles < (unsigned)(2 * sms) ? max_tiles : (unsigned)(2 * sms);
  const int first_out = p.steps - p.aux_loss_steps - 1 > 0 ? p.steps - p.aux_loss_steps - 1 : 0;
  for (int s = 0; s < p.steps; ++s) {
    if (s > 0) {
      const int prev_slot = (s - 1) >= first_out ? (s - 1) - first_out : -1;
      if ((rc = node_update(prev_slot)) != PGMP_OK) return rc;
    }
    if ((rc = mpn_node_tables_tc(p, w, w.h, st)) != PGMP_OK) return rc;
    if (E > 0) {
      PGMP_LAUNCH(edge_step_tc_kernel, grid, kTile, kTcSmemBytes, st, a);
      if (s >= first_out && (rc = mpn_edge_head(p, w, s - first_out, st)) != PGMP_OK) return rc;
    }
  }
  return node_update((p.steps - 1) - first_out);
}

}  // namespace pgmp

extern "C" int pgmp_selftest_umma(const float* a, const float* w, float* d, pgmp_stream_t stream) {
  using namespace pgmp;
  if (!a || !w || !d) return set_error(PGMP_ERR_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGMP_CUDA(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  PGMP_LAUNCH(selftest_umma_kernel, 1, kTile, kTcSmemBytes, st, a, w, d);
  return PGMP_OK;
}
