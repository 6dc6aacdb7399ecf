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
#include <cstdlib>

#include "mpn_common.cuh"
#include "umma.cuh"

namespace pgmp {

int mpn_embed(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st);
int mpn_node_tables(const pgmp_mpn_params& p, const MpnWorkspace& w, const float* h, cudaStream_t st);
int mpn_node_update(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st);
int mpn_edge_head(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, bool image, cudaStream_t st);
int mpn_node_tables_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, const float* h_img, cudaStream_t st);
int mpn_node_image(const MpnWorkspace& w, const float* h, int64_t N, float* img, cudaStream_t st);
int mpn_edge_embed_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st, bool* done);
int mpn_embed_nodes(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st);
int mpn_node_embed_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st, bool* done);
int mpn_node_update_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st);
int mpn_node_update_hier(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st);

namespace {

using namespace umma;

constexpr int kATile = kTile * 128;        // bytes of one [128][64] bf16 operand tile
constexpr int kWTile = kD * 128;           // bytes of one [64][64] bf16 weight tile
constexpr int kAddTile = kTile * kD * 4;   // bytes of one [128][64] fp32 staging tile
constexpr int kTmemCols = 64;              // selftest kernel
constexpr int kWgThreads = 128;            // one warpgroup = one tile in flight
constexpr int kEdgeThreads = 2 * kWgThreads;
constexpr int kEdgeTmemCols = 256;         // 2 warpgroups x (message / hidden accumulator + head accumulator)

struct EdgeTcArgs {
  const int32_t* slot_edge; const int32_t* slot_src; const int32_t* slot_dst;
  const int2* slot_run; const int32_t* tile_seg;      // step-invariant run bookkeeping (mpn_prep.cu: slot_runs_kernel)
  const int32_t* group_start; const int32_t* group_pstart;
  float* g; const float* c0; const float* tab_p; const float* tab_q; const float* tab_r;
  const __nv_bfloat16* w1; const __nv_bfloat16* w2; const __nv_bfloat16* wm;   // [2][64][64], [2][64][64], [Tm][2][64][64]
  const float* b2; const float* wa; const float* ba;
  float* part_val; float* part_mx; float* part_se;
  int64_t N;
  int T, per_type, aggr, attn, attn_cols;
  // fused edge head 64 -> 64 -> 32 -> 1 (NodeClassificationMPNSimple.py:84), only when with_head
  int with_head;
  const __nv_bfloat16* wh1; const __nv_bfloat16* wh2;   // [2][64][64], [2][32][64]
  const float* bh1; const float* bh2; const float* wh3; const float* bh3;
  float* edge_logits;
  int one_group;
};

// ---- shared-memory map of the edge kernel (offsets from the 1024-aligned base) ----------------------
constexpr int kOffW1 = 0;                              // hi, lo
constexpr int kOffW2 = kOffW1 + 2 * kWTile;
constexpr int kOffWh1 = kOffW2 + 2 * kWTile;
constexpr int kOffWh2 = kOffWh1 + 2 * kWTile;          // [32][64] hi, lo
constexpr int kOffConst = kOffWh2 + 2 * (kWTile / 2);  // b2[64] bh1[64] bh2[32] wh3[32] (floats), group_start[18], group_pstart[18]
constexpr int kOffWg = kOffConst + 1024;               // per-warpgroup regions follow
constexpr int kWgA = 0;                                // A hi, lo
constexpr int kWgAdd = kWgA + 2 * kATile;              // fp32 staging: g in, C+P+Q, R, then the messages
constexpr int kWgWm = kWgAdd + kAddTile;               // message weights hi, lo
constexpr int kWgMisc = kWgWm + 2 * kWTile;            // dst[128] src/prow[128] att[128] wa[64] bars[3] split
constexpr int kWgBytes = kWgMisc + 2048;
constexpr size_t kEdgeSmemBytes = kOffWg + 2 * kWgBytes + 64 + 1024;
static_assert(kOffWg % 1024 == 0 && kWgBytes % 1024 == 0 && kWgAdd % 1024 == 0 && kWgWm % 1024 == 0, "operand tiles must be 1024-byte aligned");

// ---- the step kernel: TWO threads per edge slot (column halves), 2 tiles in flight = 512 threads = 16 warps per SM ------
// Thread (row, half) owns columns [32 half, 32 half + 32) of its row -- warps w and w + 4 of a tile group share the TMEM
// lane quarter w & 3, so both may read it.  Against one thread per row (round 1: a 64-wide accumulator row per thread,
// 255 registers, 8 warps per SM) this halves the registers and the instruction stream per thread and doubles the warps
// the schedulers can switch between (ncu warps active 12.5 % -> 25 %, issue slots 36 % -> 43 % busy).  Row-wise quantities
// that need the whole row (attention logit) are summed through shared memory.
// Measured with in-kernel phase timers (17 000 cycles per tile, round 2): ~3 100 issuing the P / Q gathers (16-byte
// chunks of 32 different rows per request: the L1 data pipe is the busiest unit, 55 %), ~1 100 per product between its
// operands being ready and its result being visible (x 3), ~3 900 in the three epilogues, ~2 000 in the run scan,
// ~2 400 in the run reduction; with the three table gathers switched off (timing experiment) a launch takes 0.197 ms
// instead of 0.227 ms, i.e. the chain of dependent phases with two tiles in flight per SM is the limit, not the gathers.
// Tried and rejected, with numbers, in DESIGN.md: cooperative (line-wide) table gathers
// through the staging tile, requesting the next tile's features a whole tile early, L2 prefetch of the next tile's rows.
constexpr int kTgThreads = 256;                 // one tile group = one tile in flight
constexpr int kEdge2Threads = 2 * kTgThreads;
constexpr int kWg2Misc = kWgMisc;               // cw[128] (int2) wa[64] bars[4] scan[12] seg[32] att0[128] att1[128]
constexpr int kWg2Bytes = kWgMisc + 3072;
constexpr size_t kEdge2SmemBytes = kOffWg + 2 * kWg2Bytes + 64 + 1024;
static_assert(kWg2Bytes % 1024 == 0, "operand tiles must be 1024-byte aligned");
static_assert(kEdge2SmemBytes <= 227 * 1024, "edge step: shared memory");

#ifdef PGMP_TIMELINE
__device__ long long g_step_tl[4][32][32];      // development aid: clock64 at 16 points of a tile, 2 CTAs x 2 tile groups
#define TL(k) do { if (tl_on && tl_i < 32) g_step_tl[tl_slot][tl_i][k] = clock64(); } while (0)
#else
#define TL(k) do { } while (0)
#endif

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  tmem_ld16(taddr, v);
  tmem_ld16(taddr + 16, v + 16);
  tmem_ld_wait();
}

__global__ void __launch_bounds__(kEdge2Threads, 1) edge_step_tc_kernel(const EdgeTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(base);
  const int tid = threadIdx.x, tg = tid >> 8, tt = tid & 255, warp = tid >> 5;
  const int row = tt & 127, half = tt >> 7, c0col = 32 * half;      // this thread: row of the tile, columns [c0col, c0col + 32)
  const uint32_t w1_hi = sbase + kOffW1, w1_lo = w1_hi + kWTile, w2_hi = sbase + kOffW2, w2_lo = w2_hi + kWTile;
  const uint32_t wh1_hi = sbase + kOffWh1, wh1_lo = wh1_hi + kWTile, wh2_hi = sbase + kOffWh2, wh2_lo = wh2_hi + kWTile / 2;
  float* s_const = reinterpret_cast<float*>(base + kOffConst);
  float* s_b2 = s_const; float* s_bh1 = s_const + 64; float* s_bh2 = s_const + 128; float* s_wh3 = s_const + 160;
  int* s_gstart = reinterpret_cast<int*>(s_const + 192);     // [T + 1] first slot of each type group
  int* s_gpstart = s_gstart + 20;                            // [T + 1] first part row of each type group
  uint8_t* wgb = base + kOffWg + tg * kWg2Bytes;
  const uint32_t a_hi = smem_u32(wgb) + kWgA, a_lo = a_hi + kATile;
  const uint32_t add_a = smem_u32(wgb) + kWgAdd;
  const uint32_t wm_hi = smem_u32(wgb) + kWgWm, wm_lo = wm_hi + kWTile;
  int2* s_cw = reinterpret_cast<int2*>(wgb + kWg2Misc);       // per row: .x = part row stored by the last row of a run, else -1;
  float* s_wa = reinterpret_cast<float*>(s_cw + kTile);       //          .y = the row's softmax weight (float bits)
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_wa + kD);   // MMA completions
  uint64_t* g_bar = bar + 1;                                  // bulk load of the edge-feature image
  uint64_t* c_bar = bar + 2;                                  // bulk load of the C image
  float* s_wfirst = reinterpret_cast<float*>(bar + 4);        // per warp: max logit of its first / last run segment,
  float* s_wlast = s_wfirst + 4;                              // flags: bit 0 = lane 0 continues the previous warp's run,
  int* s_wflag = reinterpret_cast<int*>(s_wlast + 4);         //        bit 1 = the whole warp is one segment
  int* s_seg = s_wflag + 4;                                   // [32] first row of the k-th row segment of the run reduction
  float* s_att0 = reinterpret_cast<float*>(s_seg + 32);       // the two column halves' shares of the attention logit
  float* s_att1 = s_att0 + kTile;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + kOffWg + 2 * kWg2Bytes);

  if (warp == 0) tmem_alloc<kEdgeTmemCols>(tmem_slot);
  if (tt == 0) { mbar_init(bar, 1); mbar_init(g_bar, 1); mbar_init(c_bar, 1); }
  fence_barrier_init();
  load_weight_tile_a(w1_hi, a.w1, kD, kD, tid, kEdge2Threads);
  load_weight_tile_a(w1_lo, a.w1 + kD * kD, kD, kD, tid, kEdge2Threads);
  load_weight_tile_a(w2_hi, a.w2, kD, kD, tid, kEdge2Threads);
  load_weight_tile_a(w2_lo, a.w2 + kD * kD, kD, kD, tid, kEdge2Threads);
  if (a.with_head) {
    load_weight_tile_a(wh1_hi, a.wh1, kD, kD, tid, kEdge2Threads);
    load_weight_tile_a(wh1_lo, a.wh1 + kD * kD, kD, kD, tid, kEdge2Threads);
    load_weight_tile_a(wh2_hi, a.wh2, 32, kD, tid, kEdge2Threads);
    load_weight_tile_a(wh2_lo, a.wh2 + 32 * kD, 32, kD, tid, kEdge2Threads);
    if (tid < kD) s_bh1[tid] = a.bh1[tid];
    if (tid < 32) { s_bh2[tid] = a.bh2[tid]; s_wh3[tid] = a.wh3[tid]; }
  }
  if (tid < kD) s_b2[tid] = a.b2[tid];
  if (tid <= a.T) { s_gstart[tid] = a.group_start[tid]; s_gpstart[tid] = a.group_pstart[tid]; }
  if (tt < kD) s_wa[tt] = 0.f;
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot + (uint32_t)(tg * 128);   // this tile group's 128 columns
  const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // ... at this warp's lanes
  const int bar_id = 1 + tg, scan_bar_id = 3 + tg;
  uint32_t phase = 0, ld_phase = 0;
  int cur_tm = -1, cur_col = -1;
  float att_bias = 0.f;
  bool g_issued = false;        // the tile's edge features were requested during the previous tile
  const bool copier_warp = (tt >> 5) == 1;   // one elected lane of this warp owns the bulk copies (loads of g / C, write-back of g')
  uint8_t* __restrict__ g_img = reinterpret_cast<uint8_t*>(a.g);

  const int total_tiles = a.group_start[a.T] >> 7;
#ifdef PGMP_TIMELINE
  const int units = a.one_group ? gridDim.x : 2 * gridDim.x;      // development aid: one tile group per SM idle
  const int per_unit = (total_tiles + units - 1) / units;
  const int tile_begin = a.one_group ? (tg ? total_tiles : blockIdx.x * per_unit) : (blockIdx.x * 2 + tg) * per_unit;
#else
  const int units = 2 * gridDim.x;
  const int per_unit = (total_tiles + units - 1) / units;
  const int tile_begin = (blockIdx.x * 2 + tg) * per_unit;
#endif
  const int tile_end = min(tile_begin + per_unit, total_tiles);

  int n_src = -1, n_dst = -1, n_seg = 0;     // per-row indices, prefetched one tile ahead (both halves of a row hold them)
  int2 n_run = make_int2(-1, 0);
  if (tile_begin < tile_end) {
    const int64_t sl = (int64_t)tile_begin * kTile + row;
    n_src = a.slot_src[sl]; n_dst = a.slot_dst[sl]; n_run = a.slot_run[sl];
    if (tt < 32) n_seg = a.tile_seg[tile_begin * 32 + tt];
  }
  int t = 0;                    // source type of the tile: monotone over the tiles of a tile group
#ifdef PGMP_TIMELINE
  const bool tl_on = tt == 0 && !a.with_head && (blockIdx.x == 3 || blockIdx.x == 100);
  const int tl_slot = (blockIdx.x == 100 ? 2 : 0) + tg;
#endif
  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const int slot0 = tile * kTile;
#ifdef PGMP_TIMELINE
    const int tl_i = tile - tile_begin;
#endif
    TL(0);
    while (t + 1 < a.T && slot0 >= s_gstart[t + 1]) ++t;
    const int tm = a.per_type ? t : 0;
    const int col = a.attn == PGMP_ATTN_PER_TYPE ? t : 0;
    if (copier_warp && elect_one()) {
      if (!g_issued) {
        mbar_expect_tx(g_bar, 2 * kATile);
        bulk_load(a_hi, g_img + (size_t)tile * (2 * kATile), 2 * kATile, g_bar);
      }
      if (a.c0) {
        mbar_expect_tx(c_bar, kAddTile);
        bulk_load(add_a, a.c0 + (size_t)slot0 * kD, kAddTile, c_bar);
        if (tile + 1 < tile_end) bulk_prefetch_l2(a.c0 + (size_t)(slot0 + kTile) * kD, kAddTile);
      }
      if (tile + 1 < tile_end) bulk_prefetch_l2(g_img + (size_t)(tile + 1) * (2 * kATile), 2 * kATile);
    }
    const int src = n_src, dst = n_dst, e = dst;       // pad rows have src = dst = -1
    const int2 run = n_run;
    if (tt < 32) s_seg[tt] = n_seg;
    if (tile + 1 < tile_end) {   // next tile's indices
      const int64_t sl = (int64_t)slot0 + kTile + row;
      n_src = a.slot_src[sl]; n_dst = a.slot_dst[sl]; n_run = a.slot_run[sl];
      if (tt < 32) n_seg = a.tile_seg[(tile + 1) * 32 + tt];
    }
    if (tm != cur_tm) {     // all MMAs of the previous tile have completed
      load_weight_tile_a(wm_hi, a.wm + (size_t)tm * 2 * kD * kD, kD, kD, tt, kTgThreads);
      load_weight_tile_a(wm_lo, a.wm + (size_t)tm * 2 * kD * kD + kD * kD, kD, kD, tt, kTgThreads);
      fence_async_smem();
      cur_tm = tm;
    }
    if (a.attn && col != cur_col) {
      if (tt < kD) s_wa[tt] = a.wa[tt * a.attn_cols + col];
      att_bias = __ldg(a.ba + col);
      cur_col = col;
    }
    // ---- P[dst] and Q[src]: this thread's half of its two table rows (swizzled tile images: the logical 32-byte pair m
    //      of node n sits at pair position m ^ (n & 7), umma::table_index), four 256-bit loads per half row, requested
    //      before the first product is waited for, summed on arrival.
    //      Pad rows (e < 0) read row 0: finite values that nothing consumes.
    //      (Measured alternatives, per 10 steps: rows gathered cooperatively -- 16 threads per 256-byte row through the
    //      staging tile -- 3.28 ms instead of 2.26 ms: the L1 requests drop 8x but the latency lands on the critical
    //      path; the NEXT tile's rows requested before the run reduction: 2.44 ms (Q only) / 3.38 ms (P and Q, spills)
    //      instead of 2.14 ms: the requests queue in front of the reduction's shared-memory reads.)
    float4 pq[8], qv[8];
    {
      // (rows are 256-byte aligned: the swizzled pair position (4 half + q) ^ x is one XOR on the low address bits)
      const int xd = e >= 0 ? dst & 7 : 0, xs = e >= 0 ? src & 7 : 0;
      const uintptr_t p0 = reinterpret_cast<uintptr_t>(a.tab_p + (size_t)(e >= 0 ? dst : 0) * kD) + (uintptr_t)(((4 * half) ^ xd) << 5);
      const uintptr_t q0 = reinterpret_cast<uintptr_t>(a.tab_q + (size_t)(e >= 0 ? src : 0) * kD) + (uintptr_t)(((4 * half) ^ xs) << 5);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float8 pv = ldg256(reinterpret_cast<const float*>(p0 ^ (uintptr_t)(q << 5)));
        const float8 qq = ldg256(reinterpret_cast<const float*>(q0 ^ (uintptr_t)(q << 5)));
        pq[2 * q] = pv.a; pq[2 * q + 1] = pv.b;
        qv[2 * q] = qq.a; qv[2 * q + 1] = qq.b;
      }
    }
    TL(1);
    if (tt < 32 && elect_one()) {        // the edge features have landed: first product
      mbar_wait(g_bar, ld_phase);
      fence_after_sync();
      issue_gemm_x3<kD>(tmem, a_hi, a_lo, 0, w1_hi, w1_lo, 0, 1, false);
      mma_commit(bar);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 lo2 = add2(make_float2(pq[q].x, pq[q].y), make_float2(qv[q].x, qv[q].y));
      const float2 hi2 = add2(make_float2(pq[q].z, pq[q].w), make_float2(qv[q].z, qv[q].w));
      pq[q] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
    }
    TL(2);
    if (a.c0) mbar_wait(c_bar, ld_phase);          // the C image sits in the staging tile
    ld_phase ^= 1;
    TL(3);
    mbar_wait(bar, phase);
    phase ^= 1;
    fence_after_sync();
    TL(4);
    float d[32];
    tmem_ld32(tmem_row + (uint32_t)c0col, d);
    {   // hidden = ReLU(acc + C + P + Q) -> tensor memory (TS-form A operand: element k of a row in half k & 1 of column
        // k / 2): this thread's 32 elements are 16 columns of the hi block (64 ..) and 16 of the lo block (96 ..).
        // Packed adds; the ReLU is part of the bf16 split.
      uint32_t h[16], l[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float2 v0 = make_float2(pq[q].x, pq[q].y), v1 = make_float2(pq[q].z, pq[q].w);
        if (a.c0) {
          const float4 c = lds128f(add_a + 4 * stage_index(row, c0col + 4 * q));
          v0 = add2(v0, make_float2(c.x, c.y));
          v1 = add2(v1, make_float2(c.z, c.w));
        }
        split2_relu(add2(make_float2(d[4 * q + 0], d[4 * q + 1]), v0), h[2 * q], l[2 * q]);
        split2_relu(add2(make_float2(d[4 * q + 2], d[4 * q + 3]), v1), h[2 * q + 1], l[2 * q + 1]);
      }
      tmem_st16(tmem_row + 64 + 16 * half, h);
      tmem_st16(tmem_row + 96 + 16 * half, l);
      tmem_st_wait();
    }
    TL(5);
    fence_before_sync();
    named_bar_sync(bar_id, kTgThreads);
    TL(6);
    if (tt >= kTgThreads - 32 && elect_one()) {      // issued by the group's last warp: the others go on to their R rows
      fence_after_sync();
      issue_gemm_x3_ts<kD>(tmem, tmem + 64, tmem + 96, w2_hi, w2_lo, false);
      mma_commit(bar);
    }
    // ---- R[type][dst]: this thread's half row, requested now, consumed by the third epilogue
    float4 rv[8];
    {
      const int xd = e >= 0 ? dst & 7 : 0;
      const uintptr_t r0 = reinterpret_cast<uintptr_t>(a.tab_r + ((size_t)t * a.N + (e >= 0 ? dst : 0)) * kD) + (uintptr_t)(((4 * half) ^ xd) << 5);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float8 r8 = ldg256(reinterpret_cast<const float*>(r0 ^ (uintptr_t)(q << 5)));
        rv[2 * q] = r8.a; rv[2 * q + 1] = r8.b;
      }
    }
    TL(7);
    mbar_wait(bar, phase);
    phase ^= 1;
    fence_after_sync();
    TL(8);
    tmem_ld32(tmem_row + (uint32_t)c0col, d);
    {   // g' = ReLU(acc + b2): attention logit share (two packed chains) and the operand tile (4 of the row's 8
        // sixteen-byte chunks per thread)
      float2 at0 = make_float2(0.f, 0.f), at1 = at0;
      const uint32_t b2_a = smem_u32(s_b2) + 4u * (uint32_t)c0col, wa_a = smem_u32(s_wa) + 4u * (uint32_t)c0col;
      const uint32_t off0 = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128) + (((uint32_t)(4 * half) ^ (uint32_t)(row & 7)) << 4);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 b = lds128f(b2_a + 32 * c + 16 * i), wv = lds128f(wa_a + 32 * c + 16 * i);
          float2 v0 = add2(make_float2(d[8 * c + 4 * i + 0], d[8 * c + 4 * i + 1]), make_float2(b.x, b.y));
          float2 v1 = add2(make_float2(d[8 * c + 4 * i + 2], d[8 * c + 4 * i + 3]), make_float2(b.z, b.w));
          v0.x = fmaxf(v0.x, 0.f); v0.y = fmaxf(v0.y, 0.f); v1.x = fmaxf(v1.x, 0.f); v1.y = fmaxf(v1.y, 0.f);
          at0 = fma2(v0, make_float2(wv.x, wv.y), at0);
          at1 = fma2(v1, make_float2(wv.z, wv.w), at1);
          split2_pos(v0, h[2 * i], l[2 * i]);
          split2_pos(v1, h[2 * i + 1], l[2 * i + 1]);
        }
        const uint32_t off = off0 ^ (uint32_t)(c << 4);
        sts128(a_hi + off, h[0], h[1], h[2], h[3]);
        sts128(a_lo + off, l[0], l[1], l[2], l[3]);
      }
      (half ? s_att1 : s_att0)[row] = (at0.x + at0.y) + (at1.x + at1.y);
    }
    TL(9);
    fence_before_sync();
    fence_async_smem();
    named_bar_sync(bar_id, kTgThreads);
    TL(10);
    if (tt >= kTile && tt < kTile + 32 && elect_one()) {   // issued by a warp of column half 1: half 0 runs the scan meanwhile
      fence_after_sync();
      issue_gemm_x3<kD>(tmem, a_hi, a_lo, 0, wm_hi, wm_lo, 0, 1, false);
      if (a.with_head) issue_gemm_x3<kD>(tmem + 64, a_hi, a_lo, 0, wh1_hi, wh1_lo, 0, 1, false);
      mma_commit(bar);
    }
    // ---- write back g' as the bf16 hi/lo tile image (what the next step's MMA consumes): one bulk copy
    TL(16);
    if (copier_warp && elect_one()) bulk_store(g_img + (size_t)tile * (2 * kATile), a_hi, 2 * kATile);
    // ---- runs of equal targets (a bin, or the part of it inside this tile), handled by the half-0 thread of every row:
    //      maximum logit of every run -- first within the warp (segmented shuffle scan), then across the four warps
    //      through shared memory -- the softmax weight of every row, and for the last row of a run its part row
    if (half == 0) {
      const float att = att_bias + (s_att0[row] + s_att1[row]);
      const int lane = row & 31, wq = row >> 5;
      float run_max = att, wgt;
      if (a.attn) {
        const int seg_first = run.y & 31, seg_last = (run.y >> 5) & 31;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
          const float o = __shfl_down_sync(0xffffffffu, run_max, dd);
          if (lane + dd <= seg_last) run_max = fmaxf(run_max, o);
        }
        run_max = __shfl_sync(0xffffffffu, run_max, seg_first);
        if (lane == 0) { s_wfirst[wq] = run_max; s_wflag[wq] = ((run.y >> 10) & 1 ? 0 : 1) | (seg_last == 31 ? 2 : 0); }
        if (lane == 31) s_wlast[wq] = run_max;
        TL(17);
        named_bar_sync(scan_bar_id, kTile);
        TL(18);
        if (seg_last == 31)        // the run may continue in the following warps
          for (int k2 = wq + 1; k2 < 4 && (s_wflag[k2] & 1); ++k2) {
            run_max = fmaxf(run_max, s_wfirst[k2]);
            if (!(s_wflag[k2] & 2)) break;
          }
        if (seg_first == 0 && (s_wflag[wq] & 1))   // ... and may have begun in the preceding ones
          for (int k2 = wq - 1; k2 >= 0; --k2) {
            run_max = fmaxf(run_max, s_wlast[k2]);
            if ((s_wflag[k2] & 3) != 3) break;
          }
        wgt = e >= 0 ? __expf(att - run_max) : 0.f;
        if (run.x >= 0) a.part_mx[run.x] = run_max;
      } else {
        wgt = e >= 0 ? 1.f : 0.f;
      }
      TL(19);
      s_cw[row] = make_int2(run.x, __float_as_int(wgt));
      TL(20);
    }
    TL(11);
    mbar_wait(bar, phase);
    phase ^= 1;
    fence_after_sync();
    TL(12);
    // the operand tiles are free (no head): fetch the next tile's edge features behind the third epilogue and the reduction
    g_issued = false;
    if (!a.with_head && tile + 1 < tile_end) {
      if (copier_warp && elect_one()) {
        bulk_wait_read();
        mbar_expect_tx(g_bar, 2 * kATile);
        bulk_load(a_hi, g_img + (size_t)(tile + 1) * (2 * kATile), 2 * kATile, g_bar);
      }
      g_issued = true;
    }
    tmem_ld32(tmem_row + (uint32_t)c0col, d);
    // ---- message m = ReLU(d + R) goes to this thread's half of its staging row (the C row was consumed by the first epilogue)
    const uint32_t m_a0 = add_a + 4 * stage_index(row, c0col);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float2 m0 = add2(make_float2(d[4 * q + 0], d[4 * q + 1]), make_float2(rv[q].x, rv[q].y));
      const float2 m1 = add2(make_float2(d[4 * q + 2], d[4 * q + 3]), make_float2(rv[q].z, rv[q].w));
      sts128f(m_a0 ^ (uint32_t)(q << 4),
              make_float4(fmaxf(m0.x, 0.f), fmaxf(m0.y, 0.f), fmaxf(m1.x, 0.f), fmaxf(m1.y, 0.f)));
    }
    if (a.with_head) {   // head layer 1 epilogue -> A, layer 2 on the tensor cores
      tmem_ld32(tmem_row + 64 + (uint32_t)c0col, d);
#pragma unroll
      for (int o = 0; o < 32; ++o) d[o] = fmaxf(d[o] + s_bh1[c0col + o], 0.f);
      if (copier_warp && elect_one()) bulk_wait_read();           // the write-back has finished reading the operand tiles
      named_bar_sync(bar_id, kTgThreads);
      {
        const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        const uint32_t x = (uint32_t)(row & 7);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) split2(d[8 * c + 2 * i], d[8 * c + 2 * i + 1], h[i], l[i]);
          const uint32_t off = row_off + (((uint32_t)(4 * half + c) ^ x) << 4);
          sts128(a_hi + off, h[0], h[1], h[2], h[3]);
          sts128(a_lo + off, l[0], l[1], l[2], l[3]);
        }
      }
      fence_before_sync();
      fence_async_smem();
      named_bar_sync(bar_id, kTgThreads);
      if (tt < 32 && elect_one()) {
        fence_after_sync();
        issue_gemm_x3<32>(tmem + 64, a_hi, a_lo, 0, wh2_hi, wh2_lo, 0, 1, false);
        mma_commit(bar);
      }
    } else {
      TL(13);
      named_bar_sync(bar_id, kTgThreads);
      TL(14);
    }
    TL(21);
    // ---- reduce every run: thread = (two column quads c8 and c8 + 8, one of 32 row segments); the segments meet at run
    //      boundaries (tile_seg), so every run is reduced by one thread per column group, rows in order (deterministic),
    //      results stored as 256-byte part rows.
    {
      const int seg = tt >> 3;
      const uint32_t c8x = (uint32_t)(tt & 7) << 4;
      const int r_begin = s_seg[seg], r_end = seg == 31 ? kTile : s_seg[seg + 1];
      const uint32_t cw_a = smem_u32(s_cw);
      float* const out = a.part_val + 4 * (tt & 7);
      if (a.aggr == PGMP_AGGR_MAX && !a.attn) {
        float4 u = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), v = u;
#pragma unroll 2
        for (int r = r_begin; r < r_end; ++r) {
          const uint2 cw = lds64(cw_a + 8u * (uint32_t)r);
          const uint32_t ra = add_a + (uint32_t)r * 256u, x = ((uint32_t)r << 4) & 0xf0u;
          const float4 m0 = lds128f(ra + (c8x ^ x)), m1 = lds128f(ra + ((c8x | 0x80u) ^ x));
          if (cw.y != 0u) {     // weight 1 (valid row) or 0 (pad row)
            u.x = fmaxf(u.x, m0.x); u.y = fmaxf(u.y, m0.y); u.z = fmaxf(u.z, m0.z); u.w = fmaxf(u.w, m0.w);
            v.x = fmaxf(v.x, m1.x); v.y = fmaxf(v.y, m1.y); v.z = fmaxf(v.z, m1.z); v.w = fmaxf(v.w, m1.w);
          }
          if ((int)cw.x >= 0) {
            float* o = out + (size_t)cw.x * kD;
            *reinterpret_cast<float4*>(o) = u;
            *reinterpret_cast<float4*>(o + 32) = v;
            u = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY); v = u;
          }
        }
      } else {
        float se = 0.f;
        float2 u0 = make_float2(0.f, 0.f), u1 = u0, v0 = u0, v1 = u0;
#pragma unroll 2
        for (int r = r_begin; r < r_end; ++r) {
          const uint2 cw = lds64(cw_a + 8u * (uint32_t)r);
          const uint32_t ra = add_a + (uint32_t)r * 256u, x = ((uint32_t)r << 4) & 0xf0u;
          const float4 m0 = lds128f(ra + (c8x ^ x)), m1 = lds128f(ra + ((c8x | 0x80u) ^ x));
          const float wv = __uint_as_float(cw.y);
          const float2 w2 = make_float2(wv, wv);
          se += wv;
          u0 = fma2(w2, make_float2(m0.x, m0.y), u0);
          u1 = fma2(w2, make_float2(m0.z, m0.w), u1);
          v0 = fma2(w2, make_float2(m1.x, m1.y), v0);
          v1 = fma2(w2, make_float2(m1.z, m1.w), v1);
          if ((int)cw.x >= 0) {
            float* o = out + (size_t)cw.x * kD;
            *reinterpret_cast<float4*>(o) = make_float4(u0.x, u0.y, u1.x, u1.y);
            *reinterpret_cast<float4*>(o + 32) = make_float4(v0.x, v0.y, v1.x, v1.y);
            if (a.attn && (tt & 7) == 0) a.part_se[cw.x] = se;
            u0 = make_float2(0.f, 0.f); u1 = u0; v0 = u0; v1 = u0;
            se = 0.f;
          }
        }
      }
    }
    if (a.with_head) {   // head layer 2 epilogue and the final 32 -> 1 dot product (the half-0 thread of every row)
      mbar_wait(bar, phase);
      phase ^= 1;
      fence_after_sync();
      if (half == 0) {
        float hv[32];
        tmem_ld32(tmem_row + 64, hv);
        float logit = __ldg(a.bh3);
#pragma unroll
        for (int o = 0; o < 32; ++o) logit = fmaf(fmaxf(hv[o] + s_bh2[o], 0.f), s_wh3[o], logit);
        if (e >= 0) a.edge_logits[a.slot_edge[slot0 + row]] = logit;
      }
    }
    TL(15);
    fence_before_sync();
    named_bar_sync(bar_id, kTgThreads);   // the next tile overwrites the staging / operand tiles
  }
  if (copier_warp && elect_one()) bulk_wait_all();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<kEdgeTmemCols>(*tmem_slot);
}

// fp32 row-major edge features [S][64] -> per-tile bf16 hi/lo operand images (in place, one CTA per tile)
__global__ void __launch_bounds__(kWgThreads) g_to_image_kernel(float* __restrict__ g, const int32_t* __restrict__ group_start, int T) {
  const int tile = blockIdx.x;
  if ((int64_t)tile * kTile >= group_start[T]) return;
  float4 v[16];
  const float4* __restrict__ src = reinterpret_cast<const float4*>(g + (size_t)tile * kTile * kD);
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = src[threadIdx.x + k * kWgThreads];
  __syncthreads();   // the whole tile is in registers before any byte of it is overwritten
  uint8_t* __restrict__ img = reinterpret_cast<uint8_t*>(g) + (size_t)tile * (2 * kATile);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int idx = threadIdx.x + k * kWgThreads;
    store_split4(img, img + kATile, idx >> 4, idx & 15, v[k]);
  }
}

// fp32 row-major C rows [S][64] -> per-tile swizzled staging images (in place, one CTA per tile); only after the
// SIMT embedding fallback, the tensor-core embedding writes the images directly
__global__ void __launch_bounds__(kWgThreads) c_to_image_kernel(float* __restrict__ c0, const int32_t* __restrict__ group_start, int T) {
  const int tile = blockIdx.x;
  if ((int64_t)tile * kTile >= group_start[T]) return;
  float4 v[16];
  float4* __restrict__ t4 = reinterpret_cast<float4*>(c0 + (size_t)tile * kTile * kD);
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = t4[threadIdx.x + k * kWgThreads];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int idx = threadIdx.x + k * kWgThreads;
    t4[stage_index(idx >> 4, (idx & 15) * 4) >> 2] = v[k];
  }
}

// ---- selftest ------------------------------------------------------------------------------------------
struct TcSmem {
  uint8_t* a_hi; uint8_t* a_lo; uint8_t* w1_hi; uint8_t* w1_lo;
  uint64_t* bar; uint32_t* tmem;
};
constexpr size_t kTcSmemBytes = 2 * kATile + 2 * kWTile + 64 + 1024;
__device__ __forceinline__ TcSmem carve_smem(uint8_t* raw) {
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  TcSmem s;
  s.a_hi = base; s.a_lo = base + kATile;
  s.w1_hi = base + 2 * kATile; s.w1_lo = s.w1_hi + kWTile;
  s.bar = reinterpret_cast<uint64_t*>(s.w1_lo + kWTile);
  s.tmem = reinterpret_cast<uint32_t*>(s.bar + 1);
  return s;
}
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

// D = A . W^T with the A operand in tensor memory (pgmp_selftest_umma_ts): the threads split their row of A into
// bf16 hi / lo pairs and store them with tcgen05.st (hi in columns [64, 96), lo in [96, 128)); B stays in shared memory
__global__ void __launch_bounds__(kTile) selftest_umma_ts_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                                  float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  const TcSmem s = carve_smem(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc<128>(s.tmem);
  if (tid == 0) {
    mbar_init(s.bar, 1);
    fence_barrier_init();
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
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) split2(A[tid * kD + 2 * k], A[tid * kD + 2 * k + 1], hi[k], lo[k]);
  tmem_st16(tmem + lane_base + 64, hi);
  tmem_st16(tmem + lane_base + 80, hi + 16);
  tmem_st16(tmem + lane_base + 96, lo);
  tmem_st16(tmem + lane_base + 112, lo + 16);
  tmem_st_wait();
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  if (tid < 32 && elect_one()) {
    fence_after_sync();
    constexpr uint32_t idesc = idesc_bf16(128, kD);
    const uint64_t wh = smem_desc_sw128(smem_u32(s.w1_hi)), wl = smem_desc_sw128(smem_u32(s.w1_lo));
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {        // K = 16 per instruction = 8 columns of the A operand
      mma_bf16_ts(tmem, tmem + 64 + 8 * k, wh + 2 * k, idesc, acc);
      acc = 1;
      mma_bf16_ts(tmem, tmem + 64 + 8 * k, wl + 2 * k, idesc, 1u);
      mma_bf16_ts(tmem, tmem + 96 + 8 * k, wh + 2 * k, idesc, 1u);
    }
    mma_commit(s.bar);
  }
  mbar_wait(s.bar, 0);
  fence_after_sync();
  float d[kD];
  tmem_ld64(tmem, 0, d);
  for (int o = 0; o < kD; ++o) D[tid * kD + o] = d[o];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

}  // namespace

int mpn_forward_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st) {
  if (!p.tc_w1_e || !p.tc_w2 || !p.tc_wm_e || !p.tc_wtab || (p.has_update_mlp && !p.update_hier && !p.tc_wu))
    return set_error(PGMP_ERR_INVALID, "PGMP_PRECISION_TC needs the bf16 hi/lo weights");
  // the node update runs on the tensor cores when it is a matrix product (update MLP), else it is a plain merge
  auto node_update = [&](int out_slot) {
    if (p.update_hier) {            // hierarch_mlp update: fp32 SIMT kernel, then the operand image of h
      const int r = mpn_node_update_hier(p, w, out_slot, st);
      return r != PGMP_OK ? r : mpn_node_image(w, w.h, p.num_nodes, w.h_img, st);
    }
    if (p.has_update_mlp) return mpn_node_update_tc(p, w, out_slot, st);
    const int r = mpn_node_update(p, w, out_slot, st);          // plain merge (agnostic layer without update MLP)
    return r != PGMP_OK ? r : mpn_node_image(w, w.h, p.num_nodes, w.h_img, st);
  };
  const int64_t N = p.num_nodes, E = p.num_edges;
  int rc;
  bool emb_tc = false, nemb_tc = false;
  if (E > 0 && (rc = mpn_edge_embed_tc(p, w, st, &emb_tc)) != PGMP_OK) return rc;
  if (emb_tc) {
    if ((rc = mpn_node_embed_tc(p, w, st, &nemb_tc)) != PGMP_OK) return rc;      // writes h0 and its operand image
    if (!nemb_tc && (rc = mpn_embed_nodes(p, w, st)) != PGMP_OK) return rc;
  } else {          // unusual embedding shapes: SIMT chain, then convert the edge features to operand images
    if ((rc = mpn_embed(p, w, st)) != PGMP_OK) return rc;
    if (E > 0) PGMP_LAUNCH(g_to_image_kernel, (unsigned)(w.max_slots / kTile), kWgThreads, 0, st, w.g, w.group_start, p.num_types);
    if (E > 0 && p.skip) PGMP_LAUNCH(c_to_image_kernel, (unsigned)(w.max_slots / kTile), kWgThreads, 0, st, w.c0, w.group_start, p.num_types);
  }
  if (!nemb_tc && (rc = mpn_node_image(w, w.h0, N, w.h0_img, st)) != PGMP_OK) return rc;
  PGMP_CUDA(cudaFuncSetAttribute(edge_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEdge2SmemBytes));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  EdgeTcArgs a;
  a.slot_edge = w.slot_edge; a.slot_src = w.slot_src; a.slot_dst = w.slot_dst;
  a.slot_run = w.slot_run; a.tile_seg = w.tile_seg;
  a.group_start = w.group_start; a.group_pstart = w.group_pstart;
  a.g = w.g; a.c0 = p.skip ? w.c0 : nullptr; a.tab_p = w.tab_p; a.tab_q = w.tab_q; a.tab_r = w.tab_r;
  a.w1 = static_cast<const __nv_bfloat16*>(p.tc_w1_e); a.w2 = static_cast<const __nv_bfloat16*>(p.tc_w2);
  a.wm = static_cast<const __nv_bfloat16*>(p.tc_wm_e);
  a.b2 = p.b2; a.wa = p.wa; a.ba = p.ba;
  a.part_val = w.part_val; a.part_mx = w.part_mx; a.part_se = w.part_se;
  a.N = N; a.T = p.num_types; a.per_type = p.per_type; a.aggr = p.aggr; a.attn = p.attn;
  a.attn_cols = p.attn == PGMP_ATTN_PER_TYPE ? 17 : 1;
  // fused tensor-core edge head when the head is the reference's 64 -> 64 -> 32 -> 1 chain (else a SIMT pass)
  const pgmp_mlp& eh = p.edge_head;
  const bool fused_head = p.tc_wh1 && p.tc_wh2 && eh.n_layers == 3 && eh.dims[0] == 64 && eh.dims[1] == 64 &&
                          eh.dims[2] == 32 && eh.dims[3] == 1 && eh.relu[0] && eh.relu[1] && !eh.relu[2] &&
                          !eh.post_relu && !eh.post_scale;
  a.wh1 = static_cast<const __nv_bfloat16*>(p.tc_wh1); a.wh2 = static_cast<const __nv_bfloat16*>(p.tc_wh2);
  a.bh1 = eh.bias[0]; a.bh2 = eh.bias[1]; a.wh3 = eh.wt[2]; a.bh3 = eh.bias[2];
  const unsigned max_units = (unsigned)ceil_div<uint64_t>(w.max_slots / kTile, 2);
  unsigned grid = max_units < (unsigned)sms ? max_units : (unsigned)sms;
#ifdef PGMP_TIMELINE
  if (const char* gs = getenv("PGMP_STEP_GRID")) grid = (unsigned)atoi(gs);      // development aid: scaling experiments
  a.one_group = getenv("PGMP_STEP_ONE_GROUP") != nullptr;
#else
  a.one_group = 0;
#endif
  const int first_out = p.steps - p.aux_loss_steps - 1 > 0 ? p.steps - p.aux_loss_steps - 1 : 0;
  for (int s = 0; s < p.steps; ++s) {
    if (s > 0) {
      const int prev_slot = (s - 1) >= first_out ? (s - 1) - first_out : -1;
      if ((rc = node_update(prev_slot)) != PGMP_OK) return rc;
    }
    if ((rc = mpn_node_tables_tc(p, w, s == 0 ? w.h0_img : w.h_img, st)) != PGMP_OK) return rc;
    if (E > 0) {
      const bool out = s >= first_out;
      a.with_head = out && fused_head;
      a.edge_logits = out ? p.edge_logits + (size_t)(s - first_out) * E : nullptr;
      PGMP_LAUNCH(edge_step_tc_kernel, grid, kEdge2Threads, kEdge2SmemBytes, st, a);
      if (out && !fused_head && (rc = mpn_edge_head(p, w, s - first_out, true, st)) != PGMP_OK) return rc;
    }
  }
  return node_update((p.steps - 1) - first_out);
}

}  // namespace pgmp

#ifdef PGMP_TIMELINE
extern "C" int pgmp_debug_step_timeline(long long* out) {
  return cudaMemcpyFromSymbol(out, pgmp::g_step_tl, sizeof(long long) * 4 * 32 * 32) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int pgmp_selftest_umma(const float* a, const float* w, float* d, pgmp_stream_t stream) {
  using namespace pgmp;
  if (!a || !w || !d) return set_error(PGMP_ERR_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGMP_CUDA(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  PGMP_LAUNCH(selftest_umma_kernel, 1, kTile, kTcSmemBytes, st, a, w, d);
  return PGMP_OK;
}

extern "C" int pgmp_selftest_umma_ts(const float* a, const float* w, float* d, pgmp_stream_t stream) {
  using namespace pgmp;
  if (!a || !w || !d) return set_error(PGMP_ERR_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGMP_CUDA(cudaFuncSetAttribute(selftest_umma_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes));
  PGMP_LAUNCH(selftest_umma_ts_kernel, 1, kTile, kTcSmemBytes, st, a, w, d);
  return PGMP_OK;
}
