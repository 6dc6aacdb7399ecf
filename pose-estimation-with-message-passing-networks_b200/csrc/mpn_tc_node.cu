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

int mpn_node_heads_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, float* node_logits, float* class_logits,
                      cudaStream_t st, bool* done);

namespace {

using namespace umma;

constexpr int kImage = kTile * kD * 4;   // bytes of one node-tile image: hi tile (16 KB) then lo tile (16 KB)
constexpr int kHalf = kTile * 128;       // bytes of one [128][64] bf16 tile
constexpr int kWBlock = kD * 128;        // bytes of one [64][64] bf16 tile

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
// Per-node tables.  Work item = (node tile, output chunk); chunk 0 / 1 are the target / source columns of
// mlp_edge.0, chunk 2 + t the node columns of message MLP t.  Persistent CTAs (one per SM) walk a contiguous,
// tile-major range of items.  Warp 8 is the producer: it fetches the A operand ([h0 ; h] tile images) once per tile
// and the pre-swizzled weight images three items ahead (bulk copies), and issues the products into two alternating
// TMEM accumulators.  Warps 0-3 and 4-7 are two epilogue groups, one per accumulator / staging buffer (even and odd
// items): bias, fp32 rows staged as a swizzled tile image, one bulk copy per item into the table (the step kernel
// gathers 16-byte chunks with the same swizzle).  One group alone left the tensor pipe 26 % busy: with a single warp per
// scheduler every latency of the epilogue (bias loads, TMEM load, barrier) was exposed.
constexpr int kTabA = 2 * kImage;                 // [h0 image][h image]
constexpr int kTabW = 2 * 2 * kWBlock;            // one weight buffer: 2 K-blocks x (hi, lo)
constexpr int kTabWBufs = 3;
constexpr int kTabStage = kTile * kD * 4;
constexpr size_t kTabPayload = kTabA + kTabWBufs * kTabW + 2 * kTabStage;
constexpr size_t kTabSmem = kTabPayload + 128 + 1024;
constexpr int kTabThreads = 2 * kTile + 32;

__global__ void __launch_bounds__(kTabThreads, 1) node_tables_tc_kernel(
    const float* __restrict__ h0_img, const float* __restrict__ h_img, int64_t N, int skip, int per_type, int n_chunks,
    int total_items, const __nv_bfloat16* __restrict__ wtab, const float* __restrict__ b1,
    const float* __restrict__ bm, float* __restrict__ tab_p, float* __restrict__ tab_q, float* __restrict__ tab_r) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(base);
  const uint32_t a0 = sb;                               // K-block b: hi at a0 + b * kImage, lo at + kHalf
  const uint32_t wbuf = sb + kTabA;                     // buffer i at + i * kTabW; K-block b: hi at + b * 2 * kWBlock, lo + kWBlock
  const uint32_t stage = wbuf + kTabWBufs * kTabW;      // buffer i at + i * kTabStage
  uint64_t* a_bar = reinterpret_cast<uint64_t*>(base + kTabPayload);   // A operand landed
  uint64_t* w_full = a_bar + 1;                         // [3] weight image landed
  uint64_t* w_free = a_bar + 4;                         // [3] the product reading the buffer has completed
  uint64_t* d_full = a_bar + 7;                         // [2] accumulator complete
  uint64_t* d_free = a_bar + 9;                         // [2] accumulator drained by the epilogue (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 11);
  const int tid = threadIdx.x;
  if (tid < 32) tmem_alloc<128>(tmem_slot);
  if (tid == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(a_bar + i, 1);
    mbar_init(d_free, kTile);
    mbar_init(d_free + 1, kTile);
    fence_barrier_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int kb = skip ? 2 : 1;
  const uint32_t w_bytes = (uint32_t)kb * 2 * kWBlock;
  const int per = (total_items + gridDim.x - 1) / gridDim.x;
  const int j0 = blockIdx.x * per, j1 = min(j0 + per, total_items);

  if (tid >= 2 * kTile) {
   if (elect_one()) {
    // ---------------- producer (one elected lane of warp 8) ----------------
    auto issue_w = [&](int j) {                         // weight image of item j -> buffer (j - j0) % 3
      const int i = j - j0, buf = i % kTabWBufs;
      if (i >= kTabWBufs) mbar_wait(w_free + buf, (uint32_t)(i / kTabWBufs - 1) & 1u);
      const int c = j % n_chunks;
      const int wc = c < 2 ? c : 2 + (per_type ? c - 2 : 0);
      mbar_expect_tx(w_full + buf, w_bytes);
      bulk_load(wbuf + buf * kTabW, wtab + (size_t)wc * kb * 2 * kD * kD, w_bytes, w_full + buf);
    };
    int cur_tile = -1;
    uint32_t a_phase = 0;
    for (int j = j0; j < j1 && j < j0 + kTabWBufs - 1; ++j) issue_w(j);
    for (int j = j0; j < j1; ++j) {
      const int i = j - j0, wb = i % kTabWBufs, db = i & 1;
      const int tile = j / n_chunks;
      if (tile != cur_tile) {                           // [h0 ; h] (NodeClassificationMPNSimple.py:77) or [h]
        if (i > 0) mbar_wait(d_full + ((i - 1) & 1), (uint32_t)((i - 1) >> 1) & 1u);   // every product on the old tile is done
        mbar_expect_tx(a_bar, (uint32_t)kb * kImage);
        for (int b = 0; b < kb; ++b)
          bulk_load(a0 + b * kImage, reinterpret_cast<const uint8_t*>((skip && b == 0) ? h0_img : h_img) + (size_t)tile * kImage,
                    kImage, a_bar);
        mbar_wait(a_bar, a_phase);
        a_phase ^= 1;
        cur_tile = tile;
      }
      mbar_wait(w_full + wb, (uint32_t)(i / kTabWBufs) & 1u);
      if (i >= 2) mbar_wait(d_free + db, (uint32_t)((i >> 1) - 1) & 1u);
      fence_after_sync();
      issue_gemm_x3<kD>(tmem + db * 64, a0, a0 + kHalf, kImage, wbuf + wb * kTabW, wbuf + wb * kTabW + kWBlock, 2 * kWBlock, kb, false);
      mma_commit(d_full + db);
      mma_commit(w_free + wb);
      // the weight image two items ahead reuses the buffer of item j - 1: requested AFTER this item's product is queued,
      // so waiting for that buffer does not keep the tensor pipe idle
      if (j + kTabWBufs - 1 < j1) issue_w(j + kTabWBufs - 1);
    }
   }
  } else {
    // ---------------- epilogue: group eg takes the items with (j - j0) % 2 == eg ----------------
    const int eg = tid >> 7, row = tid & (kTile - 1);
    const bool store_lane = (row < 32) && elect_one();  // issues this group's bulk stores (always the same lane)
    const uint32_t st = stage + eg * kTabStage;
    for (int j = j0 + eg; j < j1; j += 2) {
      const int i = j - j0;
      const int c = j % n_chunks, tile = j / n_chunks;
      const float* __restrict__ bias = c == 0 ? (skip ? nullptr : b1) : (c == 1 ? nullptr : bm + (size_t)(per_type ? c - 2 : 0) * kD);
      float* __restrict__ dst = c == 0 ? tab_p : (c == 1 ? tab_q : tab_r + (size_t)(c - 2) * N * kD);
      float4 bv[kD / 4];
#pragma unroll
      for (int q = 0; q < kD / 4; ++q) bv[q] = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (store_lane) bulk_wait_read();                 // this group's previous store has finished reading the staging buffer
      mbar_wait(d_full + eg, (uint32_t)(i >> 1) & 1u);
      fence_after_sync();
      float d[kD];
      tmem_ld64(tmem + eg * 64, 0, d);
      fence_before_sync();
      mbar_arrive(d_free + eg);                         // the accumulator may be overwritten
      named_bar_sync(1 + eg, kTile);                    // ... and the staging buffer is free
#pragma unroll
      for (int q = 0; q < kD / 4; ++q)
        sts128f(st + 4 * table_index(row, 4 * q),
                make_float4(d[4 * q] + bv[q].x, d[4 * q + 1] + bv[q].y, d[4 * q + 2] + bv[q].z, d[4 * q + 3] + bv[q].w));
      fence_async_smem();
      named_bar_sync(1 + eg, kTile);
      if (store_lane) {
        const int64_t row0 = (int64_t)tile * kTile;
        const int64_t rows = N - row0 < kTile ? N - row0 : kTile;
        bulk_store(dst + row0 * kD, st, (uint32_t)rows * kD * 4);
      }
    }
    if (store_lane) bulk_wait_all();
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<128>(tmem);
}

// ------------------------------------------------------------------------------------------------
// Node update.  CTA = (node tile, type group); thread = node.  Per type t of the group: U[node, t, :] (the merged
// parts of bin (t, node)) becomes the A operand, Wu_t the B operand, all types accumulate into one TMEM tile.
// The first part row of every bin is fetched cooperatively -- half a warp per 256-byte row -- into a swizzled
// shared-memory tile, one type ahead of the merge; the bin bookkeeping is requested two types ahead.  A bin that
// straddles two tiles of the step kernel (about 3 %) has a second part: its owner thread requests that row and the
// two (max, sum) pairs into registers together with the prefetch, so no lane pays a dependent chain of loads.
// The A operand goes to TENSOR memory (TS form, `tcgen05.st` of the bf16 hi / lo pairs) and both it and the weight tile
// are double-buffered: the product of type t is only waited for when type t + 2 needs its buffers, so the ~1 100-cycle
// issue-to-completion latency of a product is off the per-type chain (it used to end every iteration).
constexpr int kUpdRows = kTile * kD * 4;                      // one row buffer: [128][64] fp32, swizzled rows
constexpr size_t kUpdPayload = 2 * 2 * kWBlock + 2 * kUpdRows + kTile * 4;   // 2 x W (hi, lo), 2 row buffers, part ids
constexpr int kUpdTmemCols = 256;                             // accumulator [0, 64), A operands (hi 32 + lo 32) at 64 and 128
// Two CTAs per SM need 2 x (this + 1 KB system reservation) <= 228 KB, which leaves no room for alignment slack: the
// kernel relies on the 1024-byte alignment of the dynamic shared-memory window (declared, and checked with a trap).
constexpr size_t kUpdSmem = kUpdPayload + 64;
static_assert(2 * (kUpdSmem + 1024) <= 228 * 1024, "node update: two CTAs per SM");

__global__ void __launch_bounds__(kTile) node_update_tc_kernel(AggrView av, int64_t N, int64_t Np, int T, int groups,
                                                               const __nv_bfloat16* __restrict__ wu,
                                                               float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (smem_u32(smem_raw) & 1023u) __trap();                   // no slack is allocated for re-aligning (see kUpdSmem)
  const uint32_t wbuf = smem_u32(smem_raw);                   // weight buffer i: hi at + i * 2 * kWBlock, lo + kWBlock
  const uint32_t rows = wbuf + 2 * 2 * kWBlock;               // row buffer i at + i * kUpdRows
  int* s_part = reinterpret_cast<int*>(smem_raw + 2 * 2 * kWBlock + 2 * kUpdRows);   // [128] part row or -1
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kUpdPayload);   // [2] product on buffer i complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, grp = blockIdx.y;
  if ((tid >> 5) == 0) tmem_alloc<kUpdTmemCols>(tmem_slot);
  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_barrier_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_lane = tmem + ((uint32_t)((tid >> 5) * 32) << 16);
  const int per = (T + groups - 1) / groups;
  const int t0 = grp * per, t1 = min(t0 + per, T);
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t row = row0 + tid;
  const int64_t srow = row < N ? row : N - 1;
  const uint32_t r0 = (uint32_t)(tid >> 4), c4 = (uint32_t)(tid & 15);
  float u[kD];

  struct Bin { int cnt, ls, lp; };
  struct Extra { float mx0, se0, mx1, se1; float4 v[kD / 4]; };   // what the owner of a bin holds besides the shared row
  auto load_bin = [&](int t) {
    Bin b{0, 0, 0};
    if (t < t1) {
      const int64_t bin = (int64_t)t * N + srow;
      b.cnt = av.bin_count[bin]; b.ls = av.bin_lstart[bin]; b.lp = av.bin_lpart[bin];
    }
    return b;
  };
  auto parts_of = [](const Bin& b) { return b.cnt ? ((b.ls + b.cnt - 1) >> 7) - (b.ls >> 7) + 1 : 0; };
  // owner thread: publish the first part row of its bin (-1: empty, or more than two parts), request the scalars
  // and, for a two-part bin, the second row
  auto publish = [&](int t, const Bin& b, Extra& x) {
    int part = -1;
    const int np = t < t1 ? parts_of(b) : 0;
    x.mx0 = x.mx1 = 0.f; x.se0 = 1.f; x.se1 = 0.f;
    if (np == 1 || np == 2) {
      part = av.group_pstart[t] + b.lp;
      if (av.attn) x.se0 = __ldg(av.part_se + part);
      if (np == 2) {
        if (av.attn) { x.mx0 = __ldg(av.part_mx + part); x.mx1 = __ldg(av.part_mx + part + 1); x.se1 = __ldg(av.part_se + part + 1); }
        const float4* __restrict__ v4 = reinterpret_cast<const float4*>(av.part_val + (size_t)(part + 1) * kD);
#pragma unroll
        for (int q = 0; q < kD / 4; ++q) x.v[q] = __ldg(v4 + q);
      }
    }
    s_part[tid] = part;
  };
  // all threads: fetch the published rows of type t, 16 bytes per thread and row, rows r0 + 8k
  auto fetch = [&](int t) {
    if (t < t1) {
      const uint32_t dst = rows + (uint32_t)(t & 1) * kUpdRows;
#pragma unroll
      for (uint32_t k = 0; k < 16; ++k) {
        const uint32_t r = r0 + 8 * k;
        const int part = s_part[r];
        if (part >= 0) cp_async16(dst + r * 256u + ((c4 ^ (r & 15u)) << 4), av.part_val + (size_t)part * kD + 4 * c4);
      }
    }
    cp_async_commit();
  };

  Bin cur = load_bin(t0), nxt = load_bin(t0 + 1);
  Extra x_cur, x_nxt;
  publish(t0, cur, x_cur);
  __syncthreads();
  fetch(t0);
  for (int t = t0; t < t1; ++t) {
    const int i = t - t0, buf = i & 1;
    const uint32_t w_hi = wbuf + (uint32_t)buf * 2 * kWBlock, w_lo = w_hi + kWBlock;
    if (i >= 2) {                                 // the product of type t - 2 has released this weight / operand buffer
      mbar_wait(bars + buf, (uint32_t)((i >> 1) - 1) & 1u);
      fence_after_sync();
    }
    cp_async_weight_tile(w_hi, wu + (size_t)t * 2 * kD * kD, kD, kD);
    cp_async_weight_tile(w_lo, wu + (size_t)t * 2 * kD * kD + kD * kD, kD, kD);
    cp_async_commit();
    publish(t + 1, nxt, x_nxt);
    const Bin after = load_bin(t + 2);
    __syncthreads();                              // part ids of type t + 1 visible; its row buffer is no longer read
    fetch(t + 1);
    cp_async_wait_group<1>();                     // this thread's copies of type t (rows, weights) have landed
    __syncthreads();                              // ... and everybody else's
    const int np = parts_of(cur);
    if (np == 1 || np == 2) {
      // scale of the shared row (part 0) and of the register row (part 1); layers.py:242-251 per-(target, type) softmax
      float sc0 = 1.f, sc1 = 1.f, post = 1.f;
      if (av.attn) {
        if (np == 2) {
          const float M = fmaxf(x_cur.mx0, x_cur.mx1);
          sc0 = __expf(x_cur.mx0 - M); sc1 = __expf(x_cur.mx1 - M);
        }
        post = 1.f / (fmaf(x_cur.se1, np == 2 ? sc1 : 0.f, x_cur.se0 * sc0) + 1e-12f);   // torch_scatter softmax eps
      } else if (av.aggr == PGMP_AGGR_MEAN) {
        post = 1.f / (float)cur.cnt;
      }
      const uint32_t src = rows + (uint32_t)(t & 1) * kUpdRows + (uint32_t)tid * 256u;
      const uint32_t x = (uint32_t)(tid & 15);
      const bool use_max = av.aggr == PGMP_AGGR_MAX && !av.attn;
#pragma unroll
      for (uint32_t q = 0; q < kD / 4; ++q) {
        float4 v = lds128f(src + ((q ^ x) << 4));
        if (np == 2) {
          const float4 w = x_cur.v[q];
          if (use_max) {
            v.x = fmaxf(v.x, w.x); v.y = fmaxf(v.y, w.y); v.z = fmaxf(v.z, w.z); v.w = fmaxf(v.w, w.w);
          } else {
            v.x = fmaf(w.x, sc1, v.x * sc0); v.y = fmaf(w.y, sc1, v.y * sc0);
            v.z = fmaf(w.z, sc1, v.z * sc0); v.w = fmaf(w.w, sc1, v.w * sc0);
          }
        }
        u[4 * q + 0] = v.x * post; u[4 * q + 1] = v.y * post; u[4 * q + 2] = v.z * post; u[4 * q + 3] = v.w * post;
      }
    } else if (np == 0) {
#pragma unroll
      for (int o = 0; o < kD; ++o) u[o] = 0.f;
    } else {
      merge_parts(av, t, srow, N, u);             // a bin of more than 128 edges spread over three or more tiles
    }
    store_split_row_tmem(tmem_lane + 64 + 64 * buf, tmem_lane + 96 + 64 * buf, u);
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (tid < 32 && elect_one()) {
      fence_after_sync();
      issue_gemm_x3_ts<kD>(tmem, tmem + 64 + 64 * buf, tmem + 96 + 64 * buf, w_hi, w_lo, t > t0);
      mma_commit(bars + buf);
    }
    cur = nxt; nxt = after; x_cur = x_nxt;
  }
  cp_async_wait_all();
  if (t1 > t0) {
    const int last = t1 - t0 - 1;                 // the last commit covers every product issued before it
    mbar_wait(bars + (last & 1), (uint32_t)(last >> 1) & 1u);
    fence_after_sync();
    tmem_ld64(tmem, 0, u);
  } else {
#pragma unroll
    for (int o = 0; o < kD; ++o) u[o] = 0.f;
  }
  float4* __restrict__ o4 = reinterpret_cast<float4*>(partial + ((size_t)grp * Np + row) * kD);
#pragma unroll
  for (int q = 0; q < kD / 4; ++q) o4[q] = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
  fence_before_sync();
  __syncthreads();
  if ((tid >> 5) == 0) tmem_dealloc<kUpdTmemCols>(tmem);
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
  for (int g0 = 1; g0 < groups; g0 += 4) {       // four groups' loads in flight; the sum keeps the fixed group order
    float4 w[4][4];
#pragma unroll
    for (int gg = 0; gg < 4; ++gg)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = base + threadIdx.x + k * THREADS;
        w[gg][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g0 + gg < groups)
          w[gg][k] = __ldg(reinterpret_cast<const float4*>(partial + ((size_t)(g0 + gg) * Np + row0 + (idx >> 4)) * kD + 4 * (idx & 15)));
      }
#pragma unroll
    for (int gg = 0; gg < 4; ++gg)
      if (g0 + gg < groups) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[k].x += w[gg][k].x; v[k].y += w[gg][k].y; v[k].z += w[gg][k].z; v[k].w += w[gg][k].w; }
      }
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
  const int total = (int)ceil_div<int64_t>(p.num_nodes, kTile) * n_chunks;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  PGMP_LAUNCH(node_tables_tc_kernel, (unsigned)(total < sms ? total : sms), kTabThreads, kTabSmem, st, w.h0_img, h_img, p.num_nodes,
              p.skip, p.per_type, n_chunks, total, static_cast<const __nv_bfloat16*>(p.tc_wtab), p.b1, p.bm, w.tab_p, w.tab_q,
              w.tab_r);
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
  bool tc_heads = out_slot >= 0 && p.tc_wheads != nullptr;
  if (tc_heads) {      // plain finish, then the heads on the tensor cores (when they have the reference's shape)
    PGMP_LAUNCH((node_finish_kernel<kFinThreads>), tiles, kFinThreads, 0, st, w.upd_partial, N, Np, groups, p.bu, w.h, w.h_img,
                0, p.node_head, p.class_head, nl, cl);
    const int rc = mpn_node_heads_tc(p, w, nl, cl, st, &tc_heads);
    if (rc != PGMP_OK) return rc;
    if (tc_heads) return PGMP_OK;
  }
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
