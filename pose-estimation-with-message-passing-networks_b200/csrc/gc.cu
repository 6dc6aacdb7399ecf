// Graph constructor kernels (sm_100a): heatmap NMS + per-joint top-k / threshold candidates,
// candidate graph (symmetric kNN-50 or fully connected) as a CSR / COO edge index, node-feature
// gather and edge attributes.  Semantics: SURVEY.md Appendix B.1-B.3; reference
// src/graph_constructor/ConstructGraph.py (CG.py) and src/Utils/Utils.py:15-20.
//
// Pipeline (all stream-ordered, no host sync inside):
//   detect: nms_candidates -> select_detections -> layout_nodes -> [knn_adjacency] -> row_degrees
//           -> totals           (host reads the counts once)
//   emit:   emit_nodes, gather_features, emit_edges_{knn,fully}, edge_attr
#include <cstdlib>

#include "common.cuh"

namespace pgmp {
namespace {

constexpr uint32_t kFull = 0xffffffffu;

struct GcWorkspace {
  uint64_t* cand_keys;    // [B*J][cand_capacity]  (score bits << 32 | ~flat index)
  uint32_t* cand_count;   // [B*J]
  int32_t* det_idx;       // [B*J][max_det]  flat pixel index, block 1 then block 2, each index-sorted
  int32_t* det_n1;        // [B*J] size of the top-k block
  int32_t* det_n2;        // [B*J] size of the threshold-extras block
  int32_t* node_xyt;      // [B][max_nodes]  x | y << 12 | type << 24
  float* node_score;      // [B][max_nodes]
  int32_t* node_count;    // [B]
  uint32_t* adj;          // [B][max_nodes][max_nodes/32]  symmetric adjacency bits (kNN only)
  int32_t* rowptr;        // [B][max_nodes + 1]
  int64_t* node_base;     // [B + 1]
  int64_t* edge_base;     // [B + 1]
  int32_t* gnode_xyt;     // [B*max_nodes] packed positions by global node id (written by emit)
  uint32_t* flags;        // [1]
  uint64_t bytes;
};

GcWorkspace carve(const pgmp_gc_params& p) {
  Carver c(p.workspace);
  GcWorkspace w;
  const uint64_t bj = (uint64_t)p.batch * p.num_joints;
  w.cand_keys = c.take<uint64_t>(bj * p.cand_capacity);
  w.cand_count = c.take<uint32_t>(bj);
  w.det_idx = c.take<int32_t>(bj * p.max_det_per_type);
  w.det_n1 = c.take<int32_t>(bj);
  w.det_n2 = c.take<int32_t>(bj);
  w.node_xyt = c.take<int32_t>((uint64_t)p.batch * p.max_nodes);
  w.node_score = c.take<float>((uint64_t)p.batch * p.max_nodes);
  w.node_count = c.take<int32_t>(p.batch);
  w.adj = c.take<uint32_t>(p.graph_type == PGMP_GRAPH_KNN ? (uint64_t)p.batch * p.max_nodes * (p.max_nodes / 32) : 0);
  w.rowptr = c.take<int32_t>((uint64_t)p.batch * (p.max_nodes + 1));
  w.node_base = c.take<int64_t>(p.batch + 1);
  w.edge_base = c.take<int64_t>(p.batch + 1);
  w.gnode_xyt = c.take<int32_t>((uint64_t)p.batch * p.max_nodes);
  w.flags = c.take<uint32_t>(1);
  w.bytes = c.bytes();
  return w;
}

__device__ __forceinline__ int pack_xyt(int x, int y, int t) { return x | (y << 12) | (t << 24); }
__device__ __forceinline__ int px(int v) { return v & 0xfff; }
__device__ __forceinline__ int py(int v) { return (v >> 12) & 0xfff; }
__device__ __forceinline__ int pt(int v) { return (v >> 24) & 0xff; }

// ------------------------------------------------------------------------------------------------
// K1: max-pool NMS (Utils.py:15-20) fused with candidate extraction -- shared-memory row ring fed by the
// bulk-copy (TMA) engine, threshold-first.
//
// One CTA = one (image, joint, strip of rows, <= 1024-column tile).  Rows stream through a ring of 4 stages x
// 8 rows in shared memory: one elected thread issues `cp.async.bulk` copies (a whole stage is ONE contiguous copy
// when the tile spans the map's width) three stages ahead, completion on an mbarrier per stage; no thread ever
// touches a global load or an address computation for the heatmap.
// A pixel matters only if it is a positive window maximum AND can end up in the result: score >= the strip's running
// cut = min(DETECT_THRESHOLD, k-th largest candidate found so far).  So the common path of a warp-row (128 pixels) is:
// one LDS.128, the max of the four values, one compare -- and a single warp-wide REDUX over the 8 rows of a stage.
// Only warp-rows that hold a pixel >= cut run the (2R+1)^2 window maximum, straight from the ring (vertical maxima
// of 4 + 2R columns, then the horizontal one): no register ring, no shuffles, ~70 instructions instead of ~160,
// and only where it matters.  Candidates go to a shared-memory list; whenever it has grown by >= k entries the CTA
// re-selects its k-th largest score (32-bit radix select), raises the cut and compacts the list.
// Zero padding is equivalent to the reference's -inf padding because candidates are positive.
// ------------------------------------------------------------------------------------------------
constexpr int kNmsStageRows = 8;     // >= 2 R for every supported R (pool kernel <= 9)
constexpr int kNmsStages = 4;       // stages s, s + 1 are read while s + 2, s + 3 load (3 stages / 4 CTAs per SM measured slower: 0.184 vs 0.175 ms)
constexpr int kNmsRingRows = kNmsStageRows * kNmsStages;
constexpr int kNmsListCap = 512;     // per-CTA shared-memory candidate queue (consumer warps -> producer warp), a ring
constexpr int kNmsKeepCap = 256;     // the producer warp's list of survivors (>= the running cut)
constexpr int kNmsClaimMargin = 256; // a consumer thread claims a queue slot only while this many are free (<= 256 consumer threads race)

// Scoremap assembly fused into the loader (pgmp_gc_detect_fused): the map the NMS runs on is
//   A_t = (stage2_t + up(stage1_t)) / 2 (avg) or up(stage1_t) (small);  map = A_0, or with the flipped image's outputs
//   map[j][y][x] = (A_0[j][y][x] + A_1[flip[j]][y][W - 1 - x]) / 2
// with ATen's bilinear arithmetic (common.cuh) -- the same operations in the same order as assemble.cu, so the values
// are bit-identical to the materialised map.
constexpr int kNmsFuseLoaders = 8;   // loader warps of the fused kernel (each evaluates S / 4 rows of a stage)
struct AsmDev {
  const float* s1[2]; const float* s2[2];
  int C1, h, w, mode, terms;
  float scale_y, scale_x;
  int flip[32];
  float* out;
};

__device__ __forceinline__ float asm_term(const AsmDev& A, int t, int b, int j, int J, int H, int W, int y, int x) {
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  bilinear_source_index(A.scale_y, y, A.h, y0, y1, wy0, wy1);
  bilinear_source_index(A.scale_x, x, A.w, x0, x1, wx0, wx1);
  const float* __restrict__ plane = A.s1[t] + ((size_t)b * A.C1 + j) * A.h * A.w;
  const float* __restrict__ r0 = plane + (size_t)y0 * A.w;
  const float* __restrict__ r1 = plane + (size_t)y1 * A.w;
  float v = bilinear_combine(wx0, wx1, wy0, wy1, __ldg(r0 + x0), __ldg(r0 + x1), __ldg(r1 + x0), __ldg(r1 + x1));
  if (A.mode == PGMP_ASSEMBLE_AVG) v = __fmul_rn(__fadd_rn(__ldg(A.s2[t] + (((size_t)b * J + j) * H + y) * W + x), v), 0.5f);
  return v;
}
// one pixel of the assembled map (the detections' scores when the map is not materialised)
__device__ __forceinline__ float asm_value(const AsmDev& A, int b, int j, int J, int H, int W, int y, int x) {
  float v = asm_term(A, 0, b, j, J, H, W, y, x);
  if (A.terms == 2) v = __fmul_rn(__fadd_rn(v, asm_term(A, 1, b, A.flip[j], J, H, W, y, W - 1 - x)), 0.5f);
  return v;
}

struct NmsArgs {
  AsmDev asmb;
  const float* scoremaps; const float* mask;
  int J, H, W, xtiles, rows_per_cta, pitch, top_k, use_thr;
  int src_rows, exact2x;                          // FUSE: half-resolution rows a ring stage can depend on; H = 2 h and W = 2 w
  float thr;
  uint64_t* cand_keys; uint32_t* cand_count; int cand_cap; uint32_t* flags;
};

__device__ __forceinline__ uint32_t nms_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 nms_lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void nms_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void nms_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nms_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (!done && ++spins > (1u << 24)) __trap();   // a lost copy must fail loudly, never hang the GPU
  } while (!done);
}
__device__ __forceinline__ void nms_bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(bar) : "memory");
}

// cold path: the CTA's shared list is full -> straight to the global list (already filtered by the running cut)
__device__ __noinline__ void spill_candidate(uint64_t key, uint64_t* __restrict__ gkeys, uint32_t* __restrict__ gcount,
                                             int cand_cap, uint32_t* __restrict__ flags) {
  const uint32_t g = atomicAdd(gcount, 1u);
  if (g < (uint32_t)cand_cap) gkeys[g] = key; else atomicOr(flags, (uint32_t)PGMP_GC_FLAG_CAND_OVERFLOW);
}

// Hand this thread's candidates (up to 4, usually none) to the producer warp through the CTA's shared queue: a slot is
// claimed with one shared atomic while the ring has room (claimed - consumed stays below the capacity, so a claimed
// slot has always been consumed and zeroed), else the candidate goes straight to the global list.  Key = score bits
// << 32 | ~flat index: one 64-bit compare orders by score desc / index asc; a zero key means "empty slot".
__device__ __forceinline__ void emit_candidates(const float (&sc)[4], uint32_t flat0, uint32_t s_keys_a, uint32_t s_cnt_a,
                                                uint64_t* __restrict__ gkeys, uint32_t* __restrict__ gcount, int cand_cap,
                                                uint32_t* __restrict__ flags) {
  uint32_t vm = (sc[0] > 0.f ? 1u : 0u) | (sc[1] > 0.f ? 2u : 0u) | (sc[2] > 0.f ? 4u : 0u) | (sc[3] > 0.f ? 8u : 0u);
  while (vm) {
    const int i = __ffs(vm) - 1;
    vm &= vm - 1;
    const float v = i == 0 ? sc[0] : (i == 1 ? sc[1] : (i == 2 ? sc[2] : sc[3]));
    const uint32_t hi = __float_as_uint(v), lo = ~(flat0 + (uint32_t)i);
    uint32_t claimed, consumed;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(claimed) : "r"(s_cnt_a) : "memory");
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(consumed) : "r"(s_cnt_a + 16u) : "memory");   // s_misc[4]
    if (claimed - consumed < (uint32_t)(kNmsListCap - kNmsClaimMargin)) {
      uint32_t pos;
      asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(pos) : "r"(s_cnt_a) : "memory");   // (a bound never reached: ptxas warp-aggregates add and inc 0xffffffff)
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_keys_a + (pos & (uint32_t)(kNmsListCap - 1)) * 8u), "r"(lo), "r"(hi) : "memory");
    } else {
      spill_candidate(((uint64_t)hi << 32) | (uint64_t)lo, gkeys, gcount, cand_cap, flags);
    }
  }
}

// Append the list entries with score bits >= cut to the (image, joint) list in global memory, one atomic per warp chunk.
__device__ __forceinline__ void nms_write_global(const uint64_t* s_keys, uint32_t n, uint32_t cut, uint64_t* __restrict__ gkeys,
                                                 uint32_t* __restrict__ gcount, int cand_cap, uint32_t* __restrict__ flags) {
  const int t = threadIdx.y * blockDim.x + threadIdx.x, lane = t & 31, nt = blockDim.x * blockDim.y;
  for (uint32_t base = 0; base < n; base += nt) {
    const uint32_t i = base + t;
    const bool keep = i < n && (uint32_t)(s_keys[i < n ? i : 0] >> 32) >= cut;
    const uint32_t m = __ballot_sync(kFull, keep);
    if (m == 0) continue;
    uint32_t g = 0;
    if (lane == 0) g = atomicAdd(gcount, (uint32_t)__popc(m));
    g = __shfl_sync(kFull, g, 0) + __popc(m & ((1u << lane) - 1u));
    if (keep) {
      if (g < (uint32_t)cand_cap) gkeys[g] = s_keys[i]; else atomicOr(flags, (uint32_t)PGMP_GC_FLAG_CAND_OVERFLOW);
    }
  }
}

template <bool VEC>
__device__ __forceinline__ float4 nms_ldg4(const float* __restrict__ p, int n) {   // n = valid elements (scalar path)
  if (VEC) return __ldg(reinterpret_cast<const float4*>(p));
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n > 0) v.x = __ldg(p);
  if (n > 1) v.y = __ldg(p + 1);
  if (n > 2) v.z = __ldg(p + 2);
  if (n > 3) v.w = __ldg(p + 3);
  return v;
}

__device__ __forceinline__ float2 nms_lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}

__device__ __forceinline__ void nms_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// One warp: the k-th largest score (bit pattern) among the first n list entries, n >= k (entries whose 8-byte store
// has not landed yet read as 0 and only lower the result).  Same radix select as nms_kth_largest, warp-synchronous.
__device__ __forceinline__ uint32_t nms_kth_largest_warp(const volatile uint64_t* s_keys, uint32_t n, int k, uint32_t* s_hist) {
  const int lane = threadIdx.x & 31;
  uint32_t prefix = 0, rem = (uint32_t)k;
  for (int shift = 24; shift >= 0; shift -= 8) {
#pragma unroll
    for (int q = 0; q < 8; ++q) s_hist[lane * 8 + q] = 0;
    __syncwarp();
    const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
    for (uint32_t i = lane; i < n; i += 32) {
      const uint32_t sc = (uint32_t)(s_keys[i] >> 32);
      if ((sc & himask) == prefix) atomicAdd(&s_hist[(sc >> shift) & 0xff], 1u);
    }
    __syncwarp();
    uint32_t mine = 0;                     // lane L owns bins [8 (31 - L), 8 (31 - L) + 8): lane 0 = the highest bins
    const int b0 = 8 * (31 - lane);
#pragma unroll
    for (int q = 0; q < 8; ++q) mine += s_hist[b0 + q];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t hit = __ballot_sync(kFull, incl >= rem);
    if (hit == 0) return 0;                // (cannot happen for n >= k; never return an unproven bound)
    const int owner = __ffs(hit) - 1;
    uint32_t r2 = rem - (incl - mine);
    int d = b0 + 7;
    if (lane == owner) {
      for (; d > b0; --d) {
        if (s_hist[d] >= r2) break;
        r2 -= s_hist[d];
      }
    }
    prefix |= (uint32_t)__shfl_sync(kFull, d, owner) << shift;
    rem = __shfl_sync(kFull, r2, owner);
    __syncwarp();
  }
  return prefix;
}

__device__ __forceinline__ float nms_max3(float a, float b, float c) {   // FMNMX3
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// a.x >= b.x || a.y >= b.y || ... as one predicate chain (no SEL / LOP per comparison, no branches)
__device__ __forceinline__ bool nms_any_ge(float4 a, float4 b) {
  uint32_t r;
  asm("{\n .reg .pred p;\n setp.ge.f32 p, %1, %5;\n setp.ge.or.f32 p, %2, %6, p;\n setp.ge.or.f32 p, %3, %7, p;\n"
      " setp.ge.or.f32 p, %4, %8, p;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(r) : "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w));
  return r != 0;
}

// blockDim = 32 x (column bands of the tile x RW row groups + 2).  Consumer warp (band, row group) owns 128 columns
// and RPW = 8 / RW consecutive rows of every stage.  The last two warps serve them:
//   LOADER    issues the bulk copies as ring stages are released (one `empty` mbarrier per stage, one arrival per
//             consumer warp) and zero-fills the ring rows that lie outside the image, so consumers never test a row;
//   SELECTOR  drains the candidate queue, keeps the survivors and re-selects the running cut.
// Consumer warps never meet at a CTA-wide barrier until the strip ends: a warp that runs into a blob does not hold
// up the others, and neither the copies nor the cut wait for each other.
// FUSE: the loader is kNmsFuseLoaders warps that EVALUATE the rows of the map (AsmDev) instead of copying them.
template <int R, int RPW, bool VEC, bool MASK, bool FUSE>
__global__ void __launch_bounds__(FUSE ? 288 + 32 * kNmsFuseLoaders : 320, FUSE ? 2 : 3) nms_candidates_kernel(const NmsArgs a) {
  constexpr int S = kNmsStageRows, RW = S / RPW, HALO = R > 0 ? 1 : 0;
  extern __shared__ __align__(128) uint8_t nms_smem[];
  const int P = a.pitch, H = a.H, W = a.W;
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(nms_smem + (size_t)kNmsRingRows * P * 4);   // the queue (ring)
  uint64_t* s_keep = s_keys + kNmsListCap;                               // the selector's survivors
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(s_keep + kNmsKeepCap);
  uint64_t* s_bars = reinterpret_cast<uint64_t*>(s_hist + 256);          // full[4], empty[4], raw[4] (FUSE: the stage's bulk copies)
  // [0] queue entries claimed [1] effective cut (score bits) [4] queue entries consumed [5] survivors [6] consumer warps done
  // [7] next work item
  uint32_t* s_misc = reinterpret_cast<uint32_t*>(s_bars + 3 * kNmsStages);
  int2* s_ci = reinterpret_cast<int2*>(s_misc + 16);                     // FUSE: per map column the two source columns ...
  float2* s_cw = reinterpret_cast<float2*>(s_ci + P);                    // ... and their weights
  const uint32_t ring_a = nms_smem_u32(nms_smem), keys_a = nms_smem_u32(s_keys), cnt_a = nms_smem_u32(s_misc);
  const uint32_t full_a = nms_smem_u32(s_bars), empty_a = full_a + 8u * kNmsStages;

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr int NL = FUSE ? kNmsFuseLoaders : 1;                         // loader warps
  const int n_cons = (int)(blockDim.x >> 5) - 1 - NL;                    // consumer warps
  const int bands = n_cons / RW;
  const int b = blockIdx.z, j = blockIdx.y;
  const int strip = blockIdx.x / a.xtiles, xt = blockIdx.x - strip * a.xtiles;
  const int bj = b * a.J + j;
  const float* __restrict__ map = a.scoremaps + (size_t)bj * H * W;
  const float* __restrict__ mk = MASK ? a.mask + (size_t)b * H * W : nullptr;
  uint64_t* __restrict__ gkeys = a.cand_keys + (size_t)bj * a.cand_cap;
  uint32_t* __restrict__ gcount = &a.cand_count[bj];
  const int y0 = strip * a.rows_per_cta, y_end = min(y0 + a.rows_per_cta, H);
  const int tile_cols = 128 * bands;
  const int x0 = xt * tile_cols;
  const int gx0 = x0 > 0 ? x0 - 4 : 0;                                   // first map column held in a ring row
  const int gcount_cols = min(W, x0 + tile_cols + 4) - gx0;               // columns per ring row
  const uint32_t rowb = (uint32_t)P * 4u;
  const int ybase = y0 - R;                                              // map row of ring row 0 (may be negative)
  const int y_need = min(H, y_end + R);                                  // in-image rows [max(0, ybase), y_need) are read ...
  const int n_iter = (y_end - y0 + S - 1) / S;
  const int n_stage = (y_end + R - ybase + S - 1) / S;                    // ... out-of-image rows up to y_end + R as zeros
  const uint32_t thr_bits = __float_as_uint(a.thr);

  for (int i = t; i < kNmsListCap; i += blockDim.x) s_keys[i] = 0ull;     // 0 = "empty slot"
  if (t == 0) {
    for (int i = 0; i < kNmsStages; ++i) {
      nms_mbar_init(full_a + 8u * i, (uint32_t)NL); nms_mbar_init(empty_a + 8u * i, (uint32_t)n_cons);
      nms_mbar_init(empty_a + 8u * (kNmsStages + i), 1);
    }
    s_misc[0] = 0; s_misc[4] = 0; s_misc[5] = 0; s_misc[6] = 0; s_misc[7] = 0;
    s_misc[1] = 1u;      // smallest positive float: "x >= cut" == "x > 0" until a real cut exists
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (FUSE) {
    for (int x = t; x < W; x += blockDim.x) {
      int i0, i1;
      float l0, l1;
      bilinear_source_index(a.asmb.scale_x, x, a.asmb.w, i0, i1, l0, l1);
      s_ci[x] = make_int2(i0, i1);
      s_cw[x] = make_float2(l0, l1);
    }
  }
  __syncthreads();

  if (FUSE && warp >= n_cons && warp < n_cons + NL) {
    // ---- ASSEMBLING LOADERS.  Global memory is only touched by the bulk-copy engine: per ring stage one copy brings the
    //      full-resolution rows (stage 2) straight into the ring slot and one the few half-resolution rows they depend
    //      on (stage 1) into a double-buffered source tile, two stages ahead; the loader warps (warp lw: rows lw, lw + NL,
    //      ... of the stage) then turn the slot into the assembled rows IN PLACE -- source values, column indices /
    //      weights and the stage-2 value all come from shared memory -- and hand it to the consumers.  One tile spans
    //      the map's width (launch_nms), ring column = map column.  Only the flipped image's stage-2 values (mirrored
    //      columns) are read with ordinary loads.
    const AsmDev& A = a.asmb;
    const int lw = warp - n_cons;
    const int sw = A.w, src_cap = a.src_rows * sw;               // floats per source tile
    float* s_src = reinterpret_cast<float*>(s_cw + P);           // [2 buffers][terms][src_rows][w]
    const uint32_t src_a = nms_smem_u32(s_src), raw_a = empty_a + 8u * kNmsStages;
    const bool avg = A.mode == PGMP_ASSEMBLE_AVG, two = A.terms == 2;
    const int jf = two ? A.flip[j] : j;
    const float* __restrict__ plane0 = A.s1[0] + ((size_t)b * A.C1 + j) * A.h * sw;
    const float* __restrict__ plane1 = A.s1[two ? 1 : 0] + ((size_t)b * A.C1 + jf) * A.h * sw;
    const float* __restrict__ full0 = avg ? A.s2[0] + (size_t)bj * H * W : nullptr;
    const float* __restrict__ full1 = avg && two ? A.s2[1] + ((size_t)b * a.J + jf) * H * W : nullptr;
    auto src_first = [&](int y) { int i0, i1; float l0, l1; bilinear_source_index(A.scale_y, y, A.h, i0, i1, l0, l1); return i0; };
    auto src_last = [&](int y) { int i0, i1; float l0, l1; bilinear_source_index(A.scale_y, y, A.h, i0, i1, l0, l1); return i1; };
    auto issue = [&](int L) {                                    // one thread: the copies of stage L
      const unsigned stg = (unsigned)L % kNmsStages;
      if (L >= kNmsStages) {
        nms_mbar_wait(empty_a + 8u * stg, (uint32_t)(L / kNmsStages + 1) & 1u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot was last written by ordinary stores
      }
      const int ys = ybase + L * S, r_lo = max(ys, 0), nrows = min(ys + S, y_need) - r_lo;
      const uint32_t bar = raw_a + 8u * stg;
      if (nrows <= 0) { nms_mbar_arrive(bar); return; }
      const int sy = src_first(r_lo), ns = src_last(r_lo + nrows - 1) - sy + 1;
      const uint32_t sbytes = (uint32_t)(ns * sw) * 4u, fbytes = avg ? (uint32_t)nrows * rowb : 0u;
      nms_mbar_expect_tx(bar, fbytes + sbytes * (two ? 2u : 1u));
      if (avg) nms_bulk_load(ring_a + (uint32_t)((unsigned)(r_lo - ybase) % kNmsRingRows) * rowb, full0 + (size_t)r_lo * W, fbytes, bar);
      const uint32_t sdst = src_a + (uint32_t)((L & 1) * 2 * src_cap) * 4u;
      nms_bulk_load(sdst, plane0 + (size_t)sy * sw, sbytes, bar);
      if (two) nms_bulk_load(sdst + (uint32_t)src_cap * 4u, plane1 + (size_t)sy * sw, sbytes, bar);
    };
    if (lw == 0 && lane == 0) {
      issue(0);
      if (n_stage > 1) issue(1);
    }
    for (int L = 0; L < n_stage; ++L) {
      const unsigned stg = (unsigned)L % kNmsStages;
      const int ys = ybase + L * S, r_lo = max(ys, 0), nrows = min(ys + S, y_need) - r_lo;
      nms_mbar_wait(raw_a + 8u * stg, (uint32_t)(L / kNmsStages) & 1u);
      const int sy = nrows > 0 ? src_first(r_lo) : 0;
      float* slot = reinterpret_cast<float*>(nms_smem) + (size_t)(stg * S) * P;
      const float* __restrict__ src0 = s_src + (size_t)(L & 1) * 2 * src_cap;
      const float* __restrict__ src1 = src0 + src_cap;
      for (int r = lw; r < S; r += NL) {
        const int y = ys + r;
        float4* dst = reinterpret_cast<float4*>(slot + (size_t)r * P);
        if ((unsigned)y >= (unsigned)H) {             // rows above / below the image read as zeros
          for (int i = lane; i < P / 4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          continue;
        }
        if (y >= y_need) continue;
        int y0s, y1s;
        float wy0, wy1;
        bilinear_source_index(A.scale_y, y, A.h, y0s, y1s, wy0, wy1);
        const int o0 = (y0s - sy) * sw, o1 = (y1s - sy) * sw;
        float4* __restrict__ gout = A.out ? reinterpret_cast<float4*>(A.out + ((size_t)bj * H + y) * W) : nullptr;
        const float* __restrict__ f1row = full1 ? full1 + (size_t)y * W : nullptr;
        // up-sampled stage 1 at map columns x .. x + 3.  Exactly doubled maps (H = 2 h, W = 2 w: HigherHRNet) need the four
        // source columns x / 2 - 1 .. x / 2 + 2 only, with weights 1/4, 3/4 (what bilinear_source_index yields there:
        // scale 1/2 makes every step exact; column 0 has weights 1, 0) -- 6 shared loads; other sizes go through the
        // per-column index / weight tables.
        auto up4 = [&](const float* __restrict__ sp, int x, float (&o)[4]) {
          const float* __restrict__ p0 = sp + o0; const float* __restrict__ p1 = sp + o1;
          if (a.exact2x) {
            const int m2 = x >> 1, ia = max(m2 - 1, 0), id = min(m2 + 2, sw - 1);
            const float2 t = *reinterpret_cast<const float2*>(p0 + m2), u = *reinterpret_cast<const float2*>(p1 + m2);
            const float ta = p0[ia], td = p0[id], ua = p1[ia], ud = p1[id];
            if (x == 0) o[0] = bilinear_combine(1.f, 0.f, wy0, wy1, t.x, t.y, u.x, u.y);
            else o[0] = bilinear_combine(0.25f, 0.75f, wy0, wy1, ta, t.x, ua, u.x);
            o[1] = bilinear_combine(0.75f, 0.25f, wy0, wy1, t.x, t.y, u.x, u.y);
            o[2] = bilinear_combine(0.25f, 0.75f, wy0, wy1, t.x, t.y, u.x, u.y);
            o[3] = bilinear_combine(0.75f, 0.25f, wy0, wy1, t.y, td, u.y, ud);
            return;
          }
          const int4 ca = *reinterpret_cast<const int4*>(s_ci + x), cb = *reinterpret_cast<const int4*>(s_ci + x + 2);
          const float4 wa = *reinterpret_cast<const float4*>(s_cw + x), wb = *reinterpret_cast<const float4*>(s_cw + x + 2);
          o[0] = bilinear_combine(wa.x, wa.y, wy0, wy1, p0[ca.x], p0[ca.y], p1[ca.x], p1[ca.y]);
          o[1] = bilinear_combine(wa.z, wa.w, wy0, wy1, p0[ca.z], p0[ca.w], p1[ca.z], p1[ca.w]);
          o[2] = bilinear_combine(wb.x, wb.y, wy0, wy1, p0[cb.x], p0[cb.y], p1[cb.x], p1[cb.y]);
          o[3] = bilinear_combine(wb.z, wb.w, wy0, wy1, p0[cb.z], p0[cb.w], p1[cb.z], p1[cb.w]);
        };
        auto group = [&](int c4) -> float4 {          // the assembled map at columns 4 c4 .. 4 c4 + 3 of row y
          const int x = 4 * c4;
          float4 fv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (f1row) fv = __ldg(reinterpret_cast<const float4*>(f1row + (W - 4 - x)));   // requested first: the one global load
          float o[4];
          up4(src0, x, o);
          if (avg) {
            const float4 v = dst[c4];                                         // the stage-2 values the bulk copy put there
            o[0] = __fmul_rn(__fadd_rn(v.x, o[0]), 0.5f); o[1] = __fmul_rn(__fadd_rn(v.y, o[1]), 0.5f);
            o[2] = __fmul_rn(__fadd_rn(v.z, o[2]), 0.5f); o[3] = __fmul_rn(__fadd_rn(v.w, o[3]), 0.5f);
          }
          if (two) {                                  // the flipped image's map, mirrored: columns W - 1 - (x + q)
            float f[4];
            up4(src1, W - 4 - x, f);
            if (avg) {
              f[0] = __fmul_rn(__fadd_rn(fv.x, f[0]), 0.5f); f[1] = __fmul_rn(__fadd_rn(fv.y, f[1]), 0.5f);
              f[2] = __fmul_rn(__fadd_rn(fv.z, f[2]), 0.5f); f[3] = __fmul_rn(__fadd_rn(fv.w, f[3]), 0.5f);
            }
            o[0] = __fmul_rn(__fadd_rn(o[0], f[3]), 0.5f); o[1] = __fmul_rn(__fadd_rn(o[1], f[2]), 0.5f);
            o[2] = __fmul_rn(__fadd_rn(o[2], f[1]), 0.5f); o[3] = __fmul_rn(__fadd_rn(o[3], f[0]), 0.5f);
          }
          return make_float4(o[0], o[1], o[2], o[3]);
        };
        int c4 = lane;
        if (!two)                                     // two groups per pass: the second group's loads are not held up by the first's store
          for (; c4 + 32 < W / 4; c4 += 64) {
            const float4 va = group(c4), vb = group(c4 + 32);
            dst[c4] = va; dst[c4 + 32] = vb;
            if (gout) { gout[c4] = va; gout[c4 + 32] = vb; }
          }
        for (; c4 < W / 4; c4 += 32) {
          const float4 v = group(c4);
          dst[c4] = v;
          if (gout) gout[c4] = v;
        }
      }
      __syncwarp();
      if (lane == 0) nms_mbar_arrive(full_a + 8u * stg);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NL) : "memory");   // every loader is done with this stage's source tile
      if (lw == 0 && lane == 0 && L + 2 < n_stage) issue(L + 2);
    }
  } else if (!FUSE && warp == n_cons) {
    // ---- LOADER: stage L = map rows [ybase + L S, ybase + (L + 1) S) -> ring rows (L mod 4) S ...
    for (int L = 0; L < n_stage; ++L) {
      if (L >= kNmsStages) nms_mbar_wait(empty_a + 8u * ((unsigned)L % kNmsStages), (uint32_t)(L / kNmsStages + 1) & 1u);
      const int ys = ybase + L * S;
      const int r_lo = max(ys, 0);
      const int nrows = min(ys + S, y_need) - r_lo;
      const uint32_t bar = full_a + 8u * ((unsigned)L % kNmsStages);
      float* slot = reinterpret_cast<float*>(nms_smem) + (size_t)(((unsigned)L % kNmsStages) * S) * P;
      if (ys < 0 || ys + S > H) {                   // rows above / below the image read as zeros (= the -inf padding, scores > 0)
        for (int r = 0; r < S; ++r)
          if ((unsigned)(ys + r) >= (unsigned)H)
            for (int i = lane; i < P; i += 32) slot[(size_t)r * P + i] = 0.f;
        __syncwarp();
      }
      if (nrows <= 0) {
        if (lane == 0) nms_mbar_arrive(bar);
      } else if (VEC) {
        if (lane == 0) {
          const uint32_t dst = ring_a + (uint32_t)((unsigned)(r_lo - ybase) % kNmsRingRows) * rowb;
          if (a.xtiles == 1) {                      // the tile spans the width: rows are contiguous in memory and in the ring
            const uint32_t bytes = (uint32_t)nrows * rowb;
            nms_mbar_expect_tx(bar, bytes);
            nms_bulk_load(dst, map + (size_t)r_lo * W, bytes, bar);
          } else {
            const uint32_t bytes = (uint32_t)gcount_cols * 4u;
            nms_mbar_expect_tx(bar, bytes * (uint32_t)nrows);
            for (int r = 0; r < nrows; ++r) nms_bulk_load(dst + (uint32_t)r * rowb, map + (size_t)(r_lo + r) * W + gx0, bytes, bar);
          }
        }
      } else {                                      // unaligned maps: the warp copies, zero-filling up to the pitch
        float* ring = slot + (size_t)(r_lo - ys) * P;
        for (int i = lane; i < nrows * P; i += 32) {
          const int r = i / P, x = i - r * P;
          ring[i] = x < gcount_cols ? __ldg(map + (size_t)(r_lo + r) * W + gx0 + x) : 0.f;
        }
        __syncwarp();
        if (lane == 0) nms_mbar_arrive(bar);
      }
      __syncwarp();
    }
  } else if (warp == n_cons + NL) {
    // ---- SELECTOR: drain the candidate queue.  Entries are taken in claim order up to the first one whose 8-byte store
    //      has not landed yet; survivors (>= cut) join the keep list, the slots are zeroed and released.
    const uint32_t trigger = (uint32_t)min(kNmsKeepCap - 64, max(3 * a.top_k, 96));   // survivors that start a re-selection
    const uint32_t lt_mask = (1u << lane) - 1u;
    volatile uint64_t* q = s_keys;
    uint32_t rd = 0, kc = 0, cut = 1u;
    for (;;) {
      const uint32_t done = *reinterpret_cast<volatile uint32_t*>(&s_misc[6]);     // (read before the queue: nothing is lost)
      const uint32_t claimed = *reinterpret_cast<volatile uint32_t*>(&s_misc[0]);
      if (claimed == rd) {
        if (done == (uint32_t)n_cons) break;
        __nanosleep(256);
        continue;
      }
      const uint32_t i = rd + (uint32_t)lane;
      const uint64_t key = (int32_t)(claimed - i) > 0 ? q[i & (kNmsListCap - 1)] : 0ull;
      const uint32_t landed = __ballot_sync(kFull, key != 0ull);
      const int len = landed == kFull ? 32 : __ffs(~landed) - 1;                   // leading run of landed entries
      const bool mine = lane < len;
      const bool keep = mine && (uint32_t)(key >> 32) >= cut;
      const uint32_t km = __ballot_sync(kFull, keep);
      if (keep) s_keep[kc + __popc(km & lt_mask)] = key;
      if (mine) q[i & (kNmsListCap - 1)] = 0ull;
      kc += __popc(km);
      rd += (uint32_t)len;
      __threadfence_block();
      __syncwarp();
      if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&s_misc[4]) = rd;
      if (kc >= trigger) {
        if (kc > (uint32_t)a.top_k) {          // the k-th largest survivor is a proven lower bound of the strip's k-th largest
          const uint32_t kth = nms_kth_largest_warp(s_keep, kc, a.top_k, s_hist);
          uint32_t eff = max(kth, cut);                                            // cuts only rise
          if (a.use_thr) eff = min(eff, max(thr_bits, 1u));                        // scores >= DETECT_THRESHOLD always matter
          cut = eff;
          if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&s_misc[1]) = cut;
          uint32_t w = 0;                                                          // in-place compaction, 32 entries at a time
          for (uint32_t base = 0; base < kc; base += 32) {
            const uint32_t e = base + (uint32_t)lane;
            const uint64_t k2 = e < kc ? s_keep[e] : 0ull;
            const bool kp = e < kc && (uint32_t)(k2 >> 32) >= cut;
            const uint32_t m2 = __ballot_sync(kFull, kp);
            if (kp) s_keep[w + __popc(m2 & lt_mask)] = k2;
            w += __popc(m2);
            __syncwarp();
          }
          kc = w;
        }
        if (kc >= trigger) {                   // ties / many scores above the threshold: hand the survivors over, keep the cut
          for (uint32_t base = 0; base < kc; base += 32) {
            const uint32_t e = base + (uint32_t)lane;
            const uint32_t m2 = __ballot_sync(kFull, e < kc);
            uint32_t g = 0;
            if (lane == 0) g = atomicAdd(gcount, (uint32_t)__popc(m2));
            g = __shfl_sync(kFull, g, 0) + (uint32_t)lane;
            if (e < kc) {
              if (g < (uint32_t)a.cand_cap) gkeys[g] = s_keep[e]; else atomicOr(a.flags, (uint32_t)PGMP_GC_FLAG_CAND_OVERFLOW);
            }
          }
          kc = 0;
        }
      }
    }
    if (lane == 0) s_misc[5] = kc;
  } else {
    // ---- CONSUMERS
    // Work items (stage, column band, row group) are handed out in order by one shared counter: a warp that runs into
    // blobs takes fewer items, so no warp waits for a slower one (the only coupling is the ring's depth).
    const int n_items = n_iter * n_cons;
    const uint32_t inv_cons = 0xffffffffu / (uint32_t)n_cons + 1u, inv_bands = 0xffffffffu / (uint32_t)bands + 1u;
    for (;;) {
      int item = 0;
      if (lane == 0) item = (int)atomicAdd(&s_misc[7], 1u);
      item = __shfl_sync(kFull, item, 0);
      if (item >= n_items) break;
      const int s = (int)__umulhi((uint32_t)item, inv_cons), unit = item - s * n_cons;   // item / n_cons (item < 2^16)
      const int wr = bands == 1 ? unit : (int)__umulhi((uint32_t)unit, inv_bands);        // unit / bands
      const int tx = (unit - wr * bands) * 32 + lane;                                     // row group wr, column thread tx
      const int c = x0 + 4 * tx;                                           // this thread's first column
      const bool has_v = c < W, has_l = has_v && c >= 4, has_r = c + 4 < W;
      // (threads beyond the map's width read column gx0 instead and never report anything: no zero-filled registers)
      const uint32_t col_a = ring_a + (uint32_t)(has_v ? c - gx0 : 0) * 4u;  // + ring row * P * 4
      nms_mbar_wait(full_a + 8u * ((unsigned)s % kNmsStages), (uint32_t)(s / kNmsStages) & 1u);
      if (s + 1 < n_stage) nms_mbar_wait(full_a + 8u * ((unsigned)(s + 1) % kNmsStages), (uint32_t)((s + 1) / kNmsStages) & 1u);
      const float cutf = __uint_as_float(*reinterpret_cast<volatile uint32_t*>(&s_misc[1]));
      const int yr = y0 + s * S + wr * RPW;         // the item's first row
      const int rr0 = (int)((unsigned)(yr - HALO - ybase) % kNmsRingRows);   // ring row of its first (halo) row
      // ---- common path: which of the warp's rows hold a pixel that can matter -- score >= cut and, cheaply, not below
      //      the pixel above / below it (rows of a blob other than its ridge fail here); one REDUX for the rows
      uint32_t pm = 0;
      {
        float4 v[RPW + 2 * HALO];
#pragma unroll
        for (int i = 0; i < RPW + 2 * HALO; ++i) {
          v[i] = nms_lds128(col_a + (uint32_t)(rr0 + i - (rr0 + i >= kNmsRingRows ? kNmsRingRows : 0)) * rowb);
        }
#pragma unroll
        for (int i = 0; i < RPW; ++i) {
          const float4 x = v[i + HALO];
          bool p;
          if (MASK) {
            float4 sc = x;
            if (has_v && yr + i < y_end) {
              const float4 m4 = nms_ldg4<VEC>(mk + (size_t)(yr + i) * W + c, W - c);
              sc.x *= m4.x; sc.y *= m4.y; sc.z *= m4.z; sc.w *= m4.w;
            }
            if (R > 0) {
              const float4 u = v[i], d = v[i + 2 * HALO];
              p = (sc.x >= cutf && x.x >= fmaxf(u.x, d.x)) || (sc.y >= cutf && x.y >= fmaxf(u.y, d.y)) ||
                  (sc.z >= cutf && x.z >= fmaxf(u.z, d.z)) || (sc.w >= cutf && x.w >= fmaxf(u.w, d.w));
            } else {
              p = fmaxf(fmaxf(sc.x, sc.y), fmaxf(sc.z, sc.w)) >= cutf;
            }
          } else if (R > 0) {
            const float4 u = v[i], d = v[i + 2 * HALO];
            // ... nor below its neighbours inside this thread's four columns
            p = nms_any_ge(x, make_float4(fmaxf(nms_max3(u.x, d.x, cutf), x.y), nms_max3(nms_max3(u.y, d.y, cutf), x.x, x.z),
                                          nms_max3(nms_max3(u.z, d.z, cutf), x.y, x.w), fmaxf(nms_max3(u.w, d.w, cutf), x.z)));
          } else {
            p = nms_max3(x.x, x.y, fmaxf(x.z, x.w)) >= cutf;
          }
          pm |= (p ? 1u : 0u) << i;
        }
      }
      uint32_t rows = __reduce_or_sync(kFull, has_v ? pm : 0u);
      if (yr + RPW > y_end) rows &= (1u << max(y_end - yr, 0)) - 1u;        // the strip ends inside this stage
      // ---- rows with such a pixel in this warp's 128 columns: window maxima straight from the ring
      while (rows) {
        const int r = __ffs(rows) - 1;
        rows &= rows - 1;
        const int y = yr + r;
        const int rr1 = rr0 + HALO + r - R;          // ring row of the window's first row (+ 32: positive)
        // column maxima over the 2R+1 rows, two rows per 3-input max; the centre row's own 4 values are kept
        float cm[12], xc[4] = {0.f, 0.f, 0.f, 0.f};
        auto load_row = [&](int dy, float (&e)[12]) {
          const uint32_t ra = col_a + (uint32_t)((unsigned)(rr1 + dy + kNmsRingRows) % kNmsRingRows) * rowb;
          float4 l = make_float4(0.f, 0.f, 0.f, 0.f), rr = l;
          const float4 v = nms_lds128(ra);
          if (R > 2) {
            if (has_l) l = nms_lds128(ra - 16u);
            if (has_r) rr = nms_lds128(ra + 16u);
          } else if (R > 0) {
            if (has_l) { const float2 q = nms_lds64(ra - 8u); l.z = q.x; l.w = q.y; }
            if (has_r) { const float2 q = nms_lds64(ra + 16u); rr.x = q.x; rr.y = q.y; }
          }
          e[0] = l.x; e[1] = l.y; e[2] = l.z; e[3] = l.w; e[4] = v.x; e[5] = v.y; e[6] = v.z; e[7] = v.w;
          e[8] = rr.x; e[9] = rr.y; e[10] = rr.z; e[11] = rr.w;
          if (dy == R) { xc[0] = v.x; xc[1] = v.y; xc[2] = v.z; xc[3] = v.w; }
        };
        load_row(0, cm);
#pragma unroll
        for (int dy = 1; dy + 1 <= 2 * R; dy += 2) {
          float ea[12], eb[12];
          load_row(dy, ea);
          load_row(dy + 1, eb);
#pragma unroll
          for (int i = 4 - R; i < 8 + R; ++i) cm[i] = nms_max3(cm[i], ea[i], eb[i]);
        }
        float sc[4];
        float4 m4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (MASK) { if (has_v) m4 = nms_ldg4<VEC>(mk + (size_t)y * W + c, W - c); }   // CG.py:1163-1165
        const float mq[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float m = cm[4 + q - R];
#pragma unroll
          for (int d = -R + 1; d + 1 <= R; d += 2) m = nms_max3(m, cm[4 + q + d], cm[4 + q + d + 1]);
          const float x = xc[q];
          // (out-of-image rows / columns hold zeros or are not loaded, so x > 0 already excludes them)
          float v = (x > 0.f && x == m) ? (MASK ? x * mq[q] : x) : 0.f;
          sc[q] = (has_v && v >= cutf) ? v : 0.f;
        }
        emit_candidates(sc, (uint32_t)(y * W + c), keys_a, cnt_a, gkeys, gcount, a.cand_cap, a.flags);
      }
      __syncwarp();
      if (lane == 0) nms_mbar_arrive(empty_a + 8u * ((unsigned)s % kNmsStages));   // one of the n_cons items of stage s is done
    }
    if (lane == 0) atomicAdd(&s_misc[6], 1u);
  }
  // ---- end of the strip: what can matter globally -- the survivors and whatever still sits in the queue, >= the cut
  //      (a proven lower bound of the strip's k-th largest score, already capped by the threshold)
  __syncthreads();
  const uint32_t cut = s_misc[1];
  nms_write_global(s_keys, (uint32_t)kNmsListCap, cut, gkeys, gcount, a.cand_cap, a.flags);   // empty slots are zero < cut
  nms_write_global(s_keep, s_misc[5], cut, gkeys, gcount, a.cand_cap, a.flags);
}

// ------------------------------------------------------------------------------------------------
// K2: per (image, joint) selection (CG.py:1166-1195).  Sort the candidate keys descending
// (score desc, flat index asc -- the build's top-k tie rule), block 1 = first min(top_k, n),
// block 2 = the following entries with score >= threshold (cat_unique, CG.py:1182,1199-1209),
// then order each block by flat index = the reference's (type, y, x) nonzero() order.
// ------------------------------------------------------------------------------------------------
__device__ void bitonic_sort_asc(uint64_t* s, int P) {
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint64_t a = s[i], c = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > c) == up) { s[i] = c; s[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// No-threshold path with fewer than k positive maxima (clean or crowd-masked maps): torch.topk fills the block with
// zero-score pixels (CG.py:1187; `+ 1e-10` makes them nonzero, :1189) -- by this build's tie rule the zero-score
// pixels with the lowest flat indices.  A pixel's score x * nms * mask is zero unless it is a window maximum
// (-inf padding, Utils.py:17) with x != 0 and mask != 0.  One warp scans the map from index 0.
template <int DUMMY = 0>
__device__ void pad_with_zero_pixels(const float* __restrict__ map, const float* __restrict__ mask, int H, int W, int R,
                                     int need, int* s_pad, int* s_npad) {
  const int lane = threadIdx.x & 31;
  int found = 0;
  for (int base = 0; base < H * W && found < need; base += 32) {
    const int i = base + lane;
    bool zero = false;
    if (i < H * W) {
      const int y = i / W, x = i - y * W;
      const float v = map[i];
      bool is_max = true;
      for (int dy = -R; dy <= R && is_max; ++dy)
        for (int dx = -R; dx <= R; ++dx) {
          const int yy = y + dy, xx = x + dx;
          if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W && map[yy * W + xx] > v) { is_max = false; break; }
        }
      const float sc = is_max ? (mask ? v * mask[i] : v) : 0.f;
      zero = sc == 0.f;
    }
    const uint32_t m = __ballot_sync(kFull, zero);
    const int pos = found + __popc(m & ((1u << lane) - 1u));
    if (zero && pos < need) s_pad[pos] = i;
    found += __popc(m);
  }
  if (lane == 0) *s_npad = min(found, need);
}

__global__ void __launch_bounds__(256) select_detections_kernel(
    const uint64_t* __restrict__ cand_keys, const uint32_t* __restrict__ cand_count, int cand_cap, int top_k,
    int use_thr, float thr, int max_det, const float* __restrict__ scoremaps, const float* __restrict__ mask, int J, int H,
    int W, int R, int32_t* __restrict__ det_idx, int32_t* __restrict__ det_n1, int32_t* __restrict__ det_n2,
    uint32_t* __restrict__ flags) {
  extern __shared__ uint64_t s[];
  __shared__ int s_n2, s_npad;
  __shared__ int s_pad[256];                 // top_k <= max_det_per_type; the no-threshold k is 20
  const int bj = blockIdx.x;
  const int n = (int)min(cand_count[bj], (uint32_t)cand_cap);
  int P = 1;
  while (P < n || P < top_k) P <<= 1;
  const uint64_t* __restrict__ g = cand_keys + (size_t)bj * cand_cap;
  for (int i = threadIdx.x; i < P; i += blockDim.x) s[i] = i < n ? ~g[i] : ~0ull;   // ~key ascending = key descending
  if (threadIdx.x == 0) { s_n2 = 0; s_npad = 0; }
  __syncthreads();
  bitonic_sort_asc(s, P);
  int n1 = min(top_k, n);
  if (use_thr) {
    const uint32_t tb = __float_as_uint(thr);
    int c = 0;
    for (int i = n1 + threadIdx.x; i < n; i += blockDim.x) c += ((uint32_t)((~s[i]) >> 32) >= tb) ? 1 : 0;
    if (c) atomicAdd(&s_n2, c);
  } else if (n1 < top_k && threadIdx.x < 32) {
    pad_with_zero_pixels(scoremaps + (size_t)bj * H * W, mask ? mask + (size_t)(bj / J) * H * W : nullptr, H, W, R,
                         min(top_k - n1, 256), s_pad, &s_npad);
  }
  __syncthreads();
  const int npad = s_npad;
  if (!use_thr && n1 + npad < top_k && threadIdx.x == 0) atomicOr(flags, (uint32_t)PGMP_GC_FLAG_TOO_FEW);   // map smaller than k
  int n2 = s_n2;
  if (n1 + npad + n2 > max_det) {
    if (threadIdx.x == 0) atomicOr(flags, (uint32_t)PGMP_GC_FLAG_DET_OVERFLOW);
    n2 = max(max_det - n1 - npad, 0);
  }
  const int total = min(n1 + npad + n2, max_det);
  __syncthreads();
  // second key: block << 33 | flat << 1 | zero-score pad; the pads belong to block 1 and sort among it by flat index
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t key = ~0ull;
    if (i < n1) key = (uint64_t)(~(uint32_t)(~s[i])) << 1;                  // low 32 bits of the candidate key hold ~flat
    else if (i < n1 + npad) key = ((uint64_t)(uint32_t)s_pad[i - n1] << 1) | 1ull;
    else if (i < total) key = (1ull << 33) | ((uint64_t)(~(uint32_t)(~s[i - npad])) << 1);
    s[i] = key;
  }
  __syncthreads();
  bitonic_sort_asc(s, P);
  for (int i = threadIdx.x; i < total; i += blockDim.x)
    det_idx[(size_t)bj * max_det + i] = (int32_t)(((uint32_t)(s[i] >> 1) & 0x7fffffffu) | ((uint32_t)(s[i] & 1ull) << 31));
  if (threadIdx.x == 0) { det_n1[bj] = min(n1 + npad, total); det_n2[bj] = total - min(n1 + npad, total); }
}

// K3: per-image node order: top-k blocks of types 0..J-1, then the extras blocks of types 0..J-1
// (CG.py:1180-1183); scores re-read from the scoremap (exact fp32 copy, x * mask).
__global__ void __launch_bounds__(256) layout_nodes_kernel(
    const float* __restrict__ scoremaps, const float* __restrict__ mask, int J, int H, int W, int use_thr,
    int max_det, int max_nodes, const int32_t* __restrict__ det_idx, const int32_t* __restrict__ det_n1,
    const int32_t* __restrict__ det_n2, int32_t* __restrict__ node_xyt, float* __restrict__ node_score,
    int32_t* __restrict__ node_count, uint32_t* __restrict__ flags, const AsmDev A, const int fused) {
  extern __shared__ int32_t s_start[];   // [2][J]
  __shared__ int s_total;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int off = 0;
    for (int j = 0; j < J; ++j) { s_start[j] = off; off += det_n1[b * J + j]; }
    for (int j = 0; j < J; ++j) { s_start[J + j] = off; off += det_n2[b * J + j]; }
    if (off > max_nodes) { atomicOr(flags, (uint32_t)PGMP_GC_FLAG_NODE_OVERFLOW); off = 0; }
    s_total = off;
    node_count[b] = off;
  }
  __syncthreads();
  if (s_total == 0) return;
  for (int j = 0; j < J; ++j) {
    const int bj = b * J + j;
    const int n1 = det_n1[bj], n2 = det_n2[bj];
    const float* __restrict__ map = scoremaps + (size_t)bj * H * W;
    for (int i = threadIdx.x; i < n1 + n2; i += blockDim.x) {
      const int packed = det_idx[(size_t)bj * max_det + i];
      const int flat = packed & 0x7fffffff;                 // bit 31: a zero-score pixel that pads the top-k block
      const int y = flat / W, x = flat - y * W;
      const int node = i < n1 ? s_start[j] + i : s_start[J + j] + (i - n1);
      float s = packed < 0 ? 0.f : (fused ? asm_value(A, b, j, J, H, W, y, x) : map[flat]);
      if (mask && packed >= 0) s = s * mask[(size_t)b * H * W + flat];
      if (!use_thr) s = __fadd_rn(s, 1e-10f);   // CG.py:1189
      node_xyt[(size_t)b * max_nodes + node] = pack_xyt(x, y, j);
      node_score[(size_t)b * max_nodes + node] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4: kNN-50 on integer pixel coordinates (CG.py:363-368), one warp per query node.  The k nearest
// other nodes by (squared distance asc, index asc) are found with a binary search on the squared
// distance (exact integers) plus an index-ordered tie pass; edges are recorded in a symmetric
// adjacency bit matrix, which is to_undirected + coalesce + remove_self_loops in one step.
// ------------------------------------------------------------------------------------------------
constexpr int kKnnWarps = 8;

__global__ void __launch_bounds__(kKnnWarps * 32) knn_adjacency_kernel(
    const int32_t* __restrict__ node_xyt, const int32_t* __restrict__ node_count, uint32_t* __restrict__ adj,
    int max_nodes, int k) {
  extern __shared__ int32_t s_xy[];
  const int b = blockIdx.y;
  const int N = node_count[b];
  if ((int)blockIdx.x * kKnnWarps >= N) return;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s_xy[i] = node_xyt[(size_t)b * max_nodes + i] & 0xffffff;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * kKnnWarps + (threadIdx.x >> 5);
  const int kk = min(k, N - 1);
  if (q >= N || kk <= 0) return;
  const int qx = px(s_xy[q]), qy = py(s_xy[q]);
  auto dist2 = [&](int c) {
    const int v = s_xy[c];
    const int dx = px(v) - qx, dy = py(v) - qy;
    return dx * dx + dy * dy;
  };
  // squared distances of this lane's candidates (c = lane + 32 i) cached in registers when the image has
  // at most 32 * kKnnCache nodes; the query itself and out-of-range slots get +inf
  constexpr int kKnnCache = 32;
  const bool cached = N <= 32 * kKnnCache;
  int dc[kKnnCache];
  int dmax = 0;
#pragma unroll
  for (int i = 0; i < kKnnCache; ++i) {
    const int c = lane + 32 * i;
    dc[i] = 0x7fffffff;
    if (cached && c < N && c != q) {
      dc[i] = dist2(c);
      dmax = max(dmax, dc[i]);
    }
  }
  int lo = 0, hi = 1 << 25;   // 2 * 4095^2 < 2^25
  if (cached) hi = __reduce_max_sync(kFull, dmax);
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    int cnt = 0;
    if (cached) {
#pragma unroll
      for (int i = 0; i < kKnnCache; ++i) cnt += dc[i] <= mid ? 1 : 0;
    } else {
      for (int c = lane; c < N; c += 32) cnt += (c != q && dist2(c) <= mid) ? 1 : 0;
    }
    cnt = __reduce_add_sync(kFull, cnt);
    if (cnt >= kk) hi = mid; else lo = mid + 1;
  }
  const int D = lo;
  int less = 0;
  for (int c = lane; c < N; c += 32) less += (c != q && dist2(c) < D) ? 1 : 0;
  less = __reduce_add_sync(kFull, less);
  const int need = kk - less;   // ties at distance D taken in index order
  const int words = max_nodes >> 5;
  uint32_t* __restrict__ rows = adj + (size_t)b * max_nodes * words;
  int seen_eq = 0;
  for (int base = 0; base < N; base += 32) {
    const int c = base + lane;
    const bool valid = c < N && c != q;
    const int d2 = valid ? dist2(c) : 0;
    const bool eq = valid && d2 == D;
    const uint32_t eqm = __ballot_sync(kFull, eq);
    const bool take = valid && (d2 < D || (eq && seen_eq + __popc(eqm & ((1u << lane) - 1u)) < need));
    seen_eq += __popc(eqm);
    const uint32_t tm = __ballot_sync(kFull, take);
    if (lane == 0 && tm) atomicOr(&rows[(size_t)q * words + (base >> 5)], tm);
    if (take) atomicOr(&rows[(size_t)c * words + (q >> 5)], 1u << (q & 31));
  }
}

// K5: row degrees -> per-image CSR row pointer and edge count (fully: N-1 per row, CG.py:376-381).
__global__ void __launch_bounds__(256) row_degrees_kernel(
    const uint32_t* __restrict__ adj, const int32_t* __restrict__ node_count, int max_nodes, int fully,
    int32_t* __restrict__ rowptr) {
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int b = blockIdx.x;
  const int N = node_count[b];
  const int words = max_nodes >> 5, used = (N + 31) >> 5;
  int32_t* __restrict__ rp = rowptr + (size_t)b * (max_nodes + 1);
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < N; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int deg = 0;
    if (i < N) {
      if (fully) {
        deg = N - 1;
      } else {
        const uint32_t* __restrict__ row = adj + ((size_t)b * max_nodes + i) * words;
        for (int w = 0; w < used; ++w) deg += __popc(row[w]);
      }
    }
    int incl = deg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_warp[w];
    const int carry = s_carry;
    if (i < N) rp[i] = carry + woff + incl - deg;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + woff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) rp[N] = s_carry;
}

// K6: exclusive scans over images + the counts the host reads back.
__global__ void totals_kernel(const int32_t* __restrict__ node_count, const int32_t* __restrict__ rowptr,
                              int B, int max_nodes, int64_t* __restrict__ node_base,
                              int64_t* __restrict__ edge_base, const uint32_t* __restrict__ flags,
                              int64_t* __restrict__ counts) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t n = 0, e = 0;
  for (int b = 0; b < B; ++b) {
    node_base[b] = n;
    edge_base[b] = e;
    const int nb = node_count[b];
    const int eb = rowptr[(size_t)b * (max_nodes + 1) + nb];
    counts[2 + b] = nb;
    counts[2 + B + b] = eb;
    n += nb;
    e += eb;
  }
  node_base[B] = n;
  edge_base[B] = e;
  counts[0] = n;
  counts[1] = e;
  counts[2 + 2 * B] = (int64_t)flags[0];
}

// ---------------------------------------------------------------------------------------- emit
__device__ __forceinline__ int find_image(const int64_t* __restrict__ base, int B, int64_t g) {
  int lo = 0, hi = B - 1;   // largest b with base[b] <= g
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (base[mid] <= g) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) emit_nodes_kernel(
    int64_t total_nodes, int B, int J, int H, int W, int max_nodes, const int64_t* __restrict__ node_base,
    const int32_t* __restrict__ node_xyt, const float* __restrict__ node_score, const float* __restrict__ tagmaps,
    int tag_dim, int64_t* __restrict__ joint_det, float* __restrict__ joint_scores,
    int64_t* __restrict__ batch_index, float* __restrict__ joint_tags, int32_t* __restrict__ gnode_xyt) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_nodes) return;
  const int b = find_image(node_base, B, g);
  const int n = (int)(g - node_base[b]);
  const int v = node_xyt[(size_t)b * max_nodes + n];
  const int x = px(v), y = py(v), t = pt(v);
  joint_det[g * 3 + 0] = x;
  joint_det[g * 3 + 1] = y;
  joint_det[g * 3 + 2] = t;
  joint_scores[g] = node_score[(size_t)b * max_nodes + n];
  batch_index[g] = b;
  gnode_xyt[g] = v;
  if (joint_tags) {   // CG.py:103
    const float* __restrict__ src = tagmaps + ((((size_t)b * J + t) * H + y) * W + x) * tag_dim;
    for (int d = 0; d < tag_dim; ++d) joint_tags[g * tag_dim + d] = src[d];
  }
}

// x[n, :] = features[b, :, y, x] (CG.py:265,269): one warp per node, lanes over channels.  With the
// reference's NCHW maps every element is its own 32-byte sector; channels-last maps coalesce.
__global__ void __launch_bounds__(256) gather_features_kernel(
    int64_t total_nodes, int B, const int64_t* __restrict__ node_base, const int32_t* __restrict__ gnode_xyt,
    const float* __restrict__ feat, int64_t sb, int64_t sc, int64_t sy, int64_t sx, int C,
    float* __restrict__ out) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= total_nodes) return;
  const int lane = threadIdx.x & 31;
  const int b = find_image(node_base, B, g);
  const int v = gnode_xyt[g];
  const float* __restrict__ src = feat + b * sb + py(v) * sy + px(v) * sx;
  for (int c = lane; c < C; c += 32) out[g * C + c] = __ldg(src + c * sc);
}

__global__ void __launch_bounds__(256) emit_edges_knn_kernel(
    int64_t total_nodes, int64_t total_edges, int B, int max_nodes, const int64_t* __restrict__ node_base,
    const int64_t* __restrict__ edge_base, const int32_t* __restrict__ rowptr, const uint32_t* __restrict__ adj,
    const int32_t* __restrict__ node_count, int64_t* __restrict__ edge_index) {
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= total_nodes) return;
  const int lane = threadIdx.x & 31;
  const int b = find_image(node_base, B, g);
  const int i = (int)(g - node_base[b]);
  const int words = max_nodes >> 5, used = (node_count[b] + 31) >> 5;
  const uint32_t* __restrict__ row = adj + ((size_t)b * max_nodes + i) * words;
  int64_t pos = edge_base[b] + rowptr[(size_t)b * (max_nodes + 1) + i];
  for (int w0 = 0; w0 < used; w0 += 32) {
    const int w = w0 + lane;
    uint32_t bits = w < used ? row[w] : 0u;
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += v;
    }
    int64_t p = pos + incl - cnt;
    while (bits) {
      const int bit = __ffs(bits) - 1;
      bits &= bits - 1;
      edge_index[p] = g;                                           // source
      edge_index[total_edges + p] = node_base[b] + (w << 5) + bit; // target
      ++p;
    }
    pos += __shfl_sync(kFull, incl, 31);
  }
}

__global__ void __launch_bounds__(256) emit_edges_fully_kernel(
    int64_t total_edges, int B, const int64_t* __restrict__ node_base, const int64_t* __restrict__ edge_base,
    int64_t* __restrict__ edge_index) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total_edges) return;
  const int b = find_image(edge_base, B, e);
  const int64_t n = node_base[b + 1] - node_base[b];
  const int64_t le = e - edge_base[b];
  const int64_t i = le / (n - 1), r = le - i * (n - 1);
  const int64_t j = r + (r >= i ? 1 : 0);
  edge_index[e] = node_base[b] + i;
  edge_index[total_edges + e] = node_base[b] + j;
}

// edge_attr[e] = [ (x_dst - x_src)/norm, (y_dst - y_src)/norm, two_hot(type_src, type_dst) ]  (CG.py:305-325);
// IEEE division keeps the fp32 result identical to torch's.
// one warp = 32 consecutive edges: the lanes fetch the endpoint records of their own edge once, then write the
// 32 x F block element-wise (fully coalesced), taking the records of element i's edge from lane i / F by shuffle
__global__ void __launch_bounds__(256) edge_attr_kernel(
    int64_t total_edges, int F, int J, int feats, float norm, const int64_t* __restrict__ edge_index,
    const int32_t* __restrict__ gnode_xyt, float* __restrict__ edge_attr) {
  const int lane = threadIdx.x & 31;
  const int64_t e0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (e0 >= total_edges) return;
  const int64_t e = e0 + lane;
  int s = 0, d = 0;
  if (e < total_edges) { s = gnode_xyt[edge_index[e]]; d = gnode_xyt[edge_index[total_edges + e]]; }
  const int n_el = (int)min((int64_t)32, total_edges - e0) * F;
  float* __restrict__ out = edge_attr + e0 * F;
  const int pos = (feats & PGMP_EDGE_FEAT_POSITION) ? 2 : 0;
  for (int base = 0; base < n_el; base += 32) {
    const int i = base + lane;
    const int le = min(i / F, 31);
    const int f = i - le * F;
    const int ss = __shfl_sync(kFull, s, le), dd = __shfl_sync(kFull, d, le);
    float v;
    if (f < pos) v = __fdiv_rn((float)(f == 0 ? px(dd) - px(ss) : py(dd) - py(ss)), norm);   // CG.py:305-317
    else v = (f - pos == pt(ss) || f - pos == pt(dd)) ? 1.f : 0.f;                            // two-hot, :323-325
    if (i < n_el) out[i] = v;
  }
}

int validate(const pgmp_gc_params* p) {
  if (!p) return set_error(PGMP_ERR_INVALID, "null params");
  if (p->batch <= 0 || p->num_joints <= 0 || p->num_joints > 64 || p->height <= 0 || p->width <= 0)
    return set_error(PGMP_ERR_INVALID, "bad scoremap shape [%d,%d,%d,%d]", p->batch, p->num_joints, p->height, p->width);
  if (p->height > 4096 || p->width > 4096) return set_error(PGMP_ERR_INVALID, "maps larger than 4096 px are not supported");
  if (p->pool_kernel % 2 != 1 || p->pool_kernel < 1 || p->pool_kernel > 9)
    return set_error(PGMP_ERR_INVALID, "pool_kernel must be odd and <= 9 (Utils.py:16), got %d", p->pool_kernel);
  if (p->use_threshold && !(p->threshold > 0.f))
    return set_error(PGMP_ERR_INVALID, "DETECT_THRESHOLD must be > 0 (candidates are positive maxima), got %g", p->threshold);
  if (p->top_k <= 0 || p->top_k > p->max_det_per_type)
    return set_error(PGMP_ERR_INVALID, "top_k %d must be in [1, max_det_per_type %d]", p->top_k, p->max_det_per_type);
  if (p->max_nodes <= 0 || p->max_nodes % 32 != 0 || p->max_nodes > 16384)
    return set_error(PGMP_ERR_INVALID, "max_nodes must be a positive multiple of 32 <= 16384, got %d", p->max_nodes);
  if (p->cand_capacity < p->top_k || p->cand_capacity > 16384)
    return set_error(PGMP_ERR_INVALID, "cand_capacity %d must be in [top_k, 16384]", p->cand_capacity);
  if (p->graph_type != PGMP_GRAPH_KNN && p->graph_type != PGMP_GRAPH_FULLY)
    return set_error(PGMP_ERR_INVALID, "graph_type %d", p->graph_type);
  if (p->edge_features == 0 || (p->edge_features & ~3)) return set_error(PGMP_ERR_INVALID, "edge_features %d", p->edge_features);
  if (!p->workspace) return set_error(PGMP_ERR_INVALID, "null device pointer");
  return PGMP_OK;
}

int nms_strip_rows(const pgmp_gc_params& p, int xtiles) {
  // rows per CTA: whole maps give the sharpest running cut (measured: 0.175 ms against 0.198 ms with half maps at
  // 32 x 17 x 512 x 512); split maps into strips only as far as it takes to fill the 3 CTA slots of every SM
  const char* env = getenv("PGMP_NMS_STRIP_ROWS");   // tuning / test hook
  const int forced = env ? atoi(env) : 0;
  int rows;
  if (forced > 0) {
    rows = forced;
  } else {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t maps = (int64_t)p.batch * p.num_joints * xtiles;
    int64_t strips = (3 * sms + maps - 1) / maps;
    if (strips > ceil_div(p.height, 32)) strips = ceil_div(p.height, 32);
    if (strips < 1) strips = 1;
    rows = ceil_div(p.height, (int)strips);
  }
  rows = round_up(rows, kNmsStageRows);
  return rows < kNmsStageRows ? kNmsStageRows : rows;
}

template <int R>
int launch_nms(const pgmp_gc_params& p, const GcWorkspace& w, cudaStream_t st, const AsmDev* fuse) {
  const bool vec = (p.width % 4 == 0) && (fuse || reinterpret_cast<uintptr_t>(p.scoremaps) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(p.mask) % 16 == 0);
  const int chunks = ceil_div(p.width, 4);
  const int xtiles = ceil_div(chunks, 256);
  const int threads = round_up(ceil_div(chunks, xtiles), 32);     // column threads of a tile (32 per 128-column band)
  const int bands = threads / 32;
  const int rw = bands <= 4 ? 2 : 1;                              // row groups: consumer warps = bands x rw <= 8 (+ loader and selector warps)
  NmsArgs a;
  if (fuse) {
    if (!vec || xtiles != 1) return set_error(PGMP_ERR_INVALID, "fused assembly needs width %% 4 == 0, width <= 1024 and a 16-byte aligned mask");
    a.asmb = *fuse;
  } else {
    a.asmb = AsmDev{};
  }
  a.scoremaps = fuse ? nullptr : p.scoremaps; a.mask = p.mask; a.J = p.num_joints; a.H = p.height; a.W = p.width; a.xtiles = xtiles;
  a.rows_per_cta = nms_strip_rows(p, xtiles);
  // ring row pitch (floats): the map's width when one tile spans it (a stage is one contiguous bulk copy), else the
  // tile's columns + 4 either side; unaligned maps: rounded up and zero-filled by the copying warp
  const int cols = xtiles == 1 ? p.width : 4 * threads + 8;
  a.pitch = vec ? cols : round_up(cols, 4) + 4;
  a.top_k = p.top_k; a.use_thr = p.use_threshold; a.thr = p.threshold;
  a.cand_keys = w.cand_keys; a.cand_count = w.cand_count; a.cand_cap = p.cand_capacity; a.flags = w.flags;
  a.src_rows = 0; a.exact2x = 0;
  if (fuse) {
    if (fuse->w % 4) return set_error(PGMP_ERR_INVALID, "fused assembly needs a stage-1 width that is a multiple of 4 (bulk copies of whole rows)");
    a.src_rows = (int)((double)fuse->h / p.height * kNmsStageRows) + 3;
    if (a.src_rows > fuse->h) a.src_rows = fuse->h;
    a.exact2x = (p.height == 2 * fuse->h && p.width == 2 * fuse->w && fuse->w >= 2) ? 1 : 0;
  }
  const size_t smem = (size_t)kNmsRingRows * a.pitch * 4 + (size_t)(kNmsListCap + kNmsKeepCap) * 8 + 256 * 4 + 3 * kNmsStages * 8 + 64 +
                      (fuse ? (size_t)a.pitch * 16 + (size_t)2 * 2 * a.src_rows * fuse->w * 4 : 0);
  if (smem > 227 * 1024)
    return set_error(PGMP_ERR_INVALID, "NMS row ring of %zu bytes does not fit in shared memory (fused assembly of %d x %d maps from %d x %d: "
                     "run pgmp_gc_assemble_scoremaps + pgmp_gc_detect)", smem, p.height, p.width, fuse ? fuse->h : 0, fuse ? fuse->w : 0);
  const dim3 grid(ceil_div(p.height, a.rows_per_cta) * xtiles, p.num_joints, p.batch);
  const int block = 32 * (bands * rw + 1 + (fuse ? kNmsFuseLoaders : 1));
#define PGMP_NMS_LAUNCH3(RPW, V, M, F)                                                                                             \
  do {                                                                                                                             \
    PGMP_CUDA(cudaFuncSetAttribute(nms_candidates_kernel<R, RPW, V, M, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    PGMP_LAUNCH((nms_candidates_kernel<R, RPW, V, M, F>), grid, block, smem, st, a);                                               \
  } while (0)
#define PGMP_NMS_LAUNCH2(RPW, V, M) PGMP_NMS_LAUNCH3(RPW, V, M, false)
  if (fuse) {
    if (rw == 2) { if (p.mask) PGMP_NMS_LAUNCH3(4, true, true, true); else PGMP_NMS_LAUNCH3(4, true, false, true); }
    else { if (p.mask) PGMP_NMS_LAUNCH3(8, true, true, true); else PGMP_NMS_LAUNCH3(8, true, false, true); }
    return PGMP_OK;
  }
#define PGMP_NMS_LAUNCH(V, M)                                             \
  do {                                                                    \
    if (rw == 2) PGMP_NMS_LAUNCH2(4, V, M); else PGMP_NMS_LAUNCH2(8, V, M); \
  } while (0)
  if (vec) { if (p.mask) PGMP_NMS_LAUNCH(true, true); else PGMP_NMS_LAUNCH(true, false); }
  else { if (p.mask) PGMP_NMS_LAUNCH(false, true); else PGMP_NMS_LAUNCH(false, false); }
#undef PGMP_NMS_LAUNCH
#undef PGMP_NMS_LAUNCH2
#undef PGMP_NMS_LAUNCH3
  return PGMP_OK;
}

}  // namespace
}  // namespace pgmp

using namespace pgmp;

extern "C" uint64_t pgmp_gc_workspace_bytes(const pgmp_gc_params* p) {
  if (!p) return 0;
  pgmp_gc_params q = *p;
  q.workspace = nullptr;
  return carve(q).bytes;
}

static int gc_detect(const pgmp_gc_params* p_in, const AsmDev* fuse, int64_t* counts, pgmp_stream_t stream) {
  pgmp_gc_params q;
  if (p_in && fuse) {           // the map the later kernels may read: the materialised one, if any
    q = *p_in;
    q.scoremaps = fuse->out;
    p_in = &q;
  }
  const pgmp_gc_params* p = p_in;
  int rc = validate(p);
  if (rc != PGMP_OK) return rc;
  if (!counts) return set_error(PGMP_ERR_INVALID, "null counts");
  if (!fuse && !p->scoremaps) return set_error(PGMP_ERR_INVALID, "null device pointer");
  const bool lazy = fuse && !fuse->out;                   // scores re-evaluated at the detections
  if (lazy && !p->use_threshold)
    return set_error(PGMP_ERR_INVALID, "the no-threshold path pads with zero-score pixels of the map: pass scoremaps_out");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GcWorkspace w = carve(*p);
  if (w.bytes > p->workspace_bytes)
    return set_error(PGMP_ERR_INVALID, "workspace too small: %llu < %llu", (unsigned long long)p->workspace_bytes,
                     (unsigned long long)w.bytes);
  const int B = p->batch, J = p->num_joints;
  PGMP_CUDA(cudaMemsetAsync(w.cand_count, 0, sizeof(uint32_t) * B * J, st));
  PGMP_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(uint32_t), st));
  switch (p->pool_kernel / 2) {
    case 0: rc = launch_nms<0>(*p, w, st, fuse); break;
    case 1: rc = launch_nms<1>(*p, w, st, fuse); break;
    case 2: rc = launch_nms<2>(*p, w, st, fuse); break;
    case 3: rc = launch_nms<3>(*p, w, st, fuse); break;
    default: rc = launch_nms<4>(*p, w, st, fuse); break;
  }
  if (rc != PGMP_OK) return rc;
  int P = 1;
  while (P < p->cand_capacity || P < p->top_k) P <<= 1;
  const size_t sel_smem = sizeof(uint64_t) * P;
  if (sel_smem > 48 * 1024)
    PGMP_CUDA(cudaFuncSetAttribute(select_detections_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  PGMP_LAUNCH(select_detections_kernel, B * J, 256, sel_smem, st, w.cand_keys, w.cand_count, p->cand_capacity, p->top_k,
              p->use_threshold, p->threshold, p->max_det_per_type, p->scoremaps, p->mask, J, p->height, p->width,
              p->pool_kernel / 2, w.det_idx, w.det_n1, w.det_n2, w.flags);
  PGMP_LAUNCH(layout_nodes_kernel, B, 256, sizeof(int32_t) * 2 * J, st, p->scoremaps, p->mask, J, p->height, p->width,
              p->use_threshold, p->max_det_per_type, p->max_nodes, w.det_idx, w.det_n1, w.det_n2, w.node_xyt,
              w.node_score, w.node_count, w.flags, lazy ? *fuse : AsmDev{}, lazy ? 1 : 0);
  const int fully = p->graph_type == PGMP_GRAPH_FULLY;
  if (!fully) {
    PGMP_CUDA(cudaMemsetAsync(w.adj, 0, sizeof(uint32_t) * (size_t)B * p->max_nodes * (p->max_nodes / 32), st));
    const size_t knn_smem = sizeof(int32_t) * p->max_nodes;
    if (knn_smem > 48 * 1024)
      PGMP_CUDA(cudaFuncSetAttribute(knn_adjacency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)knn_smem));
    PGMP_LAUNCH(knn_adjacency_kernel, dim3(ceil_div(p->max_nodes, kKnnWarps), B), kKnnWarps * 32, knn_smem, st,
                w.node_xyt, w.node_count, w.adj, p->max_nodes, p->knn_k);
  }
  PGMP_LAUNCH(row_degrees_kernel, B, 256, 0, st, w.adj, w.node_count, p->max_nodes, fully, w.rowptr);
  PGMP_LAUNCH(totals_kernel, 1, 32, 0, st, w.node_count, w.rowptr, B, p->max_nodes, w.node_base, w.edge_base, w.flags,
              counts);
  return PGMP_OK;
}

extern "C" int pgmp_gc_detect(const pgmp_gc_params* p, int64_t* counts, pgmp_stream_t stream) {
  return gc_detect(p, nullptr, counts, stream);
}

extern "C" int pgmp_gc_detect_fused(const pgmp_gc_params* p, const pgmp_gc_assembly* a, int64_t* counts, pgmp_stream_t stream) {
  if (!p || !a) return set_error(PGMP_ERR_INVALID, "null params");
  if (a->mode != PGMP_ASSEMBLE_AVG && a->mode != PGMP_ASSEMBLE_SMALL) return set_error(PGMP_ERR_INVALID, "mode %d", a->mode);
  if (a->n_terms != 1 && a->n_terms != 2) return set_error(PGMP_ERR_INVALID, "n_terms %d", a->n_terms);
  if (a->h <= 0 || a->w <= 0 || a->channels1 < p->num_joints || p->num_joints > 32)
    return set_error(PGMP_ERR_INVALID, "bad stage sizes (channels1 %d, %d x %d, %d joints)", a->channels1, a->h, a->w, p->num_joints);
  AsmDev d{};
  for (int t = 0; t < a->n_terms; ++t) {
    if (!a->stage1[t] || (a->mode == PGMP_ASSEMBLE_AVG && !a->stage2[t])) return set_error(PGMP_ERR_INVALID, "null stage pointer (term %d)", t);
    if (reinterpret_cast<uintptr_t>(a->stage2[t]) % 16) return set_error(PGMP_ERR_INVALID, "stage2 must be 16-byte aligned");
    d.s1[t] = a->stage1[t]; d.s2[t] = a->stage2[t];
  }
  if (reinterpret_cast<uintptr_t>(a->scoremaps_out) % 16) return set_error(PGMP_ERR_INVALID, "scoremaps_out must be 16-byte aligned");
  d.C1 = a->channels1; d.h = a->h; d.w = a->w; d.mode = a->mode; d.terms = a->n_terms;
  d.scale_y = (float)a->h / (float)p->height; d.scale_x = (float)a->w / (float)p->width;
  for (int j = 0; j < 32; ++j) {
    d.flip[j] = j;
    if (a->n_terms == 2 && j < p->num_joints) {
      if (a->flip_index[j] < 0 || a->flip_index[j] >= p->num_joints) return set_error(PGMP_ERR_INVALID, "flip_index[%d] = %d", j, a->flip_index[j]);
      d.flip[j] = a->flip_index[j];
    }
  }
  d.out = a->scoremaps_out;
  return gc_detect(p, &d, counts, stream);
}

extern "C" int pgmp_gc_emit(const pgmp_gc_params* p, const pgmp_gc_outputs* o, pgmp_stream_t stream) {
  int rc = validate(p);
  if (rc != PGMP_OK) return rc;
  if (!o) return set_error(PGMP_ERR_INVALID, "null outputs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GcWorkspace w = carve(*p);
  const int B = p->batch, J = p->num_joints;
  const int64_t N = o->total_nodes, E = o->total_edges;
  if (N < 0 || E < 0 || N > (int64_t)B * p->max_nodes) return set_error(PGMP_ERR_INVALID, "bad totals N=%lld E=%lld", (long long)N, (long long)E);
  if (N == 0) return PGMP_OK;
  if (!o->joint_det || !o->joint_scores || !o->batch_index) return set_error(PGMP_ERR_INVALID, "null node outputs");
  if (o->joint_tags && (!o->tagmaps || o->tag_dim <= 0)) return set_error(PGMP_ERR_INVALID, "joint_tags without tagmaps");
  PGMP_LAUNCH(emit_nodes_kernel, (unsigned)ceil_div<int64_t>(N, 256), 256, 0, st, N, B, J, p->height, p->width,
              p->max_nodes, w.node_base, w.node_xyt, w.node_score, o->tagmaps, o->tag_dim, o->joint_det,
              o->joint_scores, o->batch_index, o->joint_tags, w.gnode_xyt);
  if (o->x) {
    if (!o->features || o->channels <= 0) return set_error(PGMP_ERR_INVALID, "x without features");
    PGMP_LAUNCH(gather_features_kernel, (unsigned)ceil_div<int64_t>(N * 32, 256), 256, 0, st, N, B, w.node_base,
                w.gnode_xyt, o->features, o->feat_stride_b, o->feat_stride_c, o->feat_stride_y, o->feat_stride_x,
                o->channels, o->x);
  }
  if (E == 0) return PGMP_OK;
  if (!o->edge_index) return set_error(PGMP_ERR_INVALID, "null edge_index");
  if (p->graph_type == PGMP_GRAPH_FULLY) {
    PGMP_LAUNCH(emit_edges_fully_kernel, (unsigned)ceil_div<int64_t>(E, 256), 256, 0, st, E, B, w.node_base, w.edge_base,
                o->edge_index);
  } else {
    PGMP_LAUNCH(emit_edges_knn_kernel, (unsigned)ceil_div<int64_t>(N * 32, 256), 256, 0, st, N, E, B, p->max_nodes,
                w.node_base, w.edge_base, w.rowptr, w.adj, w.node_count, o->edge_index);
  }
  if (o->edge_attr) {
    const int F = ((p->edge_features & PGMP_EDGE_FEAT_POSITION) ? 2 : 0) + ((p->edge_features & PGMP_EDGE_FEAT_TYPE) ? J : 0);
    PGMP_LAUNCH(edge_attr_kernel, (unsigned)ceil_div<int64_t>(E, 256), 256, 0, st, E, F, J, p->edge_features,
                p->norm_factor, o->edge_index, w.gnode_xyt, o->edge_attr);
  }
  return PGMP_OK;
}
