// Graph constructor kernels (sm_100a): heatmap NMS + per-joint top-k / threshold candidates,
// candidate graph (symmetric kNN-50 or fully connected) as a CSR / COO edge index, node-feature
// gather and edge attributes.  Semantics: SURVEY.md Appendix B.1-B.3; reference
// src/graph_constructor/ConstructGraph.py (CG.py) and src/Utils/Utils.py:15-20.
//
// Pipeline (all stream-ordered, no host sync inside):
//   detect: nms_candidates -> select_detections -> layout_nodes -> [knn_adjacency] -> row_degrees
//           -> totals           (host reads the counts once)
//   emit:   emit_nodes, gather_features, emit_edges_{knn,fully}, edge_attr
#include "common.cuh"

namespace pgmp {
namespace {

constexpr int kNmsRows = 64;        // output rows per CTA strip
constexpr int kCtaCandCap = 2048;   // per-CTA shared-memory candidate list
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
// K1: max-pool NMS (Utils.py:15-20) fused with candidate extraction.
// One CTA = one (image, joint, 32-row strip, <=1024-column tile), 4 columns per thread.  Rows stream
// top to bottom through registers: the horizontal (2R+1)-max is taken across lanes with warp
// shuffles, the vertical one over a register ring of row maxima, so every heatmap element is
// loaded once per strip (+ R halo rows above and below).  A pixel is a candidate iff it is
// positive and equals its window maximum; zero padding is then equivalent to the reference's -inf
// padding.  Candidates go to a shared-memory list; at the end of the strip the CTA keeps only what
// can matter globally -- its own top_k scores and everything >= threshold -- and appends those to
// the (image, joint) list in global memory with one atomic per warp chunk.
// ------------------------------------------------------------------------------------------------
// cold path: the CTA's shared list is full -> unfiltered straight to the global list
__device__ __noinline__ void spill_candidate(uint64_t key, uint64_t* __restrict__ gkeys, uint32_t* __restrict__ gcount,
                                             int cand_cap, uint32_t* __restrict__ flags) {
  const uint32_t g = atomicAdd(gcount, 1u);
  if (g < (uint32_t)cand_cap) gkeys[g] = key; else atomicOr(flags, (uint32_t)PGMP_GC_FLAG_CAND_OVERFLOW);
}

// Append this thread's candidates (up to 4, usually none: a warp-row of 128 noisy heatmap pixels holds about five
// local maxima) to the CTA's shared list: one shared atomic per candidate hands out the slot -- the order inside
// the list is irrelevant, keys are unique -- and only lanes that own a candidate execute anything.  A full list
// spills unfiltered to the global list.  Key = score bits << 32 | ~flat index: one 64-bit compare orders by
// score desc / index asc.
__device__ __forceinline__ void emit_candidates(const float (&sc)[4], uint32_t flat0, uint32_t s_keys_a, uint32_t s_cnt_a,
                                                uint64_t* __restrict__ gkeys, uint32_t* __restrict__ gcount, int cand_cap,
                                                uint32_t* __restrict__ flags) {
  // a loop over the thread's candidates, not four predicated blocks: its trip count across the warp is the largest
  // number of candidates any lane holds (almost always 0 or 1).  The list is addressed in the shared state space
  // (32-bit addresses formed once per kernel) and the slot comes from a plain per-lane shared atomic: the
  // compiler's warp-aggregated form (vote + leader election + two popcounts + shuffle) costs more instructions
  // than the handful of same-address conflicts it saves.
  uint32_t vm = (sc[0] > 0.f ? 1u : 0u) | (sc[1] > 0.f ? 2u : 0u) | (sc[2] > 0.f ? 4u : 0u) | (sc[3] > 0.f ? 8u : 0u);
  while (vm) {
    const int i = __ffs(vm) - 1;
    vm &= vm - 1;
    const float v = i == 0 ? sc[0] : (i == 1 ? sc[1] : (i == 2 ? sc[2] : sc[3]));
    const uint32_t hi = __float_as_uint(v), lo = ~(flat0 + (uint32_t)i);
    uint32_t pos;
    asm volatile("atom.shared.inc.u32 %0, [%1], 0x7fffffff;" : "=r"(pos) : "r"(s_cnt_a) : "memory");   // inc with a bound never reached: ptxas warp-aggregates add (and inc 0xffffffff)
    if (pos < kCtaCandCap) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_keys_a + pos * 8u), "r"(lo), "r"(hi) : "memory");
    else spill_candidate(((uint64_t)hi << 32) | (uint64_t)lo, gkeys, gcount, cand_cap, flags);
  }
}

// End of a strip: keep only what can matter globally -- the CTA's own top_k scores (32-bit radix select in
// shared memory, 4 passes of 8 bits) and everything >= threshold -- and append it to the (image, joint) list
// with one atomic per warp chunk.
__device__ __forceinline__ void flush_candidates(const uint64_t* s_keys, uint32_t* s_hist, const uint32_t* s_cnt_p,
                                                 uint32_t* s_prefix_p, uint32_t* s_remaining_p, int top_k, int use_thr,
                                                 float thr, uint64_t* __restrict__ gkeys, uint32_t* __restrict__ gcount,
                                                 int cand_cap, uint32_t* __restrict__ flags) {
  const int t = threadIdx.x, lane = t & 31;
  __syncthreads();
  const uint32_t n = min(*s_cnt_p, (uint32_t)kCtaCandCap);
  if (n == 0) return;
  uint32_t cut = 0;  // keep entries with score bits >= cut
  if (n > (uint32_t)top_k) {
    // the CTA's top_k-th largest score: 32-bit radix select, 4 passes of 8 bits; the 256-bin suffix scan of a
    // pass is done by one warp (8 bins per lane)
    if (t == 0) { *s_prefix_p = 0; *s_remaining_p = (uint32_t)top_k; }
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = t; i < 256; i += blockDim.x) s_hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = *s_prefix_p;
      const uint32_t himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
      for (uint32_t i = t; i < n; i += blockDim.x) {
        const uint32_t sc = (uint32_t)(s_keys[i] >> 32);
        if ((sc & himask) == prefix) atomicAdd(&s_hist[(sc >> shift) & 0xff], 1u);
      }
      __syncthreads();
      if (t < 32) {
        const uint32_t rem = *s_remaining_p;
        uint32_t mine = 0;                     // lane L owns bins [8 (31 - L), 8 (31 - L) + 8): lane 0 = the highest bins
        const int b0 = 8 * (31 - lane);
#pragma unroll
        for (int q = 0; q < 8; ++q) mine += s_hist[b0 + q];
        uint32_t incl = mine;                  // inclusive scan from the highest bins down
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += v;
        }
        const uint32_t hit = __ballot_sync(kFull, incl >= rem);   // always non-empty: the matching entries number >= rem
        const int owner = __ffs(hit) - 1;
        if (lane == owner) {
          uint32_t r2 = rem - (incl - mine);
          int d = b0 + 7;
          for (; d > b0; --d) {
            if (s_hist[d] >= r2) break;
            r2 -= s_hist[d];
          }
          *s_prefix_p = prefix | ((uint32_t)d << shift);
          *s_remaining_p = r2;
        }
      }
      __syncthreads();
    }
    cut = *s_prefix_p;
  }
  if (use_thr) cut = min(cut, __float_as_uint(thr));   // positive floats order like their bit patterns
  for (uint32_t base = 0; base < n; base += blockDim.x) {
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

struct RowRegs {
  float4 v, l, r;  // own 4 columns, left / right neighbour's 4 columns (only lanes 0 / 31 load l / r)
};

// per-thread load plan of a strip: column offsets and which of the three 4-column chunks exist / are needed
struct LoadPlan {
  int off_v, off_l, off_r;      // element offsets inside a row
  bool has_v, has_l, has_r;
  int n_v, n_l, n_r;            // valid elements of each chunk (scalar path)
};

template <bool VEC>
__device__ __forceinline__ float4 load_chunk(const float* __restrict__ p, bool has, int n) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!has) return v;
  if (VEC) return __ldg(reinterpret_cast<const float4*>(p));
  v.x = __ldg(p);
  if (n > 1) v.y = __ldg(p + 1);
  if (n > 2) v.z = __ldg(p + 2);
  if (n > 3) v.w = __ldg(p + 3);
  return v;
}

// one address per row and thread: the own chunk at map + off_v + yy * W, the neighbour chunks 4 elements either side
template <bool VEC>
__device__ __forceinline__ RowRegs load_row(const float* __restrict__ map, int yy, int H, int W, const LoadPlan& lp) {
  RowRegs q;
  const bool ok = (unsigned)yy < (unsigned)H;
  const float* __restrict__ p = map + lp.off_v + (ok ? yy * W : 0);   // a map has at most 4096 x 4096 elements: 32-bit offsets
  q.v = load_chunk<VEC>(p, ok && lp.has_v, lp.n_v);
  q.l = load_chunk<VEC>(p - 4, ok && lp.has_l, lp.n_l);
  q.r = load_chunk<VEC>(p + 4, ok && lp.has_r, lp.n_r);
  return q;
}

// Register state of a strip: rings of the last K = 2R+1 rows (raw values and horizontal maxima).  The ring
// slot of a row is its phase PH = (row - first row) mod K, a template parameter, so the rings never shift.
template <int R>
struct NmsRings {
  float hm[2 * R + 1][4];
  float raw[2 * R + 1][4];
};

struct NmsCtx {
  LoadPlan lp;
  const float* map; const float* mk;
  int H, W, tt, lane, y0, y_end;
  uint32_t s_keys_a, s_cnt_a;   // shared-state-space addresses of the CTA's candidate list and its counter
  uint64_t* gkeys; uint32_t* gcount; int cand_cap; uint32_t* flags;
};

template <int R, int PH, bool VEC, bool MASK>
__device__ __forceinline__ void nms_row_step(NmsRings<R>& rg, const NmsCtx& c, const RowRegs& cur, int yy) {
  constexpr int K = 2 * R + 1;
  // only the R columns next to this thread's chunk are needed from each neighbour
  float4 l = cur.l, r = cur.r;
  {
    const float lw = __shfl_up_sync(kFull, cur.v.w, 1), rx = __shfl_down_sync(kFull, cur.v.x, 1);
    if (c.lane != 0) l.w = lw;
    if (c.lane != 31) r.x = rx;
    if (R >= 2) {
      const float lz = __shfl_up_sync(kFull, cur.v.z, 1), ry = __shfl_down_sync(kFull, cur.v.y, 1);
      if (c.lane != 0) l.z = lz;
      if (c.lane != 31) r.y = ry;
    }
    if (R >= 3) {
      const float ly = __shfl_up_sync(kFull, cur.v.y, 1), rz = __shfl_down_sync(kFull, cur.v.z, 1);
      if (c.lane != 0) l.y = ly;
      if (c.lane != 31) r.z = rz;
    }
    if (R >= 4) {
      const float lx = __shfl_up_sync(kFull, cur.v.x, 1), rw = __shfl_down_sync(kFull, cur.v.w, 1);
      if (c.lane != 0) l.x = lx;
      if (c.lane != 31) r.w = rw;
    }
  }
  const float ext[12] = {l.x, l.y, l.z, l.w, cur.v.x, cur.v.y, cur.v.z, cur.v.w, r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float m = ext[4 + q - R];
#pragma unroll
    for (int d = -R + 1; d <= R; ++d) m = fmaxf(m, ext[4 + q + d]);
    rg.hm[PH][q] = m;
    rg.raw[PH][q] = ext[4 + q];
  }
  const int yc = yy - R;
  if (yc >= c.y0 && yc < c.y_end) {   // uniform for the CTA
    constexpr int CS = (PH + K - R) % K;   // ring slot of the centre row
    float sc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x = rg.raw[CS][q];
      float m = rg.hm[0][q];
#pragma unroll
      for (int i = 1; i < K; ++i) m = fmaxf(m, rg.hm[i][q]);
      sc[q] = 0.f;
      // (out-of-image columns hold zeros, so x > 0 already excludes them)
      if (x > 0.f && x == m) sc[q] = MASK ? x * __ldg(c.mk + (yc * c.W + 4 * c.tt + q)) : x;   // CG.py:1163-1165
    }
    emit_candidates(sc, (uint32_t)(yc * c.W + 4 * c.tt), c.s_keys_a, c.s_cnt_a, c.gkeys, c.gcount, c.cand_cap, c.flags);
  }
}

// K rows per outer iteration, one statically-phased step each
template <int R, int PH, bool VEC, bool MASK>
__device__ __forceinline__ void nms_rows(NmsRings<R>& rg, const NmsCtx& c, RowRegs& nxt, int yy, int y_last) {
  constexpr int K = 2 * R + 1;
  if constexpr (PH < K) {
    if (yy < y_last) {
      const RowRegs cur = nxt;
      nxt = load_row<VEC>(c.map, yy + 1 < y_last ? yy + 1 : -1, c.H, c.W, c.lp);   // prefetch
      nms_row_step<R, PH, VEC, MASK>(rg, c, cur, yy);
      nms_rows<R, PH + 1, VEC, MASK>(rg, c, nxt, yy + 1, y_last);
    }
  }
}

template <int R, bool VEC>
__global__ void __launch_bounds__(256) nms_candidates_kernel(
    const float* __restrict__ scoremaps, const float* __restrict__ mask, int J, int H, int W, int xtiles, int top_k,
    int use_thr, float thr, uint64_t* __restrict__ cand_keys, uint32_t* __restrict__ cand_count, int cand_cap,
    uint32_t* __restrict__ flags) {
  constexpr int K = 2 * R + 1;
  __shared__ uint64_t s_keys[kCtaCandCap];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_cnt;
  __shared__ uint32_t s_prefix, s_remaining;

  const int b = blockIdx.z, j = blockIdx.y, y0 = (blockIdx.x / xtiles) * kNmsRows;
  const int t = threadIdx.x;
  const int bj = b * J + j;
  NmsCtx c;
  c.map = scoremaps + (size_t)bj * H * W;
  c.mk = mask ? mask + (size_t)b * H * W : nullptr;
  c.H = H; c.W = W; c.lane = t & 31;
  c.tt = (blockIdx.x % xtiles) * blockDim.x + t;   // absolute 4-column chunk index
  c.y0 = y0; c.y_end = min(y0 + kNmsRows, H);
  {
    const int chunks = (W + 3) >> 2;
    LoadPlan& lp = c.lp;
    lp.off_v = 4 * c.tt; lp.off_l = 4 * (c.tt - 1); lp.off_r = 4 * (c.tt + 1);
    lp.has_v = c.tt < chunks;
    lp.has_l = c.lane == 0 && c.tt > 0 && c.tt - 1 < chunks;
    lp.has_r = c.lane == 31 && c.tt + 1 < chunks;
    lp.n_v = min(4, W - lp.off_v); lp.n_l = min(4, W - lp.off_l); lp.n_r = min(4, W - lp.off_r);
  }
  c.s_keys_a = (uint32_t)__cvta_generic_to_shared(s_keys); c.s_cnt_a = (uint32_t)__cvta_generic_to_shared(&s_cnt);
  c.gkeys = cand_keys + (size_t)bj * cand_cap; c.gcount = &cand_count[bj]; c.cand_cap = cand_cap; c.flags = flags;
  if (t == 0) s_cnt = 0;
  __syncthreads();

  NmsRings<R> rg;
#pragma unroll
  for (int i = 0; i < K; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) { rg.hm[i][q] = 0.f; rg.raw[i][q] = 0.f; }
  const int y_last = c.y_end + R;
  RowRegs nxt = load_row<VEC>(c.map, y0 - R, H, W, c.lp);
  if (mask) {
    for (int yy = y0 - R; yy < y_last; yy += K) nms_rows<R, 0, VEC, true>(rg, c, nxt, yy, y_last);
  } else {
    for (int yy = y0 - R; yy < y_last; yy += K) nms_rows<R, 0, VEC, false>(rg, c, nxt, yy, y_last);
  }
  flush_candidates(s_keys, s_hist, &s_cnt, &s_prefix, &s_remaining, top_k, use_thr, thr, c.gkeys, c.gcount, cand_cap, flags);
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

__global__ void __launch_bounds__(256) select_detections_kernel(
    const uint64_t* __restrict__ cand_keys, const uint32_t* __restrict__ cand_count, int cand_cap, int top_k,
    int use_thr, float thr, int max_det, int32_t* __restrict__ det_idx, int32_t* __restrict__ det_n1,
    int32_t* __restrict__ det_n2, uint32_t* __restrict__ flags) {
  extern __shared__ uint64_t s[];
  __shared__ int s_n2;
  const int bj = blockIdx.x;
  const int n = (int)min(cand_count[bj], (uint32_t)cand_cap);
  int P = 1;
  while (P < n) P <<= 1;
  const uint64_t* __restrict__ g = cand_keys + (size_t)bj * cand_cap;
  for (int i = threadIdx.x; i < P; i += blockDim.x) s[i] = i < n ? ~g[i] : ~0ull;   // ~key ascending = key descending
  if (threadIdx.x == 0) s_n2 = 0;
  __syncthreads();
  bitonic_sort_asc(s, P);
  const int n1 = min(top_k, n);
  if (use_thr) {
    const uint32_t tb = __float_as_uint(thr);
    int c = 0;
    for (int i = n1 + threadIdx.x; i < n; i += blockDim.x) c += ((uint32_t)((~s[i]) >> 32) >= tb) ? 1 : 0;
    if (c) atomicAdd(&s_n2, c);
  } else if (n1 < top_k && threadIdx.x == 0) {
    atomicOr(flags, (uint32_t)PGMP_GC_FLAG_TOO_FEW);
  }
  __syncthreads();
  int n2 = s_n2;
  if (n1 + n2 > max_det) {
    if (threadIdx.x == 0) atomicOr(flags, (uint32_t)PGMP_GC_FLAG_DET_OVERFLOW);
    n2 = max_det - n1;
  }
  const int total = n1 + n2;
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const uint32_t flat = ~(uint32_t)(~s[i]);   // low 32 bits of the key hold ~flat
    s[i] = i < total ? (((uint64_t)(i >= n1 ? 1u : 0u) << 32) | flat) : ~0ull;
  }
  __syncthreads();
  bitonic_sort_asc(s, P);
  for (int i = threadIdx.x; i < total; i += blockDim.x) det_idx[(size_t)bj * max_det + i] = (int32_t)(uint32_t)s[i];
  if (threadIdx.x == 0) { det_n1[bj] = n1; det_n2[bj] = n2; }
}

// K3: per-image node order: top-k blocks of types 0..J-1, then the extras blocks of types 0..J-1
// (CG.py:1180-1183); scores re-read from the scoremap (exact fp32 copy, x * mask).
__global__ void __launch_bounds__(256) layout_nodes_kernel(
    const float* __restrict__ scoremaps, const float* __restrict__ mask, int J, int H, int W, int use_thr,
    int max_det, int max_nodes, const int32_t* __restrict__ det_idx, const int32_t* __restrict__ det_n1,
    const int32_t* __restrict__ det_n2, int32_t* __restrict__ node_xyt, float* __restrict__ node_score,
    int32_t* __restrict__ node_count, uint32_t* __restrict__ flags) {
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
      const int flat = det_idx[(size_t)bj * max_det + i];
      const int y = flat / W, x = flat - y * W;
      const int node = i < n1 ? s_start[j] + i : s_start[J + j] + (i - n1);
      float s = map[flat];
      if (mask) s = s * mask[(size_t)b * H * W + flat];
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
  if (!p->scoremaps || !p->workspace) return set_error(PGMP_ERR_INVALID, "null device pointer");
  return PGMP_OK;
}

template <int R>
int launch_nms(const pgmp_gc_params& p, const GcWorkspace& w, cudaStream_t st) {
  const bool vec = (p.width % 4 == 0) && (reinterpret_cast<uintptr_t>(p.scoremaps) % 16 == 0);
  const int chunks = ceil_div(p.width, 4);
  const int threads = min(round_up(chunks, 32), 256);
  const int xtiles = ceil_div(chunks, threads);
  const dim3 grid(ceil_div(p.height, kNmsRows) * xtiles, p.num_joints, p.batch);
  if (vec) {
    PGMP_LAUNCH((nms_candidates_kernel<R, true>), grid, threads, 0, st, p.scoremaps, p.mask, p.num_joints, p.height,
                p.width, xtiles, p.top_k, p.use_threshold, p.threshold, w.cand_keys, w.cand_count, p.cand_capacity, w.flags);
  } else {
    PGMP_LAUNCH((nms_candidates_kernel<R, false>), grid, threads, 0, st, p.scoremaps, p.mask, p.num_joints, p.height,
                p.width, xtiles, p.top_k, p.use_threshold, p.threshold, w.cand_keys, w.cand_count, p.cand_capacity, w.flags);
  }
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

extern "C" int pgmp_gc_detect(const pgmp_gc_params* p, int64_t* counts, pgmp_stream_t stream) {
  int rc = validate(p);
  if (rc != PGMP_OK) return rc;
  if (!counts) return set_error(PGMP_ERR_INVALID, "null counts");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const GcWorkspace w = carve(*p);
  if (w.bytes > p->workspace_bytes)
    return set_error(PGMP_ERR_INVALID, "workspace too small: %llu < %llu", (unsigned long long)p->workspace_bytes,
                     (unsigned long long)w.bytes);
  const int B = p->batch, J = p->num_joints;
  PGMP_CUDA(cudaMemsetAsync(w.cand_count, 0, sizeof(uint32_t) * B * J, st));
  PGMP_CUDA(cudaMemsetAsync(w.flags, 0, sizeof(uint32_t), st));
  switch (p->pool_kernel / 2) {
    case 0: rc = launch_nms<0>(*p, w, st); break;
    case 1: rc = launch_nms<1>(*p, w, st); break;
    case 2: rc = launch_nms<2>(*p, w, st); break;
    case 3: rc = launch_nms<3>(*p, w, st); break;
    default: rc = launch_nms<4>(*p, w, st); break;
  }
  if (rc != PGMP_OK) return rc;
  int P = 1;
  while (P < p->cand_capacity) P <<= 1;
  const size_t sel_smem = sizeof(uint64_t) * P;
  if (sel_smem > 48 * 1024)
    PGMP_CUDA(cudaFuncSetAttribute(select_detections_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
  PGMP_LAUNCH(select_detections_kernel, B * J, 256, sel_smem, st, w.cand_keys, w.cand_count, p->cand_capacity, p->top_k,
              p->use_threshold, p->threshold, p->max_det_per_type, w.det_idx, w.det_n1, w.det_n2, w.flags);
  PGMP_LAUNCH(layout_nodes_kernel, B, 256, sizeof(int32_t) * 2 * J, st, p->scoremaps, p->mask, J, p->height, p->width,
              p->use_threshold, p->max_det_per_type, p->max_nodes, w.det_idx, w.det_n1, w.det_n2, w.node_xyt,
              w.node_score, w.node_count, w.flags);
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
