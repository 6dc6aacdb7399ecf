// fp32 SIMT implementation of the message-passing network (parity mode, PGMP_PRECISION_FP32) and
// of the small per-node stages every precision mode shares (embeddings, per-node tables, node
// update, heads).  One thread owns one row (edge slot or node); activations sit transposed in
// shared memory ([feature][row], conflict-free), weights are streamed through shared memory and
// broadcast.  Reference: NodeClassificationMPNSimple.py:62-97, layers.py:8-29, 32-86, 157-274.
//
// Algebra used (exact in real arithmetic, fp32 summation order differs from the reference):
//   mlp_edge.0 [x_i ; x_j ; e0 ; e] = W1_dst x_i + W1_src x_j + W1_e0 e0 + W1_e e + b1
//     -> per-node tables P = W1_dst x, Q = W1_src x (one small GEMM per step instead of per edge),
//        per-edge constant C = W1_e0 e0 + b1 (the weights are shared by all steps), so the per-edge
//        work of a step is one 64x64 product instead of a 384x64 one.
//   mlp_node[t] [x_i ; e'] = Wm_x[t] x_i + Wm_e[t] e' + bm[t]  -> table R[t] = Wm_x[t] x + bm[t].
#include "mpn_common.cuh"
#include "simt_mlp.cuh"

namespace pgmp {

int mpn_node_update_hier(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st);

namespace {

// ------------------------------------------------------------------------------------------------
// Embedding MLPs (layers.py:8-29; NodeClassificationMPNSimple.py:65-66).  General widths up to DMAX,
// two ping-pong buffers.  `row_map` (slot -> edge id) reorders the edge embedding into slot order;
// `extra_wt` appends the per-edge constant C = W1_e0 g0 + b1 as a second output.
// ------------------------------------------------------------------------------------------------
template <int DMAX>
__global__ void __launch_bounds__(kTile) mlp_chain_kernel(
    const pgmp_mlp mlp, const float* __restrict__ in, int64_t in_sr, int64_t in_sc,
    const int32_t* __restrict__ row_map, int64_t M, float* __restrict__ out, const float* __restrict__ extra_wt,
    const float* __restrict__ extra_b, float* __restrict__ out2) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;
  float* bufB = bufA + DMAX * kTileP;
  float* ws = bufB + DMAX * kTileP;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int K0 = mlp.dims[0];
  for (int idx = threadIdx.x; idx < kTile * K0; idx += blockDim.x) {
    const int r = idx / K0, c = idx - r * K0;
    const int64_t row = row0 + r;
    int64_t srow = -1;
    if (row < M) srow = row_map ? (int64_t)row_map[row] : row;
    bufA[(size_t)c * kTileP + r] = srow >= 0 ? in[srow * in_sr + c * in_sc] : 0.f;
  }
  float* cur = bufA;
  float* nxt = bufB;
  float acc[kD];
  for (int l = 0; l < mlp.n_layers; ++l) {
    const int K = mlp.dims[l], O = mlp.dims[l + 1];
    for (int o0 = 0; o0 < O; o0 += kD) {
      init_bias(acc, mlp.bias[l], O, o0);
      matvec64(acc, cur, K, mlp.wt[l], O, o0, ws);
      put_col(nxt, acc, O, o0, mlp.relu[l] != 0);
    }
    float* tmp = cur; cur = nxt; nxt = tmp;
  }
  const int O = mlp.dims[mlp.n_layers];
  if (mlp.post_relu || mlp.post_scale) {
    for (int o = 0; o < O; ++o) {
      float v = cur[(size_t)o * kTileP + threadIdx.x];
      if (mlp.post_relu) v = fmaxf(v, 0.f);
      if (mlp.post_scale) v = fmaf(v, __ldg(mlp.post_scale + o), __ldg(mlp.post_shift + o));
      cur[(size_t)o * kTileP + threadIdx.x] = v;
    }
  }
  __syncthreads();
  store_tile_rowmajor(cur, out, row0, M, O);
  if (extra_wt) {   // O == kD here
    init_bias(acc, extra_b, kD, 0);
    matvec64(acc, cur, kD, extra_wt, kD, 0, ws);
    put_col(nxt, acc, kD, 0, false);
    __syncthreads();
    store_tile_rowmajor(nxt, out2, row0, M, kD);
  }
}

// ------------------------------------------------------------------------------------------------
// Per-node tables for one step: blockIdx.y = 0 -> P, 1 -> Q, 2 + t -> R[t].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile) node_tables_kernel(
    const float* __restrict__ h0, const float* __restrict__ h, int64_t N, int skip, const float* __restrict__ w1_dst,
    const float* __restrict__ w1_src, const float* __restrict__ b1, const float* __restrict__ wm_x,
    const float* __restrict__ bm, int per_type, float* __restrict__ tab_p, float* __restrict__ tab_q,
    float* __restrict__ tab_r) {
  extern __shared__ __align__(16) float smem[];
  float* in = smem;                       // [nd][kTileP]
  float* outb = in + 2 * kD * kTileP;     // [64][kTileP]
  float* ws = outb + kD * kTileP;
  const int nd = skip ? 2 * kD : kD;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  for (int idx = threadIdx.x; idx < kTile * nd; idx += blockDim.x) {
    const int r = idx / nd, c = idx - r * nd;
    const int64_t row = row0 + r;
    float v = 0.f;
    if (row < N) v = skip ? (c < kD ? h0[row * kD + c] : h[row * kD + c - kD]) : h[row * kD + c];   // [h0 ; h], :77
    in[(size_t)c * kTileP + r] = v;
  }
  const int which = blockIdx.y;
  const float* Wt;
  const float* bias = nullptr;
  float* dst;
  if (which == 0) { Wt = w1_dst; bias = skip ? nullptr : b1; dst = tab_p; }
  else if (which == 1) { Wt = w1_src; dst = tab_q; }
  else {
    const int t = which - 2;
    const int tm = per_type ? t : 0;
    Wt = wm_x + (size_t)tm * nd * kD;
    bias = bm + (size_t)tm * kD;
    dst = tab_r + (size_t)t * N * kD;
  }
  float acc[kD];
  init_bias(acc, bias, kD, 0);
  matvec64(acc, in, nd, Wt, kD, 0, ws);
  put_col(outb, acc, kD, 0, false);
  __syncthreads();
  store_tile_rowmajor(outb, dst, row0, N, kD);
}

// Node update (layers.py:83-86, 253-258) + node / class heads when this step is reported
// (NodeClassificationMPNSimple.py:81-83, 93-94).
__global__ void __launch_bounds__(kTile) node_update_kernel(
    AggrView av, int64_t N, int T, int has_update, const float* __restrict__ wu, const float* __restrict__ bu,
    float* __restrict__ h, int with_heads, const pgmp_mlp node_head, const pgmp_mlp class_head,
    float* __restrict__ node_logits, float* __restrict__ class_logits) {
  extern __shared__ __align__(16) float smem[];
  float* bufA = smem;                   // [64][kTileP]
  float* bufB = bufA + kD * kTileP;
  float* ws = bufB + kD * kTileP;
  const int64_t row0 = (int64_t)blockIdx.x * kTile;
  const int64_t row = row0 + threadIdx.x;
  const int64_t srow = row < N ? row : N - 1;   // out-of-range threads mirror the last node, never store
  float u[kD], acc[kD];
  if (has_update) {
    init_bias(acc, bu, kD, 0);
    for (int t = 0; t < T; ++t) {
      merge_parts(av, t, srow, N, u);
      __syncthreads();                  // previous matvec finished reading bufA
      put_col(bufA, u, kD, 0, false);
      matvec64(acc, bufA, kD, wu + (size_t)t * kD * kD, kD, 0, ws);   // update_mlp columns of type t, :255-257
    }
    __syncthreads();
    put_col(bufA, acc, kD, 0, true);
  } else {
    merge_parts(av, 0, srow, N, u);
    put_col(bufA, u, kD, 0, false);
  }
  __syncthreads();
  store_tile_rowmajor(bufA, h, row0, N, kD);
  if (!with_heads) return;
  run_small_chain(node_head, bufA, bufB, ws);
  if (row < N) node_logits[row] = bufB[threadIdx.x];
  __syncthreads();
  run_small_chain(class_head, bufA, bufB, ws);
  __syncthreads();
  const int J = class_head.dims[class_head.n_layers];
  store_tile_rowmajor(bufB, class_logits, row0, N, J);
}

// ------------------------------------------------------------------------------------------------
// One message-passing step over a 128-slot tile (fp32 SIMT).
// ------------------------------------------------------------------------------------------------
struct EdgeStepArgs {
  const int32_t* slot_edge; const int32_t* slot_src; const int32_t* slot_dst;
  const int32_t* group_start; const int32_t* group_pstart; const int32_t* bin_lstart; const int32_t* bin_lpart;
  float* g; const float* c0; const float* tab_p; const float* tab_q; const float* tab_r;
  const float* w1_e; const float* w2; const float* b2; const float* wm_e; const float* wa; const float* ba;
  float* part_val; float* part_mx; float* part_se;
  int64_t N, E;
  int T, per_type, aggr, attn, attn_cols, with_head;
  pgmp_mlp edge_head;
  float* edge_logits;
};

__global__ void __launch_bounds__(kTile) edge_step_kernel(const EdgeStepArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* buf = smem;                    // [64][kTileP]
  float* ws = buf + kD * kTileP;        // [32][64]
  float* s_a = ws + kWs;                // [128] attention logits
  int* s_dst = reinterpret_cast<int*>(s_a + kTile);
  const int tile = blockIdx.x;
  const int64_t slot0 = (int64_t)tile * kTile;
  if (slot0 >= a.group_start[a.T]) return;
  int t = 0;
  while (t + 1 < a.T && slot0 >= a.group_start[t + 1]) ++t;
  const int tm = a.per_type ? t : 0;
  const int64_t slot = slot0 + threadIdx.x;
  const int e = a.slot_edge[slot];
  const int src = a.slot_src[slot], dst = a.slot_dst[slot];

  float acc[kD];
  // hidden = ReLU(W1_e g + C + P[dst] + Q[src])     (layers.py:171-175, 214)
  if (e >= 0) {
    const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.tab_p + (size_t)dst * kD);
    const float4* __restrict__ q4 = reinterpret_cast<const float4*>(a.tab_q + (size_t)src * kD);
    const float4* __restrict__ c4 = a.c0 ? reinterpret_cast<const float4*>(a.c0 + slot * kD) : nullptr;
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
      const float4 p = p4[q], s = q4[q];
      float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4) c = c4[q];
      acc[4 * q + 0] = c.x + p.x + s.x;
      acc[4 * q + 1] = c.y + p.y + s.y;
      acc[4 * q + 2] = c.z + p.z + s.z;
      acc[4 * q + 3] = c.w + p.w + s.w;
    }
  } else {
#pragma unroll
    for (int o = 0; o < kD; ++o) acc[o] = 0.f;
  }
  load_tile_rowmajor(buf, a.g, slot0, slot0 + kTile, kD);
  matvec64(acc, buf, kD, a.w1_e, kD, 0, ws);
  __syncthreads();
  put_col(buf, acc, kD, 0, true);
  // e' = ReLU(W2 hidden + b2)
  init_bias(acc, a.b2, kD, 0);
  matvec64(acc, buf, kD, a.w2, kD, 0, ws);
  float att = 0.f;
  if (a.attn) {   // attention logit from the updated edge feature (layers.py:245)
    const int col = a.attn == PGMP_ATTN_PER_TYPE ? t : 0;
    att = __ldg(a.ba + col);
#pragma unroll
    for (int o = 0; o < kD; ++o) att = fmaf(fmaxf(acc[o], 0.f), __ldg(a.wa + o * a.attn_cols + col), att);
  }
  __syncthreads();
  put_col(buf, acc, kD, 0, true);
  __syncthreads();
  store_tile_rowmajor(buf, a.g, slot0, slot0 + kTile, kD);
  // message m = ReLU(Wm_e[t] e' + R[t][dst])         (layers.py:222-224, 264-274 / :78-81)
  if (e >= 0) {
    const float4* __restrict__ r4 = reinterpret_cast<const float4*>(a.tab_r + ((size_t)t * a.N + dst) * kD);
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
      const float4 r = r4[q];
      acc[4 * q + 0] = r.x; acc[4 * q + 1] = r.y; acc[4 * q + 2] = r.z; acc[4 * q + 3] = r.w;
    }
  } else {
#pragma unroll
    for (int o = 0; o < kD; ++o) acc[o] = 0.f;
  }
  matvec64(acc, buf, kD, a.wm_e + (size_t)tm * kD * kD, kD, 0, ws);
  __syncthreads();
  put_col(buf, acc, kD, 0, true);
  s_a[threadIdx.x] = att;
  s_dst[threadIdx.x] = e >= 0 ? dst : -1;
  __syncthreads();
  // reduce every run of equal targets (one bin, or the part of it inside this tile)
  if (e >= 0 && (threadIdx.x == 0 || s_dst[threadIdx.x - 1] != dst)) {
    int r1 = threadIdx.x;
    while (r1 + 1 < kTile && s_dst[r1 + 1] == dst) ++r1;
    const int64_t bin = (int64_t)t * a.N + dst;
    const int first_slot = a.group_start[t] + a.bin_lstart[bin];
    const int64_t prow = (int64_t)a.group_pstart[t] + a.bin_lpart[bin] + (tile - (first_slot >> 7));
    float u[kD];
    if (a.attn) {
      float mx = -INFINITY;
      for (int r = threadIdx.x; r <= r1; ++r) mx = fmaxf(mx, s_a[r]);
      float se = 0.f;
#pragma unroll
      for (int o = 0; o < kD; ++o) u[o] = 0.f;
      for (int r = threadIdx.x; r <= r1; ++r) {
        const float wgt = __expf(s_a[r] - mx);
        se += wgt;
#pragma unroll
        for (int o = 0; o < kD; ++o) u[o] = fmaf(wgt, buf[(size_t)o * kTileP + r], u[o]);
      }
      a.part_mx[prow] = mx;
      a.part_se[prow] = se;
    } else {
#pragma unroll
      for (int o = 0; o < kD; ++o) u[o] = buf[(size_t)o * kTileP + threadIdx.x];
      for (int r = threadIdx.x + 1; r <= r1; ++r) {
#pragma unroll
        for (int o = 0; o < kD; ++o) {
          const float v = buf[(size_t)o * kTileP + r];
          u[o] = a.aggr == PGMP_AGGR_MAX ? fmaxf(u[o], v) : u[o] + v;
        }
      }
    }
    float4* __restrict__ o4 = reinterpret_cast<float4*>(a.part_val + prow * kD);
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) o4[q] = make_float4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
  }
  if (!a.with_head) return;
  // edge head on e' (NodeClassificationMPNSimple.py:84): reload the tile this CTA just stored
  __syncthreads();
  load_tile_rowmajor(buf, a.g, slot0, slot0 + kTile, kD);
  run_small_chain(a.edge_head, buf, buf, ws);
  if (e >= 0) a.edge_logits[e] = buf[threadIdx.x];
}

// Edge head (NodeClassificationMPNSimple.py:84) as its own pass over the slot-ordered edge features
// (used by the tensor-core mode, whose step kernel does not carry the head).
__global__ void __launch_bounds__(kTile) edge_head_kernel(const float* __restrict__ g, const int32_t* __restrict__ slot_edge,
                                                          const int32_t* __restrict__ group_start, int T, int image,
                                                          const pgmp_mlp head, float* __restrict__ edge_logits) {
  extern __shared__ __align__(16) float smem[];
  float* buf = smem;
  float* ws = buf + kD * kTileP;
  const int64_t slot0 = (int64_t)blockIdx.x * kTile;
  if (slot0 >= group_start[T]) return;
  if (image) {   // tensor-core mode keeps the edge features as bf16 hi/lo SWIZZLE_128B tile images
    const uint16_t* __restrict__ img = reinterpret_cast<const uint16_t*>(g) + (size_t)blockIdx.x * kTile * kD * 2;
    for (int idx = threadIdx.x; idx < kTile * kD; idx += blockDim.x) {
      const int r = idx >> 6, c = idx & 63;
      const int off = ((r >> 3) * 1024 + (r & 7) * 128 + (((c >> 3) ^ (r & 7)) << 4)) / 2 + (c & 7);
      const float hi = __uint_as_float((uint32_t)img[off] << 16), lo = __uint_as_float((uint32_t)img[off + kTile * kD] << 16);
      buf[(size_t)c * kTileP + r] = hi + lo;
    }
  } else {
    load_tile_rowmajor(buf, g, slot0, slot0 + kTile, kD);
  }
  run_small_chain(head, buf, buf, ws);
  const int e = slot_edge[slot0 + threadIdx.x];
  if (e >= 0) edge_logits[e] = buf[threadIdx.x];
}

int check_small(const pgmp_mlp& m, const char* name, int in_dim, int out_dim) {
  if (m.n_layers < 1 || m.n_layers > PGMP_MAX_LAYERS) return set_error(PGMP_ERR_INVALID, "%s: %d layers", name, m.n_layers);
  if (m.dims[0] != in_dim) return set_error(PGMP_ERR_INVALID, "%s: input width %d != %d", name, m.dims[0], in_dim);
  if (out_dim > 0 && m.dims[m.n_layers] != out_dim)
    return set_error(PGMP_ERR_INVALID, "%s: output width %d != %d", name, m.dims[m.n_layers], out_dim);
  for (int l = 0; l <= m.n_layers; ++l)
    if (m.dims[l] < 1 || m.dims[l] > kD) return set_error(PGMP_ERR_INVALID, "%s: layer width %d not in [1,64]", name, m.dims[l]);
  return PGMP_OK;
}

}  // namespace

int mpn_embed(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st);
int mpn_embed_impl(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st, bool with_edges);
int mpn_embed_nodes(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st) { return mpn_embed_impl(p, w, st, false); }

// Embeddings shared by both precision modes: h0 = node_embedding(x); g = edge_embedding(edge_attr) in
// slot order; C = W1_e0 g + b1.
int mpn_embed(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st) { return mpn_embed_impl(p, w, st, true); }

int mpn_embed_impl(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st, bool with_edges) {
  const int64_t N = p.num_nodes;
  auto maxdim = [](const pgmp_mlp& m) { int d = 0; for (int l = 0; l <= m.n_layers; ++l) d = d > m.dims[l] ? d : m.dims[l]; return d; };
  for (const pgmp_mlp* m : {&p.node_emb, &p.edge_emb}) {
    if (m->n_layers < 1 || m->n_layers > PGMP_MAX_LAYERS || m->dims[m->n_layers] != kD || maxdim(*m) > 128)
      return set_error(PGMP_ERR_INVALID, "embedding MLP: need 1..%d layers, widths <= 128, output width 64", PGMP_MAX_LAYERS);
  }
  const size_t smem128 = sizeof(float) * (2 * 128 * kTileP + kWs), smem64 = sizeof(float) * (2 * kD * kTileP + kWs);
  PGMP_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem128));
  PGMP_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
  const unsigned ntiles = (unsigned)ceil_div<int64_t>(N, kTile);
  if (maxdim(p.node_emb) > kD) {
    PGMP_LAUNCH((mlp_chain_kernel<128>), ntiles, kTile, smem128, st, p.node_emb, p.x, p.x_stride_n, p.x_stride_c,
                (const int32_t*)nullptr, N, w.h0, (const float*)nullptr, (const float*)nullptr, (float*)nullptr);
  } else {
    PGMP_LAUNCH((mlp_chain_kernel<64>), ntiles, kTile, smem64, st, p.node_emb, p.x, p.x_stride_n, p.x_stride_c,
                (const int32_t*)nullptr, N, w.h0, (const float*)nullptr, (const float*)nullptr, (float*)nullptr);
  }
  if (with_edges && p.num_edges > 0) {
    const unsigned etiles = (unsigned)(w.max_slots / kTile);
    const int64_t F = p.edge_emb.dims[0];
    const float* ew = p.skip ? p.w1_e0 : nullptr;
    if (maxdim(p.edge_emb) > kD) {
      PGMP_LAUNCH((mlp_chain_kernel<128>), etiles, kTile, smem128, st, p.edge_emb, p.edge_attr, F, (int64_t)1,
                  (const int32_t*)w.slot_edge, (int64_t)w.max_slots, w.g, ew, p.b1, w.c0);
    } else {
      PGMP_LAUNCH((mlp_chain_kernel<64>), etiles, kTile, smem64, st, p.edge_emb, p.edge_attr, F, (int64_t)1,
                  (const int32_t*)w.slot_edge, (int64_t)w.max_slots, w.g, ew, p.b1, w.c0);
    }
  }
  return PGMP_OK;
}

// out[M][dims[-1]] = mlp(in[M][dims[0]]) for a 64-wide-or-narrower chain (heads)
int mpn_run_mlp_rows(const pgmp_mlp& mlp, const float* in, int64_t M, float* out, cudaStream_t st) {
  const size_t smem64 = sizeof(float) * (2 * kD * kTileP + kWs);
  PGMP_CUDA(cudaFuncSetAttribute(mlp_chain_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem64));
  PGMP_LAUNCH((mlp_chain_kernel<64>), (unsigned)ceil_div<int64_t>(M, kTile), kTile, smem64, st, mlp, in, (int64_t)mlp.dims[0],
              (int64_t)1, (const int32_t*)nullptr, M, out, (const float*)nullptr, (const float*)nullptr, (float*)nullptr);
  return PGMP_OK;
}

int mpn_node_tables(const pgmp_mpn_params& p, const MpnWorkspace& w, const float* h, cudaStream_t st) {
  const size_t smem = sizeof(float) * (3 * kD * kTileP + kWs);
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  PGMP_LAUNCH(node_tables_kernel, dim3((unsigned)ceil_div<int64_t>(p.num_nodes, kTile), 2 + p.num_types), kTile, smem, st,
              w.h0, h, p.num_nodes, p.skip, p.w1_dst, p.w1_src, p.b1, p.wm_x, p.bm, p.per_type, w.tab_p, w.tab_q,
              w.tab_r);
  return PGMP_OK;
}

int mpn_node_update(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st) {
  const size_t smem = sizeof(float) * (2 * kD * kTileP + kWs);
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  AggrView av{w.bin_count, w.bin_lstart, w.bin_lpart, w.group_pstart, w.part_val, w.part_mx, w.part_se, p.aggr, p.attn};
  const int64_t N = p.num_nodes;
  float* nl = out_slot >= 0 ? p.node_logits + (size_t)out_slot * N : nullptr;
  float* cl = out_slot >= 0 ? p.class_logits + (size_t)out_slot * N * p.num_classes : nullptr;
  PGMP_LAUNCH(node_update_kernel, (unsigned)ceil_div<int64_t>(N, kTile), kTile, smem, st, av, N, p.num_types,
              p.has_update_mlp, p.wu, p.bu, w.h, out_slot >= 0 ? 1 : 0, p.node_head, p.class_head, nl, cl);
  return PGMP_OK;
}

int mpn_edge_head(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, bool image, cudaStream_t st) {
  const size_t smem = sizeof(float) * (kD * kTileP + kWs);
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(edge_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  PGMP_LAUNCH(edge_head_kernel, (unsigned)(w.max_slots / kTile), kTile, smem, st, w.g, w.slot_edge, w.group_start,
              p.num_types, image ? 1 : 0, p.edge_head, p.edge_logits + (size_t)out_slot * p.num_edges);
  return PGMP_OK;
}

int mpn_validate_heads(const pgmp_mpn_params& p) {
  int rc;
  if ((rc = check_small(p.edge_head, "edge_classification", kD, 1)) != PGMP_OK) return rc;
  if ((rc = check_small(p.node_head, "node_classification", kD, 1)) != PGMP_OK) return rc;
  if ((rc = check_small(p.class_head, "classification", kD, p.num_classes)) != PGMP_OK) return rc;
  return PGMP_OK;
}

int mpn_forward_simt(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st) {
  const int64_t N = p.num_nodes, E = p.num_edges;
  int rc;
  if ((rc = mpn_embed(p, w, st)) != PGMP_OK) return rc;
  PGMP_CUDA(cudaMemcpyAsync(w.h, w.h0, sizeof(float) * N * kD, cudaMemcpyDeviceToDevice, st));
  const size_t smem_edge = sizeof(float) * (kD * kTileP + kWs + kTile) + sizeof(int) * kTile;
  PGMP_CUDA(cudaFuncSetAttribute(edge_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_edge));
  EdgeStepArgs a;
  a.slot_edge = w.slot_edge; a.slot_src = w.slot_src; a.slot_dst = w.slot_dst;
  a.group_start = w.group_start; a.group_pstart = w.group_pstart; a.bin_lstart = w.bin_lstart; a.bin_lpart = w.bin_lpart;
  a.g = w.g; a.c0 = p.skip ? w.c0 : nullptr; a.tab_p = w.tab_p; a.tab_q = w.tab_q; a.tab_r = w.tab_r;
  a.w1_e = p.w1_e; a.w2 = p.w2; a.b2 = p.b2; a.wm_e = p.wm_e; a.wa = p.wa; a.ba = p.ba;
  a.part_val = w.part_val; a.part_mx = w.part_mx; a.part_se = w.part_se;
  a.N = N; a.E = E; a.T = p.num_types; a.per_type = p.per_type; a.aggr = p.aggr; a.attn = p.attn;
  a.attn_cols = p.attn == PGMP_ATTN_PER_TYPE ? 17 : 1;
  a.edge_head = p.edge_head;
  const int first_out = p.steps - p.aux_loss_steps - 1 > 0 ? p.steps - p.aux_loss_steps - 1 : 0;   // NodeClassificationMPNSimple.py:81
  for (int s = 0; s < p.steps; ++s) {
    if (s > 0) {
      const int prev_slot = (s - 1) >= first_out ? (s - 1) - first_out : -1;
      if ((rc = (p.update_hier ? mpn_node_update_hier(p, w, prev_slot, st) : mpn_node_update(p, w, prev_slot, st))) != PGMP_OK) return rc;
    }
    if ((rc = mpn_node_tables(p, w, w.h, st)) != PGMP_OK) return rc;
    const int slot = s >= first_out ? s - first_out : -1;
    a.with_head = slot >= 0;
    a.edge_logits = slot >= 0 ? p.edge_logits + (size_t)slot * E : nullptr;
    if (E > 0) PGMP_LAUNCH(edge_step_kernel, (unsigned)(w.max_slots / kTile), kTile, smem_edge, st, a);
  }
  return p.update_hier ? mpn_node_update_hier(p, w, (p.steps - 1) - first_out, st)
                       : mpn_node_update(p, w, (p.steps - 1) - first_out, st);
}

}  // namespace pgmp
