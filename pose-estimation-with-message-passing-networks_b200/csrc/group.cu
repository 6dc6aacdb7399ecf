// Grouping tail: sigmoid / node threshold / subgraph -> multicut weights -> GAEC (greedy additive
// edge contraction) -> connected-component labels -> per-person keypoint selection.
//
// Reference: src/valid.py:109-111, src/Utils/Utils.py:1448-1451 (threshold + subgraph), :499-514
// (pred_to_person), src/Utils/correlation_clustering/correlation_clustering_utils.py:99-136, 187-256
// (weights w = (p_ab + p_ba)/2 - 0.5 on kept pairs a < b), Utils.py:672-743 (graph_cluster_to_persons).
// The GAEC solver itself is the reference's missing native module (andres_graph_wrapper); this
// kernel follows the same published algorithm and tie rule as oracle/grouping.py: contract the live
// edge of maximal weight while it is >= 0; ties -> smallest (a, b), a < b, clusters named by their
// smallest member; weights accumulate in fp64.
//
// One CTA per image.  The cluster graph is a dense fp64 matrix in the caller's workspace (absent
// edges = -inf); every row caches its best upper-triangle entry, a contraction updates O(N) entries
// in parallel and only rows whose cached best was touched are rescanned (one warp per row).
#include "common.cuh"

#include <math_constants.h>

namespace pgmp {
namespace {

constexpr int kThreads = 512;    // an image has a few hundred nodes: more threads only make the ~5 barriers per contraction dearer
constexpr uint32_t kFull = 0xffffffffu;

struct GroupWs {
  float* adj32;      // [B][M*M] dense sum of kept edge probabilities (to_dense_adj)
  double* w;         // [B][M*M] cluster-graph weights, -inf = no edge
  int* mult;         // [B][M*M] multiplicity of the undirected edge a < b in the kept edge list
  double* best_val;  // [B][M]
  int* best_col;     // [B][M]
  int* rep;          // [B][M]
  int* flags;        // [B][M] bit0 alive, bit1 dirty, bit2 kept (node threshold)
  float* p_node;     // [B][M]
  int* type;         // [B][M] re-assigned joint type (argmax of the class head)
  int* comp;         // [B][M] component id per node
  int* dirty;        // [B][M] rows to rescan after a contraction (used when the bookkeeping does not fit shared memory)
  int* csize;        // [B][M] nodes per component root
  unsigned long long* pkey;   // [B][M*J] per (component root, type): best (score bits << 32 | ~node) of its nodes
  uint64_t bytes;
};

GroupWs carve(const pgmp_group_params& p, int M) {
  Carver c(p.workspace);
  GroupWs w;
  const uint64_t B = p.batch, mm = (uint64_t)M * M;
  w.adj32 = c.take<float>(B * mm);
  w.w = c.take<double>(B * mm);
  w.mult = c.take<int>(B * mm);
  w.best_val = c.take<double>(B * M);
  w.best_col = c.take<int>(B * M);
  w.rep = c.take<int>(B * M);
  w.flags = c.take<int>(B * M);
  w.p_node = c.take<float>(B * M);
  w.type = c.take<int>(B * M);
  w.comp = c.take<int>(B * M);
  w.dirty = c.take<int>(B * M);
  w.csize = c.take<int>(B * M);
  w.pkey = c.take<unsigned long long>(B * M * (uint64_t)p.num_joints);
  w.bytes = c.bytes();
  return w;
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// better = larger value, then smaller column
__device__ __forceinline__ bool better(double v, int c, double bv, int bc) { return v > bv || (v == bv && c < bc); }

// one warp rescans row r over alive columns > r
__device__ void rescan_row(const double* __restrict__ W, const int* __restrict__ flags, int n, int r,
                           double* __restrict__ best_val, int* __restrict__ best_col) {
  const int lane = threadIdx.x & 31;
  double bv = -CUDART_INF;
  int bc = 0x7fffffff;
  // sixteen independent loads per lane in flight (a whole row of up to 512 entries in ONE L2 round trip): the rescan of
  // the merged row sits on the critical path of every contraction (measured: the other warps wait for it at the barrier)
  for (int c0 = r + 1 + lane; c0 < n; c0 += 512) {
    double v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int c = c0 + 32 * k;
      v[k] = (c < n && (flags[c] & 1)) ? W[(size_t)r * n + c] : -CUDART_INF;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (v[k] != -CUDART_INF && better(v[k], c0 + 32 * k, bv, bc)) { bv = v[k]; bc = c0 + 32 * k; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ov = __shfl_xor_sync(kFull, bv, o);
    const int oc = __shfl_xor_sync(kFull, bc, o);
    if (better(ov, oc, bv, bc)) { bv = ov; bc = oc; }
  }
  if (lane == 0) { best_val[r] = bv; best_col[r] = bc; }
}

__global__ void __launch_bounds__(kThreads) group_kernel(const pgmp_group_params p, const GroupWs ws, int M, int use_smem) {
  __shared__ double s_val[32];
  __shared__ int s_row[32];
  __shared__ int s_u, s_v, s_any_lower, s_count, s_kept;
  __shared__ double s_best;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n0 = p.node_offsets[b], e0 = p.edge_offsets[b], e1 = p.edge_offsets[b + 1];
  const int n = (int)(p.node_offsets[b + 1] - n0);
  const int J = p.num_joints;
  float* __restrict__ A = ws.adj32 + (size_t)b * M * M;
  double* __restrict__ W = ws.w + (size_t)b * M * M;
  int* __restrict__ mult = ws.mult + (size_t)b * M * M;
  // the per-node bookkeeping of the greedy contraction lives in shared memory when it fits (24 bytes per node): every
  // contraction step reads it in three dependent phases, an L2 round trip each when it sits in the workspace
  extern __shared__ __align__(16) unsigned char s_dyn[];
  double* __restrict__ best_val = use_smem ? reinterpret_cast<double*>(s_dyn) : ws.best_val + (size_t)b * M;
  int* __restrict__ best_col = use_smem ? reinterpret_cast<int*>(s_dyn + (size_t)8 * M) : ws.best_col + (size_t)b * M;
  int* __restrict__ rep = use_smem ? best_col + M : ws.rep + (size_t)b * M;
  int* __restrict__ flags = use_smem ? best_col + 2 * M : ws.flags + (size_t)b * M;
  int* __restrict__ dirty = use_smem ? best_col + 3 * M : ws.dirty + (size_t)b * M;   // rows to rescan after a contraction
  __shared__ int s_ndirty;
  float* __restrict__ pn = ws.p_node + (size_t)b * M;
  int* __restrict__ typ = ws.type + (size_t)b * M;
  int* __restrict__ comp = ws.comp + (size_t)b * M;
  if (n <= 0 || n > M) {
    if (tid == 0) { p.num_components[b] = 0; p.num_persons[b] = 0; p.mutants[b] = 0; p.num_kept_edges[b] = n > M ? -1 : 0; }
    return;
  }
  if (tid == 0) { s_any_lower = 0; s_kept = 0; }
  // ---- nodes: probability, threshold (Utils.py:1450), class argmax (:699-702)
  for (int i = tid; i < n; i += kThreads) {
    const float pr = sigmoidf_(p.node_logits[n0 + i]);
    pn[i] = pr;
    flags[i] = 1 | (pr > p.node_threshold ? 4 : 0);
    rep[i] = i;
    int t = (int)p.joint_det[(n0 + i) * 3 + 2];
    if (p.class_logits) {
      const float* __restrict__ cl = p.class_logits + (n0 + i) * J;
      float bv = cl[0];
      t = 0;
      for (int j = 1; j < J; ++j)
        if (cl[j] > bv) { bv = cl[j]; t = j; }
    }
    typ[i] = t;
  }
  for (size_t i = tid; i < (size_t)n * n; i += kThreads) { A[i] = 0.f; W[i] = -CUDART_INF; mult[i] = 0; }
  __syncthreads();
  // ---- dense matrix of kept edge probabilities (subgraph + to_dense_adj, correlation_clustering_utils.py:117)
  for (int64_t e = e0 + tid; e < e1; e += kThreads) {
    const int s = (int)(p.edge_index[e] - n0), d = (int)(p.edge_index[p.num_edges + e] - n0);
    if (s < 0 || s >= n || d < 0 || d >= n) continue;
    if (!((flags[s] & 4) && (flags[d] & 4))) continue;
    atomicAdd(&s_kept, 1);
    atomicAdd(&A[(size_t)s * n + d], sigmoidf_(p.edge_logits[e]));
    if (s < d) atomicAdd(&mult[(size_t)s * n + d], 1); else s_any_lower = 1;
  }
  __syncthreads();
  if (tid == 0) p.num_kept_edges[b] = s_kept;
  if (p.cc_method == PGMP_CC_GREEDY) {
    // ---- CC_METHOD "greedy" (greedy_person_construction, Utils.py:517-626).  Symmetric float64 adjacency
    //      (p_ab + p_ba) / 2 with a unit diagonal (:531-534); types in order, within a type the nodes in order: an
    //      unclaimed node with score >= 0.5 becomes the core of a person and claims, per other type, the node it is
    //      connected to most strongly -- also one that is already claimed, unless its owner's edge is stronger (:556-583).
    __shared__ int s_tstart[66];
    __shared__ int s_ncand;
    for (size_t i = tid; i < (size_t)n * n; i += kThreads) {
      const int r = (int)(i / n), c = (int)(i - (size_t)r * n);
      W[i] = r == c ? 1.0 : ((double)A[i] + (double)A[(size_t)c * n + r]) / 2.0;
    }
    for (int i = tid; i < n; i += kThreads) rep[i] = -1;                 // taken_joints
    int* __restrict__ tlist = comp;                                       // nodes ordered by (type, index)
    int* __restrict__ order = dirty;                                      // ... those with score >= 0.5: the core candidates
    if (warp == 0) {
      int nt = 0, nc = 0;
      for (int t = 0; t < J; ++t) {
        if (lane == 0) s_tstart[t] = nt;
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int i = i0 + lane;
          const bool is_t = i < n && typ[i] == t;
          const bool is_c = is_t && !(pn[i] < 0.5f);
          const uint32_t mt = __ballot_sync(kFull, is_t), mc = __ballot_sync(kFull, is_c);
          if (is_t) tlist[nt + __popc(mt & ((1u << lane) - 1u))] = i;
          if (is_c) order[nc + __popc(mc & ((1u << lane) - 1u))] = i;
          nt += __popc(mt);
          nc += __popc(mc);
        }
      }
      if (lane == 0) { s_tstart[J] = nt; s_ncand = nc; }
    }
    __syncthreads();
    const int ncand = s_ncand;
    for (int k = 0; k < ncand; ++k) {
      const int i = order[k];
      // claimed by an earlier core -> not a core itself.  Every thread decides on its own: rep[i] changes in this
      // iteration only by thread 0 marking i as its own core below (claims go to nodes of other types), which the
      // test tolerates -- a thread that reads late must not skip the iteration and its barrier.
      { const int r = rep[i]; if (r != -1 && r != i) continue; }
      const int t = typ[i];
      if (tid == 0) rep[i] = i;
      for (int j = warp; j < J; j += kThreads / 32) {
        if (j == t) continue;
        double bv = 0.0;                                                   // masked row: entries of other types are 0 (:566-569)
        int bi = 0x7fffffff;
        for (int q = s_tstart[j] + lane; q < s_tstart[j + 1]; q += 32) {
          const int c = tlist[q];
          const double v = W[(size_t)i * n + c];
          if (v > bv || (v == bv && v > 0.0 && c < bi)) { bv = v; bi = c; }   // first maximum
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(kFull, bv, o);
          const int oi = __shfl_xor_sync(kFull, bi, o);
          if (ov > bv || (ov == bv && ov > 0.0 && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0 && bv != 0.0) {                                      // :570-571 (the target is of another type: never i itself)
          const int owner = rep[bi];
          if (owner == -1 || !(W[(size_t)owner * n + bi] > bv)) rep[bi] = i;   // :573-583
        }
      }
      __syncthreads();
    }
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) p.person_labels[n0 + i] = rep[i];
    if (tid == 0) p.num_components[b] = 0;
  }
  // ---- multicut weights on pairs a < b present in the kept edge list (:221-227)
  const int any_lower = s_any_lower;
  if (p.cc_method != PGMP_CC_GREEDY)
  for (size_t i = tid; i < (size_t)n * n; i += kThreads) {
    const int a = (int)(i / n), c = (int)(i - (size_t)a * n);
    if (a < c && mult[i] > 0) {
      float v = A[i] + A[(size_t)c * n + a];
      if (any_lower) v = v / 2.f;                                   // :118-121 (else mirrored, :114-117)
      const double w = (double)(v - 0.5f) * (double)mult[i];
      W[i] = w;
      W[(size_t)c * n + a] = w;
    }
  }
  __syncthreads();
  if (p.cc_method == PGMP_CC_GAEC)
    for (int r = warp; r < n; r += kThreads / 32) rescan_row(W, flags, n, r, best_val, best_col);
  __syncthreads();
  // ---- CC_METHOD "threshold" (Utils.py:508-509): the kept edges with probability above the edge threshold are the
  //      solution; min-label propagation over the image's edge list + pointer jumping gives every node the smallest
  //      node of its component as representative (the invariant the greedy contraction below also keeps)
  if (p.cc_method == PGMP_CC_THRESHOLD) {
    __shared__ int s_changed;
    for (;;) {
      __syncthreads();
      if (tid == 0) s_changed = 0;
      __syncthreads();
      for (int64_t e = e0 + tid; e < e1; e += kThreads) {
        const int s = (int)(p.edge_index[e] - n0), d = (int)(p.edge_index[p.num_edges + e] - n0);
        if (s < 0 || s >= n || d < 0 || d >= n) continue;
        if (!((flags[s] & 4) && (flags[d] & 4))) continue;
        if (!(sigmoidf_(p.edge_logits[e]) > p.edge_threshold)) continue;
        const int ls = rep[s], ld = rep[d];
        if (ls != ld) {
          const int m = min(ls, ld);
          atomicMin(&rep[s], m);
          atomicMin(&rep[d], m);
          s_changed = 1;
        }
      }
      __syncthreads();
      for (int i = tid; i < n; i += kThreads) {       // pointer jumping
        int r = rep[i];
        while (rep[r] != r) r = rep[r];
        rep[i] = r;
      }
      __syncthreads();
      if (!s_changed) break;
    }
  }
  // ---- GAEC
  for (int it = 0; it < n && p.cc_method == PGMP_CC_GAEC; ++it) {
    // global best over rows: larger value, then smaller row (rows cache their smallest best column)
    double bv = -CUDART_INF;
    int br = 0x7fffffff;
    for (int r = tid; r < n; r += kThreads)
      if ((flags[r] & 1) && better(best_val[r], r, bv, br)) { bv = best_val[r]; br = r; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(kFull, bv, o);
      const int orow = __shfl_xor_sync(kFull, br, o);
      if (better(ov, orow, bv, br)) { bv = ov; br = orow; }
    }
    if (lane == 0) { s_val[warp] = bv; s_row[warp] = br; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < kThreads / 32 ? s_val[lane] : -CUDART_INF;
      br = lane < kThreads / 32 ? s_row[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(kFull, bv, o);
        const int orow = __shfl_xor_sync(kFull, br, o);
        if (better(ov, orow, bv, br)) { bv = ov; br = orow; }
      }
      if (lane == 0) { s_best = bv; s_u = br; s_v = br < n ? best_col[br] : 0; }
    }
    __syncthreads();
    if (!(s_best >= 0.0)) break;                                    // "there must be negative weights", :213,222
    const int u = s_u, v = s_v;                                     // u < v: v is contracted into u
    if (tid == 0) { flags[v] = 0; flags[u] |= 2; dirty[0] = u; s_ndirty = 1; }   // (the loop below skips q == u, q == v)
    __syncthreads();
    for (int q = tid; q < n; q += kThreads) {
      if (rep[q] == v) rep[q] = u;
      if (q == u || q == v || !(flags[q] & 1)) continue;
      const double wv = W[(size_t)v * n + q];
      // entries (q, v) disappear: rows q < v whose best was column v must rescan
      bool dq = (flags[q] & 2) != 0;
      if (q < v && best_col[q] == v) dq = true;
      if (wv != -CUDART_INF) {
        const double wu = W[(size_t)u * n + q];
        const double nw = (wu == -CUDART_INF) ? wv : wu + wv;
        W[(size_t)u * n + q] = nw;
        W[(size_t)q * n + u] = nw;
        if (q < u) {                                                // entry (q, u) lives in row q
          if (better(nw, u, best_val[q], best_col[q])) { best_val[q] = nw; best_col[q] = u; dq = false; }
          else if (best_col[q] == u) dq = true;
        }
      }
      // only row q's own thread decides whether row q is dirty: the rows to rescan go to a list, so that the warps
      // below do not each walk all n flags (measured: that walk and its barrier were 30 % of the kernel)
      if (dq) { flags[q] |= 2; dirty[atomicAdd(&s_ndirty, 1)] = q; }
    }
    __syncthreads();
    const int nd = s_ndirty;
    for (int i = warp; i < nd; i += kThreads / 32) {
      const int r = dirty[i];
      rescan_row(W, flags, n, r, best_val, best_col);
      if (lane == 0) flags[r] &= ~2;
    }
    __syncthreads();
  }
  __syncthreads();
  // ---- component labels in order of the smallest member (scipy connected_components, Utils.py:688-691)
  if (p.cc_method != PGMP_CC_GREEDY) {
    if (tid == 0) {
      int c = 0;
      for (int i = 0; i < n; ++i)
        if (rep[i] == i) comp[i] = c++;
      s_count = c;
      p.num_components[b] = c;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kThreads) p.person_labels[n0 + i] = comp[rep[i]];
  }
  __syncthreads();
  // ---- persons (Utils.py:692-741): components with more than one node, per type the node with the best score.  Every
  //      node votes for its (component root, type) slot with score bits << 32 | ~node: the largest key is the first
  //      maximum of :718; then one warp walks the roots in order (persons are numbered by their smallest node).
  int* __restrict__ csize = ws.csize + (size_t)b * M;
  unsigned long long* __restrict__ pkey = ws.pkey + (size_t)b * M * J;
  for (int i = tid; i < n; i += kThreads) csize[i] = 0;
  for (int i = tid; i < n * J; i += kThreads) pkey[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < n; i += kThreads) {
    if (rep[i] < 0) continue;                                      // greedy: a node no person has claimed
    atomicAdd(&csize[rep[i]], 1);
    const float sc = pn[i] > 0.f ? pn[i] : 0.f;
    atomicMax(&pkey[(size_t)rep[i] * J + typ[i]], ((unsigned long long)__float_as_uint(sc) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i));
  }
  __syncthreads();
  if (warp == 0) {
    int n_person = 0, mutant = 0;
    double* __restrict__ out = p.persons + (size_t)b * p.max_persons * J * 3;
    for (int r0 = 0; r0 < n; r0 += 32) {
      const int rr = r0 + lane;
      const int sz = rr < n ? csize[rr] : 0;                        // (only cluster ids -- roots / core nodes -- have members)
      if (__any_sync(kFull, sz > J)) mutant = 1;                     // :703-706
      uint32_t roots = __ballot_sync(kFull, sz > 1);                 // :708
      while (roots) {
        const int r = r0 + __ffs(roots) - 1;
        roots &= roots - 1;
        int valid = 0;
        for (int t = lane; t < J; t += 32) {
          const unsigned long long key = pkey[(size_t)r * J + t];
          double x = 0, y = 0, sc = 0;
          if (key != 0ull) {
            const int bi = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
            const float best = __uint_as_float((unsigned)(key >> 32));
            x = (double)p.joint_det[(n0 + bi) * 3 + 0];
            y = (double)p.joint_det[(n0 + bi) * 3 + 1];
            sc = (double)best;
            if (best > 0.f) valid = 1;
          }
          if (n_person < p.max_persons) {
            out[((size_t)n_person * J + t) * 3 + 0] = x;
            out[((size_t)n_person * J + t) * 3 + 1] = y;
            out[((size_t)n_person * J + t) * 3 + 2] = sc;
          }
        }
        valid = __any_sync(kFull, valid);
        if (valid) ++n_person;                                      // :725
      }
    }
    if (lane == 0) { p.num_persons[b] = n_person; p.mutants[b] = p.cc_method == PGMP_CC_GREEDY ? 0 : mutant; }   // :506
  }
}

int max_nodes_of(const pgmp_group_params& p) { return p.max_nodes_per_image; }

}  // namespace
}  // namespace pgmp

using namespace pgmp;

extern "C" uint64_t pgmp_group_workspace_bytes(const pgmp_group_params* p) {
  if (!p || p->batch <= 0 || p->max_nodes_per_image <= 0) return 0;
  pgmp_group_params q = *p;
  q.workspace = nullptr;
  return carve(q, q.max_nodes_per_image).bytes;
}

extern "C" int pgmp_group_persons(const pgmp_group_params* p, pgmp_stream_t stream) {
  if (!p) return set_error(PGMP_ERR_INVALID, "null params");
  if (p->batch <= 0 || p->num_joints <= 0 || p->num_joints > 64 || p->max_nodes_per_image <= 0 || p->max_persons <= 0)
    return set_error(PGMP_ERR_INVALID, "bad sizes");
  if (!p->node_offsets || !p->edge_offsets || !p->joint_det || !p->node_logits || !p->person_labels ||
      !p->num_components || !p->num_kept_edges || !p->persons || !p->num_persons || !p->mutants || !p->workspace)
    return set_error(PGMP_ERR_INVALID, "null device pointer");
  if (p->num_edges > 0 && (!p->edge_index || !p->edge_logits)) return set_error(PGMP_ERR_INVALID, "null edge pointer");
  const GroupWs w = carve(*p, p->max_nodes_per_image);
  if (w.bytes > p->workspace_bytes) return set_error(PGMP_ERR_INVALID, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t book = (size_t)24 * p->max_nodes_per_image;
  const int use_smem = book <= 160 * 1024;
  if (use_smem && book > 48 * 1024)
    PGMP_CUDA(cudaFuncSetAttribute(group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)book));
  PGMP_LAUNCH(group_kernel, p->batch, kThreads, use_smem ? book : 0, st, *p, w, p->max_nodes_per_image, use_smem);
  return PGMP_OK;
}
