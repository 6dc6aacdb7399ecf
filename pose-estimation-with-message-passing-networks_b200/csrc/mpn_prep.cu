// Graph bookkeeping for the message-passing kernels: counting sort of the caller's edges into
// (source type, target node) bins -> slot order (see mpn_common.cuh).  Replaces the gathers /
// boolean-mask loops of torch_geometric's propagate and layers.py:234-251, 268-274.
#include "mpn_common.cuh"

namespace pgmp {
namespace {

constexpr uint32_t kFull = 0xffffffffu;

__global__ void __launch_bounds__(256) node_types_kernel(const int64_t* __restrict__ types, int64_t N, int per_type,
                                                          int T, int32_t* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int t = 0;
  if (per_type) {
    const int64_t v = types[n];
    t = v < 0 ? 0 : (v >= T ? T - 1 : (int)v);   // memory safety only; valid inputs are in [0, T)
  }
  out[n] = t;
}

__global__ void __launch_bounds__(256) count_bins_kernel(const int64_t* __restrict__ edge_index, int64_t E, int64_t N,
                                                          const int32_t* __restrict__ node_type,
                                                          int32_t* __restrict__ bin_count, int32_t* __restrict__ status) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t src = edge_index[e], dst = edge_index[E + e];
  if ((uint64_t)src >= (uint64_t)N || (uint64_t)dst >= (uint64_t)N) {   // the reference raises an IndexError here:
    atomicOr(status, PGMP_MPN_STATUS_BAD_EDGE);                           // the edge is dropped and the status word says so
    return;
  }
  atomicAdd(&bin_count[(int64_t)node_type[src] * N + dst], 1);
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(kFull, wi, o);
      if (lane >= o) wi += u;
    }
    s_warp[lane] = wi - w;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  const int r = s_warp[warp] + incl - v;
  __syncthreads();
  return r;
}

// one CTA per type group: exclusive scans of bin sizes and of part counts inside the group
__global__ void __launch_bounds__(1024) scan_groups_kernel(int64_t N, const int32_t* __restrict__ bin_count,
                                                            int32_t* __restrict__ bin_lstart,
                                                            int32_t* __restrict__ bin_lpart,
                                                            int32_t* __restrict__ group_total,
                                                            int32_t* __restrict__ group_parts) {
  __shared__ int s_warp[32];
  __shared__ int s_total;
  const int t = blockIdx.x;
  const int64_t chunk = ceil_div<int64_t>(N, blockDim.x);
  const int64_t b0 = (int64_t)t * N + threadIdx.x * chunk;
  const int64_t b1 = min(b0 + chunk, (int64_t)(t + 1) * N);
  int sum = 0;
  for (int64_t b = b0; b < b1; ++b) sum += bin_count[b];
  int run = block_exclusive_scan(sum, s_warp, &s_total);
  if (threadIdx.x == 0) group_total[t] = s_total;
  int psum = 0;
  for (int64_t b = b0; b < b1; ++b) {
    const int c = bin_count[b];
    bin_lstart[b] = run;
    const int np = c ? ((run + c - 1) >> 7) - (run >> 7) + 1 : 0;   // tiles the bin touches
    bin_lpart[b] = np;
    psum += np;
    run += c;
  }
  int prun = block_exclusive_scan(psum, s_warp, &s_total);
  if (threadIdx.x == 0) group_parts[t] = s_total;
  for (int64_t b = b0; b < b1; ++b) {
    const int np = bin_lpart[b];
    bin_lpart[b] = prun;
    prun += np;
  }
}

__global__ void group_offsets_kernel(int T, const int32_t* __restrict__ group_total,
                                     const int32_t* __restrict__ group_parts, int32_t* __restrict__ group_start,
                                     int32_t* __restrict__ group_pstart) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int s = 0, q = 0;
  for (int t = 0; t < T; ++t) {
    group_start[t] = s;
    group_pstart[t] = q;
    s += round_up(group_total[t], kTile);
    q += group_parts[t];
  }
  group_start[T] = s;
  group_pstart[T] = q;
}

__global__ void __launch_bounds__(256) scatter_slots_kernel(const int64_t* __restrict__ edge_index, int64_t E,
                                                             int64_t N, const int32_t* __restrict__ node_type,
                                                             const int32_t* __restrict__ bin_lstart,
                                                             const int32_t* __restrict__ group_start,
                                                             int32_t* __restrict__ bin_cursor,
                                                             int32_t* __restrict__ slot_edge) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t src = edge_index[e], dst = edge_index[E + e];
  if ((uint64_t)src >= (uint64_t)N || (uint64_t)dst >= (uint64_t)N) return;   // dropped, flagged by count_bins_kernel
  const int t = node_type[src];
  const int64_t bin = (int64_t)t * N + dst;
  const int pos = atomicAdd(&bin_cursor[bin], 1);
  slot_edge[group_start[t] + bin_lstart[bin] + pos] = (int32_t)e;
}

// order every bin by edge id (deterministic reduction order) and fill the slot endpoints
__global__ void __launch_bounds__(256) finish_bins_kernel(const int64_t* __restrict__ edge_index, int64_t E,
                                                           int64_t N, int T, const int32_t* __restrict__ bin_count,
                                                           const int32_t* __restrict__ bin_lstart,
                                                           const int32_t* __restrict__ group_start,
                                                           int32_t* __restrict__ slot_edge,
                                                           int32_t* __restrict__ slot_src,
                                                           int32_t* __restrict__ slot_dst) {
  const int64_t bin = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= (int64_t)T * N) return;
  const int c = bin_count[bin];
  if (c == 0) return;
  const int t = (int)(bin / N);
  int32_t* __restrict__ s = slot_edge + group_start[t] + bin_lstart[bin];
  for (int i = 1; i < c; ++i) {   // insertion sort; bins hold a handful of edges
    const int32_t v = s[i];
    int j = i - 1;
    while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; }
    s[j + 1] = v;
  }
  const int64_t base = s - slot_edge;
  for (int i = 0; i < c; ++i) {
    const int64_t e = s[i];
    slot_src[base + i] = (int32_t)edge_index[e];
    slot_dst[base + i] = (int32_t)edge_index[E + e];
  }
}

// Step-invariant run bookkeeping of the tensor-core step kernel, one CTA per 128-slot tile: a run = the consecutive rows
// of one (type, target) bin inside the tile (pad rows are runs of their own).  Per row: the part row its run is stored to
// (last row of the run only), the run's extent inside the row's warp (for the segmented maximum of the attention
// logits); per tile: the rows at which the 32 segments of the run reduction begin (segment k starts at the first run
// start at or after row 4 k, so that every run is reduced by one thread per column group, rows in order).
__global__ void __launch_bounds__(kTile) slot_runs_kernel(int64_t N, int T, const int32_t* __restrict__ slot_dst,
                                                           const int32_t* __restrict__ bin_lstart,
                                                           const int32_t* __restrict__ bin_lpart,
                                                           const int32_t* __restrict__ group_start,
                                                           const int32_t* __restrict__ group_pstart,
                                                           int2* __restrict__ slot_run, int32_t* __restrict__ tile_seg) {
  __shared__ unsigned s_mask[4];
  const int tile = blockIdx.x, row = threadIdx.x, lane = row & 31;
  const int slot0 = tile * kTile;
  if (slot0 >= group_start[T]) return;
  int t = 0;
  while (t + 1 < T && slot0 >= group_start[t + 1]) ++t;
  const int64_t sl = (int64_t)slot0 + row;
  const int dst = slot_dst[sl];
  const bool valid = dst >= 0;
  const bool is_start = !valid || row == 0 || slot_dst[sl - 1] != dst;
  const bool is_last = valid && (row == kTile - 1 || slot_dst[sl + 1] != dst);
  int ctl = -1;
  if (is_last) {
    const int64_t bin = (int64_t)t * N + dst;
    const int first_slot = group_start[t] + bin_lstart[bin];
    ctl = group_pstart[t] + bin_lpart[bin] + (tile - (first_slot >> 7));
  }
  const unsigned starts = __ballot_sync(kFull, is_start);
  const unsigned upto = kFull >> (31 - lane);
  const int seg_first = 31 - __clz((starts | 1u) & upto);
  const unsigned above = starts & ~upto;
  const int seg_last = above ? __ffs(above) - 2 : 31;
  slot_run[sl] = make_int2(ctl, seg_first | seg_last << 5 | (is_start ? 1 << 10 : 0) | (valid ? 1 << 11 : 0));
  const unsigned vstarts = __ballot_sync(kFull, valid && is_start);
  if (lane == 0) s_mask[row >> 5] = vstarts;
  __syncthreads();
  if (row < 32) {
    int first = row == 0 ? 0 : kTile;
    if (row > 0) {
      const int from = 4 * row;
      for (int w = from >> 5; w < 4; ++w) {
        unsigned m = s_mask[w];
        if (w == (from >> 5)) m &= kFull << (from & 31);
        if (m) { first = 32 * w + __ffs(m) - 1; break; }
      }
    }
    tile_seg[tile * 32 + row] = first;
  }
}

}  // namespace

int mpn_prepare_graph(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st) {
  const int64_t N = p.num_nodes, E = p.num_edges;
  const int T = p.num_types;
  const int64_t bins = (int64_t)T * N;
  PGMP_CUDA(cudaMemsetAsync(w.status, 0, sizeof(int32_t), st));
  PGMP_CUDA(cudaMemsetAsync(w.bin_count, 0, sizeof(int32_t) * bins, st));
  PGMP_CUDA(cudaMemsetAsync(w.bin_cursor, 0, sizeof(int32_t) * bins, st));
  PGMP_CUDA(cudaMemsetAsync(w.slot_edge, 0xff, sizeof(int32_t) * w.max_slots, st));
  PGMP_CUDA(cudaMemsetAsync(w.slot_src, 0xff, sizeof(int32_t) * w.max_slots, st));
  PGMP_CUDA(cudaMemsetAsync(w.slot_dst, 0xff, sizeof(int32_t) * w.max_slots, st));
  PGMP_LAUNCH(node_types_kernel, (unsigned)ceil_div<int64_t>(N, 256), 256, 0, st, p.node_types, N, p.per_type, T,
              w.node_type);
  if (E > 0)
    PGMP_LAUNCH(count_bins_kernel, (unsigned)ceil_div<int64_t>(E, 256), 256, 0, st, p.edge_index, E, N, w.node_type,
                w.bin_count, w.status);
  PGMP_LAUNCH(scan_groups_kernel, T, 1024, 0, st, N, w.bin_count, w.bin_lstart, w.bin_lpart, w.group_total,
              w.group_parts);
  PGMP_LAUNCH(group_offsets_kernel, 1, 32, 0, st, T, w.group_total, w.group_parts, w.group_start, w.group_pstart);
  if (E > 0) {
    PGMP_LAUNCH(scatter_slots_kernel, (unsigned)ceil_div<int64_t>(E, 256), 256, 0, st, p.edge_index, E, N,
                w.node_type, w.bin_lstart, w.group_start, w.bin_cursor, w.slot_edge);
    PGMP_LAUNCH(finish_bins_kernel, (unsigned)ceil_div<int64_t>(bins, 256), 256, 0, st, p.edge_index, E, N, T,
                w.bin_count, w.bin_lstart, w.group_start, w.slot_edge, w.slot_src, w.slot_dst);
    if (p.precision == PGMP_PRECISION_TC)
      PGMP_LAUNCH(slot_runs_kernel, (unsigned)(w.max_slots / kTile), kTile, 0, st, N, T, w.slot_dst, w.bin_lstart,
                  w.bin_lpart, w.group_start, w.group_pstart, w.slot_run, w.tile_seg);
  }
  return PGMP_OK;
}

}  // namespace pgmp
