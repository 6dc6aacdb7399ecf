// Data layout shared by the message-passing kernels (mpn_prep.cu, mpn_simt.cu, mpn_tc.cu).
//
// Edge features live in "slot" order, not in the caller's edge order: edges are grouped by
// (source type, target node) -- type-major -- so that
//   * a 128-row tile has one source type  -> one message weight matrix per tile (layers.py:264-274)
//   * the edges of one (target, source type) pair are consecutive -> the per-(node, type) softmax /
//     sum / max of layers.py:234-251 is a reduction over consecutive rows, no atomics
// Every type group is padded to a multiple of 128 slots (pad slots have edge = -1).  Within a bin the
// slots are ordered by the caller's edge id, which makes every floating-point reduction order
// deterministic.  A bin whose rows straddle a tile boundary is reduced per tile into "parts" that
// the node update merges (flash-style for the attention softmax).
#pragma once

#include "common.cuh"

namespace pgmp {

constexpr int kD = 64;          // NODE_FEATURE_DIM = EDGE_FEATURE_DIM = EDGE_FEATURE_HIDDEN
constexpr int kTile = 128;      // rows (slots / nodes) per CTA tile
constexpr int kTileP = kTile + 1;  // padded row count of transposed shared-memory tiles

struct MpnWorkspace {
  int32_t* status;        // [1] FIRST word of the workspace: PGMP_MPN_STATUS_* bits set by the forward (0 = input was well-formed)
  // graph bookkeeping
  int32_t* node_type;     // [N] clamped to [0, T)
  int32_t* bin_count;     // [T*N] edges per (type, target)
  int32_t* bin_cursor;    // [T*N] scatter cursors
  int32_t* bin_lstart;    // [T*N] first slot of the bin relative to its type group
  int32_t* bin_lpart;     // [T*N] first part row of the bin relative to its type group
  int32_t* group_total;   // [T] slots used by the group (before padding)
  int32_t* group_parts;   // [T]
  int32_t* group_start;   // [T+1] first slot of the group (multiple of 128); [T] = total padded slots
  int32_t* group_pstart;  // [T+1] first part row of the group; [T] = total parts
  int32_t* slot_edge;     // [S] caller's edge id or -1
  int32_t* slot_src;      // [S]
  int32_t* slot_dst;      // [S]
  int2* slot_run;         // [S] tensor-core mode, step-invariant run bookkeeping of the step kernel: .x = part row stored by the
                          //     last row of a run (else -1), .y = first lane | last lane << 5 of the row's run inside its warp
                          //     | run start << 10 | valid << 11
  int32_t* tile_seg;      // [S / 128][32] first row of the k-th row segment of the tile's run reduction
  // features
  float* h0;              // [N][64] node embedding
  float* h;               // [N][64] current node feature
  float* h0_img;          // tensor-core mode: bf16 hi/lo operand images of h0 / h, one 32 KB image per 128 nodes
  float* h_img;
  float* upd_partial;     // tensor-core mode: [groups][ceil128(N)][64] partial node updates (type groups)
  float* g;               // [S][64] current edge feature (slot order), updated in place
  float* c0;              // [S][64] W1_e0 * g0 + b1 (skip only)
  float* tab_p;           // [N][64] W1_dst * x_i (+ b1 when !skip)
  float* tab_q;           // [N][64] W1_src * x_j
  float* tab_r;           // [T][N][64] Wm_x[t] * x_i + bm[t]
  float* part_val;        // [P][64] per-part reduced message (weighted sum / sum / max)
  float* part_mx;         // [P] per-part max attention logit
  float* part_se;         // [P] per-part sum of exp(logit - max)
  uint64_t max_slots, max_parts;
  uint64_t bytes;
};

// type groups of the tensor-core node update: (node tiles x groups) CTAs should fill the GPU once -- two CTAs per SM
// are resident -- so small graphs get one group per type (parallelism) and large ones few groups (less traffic)
inline int mpn_update_groups(const pgmp_mpn_params& p) {
  const int64_t tiles = (p.num_nodes + kTile - 1) / kTile;
  int64_t g = tiles > 0 ? (2 * 148) / tiles : p.num_types;
  if (g < 1) g = 1;
  if (g > p.num_types) g = p.num_types;
  const int64_t per = (p.num_types + g - 1) / g;          // types per group; drop the groups that would be empty
  return (int)((p.num_types + per - 1) / per);
}

inline MpnWorkspace carve_mpn(const pgmp_mpn_params& p) {
  Carver c(p.workspace);
  MpnWorkspace w;
  const uint64_t N = (uint64_t)p.num_nodes, E = (uint64_t)p.num_edges, T = (uint64_t)p.num_types;
  w.max_slots = round_up<uint64_t>(E, kTile) + T * kTile;       // every group padded to 128
  const uint64_t bins = T * N;
  w.max_parts = (E < bins ? E : bins) + w.max_slots / kTile;      // non-empty bins + tile crossings
  w.status = c.take<int32_t>(1);
  w.node_type = c.take<int32_t>(N);
  w.bin_count = c.take<int32_t>(bins);
  w.bin_cursor = c.take<int32_t>(bins);
  w.bin_lstart = c.take<int32_t>(bins);
  w.bin_lpart = c.take<int32_t>(bins);
  w.group_total = c.take<int32_t>(T);
  w.group_parts = c.take<int32_t>(T);
  w.group_start = c.take<int32_t>(T + 1);
  w.group_pstart = c.take<int32_t>(T + 1);
  w.slot_edge = c.take<int32_t>(w.max_slots);
  w.slot_src = c.take<int32_t>(w.max_slots);
  w.slot_dst = c.take<int32_t>(w.max_slots);
  w.slot_run = c.take<int2>(p.precision == PGMP_PRECISION_TC ? w.max_slots : 0);
  w.tile_seg = c.take<int32_t>(p.precision == PGMP_PRECISION_TC ? w.max_slots / kTile * 32 : 0);
  w.h0 = c.take<float>(N * kD);
  w.h = c.take<float>(N * kD);
  const uint64_t Np = round_up<uint64_t>(N, kTile);
  const bool tc = p.precision == PGMP_PRECISION_TC;
  w.h0_img = c.take<float>(tc ? Np * kD : 0);
  w.h_img = c.take<float>(tc ? Np * kD : 0);
  w.upd_partial = c.take<float>(tc ? (uint64_t)mpn_update_groups(p) * Np * kD : 0);
  w.g = c.take<float>(w.max_slots * kD);
  w.c0 = c.take<float>(p.skip ? w.max_slots * kD : 0);
  w.tab_p = c.take<float>(N * kD);
  w.tab_q = c.take<float>(N * kD);
  w.tab_r = c.take<float>(T * N * kD);
  w.part_val = c.take<float>(w.max_parts * kD);
  w.part_mx = c.take<float>(w.max_parts);
  w.part_se = c.take<float>(w.max_parts);
  w.bytes = c.bytes();
  return w;
}

// host-side launchers implemented in the .cu files
int mpn_prepare_graph(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st);
int mpn_forward_simt(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st);

}  // namespace pgmp
