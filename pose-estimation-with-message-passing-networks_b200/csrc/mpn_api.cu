// C-ABI entry points of the message-passing network (include/pgmp.h).
#include "mpn_common.cuh"

namespace pgmp {
int mpn_validate_heads(const pgmp_mpn_params& p);
int mpn_forward_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st);

static int validate_mpn(const pgmp_mpn_params* p) {
  if (!p) return set_error(PGMP_ERR_INVALID, "null params");
  if (p->num_nodes <= 0 || p->num_edges < 0) return set_error(PGMP_ERR_INVALID, "bad graph size N=%lld E=%lld", (long long)p->num_nodes, (long long)p->num_edges);
  if (p->num_nodes > (1ll << 30) || p->num_edges > (1ll << 30)) return set_error(PGMP_ERR_INVALID, "graph too large for 32-bit slots");
  if (p->dim != kD) return set_error(PGMP_ERR_INVALID, "NODE/EDGE_FEATURE_DIM must be 64, got %d", p->dim);
  if (p->steps < 1) return set_error(PGMP_ERR_INVALID, "STEPS must be >= 1");
  if (p->aux_loss_steps < 0) return set_error(PGMP_ERR_INVALID, "AUX_LOSS_STEPS < 0");
  if (p->per_type) {
    if (p->num_types < 1 || p->num_types > 17 || p->num_type_mlps != 17) return set_error(PGMP_ERR_INVALID, "per_type: num_types %d, mlps %d", p->num_types, p->num_type_mlps);
    if (p->update_hier) {
      if (!p->hier || (p->num_types != 17 && p->num_types != 14))
        return set_error(PGMP_ERR_INVALID, "hierarch_mlp needs its weights and 17 or 14 joint types (layers.py:96)");
    } else if (!p->has_update_mlp || !p->wu) {
      return set_error(PGMP_ERR_INVALID, "per_type needs update_mlp");
    }
  } else {
    if (p->num_types != 1 || p->num_type_mlps != 1) return set_error(PGMP_ERR_INVALID, "agnostic: num_types must be 1");
    if (p->attn != PGMP_ATTN_NONE) return set_error(PGMP_ERR_INVALID, "attention aggregation needs per_type");
  }
  if (p->aggr < 0 || p->aggr > 2 || p->attn < 0 || p->attn > 2) return set_error(PGMP_ERR_INVALID, "aggr %d attn %d", p->aggr, p->attn);
  if (p->attn && (!p->wa || !p->ba)) return set_error(PGMP_ERR_INVALID, "attention without attn_net weights");
  if (p->update_hier && !p->per_type) return set_error(PGMP_ERR_INVALID, "hierarch_mlp needs per_type");
  if (p->has_update_mlp && !p->update_hier && (!p->wu || !p->bu)) return set_error(PGMP_ERR_INVALID, "update_mlp weights missing");
  if (p->skip && !p->w1_e0) return set_error(PGMP_ERR_INVALID, "skip without w1_e0");
  if (!p->x || !p->node_types || !p->w1_dst || !p->w1_src || !p->w1_e || !p->b1 || !p->w2 || !p->b2 || !p->wm_x ||
      !p->wm_e || !p->bm || !p->node_logits || !p->class_logits || !p->workspace)
    return set_error(PGMP_ERR_INVALID, "null device pointer");
  if (p->num_edges > 0 && (!p->edge_attr || !p->edge_index || !p->edge_logits)) return set_error(PGMP_ERR_INVALID, "null edge pointer");
  if (p->num_classes < 1 || p->num_classes > kD) return set_error(PGMP_ERR_INVALID, "num_classes %d", p->num_classes);
  if (p->precision != PGMP_PRECISION_FP32 && p->precision != PGMP_PRECISION_TC) return set_error(PGMP_ERR_INVALID, "precision %d", p->precision);
  return mpn_validate_heads(*p);
}
}  // namespace pgmp

using namespace pgmp;

extern "C" uint64_t pgmp_mpn_workspace_bytes(const pgmp_mpn_params* p) {
  if (!p || p->num_nodes <= 0 || p->num_edges < 0 || p->num_types < 1) return 0;
  pgmp_mpn_params q = *p;
  q.workspace = nullptr;
  return carve_mpn(q).bytes;
}

extern "C" int pgmp_mpn_forward(const pgmp_mpn_params* p, pgmp_stream_t stream) {
  int rc = validate_mpn(p);
  if (rc != PGMP_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const MpnWorkspace w = carve_mpn(*p);
  if (w.bytes > p->workspace_bytes)
    return set_error(PGMP_ERR_INVALID, "workspace too small: %llu < %llu", (unsigned long long)p->workspace_bytes,
                     (unsigned long long)w.bytes);
  if ((rc = mpn_prepare_graph(*p, w, st)) != PGMP_OK) return rc;
  if (p->precision == PGMP_PRECISION_TC) return mpn_forward_tc(*p, w, st);
  return mpn_forward_simt(*p, w, st);
}
