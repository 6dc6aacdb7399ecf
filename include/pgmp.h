/*
 * pgmp.h -- C ABI of libpgmp.so, the B200-native (sm_100a) post-backbone grouping path of
 * nibox/Pose-Estimation-with-Message-Passing-Networks.
 *
 * The reference has no FFI of its own (it is pure Python over torch / torch_geometric /
 * torch_scatter / torch_cluster and one missing native module); each entry point below names the
 * reference Python interface it stands behind (paths relative to the reference root).  The Python
 * host (pgmp_b200/_native.py, ctypes) is the only caller; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer marked "device" is a CUDA device pointer on the
 *     current device; the caller (torch) owns every buffer including workspaces (size queries
 *     below); the library never frees or retains a pointer after the call returns.
 *   - all work is enqueued on `stream` (a cudaStream_t); no entry point synchronises.
 *   - return value: 0 = ok, negative = error (PGMP_ERR_*); text via pgmp_last_error()
 *     (thread-local).  No exception crosses the ABI.  There is no CPU fallback.
 */
#ifndef PGMP_H_
#define PGMP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGMP_VERSION 100

#define PGMP_OK 0
#define PGMP_ERR_INVALID (-1)     /* bad argument / unsupported combination */
#define PGMP_ERR_CUDA (-2)        /* a CUDA runtime call failed */

typedef void* pgmp_stream_t;      /* cudaStream_t */

int pgmp_version(void);
const char* pgmp_last_error(void);
/* cumulative number of kernels this library has launched in this process (bench.py: gpu_launches) */
uint64_t pgmp_kernel_launches(void);
/* Measurement support (the reference only has unsynchronised time.clock() pairs,
 * src/Models/PoseEstimation/PoseEstimation.py:206,239): while enabled, every kernel launch is bracketed by
 * CUDA events on its stream.  pgmp_profile_collect synchronises the device and writes one line per kernel,
 * "<name> <launches> <total_ms>\n", into `out` (returns bytes written) and clears the records. */
void pgmp_profile_enable(int on);
int pgmp_profile_collect(char* out, int size);

/* ------------------------------------------------------------------------------------------------
 * Graph constructor -- replaces NaiveGraphConstructor.construct_graph(), inference branch
 * (src/graph_constructor/ConstructGraph.py:46-68, 100-103, 206-249) including
 * joint_det_from_scoremap (:1161-1196), non_maximum_suppression (src/Utils/Utils.py:15-20),
 * cat_unique (:1199-1209), knn_mpn_graph (:363-368), fully_connected_mpn_graph (:376-381) and
 * _construct_mpn_graph (:251-325).
 *
 * Two calls with ONE host read in between (the reference syncs ~4x per image):
 *   pgmp_gc_detect : NMS + per-joint top-k / threshold candidates + candidate graph degrees
 *                    -> device counts; the host reads `counts` to size the exact outputs
 *   pgmp_gc_emit   : joint_det, scores, batch index, tags, node features, edge index, edge attr
 * ---------------------------------------------------------------------------------------------- */

#define PGMP_GRAPH_KNN 0
#define PGMP_GRAPH_FULLY 1
#define PGMP_EDGE_FEAT_POSITION 1
#define PGMP_EDGE_FEAT_TYPE 2

/* bits of counts[2 + 2*batch] (see pgmp_gc_detect) */
#define PGMP_GC_FLAG_CAND_OVERFLOW 1   /* more NMS maxima than cand_capacity for some (image, joint) */
#define PGMP_GC_FLAG_DET_OVERFLOW 2    /* more detections than max_det_per_type for some (image, joint) */
#define PGMP_GC_FLAG_NODE_OVERFLOW 4   /* more nodes than max_nodes for some image */
#define PGMP_GC_FLAG_TOO_FEW 8         /* no-threshold path: the map has fewer than top_k pixels (CG.py:1193 assert); fewer than top_k positive maxima is NOT an error: the block is padded with zero-score pixels like torch.topk, CG.py:1187-1189 */

typedef struct pgmp_gc_params {
  int32_t batch, num_joints, height, width;  /* scoremaps [B,J,H,W] float32, contiguous */
  int32_t pool_kernel;                       /* GC.POOL_KERNEL_SIZE, odd, <= 9 */
  int32_t top_k;                             /* GC.HYBRID_K (threshold path) or 20 (CG.py:1185) */
  int32_t use_threshold;                     /* 1: DETECT_THRESHOLD <= 1.5 (CG.py:28); 0: no-threshold path */
  float threshold;                           /* GC.DETECT_THRESHOLD, must be > 0 */
  int32_t graph_type;                        /* PGMP_GRAPH_* (GC.GRAPH_TYPE) */
  int32_t knn_k;                             /* 50 (CG.py:365) */
  int32_t edge_features;                     /* PGMP_EDGE_FEAT_* bit set (GC.EDGE_FEATURES_TO_USE) */
  float norm_factor;                         /* max(H,W) if GC.NORM_NODE_DISTANCE else 1 (CG.py:311-314) */
  int32_t cand_capacity;                     /* NMS-maxima list capacity per (image, joint) */
  int32_t max_det_per_type;                  /* detections kept per (image, joint) */
  int32_t max_nodes;                         /* nodes per image, multiple of 32 */
  const float* scoremaps;                    /* device */
  const float* mask;                         /* device [B,H,W] float32 crowd mask or NULL (GC.MASK_CROWDS) */
  void* workspace;                           /* device, pgmp_gc_workspace_bytes() bytes, 256-B aligned */
  uint64_t workspace_bytes;
} pgmp_gc_params;

uint64_t pgmp_gc_workspace_bytes(const pgmp_gc_params* p);

/* counts: device int64[2 + 2*batch + 1] = { sum N, sum E, N_0..N_{B-1}, E_0..E_{B-1}, flags }. */
int pgmp_gc_detect(const pgmp_gc_params* p, int64_t* counts, pgmp_stream_t stream);

typedef struct pgmp_gc_outputs {
  int64_t total_nodes, total_edges;          /* as read back from counts[0..1] */
  const float* features;                     /* device (or pinned host, read in place) [B,C,H,W] float32, arbitrary strides (elements) */
  int64_t feat_stride_b, feat_stride_c, feat_stride_y, feat_stride_x;
  int32_t channels;
  const float* tagmaps;                      /* device (or pinned host) [B,J,H,W] or [B,J,H,W,T] float32 contiguous */
  int32_t tag_dim;                           /* T (1 for [B,J,H,W]) */
  float* x;                                  /* [sum N, C]            CG.py:265,269 */
  float* edge_attr;                          /* [sum E, F]            CG.py:305-325 */
  int64_t* edge_index;                       /* [2, sum E] global ids CG.py:222-228 */
  int64_t* joint_det;                        /* [sum N, 3] (x,y,type) CG.py:1180-1182 */
  float* joint_scores;                       /* [sum N]               CG.py:1183 */
  int64_t* batch_index;                      /* [sum N]               CG.py:207 */
  float* joint_tags;                         /* [sum N, T]            CG.py:103 */
} pgmp_gc_outputs;

int pgmp_gc_emit(const pgmp_gc_params* p, const pgmp_gc_outputs* o, pgmp_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Message-passing network -- replaces NodeClassificationMPNSimple.forward
 * (src/Models/MessagePassingNetwork/NodeClassificationMPNSimple.py:62-97) with MPLayer
 * (layers.py:32-86) or TypeAwareMPNLayer (layers.py:157-274), _make_mlp embeddings and heads
 * (layers.py:8-29), eval-mode BatchNorm folded by the host.
 * ---------------------------------------------------------------------------------------------- */

#define PGMP_MAX_LAYERS 6
#define PGMP_AGGR_ADD 0
#define PGMP_AGGR_MAX 1
#define PGMP_AGGR_MEAN 2
#define PGMP_ATTN_NONE 0       /* AGGR_SUB "None" */
#define PGMP_ATTN_SHARED 1     /* "node_edge_attn" */
#define PGMP_ATTN_PER_TYPE 2   /* "node_edge_attn_per_type" */
#define PGMP_PRECISION_FP32 0  /* SIMT fp32 everywhere (parity mode) */
#define PGMP_PRECISION_TC 1    /* tcgen05 tensor-core message-passing steps, fp32 accumulation */

/* One _make_mlp chain with eval BatchNorm folded into the following Linear. */
/* Node features at the candidates straight from the backbone feature map: replaces
 * self.feature_gather(feat) (nn.Conv2d(Cin, Cout, 3, 1, 1), PoseEstimation.py:64-66, 79, 341), the bilinear
 * interpolate(..., size, align_corners=False) of PoseEstimation.py:442-450 and the gather of ConstructGraph.py:265,269 */
typedef struct pgmp_gather_conv_params {
  const float* features;                     /* device (or pinned host) [B, Cin, h, w] float32, arbitrary strides (elements) */
  int64_t feat_stride_b, feat_stride_c, feat_stride_y, feat_stride_x;
  int32_t cin, height, width;                /* Cin, h, w of the backbone feature map */
  int32_t cout;                              /* output channels (<= 128) */
  int32_t out_height, out_width;             /* size the reference interpolates to (the heatmap size) */
  const float* weight_t;                     /* device [(ky * 3 + kx) * Cin + ci][Cout] = conv weight transposed */
  const float* bias;                         /* device [Cout] */
  const int64_t* joint_det;                  /* device [N, 3] (x, y, type) as written by pgmp_gc_emit */
  const int64_t* batch_index;                /* device [N] */
  int64_t num_nodes;
  float* x;                                  /* device [N, Cout] */
} pgmp_gather_conv_params;
/* x = interpolate(feature_gather(feat))[:, y, x] at the candidates only; stream-ordered after pgmp_gc_emit */
int pgmp_gc_gather_conv(const pgmp_gather_conv_params* p, pgmp_stream_t stream);

/* Reverse pass of pgmp_gc_gather_conv (training end to end through ConvUpsampleFeatures: gradients of the feature_gather
 * convolution and of the backbone map, PoseEstimation.py:64-66, train.py:232).  With P[n][(ky 3 + kx) Cin + ci] the
 * interpolated 3 x 3 x Cin input patch of node n (x = bias + P weight_t):
 *   pgmp_gc_gather_conv_patches   writes P [N][9 Cin] (the forward's arithmetic); the caller forms d weight_t = P^T d x,
 *                                 d bias = sum d x and d P = d x weight_t^T with its library GEMM;
 *   pgmp_gc_gather_conv_backward  adds d P into d_features ([B, Cin, h, w], the strides of p->features, zero-filled by the
 *                                 caller) through the transposed interpolation; no atomics: one CTA per image walks its
 *                                 nodes in node order, so overlapping neighbourhoods are summed in a fixed order.
 * p->x / weight_t / bias are not read by either call. */
int pgmp_gc_gather_conv_patches(const pgmp_gather_conv_params* p, float* patches, pgmp_stream_t stream);
int pgmp_gc_gather_conv_backward(const pgmp_gather_conv_params* p, const float* d_patches, int32_t batch, float* d_features,
                                 pgmp_stream_t stream);

/* Scoremap assembly in front of the NMS -- hr_process_output (src/Models/HigherHRNet/hrnet.py:587-611):
 *   up = interpolate(stage1 [B, C1, h, w], size = (H, W), bilinear, align_corners = False)
 *   scoremaps [B, J, H, W] = (stage2 + up[:, :J]) / 2 (PGMP_ASSEMBLE_AVG) or up[:, :J] (PGMP_ASSEMBLE_SMALL)
 *   tags [B, C1 - J, H, W] = up[:, J:]                (skipped when tags == NULL)
 * in one pass over the outputs.  All pointers device, contiguous float32. */
#define PGMP_ASSEMBLE_AVG 0
#define PGMP_ASSEMBLE_SMALL 1
int pgmp_gc_assemble_scoremaps(const float* stage1, const float* stage2, int32_t batch, int32_t channels1, int32_t num_joints,
                               int32_t h, int32_t w, int32_t H, int32_t W, int32_t mode, float* scoremaps, float* tags,
                               pgmp_stream_t stream);

/* The same assembly fused into the detection: pgmp_gc_detect_fused() is pgmp_gc_detect() on the scoremaps
 *   A_t = (stage2[t] + up(stage1[t])[:, :J]) / 2      (mode AVG; SMALL: up(stage1[t])[:, :J])
 *   scoremaps = A_0                                                    (n_terms 1: hr_process_output, hrnet.py:587-611)
 *   scoremaps[b, j, y, x] = (A_0[b, j, y, x] + A_1[b, flip_index[j], y, W - 1 - x]) / 2
 *                                                                      (n_terms 2: the FLIP_TEST average of one scale,
 *                                                                       PoseEstimation.py:343-364, 377-402, multi_scales_testing.py:162)
 * evaluated by the NMS kernel's loader warps straight into its shared-memory row ring: the assembled map is never
 * written to or re-read from HBM unless scoremaps_out asks for it (the pose-assembly tail reads it).  The scores of the
 * detections are re-evaluated at their pixels with the same arithmetic.  p->scoremaps is ignored; width % 4 == 0,
 * width <= 1024 and the threshold path (use_threshold 1) or a non-NULL scoremaps_out are required (PGMP_ERR_INVALID
 * otherwise: run pgmp_gc_assemble_scoremaps + pgmp_gc_detect).  Bit-identical to those two calls. */
typedef struct pgmp_gc_assembly {
  const float* stage1[2];                    /* device [B, channels1, h, w]; [1] = outputs for the flipped image or NULL */
  const float* stage2[2];                    /* device [B, num_joints, H, W] (mode AVG) */
  int32_t channels1, h, w, mode, n_terms;
  int32_t flip_index[32];                    /* n_terms 2: joint permutation (FLIP_CONFIG), num_joints entries */
  float* scoremaps_out;                      /* optional device [B, J, H, W] */
} pgmp_gc_assembly;
int pgmp_gc_detect_fused(const pgmp_gc_params* p, const pgmp_gc_assembly* a, int64_t* counts, pgmp_stream_t stream);
/* ... and the tags of the detections (ConstructGraph.py:103) without the up-sampled tag maps: joint_tags[n] =
 * up(stage1)[batch_index[n], num_joints + type, y, x] for the (x, y, type) rows pgmp_gc_emit wrote (tag dimension 1:
 * stage 1 = num_joints heatmaps + num_joints tag maps); bit-identical to indexing the up-sampled maps. */
int pgmp_gc_gather_stage_tags(const float* stage1, int32_t channels1, int32_t num_joints, int32_t h, int32_t w, int32_t H,
                              int32_t W, const int64_t* joint_det, const int64_t* batch_index, int64_t num_nodes,
                              float* joint_tags, pgmp_stream_t stream);

/* Reverse of the node-feature gather of pgmp_gc_emit (x[n, :] = features[b, :, y, x], ConstructGraph.py:265, 269) under
 * autograd -- end-to-end training, train.py:232: d_features[b, :, y, x] = sum of grad_x[n, :] over the nodes at that pixel
 * (candidates of different joint types can share one), summed in node order without atomics.  d_features ([B, C, H, W],
 * strides in elements) must be zero-filled by the caller; pixels without a node are not touched.  joint_det / batch_index are
 * the arrays pgmp_gc_emit wrote (batch_index ascending: the nodes of an image are contiguous). */
int pgmp_gc_gather_backward(const float* grad_x, const int64_t* joint_det, const int64_t* batch_index, int64_t num_nodes,
                            int32_t channels, float* d_features, int64_t stride_b, int64_t stride_c, int64_t stride_y,
                            int64_t stride_x, pgmp_stream_t stream);

typedef struct pgmp_mlp {
  int32_t n_layers;
  int32_t dims[PGMP_MAX_LAYERS + 1];   /* dims[0] = input width, dims[l+1] = output width of layer l */
  int32_t relu[PGMP_MAX_LAYERS];       /* ReLU after layer l */
  const float* wt[PGMP_MAX_LAYERS];    /* device [dims[l]][dims[l+1]] = Linear.weight transposed */
  const float* bias[PGMP_MAX_LAYERS];  /* device [dims[l+1]] */
  int32_t post_relu;                   /* END_WITH_RELU */
  const float* post_scale;             /* trailing BatchNorm as scale/shift, or NULL */
  const float* post_shift;
} pgmp_mlp;

typedef struct pgmp_mpn_params {
  int64_t num_nodes, num_edges;
  const float* x;                      /* device [N, node_in], strides in elements */
  int64_t x_stride_n, x_stride_c;
  const float* edge_attr;              /* device [E, edge_in] contiguous */
  const int64_t* edge_index;           /* device [2, E]; row 0 = source j, row 1 = target i (layers.py:210) */
  const int64_t* node_types;           /* device [N], already mapped by sum_node_types (utils.py:6-19) */

  int32_t dim;                         /* NODE_FEATURE_DIM = EDGE_FEATURE_DIM = EDGE_FEATURE_HIDDEN, must be 64 */
  int32_t per_type;                    /* 0: MPLayer, 1: TypeAwareMPNLayer */
  int32_t num_types;                   /* aggregation slots per node (1 if !per_type) */
  int32_t num_type_mlps;               /* message weight sets (17 if per_type else 1) */
  int32_t skip, steps, aux_loss_steps; /* MPN.SKIP / STEPS / AUX_LOSS_STEPS */
  int32_t aggr;                        /* PGMP_AGGR_* */
  int32_t attn;                        /* PGMP_ATTN_* */
  int32_t has_update_mlp;              /* per_type: always 1; agnostic: USE_NODE_UPDATE_MLP */
  int32_t update_hier;                 /* 1: UPDATE_TYPE hierarch_mlp (layers.py:89-128): `hier` replaces wu / bu; num_types 17 or 14 */
  int32_t num_classes;                 /* width of the classification head */
  int32_t precision;                   /* PGMP_PRECISION_* */

  pgmp_mlp node_emb, edge_emb, edge_head, node_head, class_head;
  /* message-passing layer, all device float32, input-major ("transposed") matrices.
   * nd = dim * (skip ? 2 : 1) is the node-feature width seen by the layer. */
  const float* w1_dst;                 /* [nd][dim]   mlp_edge.0 columns of x_i (target) */
  const float* w1_src;                 /* [nd][dim]   mlp_edge.0 columns of x_j (source) */
  const float* w1_e0;                  /* [dim][dim]  mlp_edge.0 columns of the initial edge feature (skip) or NULL */
  const float* w1_e;                   /* [dim][dim]  mlp_edge.0 columns of the current edge feature */
  const float* b1;                     /* [dim] */
  const float* w2;                     /* [dim][dim]  mlp_edge.2 */
  const float* b2;
  const float* wm_x;                   /* [num_type_mlps][nd][dim]  mlp_node columns of x_i */
  const float* wm_e;                   /* [num_type_mlps][dim][dim] mlp_node columns of the updated edge feature */
  const float* bm;                     /* [num_type_mlps][dim] */
  const float* wa;                     /* [dim][attn_cols] attn_net.0 (attn_cols = 1 or 17) or NULL */
  const float* ba;                     /* [attn_cols] */
  const float* wu;                     /* [num_types*dim][dim] update_mlp.0 or NULL */
  const float* hier;                   /* hierarch_mlp: 7 first-layer, 6 second-layer and the final Linear as [out][in] weight + bias, back to back */
  const float* bu;
  /* PGMP_PRECISION_TC only: the per-edge weight matrices in Linear.weight layout [out][in] (K contiguous),
   * split as bf16 hi = bf16(W), lo = bf16(W - hi); device uint16 (bf16 bit patterns). */
  const void* tc_w1_e;                 /* [2][dim][dim]                 (hi, lo) of mlp_edge.0 current-edge columns */
  const void* tc_w2;                   /* [2][dim][dim]                 mlp_edge.2 */
  const void* tc_wm_e;                 /* [num_type_mlps][2][dim][dim]  mlp_node edge columns */
  const void* tc_wtab;                 /* [2 + num_types][K/64][2][64][64] pre-swizzled (SWIZZLE_128B) tile images of the per-node table weights */
  const void* tc_wu;                   /* [num_types][2][dim][dim]      update_mlp.0 columns of type t, or NULL */
  const void* tc_wnemb;                /* node embedding 128->128->64->64: [2][128][128], [2][64][128], [2][64][64] back to back, or NULL */
  const void* tc_wemb;                 /* [edge_emb.n_layers][2][64][64] edge embedding layers, zero-padded to 64x64, or NULL */
  const void* tc_w1_e0;                /* [2][dim][dim] mlp_edge.0 initial-edge columns (skip) or NULL */
  const void* tc_wheads;               /* node / class heads 64->64->32->{1,J}: node W1 [2][64][64], class W1 [2][64][64], node W2 [2][32][64], class W2 [2][32][64], or NULL */
  const void* tc_wh1;                  /* [2][64][64] edge head layer 0 (BatchNorm folded), or NULL */
  const void* tc_wh2;                  /* [2][32][64] edge head layer 1, or NULL */

  /* outputs: n_out = aux_loss_steps + 1 predictions (NodeClassificationMPNSimple.py:81-84) */
  float* edge_logits;                  /* [n_out][E] */
  float* node_logits;                  /* [n_out][N] */
  float* class_logits;                 /* [n_out][N][num_classes] */
  void* workspace;                     /* device, pgmp_mpn_workspace_bytes() bytes, 256-B aligned */
  uint64_t workspace_bytes;
} pgmp_mpn_params;

/* Self-test of the tcgen05 building blocks: D[128,64] = A[128,64] . W[64,64]^T (fp32 device pointers) through
 * the same bf16x3 split / SWIZZLE_128B tile writers / TMEM epilogue the message-passing kernels use. */
int pgmp_selftest_umma(const float* a, const float* w, float* d, pgmp_stream_t stream);
/* the same product with the A operand in tensor memory (tcgen05.st + the TS form of tcgen05.mma) */
int pgmp_selftest_umma_ts(const float* a, const float* w, float* d, pgmp_stream_t stream);

/* The FIRST int32 of the workspace is a status word written by pgmp_mpn_forward (stream-ordered, no host sync inside):
 * 0 = the input was well-formed; bit 0 = edge_index held a node id outside [0, num_nodes) -- the reference raises an
 * IndexError there; here such an edge is dropped (its logit is left unwritten) and the caller reads the word when it next
 * synchronises (the Python mirror raises at its next forward / check_status()). */
#define PGMP_MPN_STATUS_BAD_EDGE 1
uint64_t pgmp_mpn_workspace_bytes(const pgmp_mpn_params* p);
int pgmp_mpn_forward(const pgmp_mpn_params* p, pgmp_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training step of the message-passing network: forward of
 * NodeClassificationMPNSimple (NodeClassificationMPNSimple.py:62-97) in train() mode -- BatchNorm1d
 * with batch statistics (layers.py:13-14, 22-23), running statistics updated in place -- with the
 * type-agnostic MPLayer (layers.py:32-86; AGGR max / add / mean, SKIP, USE_NODE_UPDATE_MLP), and the
 * reverse pass torch autograd runs for the reference (train.py:232-236): gradients of every
 * parameter and of the node input x.  fp32-accurate throughout (products: tensor cores with the 3xTF32 operand split; statistics in fp64).
 * PGMP_TRAIN_TC=1 in the environment moves the E-level forward products to tcgen05 (bf16 hi / lo operand pairs, fp32
 * accumulation in tensor memory: ~1e-5 relative instead of ~1e-6; csrc/mpn_train_tc.cu).
 *
 * Parameters and their gradients are two flat fp32 device buffers with the same element offsets;
 * every matrix keeps the layout of its nn.Linear.weight, [out][in].
 * Two calls sharing one workspace (the forward leaves the activations the backward needs):
 *   pgmp_mpn_train_forward  -> logits, BatchNorm running statistics
 *   pgmp_mpn_train_backward -> grads (ACCUMULATED into `grads`: zero it first), grad_x (overwritten)
 * Reductions run in a fixed order (no floating-point atomics): results are reproducible run to run.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pgmp_mlp_train {
  int32_t n_layers;
  int32_t dims[PGMP_MAX_LAYERS + 1];       /* dims[0] = input width, dims[l+1] = output width of Linear l */
  int32_t relu[PGMP_MAX_LAYERS];           /* ReLU after Linear l */
  int32_t bn[PGMP_MAX_LAYERS];             /* BatchNorm1d after that ReLU (ReLU comes BEFORE BatchNorm, layers.py:11-14) */
  int64_t w[PGMP_MAX_LAYERS];              /* element offset of Linear.weight [dims[l+1]][dims[l]] in params / grads */
  int64_t b[PGMP_MAX_LAYERS];              /* Linear.bias */
  int64_t gamma[PGMP_MAX_LAYERS];          /* BatchNorm weight / bias (unused when bn[l] == 0) */
  int64_t beta[PGMP_MAX_LAYERS];
  float* running_mean[PGMP_MAX_LAYERS];    /* device [dims[l+1]], updated in place with momentum 0.1, or NULL */
  float* running_var[PGMP_MAX_LAYERS];     /* (unbiased batch variance goes into the running estimate) */
} pgmp_mlp_train;

typedef struct pgmp_mpn_train_params {
  int64_t num_nodes, num_edges;
  const float* x;                          /* device [N, node_emb.dims[0]] contiguous */
  const float* edge_attr;                  /* device [E, edge_emb.dims[0]] contiguous */
  const int64_t* edge_index;               /* device [2, E]; row 0 = source j, row 1 = target i */
  int32_t dim;                             /* 64 */
  int32_t skip, steps, aux_loss_steps;
  int32_t aggr;                            /* PGMP_AGGR_* */
  int32_t has_update_mlp;
  int32_t num_classes;
  const float* params;                     /* device, flat */
  float* grads;                            /* device, flat, same offsets (backward only) */
  pgmp_mlp_train node_emb, edge_emb, edge_head, node_head, class_head;
  int64_t w1, b1;                          /* mlp_edge.0 [64][2*nd + ed], nd = 64 * (skip ? 2 : 1), ed likewise; columns [x_i ; x_j ; e] (layers.py:66) */
  int64_t w2, b2;                          /* mlp_edge.2 [64][64] */
  int64_t wm, bm;                          /* mlp_node.0 [64][nd + 64]; columns [x_i ; e'] (layers.py:76) */
  int64_t wu, bu;                          /* update_mlp.0 [64][64] (has_update_mlp) */
  /* n_out = min(steps, aux_loss_steps + 1) reported steps */
  float* edge_logits;                      /* out [n_out][E] */
  float* node_logits;                      /* out [n_out][N] */
  float* class_logits;                     /* out [n_out][N][num_classes] */
  /* backward only */
  const float* d_edge_logits;              /* [n_out][E]   dL/d edge_logits */
  const float* d_node_logits;              /* [n_out][N] */
  const float* d_class_logits;             /* [n_out][N][num_classes] */
  float* grad_x;                           /* out [N, node_emb.dims[0]] or NULL */
  void* workspace;                         /* device, pgmp_mpn_train_workspace_bytes() bytes, 256-B aligned; must survive from forward to backward */
  uint64_t workspace_bytes;
  /* TypeAwareMPNLayer (layers.py:157-274), per_type != 0: wm / bm are mlp_node.mlp.0.0 [64][nd + 64] and matrix t sits
   * wm_type_stride elements further (17 matrices); wu is update_mlp.0 [64][num_types * 64]; attn = PGMP_ATTN_* with
   * attn_net.0 = wa [1 or 17][64], ba.  Both calls wait for the stream once (the sizes of the 17 source-type groups
   * are read back: the per-type products are launched per group). */
  int32_t per_type, num_types, attn, reserved_;
  const int64_t* node_types;               /* device [N], values in [0, num_types) (after NODE_TYPE_SUMMARY) */
  int64_t wm_type_stride;
  int64_t wa, ba;
} pgmp_mpn_train_params;

uint64_t pgmp_mpn_train_workspace_bytes(const pgmp_mpn_train_params* p);
int pgmp_mpn_train_forward(const pgmp_mpn_train_params* p, pgmp_stream_t stream);
int pgmp_mpn_train_backward(const pgmp_mpn_train_params* p, pgmp_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training-time label construction, host part (no device work): the detection-to-ground-truth matching of every image
 * of a batch -- src/graph_constructor/ConstructGraph.py:626-686 (EDGE_LABEL_METHOD 4), :769-942 (method 6),
 * USE_NEIGHBOURS :704-727 / :890-911 -- on the float32 similarity matrices exp(-d^2 / factor) (:773-785) the caller
 * computed.  Replaces the per-image torch / scipy.optimize.linear_sum_assignment calls of the reference's Python loop;
 * the assignment is SciPy's algorithm restated (csrc/match.cu).  All pointers are HOST memory.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pgmp_match_params {
  int32_t batch;
  int32_t method;                      /* EDGE_LABEL_METHOD: 4 or 6 */
  int32_t use_neighbours;              /* USE_NEIGHBOURS */
  int32_t num_threads;                 /* images are matched in parallel */
  float matching_radius, inclusion_radius;
  const float* sim;                    /* [batch][rows][cols]: similarity of annotated joint r to candidate c */
  int64_t sim_stride_b, sim_stride_g;  /* element strides of the image and the row */
  int32_t max_gt, max_det;             /* row strides of gt_type / det_type / ambiguous */
  const int32_t* num_gt;               /* [batch] annotated joints of the image (rows in use) */
  const int32_t* num_det;              /* [batch] candidates of the image (columns in use) */
  const int32_t* gt_type;              /* [batch][max_gt] joint type of every annotated joint */
  const int32_t* det_type;             /* [batch][max_det] joint type of every candidate */
  int32_t cap;                         /* capacity of match_row / match_col per image */
  int32_t* match_row;                  /* out [batch][cap]: annotated joint of every matched candidate (matches first, then neighbours) */
  int32_t* match_col;                  /* out [batch][cap]: the candidate */
  int32_t* num_match;                  /* out [batch] */
  uint8_t* ambiguous;                  /* out [batch][max_det] (use_neighbours): candidates claimed by more than one joint */
  /* optional per-node outputs over the whole batch (node = node_offsets[image] + candidate); NULL node_person skips them */
  const int64_t* node_offsets;         /* [batch + 1] */
  const int32_t* gt_person;            /* [batch][max_gt] person of every annotated joint */
  int64_t* node_person;                /* out [N]: person the node is matched to, else -1 */
  int64_t* node_class;                 /* out [N]: joint type of the matched annotation, else 0 */
  float* node_label;                   /* out [N]: 1 for matched nodes */
  uint8_t* node_ambiguous;             /* out [N] (use_neighbours) */
} pgmp_match_params;

int pgmp_match_labels(const pgmp_match_params* p);

/* The exponent -d^2 / factor of the similarity (:773-785) of every annotated joint (rows, in the order of the image's
 * gt[:, :, 2].nonzero()) to every candidate of its image, float32, operation for operation what the reference computes;
 * the caller applies exp() with the reference's own routine (torch.exp) and hands the result to pgmp_match_labels. */
typedef struct pgmp_label_args_params {
  int32_t batch, max_persons, num_joints, num_threads;
  float clamp_max;                     /* annotated positions are rounded and clamped to [0, clamp_max] */
  float min_arg;                       /* exponents below this are raised to it (-80: the caller thresholds the matrix anyway); -INF: none */
  const int64_t* det;                  /* host [N][3] (x, y, type) */
  const int64_t* node_offsets;         /* host [batch + 1] */
  const float* gt;                     /* host [batch][max_persons][num_joints][3] (x, y, visible) */
  const float* factors;                /* host [batch][max_persons][num_joints] */
  int32_t max_gt, max_det;             /* row counts of the padded outputs */
  float* arg;                          /* out [batch][max_gt][max_det] */
  int32_t* num_gt;                     /* out [batch] */
  int32_t* gt_type;                    /* out [batch][max_gt] */
  int32_t* gt_person;                  /* out [batch][max_gt] */
  int32_t* det_type;                   /* out [batch][max_det] */
} pgmp_label_args_params;

int pgmp_label_similarity_args(const pgmp_label_args_params* p);
/* scipy.optimize.linear_sum_assignment(cost [nr][nc], maximize): rows / cols hold min(nr, nc) entries */
int pgmp_linear_sum_assignment(const double* cost, int64_t nr, int64_t nc, int maximize, int64_t* rows, int64_t* cols);

/* ------------------------------------------------------------------------------------------------
 * Grouping tail -- replaces sigmoid/softmax (src/valid.py:109-111), the node threshold + subgraph
 * of pred_to_ann (src/Utils/Utils.py:1448-1451), pred_to_person with CC_METHOD GAEC (:499-514),
 * cluster_graph / extract_edge_matrix / cluster_andres_graph
 * (src/Utils/correlation_clustering/correlation_clustering_utils.py:21-64, 99-136, 187-256; the
 * GAEC solver itself is the reference's missing native andres_graph_wrapper) and the connected
 * component labelling + per-type winner selection of graph_cluster_to_persons (:672-743).
 * ---------------------------------------------------------------------------------------------- */
#define PGMP_CC_GAEC 0
#define PGMP_CC_THRESHOLD 1
#define PGMP_CC_GREEDY 2      /* greedy_person_construction, Utils.py:517-626: person_labels = core node of every node or -1 */

typedef struct pgmp_group_params {
  int32_t batch, num_joints;
  int64_t num_nodes, num_edges;
  float node_threshold;                /* MPN.NODE_THRESHOLD */
  int32_t cc_method;                   /* PGMP_CC_GAEC (greedy additive edge contraction), PGMP_CC_THRESHOLD (Utils.py:508-509) or PGMP_CC_GREEDY */
  float edge_threshold;                /* PGMP_CC_THRESHOLD: edges with probability > this join their ends (0.8 in the reference) */
  const int64_t* node_offsets;         /* device [B+1] prefix sums of nodes per image */
  const int64_t* edge_offsets;         /* device [B+1] prefix sums of edges per image (edges grouped by image) */
  const int64_t* edge_index;           /* device [2,E] global ids */
  const int64_t* joint_det;            /* device [N,3] */
  const float* node_logits;            /* device [N] */
  const float* edge_logits;            /* device [E] */
  const float* class_logits;           /* device [N, num_joints] or NULL */
  int64_t* person_labels;              /* out [N]: component id within the image, numbered by smallest node */
  int32_t* num_components;             /* out [B] */
  int32_t* num_kept_edges;             /* out [B]: edges whose two ends pass the node threshold (0 -> the reference returns None, Utils.py:1452,1457) */
  int32_t max_persons;                 /* capacity of `persons` per image */
  int32_t max_nodes_per_image;         /* upper bound of nodes in any image (sizes the dense cluster graph) */
  double* persons;                     /* out [B][max_persons][num_joints][3] (x, y, score), Utils.py:709-721 */
  int32_t* num_persons;                /* out [B] */
  int32_t* mutants;                    /* out [B] (Utils.py:703-706) */
  void* workspace;
  uint64_t workspace_bytes;
} pgmp_group_params;

uint64_t pgmp_group_workspace_bytes(const pgmp_group_params* p);
int pgmp_group_persons(const pgmp_group_params* p, pgmp_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pose-assembly tail of pred_to_ann (src/Utils/Utils.py:1472-1477): refine (:1026-1104) -- joints a person is missing
 * are looked up in the heatmaps where the tag is closest to the person's mean tag -- and adjust (:917-936), batched
 * over images and persons; `persons` is updated in place.
 * ---------------------------------------------------------------------------------------------- */
typedef struct pgmp_refine_params {
  int32_t batch, num_joints, height, width;
  int32_t tag_dim;                     /* T of tags [B, J, H, W, T] (1 for [B, J, H, W]) */
  int32_t max_persons;
  int32_t do_refine, do_adjust;        /* with_refine / adjustment of pred_to_ann */
  const float* scoremaps;              /* device [B, J, H, W] */
  const float* tags;                   /* device [B, J, H, W, T] */
  double* persons;                     /* device [B][max_persons][J][3] (x, y, score), as pgmp_group_persons writes them */
  const int32_t* num_persons;          /* device [B] */
  void* workspace;
  uint64_t workspace_bytes;
} pgmp_refine_params;

uint64_t pgmp_refine_workspace_bytes(const pgmp_refine_params* p);
int pgmp_refine_persons(const pgmp_refine_params* p, pgmp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PGMP_H_ */
