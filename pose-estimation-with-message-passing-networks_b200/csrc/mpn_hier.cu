// UPDATE_TYPE = "hierarch_mlp" (HierarchUpdateMlp, layers.py:89-128): the per-type aggregates U[node, t, :] go through a
// body-part tree -- 7 first-layer Linears over groups of joint types (-> 32), 6 second-layer Linears over pairs of those
// (-> 32), one final Linear (192 -> 64), ReLU after each.  fp32 SIMT, one thread per node; the first-layer outputs of
// a node tile sit in shared memory, the second layer and the final layer are accumulated on the fly.  Used by both
// precision modes (a handful of the reference's experiments select it; the default "mlp" update has its own kernels).
#include "mpn_common.cuh"
#include "simt_mlp.cuh"

namespace pgmp {
namespace {

constexpr int kH = kD / 2;   // 32

struct HierPlan {
  int n_first[7];            // joint types per first-layer group
  int first_types[7][5];
  int second[6][2];          // first-layer groups each second-layer Linear reads (layers.py:116,119)
};

__constant__ HierPlan c_plan17 = {{5, 2, 2, 2, 2, 2, 2},
                                  {{0, 1, 2, 3, 4}, {5, 6}, {7, 9}, {8, 10}, {11, 12}, {13, 15}, {14, 16}},
                                  {{0, 1}, {1, 2}, {1, 3}, {1, 4}, {4, 5}, {4, 6}}};
__constant__ HierPlan c_plan14 = {{2, 2, 2, 2, 2, 2, 2},
                                  {{0, 1}, {2, 3}, {4, 6}, {5, 7}, {8, 9}, {10, 12}, {11, 13}},
                                  {{0, 1}, {1, 2}, {1, 3}, {1, 4}, {4, 5}, {4, 6}}};

__global__ void __launch_bounds__(kTile) node_update_hier_kernel(AggrView av, int64_t N, int T, const float* __restrict__ hw,
                                                                 float* __restrict__ h) {
  extern __shared__ __align__(16) float s_out1[];       // [7][32][128]
  const HierPlan& plan = T == 17 ? c_plan17 : c_plan14;
  const int tid = threadIdx.x;
  const int64_t row = (int64_t)blockIdx.x * kTile + tid;
  const int64_t srow = row < N ? row : N - 1;
  const float* __restrict__ w = hw;
  float u[kD];
  // ---- first layer (layers.py:123-124)
  for (int gi = 0; gi < 7; ++gi) {
    const int nt = plan.n_first[gi], in = nt * kD;
    const float* __restrict__ bias = w + (size_t)kH * in;
    float acc[kH];
#pragma unroll
    for (int o = 0; o < kH; ++o) acc[o] = __ldg(bias + o);
    for (int s = 0; s < nt; ++s) {
      merge_parts(av, plan.first_types[gi][s], srow, N, u);
#pragma unroll 4
      for (int o = 0; o < kH; ++o) {
        const float4* __restrict__ wr = reinterpret_cast<const float4*>(w + (size_t)o * in + s * kD);
        float a = acc[o];
#pragma unroll
        for (int q = 0; q < kD / 4; ++q) {
          const float4 wv = __ldg(wr + q);
          a = fmaf(wv.x, u[4 * q], a); a = fmaf(wv.y, u[4 * q + 1], a); a = fmaf(wv.z, u[4 * q + 2], a); a = fmaf(wv.w, u[4 * q + 3], a);
        }
        acc[o] = a;
      }
    }
#pragma unroll
    for (int o = 0; o < kH; ++o) s_out1[(gi * kH + o) * kTile + tid] = fmaxf(acc[o], 0.f);
    w = bias + kH;
  }
  // ---- second layer (layers.py:125-126) folded into the final Linear (:128)
  const float* __restrict__ w2 = w;                                  // 6 x ([32][64] + [32])
  const float* __restrict__ wf = w2 + (size_t)6 * (kH * kD + kH);    // [64][192] + [64]
  float facc[kD];
#pragma unroll
  for (int j = 0; j < kD; ++j) facc[j] = __ldg(wf + (size_t)kD * 6 * kH + j);
  for (int i = 0; i < 6; ++i) {
    const float* __restrict__ wi = w2 + (size_t)i * (kH * kD + kH);
    float a2[kH];
#pragma unroll
    for (int o = 0; o < kH; ++o) a2[o] = __ldg(wi + kH * kD + o);
    for (int half = 0; half < 2; ++half) {
      const int gsrc = plan.second[i][half];
      for (int k = 0; k < kH; ++k) {
        const float v = s_out1[(gsrc * kH + k) * kTile + tid];
#pragma unroll
        for (int o = 0; o < kH; ++o) a2[o] = fmaf(__ldg(wi + (size_t)o * kD + half * kH + k), v, a2[o]);
      }
    }
#pragma unroll
    for (int o = 0; o < kH; ++o) a2[o] = fmaxf(a2[o], 0.f);
    for (int j = 0; j < kD; ++j) {
      const float* __restrict__ wr = wf + (size_t)j * 6 * kH + i * kH;
      float a = facc[j];
#pragma unroll
      for (int o = 0; o < kH; ++o) a = fmaf(__ldg(wr + o), a2[o], a);
      facc[j] = a;
    }
  }
  if (row < N) {
    float4* __restrict__ o4 = reinterpret_cast<float4*>(h + row * kD);
#pragma unroll
    for (int q = 0; q < kD / 4; ++q)
      o4[q] = make_float4(fmaxf(facc[4 * q], 0.f), fmaxf(facc[4 * q + 1], 0.f), fmaxf(facc[4 * q + 2], 0.f), fmaxf(facc[4 * q + 3], 0.f));
  }
}

}  // namespace

int mpn_run_mlp_rows(const pgmp_mlp& mlp, const float* in, int64_t M, float* out, cudaStream_t st);

// h <- HierarchUpdateMlp(U); node / class logits of a reported step through the generic SIMT chain
int mpn_node_update_hier(const pgmp_mpn_params& p, const MpnWorkspace& w, int out_slot, cudaStream_t st) {
  const size_t smem = sizeof(float) * 7 * kH * kTile;
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(node_update_hier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  AggrView av{w.bin_count, w.bin_lstart, w.bin_lpart, w.group_pstart, w.part_val, w.part_mx, w.part_se, p.aggr, p.attn};
  const int64_t N = p.num_nodes;
  PGMP_LAUNCH(node_update_hier_kernel, (unsigned)ceil_div<int64_t>(N, kTile), kTile, smem, st, av, N, p.num_types, p.hier, w.h);
  if (out_slot < 0) return PGMP_OK;
  int rc = mpn_run_mlp_rows(p.node_head, w.h, N, p.node_logits + (size_t)out_slot * N, st);
  if (rc != PGMP_OK) return rc;
  return mpn_run_mlp_rows(p.class_head, w.h, N, p.class_logits + (size_t)out_slot * N * p.num_classes, st);
}

}  // namespace pgmp
