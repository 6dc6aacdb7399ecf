// Scoremap assembly in front of the NMS: hr_process_output of the reference
// (src/Models/HigherHRNet/hrnet.py:587-611).  The HigherHRNet head returns a half-resolution stage
// [B, 2J, h, w] (J heatmaps + J tag maps) and a full-resolution stage [B, J, H, W]:
//     up        = interpolate(stage1, size = (H, W), mode = 'bilinear', align_corners = False)
//     scoremaps = (stage2 + up[:, :J]) / 2        ("avg")   |   up[:, :J]   ("small")
//     tags      = up[:, J:]
// One kernel instead of three full-map passes (interpolate, add, divide): every output element is written once,
// stage2 is read once, stage1 (a quarter of the size) is read through L1 / L2.  HBM-bound:
// (2J H W + 2J h w + J H W) * 4 bytes per image in "avg" mode.
// The interpolation follows ATen's area_pixel_compute_source_index / linear weights (align_corners = False):
// src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out, i1 = i0 + (i0 < in - 1), weights (1 - l, l), and
// out = wy0 * (wx0 p00 + wx1 p01) + wy1 * (wx0 p10 + wx1 p11) with every product and sum rounded separately.
#include "common.cuh"

namespace pgmp {
namespace {

struct AssembleArgs {
  const float* s1;
  const float* s2;
  int B, C1, J, h, w, H, W, mode;
  float* score;
  float* tags;
  float scale_y, scale_x;
};

__device__ __forceinline__ void source_index(float scale, int dst, int in, int& i0, int& i1, float& l0, float& l1) {
  bilinear_source_index(scale, dst, in, i0, i1, l0, l1);
}

constexpr int kAsmRows = 4;   // consecutive output rows per thread: the column indices / weights are computed once

// grid: (ceil(W / 4 / 64), ceil(H / 16), B * C1); 256 threads = 4 row groups x 64 threads; 4 columns x 4 rows per thread
__global__ void __launch_bounds__(256) assemble_kernel(const AssembleArgs a) {
  const int x0 = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;
  const int yb = (blockIdx.y * 4 + (threadIdx.x >> 6)) * kAsmRows;
  if (x0 >= a.W || yb >= a.H) return;
  const int b = blockIdx.z / a.C1, c = blockIdx.z % a.C1;
  const bool heat = c < a.J;
  if (!heat && !a.tags) return;
  int i0[4], i1[4];
  float wx0[4], wx1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) source_index(a.scale_x, min(x0 + q, a.W - 1), a.w, i0[q], i1[q], wx0[q], wx1[q]);
  const float* __restrict__ plane = a.s1 + ((size_t)b * a.C1 + c) * a.h * a.w;
  const bool vec = x0 + 3 < a.W && (a.W & 3) == 0;
#pragma unroll
  for (int r = 0; r < kAsmRows; ++r) {
    const int y = yb + r;
    if (y >= a.H) break;
    int y0, y1;
    float wy0, wy1;
    source_index(a.scale_y, y, a.h, y0, y1, wy0, wy1);
    const float* __restrict__ r0 = plane + (size_t)y0 * a.w;
    const float* __restrict__ r1 = plane + (size_t)y1 * a.w;
    float out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      out[q] = bilinear_combine(wx0[q], wx1[q], wy0, wy1, __ldg(r0 + i0[q]), __ldg(r0 + i1[q]), __ldg(r1 + i0[q]), __ldg(r1 + i1[q]));
    }
    float* __restrict__ dst;
    if (heat) {
      const size_t o = (((size_t)b * a.J + c) * a.H + y) * a.W + x0;
      dst = a.score + o;
      if (a.mode == 0) {      // (s2 + up) / 2
        const float* __restrict__ p2 = a.s2 + o;
        if (vec && (reinterpret_cast<uintptr_t>(p2) & 15) == 0) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p2));
          out[0] = __fmul_rn(__fadd_rn(v.x, out[0]), 0.5f);
          out[1] = __fmul_rn(__fadd_rn(v.y, out[1]), 0.5f);
          out[2] = __fmul_rn(__fadd_rn(v.z, out[2]), 0.5f);
          out[3] = __fmul_rn(__fadd_rn(v.w, out[3]), 0.5f);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (x0 + q < a.W) out[q] = __fmul_rn(__fadd_rn(__ldg(p2 + q), out[q]), 0.5f);
        }
      }
    } else {
      dst = a.tags + (((size_t)b * (a.C1 - a.J) + (c - a.J)) * a.H + y) * a.W + x0;
    }
    if (vec && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (x0 + q < a.W) dst[q] = out[q];
    }
  }
}

// joint_tags[n] = up(stage1)[b, J + type, y, x] at the detections only (the tag half of hr_process_output evaluated
// where ConstructGraph.py:103 reads it): the same arithmetic as assemble_kernel, 4 source values per node.
__global__ void __launch_bounds__(256) stage_tags_kernel(const float* __restrict__ s1, int C1, int J, int h, int w, float scale_y,
                                                         float scale_x, const int64_t* __restrict__ joint_det,
                                                         const int64_t* __restrict__ batch_index, int64_t N,
                                                         float* __restrict__ joint_tags) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int x = (int)joint_det[n * 3], y = (int)joint_det[n * 3 + 1], t = (int)joint_det[n * 3 + 2];
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  source_index(scale_y, y, h, y0, y1, wy0, wy1);
  source_index(scale_x, x, w, x0, x1, wx0, wx1);
  const float* __restrict__ plane = s1 + ((size_t)batch_index[n] * C1 + J + t) * h * w;
  const float* __restrict__ r0 = plane + (size_t)y0 * w;
  const float* __restrict__ r1 = plane + (size_t)y1 * w;
  joint_tags[n] = bilinear_combine(wx0, wx1, wy0, wy1, __ldg(r0 + x0), __ldg(r0 + x1), __ldg(r1 + x0), __ldg(r1 + x1));
}

}  // namespace
}  // namespace pgmp

extern "C" int pgmp_gc_gather_stage_tags(const float* stage1, int32_t channels1, int32_t num_joints, int32_t h, int32_t w,
                                         int32_t H, int32_t W, const int64_t* joint_det, const int64_t* batch_index,
                                         int64_t num_nodes, float* joint_tags, pgmp_stream_t stream) {
  using namespace pgmp;
  if (num_nodes == 0) return PGMP_OK;
  if (!stage1 || !joint_det || !batch_index || !joint_tags || num_nodes < 0) return set_error(PGMP_ERR_INVALID, "null pointer");
  if (channels1 < 2 * num_joints || h <= 0 || w <= 0 || H <= 0 || W <= 0)
    return set_error(PGMP_ERR_INVALID, "stage 1 must hold num_joints heatmaps and num_joints tag maps");
  PGMP_LAUNCH(stage_tags_kernel, (unsigned)ceil_div<int64_t>(num_nodes, 256), 256, 0, static_cast<cudaStream_t>(stream), stage1,
              channels1, num_joints, h, w, (float)h / (float)H, (float)w / (float)W, joint_det, batch_index, num_nodes, joint_tags);
  return PGMP_OK;
}

extern "C" int pgmp_gc_assemble_scoremaps(const float* stage1, const float* stage2, int32_t batch, int32_t channels1,
                                          int32_t num_joints, int32_t h, int32_t w, int32_t H, int32_t W, int32_t mode,
                                          float* scoremaps, float* tags, pgmp_stream_t stream) {
  using namespace pgmp;
  if (batch <= 0 || num_joints <= 0 || channels1 < num_joints || h <= 0 || w <= 0 || H <= 0 || W <= 0)
    return set_error(PGMP_ERR_INVALID, "bad sizes");
  if (mode != PGMP_ASSEMBLE_AVG && mode != PGMP_ASSEMBLE_SMALL) return set_error(PGMP_ERR_INVALID, "mode %d", mode);
  if (!stage1 || !scoremaps || (mode == PGMP_ASSEMBLE_AVG && !stage2)) return set_error(PGMP_ERR_INVALID, "null pointer");
  if ((int64_t)batch * channels1 > 65535 || H > 16 * 65535) return set_error(PGMP_ERR_INVALID, "batch * channels or height above 65535");
  AssembleArgs a{stage1, stage2, batch, channels1, num_joints, h, w, H, W, mode, scoremaps, tags,
                 (float)h / (float)H, (float)w / (float)W};
  const dim3 grid((unsigned)ceil_div(ceil_div(W, 4), 64), (unsigned)ceil_div(H, 4 * kAsmRows), (unsigned)(batch * channels1));
  PGMP_LAUNCH(assemble_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), a);
  return PGMP_OK;
}
