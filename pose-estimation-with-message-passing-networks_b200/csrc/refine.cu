// Pose-assembly tail after the grouping: refine (src/Utils/Utils.py:1026-1104) and adjust
// (:917-936) of pred_to_ann (:1472-1477), batched over images and persons on the device.
//
// refine: for every person the mean tag of its detected joints; for every joint type the pixel maximising
// score - round(|tag - mean tag|) over the whole map (first maximum in row-major order); a joint the person is missing
// is placed there (+0.5 and a quarter pixel towards the higher neighbour, score 0.001) if the heatmap is positive.
// The reference scans the J full maps once per person on the host (P x J x H x W); here one pass over the maps serves
// 8 persons at a time.  The float32 details that decide the arg-max are numpy's: the mean is np.mean in float32
// (pairwise-8 over a contiguous axis when the tag dimension is 1, row by row otherwise), squares and sums are rounded
// separately (no FMA), sqrt is correctly rounded, np.round is round-half-even, an empty person gives a NaN mean and
// np.argmax then returns pixel 0.  Bit-exact with oracle/refine.py, which is pinned to the reference's functions.
#include <cfloat>
#include <climits>

#include "common.cuh"

namespace pgmp {
namespace {

constexpr int kRefMaxJ = 32, kRefMaxT = 8, kRefGroup = 8, kRefChunks = 16;

struct RefineArgs {
  const float* sm;
  const float* tag;
  int B, J, H, W, T, Pmax;
  double* kp;            // [B][Pmax][J][3] (x, y, score)
  const int32_t* np;     // [B]
  float* prev;           // [B][Pmax][T] mean tags
  int32_t* ndet;         // [B][Pmax] detected joints
  float* part_val;       // [B][J][Pmax][kRefChunks]
  int32_t* part_idx;
};

// np.mean of float32 values a[0..n) along a contiguous axis: pairwise summation with 8 accumulators for n >= 8
__device__ float numpy_mean_contiguous(const float* a, int stride, int n) {
  float res;
  int i;
  if (n < 8) {
    res = 0.f;
    for (i = 0; i < n; ++i) res = __fadd_rn(res, a[i * stride]);
  } else {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j * stride];
    for (i = 8; i + 8 <= n; i += 8)
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[(i + j) * stride]);
    res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a[i * stride]);
  }
  return __fdiv_rn(res, (float)n);      // n == 0: 0 / 0 = NaN, as np.mean of an empty array
}

// one warp per (image, person): tags at the detected joints -> mean tag (Utils.py:1040-1048, 1062)
__global__ void __launch_bounds__(32) refine_mean_kernel(const RefineArgs a) {
  __shared__ float s[kRefMaxJ * kRefMaxT];
  const int b = blockIdx.x / a.Pmax, p = blockIdx.x % a.Pmax, j = threadIdx.x;
  if (p >= a.np[b]) return;
  const double* k = a.kp + (((size_t)b * a.Pmax + p) * a.J) * 3;
  const bool det = j < a.J && k[j * 3 + 2] > 0.0;
  const uint32_t mask = __ballot_sync(0xffffffffu, det);
  const int n = __popc(mask), rank = __popc(mask & ((1u << j) - 1u));
  if (det) {
    int x = (int)k[j * 3], y = (int)k[j * 3 + 1];                 // astype(np.int32)
    x = min(max(x, 0), a.W - 1);                                   // memory safety; valid inputs are inside the map
    y = min(max(y, 0), a.H - 1);
    const float* t = a.tag + ((((size_t)b * a.J + j) * a.H + y) * a.W + x) * a.T;
    for (int d = 0; d < a.T; ++d) s[rank * a.T + d] = t[d];
  }
  __syncwarp();
  if (j < a.T) {
    float m;
    if (a.T == 1) {
      m = numpy_mean_contiguous(s, 1, n);
    } else {                                                       // reduction over the outer axis: row by row
      float acc = 0.f;
      for (int i = 0; i < n; ++i) acc = __fadd_rn(acc, s[i * a.T + j]);
      m = __fdiv_rn(acc, (float)n);
    }
    a.prev[((size_t)b * a.Pmax + p) * a.T + j] = m;
  }
  if (j == 0) a.ndet[b * a.Pmax + p] = n;
}

__device__ __forceinline__ void better(float& v, int& i, float v2, int i2) {
  if (v2 > v || (v2 == v && i2 < i)) { v = v2; i = i2; }
}

// grid (chunk, joint, image): arg-max of score - round(dist to the person's mean tag) over a range of pixels, for all
// persons of the image, kRefGroup persons per pass (Utils.py:1066-1074)
__global__ void __launch_bounds__(256) refine_argmax_kernel(const RefineArgs a) {
  __shared__ float s_prev[kRefGroup * kRefMaxT];
  __shared__ float s_v[8][kRefGroup];
  __shared__ int s_i[8][kRefGroup];
  const int c = blockIdx.x, jt = blockIdx.y, b = blockIdx.z, t = threadIdx.x;
  const int P = a.np[b];
  const int HW = a.H * a.W;
  const int len = ceil_div(HW, kRefChunks), begin = c * len, end = min(begin + len, HW);
  const float* __restrict__ sm = a.sm + ((size_t)b * a.J + jt) * HW;
  const float* __restrict__ tg = a.tag + ((size_t)b * a.J + jt) * HW * a.T;
  for (int pg = 0; pg < P; pg += kRefGroup) {
    __syncthreads();
    if (t < kRefGroup * a.T) {
      const int q = t / a.T, d = t % a.T;
      s_prev[q * kRefMaxT + d] = pg + q < P ? a.prev[((size_t)b * a.Pmax + pg + q) * a.T + d] : 0.f;
    }
    __syncthreads();
    float bv[kRefGroup];
    int bi[kRefGroup];
#pragma unroll
    for (int q = 0; q < kRefGroup; ++q) { bv[q] = -FLT_MAX; bi[q] = INT_MAX; }
    for (int pix = begin + t; pix < end; pix += 256) {
      const float sc = __ldg(sm + pix);
      float tv[kRefMaxT];
      for (int d = 0; d < a.T; ++d) tv[d] = __ldg(tg + (size_t)pix * a.T + d);
#pragma unroll
      for (int q = 0; q < kRefGroup; ++q) {
        float acc = 0.f;
        for (int d = 0; d < a.T; ++d) {
          const float df = __fadd_rn(tv[d], -s_prev[q * kRefMaxT + d]);
          const float sq = __fmul_rn(df, df);
          acc = d == 0 ? sq : __fadd_rn(acc, sq);
        }
        const float v = __fadd_rn(sc, -rintf(__fsqrt_rn(acc)));
        if (v > bv[q]) { bv[q] = v; bi[q] = pix; }                 // strictly greater: the first maximum stays
      }
    }
#pragma unroll
    for (int q = 0; q < kRefGroup; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(0xffffffffu, bv[q], o);
        const int i2 = __shfl_xor_sync(0xffffffffu, bi[q], o);
        better(bv[q], bi[q], v2, i2);
      }
      if ((t & 31) == 0) { s_v[t >> 5][q] = bv[q]; s_i[t >> 5][q] = bi[q]; }
    }
    __syncthreads();
    if (t < kRefGroup && pg + t < P) {
      float v = s_v[0][t];
      int i = s_i[0][t];
      for (int w = 1; w < 8; ++w) better(v, i, s_v[w][t], s_i[w][t]);
      const size_t o = ((((size_t)b * a.J + jt) * a.Pmax) + pg + t) * kRefChunks + c;
      a.part_val[o] = v;
      a.part_idx[o] = i;
    }
  }
}

// one warp per (image, person), lane = joint: place the missing joints (Utils.py:1075-1102)
__global__ void __launch_bounds__(32) refine_apply_kernel(const RefineArgs a) {
  const int b = blockIdx.x / a.Pmax, p = blockIdx.x % a.Pmax, j = threadIdx.x;
  if (p >= a.np[b] || j >= a.J) return;
  // pred_to_ann refines only when the first person has a detected joint (Utils.py:1472)
  // (evaluated on the state BEFORE any update: the first person's detected-joint count from refine_mean_kernel)
  if (a.ndet[b * a.Pmax] == 0) return;
  double* k = a.kp + (((size_t)b * a.Pmax + p) * a.J + j) * 3;
  const size_t o = ((((size_t)b * a.J + j) * a.Pmax) + p) * kRefChunks;
  float v = a.part_val[o];
  int idx = a.part_idx[o];
  for (int c = 1; c < kRefChunks; ++c) better(v, idx, a.part_val[o + c], a.part_idx[o + c]);
  if (a.ndet[b * a.Pmax + p] == 0 || idx == INT_MAX) idx = 0;      // NaN mean tag: np.argmax returns the first NaN
  const int y = idx / a.W, x = idx % a.W;
  const float* __restrict__ m = a.sm + ((size_t)b * a.J + j) * a.H * a.W;
  const float val = m[y * a.W + x];
  double fx = x + 0.5, fy = y + 0.5;
  fx += m[y * a.W + min(x + 1, a.W - 1)] > m[y * a.W + max(x - 1, 0)] ? 0.25 : -0.25;
  fy += m[min(y + 1, a.H - 1) * a.W + x] > m[max(y - 1, 0) * a.W + x] ? 0.25 : -0.25;
  if (val > 0.f && k[2] == 0.0) {
    k[0] = fx;
    k[1] = fy;
    k[2] = 0.001;
  }
}

// adjust (Utils.py:917-936): a quarter pixel towards the higher neighbour, then the half-pixel centre offset
__global__ void __launch_bounds__(256) adjust_kernel(const RefineArgs a) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)a.B * a.Pmax * a.J;
  if (g >= total) return;
  const int j = (int)(g % a.J), p = (int)((g / a.J) % a.Pmax), b = (int)(g / ((int64_t)a.J * a.Pmax));
  if (p >= a.np[b]) return;
  double* k = a.kp + g * 3;
  if (!(k[2] > 0.0)) return;
  double cx = k[0], cy = k[1];
  int col = (int)cx, row = (int)cy;
  col = min(max(col, 0), a.W - 1);
  row = min(max(row, 0), a.H - 1);
  const float* __restrict__ m = a.sm + ((size_t)b * a.J + j) * a.H * a.W;
  cx += m[row * a.W + min(col + 1, a.W - 1)] > m[row * a.W + max(col - 1, 0)] ? 0.25 : -0.25;
  cy += m[min(row + 1, a.H - 1) * a.W + col] > m[max(0, row - 1) * a.W + col] ? 0.25 : -0.25;
  k[0] = cx + 0.5;
  k[1] = cy + 0.5;
}

struct RefineWs {
  float* prev;
  int32_t* ndet;
  float* part_val;
  int32_t* part_idx;
  uint64_t bytes;
};

RefineWs carve_refine(const pgmp_refine_params& p) {
  Carver c(p.workspace);
  RefineWs w;
  const uint64_t BP = (uint64_t)p.batch * p.max_persons;
  w.prev = c.take<float>(BP * p.tag_dim);
  w.ndet = c.take<int32_t>(BP);
  w.part_val = c.take<float>(BP * p.num_joints * kRefChunks);
  w.part_idx = c.take<int32_t>(BP * p.num_joints * kRefChunks);
  w.bytes = c.bytes();
  return w;
}

}  // namespace
}  // namespace pgmp

using namespace pgmp;

extern "C" uint64_t pgmp_refine_workspace_bytes(const pgmp_refine_params* p) {
  if (!p || p->batch <= 0 || p->max_persons <= 0 || p->num_joints <= 0 || p->tag_dim <= 0) return 0;
  pgmp_refine_params q = *p;
  q.workspace = nullptr;
  return carve_refine(q).bytes;
}

extern "C" int pgmp_refine_persons(const pgmp_refine_params* p, pgmp_stream_t stream) {
  if (!p) return set_error(PGMP_ERR_INVALID, "null params");
  if (p->batch <= 0 || p->max_persons <= 0 || p->height <= 0 || p->width <= 0) return set_error(PGMP_ERR_INVALID, "bad sizes");
  if (p->num_joints <= 0 || p->num_joints > kRefMaxJ) return set_error(PGMP_ERR_INVALID, "num_joints %d outside [1, %d]", p->num_joints, kRefMaxJ);
  if (p->tag_dim <= 0 || p->tag_dim > kRefMaxT) return set_error(PGMP_ERR_INVALID, "tag_dim %d outside [1, %d]", p->tag_dim, kRefMaxT);
  if (p->batch > 65535) return set_error(PGMP_ERR_INVALID, "batch above 65535");
  if (!p->scoremaps || !p->persons || !p->num_persons || (p->do_refine && (!p->tags || !p->workspace)))
    return set_error(PGMP_ERR_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RefineArgs a{p->scoremaps, p->tags, p->batch, p->num_joints, p->height, p->width, p->tag_dim, p->max_persons, p->persons,
               p->num_persons, nullptr, nullptr, nullptr, nullptr};
  if (p->do_refine) {
    const RefineWs w = carve_refine(*p);
    if (w.bytes > p->workspace_bytes) return set_error(PGMP_ERR_INVALID, "workspace too small");
    a.prev = w.prev; a.ndet = w.ndet; a.part_val = w.part_val; a.part_idx = w.part_idx;
    const unsigned bp = (unsigned)(p->batch * p->max_persons);
    PGMP_LAUNCH(refine_mean_kernel, bp, 32, 0, st, a);
    PGMP_LAUNCH(refine_argmax_kernel, dim3(kRefChunks, (unsigned)p->num_joints, (unsigned)p->batch), 256, 0, st, a);
    PGMP_LAUNCH(refine_apply_kernel, bp, 32, 0, st, a);
  }
  if (p->do_adjust)
    PGMP_LAUNCH(adjust_kernel, (unsigned)ceil_div<int64_t>((int64_t)p->batch * p->max_persons * p->num_joints, 256), 256, 0, st, a);
  return PGMP_OK;
}
