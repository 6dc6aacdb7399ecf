// Node features at the candidate pixels without materialising the feature maps:
//   x[n] = interpolate(feature_gather(feat), size = (H, W), bilinear, align_corners = False)[:, y_n, x_n]
// (PoseEstimation.py:64-66, 79, 341, 442-450; ConstructGraph.py:265, 269).  Convolution and interpolation are
// linear, so per candidate the 3 x 3 x Cin input patch is interpolated first (4 taps; the zero padding of the
// convolution is applied per tap) and multiplied once with the [9 Cin, Cout] weight matrix.  The reference writes
// and re-reads B x 128 x H x W floats (134 MB per 512-pixel image) for what is N x 128 outputs.
//
// One persistent CTA of 256 threads = (output channel c = tid & 127, half = tid >> 7) works on 8 candidates at a
// time: the 4 x 4 x Cin raw neighbourhood of each candidate is fetched once into shared memory (one 32-byte sector
// per element from NCHW maps -- device memory or pinned host memory), the interpolated patches P[k][8] are built
// from it, then every thread accumulates 4 candidates: per k one conflict-free weight read, one broadcast 16-byte
// patch read and 4 FMAs.  The transposed weights stay resident in shared memory (9 Cin Cout floats, 147 KB for
// 32 -> 128 channels).
#include "common.cuh"

namespace pgmp {
namespace {

constexpr int kGcThreads = 256;
constexpr int kGcCand = 8;              // candidates per pass
constexpr int kGcMaxCout = 128;         // one thread per output channel and half

struct GatherConvArgs {
  const float* feat; int64_t sb, sc, sy, sx;
  int cin, h, w, cout, out_h, out_w;
  const float* wmat; const float* bias;   // [9 cin][cout], [cout]
  const int64_t* joint_det; const int64_t* batch_index; int64_t n;
  float* x;
  int w_in_smem;
};

// ATen area_pixel_compute_source_index for align_corners = False, one axis
__device__ __forceinline__ void bilinear_tap(int dst, int in_size, int out_size, int& i0, int& i1, float& l0, float& l1) {
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  float src = __fsub_rn(__fmul_rn(__fadd_rn((float)dst, 0.5f), scale), 0.5f);
  src = fmaxf(src, 0.f);
  i0 = min((int)floorf(src), in_size - 1);
  i1 = min(i0 + 1, in_size - 1);
  l1 = __fsub_rn(src, (float)i0);
  l0 = __fsub_rn(1.f, l1);
}

// CIN > 0: compile-time channel count (power of two, 16 raw elements per thread prefetched one pass ahead in
// registers); CIN == 0: any channel count, no prefetch
template <int CIN>
__global__ void __launch_bounds__(kGcThreads, 1) gather_conv_kernel(const GatherConvArgs a) {
  extern __shared__ __align__(16) float s_mem[];
  const int cin = CIN > 0 ? CIN : a.cin;
  const int K = 9 * cin;
  float* s_w = s_mem;                                              // [K][cout] when resident
  float* s_p = s_mem + (a.w_in_smem ? (size_t)K * a.cout : 0);     // [K][8] interpolated patches
  float* s_r = s_p + (size_t)K * kGcCand;                          // [8][16][cin] raw 4 x 4 neighbourhoods
  struct Taps { int y0[kGcCand], x0[kGcCand], dy[kGcCand], dx[kGcCand], b[kGcCand]; float ly[kGcCand][2], lx[kGcCand][2]; };
  __shared__ Taps s_t[2];
  const int tid = threadIdx.x;
  if (a.w_in_smem)
    for (int i = tid; i < K * a.cout; i += kGcThreads) s_w[i] = __ldg(a.wmat + i);
  const float* __restrict__ wsrc = a.w_in_smem ? s_w : a.wmat;
  const int c = tid & (kGcMaxCout - 1), half = tid >> 7;
  const float bias = c < a.cout ? __ldg(a.bias + c) : 0.f;
  const int64_t groups = (a.n + kGcCand - 1) / kGcCand;
  constexpr int kRaw = CIN > 0 ? kGcCand * 16 * CIN / kGcThreads : 1;   // raw elements per thread and pass

  auto compute_taps = [&](int64_t grp, Taps& t) {
    if (tid < kGcCand) {
      const int64_t n = grp * kGcCand + tid;
      int y0 = 0, y1 = 0, x0 = 0, x1 = 0, b = -1;
      float ly0 = 0.f, ly1 = 0.f, lx0 = 0.f, lx1 = 0.f;
      if (grp < groups && n < a.n) {
        b = (int)a.batch_index[n];
        bilinear_tap((int)a.joint_det[n * 3 + 1], a.h, a.out_h, y0, y1, ly0, ly1);
        bilinear_tap((int)a.joint_det[n * 3 + 0], a.w, a.out_w, x0, x1, lx0, lx1);
      }
      t.b[tid] = b; t.y0[tid] = y0; t.x0[tid] = x0; t.dy[tid] = y1 - y0; t.dx[tid] = x1 - x0;
      t.ly[tid][0] = ly0; t.ly[tid][1] = ly1; t.lx[tid][0] = lx0; t.lx[tid][1] = lx1;
    }
  };
  // raw neighbourhood element i of a pass: rows y0 - 1 .. y0 + 2, columns x0 - 1 .. x0 + 2, zero outside the map
  auto raw = [&](const Taps& t, int i) {
    const int ci = CIN > 0 ? (i & (CIN - 1)) : i % cin;
    const int q = CIN > 0 ? i / CIN : i / cin;
    const int pos = q & 15, g = q >> 4;
    const int yy = t.y0[g] - 1 + (pos >> 2), xx = t.x0[g] - 1 + (pos & 3);
    float v = 0.f;
    if (t.b[g] >= 0 && (unsigned)yy < (unsigned)a.h && (unsigned)xx < (unsigned)a.w)
      v = __ldg(a.feat + t.b[g] * a.sb + ci * a.sc + yy * a.sy + xx * a.sx);
    return v;
  };

  float pre[kRaw];
  compute_taps(blockIdx.x, s_t[0]);
  __syncthreads();
  if (CIN > 0) {
#pragma unroll
    for (int j = 0; j < kRaw; ++j) pre[j] = raw(s_t[0], tid + j * kGcThreads);
  }
  int cur = 0;
  for (int64_t grp = blockIdx.x; grp < groups; grp += gridDim.x, cur ^= 1) {
    const int64_t n0 = grp * kGcCand;
    const Taps& t = s_t[cur];
    if (CIN > 0) {
#pragma unroll
      for (int j = 0; j < kRaw; ++j) s_r[tid + j * kGcThreads] = pre[j];
    } else {
      for (int i = tid; i < kGcCand * 16 * cin; i += kGcThreads) s_r[i] = raw(t, i);
    }
    compute_taps(grp + gridDim.x, s_t[cur ^ 1]);
    __syncthreads();
    // interpolated patches, taps in the order (y0,x0), (y0,x1), (y1,x0), (y1,x1)
    for (int i = tid; i < K * kGcCand; i += kGcThreads) {
      const int g = i & (kGcCand - 1), k = i >> 3;
      const int ci = CIN > 0 ? (k & (CIN - 1)) : k % cin;
      const int kk = CIN > 0 ? k / CIN : k / cin, ky = kk / 3, kx = kk - 3 * ky;
      const float* __restrict__ r = s_r + (size_t)g * 16 * cin + ci;
      float p = 0.f;
#pragma unroll
      for (int ty = 0; ty < 2; ++ty)
#pragma unroll
        for (int tx = 0; tx < 2; ++tx) {
          const int py = (ty ? t.dy[g] : 0) + ky, px = (tx ? t.dx[g] : 0) + kx;
          p = __fadd_rn(p, __fmul_rn(__fmul_rn(t.ly[g][ty], t.lx[g][tx]), r[(py * 4 + px) * cin]));
        }
      s_p[i] = p;
    }
    __syncthreads();
    if (CIN > 0 && grp + gridDim.x < groups) {       // the next pass's raw elements fly while this pass multiplies
#pragma unroll
      for (int j = 0; j < kRaw; ++j) pre[j] = raw(s_t[cur ^ 1], tid + j * kGcThreads);
    }
    if (c < a.cout) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const float4* __restrict__ p4 = reinterpret_cast<const float4*>(s_p) + half;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        const float wv = wsrc[(size_t)k * a.cout + c];
        const float4 pv = p4[2 * k];
        acc[0] = fmaf(pv.x, wv, acc[0]); acc[1] = fmaf(pv.y, wv, acc[1]);
        acc[2] = fmaf(pv.z, wv, acc[2]); acc[3] = fmaf(pv.w, wv, acc[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t n = n0 + half * 4 + j;
        if (n < a.n) a.x[n * a.cout + c] = acc[j] + bias;
      }
    }
    __syncthreads();                                 // s_p / s_r are rewritten by the next pass
  }
}

}  // namespace
}  // namespace pgmp

extern "C" int pgmp_gc_gather_conv(const pgmp_gather_conv_params* p, pgmp_stream_t stream) {
  using namespace pgmp;
  if (!p) return set_error(PGMP_ERR_INVALID, "null params");
  if (p->num_nodes < 0 || p->cin <= 0 || p->cout <= 0 || p->cout > kGcMaxCout || p->height <= 0 || p->width <= 0 ||
      p->out_height <= 0 || p->out_width <= 0)
    return set_error(PGMP_ERR_INVALID, "bad sizes (cout <= %d)", kGcMaxCout);
  if (p->num_nodes == 0) return PGMP_OK;
  if (!p->features || !p->weight_t || !p->bias || !p->joint_det || !p->batch_index || !p->x)
    return set_error(PGMP_ERR_INVALID, "null pointer");
  GatherConvArgs a;
  a.feat = p->features; a.sb = p->feat_stride_b; a.sc = p->feat_stride_c; a.sy = p->feat_stride_y; a.sx = p->feat_stride_x;
  a.cin = p->cin; a.h = p->height; a.w = p->width; a.cout = p->cout; a.out_h = p->out_height; a.out_w = p->out_width;
  a.wmat = p->weight_t; a.bias = p->bias; a.joint_det = p->joint_det; a.batch_index = p->batch_index; a.n = p->num_nodes;
  a.x = p->x;
  const size_t K = (size_t)9 * p->cin;
  const size_t work = (K * kGcCand + (size_t)kGcCand * 16 * p->cin) * sizeof(float);
  const size_t wbytes = K * p->cout * sizeof(float);
  a.w_in_smem = work + wbytes <= 200 * 1024;
  const size_t smem = work + (a.w_in_smem ? wbytes : 0);
  if (smem > 200 * 1024) return set_error(PGMP_ERR_INVALID, "cin too large for the patch buffers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t groups = (p->num_nodes + kGcCand - 1) / kGcCand;
  const unsigned grid = (unsigned)(groups < sms ? groups : sms);
  if (p->cin == 32) {            // the w32 backbone (default_config.py:48)
    PGMP_CUDA(cudaFuncSetAttribute(gather_conv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PGMP_LAUNCH(gather_conv_kernel<32>, grid, kGcThreads, smem, st, a);
  } else {
    PGMP_CUDA(cudaFuncSetAttribute(gather_conv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PGMP_LAUNCH(gather_conv_kernel<0>, grid, kGcThreads, smem, st, a);
  }
  return PGMP_OK;
}

// ------------------------------------------------------------------------------------------------
// Reverse of the plain node-feature gather x[n, :] = features[b, :, y, x] (ConstructGraph.py:265, 269) under autograd
// (end-to-end training, train.py:232): d_features[b, :, y, x] = sum of grad_x[n, :] over the nodes at that pixel.
// Candidates of different joint types can share a pixel; the first node of a pixel (in node order) sums its
// duplicates in node order and stores the result, so there is no floating-point atomic and the sum order is fixed.
// ------------------------------------------------------------------------------------------------
namespace pgmp {
namespace {

constexpr int kGbMaxChunks = 8;   // channels <= 256

__global__ void __launch_bounds__(256) gather_backward_kernel(const float* __restrict__ grad_x, const int64_t* __restrict__ joint_det,
                                                               const int64_t* __restrict__ batch_index, int64_t N, int C,
                                                               float* __restrict__ df, int64_t sb, int64_t sc, int64_t sy, int64_t sx) {
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = batch_index[n], x = joint_det[n * 3], y = joint_det[n * 3 + 1];
  // the image's nodes are contiguous (batch_index ascends): [lo, hi)
  int64_t lo = 0, hi = n;
  while (lo < hi) {       // first index with batch_index == b
    const int64_t mid = (lo + hi) >> 1;
    if (batch_index[mid] < b) lo = mid + 1; else hi = mid;
  }
  const int64_t first = lo;
  lo = n; hi = N;
  while (lo < hi) {       // first index with batch_index > b
    const int64_t mid = (lo + hi) >> 1;
    if (batch_index[mid] <= b) lo = mid + 1; else hi = mid;
  }
  const int64_t last = lo;
  float acc[kGbMaxChunks];
#pragma unroll
  for (int i = 0; i < kGbMaxChunks; ++i) acc[i] = (lane + 32 * i < C) ? grad_x[n * C + lane + 32 * i] : 0.f;
  for (int64_t base = first; base < last; base += 32) {
    const int64_t m = base + lane;
    const bool same = m < last && m != n && joint_det[m * 3] == x && joint_det[m * 3 + 1] == y;
    uint32_t hits = __ballot_sync(0xffffffffu, same);
    if (base < n && (hits & (n - base >= 32 ? 0xffffffffu : ((1u << (n - base)) - 1u)))) return;   // an earlier node owns the pixel
    while (hits) {          // later nodes at the same pixel, ascending
      const int64_t d = base + (__ffs(hits) - 1);
      hits &= hits - 1;
#pragma unroll
      for (int i = 0; i < kGbMaxChunks; ++i)
        if (lane + 32 * i < C) acc[i] += grad_x[d * C + lane + 32 * i];
    }
  }
  float* __restrict__ dst = df + b * sb + y * sy + x * sx;
#pragma unroll
  for (int i = 0; i < kGbMaxChunks; ++i)
    if (lane + 32 * i < C) dst[(int64_t)(lane + 32 * i) * sc] = acc[i];
}

}  // namespace
}  // namespace pgmp

extern "C" int pgmp_gc_gather_backward(const float* grad_x, const int64_t* joint_det, const int64_t* batch_index,
                                       int64_t num_nodes, int32_t channels, float* d_features, int64_t stride_b,
                                       int64_t stride_c, int64_t stride_y, int64_t stride_x, pgmp_stream_t stream) {
  using namespace pgmp;
  if (num_nodes < 0 || channels <= 0 || channels > 32 * kGbMaxChunks)
    return set_error(PGMP_ERR_INVALID, "bad sizes (channels <= %d)", 32 * kGbMaxChunks);
  if (num_nodes == 0) return PGMP_OK;
  if (!grad_x || !joint_det || !batch_index || !d_features) return set_error(PGMP_ERR_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PGMP_LAUNCH(gather_backward_kernel, (unsigned)ceil_div<int64_t>(num_nodes * 32, 256), 256, 0, st, grad_x, joint_det,
              batch_index, num_nodes, (int)channels, d_features, stride_b, stride_c, stride_y, stride_x);
  return PGMP_OK;
}

// ------------------------------------------------------------------------------------------------
// Reverse pass of pgmp_gc_gather_conv (end-to-end training with ConvUpsampleFeatures: the gradients of the
// feature_gather convolution and of the backbone map, PoseEstimation.py:64-66, train.py:232).
//   x = bias + P Wmat,  P[n][k] = the interpolated input patch of node n  (k = (ky 3 + kx) Cin + ci)
//   d Wmat = P^T d x,  d bias = sum_n d x,  d P = d x Wmat^T                (plain products: the caller's library GEMM)
//   d feat[b, ci, y_t + ky - 1, x_t + kx - 1] += ly[t] lx[t] d P[n][k]     (this file)
// pgmp_gc_gather_conv_patches writes P with the forward's arithmetic; pgmp_gc_gather_conv_backward scatters d P into
// the (zero-filled) map WITHOUT atomics: one CTA per image walks its nodes in node order, the threads of the CTA own
// the 16 x Cin (neighbourhood position, channel) pairs of the current node -- different addresses within a node,
// nodes strictly one after the other -- so colliding neighbourhoods are summed in a fixed order.
// ------------------------------------------------------------------------------------------------
namespace pgmp {
namespace {

__global__ void __launch_bounds__(256) gather_conv_patches_kernel(const GatherConvArgs a, float* __restrict__ patches) {
  const int K = 9 * a.cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n * K) return;
  const int64_t n = i / K;
  const int k = (int)(i - n * K), ci = k % a.cin, kk = k / a.cin, ky = kk / 3, kx = kk - 3 * ky;
  int y0, y1, x0, x1;
  float ly[2], lx[2];
  bilinear_tap((int)a.joint_det[n * 3 + 1], a.h, a.out_h, y0, y1, ly[0], ly[1]);
  bilinear_tap((int)a.joint_det[n * 3 + 0], a.w, a.out_w, x0, x1, lx[0], lx[1]);
  const float* __restrict__ f = a.feat + a.batch_index[n] * a.sb + ci * a.sc;
  float p = 0.f;
#pragma unroll
  for (int ty = 0; ty < 2; ++ty)
#pragma unroll
    for (int tx = 0; tx < 2; ++tx) {
      const int yy = (ty ? y1 : y0) + ky - 1, xx = (tx ? x1 : x0) + kx - 1;
      const float v = ((unsigned)yy < (unsigned)a.h && (unsigned)xx < (unsigned)a.w) ? __ldg(f + yy * a.sy + xx * a.sx) : 0.f;
      p = __fadd_rn(p, __fmul_rn(__fmul_rn(ly[ty], lx[tx]), v));
    }
  patches[i] = p;
}

__global__ void __launch_bounds__(512) gather_conv_scatter_kernel(const GatherConvArgs a, const float* __restrict__ d_patches,
                                                                   float* __restrict__ df) {
  const int b = blockIdx.x;
  // this image's nodes [first, last): batch_index ascends
  int64_t lo = 0, hi = a.n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a.batch_index[mid] < b) lo = mid + 1; else hi = mid; }
  const int64_t first = lo;
  hi = a.n;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a.batch_index[mid] <= b) lo = mid + 1; else hi = mid; }
  const int64_t last = lo;
  const int K = 9 * a.cin, pairs = 16 * a.cin;
  float* __restrict__ plane = df + (int64_t)b * a.sb;
  for (int64_t n = first; n < last; ++n) {
    int y0, y1, x0, x1;
    float ly[2], lx[2];
    bilinear_tap((int)a.joint_det[n * 3 + 1], a.h, a.out_h, y0, y1, ly[0], ly[1]);
    bilinear_tap((int)a.joint_det[n * 3 + 0], a.w, a.out_w, x0, x1, lx[0], lx[1]);
    const int dy = y1 - y0, dx = x1 - x0;
    const float* __restrict__ dp = d_patches + n * K;
    for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
      const int ci = i % a.cin, pos = i / a.cin, py = pos >> 2, px = pos & 3;
      const int yy = y0 - 1 + py, xx = x0 - 1 + px;
      if ((unsigned)yy >= (unsigned)a.h || (unsigned)xx >= (unsigned)a.w) continue;   // the convolution's zero padding
      float c = 0.f;
      bool any = false;
#pragma unroll
      for (int ty = 0; ty < 2; ++ty)
#pragma unroll
        for (int tx = 0; tx < 2; ++tx) {
          const int ky = py - (ty ? dy : 0), kx = px - (tx ? dx : 0);
          if ((unsigned)ky < 3u && (unsigned)kx < 3u) {
            c = __fadd_rn(c, __fmul_rn(__fmul_rn(ly[ty], lx[tx]), dp[(ky * 3 + kx) * a.cin + ci]));
            any = true;
          }
        }
      if (any) {
        float* __restrict__ q = plane + ci * a.sc + yy * a.sy + xx * a.sx;
        *q = __fadd_rn(*q, c);
      }
    }
    __syncthreads();        // the next node may touch the same pixels
  }
}

int fill_gather_conv_args(const pgmp_gather_conv_params* p, GatherConvArgs& a) {
  if (!p) return set_error(PGMP_ERR_INVALID, "null params");
  if (p->num_nodes < 0 || p->cin <= 0 || p->height <= 0 || p->width <= 0 || p->out_height <= 0 || p->out_width <= 0)
    return set_error(PGMP_ERR_INVALID, "bad sizes");
  if (p->num_nodes > 0 && (!p->features || !p->joint_det || !p->batch_index)) return set_error(PGMP_ERR_INVALID, "null pointer");
  a.feat = p->features; a.sb = p->feat_stride_b; a.sc = p->feat_stride_c; a.sy = p->feat_stride_y; a.sx = p->feat_stride_x;
  a.cin = p->cin; a.h = p->height; a.w = p->width; a.cout = p->cout; a.out_h = p->out_height; a.out_w = p->out_width;
  a.wmat = p->weight_t; a.bias = p->bias; a.joint_det = p->joint_det; a.batch_index = p->batch_index; a.n = p->num_nodes;
  a.x = p->x; a.w_in_smem = 0;
  return PGMP_OK;
}

}  // namespace
}  // namespace pgmp

extern "C" int pgmp_gc_gather_conv_patches(const pgmp_gather_conv_params* p, float* patches, pgmp_stream_t stream) {
  using namespace pgmp;
  GatherConvArgs a;
  const int rc = fill_gather_conv_args(p, a);
  if (rc != PGMP_OK) return rc;
  if (a.n == 0) return PGMP_OK;
  if (!patches) return set_error(PGMP_ERR_INVALID, "null pointer");
  const int64_t total = a.n * 9 * a.cin;
  PGMP_LAUNCH(gather_conv_patches_kernel, (unsigned)ceil_div<int64_t>(total, 256), 256, 0, static_cast<cudaStream_t>(stream), a, patches);
  return PGMP_OK;
}

extern "C" int pgmp_gc_gather_conv_backward(const pgmp_gather_conv_params* p, const float* d_patches, int32_t batch,
                                            float* d_features, pgmp_stream_t stream) {
  using namespace pgmp;
  GatherConvArgs a;
  const int rc = fill_gather_conv_args(p, a);
  if (rc != PGMP_OK) return rc;
  if (a.n == 0 || batch <= 0) return PGMP_OK;
  if (!d_patches || !d_features) return set_error(PGMP_ERR_INVALID, "null pointer");
  PGMP_LAUNCH(gather_conv_scatter_kernel, (unsigned)batch, 512, 0, static_cast<cudaStream_t>(stream), a, d_patches, d_features);
  return PGMP_OK;
}
