// tcgen05 / TMEM version of the E-level forward products of the training step (PGMP_TRAIN_TC=1; mpn_train.cu keeps the
// 3xTF32 mma.sync kernels as the parity mode).  Same operand treatment as the inference kernels (umma.cuh): every fp32
// operand is split into a bf16 hi / lo pair and A W^T ~ Ah Wh + Ah Wl + Al Wh is accumulated in fp32 in tensor memory
// (~1e-5 relative to the fp32 product).  One CTA = 128 rows at a time: the A rows are read with coalesced 16-byte loads,
// split and written as SWIZZLE_128B operand tiles, the weights (fp32 in the flat parameter buffer, different every step)
// are split by the CTA once, 12 or 24 UTCHMMA per tile are issued by one elected thread, and the epilogue -- one
// accumulator row per thread -- adds the bias and the gathered per-node table rows, applies the ReLU and stores the fp32
// row the reverse pass reads.  Two CTAs per SM overlap each other's load, product and epilogue.
#include "mpn_common.cuh"
#include "mpn_train_tc.cuh"
#include "umma.cuh"

namespace pgmp {
namespace {

using namespace umma;

constexpr int kATileB = kTile * 128;              // bytes of one [128][64] bf16 operand tile
constexpr int kWTileB = kD * 128;                 // bytes of one [64][64] bf16 weight tile
constexpr size_t lin_tc_smem(int kb) { return (size_t)kb * (2 * kATileB + 2 * kWTileB) + 64 + 1024; }
static_assert(2 * (lin_tc_smem(2) + 1024) <= 228 * 1024 && 3 * (lin_tc_smem(1) + 1024) <= 228 * 1024, "CTAs per SM");

// (168 registers: three CTAs per SM for K = 64 -- 49 KB each -- two for K = 128)
__global__ void __launch_bounds__(kTile, 3) lin_fwd_tc_kernel(const LinFwdTc q) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(base);
  const int tid = threadIdx.x;
  const int kb = q.a1 ? 2 : 1;
  const uint32_t a_t = sb;                               // K-block b: hi at a_t + b * 2 * kATileB, lo at + kATileB
  const uint32_t w_t = sb + kb * 2 * kATileB;            // K-block b: hi at w_t + b * 2 * kWTileB, lo at + kWTileB
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + kb * (2 * kATileB + 2 * kWTileB));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  if (tid < 32) tmem_alloc<64>(tmem_slot);
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  // the weights: element (o, k) -> row o of K-block k / 64 (K-major, like nn.Linear.weight)
  for (int b = 0; b < kb; ++b)
    for (int i = tid; i < kD * 16; i += kTile) {
      const int o = i >> 4, c4 = i & 15;
      const float4 v = __ldg(reinterpret_cast<const float4*>(q.W + (size_t)o * q.ldw + q.coloff + b * kD + 4 * c4));
      store_split4_a(w_t + b * 2 * kWTileB, w_t + b * 2 * kWTileB + kWTileB, o, c4, v);
    }
  fence_before_sync();
  fence_async_smem();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const int64_t tiles = (q.M + kTile - 1) / kTile;
  uint32_t phase = 0;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kTile;
    for (int b = 0; b < kb; ++b) {
      const float* __restrict__ src = b ? q.a1 : q.a0;
      const int ld = b ? q.lda1 : q.lda0;
      float4 v[16];                                      // the whole K-block in flight before the first split
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = tid + i * kTile;
        const int r = idx >> 4, c4 = idx & 15;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < q.M) v[i] = __ldg(reinterpret_cast<const float4*>(src + (r0 + r) * ld + 4 * c4));
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = tid + i * kTile;
        store_split4_a(a_t + b * 2 * kATileB, a_t + b * 2 * kATileB + kATileB, idx >> 4, idx & 15, v[i]);
      }
    }
    fence_before_sync();
    fence_async_smem();
    __syncthreads();
    if (tid < 32 && elect_one()) {
      fence_after_sync();
      issue_gemm_x3<kD>(tmem, a_t, a_t + kATileB, 2 * kATileB, w_t, w_t + kWTileB, 2 * kWTileB, kb, false);
      mma_commit(bar);
    }
    // what this thread adds in the epilogue -- bias + the gathered table rows -- is summed while the product runs
    const int64_t row = r0 + tid;
    const bool live = row < q.M;
    float4 e[kD / 4];
#pragma unroll
    for (int c = 0; c < kD / 4; ++c) e[c] = q.bias ? __ldg(reinterpret_cast<const float4*>(q.bias) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (q.add1 && live) {
      const float4* __restrict__ g1 = reinterpret_cast<const float4*>(q.add1 + q.idx1[row] * q.add1_ld);
#pragma unroll
      for (int c = 0; c < kD / 4; ++c) { const float4 t = __ldg(g1 + c); e[c].x += t.x; e[c].y += t.y; e[c].z += t.z; e[c].w += t.w; }
    }
    if (q.add2 && live) {
      const float4* __restrict__ g2 = reinterpret_cast<const float4*>(q.add2 + q.idx2[row] * kD);
#pragma unroll
      for (int c = 0; c < kD / 4; ++c) { const float4 t = __ldg(g2 + c); e[c].x += t.x; e[c].y += t.y; e[c].z += t.z; e[c].w += t.w; }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    fence_after_sync();
    float d[kD];
    tmem_ld64(tmem, 0, d);
    if (live) {
      float4* __restrict__ y4 = reinterpret_cast<float4*>(q.Y + row * q.ldy);
#pragma unroll
      for (int c = 0; c < kD / 4; ++c) {
        float4 v = make_float4(d[4 * c] + e[c].x, d[4 * c + 1] + e[c].y, d[4 * c + 2] + e[c].z, d[4 * c + 3] + e[c].w);
        if (q.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        y4[c] = v;
      }
    }
    fence_before_sync();
    __syncthreads();          // the accumulator has been read and the operand tiles are free for the next 128 rows
  }
  fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc<64>(tmem);
}

}  // namespace

int launch_lin_fwd_tc(cudaStream_t st, const LinFwdTc& q) {
  if (q.M <= 0) return PGMP_OK;
  static bool attr = false;
  if (!attr) {
    PGMP_CUDA(cudaFuncSetAttribute(lin_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_tc_smem(2)));
    attr = true;
  }
  const int kb = q.a1 ? 2 : 1;
  const int per_sm = kb == 1 ? 3 : 2;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (q.M + kTile - 1) / kTile;
  const unsigned grid = (unsigned)(tiles < (int64_t)per_sm * sms ? tiles : (int64_t)per_sm * sms);
  PGMP_LAUNCH(lin_fwd_tc_kernel, grid, kTile, lin_tc_smem(kb), st, q);
  return PGMP_OK;
}

}  // namespace pgmp
