// Minimal hand-written tcgen05 / TMEM / mbarrier layer for sm_100a (inline PTX; no CUTLASS).
//
// Operand tiles live in shared memory in the canonical K-major SWIZZLE_128B layout: rows of 128 bytes
// (64 bf16), groups of 8 rows = 1024 bytes, the 16-byte chunk c of row r stored at chunk position
// c ^ (r & 7).  Tile bases are 1024-byte aligned.  Accumulators are fp32 in tensor memory: row i of
// the 128-row tile is TMEM lane i, output column j is TMEM column j.
//
// Precision scheme "bf16x3": an fp32 operand x is split into hi = bf16(x), lo = bf16(x - hi);
// A.W ~= Ah.Wh + Ah.Wl + Al.Wh with fp32 accumulation (relative error ~2^-16 per product term),
// which keeps 10 weight-shared recurrent steps within the 1e-3 logit tolerance where plain bf16
// (measured 4e-3 .. 1e-2) and tf32 (about 1e-3) do not.
#pragma once

#include <cuda_bf16.h>
#include <stdint.h>

namespace pgmp {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 24)) __trap();   // a lost MMA completion must fail loudly, never hang the GPU
  } while (!done);
}
// ---- 1-D bulk asynchronous copies (TMA engine without a tensor map); sizes / addresses are multiples of 16 bytes
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared; completion is signalled on `bar` as `bytes` of transaction count
__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
               "r"(smem_src), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// bring a contiguous global range into L2 ahead of the bulk load that will need it
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(gsrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp (elect.sync): ptxas then knows the guarded code runs in a single thread and moves the
// MMA / bulk-copy operands to uniform registers directly instead of wrapping every instruction in a waterfall loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}

// ---- tensor memory ---------------------------------------------------------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from tensor memory (TS form): row i of A in lane i, element k of the row in bits [16 (k & 1), +16) of column k / 2
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// the whole 64-column accumulator row of this thread
__device__ __forceinline__ void tmem_ld64(uint32_t tmem_base, int col0, float (&v)[64]) {
  const uint32_t lane_base = ((threadIdx.x >> 5) & 3) * 32;
  const uint32_t t = tmem_base + (lane_base << 16) + (uint32_t)col0;
  tmem_ld16(t + 0, v + 0);
  tmem_ld16(t + 16, v + 16);
  tmem_ld16(t + 32, v + 32);
  tmem_ld16(t + 48, v + 48);
  tmem_ld_wait();
}

// ---- descriptors ------------------------------------------------------------------------------
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major, 1),
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups), [46,48) version = 1, [61,64) layout = 2
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// instruction descriptor, kind::f16: D fp32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[128 x N] (+)= A[128 x K] . W[N x K]^T with the bf16x3 split; A and W tiles are [rows][64 bf16] SW128 blocks,
// K / 64 blocks each, block b of A at a + b * a_block_bytes (same for W).  One thread issues.
template <int N>
__device__ __forceinline__ void issue_gemm_x3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_block_bytes,
                                              uint32_t w_hi, uint32_t w_lo, uint32_t w_block_bytes, int k_blocks,
                                              bool accumulate) {
  constexpr uint32_t idesc = idesc_bf16(128, N);
  uint32_t acc = accumulate ? 1u : 0u;
  for (int b = 0; b < k_blocks; ++b) {
    const uint64_t ah = smem_desc_sw128(a_hi + b * a_block_bytes), al = smem_desc_sw128(a_lo + b * a_block_bytes);
    const uint64_t wh = smem_desc_sw128(w_hi + b * w_block_bytes), wl = smem_desc_sw128(w_lo + b * w_block_bytes);
#pragma unroll
    for (int k = 0; k < 4; ++k) {   // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
      mma_bf16(tmem_d, ah + 2 * k, wh + 2 * k, idesc, acc);
      acc = 1u;
      mma_bf16(tmem_d, ah + 2 * k, wl + 2 * k, idesc, 1u);
      mma_bf16(tmem_d, al + 2 * k, wh + 2 * k, idesc, 1u);
    }
  }
}

// The same product with the A operand in tensor memory (TS form): a_hi / a_lo are TMEM addresses of [128 x 64] bf16
// operands (32 columns each, element k of a row in the 16-bit half k & 1 of column k / 2).  K = 64 only.
template <int N>
__device__ __forceinline__ void issue_gemm_x3_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo,
                                                 bool accumulate) {
  constexpr uint32_t idesc = idesc_bf16(128, N);
  uint32_t acc = accumulate ? 1u : 0u;
  const uint64_t wh = smem_desc_sw128(w_hi), wl = smem_desc_sw128(w_lo);
#pragma unroll
  for (int k = 0; k < 4; ++k) {   // UMMA_K = 16 bf16 = 8 TMEM columns of A = 2 descriptor units of W
    mma_bf16_ts(tmem_d, a_hi + 8 * k, wh + 2 * k, idesc, acc);
    acc = 1u;
    mma_bf16_ts(tmem_d, a_hi + 8 * k, wl + 2 * k, idesc, 1u);
    mma_bf16_ts(tmem_d, a_lo + 8 * k, wh + 2 * k, idesc, 1u);
  }
}

// ---- explicit shared-state-space accessors (32-bit shared addresses; avoids generic ST/LD) ----------
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
// 16-byte asynchronous global -> shared copy (LDGSTS), no registers held while in flight
__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// sub-CTA barrier: `count` threads meet on hardware barrier `id` (id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// float index of (row, col) in a swizzled [128][64] fp32 staging tile: the 16-byte chunk c of row r sits at chunk
// position c ^ (r & 15), so both row-per-thread and row-per-half-warp accesses are bank-conflict free.  The
// per-node tables and the C rows are stored in global memory as such tile images (fetched / written by bulk copies).
__device__ __forceinline__ int stage_index(int row, int col) {
  return row * 64 + ((((col >> 2) ^ (row & 15)) << 2) | (col & 3));
}

// The per-node tables (P, Q, R) use a 32-byte-granular variant: the 32-byte pair m of row r sits at pair position
// m ^ (r & 7).  Row-per-thread 16-byte shared-memory stores stay bank-conflict free (the 8 lanes of a phase hit 8
// different pairs), and the step kernel fetches a row with 256-bit loads -- half the requests of 16-byte chunks
// for gathers whose every lane touches a different row (the L1 request rate is what bounds them).
__device__ __forceinline__ int table_index(int row, int col) {
  return row * 64 + ((((col >> 3) ^ (row & 7)) << 3) | (col & 7));
}
struct float8 { float4 a, b; };
__device__ __forceinline__ float8 ldg256(const float* p) {       // 32-byte aligned, read-only path
  float8 v;
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v.a.x), "=f"(v.a.y), "=f"(v.a.z), "=f"(v.a.w), "=f"(v.b.x), "=f"(v.b.y), "=f"(v.b.z), "=f"(v.b.w)
               : "l"(p));
  return v;
}

// ---- operand tile writers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk16) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk16 ^ (row & 7)) << 4));
}

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// write 4 consecutive fp32 values (columns col4*4 .. +3 of `row`) into the hi / lo tiles
__device__ __forceinline__ void store_split4(uint8_t* hi_tile, uint8_t* lo_tile, int row, int col4, float4 v) {
  __nv_bfloat16 h0, h1, h2, h3, l0, l1, l2, l3;
  split_bf16(v.x, h0, l0);
  split_bf16(v.y, h1, l1);
  split_bf16(v.z, h2, l2);
  split_bf16(v.w, h3, l3);
  const uint32_t off = sw128_offset(row, col4 >> 1) + (uint32_t)((col4 & 1) << 3);
  *reinterpret_cast<uint2*>(hi_tile + off) = make_uint2(pack2(h0, h1), pack2(h2, h3));
  *reinterpret_cast<uint2*>(lo_tile + off) = make_uint2(pack2(l0, l1), pack2(l2, l3));
}

// write this thread's whole 64-wide row (values v[0..63]) into the hi / lo tiles (8 x 16-byte chunks each)
__device__ __forceinline__ void store_split_row(uint8_t* hi_tile, uint8_t* lo_tile, int row, const float (&v)[64]) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_bf16(v[8 * c + i], h[i], l[i]);
    const uint32_t off = sw128_offset(row, c);
    *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(pack2(h[0], h[1]), pack2(h[2], h[3]), pack2(h[4], h[5]), pack2(h[6], h[7]));
    *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
  }
}

// copy a [rows][64] bf16 row-major global block (K contiguous) into a SW128 tile, all threads of the CTA
__device__ __forceinline__ void load_weight_tile(uint8_t* tile, const __nv_bfloat16* __restrict__ src, int rows, int src_ld) {
  for (int idx = threadIdx.x; idx < rows * 8; idx += blockDim.x) {
    const int r = idx >> 3, c = idx & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * src_ld + c * 8));
    *reinterpret_cast<uint4*>(tile + sw128_offset(r, c)) = v;
  }
}

// two values at a time: one packed conversion for the hi parts, one for the lo parts (same results as split_bf16)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);          // .x (low half) = a, .y (high half) = b
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// packed fp32 pairs (FADD2 / FFMA2 of sm_100: one issue slot for two lanes of arithmetic)
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};"
      " fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
// ReLU folded into the split: hi = bf16_rz(max(v, 0)) (toward zero, so v - hi >= 0 for v >= 0 and hi = 0 for v < 0),
// lo = bf16_rn(max(v - hi, 0)).  hi + lo = max(v, 0) to 2^-17 relative (round-to-nearest hi: 2^-18); no FMNMX.
__device__ __forceinline__ void split2_relu(float2 v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.y), "f"(v.x));
  const float2 r = add2(v, make_float2(-__uint_as_float(hi << 16), -__uint_as_float(hi & 0xffff0000u)));
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}
// the same for values that are already non-negative
__device__ __forceinline__ void split2_pos(float2 v, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.y), "f"(v.x));
  const float2 r = add2(v, make_float2(-__uint_as_float(hi << 16), -__uint_as_float(hi & 0xffff0000u)));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r.y), "f"(r.x));
}
// this thread's 64-wide row as a TS-form A operand: hi pairs -> 32 columns at t_hi, lo pairs -> 32 columns at t_lo
// (t_* already carry the lane base of the warp)
__device__ __forceinline__ void store_split_row_tmem(uint32_t t_hi, uint32_t t_lo, const float (&v)[64]) {
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) split2(v[32 * part + 2 * i], v[32 * part + 2 * i + 1], h[i], l[i]);
    tmem_st16(t_hi + 16 * part, h);
    tmem_st16(t_lo + 16 * part, l);
  }
  tmem_st_wait();
}
// address-based variants (shared-space stores)
__device__ __forceinline__ void store_split4_a(uint32_t hi_tile, uint32_t lo_tile, int row, int col4, float4 v) {
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  const uint32_t off = sw128_offset(row, col4 >> 1) + (uint32_t)((col4 & 1) << 3);
  sts64(hi_tile + off, h0, h1);
  sts64(lo_tile + off, l0, l1);
}
__device__ __forceinline__ void store_split_row_a(uint32_t hi_tile, uint32_t lo_tile, int row, const float (&v)[64]) {
  const uint32_t row_off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
  const uint32_t x = (uint32_t)(row & 7);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split2(v[8 * c + 2 * i], v[8 * c + 2 * i + 1], h[i], l[i]);
    const uint32_t off = row_off + (((uint32_t)c ^ x) << 4);
    sts128(hi_tile + off, h[0], h[1], h[2], h[3]);
    sts128(lo_tile + off, l[0], l[1], l[2], l[3]);
  }
}
// fp32 values of columns col4*4 .. +3 of `row` reconstructed from the hi / lo tiles (hi + lo is exact in fp32)
__device__ __forceinline__ float4 load_joined4_a(uint32_t hi_tile, uint32_t lo_tile, int row, int col4) {
  const uint32_t off = sw128_offset(row, col4 >> 1) + (uint32_t)((col4 & 1) << 3);
  const uint2 h = lds64(hi_tile + off), l = lds64(lo_tile + off);
  float4 v;
  v.x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
  v.y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
  v.z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
  v.w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  return v;
}
// weight tile copy with `n` cooperating threads (thread index `t`)
__device__ __forceinline__ void load_weight_tile_a(uint32_t tile, const __nv_bfloat16* __restrict__ src, int rows, int src_ld,
                                                   int t, int n) {
  for (int idx = t; idx < rows * 8; idx += n) {
    const int r = idx >> 3, c = idx & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * src_ld + c * 8));
    sts128(tile + sw128_offset(r, c), v.x, v.y, v.z, v.w);
  }
}

}  // namespace umma
}  // namespace pgmp
