// Training step of the message-passing network with the type-agnostic MPLayer (include/pgmp.h,
// pgmp_mpn_train_*): forward in train() mode and the reverse pass.
//
// Reference: NodeClassificationMPNSimple.forward (NodeClassificationMPNSimple.py:62-97), MPLayer
// (layers.py:32-86), _make_mlp (layers.py:8-29; ReLU BEFORE BatchNorm1d), torch autograd for the gradients
// (train.py:232-236).  Checked against oracle/mpn_train.py, which is pinned to the reference under float64 autograd.
//
// Structure.  The same column split of mlp_edge.0 / mlp_node.0 the inference path uses keeps the per-edge work at
// 64-wide products: with x = [h0 ; h] and e = [g0 ; g]
//     hidden = ReLU(W1_e e + b1 + P[dst] + Q[src]),  P = W1_dst x, Q = W1_src x        (per node)
//     m      = ReLU(Wm_e g' + R[dst]),               R = Wm_x x + bm                   (per node)
// and in the reverse pass the gradients that flow to x through the gathers are summed per node FIRST
// (S_dst = sum over edges into n of d hidden, ...) and multiplied by the weights once per node:
//     dx = S_dst W1_dst + S_src W1_src + T_dst Wm_x,    dW1_dst = S_dst^T x, ...
// The per-node sums run over CSR bins ordered by edge id, weight gradients are reduced from per-CTA partial
// tiles in a fixed order: no floating-point atomics anywhere, results are reproducible.
// Arithmetic: the products run on the tensor cores with fp32-accurate 3xTF32 operands (see tile_mma), everything else is
// fp32 SIMT; BatchNorm statistics accumulate in fp64.  Moving the E-level products to tcgen05 (bf16x3 operand images as in
// mpn_tc.cu) is the next step.
#include "common.cuh"
#include "mpn_train_tc.cuh"

namespace pgmp {
namespace {

constexpr int kD = 64;
constexpr int BM = 64, BN = 64, BK = 16;
constexpr int kMaxSplits = 512;
constexpr int kMaxWidth = 128;
constexpr int kMaxSteps = 64, kMaxOut = 32;
constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;
constexpr uint32_t kFull = 0xffffffffu;

// ------------------------------------------------------------------ operand descriptors
struct ASeg {
  const float* p;
  int ld, w;
};
struct ASrc {   // up to two column blocks side by side (torch.cat(..., dim=1) without materialising it)
  int n;
  ASeg s[2];
};
inline ASrc src1(const float* p, int w) { return ASrc{1, {{p, w, w}, {nullptr, 0, 0}}}; }
inline ASrc src2(const float* p, const float* q, int w) { return ASrc{2, {{p, w, w}, {q, w, w}}}; }
inline int src_width(const ASrc& a) { return a.s[0].w + (a.n > 1 ? a.s[1].w : 0); }

__device__ __forceinline__ float a_load(const ASrc& a, int64_t r, int k) {
  if (k < a.s[0].w) return a.s[0].p[r * a.s[0].ld + k];
  k -= a.s[0].w;
  if (a.n > 1 && k < a.s[1].w) return a.s[1].p[r * a.s[1].ld + k];
  return 0.f;
}

// ------------------------------------------------------------------ tensor-core inner product, fp32-accurate
// The three GEMM kernels keep a 64 x 64 output tile per CTA and multiply 16-deep slices out of shared memory with
// mma.sync.m16n8k8 (tf32 operands, fp32 accumulation).  Every fp32 operand is split on the fly into
// hi = tf32(x), lo = tf32(x - hi) and a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi ("3xTF32"): the dropped lo.lo term is
// 2^-22 relative, so the result matches an fp32 FMA chain to rounding (the TIGHT tolerance of tests/test_gpu_train.py).
// Warp w of the 8 owns rows (w & 3) * 16 .. +16 and columns (w >> 2) * 32 .. +32 of the tile (4 n-blocks of 8).
constexpr int AS_LD = BK + 4;    // row-major A slice [64][16]: 20-float rows -> conflict-free fragment loads
constexpr int BS_LD = BN + 8;    // k-major slices [16][64]: 72-float rows

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// acc += A_slice . B_slice.  A_KMAJOR = false: A element (row, k) at Ar[row * AS_LD + k]; true: at Ak[k * BS_LD + row].
// B element (k, col) at Bk[k * BS_LD + col].
// acc += A_slice . B_slice.
// A_KMAJOR = false: A element (row, k) at Aop[row * AS_LD + k]; true: at Aop[k * BS_LD + row].
// B_NMAJOR = false: B element (k, col) at Bop[k * BS_LD + col]; true: at Bop[col * AS_LD + k].
template <bool A_KMAJOR, bool B_NMAJOR>
__device__ __forceinline__ void tile_mma(const float* __restrict__ Aop, const float* __restrict__ Bop, int wr, int wc, int g,
                                         int t4, float (&acc)[4][4]) {
#pragma unroll
  for (int k8 = 0; k8 < BK; k8 += 8) {
    uint32_t ah[4], al[4];
    if (A_KMAJOR) {
      split_tf32(Aop[(k8 + t4) * BS_LD + wr + g], ah[0], al[0]);
      split_tf32(Aop[(k8 + t4) * BS_LD + wr + g + 8], ah[1], al[1]);
      split_tf32(Aop[(k8 + t4 + 4) * BS_LD + wr + g], ah[2], al[2]);
      split_tf32(Aop[(k8 + t4 + 4) * BS_LD + wr + g + 8], ah[3], al[3]);
    } else {
      split_tf32(Aop[(wr + g) * AS_LD + k8 + t4], ah[0], al[0]);
      split_tf32(Aop[(wr + g + 8) * AS_LD + k8 + t4], ah[1], al[1]);
      split_tf32(Aop[(wr + g) * AS_LD + k8 + t4 + 4], ah[2], al[2]);
      split_tf32(Aop[(wr + g + 8) * AS_LD + k8 + t4 + 4], ah[3], al[3]);
    }
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      uint32_t bh[2], bl[2];
      if (B_NMAJOR) {
        split_tf32(Bop[(wc + nb * 8 + g) * AS_LD + k8 + t4], bh[0], bl[0]);
        split_tf32(Bop[(wc + nb * 8 + g) * AS_LD + k8 + t4 + 4], bh[1], bl[1]);
      } else {
        split_tf32(Bop[(k8 + t4) * BS_LD + wc + nb * 8 + g], bh[0], bl[0]);
        split_tf32(Bop[(k8 + t4 + 4) * BS_LD + wc + nb * 8 + g], bh[1], bl[1]);
      }
      mma_tf32(acc[nb], al, bh);      // small terms first
      mma_tf32(acc[nb], ah, bl);
      mma_tf32(acc[nb], ah, bh);
    }
  }
}
// accumulator element e of n-block nb: row = wr + g + 8 * (e >> 1), column = wc + nb * 8 + 2 * t4 + (e & 1)

// ------------------------------------------------------------------ operand slices into shared memory
// VEC (every row start 16-byte aligned, widths multiples of 4): 16-byte cp.async copies straight into a kStages-deep
// ring of slices, two slices in flight while one is multiplied.  Otherwise (19-wide edge attributes, 1 / 17-wide head
// outputs): scalar loads, one slice at a time.
constexpr int kStages = 3;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;      // 0 source bytes: the 16 destination bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// address of element (r, c) of a column-concatenated source; c is a multiple of 4 and the widths are too
__device__ __forceinline__ const float* src_ptr(const ASrc& a, int64_t r, int c) {
  const ASeg* s = &a.s[0];
  if (c >= s->w) { c -= s->w; s = &a.s[1]; }
  return s->p + r * s->ld + c;
}

template <bool VEC, class Issue, class Fill, class Compute>
__device__ __forceinline__ void run_slices(int n_iter, Issue issue, Fill fill, Compute compute) {
  if (VEC) {
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
      if (s < n_iter) issue(s, s);
      cp_commit();
    }
    for (int i = 0; i < n_iter; ++i) {
      cp_wait<kStages - 2>();
      __syncthreads();                      // slice i has landed for every thread; slice i - 1 is no longer read
      const int nx = i + kStages - 1;
      if (nx < n_iter) issue(nx, nx % kStages);
      cp_commit();
      compute(i % kStages);
    }
  } else {
    for (int i = 0; i < n_iter; ++i) {
      __syncthreads();
      fill(i);
      __syncthreads();
      compute(0);
    }
  }
}

// ------------------------------------------------------------------ Y = act(A W^T + b + add1[idx1] + add2[idx2])
struct FwdArgs {
  ASrc a;
  int64_t M;
  int K, O;
  const float* W;   // element (o, k) at W[o * ldw + coloff + k]
  int ldw, coloff;
  const float* bias;
  const float* add1;   // [.][64] rows gathered by idx1 (O == 64), or null
  const int64_t* idx1;
  const float* add2;
  const int64_t* idx2;
  int relu;
  float* Y;
  int ldy;
  int add1_ld;      // row stride of add1 (64, or num_types * 64 for the per-type tables)
};

template <bool VEC>
__global__ void __launch_bounds__(256) lin_fwd_kernel(const FwdArgs q) {
  constexpr int ST = VEC ? kStages : 1;
  __shared__ __align__(16) float As[ST][BM * AS_LD];     // rows of A, 16 reduction columns
  __shared__ __align__(16) float Bt[ST][BN * AS_LD];     // rows of W (outputs o), the same 16 reduction columns
  const int t = threadIdx.x;
  const int warp = t >> 5, g = (t & 31) >> 2, t4 = t & 3, wr = (warp & 3) * 16, wc = (warp >> 2) * 32;
  const int64_t r0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  auto issue = [&](int it, int st) {
    const int row = t >> 2, c = it * BK + (t & 3) * 4;
    const int64_t gr = r0 + row;
    const bool va = gr < q.M && c < q.K;
    cp_async16(&As[st][row * AS_LD + (t & 3) * 4], va ? src_ptr(q.a, gr, c) : q.W, va);
    const int o = n0 + row;
    const bool vb = o < q.O && c < q.K;
    cp_async16(&Bt[st][row * AS_LD + (t & 3) * 4], vb ? q.W + (int64_t)o * q.ldw + q.coloff + c : q.W, vb);
  };
  auto fill = [&](int it) {
    const int k = it * BK + (t & 15);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (t >> 4) + 16 * i;
      const int64_t gr = r0 + row;
      As[0][row * AS_LD + (t & 15)] = (gr < q.M && k < q.K) ? a_load(q.a, gr, k) : 0.f;
      const int o = n0 + row;
      Bt[0][row * AS_LD + (t & 15)] = (o < q.O && k < q.K) ? q.W[(int64_t)o * q.ldw + q.coloff + k] : 0.f;
    }
  };
  run_slices<VEC>(ceil_div(q.K, BK), issue, fill, [&](int st) { tile_mma<false, true>(As[st], Bt[st], wr, wc, g, t4, acc); });
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t gr = r0 + wr + g + 8 * h;
    if (gr >= q.M) continue;
    const float* e1 = q.add1 ? q.add1 + q.idx1[gr] * q.add1_ld : nullptr;
    const float* e2 = q.add2 ? q.add2 + q.idx2[gr] * kD : nullptr;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int o = n0 + wc + nb * 8 + 2 * t4 + j;
        if (o >= q.O) continue;
        float v = acc[nb][2 * h + j];
        if (q.bias) v += q.bias[o];
        if (e1) v += e1[o];
        if (e2) v += e2[o];
        if (q.relu) v = fmaxf(v, 0.f);
        q.Y[gr * q.ldy + o] = v;
      }
  }
}

// ------------------------------------------------------------------ dA = dY W, columns routed to their owners
struct Target {
  float* p;
  int ld, k0, w, add;   // columns [k0, k0 + w) of dA go to p[r * ld + (k - k0)], overwriting or adding
  const float* mask;    // same layout as p, or null: the result is zeroed where mask <= 0 (the ReLU in front of p)
};
struct BwdInArgs {
  const float* dY;
  int ldd;
  int64_t M;
  int O, K;
  const float* W;
  int ldw, coloff;
  int nt;
  Target t[2];
};

template <bool VEC>
__global__ void __launch_bounds__(256) lin_bwd_in_kernel(const BwdInArgs q) {
  constexpr int ST = VEC ? kStages : 1;
  __shared__ __align__(16) float As[ST][BM * AS_LD];     // rows of dY, 16 output columns o (the reduction index)
  __shared__ __align__(16) float Bs[ST][BK * BS_LD];     // the same 16 rows o of W, 64 input columns
  const int t = threadIdx.x;
  const int warp = t >> 5, g = (t & 31) >> 2, t4 = t & 3, wr = (warp & 3) * 16, wc = (warp >> 2) * 32;
  const int64_t r0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  auto issue = [&](int it, int st) {
    const int row = t >> 2, c = it * BK + (t & 3) * 4;
    const int64_t gr = r0 + row;
    const bool va = gr < q.M && c < q.O;
    cp_async16(&As[st][row * AS_LD + (t & 3) * 4], va ? q.dY + gr * q.ldd + c : q.W, va);
    const int ob = it * BK + (t >> 4), k = n0 + (t & 15) * 4;
    const bool vb = ob < q.O && k < q.K;
    cp_async16(&Bs[st][(t >> 4) * BS_LD + (t & 15) * 4], vb ? q.W + (int64_t)ob * q.ldw + q.coloff + k : q.W, vb);
  };
  auto fill = [&](int it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = (t >> 4) + 16 * i;
      const int64_t gr = r0 + row;
      const int o = it * BK + (t & 15);
      As[0][row * AS_LD + (t & 15)] = (gr < q.M && o < q.O) ? q.dY[gr * q.ldd + o] : 0.f;
      const int ob = it * BK + (t >> 6) + 4 * i, k = n0 + (t & 63);
      Bs[0][((t >> 6) + 4 * i) * BS_LD + (t & 63)] = (ob < q.O && k < q.K) ? q.W[(int64_t)ob * q.ldw + q.coloff + k] : 0.f;
    }
  };
  run_slices<VEC>(ceil_div(q.O, BK), issue, fill, [&](int st) { tile_mma<false, false>(As[st], Bs[st], wr, wc, g, t4, acc); });
#pragma unroll
  for (int nb = 0; nb < 4; ++nb)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = n0 + wc + nb * 8 + 2 * t4 + j;
      if (k >= q.K) continue;
      const int which = (k >= q.t[0].k0 && k < q.t[0].k0 + q.t[0].w) ? 0 : ((q.nt > 1 && k >= q.t[1].k0 && k < q.t[1].k0 + q.t[1].w) ? 1 : -1);
      if (which < 0) continue;
      float* const tp = which ? q.t[1].p : q.t[0].p;
      const float* const tm = which ? q.t[1].mask : q.t[0].mask;
      const int tld = which ? q.t[1].ld : q.t[0].ld, tk0 = which ? q.t[1].k0 : q.t[0].k0, tadd = which ? q.t[1].add : q.t[0].add;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t gr = r0 + wr + g + 8 * h;
        if (gr >= q.M) continue;
        const int64_t off = gr * tld + (k - tk0);
        float v = tadd ? tp[off] + acc[nb][2 * h + j] : acc[nb][2 * h + j];
        if (tm && !(tm[off] > 0.f)) v = 0.f;
        tp[off] = v;
      }
    }
}

// ------------------------------------------------------------------ dW partial tiles: part[s] = dY[rows_s]^T A[rows_s]
struct BwdWArgs {
  const float* dY;
  int ldd, O;
  ASrc a;
  int K;
  int64_t M, rows_per_split;
  float* part;    // [splits][O][K]
  float* partb;   // [splits][O] column sums of dY, or null
};

template <bool VEC>
__global__ void __launch_bounds__(256) lin_bwd_w_kernel(const BwdWArgs q) {
  constexpr int ST = VEC ? kStages : 1;
  __shared__ __align__(16) float Ds[ST][BK * BS_LD];     // 16 rows (the reduction index) of dY, 64 output columns
  __shared__ __align__(16) float As[ST][BK * BS_LD];     // the same 16 rows of A, 64 input columns
  const int t = threadIdx.x;
  const int warp = t >> 5, g = (t & 31) >> 2, t4 = t & 3, wr = (warp & 3) * 16, wc = (warp >> 2) * 32;
  const int k0 = blockIdx.x * BN, o0 = blockIdx.y * BN, split = blockIdx.z;
  const int64_t rb = (int64_t)split * q.rows_per_split;
  const int64_t re = min(rb + q.rows_per_split, q.M);
  const bool want_bias = q.partb && blockIdx.x == 0 && t < BN;   // thread t sums column o0 + t of dY
  float acc[4][4] = {};
  float bsum = 0.f;
  auto issue = [&](int it, int st) {
    const int64_t gr = rb + (int64_t)it * BK + (t >> 4);
    const int c = (t & 15) * 4;
    const bool vd = gr < re && o0 + c < q.O;
    cp_async16(&Ds[st][(t >> 4) * BS_LD + c], vd ? q.dY + gr * q.ldd + o0 + c : q.dY, vd);
    const bool va = gr < re && k0 + c < q.K;
    cp_async16(&As[st][(t >> 4) * BS_LD + c], va ? src_ptr(q.a, gr, k0 + c) : q.dY, va);
  };
  auto fill = [&](int it) {
    const int c = t & 63;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = (t >> 6) + 4 * i;
      const int64_t gr = rb + (int64_t)it * BK + r;
      const bool in = gr < re;
      Ds[0][r * BS_LD + c] = (in && o0 + c < q.O) ? q.dY[gr * q.ldd + o0 + c] : 0.f;
      As[0][r * BS_LD + c] = (in && k0 + c < q.K) ? a_load(q.a, gr, k0 + c) : 0.f;
    }
  };
  run_slices<VEC>((int)ceil_div<int64_t>(re - rb, BK), issue, fill, [&](int st) {
    tile_mma<true, false>(Ds[st], As[st], wr, wc, g, t4, acc);      // rows of the tile = outputs o, columns = inputs k
    if (want_bias) {
#pragma unroll
      for (int rr = 0; rr < BK; ++rr) bsum += Ds[st][rr * BS_LD + t];
    }
  });
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int o = o0 + wr + g + 8 * h;
    if (o >= q.O) continue;
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = k0 + wc + nb * 8 + 2 * t4 + j;
        if (k < q.K) q.part[((int64_t)split * q.O + o) * q.K + k] = acc[nb][2 * h + j];
      }
  }
  if (want_bias && o0 + t < q.O) q.partb[(int64_t)split * q.O + o0 + t] = bsum;
}

// dW[o][coloff + k] += sum_s part[s][o][k], db[o] += sum_s partb[s][o].  Fixed association (reproducible): a CTA owns
// 32 consecutive outputs; thread group q = 0..7 sums the splits s = q, q + 8, ... in order, the groups are added in order.
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float* __restrict__ part, const float* __restrict__ partb,
                                                            int splits, int O, int K, float* __restrict__ dW, int ldw,
                                                            int coloff, float* __restrict__ db) {
  __shared__ float sh[8][33];
  const int e = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int total = O * K;
  const int nw = ceil_div(total, 32);              // CTAs [0, nw) reduce dW, the rest db
  const bool bias = (int)blockIdx.x >= nw;
  const int idx = (bias ? (int)blockIdx.x - nw : (int)blockIdx.x) * 32 + e;
  const int n = bias ? O : total;
  const float* __restrict__ src = bias ? partb : part;
  float c0 = 0.f, c1 = 0.f;
  if (idx < n) {
    int i = grp;
    for (; i + 8 < splits; i += 16) {
      c0 += src[(int64_t)i * n + idx];
      c1 += src[(int64_t)(i + 8) * n + idx];
    }
    if (i < splits) c0 += src[(int64_t)i * n + idx];
  }
  sh[grp][e] = c0 + c1;
  __syncthreads();
  if (grp == 0 && idx < n) {
    float s = sh[0][e];
#pragma unroll
    for (int u = 1; u < 8; ++u) s += sh[u][e];
    if (bias) db[idx] += s;
    else dW[(int64_t)(idx / K) * ldw + coloff + idx % K] += s;
  }
}

// ------------------------------------------------------------------ elementwise
__global__ void __launch_bounds__(256) relu_mask_kernel(const float* gin, const float* __restrict__ y, float* gout, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gout[i] = y[i] > 0.f ? gin[i] : 0.f;
}

__global__ void __launch_bounds__(256) add_into_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

// ------------------------------------------------------------------ CSR of the edges by one endpoint, bins ordered by edge id
__global__ void __launch_bounds__(256) csr_count_kernel(const int64_t* __restrict__ key, int64_t E, int64_t N,
                                                         int32_t* __restrict__ cnt, int32_t* __restrict__ bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t k = key[e];
  if (k < 0 || k >= N) { *bad = 1; return; }
  atomicAdd(&cnt[k], 1);
}

__global__ void __launch_bounds__(1024) csr_scan_kernel(const int32_t* __restrict__ cnt, int64_t N, int32_t* __restrict__ ptr) {
  __shared__ int s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t chunk = ceil_div<int64_t>(N, blockDim.x);
  const int64_t b0 = min((int64_t)threadIdx.x * chunk, N), b1 = min(b0 + chunk, N);
  int sum = 0;
  for (int64_t b = b0; b < b1; ++b) sum += cnt[b];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(kFull, wi, o);
      if (lane >= o) wi += u;
    }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  int run = s_warp[warp] + incl - sum;
  for (int64_t b = b0; b < b1; ++b) {
    ptr[b] = run;
    run += cnt[b];
  }
  if (threadIdx.x == blockDim.x - 1) ptr[N] = run;   // exclusive prefix + own chunk = total
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const int64_t* __restrict__ key, int64_t E, int64_t N,
                                                        const int32_t* __restrict__ ptr, int32_t* __restrict__ cursor,
                                                        int32_t* __restrict__ perm) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t k = key[e];
  if (k < 0 || k >= N) return;
  perm[ptr[k] + atomicAdd(&cursor[k], 1)] = (int32_t)e;
}

__global__ void __launch_bounds__(256) csr_sort_kernel(const int32_t* __restrict__ ptr, int64_t N, int32_t* __restrict__ perm) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int32_t* s = perm + ptr[n];
  const int c = ptr[n + 1] - ptr[n];
  for (int i = 1; i < c; ++i) {
    const int32_t v = s[i];
    int j = i - 1;
    while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; }
    s[j + 1] = v;
  }
}

// ------------------------------------------------------------------ per-node reductions over the CSR bins (64 columns)
// S[n] = sum of V[e] over the bin of n, in edge order
__global__ void __launch_bounds__(256) seg_sum_kernel(const float* __restrict__ V, const int32_t* __restrict__ ptr,
                                                       const int32_t* __restrict__ perm, int64_t N, float* __restrict__ S) {
  const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 6);
  const int c = threadIdx.x & 63;
  if (n >= N) return;
  float s = 0.f;
  for (int i = ptr[n]; i < ptr[n + 1]; ++i) s += V[(int64_t)perm[i] * kD + c];
  S[n * kD + c] = s;
}

// scatter(m, dst, reduce = aggr) with empty bins = 0 (torch_scatter; layers.py:32-34 aggr of MessagePassing)
__global__ void __launch_bounds__(256) aggregate_fwd_kernel(const float* __restrict__ m, const int32_t* __restrict__ ptr,
                                                             const int32_t* __restrict__ perm, int64_t N, int aggr,
                                                             float* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 6);
  const int c = threadIdx.x & 63;
  if (n >= N) return;
  const int b = ptr[n], e = ptr[n + 1];
  float v = 0.f;
  if (e > b) {
    if (aggr == PGMP_AGGR_MAX) {
      v = m[(int64_t)perm[b] * kD + c];
      for (int i = b + 1; i < e; ++i) v = fmaxf(v, m[(int64_t)perm[i] * kD + c]);
    } else {
      for (int i = b; i < e; ++i) v += m[(int64_t)perm[i] * kD + c];
      if (aggr == PGMP_AGGR_MEAN) v /= (float)(e - b);
    }
  }
  out[n * kD + c] = v;
}

// gradient of the aggregation (max routes to the maxima of the bin, evenly among ties; oracle/mpn_train.py scatter)
// followed by the ReLU of the message MLP: dm is the gradient of its pre-activation
__global__ void __launch_bounds__(256) aggregate_bwd_kernel(const float* __restrict__ dagg, const float* __restrict__ m,
                                                             const float* __restrict__ agg, const int32_t* __restrict__ ptr,
                                                             const int32_t* __restrict__ perm, int64_t N, int aggr,
                                                             float* __restrict__ dm) {
  const int64_t n = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 6);
  const int c = threadIdx.x & 63;
  if (n >= N) return;
  const int b = ptr[n], e = ptr[n + 1];
  if (e <= b) return;
  const float g = dagg[n * kD + c];
  if (aggr == PGMP_AGGR_MAX) {
    const float mx = agg[n * kD + c];
    int ties = 0;
    for (int i = b; i < e; ++i) ties += m[(int64_t)perm[i] * kD + c] == mx;
    const float share = g / (float)max(ties, 1);
    for (int i = b; i < e; ++i) {
      const int64_t o = (int64_t)perm[i] * kD + c;
      const float v = m[o];
      dm[o] = (v == mx && v > 0.f) ? share : 0.f;      // ... and through the ReLU of the message
    }
  } else {
    const float share = aggr == PGMP_AGGR_MEAN ? g / (float)(e - b) : g;
    for (int i = b; i < e; ++i) {
      const int64_t o = (int64_t)perm[i] * kD + c;
      dm[o] = m[o] > 0.f ? share : 0.f;
    }
  }
}


// ------------------------------------------------------------------ per-type layer (TypeAwareMPNLayer, layers.py:157-274)
// Everything per edge is order-agnostic, so the per-type step runs in a TYPE-SORTED edge order (stable counting sort of
// the edges by the type of their source node): rows [type_ptr[t], type_ptr[t + 1]) use the message matrix of type t, only
// the edge logits and their gradients are (un)permuted.  Per-(target, type) bins (key = target * T + type, so that the bin
// index is the row of the [N][T * 64] update input) carry the attention softmax and the aggregation.
constexpr int kSortBlock = 1024;
constexpr int kMaxTypes = 17;

__global__ void __launch_bounds__(256) edge_type_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ node_types,
                                                         int64_t E, int64_t N, int T, int32_t* __restrict__ etype,
                                                         int32_t* __restrict__ bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t j = src[e];
  int t = 0;
  if (j < 0 || j >= N) { *bad = 1; } else {
    const int64_t v = node_types[j];
    if (v < 0 || v >= T) *bad = 1; else t = (int)v;
  }
  etype[e] = t;
}

// hist[t][block] = edges of type t in the block's 1024 edges
__global__ void __launch_bounds__(kSortBlock) type_hist_kernel(const int32_t* __restrict__ etype, int64_t E, int nblocks,
                                                                int32_t* __restrict__ hist) {
  __shared__ int s_cnt[kMaxTypes];
  if (threadIdx.x < kMaxTypes) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * kSortBlock + threadIdx.x;
  if (e < E) atomicAdd(&s_cnt[etype[e]], 1);
  __syncthreads();
  if (threadIdx.x < kMaxTypes) hist[threadIdx.x * nblocks + blockIdx.x] = s_cnt[threadIdx.x];
}

// stable scatter: position = (edges of smaller type) + (edges of the same type in earlier blocks: `base`, the exclusive scan
// of hist) + (edges of the same type earlier in this block)
__global__ void __launch_bounds__(kSortBlock) type_scatter_kernel(const int32_t* __restrict__ etype, int64_t E, int nblocks,
                                                                   const int32_t* __restrict__ base, int32_t* __restrict__ perm) {
  __shared__ int s_warp[32][kMaxTypes];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * kSortBlock + threadIdx.x;
  const int t = e < E ? etype[e] : -1;
  const unsigned same = __match_any_sync(kFull, t);
  const int rank = __popc(same & ((1u << lane) - 1u));
  for (int k = lane; k < kMaxTypes; k += 32) s_warp[warp][k] = 0;
  __syncwarp();
  if (t >= 0 && rank == 0) s_warp[warp][t] = __popc(same);
  __syncthreads();
  if (t < 0) return;
  int before = 0;
  for (int w2 = 0; w2 < warp; ++w2) before += s_warp[w2][t];
  perm[base[t * nblocks + blockIdx.x] + before + rank] = (int32_t)e;
}

__global__ void type_ptr_kernel(const int32_t* __restrict__ base, int nblocks, int32_t* __restrict__ type_ptr) {
  const int t = threadIdx.x;
  if (t <= kMaxTypes) type_ptr[t] = base[t * nblocks];      // base[kMaxTypes * nblocks] = E
}

// the graph in sorted order: endpoints, bin keys, edge attributes
__global__ void __launch_bounds__(256) sorted_graph_kernel(const int64_t* __restrict__ edge_index, const int32_t* __restrict__ etype,
                                                            const int32_t* __restrict__ perm, int64_t E, int T,
                                                            int64_t* __restrict__ s_src, int64_t* __restrict__ s_dst,
                                                            int64_t* __restrict__ s_key) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E) return;
  const int64_t e = perm[i];
  const int64_t d = edge_index[E + e];
  s_src[i] = edge_index[e];
  s_dst[i] = d;
  s_key[i] = d * T + etype[e];
}

// out[i][:] = in[perm[i]][:] (gather = 1) or out[perm[i]][:] = in[i][:] (gather = 0), rows of `width` floats
__global__ void __launch_bounds__(256) permute_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ perm,
                                                            int64_t E, int width, int gather, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= E * width) return;
  const int64_t i = idx / width;
  const int c = (int)(idx - i * width);
  const int64_t e = perm[i];
  if (gather) out[idx] = in[e * width + c]; else out[e * width + c] = in[idx];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// one warp per (target, type) bin: attention logit a = <g', wa[col]> + ba[col] (col = type or 0), softmax over the bin
// exp(a - max) / (sum + 1e-12) (torch_scatter's scatter_softmax, layers.py:242-249), U[bin] = sum alpha m.  The logits
// pass through `alpha` (written by lane 0, read by the warp after __syncwarp).
__global__ void __launch_bounds__(256) attn_fwd_kernel(const float* __restrict__ g, const float* __restrict__ m,
                                                        const float* __restrict__ wa, const float* __restrict__ ba, int per_type_col,
                                                        int T, const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm,
                                                        int64_t nbins, float* __restrict__ alpha, float* __restrict__ U) {
  const int64_t bin = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bin >= nbins) return;
  const int b = ptr[bin], e = ptr[bin + 1];
  float u0 = 0.f, u1 = 0.f;
  if (e > b) {
    const int col = per_type_col ? (int)(bin % T) : 0;
    const float w0 = wa[col * kD + lane], w1 = wa[col * kD + lane + 32], bias = ba[col];
    float mx = -INFINITY;
    for (int i = b; i < e; ++i) {
      const int64_t r = perm[i];
      const float a = warp_sum(fmaf(g[r * kD + lane], w0, g[r * kD + lane + 32] * w1)) + bias;
      if (lane == 0) alpha[r] = a;
      mx = fmaxf(mx, a);
    }
    __syncwarp();
    float se = 0.f;
    for (int i = b; i < e; ++i) se += expf(alpha[perm[i]] - mx);
    const float inv = 1.f / (se + 1e-12f);
    __syncwarp();
    for (int i = b; i < e; ++i) {
      const int64_t r = perm[i];
      const float al = expf(alpha[r] - mx) * inv;
      u0 = fmaf(al, m[r * kD + lane], u0);
      u1 = fmaf(al, m[r * kD + lane + 32], u1);
      __syncwarp();
      if (lane == 0) alpha[r] = al;
    }
  }
  U[bin * kD + lane] = u0;
  U[bin * kD + lane + 32] = u1;
}

// reverse: dm = alpha dU (through the ReLU of the message), d alpha = <m, dU>, da = alpha (d alpha - sum alpha d alpha);
// the logit's gradient goes on into g' (dg += da wa[col]) and is kept in `da` for the attn_net weight gradient
__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* __restrict__ dU, const float* __restrict__ m,
                                                        const float* __restrict__ alpha, const float* __restrict__ wa,
                                                        int per_type_col, int T, const int32_t* __restrict__ ptr,
                                                        const int32_t* __restrict__ perm, int64_t nbins,
                                                        float* __restrict__ dm, float* __restrict__ da, float* __restrict__ dg) {
  const int64_t bin = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (bin >= nbins) return;
  const int b = ptr[bin], e = ptr[bin + 1];
  if (e <= b) return;
  const int col = per_type_col ? (int)(bin % T) : 0;
  const float w0 = wa[col * kD + lane], w1 = wa[col * kD + lane + 32];
  const float d0 = dU[bin * kD + lane], d1 = dU[bin * kD + lane + 32];
  float s = 0.f;
  for (int i = b; i < e; ++i) {
    const int64_t r = perm[i];
    const float dal = warp_sum(fmaf(m[r * kD + lane], d0, m[r * kD + lane + 32] * d1));
    if (lane == 0) da[r] = dal;
    s = fmaf(alpha[r], dal, s);
  }
  __syncwarp();
  for (int i = b; i < e; ++i) {
    const int64_t r = perm[i];
    const float al = alpha[r];
    const float v = al * (da[r] - s);
    __syncwarp();
    if (lane == 0) da[r] = v;
    const float m0 = m[r * kD + lane], m1 = m[r * kD + lane + 32];
    dm[r * kD + lane] = m0 > 0.f ? al * d0 : 0.f;
    dm[r * kD + lane + 32] = m1 > 0.f ? al * d1 : 0.f;
    dg[r * kD + lane] += v * w0;
    dg[r * kD + lane + 32] += v * w1;
  }
}

// ------------------------------------------------------------------ BatchNorm1d, training mode
// column sums in fp64 over a row range: mode 0 (sum y, sum y^2); mode 1 (sum g, sum g * xhat)
__global__ void __launch_bounds__(256) bn_partial_kernel(const float* __restrict__ Y, const float* __restrict__ G,
                                                          const float* __restrict__ mean, const float* __restrict__ inv,
                                                          int64_t M, int C, int Cp, int64_t rows_per_cta,
                                                          double* __restrict__ part) {
  __shared__ double sh[2][256];
  const int c = threadIdx.x % Cp, lane = threadIdx.x / Cp, lanes = 256 / Cp;
  const int64_t rb = (int64_t)blockIdx.x * rows_per_cta, re = min(rb + rows_per_cta, M);
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    if (G == nullptr) {
      for (int64_t r = rb + lane; r < re; r += lanes) {
        const double y = Y[r * C + c];
        s1 += y;
        s2 += y * y;
      }
    } else {
      const float mu = mean[c], iv = inv[c];
      for (int64_t r = rb + lane; r < re; r += lanes) {
        const float gg = G[r * C + c];
        s1 += gg;
        s2 += (double)gg * (double)((Y[r * C + c] - mu) * iv);
      }
    }
  }
  sh[0][threadIdx.x] = s1;
  sh[1][threadIdx.x] = s2;
  __syncthreads();
  if (lane == 0 && c < C) {
    for (int l = 1; l < lanes; ++l) {
      s1 += sh[0][l * Cp + c];
      s2 += sh[1][l * Cp + c];
    }
    part[((int64_t)blockIdx.x * 2 + 0) * C + c] = s1;
    part[((int64_t)blockIdx.x * 2 + 1) * C + c] = s2;
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ part, int nparts, int C, int64_t M, float* __restrict__ mean,
                                   float* __restrict__ inv, float* __restrict__ rm, float* __restrict__ rv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < nparts; ++i) {
    s += part[((int64_t)i * 2 + 0) * C + c];
    q += part[((int64_t)i * 2 + 1) * C + c];
  }
  const double mu = s / (double)M;
  double var = q / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  inv[c] = (float)(1.0 / sqrt(var + (double)kBnEps));
  if (rm) rm[c] = (1.f - kBnMomentum) * rm[c] + kBnMomentum * (float)mu;
  if (rv) rv[c] = (1.f - kBnMomentum) * rv[c] + kBnMomentum * (float)(var * (double)M / (double)(M > 1 ? M - 1 : 1));
}

__global__ void bn_bwd_finalize_kernel(const double* __restrict__ part, int nparts, int C, int64_t M, float* __restrict__ m1,
                                       float* __restrict__ m2, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int i = 0; i < nparts; ++i) {
    s += part[((int64_t)i * 2 + 0) * C + c];
    q += part[((int64_t)i * 2 + 1) * C + c];
  }
  dbeta[c] += (float)s;
  dgamma[c] += (float)q;
  m1[c] = (float)(s / (double)M);
  m2[c] = (float)(q / (double)M);
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ Y, const float* __restrict__ mean,
                                                        const float* __restrict__ inv, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int64_t total, int C,
                                                        float* __restrict__ Z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  Z[i] = (Y[i] - mean[c]) * inv[c] * gamma[c] + beta[c];
}

// dY = inv * gamma * (g - mean(g) - xhat * mean(g * xhat)), then the ReLU in front of the BatchNorm
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ G, const float* __restrict__ Y,
                                                            const float* __restrict__ mean, const float* __restrict__ inv,
                                                            const float* __restrict__ gamma, const float* __restrict__ m1,
                                                            const float* __restrict__ m2, int64_t total, int C, int relu,
                                                            float* __restrict__ dY) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const float y = Y[i];
  const float xh = (y - mean[c]) * inv[c];
  const float d = inv[c] * gamma[c] * (G[i] - m1[c] - xh * m2[c]);
  dY[i] = (relu && !(y > 0.f)) ? 0.f : d;
}

// ------------------------------------------------------------------ workspace
struct MlpInst {
  float* y[PGMP_MAX_LAYERS];      // output of Linear (+ ReLU) l
  float* z[PGMP_MAX_LAYERS];      // BatchNorm output (bn[l])
  float* mean[PGMP_MAX_LAYERS];
  float* inv[PGMP_MAX_LAYERS];
  const float* out(const pgmp_mlp_train& m) const { return m.bn[m.n_layers - 1] ? z[m.n_layers - 1] : y[m.n_layers - 1]; }
  const float* act(const pgmp_mlp_train& m, int l) const { return m.bn[l] ? z[l] : y[l]; }
};

struct TrainWs {
  int32_t *dst_ptr, *dst_perm, *src_ptr, *src_perm, *cnt, *bad;
  MlpInst node_emb, edge_emb;
  MlpInst edge_head[kMaxOut], node_head[kMaxOut], class_head[kMaxOut];
  float *hid[kMaxSteps], *g[kMaxSteps], *m[kMaxSteps], *agg[kMaxSteps], *h[kMaxSteps];
  float *tab_p, *tab_q, *tab_r;
  float *dh, *dhp, *dh0, *tn1, *tn2, *tn3, *dg, *dg0, *be1, *be2, *ga, *gb;
  double* bn_part;
  float *bn_m1, *bn_m2, *part, *partb;
  // per-type layer
  int32_t *etype, *type_hist, *type_base, *type_perm, *type_ptr, *bin_ptr, *bin_perm;
  int64_t *s_src, *s_dst, *s_key;
  float *s_attr, *alpha[kMaxSteps], *U[kMaxSteps], *tab_rt, *dU, *dR, *da, *d_edge;
  uint64_t bytes;
};

inline int n_out_of(const pgmp_mpn_train_params& p) {
  const int first = p.steps - p.aux_loss_steps - 1 > 0 ? p.steps - p.aux_loss_steps - 1 : 0;   // NodeClassificationMPNSimple.py:81
  return p.steps - first;
}

MlpInst carve_mlp(Carver& c, const pgmp_mlp_train& m, uint64_t M, float* final_out) {
  MlpInst inst{};
  for (int l = 0; l < m.n_layers; ++l) {
    const uint64_t O = (uint64_t)m.dims[l + 1];
    const bool last = l == m.n_layers - 1;
    inst.y[l] = (last && final_out && !m.bn[l]) ? final_out : c.take<float>(M * O);
    if (m.bn[l]) {
      inst.z[l] = (last && final_out) ? final_out : c.take<float>(M * O);
      inst.mean[l] = c.take<float>(O);
      inst.inv[l] = c.take<float>(O);
    }
  }
  return inst;
}

TrainWs carve_train(const pgmp_mpn_train_params& p) {
  Carver c(p.workspace);
  TrainWs w{};
  const uint64_t N = (uint64_t)p.num_nodes, E = (uint64_t)p.num_edges;
  w.dst_ptr = c.take<int32_t>(N + 1);
  w.src_ptr = c.take<int32_t>(N + 1);
  w.dst_perm = c.take<int32_t>(E);
  w.src_perm = c.take<int32_t>(E);
  const uint64_t T = p.per_type ? (uint64_t)p.num_types : 1;
  w.cnt = c.take<int32_t>(N * T + 1);
  w.bad = c.take<int32_t>(1);
  w.node_emb = carve_mlp(c, p.node_emb, N, nullptr);
  w.edge_emb = carve_mlp(c, p.edge_emb, E, nullptr);
  const int n_out = n_out_of(p);
  for (int s = 0; s < p.steps && s < kMaxSteps; ++s) {
    w.hid[s] = c.take<float>(E * kD);
    w.g[s] = c.take<float>(E * kD);
    w.m[s] = c.take<float>(E * kD);
    w.agg[s] = c.take<float>(N * kD);
    w.h[s] = p.has_update_mlp ? c.take<float>(N * kD) : w.agg[s];
  }
  for (int k = 0; k < n_out && k < kMaxOut; ++k) {
    // per-type: the edge head runs in the sorted edge order, its output is un-permuted into edge_logits afterwards
    w.edge_head[k] = carve_mlp(c, p.edge_head, E, p.edge_logits && !p.per_type ? p.edge_logits + (uint64_t)k * E : nullptr);
    w.node_head[k] = carve_mlp(c, p.node_head, N, p.node_logits ? p.node_logits + (uint64_t)k * N : nullptr);
    w.class_head[k] = carve_mlp(c, p.class_head, N, p.class_logits ? p.class_logits + (uint64_t)k * N * p.num_classes : nullptr);
  }
  w.tab_p = c.take<float>(N * kD);
  w.tab_q = c.take<float>(N * kD);
  w.tab_r = c.take<float>(N * kD);
  w.dh = c.take<float>(N * kD);
  w.dhp = c.take<float>(N * kD);
  w.dh0 = c.take<float>(N * kD);
  w.tn1 = c.take<float>(N * kD);
  w.tn2 = c.take<float>(N * kD);
  w.tn3 = c.take<float>(N * kD);
  w.dg = c.take<float>(E * kD);
  w.dg0 = c.take<float>(E * kD);
  w.be1 = c.take<float>(E * kD);
  w.be2 = c.take<float>(E * kD);
  const uint64_t rows = N > E ? N : E;
  w.ga = c.take<float>(rows * kMaxWidth);
  w.gb = c.take<float>(rows * kMaxWidth);
  w.bn_part = c.take<double>((uint64_t)kMaxSplits * 2 * kMaxWidth);
  w.bn_m1 = c.take<float>(kMaxWidth);
  w.bn_m2 = c.take<float>(kMaxWidth);
  w.part = c.take<float>((uint64_t)kMaxSplits * kMaxWidth * kMaxWidth);
  w.partb = c.take<float>((uint64_t)kMaxSplits * kMaxWidth);
  if (p.per_type) {
    const uint64_t nb = ceil_div<uint64_t>(E > 0 ? E : 1, kSortBlock);
    w.etype = c.take<int32_t>(E);
    w.type_hist = c.take<int32_t>(kMaxTypes * nb + 1);
    w.type_base = c.take<int32_t>(kMaxTypes * nb + 1);
    w.type_perm = c.take<int32_t>(E);
    w.type_ptr = c.take<int32_t>(kMaxTypes + 1);
    w.bin_ptr = c.take<int32_t>(N * T + 1);
    w.bin_perm = c.take<int32_t>(E);
    w.s_src = c.take<int64_t>(E);
    w.s_dst = c.take<int64_t>(E);
    w.s_key = c.take<int64_t>(E);
    w.s_attr = c.take<float>(E * (uint64_t)p.edge_emb.dims[0]);
    for (int s = 0; s < p.steps && s < kMaxSteps; ++s) {
      w.alpha[s] = c.take<float>(p.attn ? E : 0);
      w.U[s] = c.take<float>(N * T * kD);
    }
    w.tab_rt = c.take<float>(N * T * kD);
    w.dU = c.take<float>(N * T * kD);
    w.dR = c.take<float>(N * T * kD);
    w.da = c.take<float>(E);
    w.d_edge = c.take<float>(E);
  }
  w.bytes = c.bytes();
  return w;
}

// ------------------------------------------------------------------ launch helpers
inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)ceil_div<int64_t>(n > 0 ? n : 1, per); }

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline bool vec_src(const ASrc& a) {
  for (int i = 0; i < a.n; ++i)
    if (!al16(a.s[i].p) || (a.s[i].ld & 3) || (a.s[i].w & 3)) return false;
  return true;
}
inline bool vec_w(const float* W, int ldw, int coloff, int K) { return al16(W) && !(ldw & 3) && !(coloff & 3) && !(K & 3); }

int launch_fwd(cudaStream_t st, const ASrc& a, int64_t M, const float* W, int ldw, int coloff, const float* bias, int O,
               int relu, float* Y, const float* add1 = nullptr, const int64_t* idx1 = nullptr, const float* add2 = nullptr,
               const int64_t* idx2 = nullptr, int ldy = 0, int add1_ld = kD) {
  if (M <= 0) return PGMP_OK;
  FwdArgs g{a, M, src_width(a), O, W, ldw, coloff, bias, add1, idx1, add2, idx2, relu, Y, ldy ? ldy : O, add1_ld};
  // PGMP_TRAIN_TC=1: the large 64-output products (the E-level ones) on the 5th-generation tensor cores, bf16x3 operand
  // split (mpn_train_tc.cu); default: the 3xTF32 mma.sync kernels below, which the fp32-level parity tests pin
  {
    const char* env = getenv("PGMP_TRAIN_TC");
    const bool blocks64 = a.s[0].w == kD && (a.n == 1 || a.s[1].w == kD);
    if (env && atoi(env) != 0 && O == kD && M >= 4096 && blocks64 && vec_src(a) && vec_w(W, ldw, coloff, g.K) && al16(Y) &&
        !(g.ldy & 3) && (!bias || al16(bias)) && (!add1 || (al16(add1) && !(add1_ld & 3))) && (!add2 || al16(add2))) {
      LinFwdTc t{a.s[0].p, a.s[0].ld, a.n > 1 ? a.s[1].p : nullptr, a.n > 1 ? a.s[1].ld : 0, M, W, ldw, coloff, bias,
                 add1, idx1, add1_ld, add2, idx2, relu, Y, g.ldy};
      return launch_lin_fwd_tc(st, t);
    }
  }
  const dim3 grid(blocks_for(M, BM), (unsigned)ceil_div(O, BN));
  if (vec_src(a) && vec_w(W, ldw, coloff, g.K))
    PGMP_LAUNCH(lin_fwd_kernel<true>, grid, 256, 0, st, g);
  else
    PGMP_LAUNCH(lin_fwd_kernel<false>, grid, 256, 0, st, g);
  return PGMP_OK;
}

int launch_bwd_in(cudaStream_t st, const float* dY, int O, int64_t M, const float* W, int ldw, int coloff, int K, Target t0,
                  const Target* t1 = nullptr, int ldd = 0) {
  if (M <= 0) return PGMP_OK;
  if (!ldd) ldd = O;
  BwdInArgs g{dY, ldd, M, O, K, W, ldw, coloff, t1 ? 2 : 1, {t0, t1 ? *t1 : Target{nullptr, 0, 0, 0, 0, nullptr}}};
  const dim3 grid(blocks_for(M, BM), (unsigned)ceil_div(K, BN));
  if (al16(dY) && !(ldd & 3) && !(O & 3) && vec_w(W, ldw, coloff, K))
    PGMP_LAUNCH(lin_bwd_in_kernel<true>, grid, 256, 0, st, g);
  else
    PGMP_LAUNCH(lin_bwd_in_kernel<false>, grid, 256, 0, st, g);
  return PGMP_OK;
}

// dW[:, coloff : coloff + K] += dY^T A (and db += column sums of dY when db != null)
int launch_bwd_w(cudaStream_t st, const TrainWs& w, const float* dY, int O, const ASrc& a, int64_t M, float* dW, int ldw,
                 int coloff, float* db, int ldd = 0) {
  if (M <= 0) return PGMP_OK;
  if (!ldd) ldd = O;
  const int K = src_width(a);
  int splits = (int)ceil_div<int64_t>(M, 256);
  if (splits > kMaxSplits) splits = kMaxSplits;
  const int64_t cap = (int64_t)kMaxSplits * kMaxWidth * kMaxWidth / ((int64_t)O * K);    // partial tiles that fit w.part
  if (splits > cap) splits = (int)(cap > 1 ? cap : 1);
  const int64_t rows = round_up<int64_t>(ceil_div<int64_t>(M, splits), BK);
  splits = (int)ceil_div<int64_t>(M, rows);
  BwdWArgs g{dY, ldd, O, a, K, M, rows, w.part, db ? w.partb : nullptr};
  const dim3 grid((unsigned)ceil_div(K, BN), (unsigned)ceil_div(O, BN), (unsigned)splits);
  if (al16(dY) && !(ldd & 3) && !(O & 3) && vec_src(a) && !(K & 3))
    PGMP_LAUNCH(lin_bwd_w_kernel<true>, grid, 256, 0, st, g);
  else
    PGMP_LAUNCH(lin_bwd_w_kernel<false>, grid, 256, 0, st, g);
  PGMP_LAUNCH(reduce_parts_kernel, (unsigned)(ceil_div(O * K, 32) + (db ? ceil_div(O, 32) : 0)), 256, 0, st, w.part,
              db ? w.partb : nullptr, splits, O, K, dW, ldw, coloff, db);
  return PGMP_OK;
}

int bn_partials(cudaStream_t st, const TrainWs& w, const float* Y, const float* G, const float* mean, const float* inv,
                int64_t M, int C, int* nparts) {
  int Cp = 32;
  while (Cp < C) Cp <<= 1;
  int n = (int)ceil_div<int64_t>(M, 256);
  if (n > kMaxSplits) n = kMaxSplits;
  if (n < 1) n = 1;
  const int64_t rows = ceil_div<int64_t>(M, n);
  n = (int)ceil_div<int64_t>(M, rows);
  PGMP_LAUNCH(bn_partial_kernel, (unsigned)n, 256, 0, st, Y, G, mean, inv, M, C, Cp, rows, w.bn_part);
  *nparts = n;
  return PGMP_OK;
}

#define PGMP_TRY(expr)                  \
  do {                                  \
    int _r = (expr);                    \
    if (_r != PGMP_OK) return _r;       \
  } while (0)

int mlp_forward(cudaStream_t st, const TrainWs& w, const pgmp_mlp_train& m, const MlpInst& inst, ASrc input, int64_t M,
                const float* params) {
  if (M <= 0) return PGMP_OK;
  ASrc cur = input;
  for (int l = 0; l < m.n_layers; ++l) {
    const int K = m.dims[l], O = m.dims[l + 1];
    PGMP_TRY(launch_fwd(st, cur, M, params + m.w[l], K, 0, params + m.b[l], O, m.relu[l], inst.y[l]));
    if (m.bn[l]) {
      int nparts = 0;
      PGMP_TRY(bn_partials(st, w, inst.y[l], nullptr, nullptr, nullptr, M, O, &nparts));
      PGMP_LAUNCH(bn_finalize_kernel, 1, kMaxWidth, 0, st, w.bn_part, nparts, O, M, inst.mean[l], inst.inv[l],
                  m.running_mean[l], m.running_var[l]);
      PGMP_LAUNCH(bn_apply_kernel, blocks_for(M * O), 256, 0, st, inst.y[l], inst.mean[l], inst.inv[l], params + m.gamma[l],
                  params + m.beta[l], M * O, O, inst.z[l]);
    }
    cur = src1(inst.act(m, l), O);
  }
  return PGMP_OK;
}

// reverse pass of one _make_mlp chain; the gradient of the input goes to `t0` / `t1` (nt = 0: not needed)
int mlp_backward(cudaStream_t st, const TrainWs& w, const pgmp_mlp_train& m, const MlpInst& inst, ASrc input, int64_t M,
                 const float* params, float* grads, const float* d_out, int nt, Target t0, Target t1) {
  if (M <= 0) return PGMP_OK;
  const float* g = d_out;
  float* bufs[2] = {w.ga, w.gb};
  int which = 0;
  for (int l = m.n_layers - 1; l >= 0; --l) {
    const int K = m.dims[l], O = m.dims[l + 1];
    if (m.bn[l]) {
      int nparts = 0;
      PGMP_TRY(bn_partials(st, w, inst.y[l], g, inst.mean[l], inst.inv[l], M, O, &nparts));
      PGMP_LAUNCH(bn_bwd_finalize_kernel, 1, kMaxWidth, 0, st, w.bn_part, nparts, O, M, w.bn_m1, w.bn_m2,
                  grads + m.gamma[l], grads + m.beta[l]);
      PGMP_LAUNCH(bn_bwd_apply_kernel, blocks_for(M * O), 256, 0, st, g, inst.y[l], inst.mean[l], inst.inv[l],
                  params + m.gamma[l], w.bn_m1, w.bn_m2, M * O, O, m.relu[l], bufs[which]);
      g = bufs[which];
      which ^= 1;
    } else if (m.relu[l]) {
      PGMP_LAUNCH(relu_mask_kernel, blocks_for(M * O), 256, 0, st, g, inst.y[l], bufs[which], M * O);
      g = bufs[which];
      which ^= 1;
    }
    const ASrc in = l == 0 ? input : src1(inst.act(m, l - 1), K);
    PGMP_TRY(launch_bwd_w(st, w, g, O, in, M, grads + m.w[l], K, 0, grads + m.b[l]));
    if (l > 0) {
      PGMP_TRY(launch_bwd_in(st, g, O, M, params + m.w[l], K, 0, K, Target{bufs[which], K, 0, K, 0}));
      g = bufs[which];
      which ^= 1;
    } else if (nt > 0) {
      PGMP_TRY(launch_bwd_in(st, g, O, M, params + m.w[l], K, 0, K, t0, nt > 1 ? &t1 : nullptr));
    }
  }
  return PGMP_OK;
}

int build_csr(cudaStream_t st, const TrainWs& w, const int64_t* key, int64_t E, int64_t N, int32_t* ptr, int32_t* perm) {
  PGMP_CUDA(cudaMemsetAsync(w.cnt, 0, sizeof(int32_t) * N, st));
  if (E > 0) PGMP_LAUNCH(csr_count_kernel, blocks_for(E), 256, 0, st, key, E, N, w.cnt, w.bad);
  PGMP_LAUNCH(csr_scan_kernel, 1, 1024, 0, st, w.cnt, N, ptr);
  PGMP_CUDA(cudaMemsetAsync(w.cnt, 0, sizeof(int32_t) * N, st));
  if (E > 0) {
    PGMP_LAUNCH(csr_fill_kernel, blocks_for(E), 256, 0, st, key, E, N, ptr, w.cnt, perm);
    PGMP_LAUNCH(csr_sort_kernel, blocks_for(N), 256, 0, st, ptr, N, perm);
  }
  return PGMP_OK;
}

int validate_mlp(const pgmp_mlp_train& m, const char* name, int in_dim, int out_dim) {
  if (m.n_layers < 1 || m.n_layers > PGMP_MAX_LAYERS) return set_error(PGMP_ERR_INVALID, "%s: %d layers", name, m.n_layers);
  for (int l = 0; l <= m.n_layers; ++l)
    if (m.dims[l] < 1 || m.dims[l] > kMaxWidth) return set_error(PGMP_ERR_INVALID, "%s: width %d outside [1, %d]", name, m.dims[l], kMaxWidth);
  if (in_dim > 0 && m.dims[0] != in_dim) return set_error(PGMP_ERR_INVALID, "%s: input width %d, expected %d", name, m.dims[0], in_dim);
  if (m.dims[m.n_layers] != out_dim) return set_error(PGMP_ERR_INVALID, "%s: output width %d, expected %d", name, m.dims[m.n_layers], out_dim);
  return PGMP_OK;
}

int validate_train(const pgmp_mpn_train_params* p, bool backward) {
  if (!p) return set_error(PGMP_ERR_INVALID, "null params");
  if (p->num_nodes <= 0 || p->num_edges < 0) return set_error(PGMP_ERR_INVALID, "bad graph size N=%lld E=%lld", (long long)p->num_nodes, (long long)p->num_edges);
  if (p->num_nodes > (1ll << 30) || p->num_edges > (1ll << 30)) return set_error(PGMP_ERR_INVALID, "graph too large for 32-bit bins");
  if (p->dim != kD) return set_error(PGMP_ERR_INVALID, "NODE/EDGE_FEATURE_DIM must be 64, got %d", p->dim);
  if (p->steps < 1 || p->steps > kMaxSteps) return set_error(PGMP_ERR_INVALID, "STEPS %d outside [1, %d]", p->steps, kMaxSteps);
  if (p->aux_loss_steps < 0 || n_out_of(*p) > kMaxOut) return set_error(PGMP_ERR_INVALID, "AUX_LOSS_STEPS %d", p->aux_loss_steps);
  if (p->aggr < 0 || p->aggr > 2) return set_error(PGMP_ERR_INVALID, "aggr %d", p->aggr);
  if (p->num_classes < 1 || p->num_classes > kMaxWidth) return set_error(PGMP_ERR_INVALID, "num_classes %d", p->num_classes);
  PGMP_TRY(validate_mlp(p->node_emb, "node_embedding", 0, kD));
  PGMP_TRY(validate_mlp(p->edge_emb, "edge_embedding", 0, kD));
  PGMP_TRY(validate_mlp(p->edge_head, "edge_classification", kD, 1));
  PGMP_TRY(validate_mlp(p->node_head, "node_classification", kD, 1));
  PGMP_TRY(validate_mlp(p->class_head, "classification", kD, p->num_classes));
  if (!p->x || !p->params || !p->node_logits || !p->class_logits || !p->workspace) return set_error(PGMP_ERR_INVALID, "null device pointer");
  if (p->num_edges > 0 && (!p->edge_attr || !p->edge_index || !p->edge_logits)) return set_error(PGMP_ERR_INVALID, "null edge pointer");
  if (p->per_type) {
    if (p->num_types < 1 || p->num_types > kMaxTypes) return set_error(PGMP_ERR_INVALID, "per_type: num_types %d", p->num_types);
    if (!p->has_update_mlp) return set_error(PGMP_ERR_INVALID, "per_type needs update_mlp (layers.py:253-258)");
    if (p->attn < 0 || p->attn > 2) return set_error(PGMP_ERR_INVALID, "attn %d", p->attn);
    if (!p->node_types) return set_error(PGMP_ERR_INVALID, "per_type without node_types");
    if ((int64_t)p->num_nodes * p->num_types > (1ll << 30)) return set_error(PGMP_ERR_INVALID, "too many (node, type) bins");
  } else if (p->attn) {
    return set_error(PGMP_ERR_INVALID, "attention aggregation needs per_type");
  }
  if (backward) {
    if (!p->grads || !p->d_node_logits || !p->d_class_logits || (p->num_edges > 0 && !p->d_edge_logits))
      return set_error(PGMP_ERR_INVALID, "backward: null gradient pointer");
  }
  return PGMP_OK;
}


// ================================================================== per-type layer: forward / backward
// sizes of the 17 source-type groups on the host (one stream wait): the per-type products are launched per group
int read_type_ptr(const TrainWs& w, cudaStream_t st, int32_t (&tp)[kMaxTypes + 1]) {
  PGMP_CUDA(cudaMemcpyAsync(tp, w.type_ptr, sizeof(tp), cudaMemcpyDeviceToHost, st));
  PGMP_CUDA(cudaStreamSynchronize(st));
  return PGMP_OK;
}

int mpn_train_forward_pt(const pgmp_mpn_train_params& p, const TrainWs& w, cudaStream_t st) {
  const int64_t N = p.num_nodes, E = p.num_edges;
  const float* P = p.params;
  const int T = p.num_types;
  const int nd = p.skip ? 2 * kD : kD, ed = nd;
  const int ld1 = 2 * nd + ed, ldm = nd + kD, ldu = T * kD;
  const int64_t bins = N * T;
  const int ain = p.edge_emb.dims[0];

  PGMP_CUDA(cudaMemsetAsync(w.bad, 0, sizeof(int32_t), st));
  int32_t tp[kMaxTypes + 1] = {};
  if (E > 0) {
    const int nb = (int)ceil_div<int64_t>(E, kSortBlock);
    PGMP_LAUNCH(edge_type_kernel, blocks_for(E), 256, 0, st, p.edge_index, p.node_types, E, N, T, w.etype, w.bad);
    PGMP_LAUNCH(type_hist_kernel, (unsigned)nb, kSortBlock, 0, st, w.etype, E, nb, w.type_hist);
    PGMP_LAUNCH(csr_scan_kernel, 1, 1024, 0, st, w.type_hist, (int64_t)kMaxTypes * nb, w.type_base);
    PGMP_LAUNCH(type_scatter_kernel, (unsigned)nb, kSortBlock, 0, st, w.etype, E, nb, w.type_base, w.type_perm);
    PGMP_LAUNCH(type_ptr_kernel, 1, 32, 0, st, w.type_base, nb, w.type_ptr);
    PGMP_LAUNCH(sorted_graph_kernel, blocks_for(E), 256, 0, st, p.edge_index, w.etype, w.type_perm, E, T, w.s_src, w.s_dst, w.s_key);
    PGMP_LAUNCH(permute_rows_kernel, blocks_for(E * ain), 256, 0, st, p.edge_attr, w.type_perm, E, ain, 1, w.s_attr);
    PGMP_TRY(read_type_ptr(w, st, tp));
  }
  const int64_t* src = w.s_src;
  const int64_t* dst = w.s_dst;
  PGMP_TRY(build_csr(st, w, dst, E, N, w.dst_ptr, w.dst_perm));
  PGMP_TRY(build_csr(st, w, src, E, N, w.src_ptr, w.src_perm));
  PGMP_TRY(build_csr(st, w, w.s_key, E, bins, w.bin_ptr, w.bin_perm));

  PGMP_TRY(mlp_forward(st, w, p.node_emb, w.node_emb, src1(p.x, p.node_emb.dims[0]), N, P));
  PGMP_TRY(mlp_forward(st, w, p.edge_emb, w.edge_emb, src1(w.s_attr, ain), E, P));
  const float* h0 = w.node_emb.out(p.node_emb);
  const float* g0 = w.edge_emb.out(p.edge_emb);
  const int first = p.steps - n_out_of(p);
  for (int s = 0; s < p.steps; ++s) {
    const float* hp = s ? w.h[s - 1] : h0;
    const float* gp = s ? w.g[s - 1] : g0;
    const ASrc xs = p.skip ? src2(h0, hp, kD) : src1(hp, kD);
    const ASrc es = p.skip ? src2(g0, gp, kD) : src1(gp, kD);
    // mlp_edge (layers.py:171-175, 214)
    PGMP_TRY(launch_fwd(st, xs, N, P + p.w1, ld1, 0, nullptr, kD, 0, w.tab_p));
    PGMP_TRY(launch_fwd(st, xs, N, P + p.w1, ld1, nd, nullptr, kD, 0, w.tab_q));
    PGMP_TRY(launch_fwd(st, es, E, P + p.w1, ld1, 2 * nd, P + p.b1, kD, 1, w.hid[s], w.tab_p, dst, w.tab_q, src));
    PGMP_TRY(launch_fwd(st, src1(w.hid[s], kD), E, P + p.w2, kD, 0, P + p.b2, kD, 1, w.g[s]));
    // per-type message (layers.py:222-224, 264-274): m = ReLU(Wm[t] [x_i ; g'] + bm[t]), t = type of the source
    for (int t = 0; t < kMaxTypes; ++t) {
      const int64_t a = tp[t], cnt = tp[t + 1] - tp[t];
      if (cnt <= 0) continue;
      const float* Wt = P + p.wm + (int64_t)t * p.wm_type_stride;
      const float* bt = P + p.bm + (int64_t)t * p.wm_type_stride;
      PGMP_TRY(launch_fwd(st, xs, N, Wt, ldm, 0, bt, kD, 0, w.tab_rt + t * kD, nullptr, nullptr, nullptr, nullptr, ldu));
      PGMP_TRY(launch_fwd(st, src1(w.g[s] + a * kD, kD), cnt, Wt, ldm, nd, nullptr, kD, 1, w.m[s] + a * kD, w.tab_rt + t * kD,
                          dst + a, nullptr, nullptr, 0, ldu));
    }
    // aggregation per (target, type) (layers.py:234-251) and the update MLP over the concatenated types (:253-258)
    if (p.attn)
      PGMP_LAUNCH(attn_fwd_kernel, blocks_for(bins, 8), 256, 0, st, w.g[s], w.m[s], P + p.wa, P + p.ba,
                  p.attn == PGMP_ATTN_PER_TYPE ? 1 : 0, T, w.bin_ptr, w.bin_perm, bins, w.alpha[s], w.U[s]);
    else
      PGMP_LAUNCH(aggregate_fwd_kernel, blocks_for(bins, 4), 256, 0, st, w.m[s], w.bin_ptr, w.bin_perm, bins, p.aggr, w.U[s]);
    PGMP_TRY(launch_fwd(st, src1(w.U[s], ldu), N, P + p.wu, ldu, 0, P + p.bu, kD, 1, w.h[s]));
    if (s >= first) {
      const int k = s - first;
      PGMP_TRY(mlp_forward(st, w, p.node_head, w.node_head[k], src1(w.h[s], kD), N, P));
      PGMP_TRY(mlp_forward(st, w, p.class_head, w.class_head[k], src1(w.h[s], kD), N, P));
      PGMP_TRY(mlp_forward(st, w, p.edge_head, w.edge_head[k], src1(w.g[s], kD), E, P));
      if (E > 0)
        PGMP_LAUNCH(permute_rows_kernel, blocks_for(E), 256, 0, st, w.edge_head[k].out(p.edge_head), w.type_perm, E, 1, 0,
                    p.edge_logits + (int64_t)k * E);
    }
  }
  return PGMP_OK;
}

int mpn_train_backward_pt(const pgmp_mpn_train_params& p, const TrainWs& w, cudaStream_t st) {
  const int64_t N = p.num_nodes, E = p.num_edges;
  const float* P = p.params;
  float* G = p.grads;
  const int T = p.num_types;
  const int nd = p.skip ? 2 * kD : kD, ed = nd;
  const int ld1 = 2 * nd + ed, ldm = nd + kD, ldu = T * kD;
  const int64_t bins = N * T;
  const int J = p.num_classes;
  const float* h0 = w.node_emb.out(p.node_emb);
  const float* g0 = w.edge_emb.out(p.edge_emb);
  const int first = p.steps - n_out_of(p);
  const int64_t nN = N * kD, nE = E * kD;
  const int64_t* src = w.s_src;
  const int64_t* dst = w.s_dst;
  int32_t tp[kMaxTypes + 1] = {};
  if (E > 0) PGMP_TRY(read_type_ptr(w, st, tp));

  float* dh = w.dh;
  float* dhp = w.dhp;
  PGMP_CUDA(cudaMemsetAsync(dh, 0, sizeof(float) * nN, st));
  PGMP_CUDA(cudaMemsetAsync(w.dh0, 0, sizeof(float) * nN, st));
  if (E > 0) {
    PGMP_CUDA(cudaMemsetAsync(w.dg, 0, sizeof(float) * nE, st));
    PGMP_CUDA(cudaMemsetAsync(w.dg0, 0, sizeof(float) * nE, st));
  }
  const Target none{nullptr, 0, 0, 0, 0, nullptr};
  for (int s = p.steps - 1; s >= 0; --s) {
    const float* hp = s ? w.h[s - 1] : h0;
    const float* gp = s ? w.g[s - 1] : g0;
    const ASrc xs = p.skip ? src2(h0, hp, kD) : src1(hp, kD);
    const ASrc es = p.skip ? src2(g0, gp, kD) : src1(gp, kD);
    if (s >= first) {
      const int k = s - first;
      if (E > 0)
        PGMP_LAUNCH(permute_rows_kernel, blocks_for(E), 256, 0, st, p.d_edge_logits + (int64_t)k * E, w.type_perm, E, 1, 1, w.d_edge);
      PGMP_TRY(mlp_backward(st, w, p.edge_head, w.edge_head[k], src1(w.g[s], kD), E, P, G, w.d_edge, 1,
                            Target{w.dg, kD, 0, kD, 1}, none));
      PGMP_TRY(mlp_backward(st, w, p.node_head, w.node_head[k], src1(w.h[s], kD), N, P, G, p.d_node_logits + (int64_t)k * N, 1,
                            Target{dh, kD, 0, kD, 1}, none));
      PGMP_TRY(mlp_backward(st, w, p.class_head, w.class_head[k], src1(w.h[s], kD), N, P, G,
                            p.d_class_logits + (int64_t)k * N * J, 1, Target{dh, kD, 0, kD, 1}, none));
    }
    // update_mlp over [N][T * 64]
    PGMP_LAUNCH(relu_mask_kernel, blocks_for(nN), 256, 0, st, dh, w.h[s], w.tn1, nN);
    PGMP_TRY(launch_bwd_w(st, w, w.tn1, kD, src1(w.U[s], ldu), N, G + p.wu, ldu, 0, G + p.bu));
    PGMP_TRY(launch_bwd_in(st, w.tn1, kD, N, P + p.wu, ldu, 0, ldu, Target{w.dU, ldu, 0, ldu, 0}));
    PGMP_CUDA(cudaMemsetAsync(dhp, 0, sizeof(float) * nN, st));
    const Target tx0{p.skip ? w.dh0 : dhp, kD, 0, kD, 1};
    const Target tx1{dhp, kD, kD, kD, 1};
    if (E > 0) {
      float* dm = w.be1;
      if (p.attn) {
        // softmax-weighted sum: dm, the logit gradients da (and their share of dg), then the attn_net weights
        PGMP_LAUNCH(attn_bwd_kernel, blocks_for(bins, 8), 256, 0, st, w.dU, w.m[s], w.alpha[s], P + p.wa,
                    p.attn == PGMP_ATTN_PER_TYPE ? 1 : 0, T, w.bin_ptr, w.bin_perm, bins, dm, w.da, w.dg);
        if (p.attn == PGMP_ATTN_PER_TYPE) {
          for (int t = 0; t < T; ++t) {
            const int64_t a = tp[t], cnt = tp[t + 1] - tp[t];
            if (cnt > 0)
              PGMP_TRY(launch_bwd_w(st, w, w.da + a, 1, src1(w.g[s] + a * kD, kD), cnt, G + p.wa + (int64_t)t * kD, kD, 0, G + p.ba + t));
          }
        } else {
          PGMP_TRY(launch_bwd_w(st, w, w.da, 1, src1(w.g[s], kD), E, G + p.wa, kD, 0, G + p.ba));
        }
      } else {
        PGMP_LAUNCH(aggregate_bwd_kernel, blocks_for(bins, 4), 256, 0, st, w.dU, w.m[s], w.U[s], w.bin_ptr, w.bin_perm, bins, p.aggr, dm);
      }
      // per-type message MLP: edge columns per edge, node columns through the per-(target, type) sums dR
      PGMP_LAUNCH(seg_sum_kernel, blocks_for(bins, 4), 256, 0, st, dm, w.bin_ptr, w.bin_perm, bins, w.dR);
      for (int t = 0; t < kMaxTypes; ++t) {
        const int64_t a = tp[t], cnt = tp[t + 1] - tp[t];
        if (cnt <= 0) continue;
        const float* Wt = P + p.wm + (int64_t)t * p.wm_type_stride;
        float* dWt = G + p.wm + (int64_t)t * p.wm_type_stride;
        float* dbt = G + p.bm + (int64_t)t * p.wm_type_stride;
        PGMP_TRY(launch_bwd_w(st, w, dm + a * kD, kD, src1(w.g[s] + a * kD, kD), cnt, dWt, ldm, nd, nullptr));
        PGMP_TRY(launch_bwd_w(st, w, w.dR + t * kD, kD, xs, N, dWt, ldm, 0, dbt, ldu));
        PGMP_TRY(launch_bwd_in(st, w.dR + t * kD, kD, N, Wt, ldm, 0, nd, tx0, p.skip ? &tx1 : nullptr, ldu));
        // the gradient of g_s is complete with this term: the epilogue also applies the ReLU of mlp_edge.2
        PGMP_TRY(launch_bwd_in(st, dm + a * kD, kD, cnt, Wt, ldm, nd, kD, Target{w.dg + a * kD, kD, 0, kD, 1, w.g[s] + a * kD}));
      }
      // mlp_edge.2
      PGMP_TRY(launch_bwd_w(st, w, w.dg, kD, src1(w.hid[s], kD), E, G + p.w2, kD, 0, G + p.b2));
      float* dhid = w.be2;
      PGMP_TRY(launch_bwd_in(st, w.dg, kD, E, P + p.w2, kD, 0, kD, Target{dhid, kD, 0, kD, 0, w.hid[s]}));
      // mlp_edge.0
      PGMP_LAUNCH(seg_sum_kernel, blocks_for(N, 4), 256, 0, st, dhid, w.dst_ptr, w.dst_perm, N, w.tn1);
      PGMP_LAUNCH(seg_sum_kernel, blocks_for(N, 4), 256, 0, st, dhid, w.src_ptr, w.src_perm, N, w.tn2);
      PGMP_TRY(launch_bwd_w(st, w, dhid, kD, es, E, G + p.w1, ld1, 2 * nd, nullptr));
      PGMP_TRY(launch_bwd_w(st, w, w.tn1, kD, xs, N, G + p.w1, ld1, 0, G + p.b1));
      PGMP_TRY(launch_bwd_w(st, w, w.tn2, kD, xs, N, G + p.w1, ld1, nd, nullptr));
      PGMP_TRY(launch_bwd_in(st, w.tn1, kD, N, P + p.w1, ld1, 0, nd, tx0, p.skip ? &tx1 : nullptr));
      PGMP_TRY(launch_bwd_in(st, w.tn2, kD, N, P + p.w1, ld1, nd, nd, tx0, p.skip ? &tx1 : nullptr));
      const Target te0{p.skip ? w.dg0 : w.dg, kD, 0, kD, p.skip ? 1 : 0};
      const Target te1{w.dg, kD, kD, kD, 0};
      PGMP_TRY(launch_bwd_in(st, dhid, kD, E, P + p.w1, ld1, 2 * nd, ed, te0, p.skip ? &te1 : nullptr));
    }
    float* tmp = dh;
    dh = dhp;
    dhp = tmp;
  }
  PGMP_LAUNCH(add_into_kernel, blocks_for(nN), 256, 0, st, w.dh0, dh, nN);
  if (E > 0) PGMP_LAUNCH(add_into_kernel, blocks_for(nE), 256, 0, st, w.dg0, w.dg, nE);
  const int xin = p.node_emb.dims[0];
  PGMP_TRY(mlp_backward(st, w, p.node_emb, w.node_emb, src1(p.x, xin), N, P, G, w.dh0, p.grad_x ? 1 : 0,
                        Target{p.grad_x, xin, 0, xin, 0}, none));
  PGMP_TRY(mlp_backward(st, w, p.edge_emb, w.edge_emb, src1(w.s_attr, p.edge_emb.dims[0]), E, P, G, w.dg0, 0, none, none));
  return PGMP_OK;
}

}  // namespace

// ================================================================== forward
int mpn_train_forward(const pgmp_mpn_train_params& p, cudaStream_t st) {
  const TrainWs w = carve_train(p);
  if (w.bytes > p.workspace_bytes)
    return set_error(PGMP_ERR_INVALID, "workspace too small: %llu < %llu", (unsigned long long)p.workspace_bytes, (unsigned long long)w.bytes);
  if (p.per_type) return mpn_train_forward_pt(p, w, st);
  const int64_t N = p.num_nodes, E = p.num_edges;
  const float* P = p.params;
  const int64_t* src = p.edge_index;
  const int64_t* dst = p.edge_index + E;
  const int nd = p.skip ? 2 * kD : kD, ed = nd;
  const int ld1 = 2 * nd + ed, ldm = nd + kD;

  PGMP_CUDA(cudaMemsetAsync(w.bad, 0, sizeof(int32_t), st));
  PGMP_TRY(build_csr(st, w, dst, E, N, w.dst_ptr, w.dst_perm));
  PGMP_TRY(build_csr(st, w, src, E, N, w.src_ptr, w.src_perm));

  PGMP_TRY(mlp_forward(st, w, p.node_emb, w.node_emb, src1(p.x, p.node_emb.dims[0]), N, P));          // :65
  PGMP_TRY(mlp_forward(st, w, p.edge_emb, w.edge_emb, src1(p.edge_attr, p.edge_emb.dims[0]), E, P));   // :66
  const float* h0 = w.node_emb.out(p.node_emb);
  const float* g0 = w.edge_emb.out(p.edge_emb);
  const int first = p.steps - n_out_of(p);
  for (int s = 0; s < p.steps; ++s) {
    const float* hp = s ? w.h[s - 1] : h0;
    const float* gp = s ? w.g[s - 1] : g0;
    const ASrc xs = p.skip ? src2(h0, hp, kD) : src1(hp, kD);      // :76-78
    const ASrc es = p.skip ? src2(g0, gp, kD) : src1(gp, kD);
    // mlp_edge (layers.py:66): hidden = ReLU(W1 [x_i ; x_j ; e] + b1), g' = ReLU(W2 hidden + b2)
    PGMP_TRY(launch_fwd(st, xs, N, P + p.w1, ld1, 0, nullptr, kD, 0, w.tab_p));
    PGMP_TRY(launch_fwd(st, xs, N, P + p.w1, ld1, nd, nullptr, kD, 0, w.tab_q));
    PGMP_TRY(launch_fwd(st, es, E, P + p.w1, ld1, 2 * nd, P + p.b1, kD, 1, w.hid[s], w.tab_p, dst, w.tab_q, src));
    PGMP_TRY(launch_fwd(st, src1(w.hid[s], kD), E, P + p.w2, kD, 0, P + p.b2, kD, 1, w.g[s]));
    // message (layers.py:74-77): m = ReLU(Wm [x_i ; g'] + bm); aggregate over the target
    PGMP_TRY(launch_fwd(st, xs, N, P + p.wm, ldm, 0, P + p.bm, kD, 0, w.tab_r));
    PGMP_TRY(launch_fwd(st, src1(w.g[s], kD), E, P + p.wm, ldm, nd, nullptr, kD, 1, w.m[s], w.tab_r, dst));
    PGMP_LAUNCH(aggregate_fwd_kernel, blocks_for(N, 4), 256, 0, st, w.m[s], w.dst_ptr, w.dst_perm, N, p.aggr, w.agg[s]);
    if (p.has_update_mlp)                                            // layers.py:79-82
      PGMP_TRY(launch_fwd(st, src1(w.agg[s], kD), N, P + p.wu, kD, 0, P + p.bu, kD, 1, w.h[s]));
    if (s >= first) {                                                // :81-84
      const int k = s - first;
      PGMP_TRY(mlp_forward(st, w, p.node_head, w.node_head[k], src1(w.h[s], kD), N, P));
      PGMP_TRY(mlp_forward(st, w, p.class_head, w.class_head[k], src1(w.h[s], kD), N, P));
      PGMP_TRY(mlp_forward(st, w, p.edge_head, w.edge_head[k], src1(w.g[s], kD), E, P));
    }
  }
  return PGMP_OK;
}

// ================================================================== backward
int mpn_train_backward(const pgmp_mpn_train_params& p, cudaStream_t st) {
  const TrainWs w = carve_train(p);
  if (w.bytes > p.workspace_bytes)
    return set_error(PGMP_ERR_INVALID, "workspace too small: %llu < %llu", (unsigned long long)p.workspace_bytes, (unsigned long long)w.bytes);
  if (p.per_type) return mpn_train_backward_pt(p, w, st);
  const int64_t N = p.num_nodes, E = p.num_edges;
  const float* P = p.params;
  float* G = p.grads;
  const int nd = p.skip ? 2 * kD : kD, ed = nd;
  const int ld1 = 2 * nd + ed, ldm = nd + kD;
  const int J = p.num_classes;
  const float* h0 = w.node_emb.out(p.node_emb);
  const float* g0 = w.edge_emb.out(p.edge_emb);
  const int first = p.steps - n_out_of(p);
  const int64_t nN = N * kD, nE = E * kD;

  float* dh = w.dh;     // gradient of h_s (complete once the heads of step s have been added)
  float* dhp = w.dhp;   // gradient of h_{s-1} collected during step s
  PGMP_CUDA(cudaMemsetAsync(dh, 0, sizeof(float) * nN, st));
  PGMP_CUDA(cudaMemsetAsync(w.dh0, 0, sizeof(float) * nN, st));
  if (E > 0) {
    PGMP_CUDA(cudaMemsetAsync(w.dg, 0, sizeof(float) * nE, st));
    PGMP_CUDA(cudaMemsetAsync(w.dg0, 0, sizeof(float) * nE, st));
  }
  const Target none{nullptr, 0, 0, 0, 0, nullptr};
  for (int s = p.steps - 1; s >= 0; --s) {
    const float* hp = s ? w.h[s - 1] : h0;
    const float* gp = s ? w.g[s - 1] : g0;
    const ASrc xs = p.skip ? src2(h0, hp, kD) : src1(hp, kD);
    const ASrc es = p.skip ? src2(g0, gp, kD) : src1(gp, kD);
    if (s >= first) {
      const int k = s - first;
      PGMP_TRY(mlp_backward(st, w, p.edge_head, w.edge_head[k], src1(w.g[s], kD), E, P, G, p.d_edge_logits + (int64_t)k * E, 1,
                            Target{w.dg, kD, 0, kD, 1}, none));
      PGMP_TRY(mlp_backward(st, w, p.node_head, w.node_head[k], src1(w.h[s], kD), N, P, G, p.d_node_logits + (int64_t)k * N, 1,
                            Target{dh, kD, 0, kD, 1}, none));
      PGMP_TRY(mlp_backward(st, w, p.class_head, w.class_head[k], src1(w.h[s], kD), N, P, G,
                            p.d_class_logits + (int64_t)k * N * J, 1, Target{dh, kD, 0, kD, 1}, none));
    }
    // update_mlp
    const float* dagg = dh;
    if (p.has_update_mlp) {
      PGMP_LAUNCH(relu_mask_kernel, blocks_for(nN), 256, 0, st, dh, w.h[s], w.tn1, nN);
      PGMP_TRY(launch_bwd_w(st, w, w.tn1, kD, src1(w.agg[s], kD), N, G + p.wu, kD, 0, G + p.bu));
      PGMP_TRY(launch_bwd_in(st, w.tn1, kD, N, P + p.wu, kD, 0, kD, Target{w.tn2, kD, 0, kD, 0}));
      dagg = w.tn2;
    }
    PGMP_CUDA(cudaMemsetAsync(dhp, 0, sizeof(float) * nN, st));
    const Target tx0{p.skip ? w.dh0 : dhp, kD, 0, kD, 1};        // columns of x = [h0 ; h_{s-1}] (or just h_{s-1})
    const Target tx1{dhp, kD, kD, kD, 1};
    if (E > 0) {
      // aggregation and the ReLU of the message
      float* dm = w.be1;
      PGMP_LAUNCH(aggregate_bwd_kernel, blocks_for(N, 4), 256, 0, st, dagg, w.m[s], w.agg[s], w.dst_ptr, w.dst_perm, N, p.aggr, dm);
      // mlp_node: edge columns per edge, node columns through the per-target sums
      PGMP_LAUNCH(seg_sum_kernel, blocks_for(N, 4), 256, 0, st, dm, w.dst_ptr, w.dst_perm, N, w.tn3);
      PGMP_TRY(launch_bwd_w(st, w, dm, kD, src1(w.g[s], kD), E, G + p.wm, ldm, nd, nullptr));
      PGMP_TRY(launch_bwd_w(st, w, w.tn3, kD, xs, N, G + p.wm, ldm, 0, G + p.bm));
      PGMP_TRY(launch_bwd_in(st, w.tn3, kD, N, P + p.wm, ldm, 0, nd, tx0, p.skip ? &tx1 : nullptr));
      // ... the gradient of g_s is complete with this term: the epilogue also applies the ReLU of mlp_edge.2
      PGMP_TRY(launch_bwd_in(st, dm, kD, E, P + p.wm, ldm, nd, kD, Target{w.dg, kD, 0, kD, 1, w.g[s]}));
      // mlp_edge.2
      PGMP_TRY(launch_bwd_w(st, w, w.dg, kD, src1(w.hid[s], kD), E, G + p.w2, kD, 0, G + p.b2));
      float* dhid = w.be2;
      PGMP_TRY(launch_bwd_in(st, w.dg, kD, E, P + p.w2, kD, 0, kD, Target{dhid, kD, 0, kD, 0, w.hid[s]}));   // + ReLU of mlp_edge.0
      // mlp_edge.0
      PGMP_LAUNCH(seg_sum_kernel, blocks_for(N, 4), 256, 0, st, dhid, w.dst_ptr, w.dst_perm, N, w.tn1);
      PGMP_LAUNCH(seg_sum_kernel, blocks_for(N, 4), 256, 0, st, dhid, w.src_ptr, w.src_perm, N, w.tn2);
      PGMP_TRY(launch_bwd_w(st, w, dhid, kD, es, E, G + p.w1, ld1, 2 * nd, nullptr));
      PGMP_TRY(launch_bwd_w(st, w, w.tn1, kD, xs, N, G + p.w1, ld1, 0, G + p.b1));
      PGMP_TRY(launch_bwd_w(st, w, w.tn2, kD, xs, N, G + p.w1, ld1, nd, nullptr));
      PGMP_TRY(launch_bwd_in(st, w.tn1, kD, N, P + p.w1, ld1, 0, nd, tx0, p.skip ? &tx1 : nullptr));
      PGMP_TRY(launch_bwd_in(st, w.tn2, kD, N, P + p.w1, ld1, nd, nd, tx0, p.skip ? &tx1 : nullptr));
      // e = [g0 ; g_{s-1}]: the gradient of g_{s-1} replaces dg (dg of step s is consumed)
      const Target te0{p.skip ? w.dg0 : w.dg, kD, 0, kD, p.skip ? 1 : 0};
      const Target te1{w.dg, kD, kD, kD, 0};
      PGMP_TRY(launch_bwd_in(st, dhid, kD, E, P + p.w1, ld1, 2 * nd, ed, te0, p.skip ? &te1 : nullptr));
    }
    float* tmp = dh;
    dh = dhp;
    dhp = tmp;
  }
  // h_{-1} = h0, g_{-1} = g0
  PGMP_LAUNCH(add_into_kernel, blocks_for(nN), 256, 0, st, w.dh0, dh, nN);
  if (E > 0) PGMP_LAUNCH(add_into_kernel, blocks_for(nE), 256, 0, st, w.dg0, w.dg, nE);
  const int xin = p.node_emb.dims[0];
  PGMP_TRY(mlp_backward(st, w, p.node_emb, w.node_emb, src1(p.x, xin), N, P, G, w.dh0, p.grad_x ? 1 : 0,
                        Target{p.grad_x, xin, 0, xin, 0}, none));
  PGMP_TRY(mlp_backward(st, w, p.edge_emb, w.edge_emb, src1(p.edge_attr, p.edge_emb.dims[0]), E, P, G, w.dg0, 0, none, none));
  return PGMP_OK;
}

}  // namespace pgmp

using namespace pgmp;

extern "C" uint64_t pgmp_mpn_train_workspace_bytes(const pgmp_mpn_train_params* p) {
  if (!p || p->num_nodes <= 0 || p->num_edges < 0 || p->steps < 1 || p->steps > kMaxSteps) return 0;
  pgmp_mpn_train_params q = *p;
  q.workspace = nullptr;
  return carve_train(q).bytes;
}

extern "C" int pgmp_mpn_train_forward(const pgmp_mpn_train_params* p, pgmp_stream_t stream) {
  const int rc = validate_train(p, false);
  if (rc != PGMP_OK) return rc;
  return mpn_train_forward(*p, static_cast<cudaStream_t>(stream));
}

extern "C" int pgmp_mpn_train_backward(const pgmp_mpn_train_params* p, pgmp_stream_t stream) {
  const int rc = validate_train(p, true);
  if (rc != PGMP_OK) return rc;
  return mpn_train_backward(*p, static_cast<cudaStream_t>(stream));
}
