// Interface of the tcgen05 forward product of the training step (mpn_train_tc.cu), used by mpn_train.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgmp {

// Y[M][64] = act(A W^T + bias + add1[idx1] + add2[idx2]) for an E-level product of the training forward.
// A = one or two column blocks of 64 floats side by side (K = 64 or 128), fp32 row-major; W element (o, k) at
// W[o * ldw + coloff + k], 64 outputs.  Same contract as launch_fwd / lin_fwd_kernel (mpn_train.cu).
struct LinFwdTc {
  const float* a0; int lda0;        // columns [0, 64)
  const float* a1; int lda1;        // columns [64, 128) or null
  int64_t M;
  const float* W; int ldw, coloff;
  const float* bias;
  const float* add1; const int64_t* idx1; int add1_ld;
  const float* add2; const int64_t* idx2;
  int relu;
  float* Y; int ldy;
};

int launch_lin_fwd_tc(cudaStream_t st, const LinFwdTc& q);

}  // namespace pgmp
