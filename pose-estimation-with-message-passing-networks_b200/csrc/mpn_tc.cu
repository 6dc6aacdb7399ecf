// tcgen05 / TMEM implementation of the message-passing steps (PGMP_PRECISION_TC) -- placeholder
// until the tensor-core kernels land; fails loudly instead of falling back.
#include "mpn_common.cuh"

namespace pgmp {
int mpn_forward_tc(const pgmp_mpn_params& p, const MpnWorkspace& w, cudaStream_t st) {
  (void)p; (void)w; (void)st;
  return set_error(PGMP_ERR_INVALID, "PGMP_PRECISION_TC is not built into this libpgmp.so");
}
}  // namespace pgmp
