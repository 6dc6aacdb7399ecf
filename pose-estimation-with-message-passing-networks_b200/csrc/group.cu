// Grouping tail (threshold -> multicut GAEC -> person assembly) -- entry points; kernels follow.
#include "common.cuh"

using namespace pgmp;

extern "C" uint64_t pgmp_group_workspace_bytes(const pgmp_group_params* p) {
  (void)p;
  return 0;
}

extern "C" int pgmp_group_persons(const pgmp_group_params* p, pgmp_stream_t stream) {
  (void)p; (void)stream;
  return set_error(PGMP_ERR_INVALID, "pgmp_group_persons: not built yet");
}
