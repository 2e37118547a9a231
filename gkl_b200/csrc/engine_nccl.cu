// engine_nccl.cu -- placeholder, filled in below
#include "engine_internal.h"
namespace gklb {
bool nccl_available() { return false; }
int sharded_compute_nccl(const std::vector<gklb_engine*>&, const gklb_pairhmm_batch*, const std::vector<int>&, double*,
                         gklb_pairhmm_stats*) {
  return fail(GKLB_ERR_STATE, "the NCCL sharding path is not built");
}
}  // namespace gklb
