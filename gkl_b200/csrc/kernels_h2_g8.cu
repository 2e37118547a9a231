// H2 sweep kernels, 8 lanes per read (reads of 65..128 rows).
#include "pairhmm_kernels.h"
namespace gklb {
void kernel_entries_h2_g8(std::vector<KernelEntry>& v) {
  GKLB_H2_ROW_ENTRIES(v, 8)
#ifdef GKLB_EXPERIMENTAL
  v.push_back(GKLB_E_H2(8, 13, 12));
#endif
}
}  // namespace gklb
