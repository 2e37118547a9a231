// H2 sweep kernels, 16 lanes per read (reads of 129..256 rows).
#include "pairhmm_kernels.h"
namespace gklb {
void kernel_entries_h2_g16(std::vector<KernelEntry>& v) {
  GKLB_H2_ROW_ENTRIES(v, 16)
#ifdef GKLB_EXPERIMENTAL
  v.push_back(GKLB_E_H2(16, 7, 12));
  v.push_back(GKLB_E_H2(16, 7, 16));
#endif
}
}  // namespace gklb
