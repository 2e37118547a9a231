// H2 sweep kernels, 32 lanes per read: reads of 257..320 rows (2 x 300 sequencing) in one pass.
#include "pairhmm_kernels.h"
namespace gklb {
void kernel_entries_h2_g32(std::vector<KernelEntry>& v) {
  v.push_back(GKLB_E_H2(32, 9, 8));
  v.push_back(GKLB_E_H2(32, 10, 8));
  v.push_back(GKLB_E_H2(32, 9, 12));
  v.push_back(GKLB_E_H2(32, 10, 12));
}
}  // namespace gklb
