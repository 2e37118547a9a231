// H2 sweep kernels, 4 lanes per read (reads of up to 64 rows).
#include "pairhmm_kernels.h"
namespace gklb {
void kernel_entries_h2_g4(std::vector<KernelEntry>& v) { GKLB_H2_ROW_ENTRIES(v, 4) }
}  // namespace gklb
