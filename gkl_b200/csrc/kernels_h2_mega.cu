// The multi-class H2 sweep kernel (one launch for every single-pass length class of a batch).
#include "pairhmm_kernels.h"
namespace gklb {
const void* h2_mega_kernel() { return reinterpret_cast<const void*>(&k_h2_mega<8>); }
}  // namespace gklb
