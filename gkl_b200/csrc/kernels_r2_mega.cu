// The multi-class range-extended rerun kernel.
#include "pairhmm_kernels.h"
namespace gklb {
const void* r2_mega_kernel() { return reinterpret_cast<const void*>(&k_r2_mega<8>); }
const void* r2_kernel(int G, int K) { return G == 4 ? r2_kernel_g4(K) : G == 8 ? r2_kernel_g8(K) : G == 16 ? r2_kernel_g16(K) : nullptr; }
}  // namespace gklb
