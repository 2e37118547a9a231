// Range-extended fp32 rerun kernels, 16 lanes per read.
#include "pairhmm_kernels.h"
namespace gklb {
const void* r2_kernel_g16(int K) {
  switch (K) {
    case 8: return reinterpret_cast<const void*>(&k_r2_list<16, 8, 8>);
    case 9: return reinterpret_cast<const void*>(&k_r2_list<16, 9, 8>);
    case 10: return reinterpret_cast<const void*>(&k_r2_list<16, 10, 8>);
    case 11: return reinterpret_cast<const void*>(&k_r2_list<16, 11, 8>);
    case 12: return reinterpret_cast<const void*>(&k_r2_list<16, 12, 8>);
    case 13: return reinterpret_cast<const void*>(&k_r2_list<16, 13, 8>);
    case 14: return reinterpret_cast<const void*>(&k_r2_list<16, 14, 8>);
    case 15: return reinterpret_cast<const void*>(&k_r2_list<16, 15, 8>);
    case 16: return reinterpret_cast<const void*>(&k_r2_list<16, 16, 8>);
    default: return nullptr;
  }
}
}  // namespace gklb
