// Range-extended fp32 rerun kernels, 8 lanes per read.
#include "pairhmm_kernels.h"
namespace gklb {
const void* r2_kernel_g8(int K) {
  switch (K) {
    case 8: return reinterpret_cast<const void*>(&k_r2_list<8, 8, 8>);
    case 9: return reinterpret_cast<const void*>(&k_r2_list<8, 9, 8>);
    case 10: return reinterpret_cast<const void*>(&k_r2_list<8, 10, 8>);
    case 11: return reinterpret_cast<const void*>(&k_r2_list<8, 11, 8>);
    case 12: return reinterpret_cast<const void*>(&k_r2_list<8, 12, 8>);
    case 13: return reinterpret_cast<const void*>(&k_r2_list<8, 13, 8>);
    case 14: return reinterpret_cast<const void*>(&k_r2_list<8, 14, 8>);
    case 15: return reinterpret_cast<const void*>(&k_r2_list<8, 15, 8>);
    case 16: return reinterpret_cast<const void*>(&k_r2_list<8, 16, 8>);
    default: return nullptr;
  }
}
}  // namespace gklb
