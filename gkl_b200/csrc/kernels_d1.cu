// fp64 kernels: useDoublePrecision (task kernels) and the rerun of flagged pairs (list kernels), per length class
// and as multi-class launches.
#include "pairhmm_kernels.h"
namespace gklb {

#define E_D1(G, K, W, M, V)                                                                              \
  KernelEntry { POL_D1, G, K, W, M, V, 1, reinterpret_cast<const void*>(&k_sweep_tasks<VD1, G, K, W, M, V>), \
                reinterpret_cast<const void*>(&k_sweep_list<VD1, G, K, W, M, V>) }

void kernel_entries_d1(std::vector<KernelEntry>& v) {
  const KernelEntry e[] = {
      E_D1(8, 4, 12, false, 3),  E_D1(8, 5, 12, false, 3),  E_D1(8, 6, 12, false, 3),  E_D1(8, 7, 12, false, 3),
      E_D1(8, 8, 8, false, 3),  E_D1(16, 5, 12, false, 3), E_D1(16, 6, 12, false, 3), E_D1(16, 7, 12, false, 3),
      E_D1(16, 8, 8, false, 3), E_D1(16, 9, 8, false, 3), E_D1(16, 10, 8, false, 3), E_D1(32, 5, 12, false, 3), E_D1(32, 6, 12, false, 3), E_D1(32, 7, 12, false, 3),
      E_D1(32, 8, 8, false, 3), E_D1(32, 9, 8, false, 3), E_D1(32, 10, 8, false, 3),  // 288 / 320 rows: 2 x 300 reads in one pass
      E_D1(32, 8, 8, true, 3),
  };
  for (const auto& x : e) v.push_back(x);
}

const void* mega_kernel(int policy, int list_mode) {
  if (policy != POL_D1) return nullptr;
  return list_mode ? reinterpret_cast<const void*>(&k_mega_list<VD1, 8>) : reinterpret_cast<const void*>(&k_mega_tasks<VD1, 8>);
}

}  // namespace gklb
