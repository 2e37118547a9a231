// pairhmm_kernels.h -- host-side registry of the compiled sweep kernels.  The instantiations are spread over
// several translation units (kernels_*.cu) so that they compile in parallel.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include <vector>

#include "pairhmm_device.cuh"
#include "pairhmm_h2.cuh"
#include "pairhmm_r2.cuh"

namespace gklb {

enum Policy { POL_F2 = 0, POL_F1 = 1, POL_D1 = 2, POL_H2 = 3 };

struct KernelEntry {
  int policy, G, K, warps, multi, var;
  int nr;                  // reads carried per lane
  const void* fn_tasks;    // task kernel (k_sweep_tasks / k_h2_tasks instantiation)
  const void* fn_list;     // k_sweep_list instantiation (fp64 only)
};

// All compiled instantiations.
const KernelEntry* kernel_table(int* n);
const KernelEntry* find_kernel(int policy, int G, int K, int warps, int multi, int var);

// Multi-class kernels (8 warps per CTA): the fp64 task kernel (list_mode 0) and rerun-list kernel (1); the H2 sweep.
const void* mega_kernel(int policy, int list_mode);
const void* h2_mega_kernel();
// Range-extended fp32 rerun (pairhmm_r2.cuh): per class (G = 4, 8, 16; K = 8..16; 8 warps) and multi-class.
const void* r2_kernel(int G, int K);
const void* r2_mega_kernel();
const void* r2_kernel_g4(int K);
const void* r2_kernel_g8(int K);
const void* r2_kernel_g16(int K);

cudaError_t launch_pack(const PackParams& p, cudaStream_t s);
cudaError_t launch_fill_panel(uint8_t* image, int n_haps, int hap0, const int64_t* hap_off, const uint8_t* bases,
                              cudaStream_t s);
cudaError_t launch_fill_pair_panel(uint8_t* image, int n_pairs, const int64_t* hap_off, const uint8_t* bases, cudaStream_t s);

// per translation unit
void kernel_entries_h2_g4(std::vector<KernelEntry>& v);
void kernel_entries_h2_g8(std::vector<KernelEntry>& v);
void kernel_entries_h2_g16(std::vector<KernelEntry>& v);
void kernel_entries_h2_g32(std::vector<KernelEntry>& v);
void kernel_entries_d1(std::vector<KernelEntry>& v);
void kernel_entries_misc(std::vector<KernelEntry>& v);

#define GKLB_E_H2(G, K, W) \
  KernelEntry { POL_H2, G, K, W, 0, 5, 1, reinterpret_cast<const void*>(&k_h2_tasks<G, K, W>), nullptr }
#define GKLB_H2_ROW_ENTRIES(v, G)                                                                             \
  v.push_back(GKLB_E_H2(G, 8, 8)); v.push_back(GKLB_E_H2(G, 9, 8)); v.push_back(GKLB_E_H2(G, 10, 8));          \
  v.push_back(GKLB_E_H2(G, 11, 8)); v.push_back(GKLB_E_H2(G, 12, 8)); v.push_back(GKLB_E_H2(G, 13, 8));        \
  v.push_back(GKLB_E_H2(G, 14, 8)); v.push_back(GKLB_E_H2(G, 15, 8)); v.push_back(GKLB_E_H2(G, 16, 8));        \
  /* up to 10 rows per lane fit 168 registers: a third warp per scheduler for single-class launches */         \
  v.push_back(GKLB_E_H2(G, 8, 12)); v.push_back(GKLB_E_H2(G, 9, 12)); v.push_back(GKLB_E_H2(G, 10, 12));

}  // namespace gklb
