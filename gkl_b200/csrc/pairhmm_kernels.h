// pairhmm_kernels.h -- host-side registry of the compiled sweep kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include "pairhmm_device.cuh"
#include "pairhmm_h2.cuh"

namespace gklb {

enum Policy { POL_F2 = 0, POL_F1 = 1, POL_D1 = 2, POL_H2 = 3 };

struct KernelEntry {
  int policy, G, K, warps, multi, var;
  int nr;                  // reads carried per lane
  const void* fn_tasks;    // k_sweep_tasks instantiation
  const void* fn_list;     // k_sweep_list instantiation (nullptr when nr != 1)
};

// All compiled instantiations.
const KernelEntry* kernel_table(int* n);
const KernelEntry* find_kernel(int policy, int G, int K, int warps, int multi, int var);

cudaError_t launch_sweep(const void* fn, const SweepParams& p, int grid, int threads, size_t smem, cudaStream_t s);
// Multi-class kernels (8 warps per CTA): POL_F2 / POL_D1 task kernels and the POL_D1 rerun-list kernel.
const void* mega_kernel(int policy, int list_mode, int warps = 8);
cudaError_t launch_mega(const void* fn, const MegaParams& m, uint32_t slot_bytes, int list_mode, int grid, int threads,
                        size_t smem, cudaStream_t s);
cudaError_t launch_h2_kernel(const void* fn, const H2Params& p, int grid, int threads, size_t smem, cudaStream_t s);
cudaError_t launch_pack(const PackParams& p, cudaStream_t s);
cudaError_t launch_fill_panel(uint8_t* image, int n_haps, int hap0, const int64_t* hap_off, const uint8_t* bases,
                              cudaStream_t s);

}  // namespace gklb
