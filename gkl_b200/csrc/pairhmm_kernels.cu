// pairhmm_kernels.cu -- instantiations of the sweep kernels and their registry.
//
// Product instantiations: packed-fp32 (VF2) and fp64 (VD1) for every length class of
// engine.cu's class table, plus the multi-pass variants for reads longer than 256 rows.
// The remaining entries (VF1, VAR 0, alternative warp counts) exist so that the design choices
// can be re-measured on the GPU (bench/sweep.py); the engine only uses them when told to
// through GKLB_FORCE_KERNEL.
#include "pairhmm_kernels.h"

namespace gklb {

// One warp per record: raw batch arenas -> top-padded class records.
__global__ void k_pack_reads(const __grid_constant__ PackParams p) {
  const int g = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (g >= p.n_rec_total) return;
  int ci = 0;
  while (ci + 1 < p.n_classes && g >= p.cls[ci + 1].rec_begin) ci++;
  const PackClass& c = p.cls[ci];
  const int rec = g - c.rec_begin;
  const int rid = c.rec_rid[rec];
  int64_t off = 0;
  int len = 0;
  if (rid >= 0) {
    off = p.read_off[rid];
    len = (int)(p.read_off[rid + 1] - off);
  }
  const int npad = c.rows - len;
  uint8_t* r = c.records + (size_t)rec * 5 * c.stride;
  for (int row = lane; row < c.stride; row += 32) {
    uint8_t b = 0, q = 0, i = 0, d = 0, cg = 0;
    if (row >= npad && row < c.rows) {
      const int64_t s = off + (row - npad);
      b = base_nibble(p.bases[s]);
      q = p.quals[s] & 127;  // avx-pairhmm-template.h:134-136,149
      i = p.ins[s] & 127;
      d = p.del[s] & 127;
      cg = p.gcp[s] & 127;
    }
    r[row] = b;
    r[c.stride + row] = q;
    r[2 * c.stride + row] = i;
    r[3 * c.stride + row] = d;
    r[4 * c.stride + row] = cg;
  }
}

// One block per haplotype of a tile: rewrite the nibbles of a panel image from device-resident bases
// (the lengths, and therefore the image layout, are those of the staged batch).
__global__ void k_fill_panel(uint8_t* image, int n_haps, int hap0, const int64_t* hap_off, const uint8_t* bases) {
  const int i = blockIdx.x;
  if (i >= n_haps) return;
  const int32_t* hpos = reinterpret_cast<const int32_t*>(image);
  const int32_t* hlen = hpos + n_haps;
  const int64_t o = hap_off[hap0 + i];
  uint8_t* dst = image + hpos[i] + 1;
  for (int c = threadIdx.x; c < hlen[i]; c += blockDim.x) dst[c] = panel_byte(bases[o + c]);
}

cudaError_t launch_fill_panel(uint8_t* image, int n_haps, int hap0, const int64_t* hap_off, const uint8_t* bases,
                              cudaStream_t s) {
  if (n_haps <= 0) return cudaSuccess;
  k_fill_panel<<<n_haps, 128, 0, s>>>(image, n_haps, hap0, hap_off, bases);
  return cudaGetLastError();
}

#define GKLB_TASKS(P, G, K, W, M, V) reinterpret_cast<const void*>(&k_sweep_tasks<P, G, K, W, M, V>)
#define GKLB_LIST(P, G, K, W, M, V) reinterpret_cast<const void*>(&k_sweep_list<P, G, K, W, M, V>)
#define E_F2(G, K, W, M, V) {POL_F2, G, K, W, M, V, 2, GKLB_TASKS(VF2, G, K, W, M, V), nullptr}
#define E_F1(G, K, W, M, V) {POL_F1, G, K, W, M, V, 1, GKLB_TASKS(VF1, G, K, W, M, V), nullptr}
#define E_D1(G, K, W, M, V) {POL_D1, G, K, W, M, V, 1, GKLB_TASKS(VD1, G, K, W, M, V), GKLB_LIST(VD1, G, K, W, M, V)}

#define E_H2(G, K, W) {POL_H2, G, K, W, false, 5, 1, reinterpret_cast<const void*>(&k_h2_tasks<G, K, W>), nullptr}

static const KernelEntry g_table[] = {
    // ---- product: packed fp32, folded recurrence with the shared-memory prior table (VAR 3) ----
    // 12 warps per SM in 168 registers without the prior prefetch (VAR 5); K = 8 spills ~60 bytes outside the steady
    // loop and is still 1-2 % faster than 8 warps with the prefetch (VAR 4); the multi-pass kernel keeps 8 warps
    E_F2(8, 4, 12, false, 5),  E_F2(8, 5, 12, false, 5),  E_F2(8, 6, 12, false, 5),  E_F2(8, 7, 12, false, 5),
    E_F2(8, 8, 12, false, 5),  E_F2(16, 5, 12, false, 5), E_F2(16, 6, 12, false, 5), E_F2(16, 7, 12, false, 5),
    E_F2(16, 8, 12, false, 5), E_F2(32, 5, 12, false, 5), E_F2(32, 6, 12, false, 5), E_F2(32, 7, 12, false, 5),
    E_F2(32, 8, 12, false, 5), E_F2(32, 8, 8, true, 4),
    // ---- product: fp64 (useDoublePrecision and the rerun of flagged pairs) ----
    E_D1(8, 4, 8, false, 3),  E_D1(8, 5, 8, false, 3),  E_D1(8, 6, 8, false, 3),  E_D1(8, 7, 8, false, 3),
    E_D1(8, 8, 8, false, 3),  E_D1(16, 5, 8, false, 3), E_D1(16, 6, 8, false, 3), E_D1(16, 7, 8, false, 3),
    E_D1(16, 8, 8, false, 3), E_D1(32, 5, 8, false, 3), E_D1(32, 6, 8, false, 3), E_D1(32, 7, 8, false, 3),
    E_D1(32, 8, 8, false, 3), E_D1(32, 8, 8, true, 3),
#ifdef GKLB_EXPERIMENTAL
    // ---- measurement only ----
    E_H2(8, 13, 8), E_H2(8, 13, 10), E_H2(8, 13, 12), E_H2(16, 7, 12), E_H2(16, 7, 16), E_H2(16, 10, 12), E_H2(16, 10, 8),
    E_H2(8, 7, 16), E_H2(32, 5, 12), E_H2(32, 5, 16),
    E_F2(16, 7, 8, false, 3), E_F2(16, 7, 8, false, 2), E_F2(16, 7, 8, false, 1), E_F2(16, 7, 8, false, 0),
    E_F2(16, 7, 8, false, 4), E_F2(16, 7, 12, false, 4), E_F2(16, 7, 8, false, 5), E_F2(32, 5, 8, false, 4), E_F2(16, 8, 8, false, 3), E_F2(16, 8, 8, false, 2), E_F2(32, 4, 12, false, 5), E_F2(8, 4, 16, false, 5),
    E_F2(16, 8, 8, false, 4), E_F2(32, 8, 8, false, 4), E_F2(8, 8, 8, false, 4),
    E_F2(16, 7, 12, false, 6), E_F2(16, 7, 8, false, 6), E_F2(16, 7, 10, false, 6), E_F2(16, 7, 10, false, 5),
    E_F1(8, 13, 8, false, 4), E_F1(8, 13, 8, false, 3), E_F1(8, 13, 12, false, 4), E_F1(16, 7, 16, false, 4),
    E_D1(16, 7, 8, false, 2), E_D1(8, 13, 8, false, 3), E_D1(32, 4, 8, false, 3), E_D1(16, 10, 8, false, 3), E_D1(8, 7, 8, false, 3),
#endif
};

const KernelEntry* kernel_table(int* n) {
  *n = (int)(sizeof(g_table) / sizeof(g_table[0]));
  return g_table;
}

const KernelEntry* find_kernel(int policy, int G, int K, int warps, int multi, int var) {
  int n;
  const KernelEntry* t = kernel_table(&n);
  for (int i = 0; i < n; i++)
    if (t[i].policy == policy && t[i].G == G && t[i].K == K && (warps <= 0 || t[i].warps == warps) &&
        t[i].multi == multi && (var < 0 || t[i].var == var))
      return &t[i];
  return nullptr;
}

const void* mega_kernel(int policy, int list_mode, int warps) {
  if (policy == POL_F2 && !list_mode && warps == 12) return reinterpret_cast<const void*>(&k_mega_tasks<VF2, 12>);
  if (policy == POL_F2 && !list_mode) return reinterpret_cast<const void*>(&k_mega_tasks<VF2, 8>);
  if (policy == POL_D1 && !list_mode) return reinterpret_cast<const void*>(&k_mega_tasks<VD1, 8>);
  if (policy == POL_D1 && list_mode) return reinterpret_cast<const void*>(&k_mega_list<VD1, 8>);
  return nullptr;
}

cudaError_t launch_mega(const void* fn, const MegaParams& m, uint32_t slot_bytes, int list_mode, int grid, int threads,
                        size_t smem, cudaStream_t s) {
  void* args_tasks[] = {const_cast<MegaParams*>(&m), &slot_bytes};
  void* args_list[] = {const_cast<MegaParams*>(&m), &slot_bytes};
  return cudaLaunchKernel(fn, dim3(grid), dim3(threads), list_mode ? args_list : args_tasks, smem, s);
}

cudaError_t launch_sweep(const void* fn, const SweepParams& p, int grid, int threads, size_t smem, cudaStream_t s) {
  void* args[] = {const_cast<SweepParams*>(&p)};
  return cudaLaunchKernel(fn, dim3(grid), dim3(threads), args, smem, s);
}

cudaError_t launch_h2_kernel(const void* fn, const H2Params& p, int grid, int threads, size_t smem, cudaStream_t s) {
  void* args[] = {const_cast<H2Params*>(&p)};
  return cudaLaunchKernel(fn, dim3(grid), dim3(threads), args, smem, s);
}

cudaError_t launch_pack(const PackParams& p, cudaStream_t s) {
  const int threads = 256;
  const long long total = (long long)p.n_rec_total * 32;
  const int grid = (int)((total + threads - 1) / threads);
  if (grid == 0) return cudaSuccess;
  k_pack_reads<<<grid, threads, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace gklb
