// pairhmm_kernels.cu -- the small kernels (read packing, panel refill), the multi-pass packed-read sweep, the
// measurement-only variants, and the registry of all compiled sweep kernels.
//
// Product instantiations: H2 (two haplotypes per lane, kernels_h2_*.cu) for every single-pass length class,
// packed-read VF2 multi-pass for reads longer than 256 rows, fp64 VD1 (kernels_d1.cu).  The remaining entries
// (VF2/VF1 single-pass variants) exist so that the design choices can be re-measured on the GPU (bench/sweep.py,
// make EXPERIMENTAL=1); the engine only uses them when told to through GKLB_FORCE_KERNEL.
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

// The same for a pair image (pairhmm_h2.cuh): one block per pair.
__global__ void k_fill_pair_panel(uint8_t* image, int n_pairs, const int64_t* hap_off, const uint8_t* bases) {
  const int q = blockIdx.x;
  if (q >= n_pairs) return;
  const int32_t* ppos = reinterpret_cast<const int32_t*>(image);
  const int32_t* lenA = ppos + n_pairs;
  const int32_t* lenB = lenA + n_pairs;
  const int32_t* idxA = lenB + n_pairs;
  const int32_t* idxB = idxA + n_pairs;
  const int a = idxA[q], b = idxB[q] >= 0 ? idxB[q] : a;
  const int64_t oa = hap_off[a], ob = hap_off[b];
  const int la = lenA[q], lb = lenB[q];
  uint8_t* dst = image + ppos[q] + 1;
  for (int c = threadIdx.x; c < la; c += blockDim.x)
    dst[c] = (uint8_t)(base_index(bases[oa + c]) | (c < lb ? base_index(bases[ob + c]) << 3 : 0));
}

cudaError_t launch_fill_panel(uint8_t* image, int n_haps, int hap0, const int64_t* hap_off, const uint8_t* bases,
                              cudaStream_t s) {
  if (n_haps <= 0) return cudaSuccess;
  k_fill_panel<<<n_haps, 128, 0, s>>>(image, n_haps, hap0, hap_off, bases);
  return cudaGetLastError();
}

cudaError_t launch_fill_pair_panel(uint8_t* image, int n_pairs, const int64_t* hap_off, const uint8_t* bases, cudaStream_t s) {
  if (n_pairs <= 0) return cudaSuccess;
  k_fill_pair_panel<<<n_pairs, 128, 0, s>>>(image, n_pairs, hap_off, bases);
  return cudaGetLastError();
}

cudaError_t launch_pack(const PackParams& p, cudaStream_t s) {
  const int threads = 256;
  const long long total = (long long)p.n_rec_total * 32;
  const int grid = (int)((total + threads - 1) / threads);
  if (grid == 0) return cudaSuccess;
  k_pack_reads<<<grid, threads, 0, s>>>(p);
  return cudaGetLastError();
}

#define GKLB_TASKS(P, G, K, W, M, V) reinterpret_cast<const void*>(&k_sweep_tasks<P, G, K, W, M, V>)
#define E_F2(G, K, W, M, V) KernelEntry{POL_F2, G, K, W, M, V, 2, GKLB_TASKS(VF2, G, K, W, M, V), nullptr}
#define E_F1(G, K, W, M, V) KernelEntry{POL_F1, G, K, W, M, V, 1, GKLB_TASKS(VF1, G, K, W, M, V), nullptr}

void kernel_entries_misc(std::vector<KernelEntry>& v) {
  // product: reads of more than 256 rows, two reads per lane, 32 x 8 rows per pass, carry line in global memory
  v.push_back(E_F2(32, 8, 8, true, 4));
#ifdef GKLB_EXPERIMENTAL
  // measurement only: the round-1 packed-read kernel (two reads per lane) and its variants
  const KernelEntry e[] = {
      E_F2(16, 7, 12, false, 5), E_F2(16, 7, 8, false, 4), E_F2(16, 7, 8, false, 3), E_F2(16, 7, 8, false, 2),
      E_F2(16, 7, 8, false, 1),  E_F2(16, 7, 8, false, 0), E_F2(32, 5, 12, false, 5), E_F1(8, 13, 8, false, 4),
  };
  for (const auto& x : e) v.push_back(x);
#endif
}

const KernelEntry* kernel_table(int* n) {
  static const std::vector<KernelEntry> table = [] {
    std::vector<KernelEntry> v;
    kernel_entries_h2_g4(v);
    kernel_entries_h2_g8(v);
    kernel_entries_h2_g16(v);
    kernel_entries_h2_g32(v);
    kernel_entries_d1(v);
    kernel_entries_misc(v);
    return v;
  }();
  *n = (int)table.size();
  return table.data();
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

}  // namespace gklb
