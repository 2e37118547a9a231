// engine.cu -- one PairHMM engine (one device, one stream): planning, staging, launches, and the
// gklb_engine_* part of the C-ABI (include/gklb_pairhmm.h).  The process-global surface the JNI layer
// calls (gklb_pairhmm_init / _compute / _done, device pool, sharding) is engine_global.cu.
//
// Replaces the native half of GKL's PairHMM binding:
//   computeLikelihoodsNative's pair loop                  pairhmm/IntelPairHmm.cc:150-169
//   JavaData::getData (testcase expansion)                pairhmm/JavaData.h:65-111
// The (read, haplotype) cross product is never materialised: reads are bucketed into length
// classes and packed on the device, haplotypes become shared-memory panel images, and pairs are
// addressed by (record, haplotype) index arithmetic inside the kernels.
//
// How a batch runs in fp32 mode (useDoublePrecision = false):
//   k_pack_reads           raw arenas -> top-padded class records
//   k_h2_tasks / k_h2_mega forward sweep, two haplotypes per lane (pairhmm_h2.cuh); reads of more than 256 rows
//                          take the multi-pass packed kernel k_sweep_tasks<VF2, 32, 8, multi>
//   k_sweep_list / k_mega_list<VD1>   fp64 rerun of the pairs whose scaled fp32 sum is < 1e-28f (IntelPairHmm.cc:159)
// useDoublePrecision = true runs k_sweep_tasks<VD1> over all pairs.  The launches of a staged batch are planned
// once (do_stage) and replayed by every run.
//
// There is no CPU compute path in this file: without a compute-capability-10 device every
// entry point fails with GKLB_ERR_NO_DEVICE.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>

#include "engine_internal.h"
#include "pairhmm_tables.h"

namespace gklb {

cudaError_t DevBuf::ensure(size_t bytes) {
  if (bytes <= cap) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
  if (e == cudaSuccess) cap = want;
  return e;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}
cudaError_t HostBuf::ensure(size_t bytes) {
  if (bytes <= cap) return cudaSuccess;
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
  const size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMallocHost(&p, want);
  if (e == cudaSuccess) cap = want;
  return e;
}
void HostBuf::release() {
  if (p) cudaFreeHost(p);
  p = nullptr;
  cap = 0;
}

}  // namespace gklb

namespace gklb {

#define CU GKLB_CU

namespace {

// Length classes.  A read goes to the first class that holds it; reads longer than the last class take n_pass
// passes of the multi-pass kernel.
//   fp64 kernels (useDoublePrecision, rerun): rows per pass = G * K from this table
struct ClassDef { int G, K; };
const ClassDef kClassesD1[] = {{8, 4},  {8, 5},  {8, 6},  {8, 7},  {8, 8},  {16, 5},  {16, 6}, {16, 7}, {16, 8},
                               {16, 9}, {16, 10}, {32, 5}, {32, 6}, {32, 7}, {32, 8}, {32, 9}, {32, 10}};
const int kNumClassesD1 = (int)(sizeof(kClassesD1) / sizeof(kClassesD1[0]));
//   fp32 H2 kernels: G = 4, 8, 16 lanes x K = 8..16 rows (32..64 rows in steps of 4, 72..128 in steps of 8,
//   144..256 in steps of 16) and G = 32 x K = 9, 10 (288 and 320 rows); cfg = gi * 9 + (K - 8), cfg 27 is unused
const int kNumClassesH2 = 30;
inline int h2_G(int cfg) { return 4 << (cfg / 9); }
inline int h2_K(int cfg) { return 8 + cfg % 9; }
const int kH2Warps = 8;
const int kMultiG = 32, kMultiK = 8;
const int kSinglePassMax = 320;  // 32 x 10: the largest single-pass class of both tables
const int kSmemMax = 232448;  // 227 KB opt-in dynamic shared memory per CTA on sm_100

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int d1_cfg_for_rows(int rows) {
  for (int i = 0; i < kNumClassesD1; i++)
    if (kClassesD1[i].G * kClassesD1[i].K >= rows) return i;
  return -1;
}

int upload_tables(gklb_engine* e) {
  const HostTables& t = host_tables();
  const size_t bytes = sizeof(float) * (kPh2prSize + kMmSize) + sizeof(double) * (kPh2prSize + kMmSize);
  CU(e->d_tables.ensure(bytes + 64));
  uint8_t* base = static_cast<uint8_t*>(e->d_tables.p);
  size_t off = 0;
  e->d_ph2pr_d = reinterpret_cast<const double*>(base + off);
  CU(cudaMemcpy(base + off, t.ph2pr_d, sizeof(t.ph2pr_d), cudaMemcpyHostToDevice));
  off += sizeof(t.ph2pr_d);
  e->d_mm_d = reinterpret_cast<const double*>(base + off);
  CU(cudaMemcpy(base + off, t.mm_d, sizeof(t.mm_d), cudaMemcpyHostToDevice));
  off += sizeof(t.mm_d);
  e->d_ph2pr_f = reinterpret_cast<const float*>(base + off);
  CU(cudaMemcpy(base + off, t.ph2pr_f, sizeof(t.ph2pr_f), cudaMemcpyHostToDevice));
  off += sizeof(t.ph2pr_f);
  e->d_mm_f = reinterpret_cast<const float*>(base + off);
  CU(cudaMemcpy(base + off, t.mm_f, sizeof(t.mm_f), cudaMemcpyHostToDevice));
  return GKLB_OK;
}

int parse_forced(gklb_engine* e) {
  e->forced = false;
  const char* s = getenv("GKLB_FORCE_KERNEL");
  if (!s || !*s) return GKLB_OK;
  char pol[8] = {0};
  int G, K, W, V;
  if (sscanf(s, "%7[^,],%d,%d,%d,%d", pol, &G, &K, &W, &V) != 5)
    return fail(GKLB_ERR_INVALID, "GKLB_FORCE_KERNEL must be policy,G,K,warps,var (got '%s')", s);
  int p = !strcmp(pol, "f2") ? POL_F2 : !strcmp(pol, "f1") ? POL_F1 : !strcmp(pol, "d1") ? POL_D1
          : !strcmp(pol, "h2") ? POL_H2 : -1;
  if (p < 0 || !find_kernel(p, G, K, W, 0, V)) return fail(GKLB_ERR_INVALID, "no compiled kernel for '%s'", s);
  e->forced = true;
  e->f_policy = p; e->f_G = G; e->f_K = K; e->f_warps = W; e->f_var = V;
  return GKLB_OK;
}

int set_kernel_attrs() {
  int n;
  const KernelEntry* t = kernel_table(&n);
  for (int i = 0; i < n; i++) {
    CU(cudaFuncSetAttribute(t[i].fn_tasks, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    if (t[i].fn_list) CU(cudaFuncSetAttribute(t[i].fn_list, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  }
  for (const void* fn : {h2_mega_kernel(), r2_mega_kernel(), mega_kernel(POL_D1, 0), mega_kernel(POL_D1, 1)})
    if (fn) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  for (int G : {4, 8, 16})
    for (int K = 8; K <= 16; K++) CU(cudaFuncSetAttribute(r2_kernel(G, K), cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  return GKLB_OK;
}

// Build the class instances of one region (appended to e->classes).
int plan_classes(gklb_engine* e, const gklb_pairhmm_batch* b, int region, int read_base) {
  const bool h2 = !e->use_double && !e->forced;
  const int n_single = h2 ? kNumClassesH2 : kNumClassesD1;
  std::vector<std::vector<int32_t>> rid(n_single), len(n_single);  // single-pass classes, by configuration
  std::vector<std::pair<int, int>> multi_inst;                     // (n_pass, instance index)
  std::vector<ClassInst> multi;
  ClassInst forced;
  for (int r = 0; r < b->n_reads; r++) {
    const int64_t len64 = b->read_off[r + 1] - b->read_off[r];
    if (len64 > (1 << 24)) return fail(GKLB_ERR_INVALID, "read %d is too long (%lld)", r, (long long)len64);
    const int L = (int)len64;
    if (e->forced && L <= e->f_G * e->f_K) {
      forced.rid.push_back(read_base + r);
      forced.len.push_back(L);
    } else if (L <= kSinglePassMax) {
      int ci = 0;
      if (h2) {
        const int gi = L <= 64 ? 0 : L <= 128 ? 1 : L <= 256 ? 2 : 3;
        ci = gi * 9 + std::max(0, (L + (4 << gi) - 1) / (4 << gi) - 8);
      } else {
        while (kClassesD1[ci].G * kClassesD1[ci].K < L) ci++;
      }
      rid[ci].push_back(read_base + r);
      len[ci].push_back(L);
    } else {
      const int cap = kMultiG * kMultiK;
      const int n_pass = (L + cap - 1) / cap;
      int inst = -1;
      for (auto& m : multi_inst)
        if (m.first == n_pass) inst = m.second;
      if (inst < 0) {
        ClassInst c;
        c.G = kMultiG; c.K = kMultiK; c.n_pass = n_pass; c.multi = true;
        c.kd = find_kernel(POL_D1, c.G, c.K, -1, 1, -1);
        c.kf = e->use_double ? c.kd : find_kernel(POL_F2, c.G, c.K, -1, 1, -1);
        c.cfg_d = kCfgMulti;
        inst = (int)multi.size();
        multi_inst.push_back({n_pass, inst});
        multi.push_back(c);
      }
      multi[inst].rid.push_back(read_base + r);
      multi[inst].len.push_back(L);
    }
  }
  if (h2) {
    // A class whose record count is not a multiple of the records per warp-task would be padded with filler records;
    // instead its remainder (the longest reads of the class) moves up to the next non-empty class, where a few
    // extra padding rows cost less than idle lane groups.  Only the largest class keeps fillers.
    int last = -1;
    for (int ci = 0; ci < n_single; ci++)
      if (!rid[ci].empty()) last = ci;
    for (int ci = 0; ci < last; ci++) {
      const int gpw = 32 / h2_G(ci);
      const int rem = (int)(rid[ci].size() % gpw);
      if (rem == 0) continue;
      int next = ci + 1;
      while (rid[next].empty()) next++;
      std::vector<int> order(rid[ci].size());
      for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return len[ci][x] < len[ci][y]; });
      std::vector<int32_t> keep_r, keep_l;
      for (size_t i = 0; i < order.size(); i++) {
        const int k = order[i];
        if (i + rem < order.size()) { keep_r.push_back(rid[ci][k]); keep_l.push_back(len[ci][k]); }
        else { rid[next].push_back(rid[ci][k]); len[next].push_back(len[ci][k]); }
      }
      rid[ci].swap(keep_r);
      len[ci].swap(keep_l);
    }
  }
  const size_t first = e->classes.size();
  for (int ci = 0; ci < n_single; ci++) {
    if (rid[ci].empty()) continue;
    ClassInst c;
    if (h2) {
      c.G = h2_G(ci); c.K = h2_K(ci);
      // launches of a single class run 12 warps per SM where the class fits 168 registers (K <= 10); the multi-class
      // kernel and the rerun kernels stay at kH2Warps whatever the entry says
      c.kf = (c.K <= 10) ? find_kernel(POL_H2, c.G, c.K, 12, 0, -1) : nullptr;
      if (!c.kf) c.kf = find_kernel(POL_H2, c.G, c.K, kH2Warps, 0, -1);
      c.cfg_f = ci;
      c.cfg_d = d1_cfg_for_rows(c.G * c.K);
      c.kd = find_kernel(POL_D1, kClassesD1[c.cfg_d].G, kClassesD1[c.cfg_d].K, -1, 0, -1);
    } else {
      c.G = kClassesD1[ci].G; c.K = kClassesD1[ci].K;
      c.kd = find_kernel(POL_D1, c.G, c.K, -1, 0, -1);
      c.kf = c.kd;
      c.cfg_d = ci;
    }
    c.rid.swap(rid[ci]);
    c.len.swap(len[ci]);
    e->classes.push_back(std::move(c));
  }
  if (!forced.rid.empty()) {
    forced.G = e->f_G; forced.K = e->f_K;
    forced.kf = find_kernel(e->f_policy, forced.G, forced.K, e->f_warps, 0, e->f_var);
    forced.cfg_d = d1_cfg_for_rows(forced.G * forced.K);
    if (forced.cfg_d >= 0)
      forced.kd = find_kernel(POL_D1, kClassesD1[forced.cfg_d].G, kClassesD1[forced.cfg_d].K, -1, 0, -1);
    e->classes.push_back(std::move(forced));
  }
  for (auto& c : multi) e->classes.push_back(std::move(c));
  for (size_t i = first; i < e->classes.size(); i++) {
    ClassInst& c = e->classes[i];
    if (!c.kf || !c.kd) return fail(GKLB_ERR_STATE, "kernel for class G=%d K=%d is not compiled", c.G, c.K);
    c.region = region;
    c.rows = c.n_pass * c.G * c.K;
    c.stride = (int)align_up((size_t)c.rows, 16);
    const int rpw = (32 / c.G) * c.kf->nr;
    while (c.rid.size() % rpw) { c.rid.push_back(-1); c.len.push_back(0); }
    c.n_rec = (int)c.rid.size();
  }
  return GKLB_OK;
}

// Per-warp slot: the task's packed records, then the prior table of 5 symbols x K rows x 32 lanes.  The geometry
// is the kernel's (the fp64 kernels of an H2 class have their own G and K).
uint32_t warp_slot_bytes(const ClassInst& c, const KernelEntry* k, bool list_mode) {
  const uint32_t rpw = (uint32_t)((32 / k->G) * k->nr);
  const uint32_t rec = (list_mode || k->multi) ? 0u : (uint32_t)align_up((size_t)rpw * 5 * c.stride, 128);
  const uint32_t tbl = (k->policy == POL_H2) ? (uint32_t)(kPriorSyms * k->K * 32 * 4)
                       : (k->var >= 3)          ? (uint32_t)(kPriorSyms * k->K * 32 * 8) : 0u;
  return rec + tbl;
}

uint32_t slot_bytes_total(const ClassInst& c, const KernelEntry* k, bool list_mode) {
  return (uint32_t)k->warps * (uint32_t)align_up((size_t)warp_slot_bytes(c, k, list_mode), 128);
}

// Shared memory left for the resident image(s) beside the largest slot area any launch of this job uses.
long long image_budget(gklb_engine* e) {
  uint32_t worst_slots = 0;
  for (auto& c : e->classes) {
    worst_slots = std::max(worst_slots, slot_bytes_total(c, c.kf, false));
    worst_slots = std::max(worst_slots, slot_bytes_total(c, c.kd, !e->use_double));
  }
  return (long long)kSmemMax - 8192 - worst_slots;  // 8 KB: mbarriers, ph2pr table, task prefix of multi-class launches
}

// Split one region's haplotypes into tiles whose images fit in `budget` bytes (appended to e->tiles).
int plan_tiles(gklb_engine* e, const gklb_pairhmm_batch* b, int region, long long budget) {
  int h = 0, index = 0;
  while (h < b->n_haps) {
    Tile t;
    t.region = region;
    t.index_in_region = index++;
    t.hap0 = h;
    size_t data = 0;
    while (h < b->n_haps) {
      const size_t len = (size_t)(b->hap_off[h + 1] - b->hap_off[h]);
      const size_t add = kHapLeftMargin + len + kHapRightMargin;
      const size_t header = align_up((size_t)20 * (t.n + 1), 16);  // the pair image's header is the larger one
      if (t.n > 0 && (long long)(header + data + add + 16) > budget) break;
      if (t.n == 0 && (long long)(header + add + 16) > budget)
        return fail(GKLB_ERR_INVALID, "haplotype %d (%zu bases) does not fit in shared memory", h, len);
      data += add;
      t.max_len = std::max(t.max_len, (int)len);
      t.n++;
      h++;
    }
    t.bytes = (uint32_t)align_up(align_up((size_t)8 * t.n, 16) + data, 16);
    // pair image: haplotypes by decreasing length, adjacent ones share a byte column
    t.order.resize(t.n);
    for (int i = 0; i < t.n; i++) t.order[i] = t.hap0 + i;
    std::stable_sort(t.order.begin(), t.order.end(), [&](int x, int y) {
      return b->hap_off[x + 1] - b->hap_off[x] > b->hap_off[y + 1] - b->hap_off[y];
    });
    t.n_pairs = (t.n + 1) / 2;
    size_t pdata = 0;
    for (int q = 0; q < t.n_pairs; q++) {
      const int a = t.order[2 * q];
      pdata += kHapLeftMargin + (size_t)(b->hap_off[a + 1] - b->hap_off[a]) + kHapRightMargin;
    }
    t.pbytes = (uint32_t)align_up(align_up((size_t)20 * t.n_pairs, 16) + pdata, 16);
    e->tiles.push_back(std::move(t));
  }
  return GKLB_OK;
}

int classes_of_region(const gklb_engine* e, int region) {
  int n = 0;
  for (auto& c : e->classes) n += (c.region == region);
  return n;
}

// Pack consecutive tiles into launch groups (their images resident together), create the (class, tile) entries and lay
// everything out in the meta block.
void plan_groups(gklb_engine* e, long long budget, size_t* meta_bytes) {
  e->groups.clear();
  e->entries.clear();
  const int max_entries = kMaxMegaClasses;
  size_t i = 0;
  while (i < e->tiles.size()) {
    Group g;
    g.tile0 = (int)i;
    g.entry0 = (int)e->entries.size();
    while (i < e->tiles.size()) {
      Tile& t = e->tiles[i];
      const int ent = classes_of_region(e, t.region);
      if (g.n_tiles > 0 && ((long long)(g.bytes + t.bytes) > budget || (long long)(g.pbytes + t.pbytes) > budget ||
                            g.n_entries + ent > max_entries))
        break;
      t.img_off = g.bytes;
      t.pimg_off = g.pbytes;
      g.bytes += t.bytes;    // multiples of 16: every image stays 16-byte aligned for the bulk copy
      g.pbytes += t.pbytes;
      for (size_t ci = 0; ci < e->classes.size(); ci++)
        if (e->classes[ci].region == t.region) {
          EntryInst en;
          en.cls = (int)ci;
          en.tile = (int)i;
          e->entries.push_back(en);
        }
      g.n_entries += ent;
      g.n_tiles++;
      i++;
    }
    // longest classes first: their tasks are the most expensive
    std::stable_sort(e->entries.begin() + g.entry0, e->entries.end(), [&](const EntryInst& x, const EntryInst& y) {
      return e->classes[x.cls].rows > e->classes[y.cls].rows;
    });
    g.meta_off = *meta_bytes;
    *meta_bytes += align_up(g.bytes, 128);
    g.pmeta_off = *meta_bytes;
    *meta_bytes += align_up(g.pbytes, 128);
    for (int k = 0; k < g.n_tiles; k++) {
      Tile& t = e->tiles[g.tile0 + k];
      t.meta_off = g.meta_off + t.img_off;
      t.pmeta_off = g.pmeta_off + t.pimg_off;
    }
    const size_t n = (size_t)g.n_entries;
    auto take = [&](size_t bytes) { const size_t o = *meta_bytes; *meta_bytes += align_up(bytes, 128); return o; };
    g.h2_cls_off = take(sizeof(H2Class) * n);
    g.h2_cfg_off = take(sizeof(int) * n);
    g.h2_end_off = take(sizeof(int) * n);
    g.r2_cls_off = take(sizeof(R2Class) * n);
    g.r2_cfg_off = take(sizeof(int) * n);
    g.dl_cls_off = take(sizeof(SweepParams) * n);
    g.dl_cfg_off = take(sizeof(int) * n);
    g.dt_cls_off = take(sizeof(SweepParams) * n);
    g.dt_cfg_off = take(sizeof(int) * n);
    g.dt_end_off = take(sizeof(int) * n);
    e->groups.push_back(g);
  }
}

// byte -> panel byte / symbol index (ConvertChar, pairhmm_common.h:49-67), tabulated: the images are built per call
struct BaseLut {
  uint8_t panel[256], index[256];
  BaseLut() {
    for (int i = 0; i < 256; i++) { panel[i] = panel_byte((uint8_t)i); index[i] = base_index((uint8_t)i); }
  }
};
const BaseLut& base_lut() {
  static const BaseLut lut;
  return lut;
}

// fill = false: header (positions, lengths) only; the bases are written on the device later
void build_tile_image(const Tile& t, const gklb_pairhmm_batch* b, uint8_t* img, bool fill) {
  const uint8_t* lut = base_lut().panel;
  memset(img, 0, t.bytes);
  int32_t* hpos = reinterpret_cast<int32_t*>(img);
  int32_t* hlen = hpos + t.n;
  size_t off = align_up((size_t)8 * t.n, 16);
  for (int i = 0; i < t.n; i++) {
    const int h = t.hap0 + i;
    const int64_t o = b->hap_off[h];
    const int len = (int)(b->hap_off[h + 1] - o);
    hpos[i] = (int32_t)(off + kHapLeftMargin - 1);  // column 0; column c is at hpos + c
    hlen[i] = len;
    uint8_t* dst = img + off + kHapLeftMargin;
    const uint8_t* src = b->hap_bases + o;
    for (int c = 0; fill && c < len; c++) dst[c] = lut[src[c]];
    off += kHapLeftMargin + len + kHapRightMargin;
  }
}

void build_pair_image(const Tile& t, const gklb_pairhmm_batch* b, uint8_t* img, bool fill) {
  const uint8_t* lut = base_lut().index;
  memset(img, 0, t.pbytes);
  int32_t* ppos = reinterpret_cast<int32_t*>(img);
  int32_t* lenA = ppos + t.n_pairs;
  int32_t* lenB = lenA + t.n_pairs;
  int32_t* idxA = lenB + t.n_pairs;
  int32_t* idxB = idxA + t.n_pairs;
  size_t off = align_up((size_t)20 * t.n_pairs, 16);
  for (int q = 0; q < t.n_pairs; q++) {
    const int a = t.order[2 * q];
    const bool has_b = 2 * q + 1 < t.n;
    const int bb = has_b ? t.order[2 * q + 1] : a;  // an odd haplotype out is paired with itself, result B dropped
    const int64_t oa = b->hap_off[a], ob = b->hap_off[bb];
    const int la = (int)(b->hap_off[a + 1] - oa), lb = (int)(b->hap_off[bb + 1] - ob);
    ppos[q] = (int32_t)(off + kHapLeftMargin - 1);
    lenA[q] = la;
    lenB[q] = lb;
    idxA[q] = a;
    idxB[q] = has_b ? bb : -1;
    uint8_t* dst = img + off + kHapLeftMargin;
    const uint8_t *sa = b->hap_bases + oa, *sb = b->hap_bases + ob;
    for (int c = 0; fill && c < lb; c++) dst[c] = (uint8_t)(lut[sa[c]] | (lut[sb[c]] << 3));
    for (int c = lb; fill && c < la; c++) dst[c] = lut[sa[c]];
    off += kHapLeftMargin + la + kHapRightMargin;
  }
}

// The range-extended fp32 rerun (pairhmm_r2.cuh) is built and tested but measured slower than the fp64 rerun it
// would replace (profiles/r2_rerun_r2_vs_fp64.json), so it is opt-in: GKLB_R2=1.
bool use_r2_env() {
  static const bool v = [] {
    const char* s = getenv("GKLB_R2");
    return s && atoi(s) == 1;
  }();
  return v;
}

int tasks_per_warp_target() {
  // aim at ~this many tasks per resident warp: enough for the dynamic queue to balance the tail, few enough that
  // the per-task constant set-up stays negligible (GKLB_TASKS_PER_WARP overrides, for measurement)
  static const int v = [] {
    const char* s = getenv("GKLB_TASKS_PER_WARP");
    return s && atoi(s) > 0 ? atoi(s) : 32;
  }();
  return v;
}

// ---- launch plan ---------------------------------------------------------------------------------------------

// The region's likelihoods, biased by the region's first read so that kernels index rows by job-wide read index.
double* region_out(gklb_engine* e, const Region& r) {
  return static_cast<double*>(e->d_out.p) + r.out_base - (int64_t)r.read_base * r.n_haps;
}

struct Sizing { int grid; size_t smem; uint32_t slot_bytes; };

// Parameters of one fp64 / packed-read (class, tile, kernel) combination.  resident: the panel is part of a group
// image (multi-class launch); otherwise the launch loads the tile's own image.  work_scale: how many warp slots the
// launch can count on for this entry (the whole GPU for a single-class launch, a share of it in a group).
Sizing fill_params(gklb_engine* e, const EntryInst& en, const KernelEntry* k, bool list_mode, bool resident,
                   long long group_units, SweepParams* out) {
  const ClassInst& c = e->classes[en.cls];
  const Tile& t = e->tiles[en.tile];
  const bool dbl = (k->policy == POL_D1);
  const HostTables& ht = host_tables();
  const Region& reg = e->regions[c.region];
  SweepParams& p = *out;
  memset(&p, 0, sizeof(p));
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  p.panel.image = dm + t.meta_off;
  p.panel.bytes = t.bytes;
  p.panel.n_haps = t.n;
  p.panel.hap0 = t.hap0;
  p.panel.n_haps_total = reg.n_haps;
  p.panel.max_hap_len = t.max_len;
  p.panel.smem_off = resident ? t.img_off : 0u;
  p.cls.records = static_cast<const uint8_t*>(e->d_records.p) + c.rec_off;
  p.cls.rec_rid = reinterpret_cast<const int32_t*>(dm + c.meta_rid);
  p.cls.rec_len = reinterpret_cast<const int32_t*>(dm + c.meta_len);
  p.cls.n_rec = c.n_rec;
  p.cls.rows = c.rows;
  p.cls.stride = c.stride;
  p.cls.n_pass = c.n_pass;
  p.ph2pr = dbl ? (const void*)e->d_ph2pr_d : (const void*)e->d_ph2pr_f;
  p.mm = dbl ? (const void*)e->d_mm_d : (const void*)e->d_mm_f;
  p.out = region_out(e, reg);
  unsigned int* counters = static_cast<unsigned int*>(e->d_counters.p);
  p.fb_count = counters + en.counter0 + 1;
  p.fb_items = e->d_fb.p ? static_cast<uint2*>(e->d_fb.p) + en.fb_off : nullptr;
  p.task_counter = counters + en.counter0 + (list_mode ? 5 : 3);
  p.carry = c.multi ? static_cast<uint8_t*>(e->d_carry.p) + c.carry_off : nullptr;
  p.carry_stride_bytes = c.carry_stride;
  p.init_const = dbl ? ht.init_d : (double)ht.init_f;
  p.log10_init = dbl ? ht.log10_init_d : (double)ht.log10_init_f;
  Sizing z;
  z.slot_bytes = warp_slot_bytes(c, k, list_mode);
  if (!list_mode) {
    const int rpw = (32 / k->G) * k->nr;
    const int n_blocks = c.n_rec / rpw;
    const int slots = e->num_sms * k->warps;
    const long long units = std::max<long long>(group_units, (long long)n_blocks * t.n);
    long long chunk = units / ((long long)tasks_per_warp_target() * slots);
    chunk = std::max(1LL, std::min(chunk, 32LL));
    if (c.multi) chunk = 1;
    chunk = std::min<long long>(chunk, t.n);
    p.hap_chunk = (int)chunk;
    p.n_chunks = (t.n + p.hap_chunk - 1) / p.hap_chunk;
    p.n_tasks = n_blocks * p.n_chunks;
    z.grid = std::min(e->num_sms, (p.n_tasks + k->warps - 1) / k->warps);
    z.smem = smem_layout(k->warps, t.bytes, z.slot_bytes, dbl ? 8 : 4).total;
  } else {
    p.list_items = p.fb_items;
    p.list_count = p.fb_count;
    z.grid = e->num_sms;
    z.smem = smem_layout(k->warps, t.bytes, z.slot_bytes, 8).total;
  }
  p.slot_bytes = z.slot_bytes;
  return z;
}

void fill_h2_common(gklb_engine* e, const uint8_t* image, uint32_t bytes, H2Common* com) {
  const HostTables& ht = host_tables();
  memset(com, 0, sizeof(*com));
  com->image = image;
  com->image_bytes = bytes;
  com->ph2pr = e->d_ph2pr_f;
  com->mm = e->d_mm_f;
  com->init_const = ht.init_f;
  com->log10_init = ht.log10_init_f;
}

void fill_h2_class(gklb_engine* e, const EntryInst& en, int warps, bool resident, long long group_units, H2Class* cls) {
  const ClassInst& c = e->classes[en.cls];
  const Tile& t = e->tiles[en.tile];
  memset(cls, 0, sizeof(*cls));
  const Region& reg = e->regions[c.region];
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  cls->records = static_cast<const uint8_t*>(e->d_records.p) + c.rec_off;
  cls->rec_rid = reinterpret_cast<const int32_t*>(dm + c.meta_rid);
  cls->rec_len = reinterpret_cast<const int32_t*>(dm + c.meta_len);
  unsigned int* counters = static_cast<unsigned int*>(e->d_counters.p);
  cls->r2_count = counters + en.counter0;
  cls->r2_items = static_cast<uint2*>(e->d_r2.p) + en.r2_off;
  cls->fb_pairs = counters + en.counter0 + 2;
  cls->fb_count = counters + en.counter0 + 1;
  cls->fb_items = e->d_fb.p ? static_cast<uint2*>(e->d_fb.p) + en.fb_off : nullptr;
  cls->use_r2 = e->r2_on ? 1 : 0;
  cls->n_rec = c.n_rec;
  cls->rows = c.rows;
  cls->stride = c.stride;
  cls->panel_off = resident ? t.pimg_off : 0u;
  cls->n_pairs = t.n_pairs;
  cls->n_haps_total = reg.n_haps;
  cls->out = region_out(e, reg);
  const int n_blocks = c.n_rec / (32 / c.G);
  const int slots = e->num_sms * warps;
  const long long units = std::max<long long>(group_units, (long long)n_blocks * t.n_pairs);
  long long chunk = units / ((long long)tasks_per_warp_target() * slots);
  chunk = std::max(1LL, std::min(chunk, 32LL));
  chunk = std::min<long long>(chunk, t.n_pairs);
  cls->pair_chunk = (int)chunk;
  cls->n_chunks = (t.n_pairs + cls->pair_chunk - 1) / cls->pair_chunk;
  cls->n_tasks = n_blocks * cls->n_chunks;
}

template <class T>
void set_params(Launch& l, const T& p) {
  l.params.resize(sizeof(T));
  memcpy(l.params.data(), &p, sizeof(T));
}

int push_launch(gklb_engine* e, Launch&& l) {
  if (l.smem > (size_t)kSmemMax) return fail(GKLB_ERR_STATE, "shared memory plan exceeds the device limit (%zu)", l.smem);
  if (l.grid > 0) e->plan.push_back(std::move(l));
  return GKLB_OK;
}

void fill_r2_class(gklb_engine* e, const EntryInst& en, bool resident, R2Class* cls) {
  const ClassInst& c = e->classes[en.cls];
  const Tile& t = e->tiles[en.tile];
  const Region& reg = e->regions[c.region];
  const bool force_fp64 = false;
  memset(cls, 0, sizeof(*cls));
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  unsigned int* counters = static_cast<unsigned int*>(e->d_counters.p);
  cls->records = static_cast<const uint8_t*>(e->d_records.p) + c.rec_off;
  cls->rec_rid = reinterpret_cast<const int32_t*>(dm + c.meta_rid);
  cls->rec_len = reinterpret_cast<const int32_t*>(dm + c.meta_len);
  cls->items = static_cast<const uint2*>(e->d_r2.p) + en.r2_off;
  cls->n_items = counters + en.counter0;
  cls->fb_items = static_cast<uint2*>(e->d_fb.p) + en.fb_off;
  cls->fb_count = counters + en.counter0 + 1;
  cls->rows = c.rows;
  cls->stride = c.stride;
  cls->panel_off = resident ? t.pimg_off : 0u;
  cls->n_pairs = t.n_pairs;
  cls->n_haps_total = reg.n_haps;
  cls->out = region_out(e, reg);
  cls->force_fp64 = force_fp64 ? 1 : 0;
  cls->debug_flags = getenv("GKLB_R2_DEBUG") ? atoi(getenv("GKLB_R2_DEBUG")) : 0;
}

uint32_t r2_slot_bytes(const ClassInst& c) { return (uint32_t)(kPriorSyms * c.K * 32 * 4); }

Launch h2_single_launch(gklb_engine* e, const EntryInst& en) {
  const ClassInst& c = e->classes[en.cls];
  const Tile& t = e->tiles[en.tile];
  H2Params p;
  memset(&p, 0, sizeof(p));
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  fill_h2_common(e, dm + t.pmeta_off, t.pbytes, &p.com);
  fill_h2_class(e, en, c.kf->warps, false, 0, &p.cls);
  p.com.slot_bytes = warp_slot_bytes(c, c.kf, false);
  p.task_counter = static_cast<unsigned int*>(e->d_counters.p) + en.counter0 + 3;
  Launch l;
  l.fn = c.kf->fn_tasks;
  l.threads = c.kf->warps * 32;
  l.smem = smem_layout(c.kf->warps, t.pbytes, p.com.slot_bytes, 4).total;
  l.grid = std::min(e->num_sms, (p.cls.n_tasks + c.kf->warps - 1) / c.kf->warps);
  l.sweep = true;
  set_params(l, p);
  snprintf(e->sweep_kernel, sizeof(e->sweep_kernel), "k_h2_tasks<%d,%d,%d>", c.G, c.K, c.kf->warps);
  return l;
}

Launch r2_single_launch(gklb_engine* e, const EntryInst& en) {
  const ClassInst& c = e->classes[en.cls];
  const Tile& t = e->tiles[en.tile];
  R2Params p;
  memset(&p, 0, sizeof(p));
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  fill_h2_common(e, dm + t.pmeta_off, t.pbytes, &p.com);
  fill_r2_class(e, en, false, &p.cls);
  p.com.slot_bytes = r2_slot_bytes(c);
  p.item_counter = static_cast<unsigned int*>(e->d_counters.p) + en.counter0 + 4;
  Launch l;
  l.fn = r2_kernel(c.G, c.K);
  l.threads = kH2Warps * 32;
  l.smem = smem_layout(kH2Warps, t.pbytes, p.com.slot_bytes, 4).total;
  l.grid = e->num_sms;
  set_params(l, p);
  return l;
}


// Build the launches of the staged job.  hm: host image of the meta block (the device-resident class arrays of the
// multi-class kernels are written into it before it is uploaded).
int build_plan(gklb_engine* e, uint8_t* hm) {
  e->plan.clear();
  e->sweep_kernel[0] = 0;
  const char* mg = getenv("GKLB_MEGA");
  const bool mega_ok = !e->forced && (mg ? atoi(mg) != 0 : true);
  unsigned int* counters = static_cast<unsigned int*>(e->d_counters.p);
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  int rc;
  for (size_t gi = 0; gi < e->groups.size(); gi++) {
    const Group& g = e->groups[gi];
    const EntryInst* ents = e->entries.data() + g.entry0;
    const int n_ent = g.n_entries;
    unsigned int* queue = counters + e->mega_counter0 + 3 * gi;
    auto cls_of = [&](const EntryInst& en) -> const ClassInst& { return e->classes[en.cls]; };
    auto tile_of = [&](const EntryInst& en) -> const Tile& { return e->tiles[en.tile]; };

    if (e->use_double) {  // ---- fp64 over all pairs ----
      if (mega_ok && n_ent > 1) {
        long long units = 0;
        for (int i = 0; i < n_ent; i++) {
          const ClassInst& c = cls_of(ents[i]);
          units += (long long)(c.n_rec / ((32 / c.kd->G) * c.kd->nr)) * tile_of(ents[i]).n;
        }
        MegaParams mp;
        memset(&mp, 0, sizeof(mp));
        SweepParams* arr = reinterpret_cast<SweepParams*>(hm + g.dt_cls_off);
        int* cfg = reinterpret_cast<int*>(hm + g.dt_cfg_off);
        int* ends = reinterpret_cast<int*>(hm + g.dt_end_off);
        uint32_t slot_bytes = 0;
        int tasks = 0;
        for (int i = 0; i < n_ent; i++) {
          const ClassInst& c = cls_of(ents[i]);
          const Sizing z = fill_params(e, ents[i], c.kd, false, true, units, &arr[i]);
          cfg[i] = c.cfg_d;
          tasks += arr[i].n_tasks;
          ends[i] = tasks;
          slot_bytes = std::max(slot_bytes, z.slot_bytes);
        }
        mp.n_classes = n_ent;
        mp.cfg = reinterpret_cast<const int*>(dm + g.dt_cfg_off);
        mp.task_end = reinterpret_cast<const int*>(dm + g.dt_end_off);
        mp.cls = reinterpret_cast<const SweepParams*>(dm + g.dt_cls_off);
        mp.queue = queue;
        mp.image = dm + g.meta_off;
        mp.image_bytes = g.bytes;
        Launch l;
        l.fn = mega_kernel(POL_D1, 0);
        l.threads = 8 * 32;
        l.smem = smem_layout(8, g.bytes, slot_bytes, 8, (uint32_t)n_ent).total;
        l.grid = std::min(e->num_sms, (tasks + 7) / 8);
        l.extra = slot_bytes;
        l.sweep = true;
        set_params(l, mp);
        snprintf(e->sweep_kernel, sizeof(e->sweep_kernel), "k_mega_tasks<VD1,8> (%d class x tile entries)", n_ent);
        if ((rc = push_launch(e, std::move(l)))) return rc;
      } else {
        for (int i = 0; i < n_ent; i++) {
          const ClassInst& c = cls_of(ents[i]);
          SweepParams p;
          const Sizing z = fill_params(e, ents[i], c.kd, false, false, 0, &p);
          Launch l;
          l.fn = c.kd->fn_tasks;
          l.threads = c.kd->warps * 32;
          l.smem = z.smem;
          l.grid = z.grid;
          l.sweep = true;
          set_params(l, p);
          snprintf(e->sweep_kernel, sizeof(e->sweep_kernel), "k_sweep_tasks<VD1,%d,%d,%d,%s>", c.kd->G, c.kd->K, c.kd->warps,
                   c.multi ? "multi" : "single");
          if ((rc = push_launch(e, std::move(l)))) return rc;
        }
      }
      continue;
    }

    // ---- fp32 forward sweep ----
    std::vector<int> h2, other;  // indices into ents
    for (int i = 0; i < n_ent; i++) (cls_of(ents[i]).kf->policy == POL_H2 && !e->forced ? h2 : other).push_back(i);
    const bool h2_mega = mega_ok && h2.size() > 1;
    if (h2_mega) {
      long long units = 0;
      for (int i : h2) units += (long long)(cls_of(ents[i]).n_rec / (32 / cls_of(ents[i]).G)) * tile_of(ents[i]).n_pairs;
      H2MegaParams mp;
      memset(&mp, 0, sizeof(mp));
      fill_h2_common(e, dm + g.pmeta_off, g.pbytes, &mp.com);
      H2Class* arr = reinterpret_cast<H2Class*>(hm + g.h2_cls_off);
      int* cfg = reinterpret_cast<int*>(hm + g.h2_cfg_off);
      int* ends = reinterpret_cast<int*>(hm + g.h2_end_off);
      int tasks = 0;
      uint32_t slot_bytes = 0;
      for (int i : h2) {
        const ClassInst& c = cls_of(ents[i]);
        const int k = mp.n_classes++;
        fill_h2_class(e, ents[i], kH2Warps, true, units, &arr[k]);
        cfg[k] = c.cfg_f;
        tasks += arr[k].n_tasks;
        ends[k] = tasks;
        slot_bytes = std::max(slot_bytes, warp_slot_bytes(c, c.kf, false));
      }
      mp.com.slot_bytes = slot_bytes;
      mp.cfg = reinterpret_cast<const int*>(dm + g.h2_cfg_off);
      mp.task_end = reinterpret_cast<const int*>(dm + g.h2_end_off);
      mp.cls = reinterpret_cast<const H2Class*>(dm + g.h2_cls_off);
      mp.queue = queue;
      Launch l;
      l.fn = h2_mega_kernel();
      l.threads = kH2Warps * 32;
      l.smem = smem_layout(kH2Warps, g.pbytes, slot_bytes, 4, (uint32_t)mp.n_classes).total;
      l.grid = std::min(e->num_sms, (tasks + kH2Warps - 1) / kH2Warps);
      l.sweep = true;
      set_params(l, mp);
      snprintf(e->sweep_kernel, sizeof(e->sweep_kernel), "k_h2_mega<8> (%d class x tile entries, %d tiles)", mp.n_classes,
               g.n_tiles);
      if ((rc = push_launch(e, std::move(l)))) return rc;
    } else {
      for (int i : h2)
        if ((rc = push_launch(e, h2_single_launch(e, ents[i])))) return rc;
    }
    for (int i : other) {  // multi-pass classes and forced measurement kernels: one launch each
      const ClassInst& c = cls_of(ents[i]);
      if (c.kf->policy == POL_H2) {
        if ((rc = push_launch(e, h2_single_launch(e, ents[i])))) return rc;
        continue;
      }
      SweepParams p;
      const Sizing z = fill_params(e, ents[i], c.kf, false, false, 0, &p);
      Launch l;
      l.fn = c.kf->fn_tasks;
      l.threads = c.kf->warps * 32;
      l.smem = z.smem;
      l.grid = z.grid;
      l.sweep = true;
      set_params(l, p);
      if (!e->sweep_kernel[0] || e->forced)
        snprintf(e->sweep_kernel, sizeof(e->sweep_kernel), "k_sweep_tasks<pol%d,%d,%d,%d,%s,var%d>", c.kf->policy, c.G, c.K,
                 c.kf->warps, c.multi ? "multi" : "single", c.kf->var);
      if ((rc = push_launch(e, std::move(l)))) return rc;
    }

    // ---- range-extended fp32 rerun of the pairs the H2 sweep flagged ----
    std::vector<int> h2all;  // every H2 entry, forced ones included (their lists have the same format)
    for (int i = 0; i < n_ent; i++)
      if (cls_of(ents[i]).kf->policy == POL_H2) h2all.push_back(i);
    if (!e->r2_on) {
      // the H2 sweep appended its flagged pairs straight to the fp64 lists
    } else if (h2_mega) {
      R2MegaParams mp;
      memset(&mp, 0, sizeof(mp));
      fill_h2_common(e, dm + g.pmeta_off, g.pbytes, &mp.com);
      R2Class* arr = reinterpret_cast<R2Class*>(hm + g.r2_cls_off);
      int* cfg = reinterpret_cast<int*>(hm + g.r2_cfg_off);
      uint32_t slot_bytes = 0;
      for (int i : h2all) {
        const ClassInst& c = cls_of(ents[i]);
        const int k = mp.n_classes++;
        fill_r2_class(e, ents[i], true, &arr[k]);
        cfg[k] = c.cfg_f;
        slot_bytes = std::max(slot_bytes, r2_slot_bytes(c));
      }
      mp.com.slot_bytes = slot_bytes;
      mp.cfg = reinterpret_cast<const int*>(dm + g.r2_cfg_off);
      mp.cls = reinterpret_cast<const R2Class*>(dm + g.r2_cls_off);
      mp.queue = queue + 1;
      Launch l;
      l.fn = r2_mega_kernel();
      l.threads = kH2Warps * 32;
      l.smem = smem_layout(kH2Warps, g.pbytes, slot_bytes, 4, (uint32_t)mp.n_classes).total;
      l.grid = e->num_sms;
      set_params(l, mp);
      if ((rc = push_launch(e, std::move(l)))) return rc;
    } else {
      for (int i : h2all)
        if ((rc = push_launch(e, r2_single_launch(e, ents[i])))) return rc;
    }

    // ---- fp64: what the rerun's guards passed on, and the flagged pairs of the multi-pass classes ----
    if (mega_ok && n_ent > 1) {
      MegaParams mp;
      memset(&mp, 0, sizeof(mp));
      SweepParams* arr = reinterpret_cast<SweepParams*>(hm + g.dl_cls_off);
      int* cfg = reinterpret_cast<int*>(hm + g.dl_cfg_off);
      uint32_t slot_bytes = 0;
      for (int i = 0; i < n_ent; i++) {
        const ClassInst& c = cls_of(ents[i]);
        const Sizing z = fill_params(e, ents[i], c.kd, true, true, 0, &arr[i]);
        cfg[i] = c.cfg_d;
        slot_bytes = std::max(slot_bytes, z.slot_bytes);
      }
      mp.n_classes = n_ent;
      mp.cfg = reinterpret_cast<const int*>(dm + g.dl_cfg_off);
      mp.cls = reinterpret_cast<const SweepParams*>(dm + g.dl_cls_off);
      mp.queue = queue + 2;
      mp.image = dm + g.meta_off;
      mp.image_bytes = g.bytes;
      Launch l;
      l.fn = mega_kernel(POL_D1, 1);
      l.threads = 8 * 32;
      l.smem = smem_layout(8, g.bytes, slot_bytes, 8, (uint32_t)n_ent).total;
      l.grid = e->num_sms;
      l.extra = slot_bytes;
      set_params(l, mp);
      if ((rc = push_launch(e, std::move(l)))) return rc;
    } else {
      for (int i = 0; i < n_ent; i++) {
        const ClassInst& c = cls_of(ents[i]);
        if (c.kf->policy == POL_D1) continue;  // forced fp64 sweep: nothing to rerun
        SweepParams p;
        const Sizing z = fill_params(e, ents[i], c.kd, true, false, 0, &p);
        Launch l;
        l.fn = c.kd->fn_list;
        l.threads = c.kd->warps * 32;
        l.smem = z.smem;
        l.grid = z.grid;
        set_params(l, p);
        if ((rc = push_launch(e, std::move(l)))) return rc;
      }
    }
  }
  return GKLB_OK;
}

}  // namespace

bool use_r2(const gklb_engine* e) { return e->r2_on; }

void read_fallback_count(gklb_engine* e) {
  int64_t fb = 0, f64 = 0;
  const unsigned int* hc = static_cast<const unsigned int*>(e->h_counters.p);
  for (auto& en : e->entries) {
    const bool h2 = e->classes[en.cls].kf->policy == POL_H2;
    fb += h2 ? hc[en.counter0 + 2] : hc[en.counter0 + 1];
    f64 += hc[en.counter0 + 1];
  }
  e->stats.fallback_pairs = fb;
  e->stats.fp64_pairs = f64;
}

int validate_batch(const gklb_pairhmm_batch* b) {
  if (!b) return fail(GKLB_ERR_INVALID, "batch is null");
  if (b->n_reads < 0 || b->n_haps < 0) return fail(GKLB_ERR_INVALID, "negative batch size");
  if (b->n_reads == 0 || b->n_haps == 0) return GKLB_OK;
  if (!b->read_off || !b->hap_off || !b->read_bases || !b->read_quals || !b->ins_gop || !b->del_gop || !b->gcp ||
      !b->hap_bases)
    return fail(GKLB_ERR_INVALID, "null pointer in batch");
  if (b->read_off[0] != 0 || b->hap_off[0] != 0) return fail(GKLB_ERR_INVALID, "offsets must start at 0");
  for (int r = 0; r < b->n_reads; r++)
    if (b->read_off[r + 1] <= b->read_off[r]) return fail(GKLB_ERR_INVALID, "read %d is empty or offsets decrease", r);
  for (int h = 0; h < b->n_haps; h++)
    if (b->hap_off[h + 1] <= b->hap_off[h]) return fail(GKLB_ERR_INVALID, "haplotype %d is empty or offsets decrease", h);
  return GKLB_OK;
}

namespace {
struct StageTimer {   // GKLB_STAGE_TIMING=1: host time of the phases of do_stage, to stderr
  bool on;
  std::chrono::steady_clock::time_point t0;
  double acc[8] = {0};
  StageTimer() : on(getenv("GKLB_STAGE_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(int i) {
    if (!on) return;
    const auto t = std::chrono::steady_clock::now();
    acc[i] += std::chrono::duration<double, std::milli>(t - t0).count();
    t0 = t;
  }
  ~StageTimer() {
    if (on)
      fprintf(stderr, "[gklb stage] validate %.3f classes %.3f tiles+groups %.3f alloc %.3f images %.3f plan %.3f copy %.3f pack %.3f ms\n",
              acc[0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6], acc[7]);
  }
};
}  // namespace

int do_stage(gklb_engine* e, const gklb_pairhmm_batch* batches, int k, bool hap_on_device) {
  StageTimer tm;
  if (!e->pending_out.empty())
    return fail(GKLB_ERR_STATE, "a submitted batch is still in flight on this engine: call gklb_engine_wait first");
  e->staged = false;
  if (k < 0 || (k > 0 && !batches)) return fail(GKLB_ERR_INVALID, "bad region array");
  if (hap_on_device && k != 1) return fail(GKLB_ERR_INVALID, "device-resident staging takes one region");
  int rc;
  for (int r = 0; r < k; r++)
    if ((rc = validate_batch(&batches[r]))) return rc;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));  // the pinned staging buffers are about to be rewritten
  e->stats = gklb_pairhmm_stats{};
  e->plan.clear();
  e->classes.clear();
  e->tiles.clear();
  e->groups.clear();
  e->regions.assign((size_t)k, Region{});
  // regions, in one job-wide numbering of reads, haplotypes and result slots; empty regions keep their slot
  int64_t total_read = 0, total_hap = 0, total_pairs = 0, n_reads_all = 0, n_haps_all = 0;
  for (int r = 0; r < k; r++) {
    const gklb_pairhmm_batch& b = batches[r];
    Region& reg = e->regions[r];
    const bool empty = b.n_reads == 0 || b.n_haps == 0;
    reg.n_reads = empty ? 0 : b.n_reads;
    reg.n_haps = empty ? 0 : b.n_haps;
    reg.read_base = (int)n_reads_all;
    reg.arena_base = total_read;
    reg.hap_base = (int)n_haps_all;
    reg.hap_arena_base = total_hap;
    reg.out_base = total_pairs;
    if (empty) continue;
    n_reads_all += b.n_reads;
    n_haps_all += b.n_haps;
    total_read += b.read_off[b.n_reads];
    total_hap += b.hap_off[b.n_haps];
    total_pairs += (int64_t)b.n_reads * b.n_haps;
    e->stats.cells += b.read_off[b.n_reads] * b.hap_off[b.n_haps];
    if (n_reads_all > 0x7fffffff || n_haps_all > 0x7fffffff) return fail(GKLB_ERR_INVALID, "job too large");
  }
  tm.lap(0);
  e->stats.pairs = total_pairs;
  if (total_pairs == 0) {
    e->staged = true;
    return GKLB_OK;
  }
  if ((rc = parse_forced(e))) return rc;

  // the haplotype panel images are built on the host: fetch the bases if they live on the device
  std::vector<uint8_t> hap_host;
  std::vector<gklb_pairhmm_batch> hb(batches, batches + k);
  if (hap_on_device) {
    hap_host.resize((size_t)total_hap);
    CU(cudaMemcpyAsync(hap_host.data(), batches[0].hap_bases, (size_t)total_hap, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    hb[0].hap_bases = hap_host.data();
  }

  for (int r = 0; r < k; r++)
    if (e->regions[r].n_reads && (rc = plan_classes(e, &batches[r], r, e->regions[r].read_base))) return rc;
  tm.lap(1);
  e->r2_on = use_r2_env();
  for (auto& c : e->classes)
    if (c.kf && c.kf->policy == POL_H2 && c.G > 16) e->r2_on = false;   // no range-extended kernels for 32 lanes
  const long long budget = image_budget(e);
  for (int r = 0; r < k; r++)
    if (e->regions[r].n_reads && (rc = plan_tiles(e, &hb[r], r, budget))) return rc;
  size_t meta_bytes = 0;
  plan_groups(e, budget, &meta_bytes);
  tm.lap(2);

  size_t rec_bytes = 0, fb_items = 0, r2_items = 0, carry_bytes = 0;
  int counters = 0;
  for (auto& en : e->entries) {
    const ClassInst& c = e->classes[en.cls];
    const Tile& t = e->tiles[en.tile];
    en.fb_off = fb_items;
    en.r2_off = r2_items;
    if (!e->use_double) {
      fb_items += (size_t)c.n_rec * t.n;
      if (c.kf->policy == POL_H2) r2_items += (size_t)c.n_rec * t.n_pairs;
    }
    en.counter0 = counters;
    counters += kEntryCounters;
  }
  for (auto& c : e->classes) {
    c.meta_rid = meta_bytes;
    meta_bytes += align_up(sizeof(int32_t) * c.n_rec, 128);
    c.meta_len = meta_bytes;
    meta_bytes += align_up(sizeof(int32_t) * c.n_rec, 128);
    c.rec_off = rec_bytes;
    rec_bytes += align_up((size_t)c.n_rec * 5 * c.stride, 128);
    if (c.multi) {
      int max_len = 0;
      for (auto& t : e->tiles)
        if (t.region == c.region) max_len = std::max(max_len, t.max_len);
      // per-warp scratch: sized for the widest CTA any kernel of this class may run with
      const int warps = std::max(12, std::max(c.kf->warps, c.kd->warps));
      c.carry_stride = (size_t)(32 / c.G) * 6 * (max_len + 2) * 8;
      c.carry_off = carry_bytes;
      carry_bytes += c.carry_stride * warps * e->num_sms;
    }
  }
  e->mega_counter0 = counters;
  counters += 3 * (int)e->groups.size();
  e->n_counters = counters;

  // host-resident jobs of moderate size: offsets + arenas are staged behind the meta block -> one host->device copy
  e->arena_pitch = align_up((size_t)total_read, 256);
  const bool consolidate = !hap_on_device && (k > 1 || (size_t)total_read * 5 <= (4u << 20));
  const size_t off_bytes_r = sizeof(int64_t) * ((size_t)n_reads_all + 1), off_bytes_h = sizeof(int64_t) * ((size_t)n_haps_all + 1);
  size_t st_read_off = 0, st_hap_off = 0, st_arena = 0, staged_bytes = meta_bytes;
  if (consolidate) {
    st_read_off = staged_bytes;
    staged_bytes += align_up(off_bytes_r, 256);
    st_hap_off = staged_bytes;
    staged_bytes += align_up(off_bytes_h, 256);
    st_arena = staged_bytes;
    staged_bytes += 5 * e->arena_pitch;
  }
  CU(e->h_meta.ensure(staged_bytes));
  CU(e->d_meta.ensure(staged_bytes));
  CU(e->d_records.ensure(rec_bytes));
  if (!consolidate) {
    CU(e->d_read_off.ensure(off_bytes_r));
    CU(e->d_hap_off.ensure(off_bytes_h));
    CU(e->d_arenas.ensure(e->arena_pitch * 5));
  }
  CU(e->d_out.ensure(sizeof(double) * (size_t)total_pairs));
  if (fb_items) CU(e->d_fb.ensure(sizeof(uint2) * fb_items));
  if (r2_items) CU(e->d_r2.ensure(sizeof(uint2) * r2_items));
  CU(e->d_counters.ensure(sizeof(unsigned int) * (size_t)counters));
  CU(e->h_counters.ensure(sizeof(unsigned int) * (size_t)counters));
  if (carry_bytes) CU(e->d_carry.ensure(carry_bytes));

  tm.lap(3);
  uint8_t* hm = static_cast<uint8_t*>(e->h_meta.p);
  for (auto& t : e->tiles) {
    build_tile_image(t, &hb[t.region], hm + t.meta_off, !e->defer_panel);
    build_pair_image(t, &hb[t.region], hm + t.pmeta_off, !e->defer_panel);
  }
  for (auto& c : e->classes) {
    memcpy(hm + c.meta_rid, c.rid.data(), sizeof(int32_t) * c.n_rec);
    memcpy(hm + c.meta_len, c.len.data(), sizeof(int32_t) * c.n_rec);
  }
  tm.lap(4);
  if ((rc = build_plan(e, hm))) return rc;  // every device buffer has its final address now
  tm.lap(5);

  cudaStream_t s = e->stream;
  uint8_t* dmw = static_cast<uint8_t*>(e->d_meta.p);
  if (consolidate) {
    int64_t* ro = reinterpret_cast<int64_t*>(hm + st_read_off);
    int64_t* ho = reinterpret_cast<int64_t*>(hm + st_hap_off);
    for (int r = 0; r < k; r++) {
      const Region& reg = e->regions[r];
      if (!reg.n_reads) continue;
      const gklb_pairhmm_batch& b = batches[r];
      for (int i = 0; i <= b.n_reads; i++) ro[reg.read_base + i] = reg.arena_base + b.read_off[i];
      for (int i = 0; i <= b.n_haps; i++) ho[reg.hap_base + i] = reg.hap_arena_base + b.hap_off[i];
      const uint8_t* src[5] = {b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp};
      for (int i = 0; i < 5; i++)
        memcpy(hm + st_arena + i * e->arena_pitch + reg.arena_base, src[i], (size_t)b.read_off[b.n_reads]);
    }
    CU(cudaMemcpyAsync(dmw, hm, staged_bytes, cudaMemcpyHostToDevice, s));
    e->p_read_off = reinterpret_cast<const int64_t*>(dmw + st_read_off);
    e->p_hap_off = reinterpret_cast<const int64_t*>(dmw + st_hap_off);
    e->p_arenas = dmw + st_arena;
  } else {  // one large region: the arenas go straight from the caller's buffers (host or device)
    const gklb_pairhmm_batch& b = batches[0];
    CU(cudaMemcpyAsync(dmw, hm, meta_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->d_read_off.p, b.read_off, off_bytes_r, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->d_hap_off.p, b.hap_off, off_bytes_h, cudaMemcpyHostToDevice, s));
    const uint8_t* src[5] = {b.read_bases, b.read_quals, b.ins_gop, b.del_gop, b.gcp};
    uint8_t* da = static_cast<uint8_t*>(e->d_arenas.p);
    for (int i = 0; i < 5; i++)
      CU(cudaMemcpyAsync(da + i * e->arena_pitch, src[i], (size_t)total_read, cudaMemcpyDefault, s));
    e->p_read_off = static_cast<const int64_t*>(e->d_read_off.p);
    e->p_hap_off = static_cast<const int64_t*>(e->d_hap_off.p);
    e->p_arenas = da;
  }

  tm.lap(6);
  // one packing launch for all classes (32 classes per launch)
  const uint8_t* dm = dmw;
  for (size_t c0 = 0; c0 < e->classes.size(); c0 += 32) {
    PackParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.read_off = e->p_read_off;
    pp.bases = e->p_arenas;
    pp.quals = e->p_arenas + e->arena_pitch;
    pp.ins = e->p_arenas + 2 * e->arena_pitch;
    pp.del = e->p_arenas + 3 * e->arena_pitch;
    pp.gcp = e->p_arenas + 4 * e->arena_pitch;
    for (size_t ci = c0; ci < std::min(e->classes.size(), c0 + 32); ci++) {
      const ClassInst& c = e->classes[ci];
      PackClass& pc = pp.cls[pp.n_classes++];
      pc.records = static_cast<uint8_t*>(e->d_records.p) + c.rec_off;
      pc.rec_rid = reinterpret_cast<const int32_t*>(dm + c.meta_rid);
      pc.rec_begin = pp.n_rec_total;
      pc.n_rec = c.n_rec;
      pc.rows = c.rows;
      pc.stride = c.stride;
      pp.n_rec_total += c.n_rec;
    }
    CU(launch_pack(pp, s));
  }
  tm.lap(7);
  e->staged = true;
  return GKLB_OK;
}

int do_run(gklb_engine* e) {
  if (!e->staged) return fail(GKLB_ERR_STATE, "nothing staged");
  CU(cudaSetDevice(e->device));
  e->stats.kernel_launches = 0;
  e->kev_used = 0;
  e->stats.n_classes = (int)e->classes.size();
  if (e->classes.empty()) return GKLB_OK;
  CU(cudaMemsetAsync(e->d_counters.p, 0, sizeof(unsigned int) * (size_t)e->n_counters, e->stream));
  for (const Launch& l : e->plan) {
    if (l.sweep) {
      while ((int)e->kev.size() < e->kev_used + 2) {
        cudaEvent_t ev;
        CU(cudaEventCreate(&ev));
        e->kev.push_back(ev);
      }
      CU(cudaEventRecord(e->kev[e->kev_used], e->stream));
    }
    void* args[2] = {const_cast<uint8_t*>(l.params.data()), const_cast<uint32_t*>(&l.extra)};
    CU(cudaLaunchKernel(l.fn, dim3(l.grid), dim3(l.threads), args, l.smem, e->stream));
    if (l.sweep) {
      CU(cudaEventRecord(e->kev[e->kev_used + 1], e->stream));
      e->kev_used += 2;
    }
    e->stats.kernel_launches++;
  }
  return GKLB_OK;
}

namespace {
// The job's result buffer -> the callers' arrays, one per region.
void scatter_results(gklb_engine* e, const uint8_t* host_src, double* const* outs) {
  for (size_t r = 0; r < e->regions.size(); r++) {
    const Region& reg = e->regions[r];
    const size_t n = (size_t)reg.n_reads * reg.n_haps;
    if (n) memcpy(outs[r], host_src + sizeof(double) * (size_t)reg.out_base, sizeof(double) * n);
  }
}
}  // namespace

int do_fetch(gklb_engine* e, double* const* outs) {
  if (!e->staged) return fail(GKLB_ERR_STATE, "nothing staged");
  if (!e->pending_out.empty()) return fail(GKLB_ERR_STATE, "a submitted batch is still in flight on this engine");
  CU(cudaSetDevice(e->device));
  if (e->classes.empty()) return GKLB_OK;
  if (!outs) return fail(GKLB_ERR_INVALID, "likelihoods is null");
  for (size_t r = 0; r < e->regions.size(); r++)
    if (e->regions[r].n_reads && !outs[r]) return fail(GKLB_ERR_INVALID, "likelihoods is null");
  // Small results travel through the engine's pinned buffer: a device->host copy into the caller's (usually
  // pageable) array is staged by the driver and costs tens of microseconds more per call than the memcpy here.
  const size_t out_bytes = sizeof(double) * (size_t)e->stats.pairs;
  const bool via_pinned = e->regions.size() > 1 || out_bytes <= ((size_t)2 << 20);
  if (via_pinned) {
    CU(e->h_out.ensure(out_bytes));
    CU(cudaMemcpyAsync(e->h_out.p, e->d_out.p, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  } else {
    CU(cudaMemcpyAsync(outs[0], e->d_out.p, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  }
  CU(cudaMemcpyAsync(e->h_counters.p, e->d_counters.p, sizeof(unsigned int) * (size_t)e->n_counters,
                     cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (via_pinned) scatter_results(e, static_cast<const uint8_t*>(e->h_out.p), outs);
  read_fallback_count(e);
  return GKLB_OK;
}

int do_compute(gklb_engine* e, const gklb_pairhmm_batch* batches, int k, double* const* outs) {
  int rc;
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev[0], e->stream));
  if ((rc = do_stage(e, batches, k, false))) return rc;
  CU(cudaEventRecord(e->ev[1], e->stream));
  if ((rc = do_run(e))) return rc;
  CU(cudaEventRecord(e->ev[2], e->stream));
  if ((rc = do_fetch(e, outs))) return rc;
  CU(cudaEventRecord(e->ev[3], e->stream));
  CU(cudaEventSynchronize(e->ev[3]));
  cudaEventElapsedTime(&e->stats.h2d_ms, e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&e->stats.kernel_ms, e->ev[1], e->ev[2]);
  cudaEventElapsedTime(&e->stats.d2h_ms, e->ev[2], e->ev[3]);
  return GKLB_OK;
}

// Asynchronous compute: everything is queued on the engine's stream and the likelihoods travel to a pinned
// buffer (a device->host copy into pageable memory would block the host until the kernels are done).
int do_submit(gklb_engine* e, const gklb_pairhmm_batch* batches, int k, double* const* outs) {
  int rc;
  if ((rc = do_stage(e, batches, k, false))) return rc;  // refuses while a job is in flight
  if ((rc = do_run(e))) return rc;
  if (e->classes.empty()) return GKLB_OK;
  if (!outs) return fail(GKLB_ERR_INVALID, "likelihoods is null");
  for (int r = 0; r < k; r++)
    if (e->regions[r].n_reads && !outs[r]) return fail(GKLB_ERR_INVALID, "likelihoods is null");
  CU(e->h_out.ensure(sizeof(double) * (size_t)e->stats.pairs));
  CU(cudaMemcpyAsync(e->h_out.p, e->d_out.p, sizeof(double) * (size_t)e->stats.pairs, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaMemcpyAsync(e->h_counters.p, e->d_counters.p, sizeof(unsigned int) * (size_t)e->n_counters,
                     cudaMemcpyDeviceToHost, e->stream));
  e->pending_out.assign(outs, outs + k);
  return GKLB_OK;
}

int do_wait(gklb_engine* e) {
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  if (e->pending_out.empty()) return GKLB_OK;
  scatter_results(e, static_cast<const uint8_t*>(e->h_out.p), e->pending_out.data());
  e->pending_out.clear();
  read_fallback_count(e);
  return GKLB_OK;
}

int fill_panels_from_device(gklb_engine* e, const uint8_t* hap_bases_dev) {
  if (!e->staged) return fail(GKLB_ERR_STATE, "nothing staged");
  if (e->regions.size() != 1) return fail(GKLB_ERR_STATE, "the staged job must have one region");
  CU(cudaSetDevice(e->device));
  uint8_t* dm = static_cast<uint8_t*>(e->d_meta.p);
  for (auto& t : e->tiles) {
    CU(launch_fill_panel(dm + t.meta_off, t.n, t.hap0, e->p_hap_off, hap_bases_dev, e->stream));
    CU(launch_fill_pair_panel(dm + t.pmeta_off, t.n_pairs, e->p_hap_off, hap_bases_dev, e->stream));
  }
  return GKLB_OK;
}

int create_engine(gklb_engine** out, int device, int use_double) {
  int n = 0;
  cudaError_t ce = cudaGetDeviceCount(&n);
  if (ce != cudaSuccess || n <= 0) return fail(GKLB_ERR_NO_DEVICE, "no CUDA device (%s)", cudaGetErrorString(ce));
  if (device < 0 || device >= n) return fail(GKLB_ERR_NO_DEVICE, "device %d out of range (%d devices)", device, n);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(GKLB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major,
                prop.minor);
  CU(cudaSetDevice(device));
  gklb_engine* e = new gklb_engine;
  e->device = device;
  e->use_double = use_double != 0;
  e->num_sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  e->stream = e->own_stream;
  for (auto& ev : e->ev) CU(cudaEventCreate(&ev));
  int rc = set_kernel_attrs();
  if (!rc) rc = upload_tables(e);
  if (rc) { delete e; return rc; }
  *out = e;
  return GKLB_OK;
}

void destroy_engine(gklb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  for (DevBuf* b : {&e->d_tables, &e->d_hap_off, &e->d_read_off, &e->d_arenas, &e->d_meta, &e->d_records, &e->d_out, &e->d_fb,
                    &e->d_r2, &e->d_counters, &e->d_carry, &e->d_xhap, &e->d_xf32, &e->d_xidx, &e->d_xval, &e->d_xcnt})
    b->release();
  for (HostBuf* b : {&e->h_xf32[0], &e->h_xf32[1], &e->h_xidx, &e->h_xval}) b->release();
  e->h_meta.release();
  e->h_counters.release();
  e->h_out.release();
  for (auto& ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : e->kev) cudaEventDestroy(ev);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

}  // namespace gklb

using namespace gklb;

extern "C" {

int gklb_engine_create(gklb_engine** out, int device, int use_double) {
  if (!out) return fail(GKLB_ERR_INVALID, "out is null");
  return create_engine(out, device, use_double);
}

int gklb_engine_destroy(gklb_engine* e) {
  destroy_engine(e);
  return GKLB_OK;
}

int gklb_engine_device(gklb_engine* e) { return e ? e->device : -1; }

int gklb_engine_set_stream(gklb_engine* e, void* cuda_stream) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  e->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : e->own_stream;
  return GKLB_OK;
}

int gklb_engine_compute(gklb_engine* e, const gklb_pairhmm_batch* batch, double* likelihoods) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  if (!batch) return fail(GKLB_ERR_INVALID, "batch is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_compute(e, batch, 1, &likelihoods);
}

int gklb_engine_compute_multi(gklb_engine* e, const gklb_pairhmm_batch* batches, int n_batches, double* const* likelihoods) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_compute(e, batches, n_batches, likelihoods);
}

int gklb_engine_submit(gklb_engine* e, const gklb_pairhmm_batch* batch, double* likelihoods) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  if (!batch) return fail(GKLB_ERR_INVALID, "batch is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_submit(e, batch, 1, &likelihoods);
}

int gklb_engine_wait(gklb_engine* e) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_wait(e);
}

int gklb_engine_stage(gklb_engine* e, const gklb_pairhmm_batch* batch) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  if (!batch) return fail(GKLB_ERR_INVALID, "batch is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_stage(e, batch, 1, false);
}

int gklb_engine_stage_multi(gklb_engine* e, const gklb_pairhmm_batch* batches, int n_batches) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_stage(e, batches, n_batches, false);
}

int gklb_engine_stage_device(gklb_engine* e, const gklb_pairhmm_batch* batch) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  if (!batch) return fail(GKLB_ERR_INVALID, "batch is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_stage(e, batch, 1, true);
}

int gklb_engine_update_haps_device(gklb_engine* e, const void* hap_bases_dev) {
  if (!e || !hap_bases_dev) return fail(GKLB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  return fill_panels_from_device(e, static_cast<const uint8_t*>(hap_bases_dev));
}

int gklb_engine_run(gklb_engine* e) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_run(e);
}

int gklb_engine_fetch(gklb_engine* e, double* likelihoods) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  if (e->regions.size() > 1) return fail(GKLB_ERR_STATE, "the staged job has several regions: use gklb_engine_compute_multi");
  return do_fetch(e, &likelihoods);
}

int gklb_engine_result_device(gklb_engine* e, void** dev_ptr) {
  if (!e || !dev_ptr) return fail(GKLB_ERR_INVALID, "null argument");
  if (!e->staged) return fail(GKLB_ERR_STATE, "nothing staged");
  *dev_ptr = e->d_out.p;
  return GKLB_OK;
}

int gklb_engine_synchronize(gklb_engine* e) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  return GKLB_OK;
}

int gklb_engine_stats(gklb_engine* e, gklb_pairhmm_stats* out) {
  if (!e || !out) return fail(GKLB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  e->stats.sweep_ms = 0;
  e->stats.sweep_launches = e->kev_used / 2;
  for (int i = 0; i + 1 < e->kev_used; i += 2) {
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e->kev[i], e->kev[i + 1]));
    e->stats.sweep_ms += ms;
  }
  *out = e->stats;
  return GKLB_OK;
}

const char* gklb_engine_sweep_kernel(gklb_engine* e) { return e ? e->sweep_kernel : ""; }

// Text description of the staged job's plan (regions, classes, tiles, launch groups, launches), for reports.
int gklb_engine_plan_info(gklb_engine* e, char* buf, int n) {
  if (!e || !buf || n <= 0) return fail(GKLB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  int at = snprintf(buf, n, "regions=%zu classes=%zu tiles=%zu groups=%zu launches=%zu\n", e->regions.size(),
                    e->classes.size(), e->tiles.size(), e->groups.size(), e->plan.size());
  for (size_t i = 0; i < e->groups.size() && at < n; i++)
    at += snprintf(buf + at, n - at, "group %zu: tiles=%d entries=%d image=%u pair_image=%u\n", i, e->groups[i].n_tiles,
                   e->groups[i].n_entries, e->groups[i].bytes, e->groups[i].pbytes);
  for (size_t i = 0; i < e->plan.size() && at < n; i++)
    at += snprintf(buf + at, n - at, "launch %zu: grid=%d threads=%d smem=%zu%s\n", i, e->plan[i].grid, e->plan[i].threads,
                   e->plan[i].smem, e->plan[i].sweep ? " sweep" : "");
  return GKLB_OK;
}

int gklb_engine_time_runs(gklb_engine* e, int iters, float* ms_per_run) {
  if (!e || !ms_per_run || iters <= 0) return fail(GKLB_ERR_INVALID, "bad argument");
  std::lock_guard<std::mutex> lk(e->mu);
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaEventRecord(e->ev[0], e->stream));
  for (int i = 0; i < iters; i++) {
    int rc = do_run(e);
    if (rc) return rc;
  }
  CU(cudaEventRecord(e->ev[1], e->stream));
  CU(cudaEventSynchronize(e->ev[1]));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
  *ms_per_run = ms / (float)iters;
  return GKLB_OK;
}

const void* gklb_pairhmm_table(int which, int* n) {
  const HostTables& t = host_tables();
  switch (which) {
    case 0: if (n) *n = kPh2prSize; return t.ph2pr_f;
    case 1: if (n) *n = kMmSize; return t.mm_f;
    case 2: if (n) *n = kPh2prSize; return t.ph2pr_d;
    case 3: if (n) *n = kMmSize; return t.mm_d;
    default: if (n) *n = 0; return nullptr;
  }
}

}  // extern "C"
