// engine.cu -- host side of the B200 PairHMM engine and its C-ABI (include/gklb_pairhmm.h).
//
// Replaces the native half of GKL's PairHMM binding:
//   initNative / computeLikelihoodsNative / doneNative   pairhmm/IntelPairHmm.cc:55-118,125-181,189-192
//   JavaData::getData (testcase expansion)                pairhmm/JavaData.h:65-111
// The (read, haplotype) cross product is never materialised: reads are bucketed into length
// classes and packed on the device, haplotypes become shared-memory panel images, and pairs are
// addressed by (record, haplotype) index arithmetic inside the kernels.
//
// There is no CPU compute path in this file: without a compute-capability-10 device every
// entry point fails with GKLB_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/gklb_pairhmm.h"
#include "pairhmm_device.cuh"
#include "pairhmm_h2.cuh"
#include "pairhmm_kernels.h"
#include "pairhmm_tables.h"

using namespace gklb;

namespace {

thread_local std::string t_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

}  // namespace

// shared with pdhmm_engine.cu
int gklb_internal_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  t_last_error = buf;
  return code;
}

namespace {

#define CU(call)                                                                                        \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(e_ == cudaErrorMemoryAllocation ? GKLB_ERR_OOM : GKLB_ERR_CUDA, "%s failed: %s", #call, \
                  cudaGetErrorString(e_));                                                              \
  } while (0)

// Length classes: rows per pass = G * K.  A read goes to the first class that holds it; reads
// longer than the last class take n_pass passes of the multi-pass kernel.
struct ClassDef { int G, K; };
const ClassDef kClasses[] = {{8, 4},  {8, 5},  {8, 6},  {8, 7},  {8, 8},  {16, 5}, {16, 6},
                             {16, 7}, {16, 8}, {32, 5}, {32, 6}, {32, 7}, {32, 8}};
const int kNumClasses = (int)(sizeof(kClasses) / sizeof(kClasses[0]));
const int kMultiG = 32, kMultiK = 8;
const int kSmemMax = 232448;  // 227 KB opt-in dynamic shared memory per CTA on sm_100

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
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
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct HostBuf {  // pinned
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct ClassInst {
  int G = 0, K = 0, n_pass = 1, rows = 0, stride = 0;
  bool multi = false;
  const KernelEntry* kf = nullptr;  // fp32 task kernel
  const KernelEntry* kd = nullptr;  // fp64 task + list kernel
  std::vector<int32_t> rid, len;    // record order
  int n_rec = 0;
  // offsets into the meta upload / device buffers
  size_t meta_rid = 0, meta_len = 0;
  size_t rec_off = 0;   // into d_records
  size_t fb_off = 0;    // into d_fb (uint2 units)
  size_t carry_off = 0, carry_stride = 0;
  int counter0 = 0;     // index of this class's first counter (fb count), then per-tile task counters
};

struct Tile {
  int hap0 = 0, n = 0, max_len = 0;
  size_t meta_off = 0;
  uint32_t bytes = 0;
  // the same haplotypes as a pair image (pairhmm_h2.cuh): sorted by length, two per byte column
  int n_pairs = 0;
  size_t pmeta_off = 0;
  uint32_t pbytes = 0;
  std::vector<int> order;   // haplotype indices (in the batch) by decreasing length; pair q = order[2q], order[2q+1]
};

}  // namespace

struct gklb_engine {
  int device = 0;
  bool use_double = false;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> kev;  // pairs of events around each forward-sweep task kernel of the last run
  int kev_used = 0;
  std::mutex mu;
  // device tables
  DevBuf d_tables;
  const float *d_ph2pr_f = nullptr, *d_mm_f = nullptr;
  const double *d_ph2pr_d = nullptr, *d_mm_d = nullptr;
  // staged batch
  bool staged = false;
  int n_reads = 0, n_haps = 0;
  std::vector<ClassInst> classes;
  std::vector<Tile> tiles;
  DevBuf d_read_off, d_hap_off, d_arenas, d_meta, d_records, d_out, d_fb, d_counters, d_carry;
  HostBuf h_meta, h_counters, h_out;
  double* pending_out = nullptr;  // destination of the batch submitted with gklb_engine_submit, until gklb_engine_wait
  size_t arena_pitch = 0;
  const int64_t* p_read_off = nullptr;   // where the staged offsets / arenas live on the device
  const int64_t* p_hap_off = nullptr;
  const uint8_t* p_arenas = nullptr;
  int n_counters = 0;
  int mega_counter0 = 0;   // first of the per-tile unified queue counters
  bool use_mega = false;   // one multi-class launch per tile instead of one launch per class
  gklb_pairhmm_stats stats{};
  // forced kernel (measurement): policy,G,K,warps,var
  bool forced = false;
  int f_policy = 0, f_G = 0, f_K = 0, f_warps = 0, f_var = 0;
};

namespace {

std::mutex g_mu;
std::vector<gklb_engine*> g_engines;   // one per device in use (GKLB_DEVICES); [0] serves small batches alone
gklb_pairhmm_stats g_last_stats{};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int class_cfg(const ClassInst& c);

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
  int p = !strcmp(pol, "f2") ? POL_F2 : !strcmp(pol, "f1") ? POL_F1 : !strcmp(pol, "d1") ? POL_D1 : !strcmp(pol, "h2") ? POL_H2 : -1;
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
  for (int pol : {POL_F2, POL_D1})
    for (int lm = 0; lm < 2; lm++) {
      for (int w : {8, 12}) {
        const void* fn = mega_kernel(pol, lm, w);
        if (fn) CU(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
      }
    }
  return GKLB_OK;
}

int validate(const gklb_pairhmm_batch* b) {
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

// Build the class instances for this batch.
int plan_classes(gklb_engine* e, const gklb_pairhmm_batch* b) {
  e->classes.clear();
  std::vector<int> inst_of_class(kNumClasses, -1);
  std::vector<std::pair<int, int>> multi_inst;  // (n_pass, inst)
  int forced_inst = -1;
  for (int r = 0; r < b->n_reads; r++) {
    const int64_t len64 = b->read_off[r + 1] - b->read_off[r];
    if (len64 > (1 << 24)) return fail(GKLB_ERR_INVALID, "read %d is too long (%lld)", r, (long long)len64);
    const int len = (int)len64;
    int inst = -1;
    if (e->forced && len <= e->f_G * e->f_K) {
      if (forced_inst < 0) {
        ClassInst c;
        c.G = e->f_G; c.K = e->f_K; c.n_pass = 1; c.multi = false;
        c.kf = find_kernel(e->f_policy, c.G, c.K, e->f_warps, 0, e->f_var);
        c.kd = find_kernel(POL_D1, c.G, c.K, -1, 0, -1);
        forced_inst = (int)e->classes.size();
        e->classes.push_back(c);
      }
      inst = forced_inst;
    } else if (len <= kClasses[kNumClasses - 1].G * kClasses[kNumClasses - 1].K) {
      int ci = 0;
      while (kClasses[ci].G * kClasses[ci].K < len) ci++;
      if (inst_of_class[ci] < 0) {
        ClassInst c;
        c.G = kClasses[ci].G; c.K = kClasses[ci].K; c.n_pass = 1; c.multi = false;
        c.kf = find_kernel(POL_F2, c.G, c.K, -1, 0, -1);
        c.kd = find_kernel(POL_D1, c.G, c.K, -1, 0, -1);
        inst_of_class[ci] = (int)e->classes.size();
        e->classes.push_back(c);
      }
      inst = inst_of_class[ci];
    } else {
      const int cap = kMultiG * kMultiK;
      const int n_pass = (len + cap - 1) / cap;
      for (auto& m : multi_inst)
        if (m.first == n_pass) inst = m.second;
      if (inst < 0) {
        ClassInst c;
        c.G = kMultiG; c.K = kMultiK; c.n_pass = n_pass; c.multi = true;
        c.kf = find_kernel(POL_F2, c.G, c.K, -1, 1, -1);
        c.kd = find_kernel(POL_D1, c.G, c.K, -1, 1, -1);
        inst = (int)e->classes.size();
        multi_inst.push_back({n_pass, inst});
        e->classes.push_back(c);
      }
    }
    e->classes[inst].rid.push_back(r);
    e->classes[inst].len.push_back(len);
  }
  for (auto& c : e->classes) {
    if (!c.kf || (!c.kd && !e->forced)) return fail(GKLB_ERR_STATE, "kernel for class G=%d K=%d is not compiled", c.G, c.K);
    c.rows = c.n_pass * c.G * c.K;
    c.stride = (int)align_up((size_t)c.rows, 16);
    const int rpw = (32 / c.G) * 2;  // covers the packed (2 reads / lane) and the fp64 (1 read / lane) kernels
    while (c.rid.size() % rpw) { c.rid.push_back(-1); c.len.push_back(0); }
    c.n_rec = (int)c.rid.size();
  }
  return GKLB_OK;
}

// Per-warp slot: the task's packed records, then (VAR 3) the prior table of 5 symbols x K rows x 32 lanes.
uint32_t warp_slot_bytes(const ClassInst& c, const KernelEntry* k, bool list_mode) {
  const uint32_t rpw = (uint32_t)((32 / c.G) * k->nr);
  const uint32_t rec = (list_mode || c.multi) ? 0u : (uint32_t)align_up((size_t)rpw * 5 * c.stride, 128);
  const uint32_t tbl = (k->policy == POL_H2) ? (uint32_t)(kPriorSyms * c.K * 32 * 4)
                       : (k->var >= 3)          ? (uint32_t)(kPriorSyms * c.K * 32 * 8) : 0u;
  return rec + tbl;
}

uint32_t slot_bytes_total(const ClassInst& c, const KernelEntry* k) {
  return (uint32_t)k->warps * (uint32_t)align_up((size_t)warp_slot_bytes(c, k, false), 128);
}

// Split the haplotypes into tiles whose panel image fits beside the largest slot area.
int plan_tiles(gklb_engine* e, const gklb_pairhmm_batch* b, size_t* meta_bytes) {
  uint32_t worst_slots = 0;
  for (auto& c : e->classes) {
    worst_slots = std::max(worst_slots, slot_bytes_total(c, c.kf));
    if (c.kd) worst_slots = std::max(worst_slots, slot_bytes_total(c, c.kd));
    // the multi-class kernel runs every class with its own warp count (up to 12) and the largest slot
    worst_slots = std::max(worst_slots, 12u * (uint32_t)align_up((size_t)warp_slot_bytes(c, c.kf, false), 128));
  }
  const long long budget = (long long)kSmemMax - 4096 - worst_slots;
  e->tiles.clear();
  int h = 0;
  while (h < b->n_haps) {
    Tile t;
    t.hap0 = h;
    size_t data = 0;
    while (h < b->n_haps) {
      const size_t len = (size_t)(b->hap_off[h + 1] - b->hap_off[h]);
      const size_t add = kHapLeftMargin + len + kHapRightMargin;
      const size_t header = align_up((size_t)8 * (t.n + 1), 16);
      if (t.n > 0 && (long long)(header + data + add + 16) > budget) break;
      if (t.n == 0 && (long long)(header + add + 16) > budget)
        return fail(GKLB_ERR_INVALID, "haplotype %d (%zu bases) does not fit in shared memory", h, len);
      data += add;
      t.max_len = std::max(t.max_len, (int)len);
      t.n++;
      h++;
    }
    t.bytes = (uint32_t)align_up(align_up((size_t)8 * t.n, 16) + data, 16);
    t.meta_off = *meta_bytes;
    *meta_bytes += align_up(t.bytes, 128);
    // pair image: haplotypes by decreasing length, adjacent ones share a byte column (never larger than t.bytes)
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
    t.pmeta_off = *meta_bytes;
    *meta_bytes += align_up(t.pbytes, 128);
    e->tiles.push_back(t);
  }
  return GKLB_OK;
}

void build_tile_image(const Tile& t, const gklb_pairhmm_batch* b, uint8_t* img) {
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
    for (int c = 0; c < len; c++) dst[c] = panel_byte(b->hap_bases[o + c]);
    off += kHapLeftMargin + len + kHapRightMargin;
  }
}

void build_pair_image(const Tile& t, const gklb_pairhmm_batch* b, uint8_t* img) {
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
    const int bb = has_b ? t.order[2 * q + 1] : a;   // an odd haplotype out is paired with itself, result B dropped
    const int64_t oa = b->hap_off[a], ob = b->hap_off[bb];
    const int la = (int)(b->hap_off[a + 1] - oa), lb = (int)(b->hap_off[bb + 1] - ob);
    ppos[q] = (int32_t)(off + kHapLeftMargin - 1);
    lenA[q] = la;
    lenB[q] = lb;
    idxA[q] = a;
    idxB[q] = has_b ? bb : -1;
    uint8_t* dst = img + off + kHapLeftMargin;
    for (int c = 0; c < la; c++)
      dst[c] = (uint8_t)(base_index(b->hap_bases[oa + c]) | (c < lb ? base_index(b->hap_bases[ob + c]) << 3 : 0));
    off += kHapLeftMargin + la + kHapRightMargin;
  }
}

int do_stage(gklb_engine* e, const gklb_pairhmm_batch* b, bool hap_on_device) {
  e->staged = false;
  int rc = validate(b);
  if (rc) return rc;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));  // the pinned staging buffers are about to be rewritten
  e->n_reads = b->n_reads;
  e->n_haps = b->n_haps;
  e->stats = gklb_pairhmm_stats{};
  if (b->n_reads == 0 || b->n_haps == 0) {
    e->classes.clear();
    e->tiles.clear();
    e->staged = true;
    return GKLB_OK;
  }
  if ((rc = parse_forced(e))) return rc;
  const int64_t total_read = b->read_off[b->n_reads];
  const int64_t total_hap = b->hap_off[b->n_haps];
  e->stats.pairs = (int64_t)b->n_reads * b->n_haps;
  e->stats.cells = total_read * total_hap;

  // the haplotype panel images are built on the host: fetch the bases if they live on the device
  std::vector<uint8_t> hap_host;
  gklb_pairhmm_batch hb = *b;
  if (hap_on_device) {
    hap_host.resize((size_t)total_hap);
    CU(cudaMemcpyAsync(hap_host.data(), b->hap_bases, (size_t)total_hap, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    hb.hap_bases = hap_host.data();
  }

  if ((rc = plan_classes(e, b))) return rc;
  size_t meta_bytes = 0;
  if ((rc = plan_tiles(e, &hb, &meta_bytes))) return rc;
  size_t rec_bytes = 0, fb_items = 0, carry_bytes = 0;
  int counters = 0;
  for (auto& c : e->classes) {
    c.meta_rid = meta_bytes;
    meta_bytes += align_up(sizeof(int32_t) * c.n_rec, 128);
    c.meta_len = meta_bytes;
    meta_bytes += align_up(sizeof(int32_t) * c.n_rec, 128);
    c.rec_off = rec_bytes;
    rec_bytes += align_up((size_t)c.n_rec * 5 * c.stride, 128);
    c.fb_off = fb_items;
    if (!e->use_double) fb_items += (size_t)c.n_rec * b->n_haps;
    c.counter0 = counters;
    counters += 1 + 2 * (int)e->tiles.size();
    if (c.multi) {
      int max_len = 0;
      for (auto& t : e->tiles) max_len = std::max(max_len, t.max_len);
      // per-warp scratch: sized for the widest CTA any kernel of this class may run with (the multi-class
      // kernel uses up to 12 warps)
      const int warps = std::max(12, std::max(c.kf->warps, c.kd ? c.kd->warps : 0));
      c.carry_stride = (size_t)(32 / c.G) * 6 * (max_len + 2) * 8;
      c.carry_off = carry_bytes;
      carry_bytes += c.carry_stride * warps * e->num_sms;
    }
  }
  e->mega_counter0 = counters;
  counters += 2 * (int)e->tiles.size();
  e->n_counters = counters;
  {
    const char* mg = getenv("GKLB_MEGA");
    const bool all_cfg = std::all_of(e->classes.begin(), e->classes.end(), [](const ClassInst& c) { return class_cfg(c) >= 0; });
    e->use_mega = !e->forced && all_cfg && (int)e->classes.size() <= kMaxMegaClasses &&
                  (mg ? atoi(mg) != 0 : e->classes.size() > 1);
  }

  // small host-resident batches: offsets + arenas are staged behind the meta block (see below)
  e->arena_pitch = align_up((size_t)total_read, 256);
  const bool consolidate = !hap_on_device && (size_t)total_read * 5 <= (4u << 20);
  size_t st_read_off = 0, st_hap_off = 0, st_arena = 0, staged_bytes = meta_bytes;
  if (consolidate) {
    st_read_off = staged_bytes;
    staged_bytes += align_up(sizeof(int64_t) * ((size_t)b->n_reads + 1), 256);
    st_hap_off = staged_bytes;
    staged_bytes += align_up(sizeof(int64_t) * ((size_t)b->n_haps + 1), 256);
    st_arena = staged_bytes;
    staged_bytes += 5 * e->arena_pitch;
  }
  CU(e->h_meta.ensure(staged_bytes));
  CU(e->d_meta.ensure(staged_bytes));
  CU(e->d_records.ensure(rec_bytes));
  CU(e->d_read_off.ensure(sizeof(int64_t) * ((size_t)b->n_reads + 1)));
  CU(e->d_hap_off.ensure(sizeof(int64_t) * ((size_t)b->n_haps + 1)));
  if (!consolidate) CU(e->d_arenas.ensure(e->arena_pitch * 5));
  CU(e->d_out.ensure(sizeof(double) * (size_t)e->stats.pairs));
  if (fb_items) CU(e->d_fb.ensure(sizeof(uint2) * fb_items));
  CU(e->d_counters.ensure(sizeof(unsigned int) * (size_t)counters));
  CU(e->h_counters.ensure(sizeof(unsigned int) * (size_t)counters));
  if (carry_bytes) CU(e->d_carry.ensure(carry_bytes));

  uint8_t* hm = static_cast<uint8_t*>(e->h_meta.p);
  for (auto& t : e->tiles) {
    build_tile_image(t, &hb, hm + t.meta_off);
    build_pair_image(t, &hb, hm + t.pmeta_off);
  }
  for (auto& c : e->classes) {
    memcpy(hm + c.meta_rid, c.rid.data(), sizeof(int32_t) * c.n_rec);
    memcpy(hm + c.meta_len, c.len.data(), sizeof(int32_t) * c.n_rec);
  }
  cudaStream_t s = e->stream;
  const size_t off_bytes_r = sizeof(int64_t) * ((size_t)b->n_reads + 1), off_bytes_h = sizeof(int64_t) * ((size_t)b->n_haps + 1);
  const uint8_t* src[5] = {b->read_bases, b->read_quals, b->ins_gop, b->del_gop, b->gcp};
  uint8_t* dmw = static_cast<uint8_t*>(e->d_meta.p);
  if (consolidate) {
    // small batch: offsets and arenas ride in the same pinned staging buffer -> one host->device copy
    memcpy(hm + st_read_off, b->read_off, off_bytes_r);
    memcpy(hm + st_hap_off, b->hap_off, off_bytes_h);
    for (int i = 0; i < 5; i++) memcpy(hm + st_arena + i * e->arena_pitch, src[i], (size_t)total_read);
    CU(cudaMemcpyAsync(dmw, hm, staged_bytes, cudaMemcpyHostToDevice, s));
    e->p_read_off = reinterpret_cast<const int64_t*>(dmw + st_read_off);
    e->p_hap_off = reinterpret_cast<const int64_t*>(dmw + st_hap_off);
    e->p_arenas = dmw + st_arena;
  } else {
    CU(cudaMemcpyAsync(dmw, hm, meta_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->d_read_off.p, b->read_off, off_bytes_r, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(e->d_hap_off.p, b->hap_off, off_bytes_h, cudaMemcpyHostToDevice, s));
    uint8_t* da = static_cast<uint8_t*>(e->d_arenas.p);
    for (int i = 0; i < 5; i++)
      CU(cudaMemcpyAsync(da + i * e->arena_pitch, src[i], (size_t)total_read, cudaMemcpyDefault, s));
    e->p_read_off = static_cast<const int64_t*>(e->d_read_off.p);
    e->p_hap_off = static_cast<const int64_t*>(e->d_hap_off.p);
    e->p_arenas = da;
  }

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
  e->staged = true;
  return GKLB_OK;
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

// Fill the per-class kernel parameters for one (class, tile, kernel) combination.
void fill_params(gklb_engine* e, const ClassInst& c, const Tile& t, const KernelEntry* k, bool list_mode, int tile_index,
                 SweepParams* out, int* grid_out, size_t* smem_out, uint32_t* slot_bytes_out) {
  const bool dbl = (k->policy == POL_D1);
  const HostTables& ht = host_tables();
  SweepParams& p = *out;
  memset(&p, 0, sizeof(p));
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  p.panel.image = dm + t.meta_off;
  p.panel.bytes = t.bytes;
  p.panel.n_haps = t.n;
  p.panel.hap0 = t.hap0;
  p.panel.n_haps_total = e->n_haps;
  p.panel.max_hap_len = t.max_len;
  p.cls.records = static_cast<const uint8_t*>(e->d_records.p) + c.rec_off;
  p.cls.rec_rid = reinterpret_cast<const int32_t*>(dm + c.meta_rid);
  p.cls.rec_len = reinterpret_cast<const int32_t*>(dm + c.meta_len);
  p.cls.n_rec = c.n_rec;
  p.cls.rows = c.rows;
  p.cls.stride = c.stride;
  p.cls.n_pass = c.n_pass;
  p.ph2pr = dbl ? (const void*)e->d_ph2pr_d : (const void*)e->d_ph2pr_f;
  p.mm = dbl ? (const void*)e->d_mm_d : (const void*)e->d_mm_f;
  p.out = static_cast<double*>(e->d_out.p);
  unsigned int* counters = static_cast<unsigned int*>(e->d_counters.p);
  p.fb_count = counters + c.counter0;
  p.fb_items = e->d_fb.p ? static_cast<uint2*>(e->d_fb.p) + c.fb_off : nullptr;
  p.task_counter = counters + c.counter0 + 1 + 2 * tile_index + (list_mode ? 1 : 0);
  p.carry = c.multi ? static_cast<uint8_t*>(e->d_carry.p) + c.carry_off : nullptr;
  p.carry_stride_bytes = c.carry_stride;
  p.init_const = dbl ? ht.init_d : (double)ht.init_f;
  p.log10_init = dbl ? ht.log10_init_d : (double)ht.log10_init_f;

  const int gpw = 32 / c.G;
  const int rpw = gpw * k->nr;
  const int slots = e->num_sms * k->warps;
  if (!list_mode) {
    const int n_blocks = c.n_rec / rpw;
    const long long tasks_per_warp = tasks_per_warp_target();
    long long chunk = ((long long)n_blocks * t.n) / (tasks_per_warp * slots);
    chunk = std::max(1LL, std::min(chunk, 32LL));
    if (c.multi) chunk = 1;
    chunk = std::min<long long>(chunk, t.n);
    p.hap_chunk = (int)chunk;
    p.n_chunks = (t.n + p.hap_chunk - 1) / p.hap_chunk;
    p.n_tasks = n_blocks * p.n_chunks;
    *grid_out = std::min(e->num_sms, (p.n_tasks + k->warps - 1) / k->warps);
    *slot_bytes_out = warp_slot_bytes(c, k, false);
    *smem_out = smem_layout(k->warps, t.bytes, *slot_bytes_out, dbl ? 8 : 4).total;
  } else {
    p.list_items = p.fb_items;
    p.list_count = p.fb_count;
    *grid_out = e->num_sms;
    *slot_bytes_out = warp_slot_bytes(c, k, true);
    *smem_out = smem_layout(k->warps, t.bytes, *slot_bytes_out, 8).total;
  }
  p.slot_bytes = *slot_bytes_out;
}

// The two-haplotypes-per-lane kernel (pairhmm_h2.cuh): tasks are (block of 32/G records) x (chunk of pairs).
int launch_h2(gklb_engine* e, const ClassInst& c, const Tile& t, const KernelEntry* k, int tile_index) {
  const HostTables& ht = host_tables();
  H2Params p;
  memset(&p, 0, sizeof(p));
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  p.panel.image = dm + t.pmeta_off;
  p.panel.bytes = t.pbytes;
  p.panel.n_pairs = t.n_pairs;
  p.panel.n_haps_total = e->n_haps;
  p.panel.max_hap_len = t.max_len;
  p.cls.records = static_cast<const uint8_t*>(e->d_records.p) + c.rec_off;
  p.cls.rec_rid = reinterpret_cast<const int32_t*>(dm + c.meta_rid);
  p.cls.rec_len = reinterpret_cast<const int32_t*>(dm + c.meta_len);
  p.cls.n_rec = c.n_rec;
  p.cls.rows = c.rows;
  p.cls.stride = c.stride;
  p.cls.n_pass = 1;
  p.ph2pr = e->d_ph2pr_f;
  p.mm = e->d_mm_f;
  p.out = static_cast<double*>(e->d_out.p);
  unsigned int* counters = static_cast<unsigned int*>(e->d_counters.p);
  p.fb_count = counters + c.counter0;
  p.fb_items = e->d_fb.p ? static_cast<uint2*>(e->d_fb.p) + c.fb_off : nullptr;
  p.task_counter = counters + c.counter0 + 1 + 2 * tile_index;
  p.init_const = ht.init_f;
  p.log10_init = ht.log10_init_f;
  const int gpw = 32 / c.G;
  const int n_blocks = c.n_rec / gpw;
  const int slots = e->num_sms * k->warps;
  long long chunk = ((long long)n_blocks * t.n_pairs) / ((long long)tasks_per_warp_target() * slots);
  chunk = std::max(1LL, std::min(chunk, 32LL));
  chunk = std::min<long long>(chunk, t.n_pairs);
  p.pair_chunk = (int)chunk;
  p.n_chunks = (t.n_pairs + p.pair_chunk - 1) / p.pair_chunk;
  p.n_tasks = n_blocks * p.n_chunks;
  const int grid = std::min(e->num_sms, (p.n_tasks + k->warps - 1) / k->warps);
  p.slot_bytes = warp_slot_bytes(c, k, false);
  const size_t smem = smem_layout(k->warps, t.pbytes, p.slot_bytes, 4).total;
  if (smem > (size_t)kSmemMax) return fail(GKLB_ERR_STATE, "shared memory plan exceeds the device limit (%zu)", smem);
  if (grid <= 0) return GKLB_OK;
  while ((int)e->kev.size() < e->kev_used + 2) {
    cudaEvent_t ev;
    CU(cudaEventCreate(&ev));
    e->kev.push_back(ev);
  }
  CU(cudaEventRecord(e->kev[e->kev_used], e->stream));
  CU(launch_h2_kernel(k->fn_tasks, p, grid, k->warps * 32, smem, e->stream));
  CU(cudaEventRecord(e->kev[e->kev_used + 1], e->stream));
  e->kev_used += 2;
  e->stats.kernel_launches++;
  return GKLB_OK;
}

int launch_one(gklb_engine* e, const ClassInst& c, const Tile& t, const KernelEntry* k, bool list_mode, int tile_index) {
  if (k->policy == POL_H2) return launch_h2(e, c, t, k, tile_index);
  SweepParams p;
  int grid;
  size_t smem;
  uint32_t slot_bytes;
  fill_params(e, c, t, k, list_mode, tile_index, &p, &grid, &smem, &slot_bytes);
  if (smem > (size_t)kSmemMax) return fail(GKLB_ERR_STATE, "shared memory plan exceeds the device limit (%zu)", smem);
  if (grid <= 0) return GKLB_OK;
  if (!list_mode) {
    while ((int)e->kev.size() < e->kev_used + 2) {
      cudaEvent_t ev;
      CU(cudaEventCreate(&ev));
      e->kev.push_back(ev);
    }
    CU(cudaEventRecord(e->kev[e->kev_used], e->stream));
  }
  CU(launch_sweep(list_mode ? k->fn_list : k->fn_tasks, p, grid, k->warps * 32, smem, e->stream));
  if (!list_mode) {
    CU(cudaEventRecord(e->kev[e->kev_used + 1], e->stream));
    e->kev_used += 2;
  }
  e->stats.kernel_launches++;
  return GKLB_OK;
}

int class_cfg(const ClassInst& c) {
  if (c.multi) return kCfgMulti;
  for (int i = 0; i < kNumClasses; i++)
    if (kClasses[i].G == c.G && kClasses[i].K == c.K) return i;
  return -1;
}

// One launch for all classes of a tile (see k_mega_tasks).  policy: POL_F2 or POL_D1.
int mega_warps(int policy, bool list_mode) {
  static const int w = [] {
    const char* v = getenv("GKLB_MEGA_WARPS");
    return v ? atoi(v) : 12;
  }();
  return (policy == POL_F2 && !list_mode && w == 12) ? 12 : 8;
}

int launch_mega_tile(gklb_engine* e, const Tile& t, int tile_index, int policy, bool list_mode) {
  const int warps = mega_warps(policy, list_mode);
  MegaParams mp;
  memset(&mp, 0, sizeof(mp));
  // longest classes first: their tasks are the most expensive, schedule them early
  std::vector<const ClassInst*> order;
  for (auto& c : e->classes) order.push_back(&c);
  std::sort(order.begin(), order.end(), [](const ClassInst* a, const ClassInst* b) { return a->rows > b->rows; });
  size_t smem = 0;
  uint32_t slot_bytes = 0;
  int tasks = 0;
  for (const ClassInst* c : order) {
    const KernelEntry* k = (policy == POL_D1) ? c->kd : c->kf;
    const int i = mp.n_classes++;
    int grid;
    size_t sm;
    uint32_t sb;
    fill_params(e, *c, t, k, list_mode, tile_index, &mp.cls[i], &grid, &sm, &sb);
    mp.cfg[i] = class_cfg(*c);
    if (mp.cfg[i] < 0) return fail(GKLB_ERR_STATE, "class G=%d K=%d has no multi-class configuration", c->G, c->K);
    tasks += mp.cls[i].n_tasks;
    mp.task_end[i] = tasks;
    slot_bytes = std::max(slot_bytes, sb);
  }
  unsigned int* counters = static_cast<unsigned int*>(e->d_counters.p);
  mp.queue = counters + e->mega_counter0 + 2 * tile_index + (list_mode ? 1 : 0);
  smem = smem_layout(warps, t.bytes, slot_bytes, policy == POL_D1 ? 8 : 4).total;
  if (smem > (size_t)kSmemMax)
    return fail(GKLB_ERR_STATE, "shared memory plan exceeds the device limit (%zu: panel %u, %d warps x %u, policy %d list %d)",
                smem, t.bytes, warps, slot_bytes, policy, (int)list_mode);
  const int grid = list_mode ? e->num_sms : std::min(e->num_sms, (tasks + warps - 1) / warps);
  if (grid <= 0) return GKLB_OK;
  if (!list_mode) {
    while ((int)e->kev.size() < e->kev_used + 2) {
      cudaEvent_t ev;
      CU(cudaEventCreate(&ev));
      e->kev.push_back(ev);
    }
    CU(cudaEventRecord(e->kev[e->kev_used], e->stream));
  }
  CU(launch_mega(mega_kernel(policy, list_mode ? 1 : 0, warps), mp, slot_bytes, list_mode ? 1 : 0, grid, warps * 32, smem, e->stream));
  if (!list_mode) {
    CU(cudaEventRecord(e->kev[e->kev_used + 1], e->stream));
    e->kev_used += 2;
  }
  e->stats.kernel_launches++;
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
  for (size_t ti = 0; ti < e->tiles.size(); ti++) {
    if (e->use_mega) {
      int rc;
      if (e->use_double) {
        if ((rc = launch_mega_tile(e, e->tiles[ti], (int)ti, POL_D1, false))) return rc;
      } else {
        if ((rc = launch_mega_tile(e, e->tiles[ti], (int)ti, POL_F2, false))) return rc;
        if ((rc = launch_mega_tile(e, e->tiles[ti], (int)ti, POL_D1, true))) return rc;
      }
      continue;
    }
    for (auto& c : e->classes) {
      int rc;
      if (e->use_double) {
        if ((rc = launch_one(e, c, e->tiles[ti], c.kd, false, (int)ti))) return rc;
      } else {
        if ((rc = launch_one(e, c, e->tiles[ti], c.kf, false, (int)ti))) return rc;
        if (c.kd && c.kf->policy != POL_D1)
          if ((rc = launch_one(e, c, e->tiles[ti], c.kd, true, (int)ti))) return rc;
      }
    }
  }
  return GKLB_OK;
}

int do_fetch(gklb_engine* e, double* out) {
  if (!e->staged) return fail(GKLB_ERR_STATE, "nothing staged");
  CU(cudaSetDevice(e->device));
  if (e->classes.empty()) return GKLB_OK;
  if (!out) return fail(GKLB_ERR_INVALID, "likelihoods is null");
  // Small results travel through the engine's pinned buffer: a device->host copy into the caller's (usually
  // pageable) array is staged by the driver and costs tens of microseconds more per call than the memcpy here.
  const size_t out_bytes = sizeof(double) * (size_t)e->stats.pairs;
  const bool via_pinned = out_bytes <= ((size_t)2 << 20) && !e->pending_out;  // the buffer also serves submit/wait
  if (via_pinned) CU(e->h_out.ensure(out_bytes));
  CU(cudaMemcpyAsync(via_pinned ? e->h_out.p : (void*)out, e->d_out.p, out_bytes, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaMemcpyAsync(e->h_counters.p, e->d_counters.p, sizeof(unsigned int) * (size_t)e->n_counters,
                     cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  if (via_pinned) memcpy(out, e->h_out.p, out_bytes);
  int64_t fb = 0;
  const unsigned int* hc = static_cast<const unsigned int*>(e->h_counters.p);
  for (auto& c : e->classes) fb += hc[c.counter0];
  e->stats.fallback_pairs = fb;
  return GKLB_OK;
}

int do_compute(gklb_engine* e, const gklb_pairhmm_batch* b, double* out) {
  int rc;
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev[0], e->stream));
  if ((rc = do_stage(e, b, false))) return rc;
  CU(cudaEventRecord(e->ev[1], e->stream));
  if ((rc = do_run(e))) return rc;
  CU(cudaEventRecord(e->ev[2], e->stream));
  if ((rc = do_fetch(e, out))) return rc;
  CU(cudaEventRecord(e->ev[3], e->stream));
  CU(cudaEventSynchronize(e->ev[3]));
  cudaEventElapsedTime(&e->stats.h2d_ms, e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&e->stats.kernel_ms, e->ev[1], e->ev[2]);
  cudaEventElapsedTime(&e->stats.d2h_ms, e->ev[2], e->ev[3]);
  return GKLB_OK;
}

// Asynchronous compute: everything is queued on the engine's stream and the likelihoods travel to a pinned
// buffer (a device->host copy into pageable memory would block the host until the kernels are done).
int do_submit(gklb_engine* e, const gklb_pairhmm_batch* b, double* out) {
  if (e->pending_out) return fail(GKLB_ERR_STATE, "a submitted batch is still in flight: call gklb_engine_wait first");
  int rc;
  if ((rc = do_stage(e, b, false))) return rc;
  if ((rc = do_run(e))) return rc;
  if (e->classes.empty()) return GKLB_OK;
  if (!out) return fail(GKLB_ERR_INVALID, "likelihoods is null");
  CU(e->h_out.ensure(sizeof(double) * (size_t)e->stats.pairs));
  CU(cudaMemcpyAsync(e->h_out.p, e->d_out.p, sizeof(double) * (size_t)e->stats.pairs, cudaMemcpyDeviceToHost, e->stream));
  CU(cudaMemcpyAsync(e->h_counters.p, e->d_counters.p, sizeof(unsigned int) * (size_t)e->n_counters,
                     cudaMemcpyDeviceToHost, e->stream));
  e->pending_out = out;
  return GKLB_OK;
}

int do_wait(gklb_engine* e) {
  CU(cudaSetDevice(e->device));
  CU(cudaStreamSynchronize(e->stream));
  if (!e->pending_out) return GKLB_OK;
  memcpy(e->pending_out, e->h_out.p, sizeof(double) * (size_t)e->stats.pairs);
  e->pending_out = nullptr;
  int64_t fb = 0;
  const unsigned int* hc = static_cast<const unsigned int*>(e->h_counters.p);
  for (auto& c : e->classes) fb += hc[c.counter0];
  e->stats.fallback_pairs = fb;
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
                    &e->d_counters, &e->d_carry})
    b->release();
  e->h_meta.release();
  e->h_counters.release();
  e->h_out.release();
  for (auto& ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : e->kev) cudaEventDestroy(ev);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

}  // namespace

extern "C" {

int gklb_pairhmm_init(int use_double, int max_threads) {
  (void)max_threads;  // GKL's non-OpenMP library ignores it as well (IntelPairHmm.cc:85-89)
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto* e : g_engines) destroy_engine(e);
  g_engines.clear();
  // GKLB_DEVICES = "all" | "0,1,2,..." : shard large batches over several GPUs inside this process
  // (the JVM is one process); GKLB_DEVICE = n : a single device (default 0)
  std::vector<int> devices;
  const char* many = getenv("GKLB_DEVICES");
  if (many && *many) {
    if (!strcmp(many, "all")) {
      int n = 0;
      cudaGetDeviceCount(&n);
      for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) devices.push_back(i);
      }
    } else {
      for (const char* q = many; *q;) {
        devices.push_back(atoi(q));
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
      }
    }
  }
  if (devices.empty()) {
    const char* dev = getenv("GKLB_DEVICE");
    devices.push_back(dev ? atoi(dev) : 0);
  }
  for (int d : devices) {
    gklb_engine* e = nullptr;
    const int rc = create_engine(&e, d, use_double);
    if (rc) {
      for (auto* x : g_engines) destroy_engine(x);
      g_engines.clear();
      return rc;
    }
    g_engines.push_back(e);
  }
  return GKLB_OK;
}

// Batches below this many cells are not worth splitting: one GPU finishes them in well under a millisecond.
static const long long kMinCellsPerDevice = 4000000000LL;

int gklb_pairhmm_compute(const gklb_pairhmm_batch* batch, double* likelihoods) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_engines.empty()) return fail(GKLB_ERR_STATE, "gklb_pairhmm_init has not been called");
  int rc = validate(batch);
  if (rc) return rc;
  int n_dev = (int)g_engines.size();
  if (batch->n_reads > 0 && batch->n_haps > 0) {
    const long long cells = (long long)batch->read_off[batch->n_reads] * (long long)batch->hap_off[batch->n_haps];
    n_dev = (int)std::max(1LL, std::min<long long>(n_dev, cells / kMinCellsPerDevice));
    n_dev = std::min(n_dev, batch->n_reads);
  } else {
    n_dev = 1;
  }
  if (n_dev <= 1) {
    rc = do_compute(g_engines[0], batch, likelihoods);
    g_last_stats = g_engines[0]->stats;
    return rc;
  }
  // Reads are sharded into contiguous ranges balanced by total length (every read meets every haplotype, so
  // cells are proportional to read length); range g gets a contiguous slab of the read-major output
  // (JavaData.h:94-105).  Each device copies its shard and the panel over its own PCIe link and writes its slab
  // straight into the caller's array: the shards never need to meet on one GPU.
  const int64_t total = batch->read_off[batch->n_reads];
  std::vector<int> cut(n_dev + 1, 0);
  {
    int r = 0;
    for (int g = 1; g < n_dev; g++) {
      const int64_t target = total * g / n_dev;
      while (r < batch->n_reads && batch->read_off[r] < target) r++;
      cut[g] = std::max(r, cut[g - 1]);
    }
    cut[n_dev] = batch->n_reads;
  }
  std::vector<std::vector<int64_t>> offs(n_dev);
  std::vector<gklb_pairhmm_batch> sub(n_dev);
  std::vector<int> rcs(n_dev, GKLB_OK);
  std::vector<std::string> errs(n_dev);
  std::vector<std::thread> th;
  for (int g = 0; g < n_dev; g++) {
    const int lo = cut[g], hi = cut[g + 1];
    const int64_t base = batch->read_off[lo];
    offs[g].resize((size_t)(hi - lo) + 1);
    for (int r = lo; r <= hi; r++) offs[g][r - lo] = batch->read_off[r] - base;
    sub[g] = *batch;
    sub[g].n_reads = hi - lo;
    sub[g].read_off = offs[g].data();
    sub[g].read_bases = batch->read_bases + base;
    sub[g].read_quals = batch->read_quals + base;
    sub[g].ins_gop = batch->ins_gop + base;
    sub[g].del_gop = batch->del_gop + base;
    sub[g].gcp = batch->gcp + base;
  }
  for (int g = 0; g < n_dev; g++) {
    th.emplace_back([&, g] {
      if (sub[g].n_reads == 0) return;
      rcs[g] = do_compute(g_engines[g], &sub[g], likelihoods + (size_t)cut[g] * batch->n_haps);
      if (rcs[g]) errs[g] = t_last_error;  // last error is thread-local: carry it to the caller's thread
    });
  }
  for (auto& t : th) t.join();
  g_last_stats = gklb_pairhmm_stats{};
  for (int g = 0; g < n_dev; g++) {
    if (rcs[g]) { t_last_error = errs[g]; return rcs[g]; }
    const gklb_pairhmm_stats& st = g_engines[g]->stats;
    g_last_stats.pairs += st.pairs;
    g_last_stats.cells += st.cells;
    g_last_stats.fallback_pairs += st.fallback_pairs;
    g_last_stats.kernel_launches += st.kernel_launches;
    g_last_stats.n_classes = std::max(g_last_stats.n_classes, st.n_classes);
    g_last_stats.h2d_ms = std::max(g_last_stats.h2d_ms, st.h2d_ms);
    g_last_stats.kernel_ms = std::max(g_last_stats.kernel_ms, st.kernel_ms);
    g_last_stats.d2h_ms = std::max(g_last_stats.d2h_ms, st.d2h_ms);
  }
  return GKLB_OK;
}

int gklb_pairhmm_done(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto* e : g_engines) destroy_engine(e);
  g_engines.clear();
  return GKLB_OK;
}

int gklb_pairhmm_devices_in_use(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return (int)g_engines.size();
}

int gklb_pairhmm_last_stats(gklb_pairhmm_stats* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out) return fail(GKLB_ERR_INVALID, "out is null");
  *out = g_last_stats;
  return GKLB_OK;
}

int gklb_engine_create(gklb_engine** out, int device, int use_double) {
  if (!out) return fail(GKLB_ERR_INVALID, "out is null");
  return create_engine(out, device, use_double);
}

int gklb_engine_destroy(gklb_engine* e) {
  destroy_engine(e);
  return GKLB_OK;
}

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
  std::lock_guard<std::mutex> lk(e->mu);
  return do_compute(e, batch, likelihoods);
}

int gklb_engine_submit(gklb_engine* e, const gklb_pairhmm_batch* batch, double* likelihoods) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_submit(e, batch, likelihoods);
}

int gklb_engine_wait(gklb_engine* e) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_wait(e);
}

int gklb_engine_stage(gklb_engine* e, const gklb_pairhmm_batch* batch) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_stage(e, batch, false);
}

int gklb_engine_stage_device(gklb_engine* e, const gklb_pairhmm_batch* batch) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_stage(e, batch, true);
}

int gklb_engine_update_haps_device(gklb_engine* e, const void* hap_bases_dev) {
  if (!e || !hap_bases_dev) return fail(GKLB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->staged) return fail(GKLB_ERR_STATE, "nothing staged");
  CU(cudaSetDevice(e->device));
  uint8_t* dm = static_cast<uint8_t*>(e->d_meta.p);
  for (auto& t : e->tiles)
    CU(launch_fill_panel(dm + t.meta_off, t.n, t.hap0, e->p_hap_off,
                         static_cast<const uint8_t*>(hap_bases_dev), e->stream));
  return GKLB_OK;
}

int gklb_engine_run(gklb_engine* e) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_run(e);
}

int gklb_engine_fetch(gklb_engine* e, double* likelihoods) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  std::lock_guard<std::mutex> lk(e->mu);
  return do_fetch(e, likelihoods);
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

const char* gklb_last_error(void) { return t_last_error.c_str(); }

const char* gklb_version(void) { return "gkl_b200 0.1 (sm_100a)"; }

int gklb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  int ok = 0;
  for (int i = 0; i < n; i++) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
  }
  return ok;
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
