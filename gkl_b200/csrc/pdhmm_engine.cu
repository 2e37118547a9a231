// pdhmm_engine.cu -- host side and C-ABI (include/gklb_pdhmm.h) of the PDHMM engine.
// Replaces pdhmm/IntelPDHMM.cc + pdhmm-implementation.h:292-396 below the JNI boundary.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <array>
#include <vector>

#include "../../include/gklb_pdhmm.h"
#include "pdhmm_device.cuh"

using namespace gklb;

extern int gklb_internal_fail(int code, const char* fmt, ...);  // engine.cu: sets gklb_last_error()

namespace {

#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess)                                                                                \
      return gklb_internal_fail(e_ == cudaErrorMemoryAllocation ? GKLB_ERR_OOM : GKLB_ERR_CUDA, "%s failed: %s", \
                                #call, cudaGetErrorString(e_));                                           \
  } while (0)

constexpr int kG = 32, kK = 4, kWarps = 8;
// k_pdhmm2 instantiations: rows per lane, warps per CTA, longest read (32 K - K: lane 0 stays all padding).
// K = 4 fits 168 registers (two 8-byte spills outside the loops), i.e. a third warp per scheduler; K = 5 / 6 serve
// 150-base reads and need the full register file.
struct V2Config { int K, warps, max_read; const void* fn; };
const V2Config kV2[] = {
    {4, 12, 32 * 4 - 4, reinterpret_cast<const void*>(&k_pdhmm2<4, 12>)},
    {5, 8, 32 * 5 - 5, reinterpret_cast<const void*>(&k_pdhmm2<5, 8>)},
    {6, 8, 32 * 6 - 6, reinterpret_cast<const void*>(&k_pdhmm2<6, 8>)},
};
constexpr int kNumV2 = (int)(sizeof(kV2) / sizeof(kV2[0]));
// k_pdhmm3 instantiations (cross layout only): lanes per read, rows per lane, warps per CTA, longest read (lane 0 of a
// read stays all padding).  Two 101-base reads per warp, or one read of up to 155 rows.
// `ids`: kinds of columns the per-pair prior table holds, one of them "no column" (haplotypes with more go to k_pdhmm2).
struct V3Config { int G, K, warps, ids, max_read; const void* fn; const void* fn_wfold; };
#define GKLB_V3(G, K, W, N)                                                          \
  {G, K, W, N, G * K - K, reinterpret_cast<const void*>(&k_pdhmm3<G, K, W, N, false>), \
   reinterpret_cast<const void*>(&k_pdhmm3<G, K, W, N, true>)}
const V3Config kV3[] = {
    GKLB_V3(16, 7, 8, 11),   // haplotypes of up to ~470 columns
    GKLB_V3(16, 7, 8, 9),
    GKLB_V3(32, 5, 8, 11),
};
constexpr int kNumV3 = (int)(sizeof(kV3) / sizeof(kV3[0]));
size_t v3_smem(const V3Config& c, size_t col_pitch) {
  return (size_t)c.warps * (((7 * col_pitch + 15) & ~(size_t)15) + (size_t)(3 + c.ids) * c.K * 32 * sizeof(double) + 64);
}
constexpr int kMaxQual = 254;
constexpr int kMmSizePd = ((kMaxQual + 1) * (kMaxQual + 2)) >> 1;
constexpr int kSmemMax = 232448;

struct PdTables {
  double q2err[kMaxQual + 1];
  double mm[kMmSizePd];
  double init_cond, log10_init;
};

// ProbabilityCache::initialize + JacobianLogTable (pdhmm-common.h:149-195, MathUtils.cc:31-109), host libm
const PdTables& pd_tables() {
  static const PdTables* t = [] {
    PdTables* p = new PdTables;
    std::vector<double> jac(80001);
    for (int k = 0; k < 80001; k++) jac[k] = log10(1.0 + pow(10.0, -k * 0.0001));
    const double inv_ln10 = 1.0 / log(10);
    for (int i = 0, off = 0; i <= kMaxQual; off += ++i)
      for (int j = 0; j <= i; j++) {
        double a = -0.1 * i, b = -0.1 * j;
        if (a > b) std::swap(a, b);
        const double diff = b - a;
        double ls = b;
        if (diff < 8.0) {
          const double v = diff * (1.0 / 0.0001);
          ls = b + jac[(v > 0.0) ? (int)(v + 0.5) : (int)(v - 0.5)];
        }
        p->mm[off + j] = pow(10, log1p(-std::min(1.0, pow(10, ls))) * inv_ln10);
      }
    for (int i = 0; i <= kMaxQual; i++) p->q2err[i] = pow(10.0, ((double)i) / -10.0);
    p->init_cond = pow(2, 1020);
    p->log10_init = log10(p->init_cond);
    return p;
  }();
  return *t;
}

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, n + n / 4 + 256);
    if (e == cudaSuccess) cap = n + n / 4 + 256;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct PdEngine {
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  Buf tables, hap, pd, rd[5], hl, rl, out, misc, carry, deferred;
  int carry_state = 1;
  PdhmmParams last{};
  bool have_last = false;
  size_t last_smem = 0;
  int last_grid = 0;
  // k_pdhmm2 (single pass, haplotype-major tasks): reads per task, blocks per haplotype, tasks
  bool use_v2 = false, allow_v2 = true;
  int v2 = 0;  // index into kV2
  // k_pdhmm3 takes the haplotypes whose rows all start NORMAL, k_pdhmm2 the deferred rest (same task space)
  bool use_v3 = false, allow_v3 = true;
  int v3 = 0;  // index into kV3
  bool wfold = false;  // the variant with the folded insertion state (wfold_ok)
  int n_deferred = 0;
  size_t v3_smem_bytes = 0;
  int v3_grid = 0;
  int read_block2 = 1, n_blocks2 = 1;   // the deferred haplotypes: small read blocks, every resident warp gets some
  unsigned int n_tasks2 = 0;
  int read_block = 1, n_blocks = 1;
  unsigned int n_tasks = 0;
  gklb_pdhmm_stats stats{};
};

// The process-global surface is a pool of engines, like the PairHMM one (engine_global.cu): init/done are reference
// counted (the reference's doneNative only frees its DP tables and several IntelPDHMM instances may live side by
// side, pdhmm/IntelPDHMM.cc:43-60,246-249), every compute call borrows an engine of its own, so concurrent Java
// threads do not serialise on one lock, and the last done frees the idle engines.
struct PdSlot { struct PdEngine* e = nullptr; bool busy = false; };
std::mutex g_mu;
std::condition_variable g_cv;
std::vector<PdSlot> g_slots;
int g_refs = 0;
bool g_inited = false;
int g_carry_state = 1;
bool g_allow_v2 = true, g_allow_v3 = true;
int g_device = 0;
PdEngine* g_last = nullptr;   // engine of the last finished compute call: what last_stats / time_runs refer to
gklb_pdhmm_stats g_last_stats{};
constexpr int kMaxPdEngines = 4;

void destroy(PdEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (Buf* b : {&e->tables, &e->hap, &e->pd, &e->rd[0], &e->rd[1], &e->rd[2], &e->rd[3], &e->rd[4], &e->hl, &e->rl,
                 &e->out, &e->misc, &e->carry, &e->deferred})
    b->release();
  for (auto& ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int launch(PdEngine* e) {
  CU(cudaMemsetAsync(e->misc.p, 0, 8, e->stream));
  if (e->use_v2) {
    const uint8_t* only = nullptr;
    if (e->use_v3) {
      // k_pdhmm3 adds the haplotypes with more kinds of columns than its prior table holds to `deferred`, so the
      // second launch always follows (a few microseconds when nothing was deferred)
      uint8_t* deferred = static_cast<uint8_t*>(e->deferred.p);
      void* args3[] = {&e->last, &e->read_block, &e->n_blocks, &e->n_tasks, &deferred};
      CU(cudaLaunchKernel(e->wfold ? kV3[e->v3].fn_wfold : kV3[e->v3].fn, dim3(e->v3_grid), dim3(kV3[e->v3].warps * 32), args3,
                          e->v3_smem_bytes, e->stream));
      e->stats.kernel_launches++;
      only = deferred;
      void* args2[] = {&e->last, &e->read_block2, &e->n_blocks2, &e->n_tasks2, &only};
      CU(cudaLaunchKernel(kV2[e->v2].fn, dim3(e->num_sms), dim3(kV2[e->v2].warps * 32), args2, e->last_smem, e->stream));
      e->stats.kernel_launches++;
      return GKLB_OK;
    }
    void* args2[] = {&e->last, &e->read_block, &e->n_blocks, &e->n_tasks, &only};
    CU(cudaLaunchKernel(kV2[e->v2].fn, dim3(e->last_grid), dim3(kV2[e->v2].warps * 32), args2, e->last_smem, e->stream));
    e->stats.kernel_launches++;
    return GKLB_OK;
  }
  void* args[] = {&e->last};
  const void* fn = (e->last.max_read <= kG * kK) ? reinterpret_cast<const void*>(&k_pdhmm<kG, kK, kWarps, false>)
                                                 : reinterpret_cast<const void*>(&k_pdhmm<kG, kK, kWarps, true>);
  CU(cudaLaunchKernel(fn, dim3(e->last_grid), dim3(kWarps * 32), args, e->last_smem, e->stream));
  e->stats.kernel_launches++;
  return GKLB_OK;
}

// pdhmm-serial.cc:228-277: a read byte that is not one of ACGTacgt, is not 'N' and differs from the haplotype byte
// fails the call with PDHMM_INPUT_DATA_ERROR when it meets a column that carries SNP alleles (and is not 'N').
// Checked on the host: one pass over the read bytes, the haplotypes only when some read holds such a byte.
struct ByteSet {
  uint64_t w[4] = {0, 0, 0, 0};
  void add(uint8_t x) { w[x >> 6] |= 1ull << (x & 63); }
  bool empty() const { return !(w[0] | w[1] | w[2] | w[3]); }
  void merge(const ByteSet& o) { for (int i = 0; i < 4; i++) w[i] |= o.w[i]; }
  // some member of *this differs from some member of o
  bool differs_from(const ByteSet& o) const {
    if (empty() || o.empty()) return false;
    int na = 0, nb = 0;
    for (int i = 0; i < 4; i++) { na += __builtin_popcountll(w[i]); nb += __builtin_popcountll(o.w[i]); }
    if (na > 1 || nb > 1) return true;
    for (int i = 0; i < 4; i++) if (w[i] != o.w[i]) return true;
    return false;
  }
};

bool unexpected_base_at_snp_column(const gklb_pdhmm_batch* b, bool cross, long long n_read_rows, long long n_hap_rows) {
  static const std::array<uint8_t, 256> plain = [] {
    std::array<uint8_t, 256> t{};
    for (const char* c = "ACGTacgtN"; *c; c++) t[(uint8_t)*c] = 1;
    return t;
  }();
  std::vector<long long> odd_reads;
  for (long long r = 0; r < n_read_rows; r++) {
    const uint8_t* x = reinterpret_cast<const uint8_t*>(b->read_bases) + r * (long long)b->max_read;
    const long long len = b->read_lengths[r];
    uint8_t odd = 0;   // branch-free byte compares: the host compiler vectorises this loop
    for (long long i = 0; i < len; i++) {
      const uint8_t u = x[i] & 0xDF;
      odd |= (uint8_t)!((u == 'A') | (u == 'C') | (u == 'G') | (u == 'T') | (x[i] == 'N'));
    }
    if (odd) odd_reads.push_back(r);
  }
  if (odd_reads.empty()) return false;
  auto odd_bytes = [&](long long r) {
    ByteSet s;
    const uint8_t* x = reinterpret_cast<const uint8_t*>(b->read_bases) + r * (long long)b->max_read;
    for (long long i = 0; i < b->read_lengths[r]; i++)
      if (!plain[x[i]]) s.add(x[i]);
    return s;
  };
  auto snp_bytes = [&](long long h) {
    ByteSet s;
    const uint8_t* y = reinterpret_cast<const uint8_t*>(b->hap_bases) + h * (long long)b->max_hap;
    const int8_t* f = b->hap_pdbases + h * (long long)b->max_hap;
    for (long long j = 0; j < b->hap_lengths[h]; j++)
      if ((f[j] & 1) && y[j] != 'N') s.add(y[j]);
    return s;
  };
  if (cross) {
    ByteSet reads, haps;
    for (long long r : odd_reads) reads.merge(odd_bytes(r));
    for (long long h = 0; h < n_hap_rows; h++) haps.merge(snp_bytes(h));
    return reads.differs_from(haps);
  }
  for (long long k : odd_reads)
    if (odd_bytes(k).differs_from(snp_bytes(k))) return true;
  return false;
}

// May k_pdhmm3 fold the insertion state (pdhmm_device.cuh, WFOLD)?  Per read: the insertion qualities span at most
// 30 dB (the folded state's scale crosses rows by that ratio), and the likelihood's lower bound -- one (mis)match, one
// gap open, gap extensions for the other rows, whatever the haplotype -- stays some 40 decades above the range the
// lowered initial condition gives up (results under about -607).  One pass over three of the read arrays.
bool wfold_ok(const gklb_pdhmm_batch* b, long long n_read_rows) {
  if (const char* w = getenv("GKLB_PDHMM_WFOLD"))
    if (!strcmp(w, "0")) return false;
  for (long long r = 0; r < n_read_rows; r++) {
    const long long len = b->read_lengths[r], o = r * (long long)b->max_read;
    const uint8_t* iq = reinterpret_cast<const uint8_t*>(b->read_ins_qual) + o;
    const uint8_t* gq = reinterpret_cast<const uint8_t*>(b->gcp) + o;
    const uint8_t* qq = reinterpret_cast<const uint8_t*>(b->read_qual) + o;
    uint8_t mn = 255, mx = 0, qmax = 0;
    unsigned gsum = 0;
    for (long long i = 0; i < len; i++) {
      mn = std::min(mn, iq[i]);
      mx = std::max(mx, iq[i]);
      qmax = std::max(qmax, qq[i]);
      gsum += gq[i];
    }
    if (mx - mn > 30 || mx > 127) return false;                      // bytes above 127 are negative qualities: error path
    if (gsum + (unsigned)qmax + (unsigned)mx + 80u > 5600u) return false;   // decades * 10, incl. log10(3 H)
  }
  return true;
}

// n_reads/n_haps > 0: cross layout; else flat with b->n pairs.
int compute(PdEngine* e, const gklb_pdhmm_batch* b, int n_reads, int n_haps, double* out) {
  if (!b || !out) return gklb_internal_fail(GKLB_ERR_INVALID, "null argument");
  const bool cross = n_haps > 0;
  const long long n = cross ? (long long)n_reads * n_haps : (long long)b->n;
  const long long n_hap_rows = cross ? n_haps : n, n_read_rows = cross ? n_reads : n;
  if (n <= 0) return gklb_internal_fail(GKLB_ERR_INVALID, "batchSize must be greater than 0");
  if (b->max_hap <= 0 || b->max_read <= 0)
    return gklb_internal_fail(GKLB_ERR_INVALID, "maxHapLength / maxReadLength must be greater than 0");
  if (!b->hap_bases || !b->hap_pdbases || !b->read_bases || !b->read_qual || !b->read_ins_qual || !b->read_del_qual ||
      !b->gcp || !b->hap_lengths || !b->read_lengths)
    return gklb_internal_fail(GKLB_ERR_INVALID, "null array in batch");
  long long cells = 0, sum_h = 0, sum_r = 0;
  for (long long k = 0; k < n_hap_rows; k++) {
    if (b->hap_lengths[k] <= 0 || b->hap_lengths[k] > b->max_hap)
      return gklb_internal_fail(GKLB_ERR_INVALID, "haplotype %lld has length %lld outside 1..%d", k,
                                (long long)b->hap_lengths[k], b->max_hap);
    sum_h += b->hap_lengths[k];
  }
  for (long long k = 0; k < n_read_rows; k++) {
    if (b->read_lengths[k] <= 0 || b->read_lengths[k] > b->max_read)
      return gklb_internal_fail(GKLB_ERR_INVALID, "read %lld has length %lld outside 1..%d", k,
                                (long long)b->read_lengths[k], b->max_read);
    sum_r += b->read_lengths[k];
  }
  if (cross) cells = sum_h * sum_r;
  else for (long long k = 0; k < n; k++) cells += b->hap_lengths[k] * b->read_lengths[k];
  if (unexpected_base_at_snp_column(b, cross, n_read_rows, n_hap_rows))
    return gklb_internal_fail(GKLB_ERR_INVALID, "Found unexpected base in alt alleles");

  constexpr int gpw = 32 / kG;
  const size_t col_pitch = ((size_t)b->max_hap + 2 * kPdMargin + 1) & ~(size_t)1;
  size_t smem = (size_t)kWarps * gpw * 7 * col_pitch;
  if (smem > (size_t)kSmemMax)
    return gklb_internal_fail(GKLB_ERR_INVALID, "maxHapLength %d does not fit in shared memory", b->max_hap);

  CU(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  CU(cudaEventRecord(e->ev[0], s));
  const size_t hb = (size_t)n_hap_rows * b->max_hap, rb = (size_t)n_read_rows * b->max_read;
  CU(e->hap.ensure(hb));
  CU(e->pd.ensure(hb));
  for (auto& r : e->rd) CU(r.ensure(rb));
  CU(e->hl.ensure(sizeof(int64_t) * n_hap_rows));
  CU(e->rl.ensure(sizeof(int64_t) * n_read_rows));
  CU(e->out.ensure(sizeof(double) * n));
  CU(e->misc.ensure(64));
  const size_t carry_stride = (size_t)gpw * 12 * ((size_t)b->max_hap + 2);
  CU(e->carry.ensure(sizeof(double) * carry_stride * kWarps * e->num_sms));
  CU(cudaMemcpyAsync(e->hap.p, b->hap_bases, hb, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(e->pd.p, b->hap_pdbases, hb, cudaMemcpyHostToDevice, s));
  const int8_t* src[5] = {b->read_bases, b->read_qual, b->read_ins_qual, b->read_del_qual, b->gcp};
  for (int i = 0; i < 5; i++) CU(cudaMemcpyAsync(e->rd[i].p, src[i], rb, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(e->hl.p, b->hap_lengths, sizeof(int64_t) * n_hap_rows, cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(e->rl.p, b->read_lengths, sizeof(int64_t) * n_read_rows, cudaMemcpyHostToDevice, s));

  const PdTables& t = pd_tables();
  PdhmmParams& p = e->last;
  p.hap_bases = static_cast<const int8_t*>(e->hap.p);
  p.hap_pdbases = static_cast<const int8_t*>(e->pd.p);
  p.read_bases = static_cast<const int8_t*>(e->rd[0].p);
  p.read_qual = static_cast<const int8_t*>(e->rd[1].p);
  p.read_ins_qual = static_cast<const int8_t*>(e->rd[2].p);
  p.read_del_qual = static_cast<const int8_t*>(e->rd[3].p);
  p.gcp = static_cast<const int8_t*>(e->rd[4].p);
  p.hap_lengths = static_cast<const int64_t*>(e->hl.p);
  p.read_lengths = static_cast<const int64_t*>(e->rl.p);
  p.n = n;
  p.n_haps = cross ? n_haps : 0;
  p.max_hap = b->max_hap;
  p.max_read = b->max_read;
  p.q2err = static_cast<const double*>(e->tables.p);
  p.mm = p.q2err + (kMaxQual + 1);
  p.init_cond = t.init_cond;
  p.log10_init = t.log10_init;
  p.out = static_cast<double*>(e->out.p);
  p.counter = static_cast<unsigned int*>(e->misc.p);
  p.error_flag = p.counter + 1;
  p.carry = static_cast<double*>(e->carry.p);
  p.carry_stride = carry_stride;
  p.carry_state = e->carry_state;
  e->last_smem = smem;
  // Reads that fit one pass take the haplotype-major kernel (the smallest rows-per-lane variant that covers them,
  // if its column tables fit in shared memory): a task is one haplotype x a block of reads, sized for about 16 tasks
  // per resident warp so that the dynamic queue balances the load.  Everything else: one pair per group, k_pdhmm.
  long long warp_items = (n + gpw - 1) / gpw;
  int warps_per_cta = kWarps;
  e->use_v2 = false;
  for (int i = kNumV2 - 1; i >= 0 && e->allow_v2; i--)
    if (b->max_read <= kV2[i].max_read && (size_t)kV2[i].warps * 7 * col_pitch <= (size_t)kSmemMax) {
      e->use_v2 = true;
      e->v2 = i;
    }
  if (e->use_v2) {
    const int w2 = kV2[e->v2].warps;
    long long tasks = n;
    e->read_block = 1;
    e->n_blocks = 1;
    if (cross) {
      const long long want = 16LL * w2 * e->num_sms;
      const long long nb = std::min<long long>(n_reads, std::max<long long>(1, (want + n_haps - 1) / n_haps));
      e->read_block = (int)((n_reads + nb - 1) / nb);
      e->n_blocks = (int)((n_reads + e->read_block - 1) / e->read_block);
      tasks = (long long)e->n_blocks * n_haps;
    }
    // Two reads per warp (k_pdhmm3) when every read fits 16 x 7 rows: the haplotypes whose last column leaves the state
    // machine outside NORMAL (the state is carried into the next row, pdhmm-serial.cc:306,370-385) stay with k_pdhmm2.
    e->use_v3 = false;
    e->n_deferred = 0;
    int v3 = -1;
    for (int i = kNumV3 - 1; i >= 0; i--)
      if (b->max_read <= kV3[i].max_read && v3_smem(kV3[i], col_pitch) <= (size_t)kSmemMax) v3 = i;
    if (const char* force = getenv("GKLB_PDHMM_V3")) {   // measurement: a given instantiation, if it fits
      const int i = atoi(force);
      if (i >= 0 && i < kNumV3 && b->max_read <= kV3[i].max_read && v3_smem(kV3[i], col_pitch) <= (size_t)kSmemMax) v3 = i;
    }
    if (cross && e->allow_v3 && v3 >= 0 && tasks <= 0xFFFFFFF0LL) {
      e->v3 = v3;
      const int reads_per_warp = 32 / kV3[v3].G;
      std::vector<uint8_t> deferred((size_t)n_haps, 0);
      for (int h = 0; h < n_haps && e->carry_state; h++) {
        const int8_t* f = b->hap_pdbases + (size_t)h * b->max_hap;
        int st = 0;
        for (long long c = 0; c < b->hap_lengths[h]; c++) {
          if (st == 2) st = 0;
          if (f[c] & 2) st = 1;
          if (f[c] & 4) st = 2;
        }
        deferred[h] = st != 0;
        e->n_deferred += st != 0;
      }
      CU(e->deferred.ensure((size_t)n_haps));
      CU(cudaMemcpyAsync(e->deferred.p, deferred.data(), (size_t)n_haps, cudaMemcpyHostToDevice, s));
      CU(cudaStreamSynchronize(s));   // `deferred` is a local
      const int w3 = kV3[v3].warps;
      const long long want = 16LL * w3 * e->num_sms;
      const long long groups = (n_reads + reads_per_warp - 1) / reads_per_warp;
      const long long nb = std::min<long long>(groups, std::max<long long>(1, (want + n_haps - 1) / n_haps));
      e->read_block = (int)((n_reads + nb - 1) / nb);
      e->read_block = (e->read_block + reads_per_warp - 1) / reads_per_warp * reads_per_warp;   // whole warps
      e->n_blocks = (int)((n_reads + e->read_block - 1) / e->read_block);
      tasks = (long long)e->n_blocks * n_haps;
      e->use_v3 = true;
      e->wfold = wfold_ok(b, n_read_rows);
      {
        const long long resident = (long long)kV2[e->v2].warps * e->num_sms;
        long long rb2 = std::max<long long>(1, (long long)n_reads * std::max(1, e->n_deferred) / (2 * resident));
        while ((n_reads + rb2 - 1) / rb2 * n_haps > 0x7FFFFFF0LL) rb2 *= 2;
        e->read_block2 = (int)rb2;
        e->n_blocks2 = (int)((n_reads + rb2 - 1) / rb2);
        e->n_tasks2 = (unsigned int)((long long)e->n_blocks2 * n_haps);
      }
      e->v3_smem_bytes = v3_smem(kV3[v3], col_pitch);
      e->v3_grid = (int)std::min<long long>(e->num_sms, (tasks + w3 - 1) / w3);
    }
    if (tasks > 0xFFFFFFF0LL) {
      e->use_v2 = false;  // the task counter is 32 bits wide
    } else {
      e->n_tasks = (unsigned int)tasks;
      warp_items = tasks;
      warps_per_cta = w2;
      smem = (size_t)w2 * 7 * col_pitch;
    }
  }
  e->last_smem = smem;
  e->last_grid = (int)std::min<long long>(e->num_sms, (warp_items + warps_per_cta - 1) / warps_per_cta);
  e->have_last = true;
  e->stats = gklb_pdhmm_stats{};
  e->stats.pairs = n;
  e->stats.cells = cells;

  CU(cudaEventRecord(e->ev[1], s));
  int rc = launch(e);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev[2], s));
  unsigned int flags[2] = {0, 0};
  CU(cudaMemcpyAsync(out, e->out.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(flags, e->misc.p, sizeof(flags), cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(e->ev[3], s));
  CU(cudaStreamSynchronize(s));
  cudaEventElapsedTime(&e->stats.h2d_ms, e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&e->stats.kernel_ms, e->ev[1], e->ev[2]);
  cudaEventElapsedTime(&e->stats.d2h_ms, e->ev[2], e->ev[3]);
  if (flags[1] & 1u)  // pdhmm-serial.cc:184-198 -> PDHMM_INPUT_DATA_ERROR -> IllegalArgumentException
    return gklb_internal_fail(GKLB_ERR_INVALID, "insertion, deletion or gcp quality is negative");
  if (flags[1] & 2u)  // pdhmm-serial.cc:432-441 -> PDHMM_FAILURE -> RuntimeException
    return gklb_internal_fail(GKLB_ERR_CUDA, "PDHMM log probability is greater than 0 or not a number");
  return GKLB_OK;
}

}  // namespace

namespace {

int create_pd_engine(PdEngine** out) {
  int n = 0;
  cudaError_t ce = cudaGetDeviceCount(&n);
  if (ce != cudaSuccess || n <= 0) return gklb_internal_fail(GKLB_ERR_NO_DEVICE, "no CUDA device (%s)", cudaGetErrorString(ce));
  const int device = g_device;
  if (device < 0 || device >= n) return gklb_internal_fail(GKLB_ERR_NO_DEVICE, "device %d out of range", device);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return gklb_internal_fail(GKLB_ERR_NO_DEVICE, "device %d is sm_%d%d; sm_100a code only", device, prop.major, prop.minor);
  CU(cudaSetDevice(device));
  PdEngine* e = new PdEngine;
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  e->carry_state = g_carry_state;
  e->allow_v2 = g_allow_v2;
  e->allow_v3 = g_allow_v3;
  CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev) CU(cudaEventCreate(&ev));
  CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(&k_pdhmm<kG, kK, kWarps, false>),
                          cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  CU(cudaFuncSetAttribute(reinterpret_cast<const void*>(&k_pdhmm<kG, kK, kWarps, true>),
                          cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  for (const V2Config& c : kV2) CU(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  for (const V3Config& c : kV3) {
    CU(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    CU(cudaFuncSetAttribute(c.fn_wfold, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  }
  const PdTables& t = pd_tables();
  CU(e->tables.ensure(sizeof(double) * (kMaxQual + 1 + kMmSizePd)));
  CU(cudaMemcpy(e->tables.p, t.q2err, sizeof(t.q2err), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(static_cast<double*>(e->tables.p) + (kMaxQual + 1), t.mm, sizeof(t.mm), cudaMemcpyHostToDevice));
  *out = e;
  return GKLB_OK;
}

int acquire_pd(PdEngine** out) {
  std::unique_lock<std::mutex> lk(g_mu);
  if (!g_inited) return gklb_internal_fail(GKLB_ERR_STATE, "gklb_pdhmm_init has not been called");
  for (;;) {
    for (auto& s : g_slots)
      if (!s.busy && s.e) { s.busy = true; *out = s.e; return GKLB_OK; }
    if ((int)g_slots.size() < kMaxPdEngines) {
      g_slots.push_back(PdSlot{nullptr, true});
      const size_t idx = g_slots.size() - 1;
      lk.unlock();
      PdEngine* e = nullptr;
      const int rc = create_pd_engine(&e);
      lk.lock();
      if (rc) {
        g_slots.erase(g_slots.begin() + idx);
        g_cv.notify_all();
        return rc;
      }
      g_slots[idx].e = e;
      *out = e;
      return GKLB_OK;
    }
    g_cv.wait(lk);
  }
}

void release_pd(PdEngine* e) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& s : g_slots)
    if (s.e == e) s.busy = false;
  g_last = e;
  g_last_stats = e->stats;
  g_cv.notify_all();
}

}  // namespace

extern "C" {

int gklb_pdhmm_init(int openmp_setting, int max_threads, int avx_level, int max_memory_mb) {
  (void)openmp_setting; (void)max_threads; (void)avx_level; (void)max_memory_mb;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    const char* dev = getenv("GKLB_DEVICE");
    const int device = dev ? atoi(dev) : 0;
    const char* rs = getenv("GKLB_PDHMM_ROW_STATE");
    const int carry = (rs && !strcmp(rs, "reset")) ? 0 : 1;
    // measurement knob: "1" = the pair-at-a-time kernel for every batch, "2" = no two-reads-per-warp kernel
    const char* kv = getenv("GKLB_PDHMM_KERNEL");
    const bool allow_v2 = !(kv && !strcmp(kv, "1"));
    const bool allow_v3 = allow_v2 && !(kv && !strcmp(kv, "2"));
    if (g_inited && (device != g_device || carry != g_carry_state || allow_v2 != g_allow_v2 || allow_v3 != g_allow_v3)) {
      // a different configuration: idle engines are rebuilt lazily with it
      for (size_t i = 0; i < g_slots.size();)
        if (!g_slots[i].busy) { if (g_last == g_slots[i].e) g_last = nullptr; destroy(g_slots[i].e); g_slots.erase(g_slots.begin() + i); } else i++;
    }
    g_device = device;
    g_carry_state = carry;
    g_allow_v2 = allow_v2;
    g_allow_v3 = allow_v3;
    g_inited = true;
    g_refs++;
  }
  PdEngine* e = nullptr;  // one engine now: a missing device must fail initialize(), not the first compute
  const int rc = acquire_pd(&e);
  if (rc) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (--g_refs == 0) g_inited = false;
    return rc;
  }
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& s : g_slots)
      if (s.e == e) s.busy = false;
    g_cv.notify_all();
  }
  return GKLB_OK;
}

int gklb_pdhmm_compute(const gklb_pdhmm_batch* batch, double* likelihoods) {
  PdEngine* e = nullptr;
  int rc = acquire_pd(&e);
  if (rc) return rc;
  rc = compute(e, batch, 0, 0, likelihoods);
  release_pd(e);
  return rc;
}

int gklb_pdhmm_compute_cross(const gklb_pdhmm_batch* operands, int32_t n_reads, int32_t n_haps, double* likelihoods) {
  if (n_reads <= 0 || n_haps <= 0) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) return gklb_internal_fail(GKLB_ERR_STATE, "gklb_pdhmm_init has not been called");
    return gklb_internal_fail(GKLB_ERR_INVALID, "empty read or haplotype array");
  }
  PdEngine* e = nullptr;
  int rc = acquire_pd(&e);
  if (rc) return rc;
  rc = compute(e, operands, n_reads, n_haps, likelihoods);
  release_pd(e);
  return rc;
}

int gklb_pdhmm_done(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_refs > 0) g_refs--;
  if (g_refs == 0) {
    for (size_t i = 0; i < g_slots.size();)
      if (!g_slots[i].busy) { if (g_last == g_slots[i].e) g_last = nullptr; destroy(g_slots[i].e); g_slots.erase(g_slots.begin() + i); } else i++;
  }
  return GKLB_OK;
}

int gklb_pdhmm_last_stats(gklb_pdhmm_stats* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out) return gklb_internal_fail(GKLB_ERR_INVALID, "out is null");
  *out = g_last_stats;
  return GKLB_OK;
}

const char* gklb_pdhmm_kernel_name(void) {
  static thread_local char name[48];
  std::lock_guard<std::mutex> lk(g_mu);
  const PdEngine* e = g_last;
  if (!e || !e->have_last) return "";
  if (e->use_v2 && e->use_v3)
    snprintf(name, sizeof(name), "k_pdhmm3<%d,%d,%d%s>", kV3[e->v3].G, kV3[e->v3].K, kV3[e->v3].warps, e->wfold ? ",wfold" : "");
  else if (e->use_v2) snprintf(name, sizeof(name), "k_pdhmm2<%d,%d>", kV2[e->v2].K, kV2[e->v2].warps);
  else snprintf(name, sizeof(name), "k_pdhmm<%d,%d,%d>", kG, kK, kWarps);
  return name;
}

int gklb_pdhmm_time_runs(int iters, float* ms_per_run) {
  PdEngine* e = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& s : g_slots)
      if (s.e == g_last && g_last && !s.busy) { s.busy = true; e = s.e; }
  }
  if (!e || !e->have_last || iters <= 0 || !ms_per_run) {
    if (e) release_pd(e);
    return gklb_internal_fail(GKLB_ERR_STATE, "nothing to time");
  }
  int rc = GKLB_OK;
  float ms = 0;
  do {
    if (cudaSetDevice(e->device) != cudaSuccess || cudaEventRecord(e->ev[0], e->stream) != cudaSuccess) { rc = GKLB_ERR_CUDA; break; }
    for (int i = 0; i < iters && rc == GKLB_OK; i++) rc = launch(e);
    if (rc) break;
    if (cudaEventRecord(e->ev[1], e->stream) != cudaSuccess || cudaEventSynchronize(e->ev[1]) != cudaSuccess) { rc = GKLB_ERR_CUDA; break; }
    cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]);
  } while (0);
  release_pd(e);
  if (rc) return gklb_internal_fail(rc, "timing run failed");
  *ms_per_run = ms / iters;
  return GKLB_OK;
}

const void* gklb_pdhmm_table(int which, int* n) {
  const PdTables& t = pd_tables();
  if (which == 0) { if (n) *n = kMaxQual + 1; return t.q2err; }
  if (which == 1) { if (n) *n = kMmSizePd; return t.mm; }
  if (n) *n = 0;
  return nullptr;
}

}  // extern "C"
