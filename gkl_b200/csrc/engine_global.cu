// engine_global.cu -- the process-global surface of the PairHMM engine: what the JNI layer (and GKL's Java
// shim above it) sees.  Replaces the native state of pairhmm/IntelPairHmm.cc:41-48 (g_use_double, g_max_threads,
// function pointers) and the bodies of initNative / computeLikelihoodsNative / doneNative (:55-118,125-181,189-192).
//
// GKL's native state is read-only after init, so several IntelPairHmm instances and several Java threads (Spark
// executors) may call computeLikelihoods concurrently, and doneNative is empty.  Here the state is a POOL of
// engines: every compute call borrows an idle engine (or creates one, up to GKLB_ENGINES_PER_DEVICE per device),
// on the least busy of the configured devices, so concurrent callers run on different GPUs when there are
// several and never wait on one global lock.  init/done are reference counted: done of the last instance frees
// the idle engines' device memory, and a later compute simply creates engines again.
// Large batches (more than ~4e9 cells per device) are sharded over reads across the configured devices inside
// the one call (GKLB_SHARD=direct: one host thread per device, each device copies its shard over its own PCIe
// link and writes its slab of the caller's array -- a shard of more than ~2e11 cells as pieces that two engines on
// the device take in turn, so that the copies of one piece run under the kernels of the other; GKLB_SHARD=nccl: see
// engine_nccl.cu).  Multi-region calls (gklb_pairhmm_compute_multi) are cut into jobs of bounded size, one per device
// when there is enough work.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <thread>

#include "engine_internal.h"

using namespace gklb;

namespace {

struct Slot {
  gklb_engine* e = nullptr;
  int device = 0;
  bool use_double = false;
  bool busy = false;
};

std::mutex g_mu;
std::condition_variable g_cv;
bool g_inited = false;
int g_refs = 0;
bool g_use_double = false;
std::vector<int> g_devices;  // as configured (a device may be listed twice: two shards on one GPU)
std::vector<Slot*> g_pool;
unsigned g_rr = 0;
gklb_pairhmm_stats g_last_stats{};

int max_engines_per_device() {
  static const int v = [] {
    const char* s = getenv("GKLB_ENGINES_PER_DEVICE");
    return s && atoi(s) > 0 ? atoi(s) : 4;
  }();
  return v;
}

std::vector<int> configured_devices() {
  // GKLB_DEVICES = "all" | "0,1,2,..." ; GKLB_DEVICE = n : a single device (default 0)
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
  return devices;
}

// Borrow an engine.  device_hint < 0: the least busy configured device.
// wait = false: give up (GKLB_ERR_STATE, no message) instead of waiting for an engine of a fully booked device
int acquire(int device_hint, gklb_engine** out, bool wait = true) {
  std::unique_lock<std::mutex> lk(g_mu);
  if (!g_inited) return fail(GKLB_ERR_STATE, "gklb_pairhmm_init has not been called");
  for (;;) {
    int d = device_hint;
    if (d < 0) {
      int best = -1, best_busy = 1 << 30;
      const size_t n = g_devices.size();
      for (size_t k = 0; k < n; k++) {
        const int cand = g_devices[(g_rr + k) % n];
        int busy = 0;
        for (Slot* s : g_pool) busy += (s->device == cand && s->busy);
        if (busy < best_busy) { best_busy = busy; best = cand; }
      }
      g_rr++;
      d = best;
    }
    Slot* idle_other = nullptr;
    int on_device = 0;
    for (Slot* s : g_pool) {
      if (s->device != d) continue;
      on_device++;
      if (s->busy) continue;
      if (s->use_double == g_use_double) {
        s->busy = true;
        *out = s->e;
        return GKLB_OK;
      }
      idle_other = s;
    }
    if (on_device < max_engines_per_device() || idle_other) {
      // create outside the lock (context creation takes a while); an idle engine of the other precision is replaced
      Slot* s = idle_other;
      gklb_engine* old = nullptr;
      const bool use_double = g_use_double;
      if (s) { old = s->e; s->e = nullptr; } else { s = new Slot; s->device = d; g_pool.push_back(s); }
      s->busy = true;
      s->use_double = use_double;
      lk.unlock();
      if (old) destroy_engine(old);
      gklb_engine* e = nullptr;
      const int rc = create_engine(&e, d, use_double);
      lk.lock();
      if (rc) {
        g_pool.erase(std::find(g_pool.begin(), g_pool.end(), s));
        delete s;
        g_cv.notify_all();
        return rc;
      }
      s->e = e;
      *out = e;
      return GKLB_OK;
    }
    if (!wait) return GKLB_ERR_STATE;
    g_cv.wait(lk);
  }
}

void release(gklb_engine* e) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (Slot* s : g_pool)
    if (s->e == e) s->busy = false;
  g_cv.notify_all();
}

void free_idle_engines_locked() {
  for (size_t i = 0; i < g_pool.size();) {
    Slot* s = g_pool[i];
    if (!s->busy) {
      destroy_engine(s->e);
      delete s;
      g_pool.erase(g_pool.begin() + i);
    } else {
      i++;
    }
  }
}

// Batches below this many cells per device are not worth splitting: one GPU finishes them in about a millisecond.
const long long kMinCellsPerDevice = 4000000000LL;
// direct sharding: a device's shard is cut into pieces of about this many cells (~30 ms of kernels), at most kMaxPieces,
// handled by two engines in turn so that the copies of one piece overlap the kernels of the other
const long long kCellsPerPiece = 100000000000LL;
const int kMaxPieces = 12;

}  // namespace

namespace gklb {
int sharded_compute_nccl(const std::vector<gklb_engine*>& engines, const gklb_pairhmm_batch* batch, const std::vector<int>& cut,
                         double* likelihoods, gklb_pairhmm_stats* stats);
bool nccl_available();
}  // namespace gklb

extern "C" {

int gklb_pairhmm_init(int use_double, int max_threads) {
  (void)max_threads;  // GKL's non-OpenMP library ignores it as well (IntelPairHmm.cc:85-89)
  std::vector<int> devices = configured_devices();
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_devices = devices;
    g_use_double = use_double != 0;
    g_inited = true;
    g_refs++;
  }
  // one engine per configured device now, so that a missing or unsupported device fails initialize(), not the
  // first computeLikelihoods (create_engine is the only place that can say GKLB_ERR_NO_DEVICE)
  std::vector<int> distinct = devices;
  std::sort(distinct.begin(), distinct.end());
  distinct.erase(std::unique(distinct.begin(), distinct.end()), distinct.end());
  for (int d : distinct) {
    gklb_engine* e = nullptr;
    const int rc = acquire(d, &e);
    if (rc) {
      std::lock_guard<std::mutex> lk(g_mu);
      g_refs--;
      if (g_refs == 0) { free_idle_engines_locked(); g_inited = false; }
      return rc;
    }
    release(e);
  }
  return GKLB_OK;
}

int gklb_pairhmm_compute(const gklb_pairhmm_batch* batch, double* likelihoods) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) return fail(GKLB_ERR_STATE, "gklb_pairhmm_init has not been called");
  }
  int rc = validate_batch(batch);
  if (rc) return rc;
  std::vector<int> devices;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    devices = g_devices;
  }
  int n_dev = (int)devices.size();
  if (batch->n_reads > 0 && batch->n_haps > 0) {
    const long long cells = (long long)batch->read_off[batch->n_reads] * (long long)batch->hap_off[batch->n_haps];
    n_dev = (int)std::max(1LL, std::min<long long>(n_dev, cells / kMinCellsPerDevice));
    n_dev = std::min(n_dev, batch->n_reads);
  } else {
    n_dev = 1;
  }
  long long batch_cells = 0;
  if (batch->n_reads > 0 && batch->n_haps > 0)
    batch_cells = (long long)batch->read_off[batch->n_reads] * (long long)batch->hap_off[batch->n_haps];
  const char* mode = getenv("GKLB_SHARD");
  const bool nccl = mode && !strcmp(mode, "nccl");
  if (n_dev <= 1 && (nccl || batch_cells < 2 * kCellsPerPiece)) {
    gklb_engine* e = nullptr;
    if ((rc = acquire(-1, &e))) return rc;
    {
      std::lock_guard<std::mutex> lk(e->mu);
      rc = do_compute(e, batch, 1, &likelihoods);
    }
    {
      std::lock_guard<std::mutex> lk(g_mu);
      g_last_stats = e->stats;
    }
    release(e);
    return rc;
  }
  // Reads are sharded into contiguous ranges balanced by total length (every read meets every haplotype, so
  // cells are proportional to read length); range g gets a contiguous slab of the read-major output
  // (JavaData.h:94-105).
  auto split = [&](int lo, int hi, int parts, std::vector<int>* cuts) {   // [lo, hi) into `parts` ranges of equal length
    cuts->assign((size_t)parts + 1, lo);
    const int64_t base = batch->read_off[lo], span = batch->read_off[hi] - base;
    int r = lo;
    for (int g = 1; g < parts; g++) {
      const int64_t target = base + span * g / parts;
      while (r < hi && batch->read_off[r] < target) r++;
      (*cuts)[g] = std::max(r, (*cuts)[g - 1]);
    }
    (*cuts)[parts] = hi;
  };
  std::vector<int> cut;
  split(0, batch->n_reads, n_dev, &cut);
  std::vector<gklb_engine*> engines(n_dev, nullptr);
  for (int g = 0; g < n_dev; g++) {
    if ((rc = acquire(n_dev == 1 ? -1 : devices[g], &engines[g]))) {
      for (int k = 0; k < g; k++) release(engines[k]);
      return rc;
    }
  }
  gklb_pairhmm_stats total_stats{};
  if (nccl) {
    rc = sharded_compute_nccl(engines, batch, cut, likelihoods, &total_stats);
  } else {
    // direct: each device copies its shard and the panel over its own PCIe link and writes its slab straight into
    // the caller's array; the shards never meet on one GPU.  A shard of more than two pieces' worth of work is cut into
    // pieces that two host threads, each with an engine of its own on that device, take in turn: the copies of one
    // piece (in and out, pageable memory on the caller's side) run under the kernels of the other.
    const int64_t hap_total = batch->hap_off[batch->n_haps];
    std::vector<std::vector<int>> pieces(n_dev);
    std::vector<gklb_engine*> second(n_dev, nullptr);
    for (int g = 0; g < n_dev; g++) {
      const long long cells_g = (long long)(batch->read_off[cut[g + 1]] - batch->read_off[cut[g]]) * hap_total;
      int np = (int)std::max(1LL, std::min<long long>(kMaxPieces, cells_g / kCellsPerPiece));
      np = std::max(1, std::min(np, cut[g + 1] - cut[g]));
      // never wait for the second engine while holding the first (concurrent callers would deadlock): without it
      // the shard runs as one piece
      if (np >= 2 && acquire(gklb_engine_device(engines[g]), &second[g], false) != GKLB_OK) { second[g] = nullptr; np = 1; }
      split(cut[g], cut[g + 1], np, &pieces[g]);
    }
    // k_first / k_last: when the device's first kernel could start and its last kernel ended, against an event recorded
    // on the device before the first piece -- with two engines taking turns the per-piece kernel times overlap (a
    // piece's kernels wait for the other engine's), so the device's kernel phase is the span, not the sum
    struct DevAcc {
      std::mutex mu; gklb_pairhmm_stats st{}; int rc = GKLB_OK; std::string err;
      cudaEvent_t origin = nullptr; float k_first = 1e30f, k_last = 0.f;
    };
    std::vector<DevAcc> acc(n_dev);
    for (int g = 0; g < n_dev; g++) {
      if (!second[g]) continue;
      if (cudaSetDevice(gklb_engine_device(engines[g])) != cudaSuccess ||
          cudaEventCreate(&acc[g].origin) != cudaSuccess ||
          cudaEventRecord(acc[g].origin, engines[g]->stream) != cudaSuccess) {
        if (acc[g].origin) cudaEventDestroy(acc[g].origin);
        acc[g].origin = nullptr;
      }
    }
    auto run_pieces = [&](int g, gklb_engine* e, int first, int stride) {
      const int np = (int)pieces[g].size() - 1;
      std::vector<int64_t> offs;
      for (int k = first; k < np; k += stride) {
        const int lo = pieces[g][k], hi = pieces[g][k + 1];
        if (hi == lo) continue;
        const int64_t base = batch->read_off[lo];
        offs.resize((size_t)(hi - lo) + 1);
        for (int r = lo; r <= hi; r++) offs[r - lo] = batch->read_off[r] - base;
        gklb_pairhmm_batch sub = *batch;
        sub.n_reads = hi - lo;
        sub.read_off = offs.data();
        sub.read_bases = batch->read_bases + base;
        sub.read_quals = batch->read_quals + base;
        sub.ins_gop = batch->ins_gop + base;
        sub.del_gop = batch->del_gop + base;
        sub.gcp = batch->gcp + base;
        double* slab = likelihoods + (size_t)lo * batch->n_haps;
        int prc;
        gklb_pairhmm_stats st;
        float k0 = 0.f, k1 = 0.f;
        {
          std::lock_guard<std::mutex> lk(e->mu);
          prc = do_compute(e, &sub, 1, &slab);
          st = e->stats;
          if (!prc && acc[g].origin &&
              (cudaEventElapsedTime(&k0, acc[g].origin, e->ev[1]) != cudaSuccess ||
               cudaEventElapsedTime(&k1, acc[g].origin, e->ev[2]) != cudaSuccess))
            k0 = k1 = 0.f;
        }
        std::lock_guard<std::mutex> lk(acc[g].mu);
        if (k1 > 0.f) { acc[g].k_first = std::min(acc[g].k_first, k0); acc[g].k_last = std::max(acc[g].k_last, k1); }
        if (prc) {   // the last error is thread-local: carry it to the caller's thread
          if (!acc[g].rc) { acc[g].rc = prc; acc[g].err = last_error_string(); }
          return;
        }
        gklb_pairhmm_stats& a = acc[g].st;
        a.pairs += st.pairs; a.cells += st.cells; a.fallback_pairs += st.fallback_pairs; a.fp64_pairs += st.fp64_pairs;
        a.kernel_launches += st.kernel_launches;
        a.n_classes = std::max(a.n_classes, st.n_classes);
        // what is not hidden under another piece's kernels: the first piece's way in, the last piece's way out
        if (k == 0) a.h2d_ms = st.h2d_ms;
        if (k == np - 1) a.d2h_ms = st.d2h_ms;
        a.kernel_ms += st.kernel_ms;   // one piece: its kernel time; several: replaced by the span below
      }
    };
    std::vector<std::thread> th;
    for (int g = 0; g < n_dev; g++) {
      const int workers = second[g] ? 2 : 1;
      th.emplace_back(run_pieces, g, engines[g], 0, workers);
      if (second[g]) th.emplace_back(run_pieces, g, second[g], 1, workers);
    }
    for (auto& t : th) t.join();
    for (auto* e : second)
      if (e) release(e);
    for (int g = 0; g < n_dev; g++) {
      if (!acc[g].origin) continue;
      if (acc[g].k_last > acc[g].k_first) acc[g].st.kernel_ms = acc[g].k_last - acc[g].k_first;
      cudaEventDestroy(acc[g].origin);
    }
    for (int g = 0; g < n_dev && rc == GKLB_OK; g++) {
      if (acc[g].rc) { set_last_error(acc[g].err); rc = acc[g].rc; break; }
      const gklb_pairhmm_stats& st = acc[g].st;
      total_stats.pairs += st.pairs;
      total_stats.cells += st.cells;
      total_stats.fallback_pairs += st.fallback_pairs;
      total_stats.fp64_pairs += st.fp64_pairs;
      total_stats.kernel_launches += st.kernel_launches;
      total_stats.n_classes = std::max(total_stats.n_classes, st.n_classes);
      total_stats.h2d_ms = std::max(total_stats.h2d_ms, st.h2d_ms);
      total_stats.kernel_ms = std::max(total_stats.kernel_ms, st.kernel_ms);
      total_stats.d2h_ms = std::max(total_stats.d2h_ms, st.d2h_ms);
    }
  }
  for (auto* e : engines) release(e);
  if (rc == GKLB_OK) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_last_stats = total_stats;
  }
  return rc;
}

// Several regions (active regions of HaplotypeCaller) in one call: the caller-side coalescing of SURVEY.md 8(f) N2.
// All regions are staged with one host->device copy and served by launches that pull from one task queue per
// launch group (engine.cu: plan_groups), so that a dozen small regions fill the GPU like one large batch;
// likelihoods[r] receives region r's matrix.  Results are bit-identical to one gklb_pairhmm_compute per region.
int gklb_pairhmm_compute_multi(const gklb_pairhmm_batch* batches, int n_batches, double* const* likelihoods) {
  std::vector<int> devices;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_inited) return fail(GKLB_ERR_STATE, "gklb_pairhmm_init has not been called");
    devices = g_devices;
  }
  if (n_batches < 0 || (n_batches > 0 && (!batches || !likelihoods))) return fail(GKLB_ERR_INVALID, "bad region array");
  if (n_batches == 0) return GKLB_OK;
  int rc;
  for (int r = 0; r < n_batches; r++)
    if ((rc = validate_batch(&batches[r]))) return rc;
  // Regions large enough to be sharded over the devices go through gklb_pairhmm_compute one by one; the others are
  // coalesced into jobs of bounded size (the staging buffer is pinned memory), balanced over the configured devices
  // when there is enough work for several, each job on an engine of its own.
  const long long kMaxJobBytes = 192LL << 20;
  std::vector<int> small;
  std::vector<long long> cells((size_t)n_batches, 0), bytes((size_t)n_batches, 0);
  gklb_pairhmm_stats total{};
  for (int r = 0; r < n_batches; r++) {
    const gklb_pairhmm_batch& b = batches[r];
    if (b.n_reads == 0 || b.n_haps == 0) continue;
    cells[r] = (long long)b.read_off[b.n_reads] * (long long)b.hap_off[b.n_haps];
    bytes[r] = 5LL * b.read_off[b.n_reads] + b.hap_off[b.n_haps] + 8LL * b.n_reads * b.n_haps;
    if (cells[r] >= kMinCellsPerDevice || bytes[r] > kMaxJobBytes / 2) {
      if ((rc = gklb_pairhmm_compute(&b, likelihoods[r]))) return rc;
      std::lock_guard<std::mutex> lk(g_mu);
      total.pairs += g_last_stats.pairs; total.cells += g_last_stats.cells; total.fallback_pairs += g_last_stats.fallback_pairs;
      total.fp64_pairs += g_last_stats.fp64_pairs; total.kernel_launches += g_last_stats.kernel_launches;
    } else {
      small.push_back(r);
    }
  }
  if (!small.empty()) {
    long long small_cells = 0, small_bytes = 0;
    for (int r : small) { small_cells += cells[r]; small_bytes += bytes[r]; }
    int n_jobs = (int)std::max<long long>(1, (small_bytes + kMaxJobBytes - 1) / kMaxJobBytes);
    // several devices: one job per device once every job still holds about a millisecond of work
    n_jobs = std::max(n_jobs, (int)std::min<long long>((long long)devices.size(), small_cells / 2000000000LL));
    n_jobs = std::max(1, std::min(n_jobs, (int)small.size()));
    // greedy balance: largest regions first, each to the lightest job (region order inside a job does not matter)
    std::vector<std::vector<int>> jobs((size_t)n_jobs);
    std::vector<long long> load((size_t)n_jobs, 0);
    std::vector<int> order = small;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cells[x] > cells[y]; });
    for (int r : order) {
      const int j = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      jobs[j].push_back(r);
      load[j] += cells[r];
    }
    std::vector<int> rcs((size_t)n_jobs, GKLB_OK);
    std::vector<std::string> errs((size_t)n_jobs);
    std::vector<gklb_pairhmm_stats> sts((size_t)n_jobs);
    auto run_job = [&](int j) {
      std::vector<gklb_pairhmm_batch> jb;
      std::vector<double*> jo;
      for (int r : jobs[j]) { jb.push_back(batches[r]); jo.push_back(likelihoods[r]); }
      gklb_engine* e = nullptr;
      rcs[j] = acquire(-1, &e);
      if (!rcs[j]) {
        {
          std::lock_guard<std::mutex> lk(e->mu);
          rcs[j] = do_compute(e, jb.data(), (int)jb.size(), jo.data());
          sts[j] = e->stats;
        }
        release(e);
      }
      if (rcs[j]) errs[j] = last_error_string();
    };
    if (n_jobs == 1) {
      run_job(0);
    } else {
      std::vector<std::thread> th;
      for (int j = 0; j < n_jobs; j++) th.emplace_back(run_job, j);
      for (auto& t : th) t.join();
    }
    for (int j = 0; j < n_jobs; j++) {
      if (rcs[j]) { set_last_error(errs[j]); return rcs[j]; }
      total.pairs += sts[j].pairs; total.cells += sts[j].cells; total.fallback_pairs += sts[j].fallback_pairs;
      total.fp64_pairs += sts[j].fp64_pairs; total.kernel_launches += sts[j].kernel_launches;
      total.n_classes = std::max(total.n_classes, sts[j].n_classes);
      total.h2d_ms = std::max(total.h2d_ms, sts[j].h2d_ms);
      total.kernel_ms = std::max(total.kernel_ms, sts[j].kernel_ms);
      total.d2h_ms = std::max(total.d2h_ms, sts[j].d2h_ms);
    }
  }
  std::lock_guard<std::mutex> lk(g_mu);
  g_last_stats = total;
  return GKLB_OK;
}

int gklb_pairhmm_done(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_refs > 0) g_refs--;
  if (g_refs == 0) free_idle_engines_locked();  // engines in use by a concurrent call are freed by a later done
  return GKLB_OK;
}

int gklb_pairhmm_devices_in_use(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_inited ? (int)g_devices.size() : 0;
}

int gklb_pairhmm_engines_alive(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return (int)g_pool.size();
}

int gklb_pairhmm_last_stats(gklb_pairhmm_stats* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!out) return fail(GKLB_ERR_INVALID, "out is null");
  *out = g_last_stats;
  return GKLB_OK;
}

int gklb_pairhmm_acquire_engine(int device, gklb_engine** out) {
  if (!out) return fail(GKLB_ERR_INVALID, "out is null");
  return acquire(device, out);
}

int gklb_pairhmm_release_engine(gklb_engine* e) {
  if (!e) return fail(GKLB_ERR_INVALID, "engine is null");
  release(e);
  return GKLB_OK;
}

}  // extern "C"
