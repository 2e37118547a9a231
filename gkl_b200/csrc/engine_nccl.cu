// engine_nccl.cu -- the NCCL form of the in-process multi-GPU sharding (GKLB_SHARD=nccl):
//
//   one process, ncclCommInitAll over the configured GPUs (kept for the life of the process), and inside ONE
//   gklb_pairhmm_compute call:
//     1. read shards go host->device straight to their owning GPU (no collective); the haplotype arena goes
//        host->device to GPU 0 only and travels to the others with one ncclBroadcast over NVLink; every GPU writes
//        the bases into its shared-memory panel images with a kernel (k_fill_panel / k_fill_pair_panel);
//     2. every GPU sweeps its shard;
//     3. results are narrowed on the device to what the reference's values really are -- an fp32 per pair
//        (IntelPairHmm.cc:164 widens a float) plus sparse (index, fp64) overrides for the pairs of the fp64 rerun
//        (:160-161) -- and gathered to GPU 0 with grouped ncclSend/ncclRecv;
//     4. GPU 0 copies the fp32 matrix and the overrides to the host, where they are widened into the caller's
//        double array.
//   The results are bit-identical to the direct path's (engine_global.cu).  The direct path, where every GPU uses
//   its own PCIe link in both directions, stays the default; bench/configs.py measures both
//   (profiles/r2_config4_nccl_vs_direct.json).
//
// NCCL is loaded with dlopen("libnccl.so.2") on first use: libgkl_pairhmm.so must stay loadable from a temp file
// without NCCL on the machine (NativeLibraryLoader.java:114-128), and without it this path reports an error.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>

#include "engine_internal.h"

namespace gklb {

namespace {

struct NcclApi {
  void* handle = nullptr;
  std::string error;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

NcclApi& api() {
  static NcclApi a = [] {
    NcclApi x;
    const char* names[] = {getenv("GKLB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      x.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (x.handle) break;
    }
    if (!x.handle) {
      x.error = "libnccl.so.2 could not be loaded (set GKLB_NCCL_LIB)";
      return x;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(x.handle, name);
      if (!p) { ok = false; x.error = std::string("missing NCCL symbol ") + name; }
      return p;
    };
    x.GetVersion = reinterpret_cast<decltype(x.GetVersion)>(sym("ncclGetVersion"));
    x.CommInitAll = reinterpret_cast<decltype(x.CommInitAll)>(sym("ncclCommInitAll"));
    x.CommDestroy = reinterpret_cast<decltype(x.CommDestroy)>(sym("ncclCommDestroy"));
    x.GetErrorString = reinterpret_cast<decltype(x.GetErrorString)>(sym("ncclGetErrorString"));
    x.Broadcast = reinterpret_cast<decltype(x.Broadcast)>(sym("ncclBroadcast"));
    x.Send = reinterpret_cast<decltype(x.Send)>(sym("ncclSend"));
    x.Recv = reinterpret_cast<decltype(x.Recv)>(sym("ncclRecv"));
    x.GroupStart = reinterpret_cast<decltype(x.GroupStart)>(sym("ncclGroupStart"));
    x.GroupEnd = reinterpret_cast<decltype(x.GroupEnd)>(sym("ncclGroupEnd"));
    if (!ok) { dlclose(x.handle); x.handle = nullptr; }
    return x;
  }();
  return a;
}

#define NC(call)                                                                                        \
  do {                                                                                                  \
    ncclResult_t r_ = (call);                                                                           \
    if (r_ != ncclSuccess) return fail(GKLB_ERR_CUDA, "%s failed: %s", #call, api().GetErrorString(r_)); \
  } while (0)
#define CU GKLB_CU

std::mutex g_mu;                  // one sharded call at a time uses the communicators
std::vector<int> g_comm_devices;  // the device list the cached communicators were built for
std::vector<ncclComm_t> g_comms;

int ensure_comms(const std::vector<int>& devices) {
  if (devices == g_comm_devices && !g_comms.empty()) return GKLB_OK;
  for (auto c : g_comms) api().CommDestroy(c);
  g_comms.assign(devices.size(), nullptr);
  g_comm_devices.clear();
  NC(api().CommInitAll(g_comms.data(), (int)devices.size(), devices.data()));
  g_comm_devices = devices;
  return GKLB_OK;
}

__global__ void k_narrow(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (float)in[i];
}

// The same for the rerun items of an H2 entry: (record, pair index in the tile | half mask << 30); the haplotype
// indices of a pair are in the header of the tile's pair image.
__global__ void k_collect_overrides_r2(const uint2* __restrict__ items, const unsigned int* __restrict__ count,
                                       const int32_t* __restrict__ rec_rid, const uint8_t* __restrict__ pair_image, int n_pairs,
                                       int n_haps, const double* __restrict__ out, uint32_t* __restrict__ idx,
                                       double* __restrict__ val, unsigned int* cursor, unsigned int cap) {
  const int32_t* idxA = reinterpret_cast<const int32_t*>(pair_image) + 3 * n_pairs;
  const int32_t* idxB = idxA + n_pairs;
  const unsigned int n = *count;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const uint2 it = items[k];
    const int q = (int)(it.y & 0x3fffffffu);
    const unsigned int mask = it.y >> 30;
    for (int x = 0; x < 2; x++) {
      if (!(mask & (1u << x))) continue;
      const int h = x == 0 ? idxA[q] : idxB[q];
      const uint32_t pair = (uint32_t)rec_rid[it.x] * (uint32_t)n_haps + (uint32_t)h;
      const unsigned int at = atomicAdd(cursor, 1u);
      if (at < cap) {  // the cursor keeps counting: the receiver sees an overflow as count > capacity
        idx[at] = pair;
        val[at] = out[pair];
      }
    }
  }
}

// The pairs the fp64 kernel produced from a (record, haplotype) list (multi-pass classes), as (pair index, value).
__global__ void k_collect_overrides(const uint2* __restrict__ items, const unsigned int* __restrict__ count,
                                    const int32_t* __restrict__ rec_rid, int n_haps, const double* __restrict__ out,
                                    uint32_t* __restrict__ idx, double* __restrict__ val, unsigned int* cursor,
                                    unsigned int cap) {
  const unsigned int n = *count;
  for (unsigned int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const uint2 it = items[k];
    const uint32_t pair = (uint32_t)rec_rid[it.x] * (uint32_t)n_haps + it.y;
    const unsigned int at = atomicAdd(cursor, 1u);
    if (at < cap) {
      idx[at] = pair;
      val[at] = out[pair];
    }
  }
}

// Enqueue the narrowing of one engine's (single-region) result: fp32 matrix + overrides of at most `cap` flagged pairs.
int enqueue_narrow(gklb_engine* e, float* f32, uint32_t* idx, double* val, unsigned int* cursor, unsigned int cap) {
  const size_t n = (size_t)e->stats.pairs;
  const int H = e->regions[0].n_haps;
  GKLB_CU(cudaMemsetAsync(cursor, 0, sizeof(unsigned int), e->stream));
  if (!n) return GKLB_OK;
  k_narrow<<<e->num_sms * 4, 256, 0, e->stream>>>(static_cast<const double*>(e->d_out.p), f32, n);
  const uint8_t* dm = static_cast<const uint8_t*>(e->d_meta.p);
  for (auto& en : e->entries) {  // every pair the plain fp32 sweep flagged carries a value computed in double
    const ClassInst& c = e->classes[en.cls];
    const Tile& t = e->tiles[en.tile];
    const unsigned int* cnt = static_cast<const unsigned int*>(e->d_counters.p) + en.counter0;
    if (c.kf->policy == POL_H2 && use_r2(e)) {  // the H2 sweep's flagged pairs are rerun items (record, pair | mask)
      k_collect_overrides_r2<<<e->num_sms, 256, 0, e->stream>>>(
          static_cast<const uint2*>(e->d_r2.p) + en.r2_off, cnt, reinterpret_cast<const int32_t*>(dm + c.meta_rid),
          dm + t.pmeta_off, t.n_pairs, H, static_cast<const double*>(e->d_out.p), idx, val, cursor, cap);
    } else if (e->d_fb.p) {
      k_collect_overrides<<<e->num_sms, 256, 0, e->stream>>>(
          static_cast<const uint2*>(e->d_fb.p) + en.fb_off, cnt + 1, reinterpret_cast<const int32_t*>(dm + c.meta_rid), H,
          static_cast<const double*>(e->d_out.p), idx, val, cursor, cap);
    }
  }
  GKLB_CU(cudaGetLastError());
  return GKLB_OK;
}

void widen(const float* src, double* dst, size_t n, int threads) {
  std::vector<std::thread> th;
  const size_t per = (n + threads - 1) / threads;
  for (int t = 0; t < threads; t++) {
    const size_t lo = std::min(n, per * t), hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] {
      for (size_t i = lo; i < hi; i++) dst[i] = (double)src[i];
    });
  }
  for (auto& t : th) t.join();
}

}  // namespace

bool nccl_available() { return api().handle != nullptr; }

int sharded_compute_nccl(const std::vector<gklb_engine*>& engines, const gklb_pairhmm_batch* batch, const std::vector<int>& cut,
                         double* likelihoods, gklb_pairhmm_stats* stats) {
  if (!api().handle) return fail(GKLB_ERR_STATE, "GKLB_SHARD=nccl: %s", api().error.c_str());
  const int n = (int)engines.size();
  const int H = batch->n_haps;
  if ((long long)(cut[1] - cut[0]) * H > 0xffffffffLL) return fail(GKLB_ERR_INVALID, "shard too large for 32-bit pair indices");
  std::lock_guard<std::mutex> lk(g_mu);
  std::vector<int> devices(n);
  for (int g = 0; g < n; g++) devices[g] = engines[g]->device;
  {
    std::vector<int> d = devices;
    std::sort(d.begin(), d.end());
    if (std::adjacent_find(d.begin(), d.end()) != d.end())
      return fail(GKLB_ERR_INVALID, "GKLB_SHARD=nccl needs distinct devices (one NCCL rank per GPU)");
  }
  int rc = ensure_comms(devices);
  if (rc) return rc;
  const size_t hap_bytes = (size_t)batch->hap_off[H];
  cudaEvent_t t0 = engines[0]->ev[0], t1 = engines[0]->ev[1], t2 = engines[0]->ev[2], t3 = engines[0]->ev[3];
  CU(cudaSetDevice(engines[0]->device));
  CU(cudaEventRecord(t0, engines[0]->stream));

  // 1. stage every shard on its GPU (its reads over its own PCIe link; panel images without bases)
  std::vector<std::vector<int64_t>> offs(n);
  std::vector<gklb_pairhmm_batch> sub(n);
  std::vector<int> rcs(n, GKLB_OK);
  std::vector<std::string> errs(n);
  {
    std::vector<std::thread> th;
    for (int g = 0; g < n; g++) {
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
      th.emplace_back([&, g] {
        gklb_engine* e = engines[g];
        std::lock_guard<std::mutex> elk(e->mu);
        e->defer_panel = true;
        rcs[g] = do_stage(e, &sub[g], 1, false);
        e->defer_panel = false;
        if (!rcs[g] && cudaSetDevice(e->device) == cudaSuccess && e->d_xhap.ensure(hap_bytes) != cudaSuccess)
          rcs[g] = fail(GKLB_ERR_OOM, "device allocation failed");
        if (rcs[g]) errs[g] = last_error_string();
      });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < n; g++)
      if (rcs[g]) { set_last_error(errs[g]); return rcs[g]; }
  }
  // the panel: host -> GPU 0, then one broadcast
  CU(cudaSetDevice(engines[0]->device));
  CU(cudaMemcpyAsync(engines[0]->d_xhap.p, batch->hap_bases, hap_bytes, cudaMemcpyHostToDevice, engines[0]->stream));
  NC(api().GroupStart());
  for (int g = 0; g < n; g++)
    NC(api().Broadcast(engines[g]->d_xhap.p, engines[g]->d_xhap.p, hap_bytes, ncclUint8, 0, g_comms[g], engines[g]->stream));
  NC(api().GroupEnd());
  CU(cudaSetDevice(engines[0]->device));
  CU(cudaEventRecord(t1, engines[0]->stream));

  // 2. sweep
  for (int g = 0; g < n; g++) {
    gklb_engine* e = engines[g];
    if ((rc = fill_panels_from_device(e, static_cast<const uint8_t*>(e->d_xhap.p)))) return rc;
    if ((rc = do_run(e))) return rc;
    CU(cudaMemcpyAsync(e->h_counters.p, e->d_counters.p, sizeof(unsigned int) * (size_t)e->n_counters,
                       cudaMemcpyDeviceToHost, e->stream));
  }
  // 3. narrow: fp32 slab + overrides (their number is known once the rerun lists are final)
  std::vector<size_t> n_pairs(n), n_ovr(n), pair_base(n);
  size_t total_pairs = 0, total_ovr = 0;
  for (int g = 0; g < n; g++) {
    gklb_engine* e = engines[g];
    CU(cudaSetDevice(e->device));
    CU(cudaStreamSynchronize(e->stream));
    read_fallback_count(e);
    const size_t fb = (size_t)e->stats.fallback_pairs;
    n_pairs[g] = (size_t)e->stats.pairs;
    n_ovr[g] = fb;
    pair_base[g] = total_pairs;
    total_pairs += n_pairs[g];
    total_ovr += fb;
  }
  CU(cudaSetDevice(engines[0]->device));
  CU(cudaEventRecord(t2, engines[0]->stream));
  gklb_engine* root = engines[0];
  for (int g = 0; g < n; g++) {
    gklb_engine* e = engines[g];
    CU(cudaSetDevice(e->device));
    // GPU 0 narrows straight into the gathered matrix and appends its overrides to the gathered lists
    CU(e->d_xf32.ensure(sizeof(float) * (g == 0 ? total_pairs : n_pairs[g])));
    CU(e->d_xidx.ensure(sizeof(uint32_t) * std::max<size_t>(1, g == 0 ? total_ovr : n_ovr[g])));
    CU(e->d_xval.ensure(sizeof(double) * std::max<size_t>(1, g == 0 ? total_ovr : n_ovr[g])));
    CU(e->d_xcnt.ensure(sizeof(unsigned int)));
    if ((rc = enqueue_narrow(e, static_cast<float*>(e->d_xf32.p), static_cast<uint32_t*>(e->d_xidx.p),
                             static_cast<double*>(e->d_xval.p), static_cast<unsigned int*>(e->d_xcnt.p), 0xffffffffu)))
      return rc;
  }
  // gather to GPU 0 over NVLink
  {
    std::vector<size_t> ovr_base(n);
    size_t at = n_ovr[0];
    for (int g = 1; g < n; g++) { ovr_base[g] = at; at += n_ovr[g]; }
    NC(api().GroupStart());
    for (int g = 1; g < n; g++) {
      gklb_engine* e = engines[g];
      if (n_pairs[g]) {
        NC(api().Send(e->d_xf32.p, n_pairs[g], ncclFloat32, 0, g_comms[g], e->stream));
        NC(api().Recv(static_cast<float*>(root->d_xf32.p) + pair_base[g], n_pairs[g], ncclFloat32, g, g_comms[0], root->stream));
      }
      if (n_ovr[g]) {
        NC(api().Send(e->d_xidx.p, n_ovr[g], ncclUint32, 0, g_comms[g], e->stream));
        NC(api().Send(e->d_xval.p, n_ovr[g], ncclFloat64, 0, g_comms[g], e->stream));
        NC(api().Recv(static_cast<uint32_t*>(root->d_xidx.p) + ovr_base[g], n_ovr[g], ncclUint32, g, g_comms[0], root->stream));
        NC(api().Recv(static_cast<double*>(root->d_xval.p) + ovr_base[g], n_ovr[g], ncclFloat64, g, g_comms[0], root->stream));
      }
    }
    NC(api().GroupEnd());
    // 4. one device->host path: the fp32 matrix in chunks through two pinned buffers, widened while the next chunk
    //    is in flight; then the overrides
    CU(cudaSetDevice(root->device));
    const size_t chunk = (size_t)16 << 20;  // floats per chunk (64 MB)
    for (auto& hbuf : root->h_xf32) CU(hbuf.ensure(sizeof(float) * std::min(chunk, total_pairs)));
    CU(root->h_xidx.ensure(sizeof(uint32_t) * std::max<size_t>(1, total_ovr)));
    CU(root->h_xval.ensure(sizeof(double) * std::max<size_t>(1, total_ovr)));
    const int wthreads = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t n_chunks = (total_pairs + chunk - 1) / chunk;
    std::vector<cudaEvent_t> done(2);
    for (auto& ev : done) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (size_t k = 0; k <= n_chunks; k++) {
      if (k < n_chunks) {
        const size_t lo = k * chunk, cnt = std::min(chunk, total_pairs - lo);
        CU(cudaMemcpyAsync(root->h_xf32[k & 1].p, static_cast<const float*>(root->d_xf32.p) + lo, sizeof(float) * cnt,
                           cudaMemcpyDeviceToHost, root->stream));
        CU(cudaEventRecord(done[k & 1], root->stream));
      }
      if (k > 0) {
        const size_t lo = (k - 1) * chunk, cnt = std::min(chunk, total_pairs - lo);
        CU(cudaEventSynchronize(done[(k - 1) & 1]));
        widen(static_cast<const float*>(root->h_xf32[(k - 1) & 1].p), likelihoods + lo, cnt, wthreads);
      }
    }
    for (auto& ev : done) cudaEventDestroy(ev);
    if (total_ovr) {
      CU(cudaMemcpyAsync(root->h_xidx.p, root->d_xidx.p, sizeof(uint32_t) * total_ovr, cudaMemcpyDeviceToHost, root->stream));
      CU(cudaMemcpyAsync(root->h_xval.p, root->d_xval.p, sizeof(double) * total_ovr, cudaMemcpyDeviceToHost, root->stream));
    }
    CU(cudaEventRecord(t3, root->stream));
    CU(cudaStreamSynchronize(root->stream));
    const uint32_t* hi = static_cast<const uint32_t*>(root->h_xidx.p);
    const double* hv = static_cast<const double*>(root->h_xval.p);
    size_t k = 0;
    for (int g = 0; g < n; g++)
      for (size_t j = 0; j < n_ovr[g]; j++, k++) likelihoods[pair_base[g] + hi[k]] = hv[k];
  }
  for (int g = 1; g < n; g++) {
    CU(cudaSetDevice(engines[g]->device));
    CU(cudaStreamSynchronize(engines[g]->stream));
  }
  *stats = gklb_pairhmm_stats{};
  for (int g = 0; g < n; g++) {
    const gklb_pairhmm_stats& st = engines[g]->stats;
    stats->pairs += st.pairs;
    stats->cells += st.cells;
    stats->fallback_pairs += st.fallback_pairs;
    stats->fp64_pairs += st.fp64_pairs;
    stats->kernel_launches += st.kernel_launches;
    stats->n_classes = std::max(stats->n_classes, st.n_classes);
  }
  CU(cudaSetDevice(root->device));
  cudaEventElapsedTime(&stats->h2d_ms, t0, t1);     // staging + broadcast, as seen by GPU 0's stream
  cudaEventElapsedTime(&stats->kernel_ms, t1, t2);  // GPU 0's kernels; the host waited for all GPUs before t2 was recorded
  cudaEventElapsedTime(&stats->d2h_ms, t2, t3);     // narrowing, gather, device->host, widening
  return GKLB_OK;
}

}  // namespace gklb

using namespace gklb;

extern "C" {

// The result of the last run as one packed device buffer (used by multi-process hosts that gather results over
// NVLink, bench.py --gpus N):  float likelihoods[pairs] | uint32 count | uint32 capacity | uint32 index[capacity] |
// (8-byte aligned) double value[capacity].  The reference's value for an unflagged pair IS an fp32 widened to double
// (IntelPairHmm.cc:164), so the fp32 matrix loses nothing; the flagged pairs (fp64 results) travel as overrides.
// count > capacity means the override list overflowed.  Asynchronous on the engine's stream.
int gklb_engine_narrow(gklb_engine* e, unsigned int capacity, void** packed_dev, size_t* packed_bytes) {
  if (!e || !packed_dev || !packed_bytes) return fail(GKLB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(e->mu);
  if (!e->staged || e->regions.size() != 1) return fail(GKLB_ERR_STATE, "a single-region job must be staged");
  GKLB_CU(cudaSetDevice(e->device));
  const size_t n = (size_t)e->stats.pairs;
  const size_t off_cnt = sizeof(float) * n, off_idx = off_cnt + 8, off_val = (off_idx + sizeof(uint32_t) * capacity + 7) / 8 * 8;
  const size_t bytes = off_val + sizeof(double) * capacity;
  GKLB_CU(e->d_xf32.ensure(bytes));
  uint8_t* base = static_cast<uint8_t*>(e->d_xf32.p);
  int rc = enqueue_narrow(e, reinterpret_cast<float*>(base), reinterpret_cast<uint32_t*>(base + off_idx),
                          reinterpret_cast<double*>(base + off_val), reinterpret_cast<unsigned int*>(base + off_cnt), capacity);
  if (rc) return rc;
  *packed_dev = base;
  *packed_bytes = bytes;
  return GKLB_OK;
}

}  // extern "C"
