// engine_internal.h -- types shared by engine.cu (one engine: planning, staging, launches) and
// engine_global.cu (the process-global surface the JNI layer calls: device pool, sharding, multi-region calls).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/gklb_pairhmm.h"
#include "pairhmm_device.cuh"
#include "pairhmm_h2.cuh"
#include "pairhmm_r2.cuh"
#include "pairhmm_kernels.h"

namespace gklb {

int fail(int code, const char* fmt, ...);
const std::string& last_error_string();
void set_last_error(const std::string& s);

#define GKLB_CU(call)                                                                                          \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      return ::gklb::fail(e_ == cudaErrorMemoryAllocation ? GKLB_ERR_OOM : GKLB_ERR_CUDA, "%s failed: %s", #call, \
                          cudaGetErrorString(e_));                                                             \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes);
  void release();
};

struct HostBuf {  // pinned
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes);
  void release();
};

// One region of a staged job: one read x haplotype batch (one active region of HaplotypeCaller = one
// computeLikelihoods call).  A job is one or several regions that share the launches.
struct Region {
  int n_reads = 0, n_haps = 0;
  int read_base = 0;        // index of the region's first read in the job's global read numbering
  int64_t arena_base = 0;   // byte offset of its reads in the job's arenas
  int hap_base = 0;         // index of its first haplotype in the job's global haplotype offsets
  int64_t hap_arena_base = 0;
  int64_t out_base = 0;     // offset of its likelihoods (doubles) in the job's result buffer
};

// One length class of one region: its reads, packed as records of `rows` rows, and the kernels that serve it.
struct ClassInst {
  int region = 0;
  int G = 0, K = 0, n_pass = 1, rows = 0, stride = 0;
  bool multi = false;
  const KernelEntry* kf = nullptr;  // forward sweep: H2 (fp32, single pass), F2 multi-pass, or fp64 tasks in use_double mode
  const KernelEntry* kd = nullptr;  // fp64: rerun-list kernel (and task kernel in use_double mode)
  int cfg_f = -1;                   // configuration index in the H2 multi-class kernel
  int cfg_d = -1;                   // configuration index in the fp64 multi-class kernels
  std::vector<int32_t> rid, len;    // record order; rid is the read's index in the job
  int n_rec = 0;
  size_t meta_rid = 0, meta_len = 0;  // offsets into the meta block
  size_t rec_off = 0;                 // into d_records
  size_t carry_off = 0, carry_stride = 0;
};

// One (class, tile) combination: the unit the kernels work on.  Its rerun lists and counters.
//   counters: +0 rerun items (H2 sweep -> range-extended rerun)   +1 fp64 items   +2 pairs flagged by the fp32 sweep
//             +3 task counter of a single-class sweep   +4 ... of the rerun   +5 ... of the fp64 list kernel
struct EntryInst {
  int cls = 0, tile = 0;
  size_t r2_off = 0;   // into d_r2 (uint2 units): capacity n_rec * n_pairs
  size_t fb_off = 0;   // into d_fb (uint2 units): capacity n_rec * n (haplotypes of the tile)
  int counter0 = 0;
};
constexpr int kEntryCounters = 6;

// A tile: as many haplotypes of one region as fit in shared memory beside the record slots, as two images.
struct Tile {
  int region = 0, index_in_region = 0;
  int hap0 = 0, n = 0, max_len = 0;  // haplotypes [hap0, hap0 + n) of the region
  size_t meta_off = 0;               // per-haplotype image (fp64 and multi-pass kernels)
  uint32_t bytes = 0, img_off = 0;   // ... its size and its offset inside the group's resident image
  // the same haplotypes as a pair image (pairhmm_h2.cuh): sorted by length, two per byte column
  int n_pairs = 0;
  size_t pmeta_off = 0;
  uint32_t pbytes = 0, pimg_off = 0;
  std::vector<int> order;  // haplotype indices (in the region) by decreasing length; pair q = order[2q], order[2q+1]
};

// A launch group: consecutive tiles (of one or several regions) whose images are resident together, so that one
// multi-class launch serves all of them from a single task queue.
struct Group {
  int tile0 = 0, n_tiles = 0;
  size_t meta_off = 0, pmeta_off = 0;  // start of the group's contiguous per-haplotype / pair images in the meta block
  uint32_t bytes = 0, pbytes = 0;
  int entry0 = 0, n_entries = 0;       // (class, tile) entries [entry0, entry0 + n_entries) of e->entries
  size_t h2_cls_off = 0, h2_cfg_off = 0, h2_end_off = 0;      // device-resident arrays of the multi-class launches
  size_t r2_cls_off = 0, r2_cfg_off = 0;                       //   range-extended rerun
  size_t dl_cls_off = 0, dl_cfg_off = 0;                       //   fp64 rerun list
  size_t dt_cls_off = 0, dt_cfg_off = 0, dt_end_off = 0;      //   fp64 tasks (use_double)
};

// One kernel launch of the staged batch's plan (built once per stage, replayed by every run).
struct Launch {
  const void* fn = nullptr;
  int grid = 0, threads = 0;
  size_t smem = 0;
  std::vector<uint8_t> params;   // the kernel's single by-value parameter struct
  uint32_t extra = 0;            // second parameter of the multi-class fp64 kernels (slot bytes)
  bool sweep = false;            // a forward-sweep launch: bracketed by events, counted in stats.sweep_ms
};

}  // namespace gklb

struct gklb_engine {
  int device = 0;
  bool use_double = false;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  bool r2_on = false;  // staged job: GKLB_R2=1 and every H2 class has an R2 kernel (G <= 16)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> kev;  // pairs of events around each forward-sweep launch of the last run
  int kev_used = 0;
  std::mutex mu;
  gklb::DevBuf d_tables;
  const float *d_ph2pr_f = nullptr, *d_mm_f = nullptr;
  const double *d_ph2pr_d = nullptr, *d_mm_d = nullptr;
  // staged batch
  bool staged = false;
  std::vector<gklb::Region> regions;
  std::vector<gklb::ClassInst> classes;
  std::vector<gklb::Tile> tiles;
  std::vector<gklb::Group> groups;
  std::vector<gklb::EntryInst> entries;
  std::vector<gklb::Launch> plan;
  gklb::DevBuf d_read_off, d_hap_off, d_arenas, d_meta, d_records, d_out, d_fb, d_r2, d_counters, d_carry;
  gklb::HostBuf h_meta, h_counters, h_out;
  gklb::DevBuf d_xhap, d_xf32, d_xidx, d_xval, d_xcnt;  // NCCL sharding path (engine_nccl.cu): panel, fp32 slab, overrides
  gklb::HostBuf h_xf32[2], h_xidx, h_xval;
  std::vector<double*> pending_out;  // destinations of the job submitted with gklb_engine_submit, until gklb_engine_wait
  size_t arena_pitch = 0;
  const int64_t* p_read_off = nullptr;  // where the staged offsets / arenas live on the device
  const int64_t* p_hap_off = nullptr;
  const uint8_t* p_arenas = nullptr;
  int n_counters = 0;
  int mega_counter0 = 0;  // first of the per-group unified queue counters (two per group)
  gklb_pairhmm_stats stats{};
  char sweep_kernel[96] = {0};  // name of the (last) forward-sweep kernel of the plan
  bool defer_panel = false;     // stage with empty panel images: the bases arrive in device memory (fill_panels_from_device)
  // forced kernel (measurement): policy,G,K,warps,var
  bool forced = false;
  int f_policy = 0, f_G = 0, f_K = 0, f_warps = 0, f_var = 0;
};

namespace gklb {

// engine.cu
int create_engine(gklb_engine** out, int device, int use_double);
void destroy_engine(gklb_engine* e);
int validate_batch(const gklb_pairhmm_batch* b);
// A job is k regions (k >= 1); outs[r] receives region r's likelihoods (double[n_reads * n_haps], read-major).
int do_stage(gklb_engine* e, const gklb_pairhmm_batch* batches, int k, bool hap_on_device);
int do_run(gklb_engine* e);
int do_fetch(gklb_engine* e, double* const* outs);
int do_compute(gklb_engine* e, const gklb_pairhmm_batch* batches, int k, double* const* outs);
int do_submit(gklb_engine* e, const gklb_pairhmm_batch* batches, int k, double* const* outs);
int do_wait(gklb_engine* e);
// stats.fallback_pairs / fp64_pairs from the counters last copied to e->h_counters
void read_fallback_count(gklb_engine* e);
bool use_r2(const gklb_engine* e);  // this job's flagged pairs go through the range-extended fp32 rerun (GKLB_R2=1)
// Write the haplotype bases of the staged (single-region) job's panel images from a device buffer laid out like the
// batch's hap_bases arena (kernels on the engine's stream).
int fill_panels_from_device(gklb_engine* e, const uint8_t* hap_bases_dev);

}  // namespace gklb
