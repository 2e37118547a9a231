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

// One length class of a staged batch: its reads, packed as records of `rows` rows, and the kernels that serve it.
struct ClassInst {
  int G = 0, K = 0, n_pass = 1, rows = 0, stride = 0;
  bool multi = false;
  const KernelEntry* kf = nullptr;  // forward sweep: H2 (fp32, single pass), F2 multi-pass, or fp64 tasks in use_double mode
  const KernelEntry* kd = nullptr;  // fp64: rerun-list kernel (and task kernel in use_double mode)
  int cfg_f = -1;                   // configuration index in the H2 multi-class kernel
  int cfg_d = -1;                   // configuration index in the fp64 multi-class kernels
  std::vector<int32_t> rid, len;    // record order
  int n_rec = 0;
  size_t meta_rid = 0, meta_len = 0;  // offsets into the meta block
  size_t rec_off = 0;                 // into d_records
  size_t fb_off = 0;                  // into d_fb (uint2 units)
  size_t carry_off = 0, carry_stride = 0;
  int counter0 = 0;  // index of this class's first counter (rerun count), then per tile: task counter, list counter
};

struct Tile {
  int hap0 = 0, n = 0, max_len = 0;
  size_t meta_off = 0;
  uint32_t bytes = 0;
  // the same haplotypes as a pair image (pairhmm_h2.cuh): sorted by length, two per byte column
  int n_pairs = 0;
  size_t pmeta_off = 0;
  uint32_t pbytes = 0;
  std::vector<int> order;  // haplotype indices by decreasing length; pair q = order[2q], order[2q+1]
  size_t cls_list_off = 0;   // meta offset of the SweepParams array of the multi-class rerun launch
  size_t cls_tasks_off = 0;  // ... and of the multi-class fp64 task launch (use_double)
};

// One kernel launch of the staged batch's plan (built once per stage, replayed by every run).
struct Launch {
  const void* fn = nullptr;
  int grid = 0, threads = 0;
  size_t smem = 0;
  std::vector<uint8_t> params;   // the kernel's single by-value parameter struct
  uint32_t extra = 0;            // second parameter of the multi-class fp64 kernels (slot bytes)
  bool has_extra = false;
  bool sweep = false;            // a forward-sweep launch: bracketed by events, counted in stats.sweep_ms
  const char* name = "";
};

}  // namespace gklb

struct gklb_engine {
  int device = 0;
  bool use_double = false;
  int num_sms = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> kev;  // pairs of events around each forward-sweep launch of the last run
  int kev_used = 0;
  std::mutex mu;
  gklb::DevBuf d_tables;
  const float *d_ph2pr_f = nullptr, *d_mm_f = nullptr;
  const double *d_ph2pr_d = nullptr, *d_mm_d = nullptr;
  // staged batch
  bool staged = false;
  int n_reads = 0, n_haps = 0;
  std::vector<gklb::ClassInst> classes;
  std::vector<gklb::Tile> tiles;
  std::vector<gklb::Launch> plan;
  gklb::DevBuf d_read_off, d_hap_off, d_arenas, d_meta, d_records, d_out, d_fb, d_counters, d_carry;
  gklb::HostBuf h_meta, h_counters, h_out;
  double* pending_out = nullptr;  // destination of the batch submitted with gklb_engine_submit, until gklb_engine_wait
  int64_t pending_pairs = 0;
  size_t arena_pitch = 0;
  const int64_t* p_read_off = nullptr;  // where the staged offsets / arenas live on the device
  const int64_t* p_hap_off = nullptr;
  const uint8_t* p_arenas = nullptr;
  int n_counters = 0;
  int mega_counter0 = 0;  // first of the per-tile unified queue counters
  gklb_pairhmm_stats stats{};
  char sweep_kernel[96] = {0};  // name of the (last) forward-sweep kernel of the plan
  // forced kernel (measurement): policy,G,K,warps,var
  bool forced = false;
  int f_policy = 0, f_G = 0, f_K = 0, f_warps = 0, f_var = 0;
};

namespace gklb {

// engine.cu
int create_engine(gklb_engine** out, int device, int use_double);
void destroy_engine(gklb_engine* e);
int validate_batch(const gklb_pairhmm_batch* b);
int do_stage(gklb_engine* e, const gklb_pairhmm_batch* b, bool hap_on_device);
int do_run(gklb_engine* e);
int do_fetch(gklb_engine* e, double* out);
int do_compute(gklb_engine* e, const gklb_pairhmm_batch* b, double* out);
int do_submit(gklb_engine* e, const gklb_pairhmm_batch* b, double* out);
int do_wait(gklb_engine* e);

}  // namespace gklb
