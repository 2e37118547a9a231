// sw_engine.cu -- host side and C-ABI (include/gklb_sw.h) of the Smith-Waterman aligner.
// Replaces smithwaterman/IntelSmithWaterman.cc + runSWOnePairBT (PairWiseSW.h:454-501) below the JNI boundary.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <numeric>
#include <vector>

#include "../../include/gklb_sw.h"
#include "sw_device.cuh"

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

constexpr int kWarps = 16;         // per CTA: 126 registers per thread, 16 warps per SM
constexpr int kCtasPerSm = 1;

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, n + n / 8 + 256);
    if (e == cudaSuccess) cap = n + n / 8 + 256;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct SwEngine {
  int device = 0, num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  Buf seq1, off1, seq2, off2, order, bt, lines, runs_scratch, runs, heads, misc;
  SwParams last{};
  bool have_last = false;
  int last_grid = 0;
  gklb_sw_stats stats{};
};

std::mutex g_mu;
SwEngine* g_sw = nullptr;

void destroy(SwEngine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (Buf* b : {&e->seq1, &e->off1, &e->seq2, &e->off2, &e->order, &e->bt, &e->lines, &e->runs_scratch, &e->runs,
                 &e->heads, &e->misc})
    b->release();
  for (auto& ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int launch(SwEngine* e) {
  CU(cudaMemsetAsync(e->misc.p, 0, 8, e->stream));  // queue, cursor
  void* args[] = {&e->last};
  CU(cudaLaunchKernel(reinterpret_cast<const void*>(&k_smith_waterman<kWarps>), dim3(e->last_grid), dim3(kWarps * 32),
                      args, 0, e->stream));
  e->stats.kernel_launches++;
  return GKLB_OK;
}

// smithwaterman_common.cc:26-58 without the sign (lengths are positive)
int itoa_len(int v) {
  int d = 0;
  while (v > 0) { v /= 10; d++; }
  return d;
}

// The tail of getCIGAR (PairWiseSW.h:411-436): elements are emitted last to first; one that does not fit into what is
// left of the buffer, or has length 0, is skipped.  Returns the number of bytes written; a NUL follows when it fits.
int write_cigar(const uint32_t* runs, int n_runs, char* cigar, int cap) {
  int cur = 0;
  for (int k = n_runs - 1; k >= 0; k--) {
    const int op = (int)(runs[k] & 15u), len = (int)(runs[k] >> 4);
    const char st = op == 0 ? 'M' : op == 1 ? 'I' : op == 2 ? 'D' : op == 9 ? 'S' : 'R';
    const int digits = itoa_len(len), expected = digits + 1;
    if (expected > 1 && cur + expected <= cap) {
      int v = len;
      for (int i = digits - 1; i >= 0; i--) { cigar[cur + i] = (char)('0' + v % 10); v /= 10; }
      cur += digits;
      cigar[cur++] = st;
    }
  }
  if (cur < cap) cigar[cur] = 0;
  return cur;  // == strnlen of what was written (:436): every byte written is a digit or a letter
}

int align_chunk(SwEngine* e, const gklb_sw_batch* b, char* cigars, int32_t pitch, int32_t* cigar_len, int32_t* offsets) {
  if (!b || !cigars || !cigar_len || !offsets) return gklb_internal_fail(GKLB_ERR_INVALID, "null argument");
  if (b->n < 0) return gklb_internal_fail(GKLB_ERR_INVALID, "negative batch size");
  if (b->n == 0) return GKLB_OK;
  if (!b->seq1 || !b->seq2 || !b->seq1_off || !b->seq2_off) return gklb_internal_fail(GKLB_ERR_INVALID, "null array in batch");
  if (b->strategy < GKLB_SW_SOFTCLIP || b->strategy > GKLB_SW_IGNORE)  // IntelSmithWaterman.java:144
    return gklb_internal_fail(GKLB_ERR_INVALID, "Strategy is invalid.");
  if (pitch <= 0) return gklb_internal_fail(GKLB_ERR_INVALID, "cigar buffer is empty");
  const int n = b->n;
  int max1 = 0, max2 = 0;
  long long cells = 0;
  size_t bt_words = 0, total_runs = 0;
  std::vector<long long> size(n);
  for (int k = 0; k < n; k++) {
    const long long l1 = b->seq1_off[k + 1] - b->seq1_off[k], l2 = b->seq2_off[k + 1] - b->seq2_off[k];
    if (l1 <= 0 || l2 <= 0) return gklb_internal_fail(GKLB_ERR_INVALID, "Cannot align empty sequences");  // .java:131
    if (l1 > GKLB_SW_MAX_SEQUENCE_LENGTH || l2 > GKLB_SW_MAX_SEQUENCE_LENGTH)
      return gklb_internal_fail(GKLB_ERR_INVALID, "Sequences exceed maximum length of %d bytes", GKLB_SW_MAX_SEQUENCE_LENGTH);
    max1 = std::max<int>(max1, (int)l1);
    max2 = std::max<int>(max2, (int)l2);
    size[k] = l1 * l2;
    cells += size[k];
    bt_words = std::max(bt_words, sw_bt_words((int)l1, (int)l2));
    total_runs += (size_t)(l1 + l2 + 2);
  }
  std::vector<int32_t> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return size[a] > size[c]; });

  CU(cudaSetDevice(e->device));
  cudaStream_t s = e->stream;
  const int line_w = (max2 + 1 + 31) & ~31, line_r = (max1 + 1 + 31) & ~31;
  const size_t lines_stride = 5 * (size_t)line_w + (size_t)line_r;
  const size_t runs_stride = (size_t)line_w + line_r + 4;
  const size_t per_warp = 4 * (bt_words + lines_stride + runs_stride);
  // scratch budget: half of what is free, at most 16 GB; asked of the driver only when the batch could come near it
  // (cudaMemGetInfo costs more than a small alignment)
  const long long max_warps = std::min<long long>((long long)e->num_sms * kCtasPerSm * kWarps, ((long long)n + 0));
  size_t budget = (size_t)1 << 30;
  if (per_warp * (size_t)std::max<long long>(max_warps, 1) > budget) {
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    budget = std::min<size_t>((size_t)16 << 30, (free_b + e->bt.cap + e->lines.cap + e->runs_scratch.cap) / 2);
  }
  long long warps = std::min<long long>((long long)e->num_sms * kCtasPerSm * kWarps, (long long)(budget / per_warp));
  if (warps < 1)
    return gklb_internal_fail(GKLB_ERR_OOM, "backtrack scratch of %zu bytes for one pair does not fit in device memory", per_warp);
  int grid = (int)std::max<long long>(1, std::min<long long>(warps / kWarps, ((long long)n + kWarps - 1) / kWarps));
  if (warps < kWarps) grid = 1;
  const size_t n_warps = (size_t)grid * kWarps;

  CU(cudaEventRecord(e->ev[0], s));
  const size_t b1 = (size_t)b->seq1_off[n], b2 = (size_t)b->seq2_off[n];
  CU(e->seq1.ensure(b1 - (size_t)b->seq1_off[0] + 16));
  CU(e->seq2.ensure(b2 - (size_t)b->seq2_off[0] + 16));
  CU(e->off1.ensure(sizeof(int64_t) * (n + 1)));
  CU(e->off2.ensure(sizeof(int64_t) * (n + 1)));
  CU(e->order.ensure(sizeof(int32_t) * n));
  CU(e->bt.ensure(4 * bt_words * n_warps));
  CU(e->lines.ensure(4 * lines_stride * n_warps));
  CU(e->runs_scratch.ensure(4 * runs_stride * n_warps));
  CU(e->runs.ensure(4 * total_runs));
  CU(e->heads.ensure(sizeof(int32_t) * 3 * (size_t)n));
  CU(e->misc.ensure(64));
  CU(cudaMemcpyAsync(e->seq1.p, b->seq1 + b->seq1_off[0], b1 - (size_t)b->seq1_off[0], cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(e->seq2.p, b->seq2 + b->seq2_off[0], b2 - (size_t)b->seq2_off[0], cudaMemcpyHostToDevice, s));
  std::vector<int64_t> o1(b->seq1_off, b->seq1_off + n + 1), o2(b->seq2_off, b->seq2_off + n + 1);
  for (auto& v : o1) v -= b->seq1_off[0];
  for (auto& v : o2) v -= b->seq2_off[0];
  CU(cudaMemcpyAsync(e->off1.p, o1.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(e->off2.p, o2.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, s));
  CU(cudaMemcpyAsync(e->order.p, order.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, s));

  SwParams& p = e->last;
  p.seq1 = static_cast<const uint8_t*>(e->seq1.p);
  p.off1 = static_cast<const int64_t*>(e->off1.p);
  p.seq2 = static_cast<const uint8_t*>(e->seq2.p);
  p.off2 = static_cast<const int64_t*>(e->off2.p);
  p.order = static_cast<const int32_t*>(e->order.p);
  p.n = n;
  p.match = b->match; p.mismatch = b->mismatch; p.open = b->open; p.extend = b->extend; p.strategy = b->strategy;
  p.bt = static_cast<uint32_t*>(e->bt.p);
  p.bt_stride = bt_words;
  p.lines = static_cast<int32_t*>(e->lines.p);
  p.lines_stride = lines_stride;
  p.line_w = line_w;
  p.line_r = line_r;
  p.runs_scratch = static_cast<uint32_t*>(e->runs_scratch.p);
  p.runs = static_cast<uint32_t*>(e->runs.p);
  p.queue = static_cast<unsigned int*>(e->misc.p);
  p.cursor = p.queue + 1;
  p.run_start = static_cast<int32_t*>(e->heads.p);
  p.run_count = p.run_start + n;
  p.offsets = p.run_count + n;
  e->last_grid = grid;
  e->have_last = true;
  e->stats = gklb_sw_stats{};
  e->stats.pairs = n;
  e->stats.cells = cells;
  e->stats.warps = (int32_t)n_warps;

  CU(cudaEventRecord(e->ev[1], s));
  int rc = launch(e);
  if (rc) return rc;
  CU(cudaEventRecord(e->ev[2], s));
  std::vector<int32_t> heads(3 * (size_t)n);
  unsigned int misc[2] = {0, 0};
  CU(cudaMemcpyAsync(heads.data(), e->heads.p, sizeof(int32_t) * heads.size(), cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(misc, e->misc.p, sizeof(misc), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  std::vector<uint32_t> runs(misc[1]);
  if (misc[1]) CU(cudaMemcpyAsync(runs.data(), e->runs.p, 4 * (size_t)misc[1], cudaMemcpyDeviceToHost, s));
  CU(cudaEventRecord(e->ev[3], s));
  CU(cudaStreamSynchronize(s));
  cudaEventElapsedTime(&e->stats.h2d_ms, e->ev[0], e->ev[1]);
  cudaEventElapsedTime(&e->stats.kernel_ms, e->ev[1], e->ev[2]);
  cudaEventElapsedTime(&e->stats.d2h_ms, e->ev[2], e->ev[3]);

  for (int k = 0; k < n; k++) {
    const long long l1 = b->seq1_off[k + 1] - b->seq1_off[k], l2 = b->seq2_off[k + 1] - b->seq2_off[k];
    const int cap = (int)std::min<long long>(pitch, 2 * std::max(l1, l2));  // IntelSmithWaterman.java:135
    const int start = heads[k], count = heads[(size_t)n + k];
    if (start < 0 || count < 0 || (size_t)start + (size_t)count > runs.size())
      return gklb_internal_fail(GKLB_ERR_CUDA, "corrupt run list for pair %d", k);
    cigar_len[k] = write_cigar(runs.data() + start, count, cigars + (size_t)k * pitch, cap);
    offsets[k] = heads[2 * (size_t)n + k];
  }
  return GKLB_OK;
}

// The run arena of one launch is indexed with 32 bits and sized for the worst case (len1 + len2 + 2 elements per
// pair), so very large batches are cut into consecutive chunks of pairs (GKLB_SW_MAX_RUN_ELEMENTS overrides the
// bound, for tests); outputs are per pair, so the chunks simply write their own rows.
int align_batch(SwEngine* e, const gklb_sw_batch* b, char* cigars, int32_t pitch, int32_t* cigar_len, int32_t* offsets) {
  if (!b || !cigars || !cigar_len || !offsets) return gklb_internal_fail(GKLB_ERR_INVALID, "null argument");
  if (b->n < 0) return gklb_internal_fail(GKLB_ERR_INVALID, "negative batch size");
  if (b->n == 0) return GKLB_OK;
  if (!b->seq1_off || !b->seq2_off) return gklb_internal_fail(GKLB_ERR_INVALID, "null array in batch");
  static const long long max_elements = [] {
    const char* v = getenv("GKLB_SW_MAX_RUN_ELEMENTS");
    return (v && atoll(v) > 0) ? atoll(v) : (1LL << 27);
  }();
  gklb_sw_stats total{};
  int k0 = 0;
  while (k0 < b->n) {
    long long elements = 0;
    int k1 = k0;
    while (k1 < b->n) {
      const long long add = (b->seq1_off[k1 + 1] - b->seq1_off[k1]) + (b->seq2_off[k1 + 1] - b->seq2_off[k1]) + 2;
      if (k1 > k0 && elements + add > max_elements) break;
      elements += add;
      k1++;
    }
    gklb_sw_batch part = *b;
    part.n = k1 - k0;
    part.seq1_off = b->seq1_off + k0;
    part.seq2_off = b->seq2_off + k0;
    const int rc = align_chunk(e, &part, cigars + (size_t)k0 * pitch, pitch, cigar_len + k0, offsets + k0);
    if (rc) return rc;
    total.pairs += e->stats.pairs;
    total.cells += e->stats.cells;
    total.h2d_ms += e->stats.h2d_ms;
    total.kernel_ms += e->stats.kernel_ms;
    total.d2h_ms += e->stats.d2h_ms;
    total.kernel_launches += e->stats.kernel_launches;
    total.warps = std::max(total.warps, e->stats.warps);
    k0 = k1;
  }
  e->stats = total;
  return GKLB_OK;
}

}  // namespace

extern "C" {

int gklb_sw_init(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_sw) return GKLB_OK;  // initNative only selects function pointers; calling it again is harmless
  int n = 0;
  cudaError_t ce = cudaGetDeviceCount(&n);
  if (ce != cudaSuccess || n <= 0) return gklb_internal_fail(GKLB_ERR_NO_DEVICE, "no CUDA device (%s)", cudaGetErrorString(ce));
  const char* dev = getenv("GKLB_DEVICE");
  const int device = dev ? atoi(dev) : 0;
  if (device < 0 || device >= n) return gklb_internal_fail(GKLB_ERR_NO_DEVICE, "device %d out of range", device);
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return gklb_internal_fail(GKLB_ERR_NO_DEVICE, "device %d is sm_%d%d; sm_100a code only", device, prop.major, prop.minor);
  CU(cudaSetDevice(device));
  SwEngine* e = new SwEngine;
  e->device = device;
  e->num_sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  for (auto& ev : e->ev) CU(cudaEventCreate(&ev));
  g_sw = e;
  return GKLB_OK;
}

int gklb_sw_align_batch(const gklb_sw_batch* batch, char* cigars, int32_t cigar_pitch, int32_t* cigar_len, int32_t* offsets) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_sw) return gklb_internal_fail(GKLB_ERR_STATE, "gklb_sw_init has not been called");
  return align_batch(g_sw, batch, cigars, cigar_pitch, cigar_len, offsets);
}

int gklb_sw_align(int32_t match, int32_t mismatch, int32_t open, int32_t extend, const uint8_t* seq1, const uint8_t* seq2,
                  int32_t len1, int32_t len2, int32_t strategy, char* cigar, int32_t cigar_len, uint32_t* cigar_count,
                  int32_t* offset) {
  if (!seq1 || !seq2 || !cigar || !cigar_count || !offset) return gklb_internal_fail(GKLB_ERR_INVALID, "Arrays aren't valid.");
  const int64_t o1[2] = {0, len1}, o2[2] = {0, len2};
  gklb_sw_batch b{1, seq1, o1, seq2, o2, match, mismatch, open, extend, strategy};
  // runSWOnePairBT writes through the caller's buffer without clearing it; the batch call zero-fills a row of
  // its own and the result is copied over
  const int cap = cigar_len > 0 ? cigar_len : 0;
  std::vector<char> row((size_t)std::max(cap, 1), 0);
  int32_t clen = 0, off = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_sw) return gklb_internal_fail(GKLB_ERR_STATE, "gklb_sw_init has not been called");
  if (cap <= 0) return gklb_internal_fail(GKLB_ERR_INVALID, "Strategy is invalid.");  // .java:144 (cigar.length <= 0)
  const int rc = align_batch(g_sw, &b, row.data(), cap, &clen, &off);
  if (rc) return rc;
  memcpy(cigar, row.data(), (size_t)clen);
  *cigar_count = (uint32_t)clen;
  *offset = off;
  return GKLB_OK;
}

int gklb_sw_done(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_sw) { destroy(g_sw); g_sw = nullptr; }
  return GKLB_OK;
}

int gklb_sw_last_stats(gklb_sw_stats* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_sw || !out) return gklb_internal_fail(GKLB_ERR_STATE, "no engine");
  *out = g_sw->stats;
  return GKLB_OK;
}

int gklb_sw_time_runs(int iters, float* ms_per_run) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_sw || !g_sw->have_last || iters <= 0 || !ms_per_run) return gklb_internal_fail(GKLB_ERR_STATE, "nothing to time");
  SwEngine* e = g_sw;
  CU(cudaSetDevice(e->device));
  CU(cudaEventRecord(e->ev[0], e->stream));
  for (int i = 0; i < iters; i++) {
    int rc = launch(e);
    if (rc) return rc;
  }
  CU(cudaEventRecord(e->ev[1], e->stream));
  CU(cudaEventSynchronize(e->ev[1]));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e->ev[0], e->ev[1]));
  *ms_per_run = ms / iters;
  return GKLB_OK;
}

}  // extern "C"
