// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Driver that links GKL's own, unmodified Smith-Waterman translation units (smithwaterman/avx2_impl.cc,
// smithwaterman/avx512_impl.cc, smithwaterman/smithwaterman_common.cc, compiled where they lie under
// /root/reference by oracle/Makefile) and runs them over a batch of pairs.  GKL's JNI wrapper
// (smithwaterman/IntelSmithWaterman.cc) cannot be built here (no jni.h); what this file restates of it:
//
//   * initNative:   AVX-512 vs AVX2 dispatch                              IntelSmithWaterman.cc:47-66
//   * alignNative:  one call of g_runSWOnePairBT per pair, cigar buffer of 2 * max(len) zeroed bytes
//                   (IntelSmithWaterman.java:135), result = (cigar, offset)   IntelSmithWaterman.cc:72-121
#ifdef linux
#include <omp.h>
#endif
#include <stdint.h>
#include <string.h>
#include <time.h>

#include <avx.h>           // reference: common/avx.h (is_avx512_supported)
#include "avx2_impl.h"     // reference: runSWOnePairBT_fp_avx2
#include "avx512_impl.h"   // reference: runSWOnePairBT_fp_avx512

static double sw_now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

extern "C" {

// engine: 0 = GKL's own dispatch, 1 = force AVX2, 2 = force AVX-512.  cigars: n rows of `pitch` bytes, zero filled
// here; row k gets pair k's CIGAR string (pitch must be >= 2 * max(len1, len2) of every pair, the Java buffer size).
// Returns 0, or the reference's SW_MEMORY_ALLOCATION_FAILED.
int gklref_sw(int n, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2, const int64_t* off2, int match,
              int mismatch, int open, int extend, int strategy, int engine, int n_threads, char* cigars, int pitch,
              int32_t* cigar_len, int32_t* offsets, int* used_avx512, double* seconds) {
  const bool avx512 = (engine == 2) || (engine == 0 && is_avx512_supported());
  int32_t (*fn)(int32_t, int32_t, int32_t, int32_t, uint8_t*, uint8_t*, int16_t, int16_t, int8_t, char*, int32_t,
                uint32_t*, int32_t*) = avx512 ? runSWOnePairBT_fp_avx512 : runSWOnePairBT_fp_avx2;
  if (used_avx512) *used_avx512 = avx512 ? 1 : 0;
  memset(cigars, 0, (size_t)n * (size_t)pitch);
  int failed = 0;
  const int threads = n_threads < 1 ? 1 : n_threads;
  const double t0 = sw_now_s();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
  for (int k = 0; k < n; k++) {
    const int len1 = (int)(off1[k + 1] - off1[k]), len2 = (int)(off2[k + 1] - off2[k]);
    const int cap = 2 * (len1 > len2 ? len1 : len2);  // IntelSmithWaterman.java:135
    uint32_t count = 0;
    int32_t offset = 0;
    const int32_t rc = fn(match, mismatch, open, extend, const_cast<uint8_t*>(seq1 + off1[k]),
                          const_cast<uint8_t*>(seq2 + off2[k]), (int16_t)len1, (int16_t)len2, (int8_t)strategy,
                          cigars + (size_t)k * pitch, cap < pitch ? cap : pitch, &count, &offset);
    if (rc != 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      failed = rc;
    }
    cigar_len[k] = (int32_t)count;
    offsets[k] = offset;
  }
  if (seconds) *seconds = sw_now_s() - t0;
  return failed;
}

}  // extern "C"
