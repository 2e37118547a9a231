// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Driver that links GKL's own, unmodified PairHMM translation units
// (pairhmm/avx_impl.cc, pairhmm/avx512_impl.cc, pairhmm/pairhmm_common.cc, compiled
// where they lie under /root/reference by oracle/Makefile) and exposes them through a
// flat C entry point.  GKL's JNI wrapper (pairhmm/IntelPairHmm.cc) cannot be built here
// (no jni.h / JVM), so this file restates exactly the part of it that is arithmetic:
//
//   * library-load state:   Context<float> g_ctxf; Context<double> g_ctxd;   IntelPairHmm.cc:44-45
//   * initNative:           FTZ on, AVX-512 vs AVX dispatch, ConvertChar::init  IntelPairHmm.cc:93-116
//   * computeLikelihoods:   the pair loop, fp32 -> fp64 fallback and log10     IntelPairHmm.cc:150-169
//   * JavaData::getData:    testcase index = r * numHaplotypes + h             JavaData.h:84-105
//
// Everything numeric comes from the reference headers included below.
#ifdef linux
#include <omp.h>
#endif
#include <math.h>
#include <stdint.h>
#include <time.h>
#include <vector>

#include <avx.h>              // reference: common/avx.h  (is_avx512_supported)
#include "pairhmm_common.h"   // reference: testcase, ConvertChar, MIN_ACCEPTED
#include "avx_impl.h"         // reference: compute_fp_avxs / compute_fp_avxd
#include "avx512_impl.h"      // reference: compute_fp_avx512s / compute_fp_avx512d
#include "Context.h"          // reference: Context<float>, Context<double>

static Context<float> g_ctxf;
static Context<double> g_ctxd;

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

extern "C" {

// engine: 0 = GKL's own dispatch (AVX-512 if supported, else AVX), 1 = force AVX, 2 = force AVX-512
// Returns 0 on success.  *used_avx512 and *seconds (pair loop only) are optional outputs.
int gklref_pairhmm(int n_reads, int n_haps, const int64_t* read_off, const uint8_t* read_bases,
                   const uint8_t* read_quals, const uint8_t* ins_gop, const uint8_t* del_gop,
                   const uint8_t* gcp, const int64_t* hap_off, const uint8_t* hap_bases,
                   int use_double, int n_threads, int engine, double* out, int* used_avx512,
                   double* seconds) {
  _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
  ConvertChar::init();

  float (*fp_float)(testcase*);
  double (*fp_double)(testcase*);
  bool avx512 = (engine == 2) || (engine == 0 && is_avx512_supported());
  if (avx512) {
    fp_float = compute_fp_avx512s;
    fp_double = compute_fp_avx512d;
  } else {
    fp_float = compute_fp_avxs;
    fp_double = compute_fp_avxd;
  }
  if (used_avx512) *used_avx512 = avx512 ? 1 : 0;

  std::vector<testcase> testcases;
  testcases.reserve((size_t)n_reads * (size_t)n_haps);
  for (int r = 0; r < n_reads; r++) {
    for (int h = 0; h < n_haps; h++) {
      testcase tc;
      tc.hap = (const char*)hap_bases + hap_off[h];
      tc.haplen = (int)(hap_off[h + 1] - hap_off[h]);
      tc.rs = (const char*)read_bases + read_off[r];
      tc.rslen = (int)(read_off[r + 1] - read_off[r]);
      tc.i = (const char*)ins_gop + read_off[r];
      tc.d = (const char*)del_gop + read_off[r];
      tc.c = (const char*)gcp + read_off[r];
      tc.q = (const char*)read_quals + read_off[r];
      testcases.push_back(tc);
    }
  }

  // GKL clamps the request to omp_get_max_threads() (IntelPairHmm.cc:72-74).  The caller here passes the
  // number of host cores it wants used; the num_threads clause below honours it even when a launcher
  // (torchrun) exported OMP_NUM_THREADS=1 for this process.
  int max_threads = n_threads < 1 ? 1 : n_threads;
  const bool g_use_double = use_double != 0;
  const long n = (long)testcases.size();

  double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel num_threads(max_threads)
#endif
  {
    // FTZ is per-thread MXCSR state.  GKL sets it on the thread that calls initNative and
    // OpenMP workers created afterwards inherit it; set it explicitly on every worker so the
    // result does not depend on when the thread pool was first spun up.
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (long i = 0; i < n; i++) {
      double result_final = 0;
      float result_float = g_use_double ? 0.0f : fp_float(&testcases[i]);
      if (result_float < MIN_ACCEPTED) {
        double result_double = fp_double(&testcases[i]);
        result_final = log10(result_double) - g_ctxd.LOG10_INITIAL_CONSTANT;
      } else {
        result_final = (double)(log10f(result_float) - g_ctxf.LOG10_INITIAL_CONSTANT);
      }
      out[i] = result_final;
    }
  }
  double t1 = now_s();
  if (seconds) *seconds = t1 - t0;
  return 0;
}

// The same arithmetic for an explicit list of (read, haplotype) pairs: out[k] is the value
// computeLikelihoods would write at index pair_r[k] * n_haps + pair_h[k].  Used to check every fp64-rerun pair of
// a batch too large to recompute in full on the host (BASELINE configs[3]).
int gklref_pairhmm_pairs(long n_pairs, const int32_t* pair_r, const int32_t* pair_h, const int64_t* read_off,
                         const uint8_t* read_bases, const uint8_t* read_quals, const uint8_t* ins_gop,
                         const uint8_t* del_gop, const uint8_t* gcp, const int64_t* hap_off, const uint8_t* hap_bases,
                         int n_threads, double* out, double* seconds) {
  _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
  ConvertChar::init();
  const bool avx512 = is_avx512_supported();
  float (*fp_float)(testcase*) = avx512 ? compute_fp_avx512s : compute_fp_avxs;
  double (*fp_double)(testcase*) = avx512 ? compute_fp_avx512d : compute_fp_avxd;
  const int max_threads = n_threads < 1 ? 1 : n_threads;
  double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel num_threads(max_threads)
#endif
  {
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
    for (long k = 0; k < n_pairs; k++) {
      const int r = pair_r[k], h = pair_h[k];
      testcase tc;
      tc.hap = (const char*)hap_bases + hap_off[h];
      tc.haplen = (int)(hap_off[h + 1] - hap_off[h]);
      tc.rs = (const char*)read_bases + read_off[r];
      tc.rslen = (int)(read_off[r + 1] - read_off[r]);
      tc.i = (const char*)ins_gop + read_off[r];
      tc.d = (const char*)del_gop + read_off[r];
      tc.c = (const char*)gcp + read_off[r];
      tc.q = (const char*)read_quals + read_off[r];
      float result_float = fp_float(&tc);
      if (result_float < MIN_ACCEPTED)
        out[k] = log10(fp_double(&tc)) - g_ctxd.LOG10_INITIAL_CONSTANT;
      else
        out[k] = (double)(log10f(result_float) - g_ctxf.LOG10_INITIAL_CONSTANT);
    }
  }
  if (seconds) *seconds = now_s() - t0;
  return 0;
}

int gklref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int gklref_avx512_supported(void) { return is_avx512_supported() ? 1 : 0; }

}  // extern "C"
