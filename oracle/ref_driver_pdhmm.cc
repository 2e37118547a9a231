// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Driver that links GKL's own, unmodified PDHMM translation units (pdhmm/MathUtils.cc,
// pdhmm/pdhmm-serial.cc, pdhmm/avx2_impl.cc, pdhmm/avx512_impl.cc, compiled where they lie under
// /root/reference by oracle/Makefile) and exposes the flat entry point that
// Java_com_intel_gkl_pdhmm_IntelPDHMM_computePDHMMNative reaches (pdhmm/IntelPDHMM.cc:144-244 ->
// pdhmm-implementation.h:365-396).  The JNI wrapper itself needs jni.h and cannot be built here.
#include <stdint.h>
#include <time.h>

#include <avx.h>                    // reference: common/avx.h
#include "pdhmm-implementation.h"   // reference: initializeNative, allocateDPTable, computePDHMM

static double now_s() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

extern "C" {

// level: 0 FASTEST_AVAILABLE, 1 SCALAR, 2 AVX2, 3 AVX512 (the ordinals of AVXLevel,
// pdhmm-implementation.h:45-51).  Flat layout: pair k uses hap[k * max_hap ...], read[k * max_read ...].
// Returns the reference's status code (pdhmm-common.h:38-42), or -1 if initialisation failed.
int gklref_pdhmm(const int8_t* hap_bases, const int8_t* hap_pdbases, const int8_t* read_bases, const int8_t* read_qual,
                 const int8_t* read_ins_qual, const int8_t* read_del_qual, const int8_t* gcp, double* result, int64_t n,
                 const int64_t* hap_lengths, const int64_t* read_lengths, int max_read, int max_hap, int level,
                 int threads, double* seconds) {
  try {
    const OpenMPSetting omp = threads > 1 ? OpenMPSetting::ENABLE : OpenMPSetting::DISABLE;
    if (!initializeNative(omp, threads, static_cast<AVXLevel>(level), 1 << 20)) return -1;
    if (allocateDPTable(max_hap, max_read) != PDHMM_SUCCESS) return PDHMM_MEMORY_ALLOCATION_FAILED;
  } catch (JavaException& e) {
    return -1;
  }
  // GKL's PDHMM never enables flush-to-zero, but its PairHMM init leaves it on for the thread that called it
  // (IntelPairHmm.cc:93-96) and the OpenMP pool threads keep whatever a previous PairHMM run set.  The PDHMM's
  // lowest likelihoods are fp64 denormals, so start from the JVM's default (FTZ off) on every pool thread.
  _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_OFF);
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 1 ? threads : 1)
  { _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_OFF); }
#endif
  const double t0 = now_s();
  const int rc = computePDHMM(hap_bases, hap_pdbases, read_bases, read_qual, read_ins_qual, read_del_qual, gcp, result, n,
                              hap_lengths, read_lengths, max_read, max_hap);
  if (seconds) *seconds = now_s() - t0;
  return rc;
}

}  // extern "C"
