/*
 * gklb_pdhmm.h -- C-ABI of the B200-native PDHMM engine (config 5; exported by libgkl_pdhmm.so, which is the
 * same binary as libgkl_pairhmm.so under the second name GKL's loader accepts, NativeLibraryLoader.java:45).
 *
 *   gklb_pdhmm_init            <- Java_com_intel_gkl_pdhmm_IntelPDHMM_initNative               pdhmm/IntelPDHMM.cc:43-60
 *   gklb_pdhmm_compute         <- Java_com_intel_gkl_pdhmm_IntelPDHMM_computePDHMMNative       pdhmm/IntelPDHMM.cc:144-244
 *                                 (flat arrays in, one log10 likelihood per pair out; replaces
 *                                 computePDHMM / avx512_impl / scalar_impl, pdhmm-implementation.h:332-396)
 *   gklb_pdhmm_compute_cross   <- Java_com_intel_gkl_pdhmm_IntelPDHMM_computeLikelihoodsNative pdhmm/IntelPDHMM.cc:62-137
 *                                 (reads x haplotypes, out[r * n_haps + h]; the reference expands the cross product
 *                                 into flat batches on the host, pdhmm/JavaData.h:177-242 -- here it is index
 *                                 arithmetic in the kernel)
 *   gklb_pdhmm_done            <- Java_com_intel_gkl_pdhmm_IntelPDHMM_doneNative               pdhmm/IntelPDHMM.cc:246-249
 *
 * Status codes are gklb_status (gklb_pairhmm.h); they map onto the reference's PDHMM codes
 * (pdhmm-common.h:38-42): OOM <-> MEMORY_ALLOCATION_FAILED, INVALID <-> INPUT_DATA_ERROR, CUDA <-> FAILURE.
 * No CPU path: without an sm_100 device every call fails with GKLB_ERR_NO_DEVICE.
 */
#ifndef GKLB_PDHMM_H
#define GKLB_PDHMM_H

#include "gklb_pairhmm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Flat batch of IntelPDHMM.computePDHMM (IntelPDHMM.java:163-204): pair k reads hap_bases / hap_pdbases at
 * [k * max_hap, k * max_hap + hap_lengths[k]) and the five read arrays at [k * max_read, ... + read_lengths[k]).
 * For gklb_pdhmm_compute_cross the same struct describes the operands once: n is unused, haplotype h sits at
 * h * max_hap (hap_lengths[n_haps]) and read r at r * max_read (read_lengths[n_reads]).  Host pointers. */
typedef struct gklb_pdhmm_batch {
  int64_t n;
  int32_t max_hap;
  int32_t max_read;
  const int8_t* hap_bases;
  const int8_t* hap_pdbases;   /* PD flag bytes: SNP=1 DEL_START=2 DEL_END=4 A=8 C=16 G=32 T=64 (pdhmm/MathUtils.h:66-75) */
  const int8_t* read_bases;
  const int8_t* read_qual;
  const int8_t* read_ins_qual;
  const int8_t* read_del_qual;
  const int8_t* gcp;
  const int64_t* hap_lengths;
  const int64_t* read_lengths;
} gklb_pdhmm_batch;

typedef struct gklb_pdhmm_stats {
  int64_t pairs;
  int64_t cells;        /* sum over pairs of read length x haplotype length */
  float h2d_ms, kernel_ms, d2h_ms;
  int32_t kernel_launches;
} gklb_pdhmm_stats;

/* The four arguments of PDHMMNativeArguments are accepted for signature compatibility; threads, AVX level and
 * memory cap have no meaning on the device.  Device: env GKLB_DEVICE.  Row-start state semantics: env
 * GKLB_PDHMM_ROW_STATE = "carry" (default: the reference's scalar path and GATK's Java) or "reset" (its AVX paths). */
GKLB_API int gklb_pdhmm_init(int openmp_setting, int max_threads, int avx_level, int max_memory_mb);
GKLB_API int gklb_pdhmm_compute(const gklb_pdhmm_batch* batch, double* likelihoods);
GKLB_API int gklb_pdhmm_compute_cross(const gklb_pdhmm_batch* operands, int32_t n_reads, int32_t n_haps,
                                      double* likelihoods);
GKLB_API int gklb_pdhmm_done(void);
GKLB_API int gklb_pdhmm_last_stats(gklb_pdhmm_stats* out);
/* Time `iters` kernel launches over the operands of the last compute call (still resident in HBM). */
GKLB_API int gklb_pdhmm_time_runs(int iters, float* ms_per_run);
/* Name of the kernel that carried the last finished compute call ("k_pdhmm3<7,8>", "k_pdhmm2<4,12>", ...); "" before. */
GKLB_API const char* gklb_pdhmm_kernel_name(void);
/* which: 0 qualToErrorProb double[255], 1 matchToMatchProb double[32640] (pdhmm-common.h:149-195) */
GKLB_API const void* gklb_pdhmm_table(int which, int* n);

#ifdef __cplusplus
}
#endif
#endif /* GKLB_PDHMM_H */
