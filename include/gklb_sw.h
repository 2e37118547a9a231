/*
 * gklb_sw.h -- C-ABI of the B200-native Smith-Waterman aligner (exported by
 * libgkl_smithwaterman.so, the same binary as libgkl_pairhmm.so under the third name GKL's loader accepts,
 * NativeLibraryLoader.java:45).  Citations are relative to /root/reference/src/main/native.
 *
 *   gklb_sw_init         <- Java_com_intel_gkl_smithwaterman_IntelSmithWaterman_initNative   smithwaterman/IntelSmithWaterman.cc:47-66
 *   gklb_sw_align        <- ..._alignNative -> runSWOnePairBT_{avx2,avx512}                   IntelSmithWaterman.cc:72-121, PairWiseSW.h:454-501
 *                           (same arguments and results as runSWOnePairBT: CIGAR string into the caller's buffer, its
 *                           length, the alignment offset)
 *   gklb_sw_align_batch  -- the batched form a GPU needs (the reference aligns one pair per JNI call): n pairs, one
 *                           set of scoring parameters and one overhang strategy, per-pair CIGAR rows and offsets
 *   gklb_sw_done         <- ..._doneNative                                                    IntelSmithWaterman.cc:128-130
 *
 * Results are bit-identical to the reference's: integer scores, the same tie-breaking among equal maxima
 * (PairWiseSW.h:225-251) and the same CIGAR construction (getCIGAR, :269-437).  No CPU path: without an sm_100
 * device every call fails with GKLB_ERR_NO_DEVICE.
 */
#ifndef GKLB_SW_H
#define GKLB_SW_H

#include "gklb_pairhmm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Overhang strategies as the Java wrapper encodes them (IntelSmithWaterman.java:153-170, smithwaterman_common.h:47-50) */
enum { GKLB_SW_SOFTCLIP = 9, GKLB_SW_INDEL = 10, GKLB_SW_LEADING_INDEL = 11, GKLB_SW_IGNORE = 12 };
#define GKLB_SW_MAX_SEQUENCE_LENGTH 32767   /* IntelSmithWaterman.java:53 */

typedef struct gklb_sw_batch {
  int32_t n;                 /* pairs */
  const uint8_t* seq1;       /* reference sequences, concatenated; pair k = [seq1_off[k], seq1_off[k+1]) */
  const int64_t* seq1_off;   /* n + 1 */
  const uint8_t* seq2;       /* alternate sequences */
  const int64_t* seq2_off;   /* n + 1 */
  int32_t match, mismatch, open, extend;   /* SWParameters */
  int32_t strategy;          /* GKLB_SW_* */
} gklb_sw_batch;

typedef struct gklb_sw_stats {
  int64_t pairs;
  int64_t cells;             /* sum of len1 * len2 */
  float h2d_ms, kernel_ms, d2h_ms;
  int32_t kernel_launches;
  int32_t warps;             /* resident warps the launch used (bounded by the backtrack scratch) */
} gklb_sw_stats;

GKLB_API int gklb_sw_init(void);
/* cigars: n rows of cigar_pitch bytes; row k receives pair k's CIGAR, cigar_len[k] bytes, followed by a NUL when
 * the pair's buffer has room for it (the rest of the row is not touched).  Each pair's buffer length is
 * min(cigar_pitch, 2 * max(len1, len2)) -- the array the Java wrapper allocates (IntelSmithWaterman.java:135);
 * elements that do not fit are dropped like getCIGAR drops them. */
GKLB_API int gklb_sw_align_batch(const gklb_sw_batch* batch, char* cigars, int32_t cigar_pitch, int32_t* cigar_len,
                                 int32_t* offsets);
/* runSWOnePairBT (PairWiseSW.h:454): returns GKLB_OK or an error; *cigar_count = strnlen of what was written. */
GKLB_API int gklb_sw_align(int32_t match, int32_t mismatch, int32_t open, int32_t extend, const uint8_t* seq1,
                           const uint8_t* seq2, int32_t len1, int32_t len2, int32_t strategy, char* cigar,
                           int32_t cigar_len, uint32_t* cigar_count, int32_t* offset);
GKLB_API int gklb_sw_done(void);
GKLB_API int gklb_sw_last_stats(gklb_sw_stats* out);
/* Time `iters` kernel launches over the batch of the last gklb_sw_align_batch call (still resident in HBM). */
GKLB_API int gklb_sw_time_runs(int iters, float* ms_per_run);

#ifdef __cplusplus
}
#endif
#endif /* GKLB_SW_H */
