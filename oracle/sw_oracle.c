/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * Plain-C restatement of GKL's Smith-Waterman with backtrack (reference smithwaterman/PairWiseSW.h; citations are
 * relative to /root/reference/src/main/native).  Pinned against the reference's own compiled code (oracle/_ref,
 * gklref_sw) on the pairs of src/test/resources/smith-waterman.SOFTCLIP.in for all four overhang strategies, and
 * against the two known answers of SmithWatermanUnitTest.java:160-190 ("1M", "1M1I").
 *
 * The reference fills the matrix along anti-diagonals with AVX vectors; cell values do not depend on the order, so
 * this file fills row by row.  What does depend on the order -- the choice among equal maxima on the last row and
 * column -- is scanned afterwards in the reference's anti-diagonal order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* smithwaterman_common.h:42-50,81-82 */
enum { SW_MATCH = 0, SW_INSERT = 1, SW_DELETE = 2, SW_INSERT_EXT = 4, SW_DELETE_EXT = 8,
       SW_SOFTCLIP = 9, SW_INDEL = 10, SW_LEADING_INDEL = 11, SW_IGNORE = 12 };
#define SW_MATRIX_MIN_CUTOFF (-100000000)
#define SW_LOW_INIT_VALUE (INT32_MIN / 2)

static int iabs(int x) { return x < 0 ? -x : x; }

/* smithwaterman_common.cc:26-58 */
static int sw_itoa(char* ptr, int number) {
  int neg = 0;
  if (number < 0) { number = -number; neg = 1; }
  int cp = number, digits = 0;
  while (cp > 0) { cp /= 10; digits++; }
  if (!ptr) return digits + neg;
  if (neg) *(ptr++) = '-';
  for (int i = digits - 1; i >= 0; i--) { ptr[i] = (char)('0' + number % 10); number /= 10; }
  return digits + neg;
}

/* One pair.  seq1 = reference (rows, i), seq2 = alternate (columns, j).  Returns 0, or 1 when out of memory. */
static int sw_one(int match, int mismatch, int open, int extend, const uint8_t* seq1, const uint8_t* seq2, int nrow,
                  int ncol, int strategy, char* cigar, int cigar_cap, int32_t* cigar_count, int32_t* offset_out) {
  const size_t W = (size_t)ncol + 1;
  uint8_t* bt = (uint8_t*)malloc((size_t)(nrow + 1) * W);
  int32_t* Hprev = (int32_t*)malloc(W * sizeof(int32_t));
  int32_t* Hcur = (int32_t*)malloc(W * sizeof(int32_t));
  int32_t* F = (int32_t*)malloc(W * sizeof(int32_t));
  int32_t* lastcol = (int32_t*)malloc((size_t)(nrow + 1) * sizeof(int32_t));
  int16_t* el = (int16_t*)malloc(((size_t)nrow + ncol + 4) * 2 * sizeof(int16_t));
  if (!bt || !Hprev || !Hcur || !F || !lastcol || !el) {
    free(bt); free(Hprev); free(Hcur); free(F); free(lastcol); free(el);
    return 1;
  }
  const int indel_edges = (strategy == SW_INDEL) || (strategy == SW_LEADING_INDEL);
  /* row 0: PairWiseSW.h:212-221 (edge cells of every anti-diagonal), F = lowInitValue (:86-89,222) */
  Hprev[0] = 0;
  for (int j = 1; j <= ncol; j++) Hprev[j] = indel_edges ? open + (j - 1) * extend : 0;
  for (int j = 0; j <= ncol; j++) F[j] = SW_LOW_INIT_VALUE;
  for (int i = 1; i <= nrow; i++) {
    Hcur[0] = indel_edges ? open + (i - 1) * extend : 0;
    int32_t E = SW_LOW_INIT_VALUE;  /* :90-93,223 */
    for (int j = 1; j <= ncol; j++) {
      /* MAIN_CODE, PairWiseSW.h:27-62 */
      const int32_t ext_h = E + extend, open_h = Hcur[j - 1] + open;
      const int32_t e11 = open_h > ext_h ? open_h : ext_h;
      int code = (open_h > ext_h) ? 0 : SW_INSERT_EXT;
      const int32_t ext_v = F[j] + extend, open_v = Hprev[j] + open;
      const int32_t f11 = ext_v > open_v ? ext_v : open_v;
      if (!(open_v > ext_v)) code |= SW_DELETE_EXT;
      const int32_t m11 = Hprev[j - 1] + (seq1[i - 1] == seq2[j - 1] ? match : mismatch);
      int32_t h11 = m11 > SW_MATRIX_MIN_CUTOFF ? m11 : SW_MATRIX_MIN_CUTOFF;
      int dir = SW_MATCH;
      if (e11 > h11) { dir = SW_INSERT; h11 = e11; }
      if (f11 > h11) { dir = SW_DELETE; h11 = f11; }
      bt[(size_t)i * W + j] = (uint8_t)(dir | code);
      E = e11;
      F[j] = f11;
      Hcur[j] = h11;
    }
    lastcol[i] = Hcur[ncol];
    int32_t* t = Hprev; Hprev = Hcur; Hcur = t;
  }
  const int32_t* lastrow = Hprev;  /* H(nrow, j) */
  /* PairWiseSW.h:225-251: candidates in anti-diagonal order; on one anti-diagonal the last-row cell comes first */
  int32_t maxScore = INT32_MIN;
  int max_i = 0, max_j = 0;
  for (int ad = 1; ad <= nrow + ncol; ad++) {
    if (ad >= nrow + 1 && (strategy == SW_SOFTCLIP || strategy == SW_IGNORE)) {
      const int i = nrow, j = ad - nrow;
      const int32_t score = lastrow[j];
      if (maxScore < score || (maxScore == score && iabs(i - j) < iabs(max_i - max_j))) {
        maxScore = score; max_i = i; max_j = j;
      }
    }
    if (ad >= ncol + 1) {
      const int i = ad - ncol, j = ncol;
      const int32_t score = lastcol[i];
      if (maxScore < score || (maxScore == score && (max_j == ncol || iabs(i - j) <= iabs(max_i - max_j)))) {
        maxScore = score; max_i = i; max_j = j;
      }
    }
  }
  /* getCIGAR, PairWiseSW.h:269-437 */
  int i, j, n_el = 0;
  if (strategy == SW_INDEL) { i = nrow; j = ncol; }
  else if (strategy == SW_LEADING_INDEL) { i = max_i; j = ncol; }
  else { i = max_i; j = max_j; }
  if (j < ncol) { el[0] = SW_SOFTCLIP; el[1] = (int16_t)(ncol - j); n_el = 1; }
  int state = 0;
  while (i > 0 && j > 0) {
    const int btr = bt[(size_t)i * W + j];
    if (state == SW_INSERT_EXT) { j--; el[n_el * 2 - 1]++; state = btr & SW_INSERT_EXT; }
    else if (state == SW_DELETE_EXT) { i--; el[n_el * 2 - 1]++; state = btr & SW_DELETE_EXT; }
    else {
      switch (btr & 3) {
        case SW_MATCH: i--; j--; el[n_el * 2] = SW_MATCH; el[n_el * 2 + 1] = 1; state = 0; n_el++; break;
        case SW_INSERT: j--; el[n_el * 2] = SW_INSERT; el[n_el * 2 + 1] = 1; state = btr & SW_INSERT_EXT; n_el++; break;
        case SW_DELETE: i--; el[n_el * 2] = SW_DELETE; el[n_el * 2 + 1] = 1; state = btr & SW_DELETE_EXT; n_el++; break;
      }
    }
  }
  int offset;
  if (strategy == SW_SOFTCLIP) {
    if (j > 0) { el[n_el * 2] = SW_SOFTCLIP; el[n_el * 2 + 1] = (int16_t)j; n_el++; }
    offset = i;
  } else if (strategy == SW_IGNORE) {
    if (j > 0) { el[n_el * 2] = el[(n_el - 1) * 2]; el[n_el * 2 + 1] = (int16_t)j; n_el++; }
    offset = (int16_t)(i - j);
  } else {
    if (i > 0) { el[n_el * 2] = SW_DELETE; el[n_el * 2 + 1] = (int16_t)i; n_el++; }
    else if (j > 0) { el[n_el * 2] = SW_INSERT; el[n_el * 2 + 1] = (int16_t)j; n_el++; }
    offset = 0;
  }
  int newId = 0;
  int16_t prev = el[0];
  for (int k = 1; k < n_el; k++) {
    const int16_t cur = el[k * 2];
    if (cur == prev) el[newId * 2 + 1] = (int16_t)(el[newId * 2 + 1] + el[k * 2 + 1]);
    else { newId++; el[newId * 2] = cur; el[newId * 2 + 1] = el[k * 2 + 1]; prev = cur; }
  }
  int cur_size = 0;
  for (int k = newId; k >= 0; k--) {
    char st;
    switch (el[2 * k]) {
      case SW_MATCH: st = 'M'; break;
      case SW_INSERT: st = 'I'; break;
      case SW_DELETE: st = 'D'; break;
      case SW_SOFTCLIP: st = 'S'; break;
      default: st = 'R'; break;
    }
    const int expected = sw_itoa(NULL, el[2 * k + 1]) + 1;
    if (cur_size >= 0 && expected > 1 && cur_size + expected <= cigar_cap) {
      cur_size += sw_itoa(cigar + cur_size, el[2 * k + 1]);
      cigar[cur_size++] = st;
    }
  }
  *cigar_count = (int32_t)strnlen(cigar, (size_t)cur_size);
  *offset_out = offset;
  free(bt); free(Hprev); free(Hcur); free(F); free(lastcol); free(el);
  return 0;
}

static double sw_now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Batch entry point, same shape as gklref_sw (oracle/ref_driver_sw.cc). */
int gklport_sw(int n, const uint8_t* seq1, const int64_t* off1, const uint8_t* seq2, const int64_t* off2, int match,
               int mismatch, int open, int extend, int strategy, int n_threads, char* cigars, int pitch,
               int32_t* cigar_len, int32_t* offsets, double* seconds) {
  memset(cigars, 0, (size_t)n * (size_t)pitch);
  int failed = 0;
  const int threads = n_threads < 1 ? 1 : n_threads;
  (void)threads;
  const double t0 = sw_now();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
  for (int k = 0; k < n; k++) {
    const int len1 = (int)(off1[k + 1] - off1[k]), len2 = (int)(off2[k + 1] - off2[k]);
    int cap = 2 * (len1 > len2 ? len1 : len2);
    if (cap > pitch) cap = pitch;
    const int rc = sw_one(match, mismatch, open, extend, seq1 + off1[k], seq2 + off2[k], len1, len2, strategy,
                          cigars + (size_t)k * pitch, cap, &cigar_len[k], &offsets[k]);
    if (rc) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      failed = rc;
    }
  }
  if (seconds) *seconds = sw_now() - t0;
  return failed;
}
