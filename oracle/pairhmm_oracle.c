/*
 * TEST INFRASTRUCTURE ONLY -- a CPU restatement ("port") of GKL's PairHMM forward
 * likelihood, used solely as the checker in tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py.  Nothing under gkl_b200/ may import, link or call it.
 *
 * Plain scalar C, written from the algorithm (not from GKL's striped SIMD code):
 *
 *   tables      ph2pr, jacobianLogTable, matchToMatchProb   reference pairhmm/Context.h:65-89,133-148,174-189
 *   set_mm_prob triangular table lookup                     reference pairhmm/Context.h:156-167,197-209
 *   base codes  A0 C1 T2 G3 N4, everything else 0           reference pairhmm/pairhmm_common.h:49-67
 *   per-row probabilities (quals & 127)                     reference pairhmm/avx-pairhmm-template.h:106-152
 *   recurrence  M/X/Y cell form of computeMXY               reference pairhmm/avx-pairhmm-template.h:208-223
 *   boundary    row 0: M=X=0, Y=INITIAL_CONSTANT/haplen      reference pairhmm/avx-pairhmm-template.h:114-121,160-202
 *   result      sum over last row of M + X                  reference pairhmm/avx-pairhmm-template.h:325-371
 *   fallback    fp32 < 1e-28f -> fp64 rerun, log10 - const  reference pairhmm/IntelPairHmm.cc:150-169
 *
 * Parity pin: tests/test_oracle.py checks this file against all 104 rows of GKL's
 * pairhmm-testdata.txt (both precisions, 1e-5 abs, as PairHmmUnitTest.dataFileTest does),
 * against the simpleTest known answer, and against GKL's own compiled AVX code
 * (oracle/_ref) on random batches.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#if defined(__SSE__)
#include <xmmintrin.h>
#endif

#define MAX_QUAL 254
#define JAC_TOL 8.0
#define JAC_STEP 0.0001
#define JAC_INV_STEP (1.0 / JAC_STEP)
#define JAC_SIZE 80001 /* (int)(8.0 / 0.0001) + 1 */
#define MM_SIZE (((MAX_QUAL + 1) * (MAX_QUAL + 2)) >> 1)
#define MIN_ACCEPTED 1e-28f

static float ph2pr_f[128], jac_f[JAC_SIZE], mm_f[MM_SIZE];
static double ph2pr_d[128], jac_d[JAC_SIZE], mm_d[MM_SIZE];
static float init_const_f, log10_init_f;
static double init_const_d, log10_init_d;
static uint8_t base_code[256];
static int tables_ready = 0;

/* Context.h:91-122, instantiated for float */
static float approx_log10_sum_f(float small, float big) {
  if (small > big) { float t = big; big = small; small = t; }
  if (isinf(small) || isinf(big)) return big;
  float diff = big - small;
  if (diff >= (float)JAC_TOL) return big;
  float v = (float)(diff * ((float)JAC_INV_STEP));
  int ind = (v > 0.0f) ? (int)(v + 0.5f) : (int)(v - 0.5f);
  return big + jac_f[ind];
}

/* Context.h:91-122, instantiated for double */
static double approx_log10_sum_d(double small, double big) {
  if (small > big) { double t = big; big = small; small = t; }
  if (isinf(small) || isinf(big)) return big;
  double diff = big - small;
  if (diff >= JAC_TOL) return big;
  double v = diff * JAC_INV_STEP;
  int ind = (v > 0.0) ? (int)(v + 0.5) : (int)(v - 0.5);
  return big + jac_d[ind];
}

void gklport_init_tables(void) {
  if (tables_ready) return;
  /* Context.h:65-72 */
  for (int k = 0; k < JAC_SIZE; k++) {
    double v = log10(1.0 + pow(10.0, -((double)k) * JAC_STEP));
    jac_f[k] = (float)v;
    jac_d[k] = v;
  }
  /* Context.h:75-89: note the truncated constant INV_LN10 = 0.434294 */
  const double INV_LN10 = 0.434294;
  for (int i = 0, offset = 0; i <= MAX_QUAL; offset += ++i) {
    for (int j = 0; j <= i; j++) {
      double ls_f = approx_log10_sum_f((float)-0.1 * (float)i, (float)-0.1 * (float)j);
      double m_f = log1p(-fmin(1.0, pow(10, ls_f))) * INV_LN10;
      mm_f[offset + j] = (float)pow(10, m_f);
      double ls_d = approx_log10_sum_d((double)-0.1 * (double)i, (double)-0.1 * (double)j);
      double m_d = log1p(-fmin(1.0, pow(10, ls_d))) * INV_LN10;
      mm_d[offset + j] = pow(10, m_d);
    }
  }
  /* Context.h:137-143 and :178-184 */
  for (int x = 0; x < 128; x++) {
    ph2pr_d[x] = pow(10.0, -((double)x) / 10.0);
    ph2pr_f[x] = powf(10.f, -((float)x) / 10.f);
  }
  init_const_d = ldexp(1.0, 1020);
  log10_init_d = log10(init_const_d);
  init_const_f = ldexpf(1.f, 120);
  log10_init_f = log10f(init_const_f);
  /* pairhmm_common.h:53-62: zero-initialised table, five letters set */
  memset(base_code, 0, sizeof(base_code));
  base_code['A'] = 0; base_code['C'] = 1; base_code['T'] = 2; base_code['G'] = 3; base_code['N'] = 4;
  tables_ready = 1;
}

/* Context.h:197-209 (float) / :156-167 (double); quals are already & 127 so the
 * MAX_QUAL < maxQual branch is unreachable. */
static inline int mm_index(int ins_q, int del_q) {
  int mn = del_q, mx = ins_q;
  if (ins_q <= del_q) { mn = ins_q; mx = del_q; }
  return ((mx * (mx + 1)) >> 1) + mn;
}

#define DEFINE_FULL_PROB(NAME, NUM, PH2PR, MMTAB, INITC)                                         \
  static NUM NAME(int rslen, int haplen, const uint8_t* rs, const uint8_t* q, const uint8_t* ig, \
                  const uint8_t* dg, const uint8_t* cg, const uint8_t* hap, NUM* work) {         \
    const int C = haplen + 1;                                                                    \
    NUM* Mp = work;          NUM* Xp = work + C;     NUM* Yp = work + 2 * C;                      \
    NUM* Mc = work + 3 * C;  NUM* Xc = work + 4 * C; NUM* Yc = work + 5 * C;                      \
    const NUM init_Y = INITC / (NUM)haplen;                                                      \
    for (int c = 0; c < C; c++) { Mp[c] = 0; Xp[c] = 0; Yp[c] = init_Y; }                        \
    NUM sumM = 0, sumX = 0;                                                                      \
    for (int r = 1; r <= rslen; r++) {                                                           \
      const int _i = ig[r - 1] & 127, _d = dg[r - 1] & 127, _c = cg[r - 1] & 127;                \
      const int _q = q[r - 1] & 127;                                                             \
      const NUM pMM = MMTAB[mm_index(_i, _d)];                                                   \
      const NUM pGAPM = (NUM)1.0 - PH2PR[_c];                                                    \
      const NUM pMX = PH2PR[_i], pXX = PH2PR[_c], pMY = PH2PR[_d], pYY = PH2PR[_c];              \
      const NUM distm = PH2PR[_q];                                                               \
      const NUM one_minus = (NUM)1.0 - distm; /* stripeINITIALIZATION :181-183 */                \
      const NUM third = distm / (NUM)3.0;                                                        \
      const uint8_t rb = base_code[rs[r - 1]];                                                   \
      Mc[0] = 0; Xc[0] = 0; Yc[0] = 0;                                                           \
      for (int c = 1; c < C; c++) {                                                              \
        const uint8_t hb = base_code[hap[c - 1]];                                                \
        const int match = (rb == hb) || (rb == 4) || (hb == 4);                                  \
        const NUM prior = match ? one_minus : third;                                             \
        Mc[c] = ((Mp[c - 1] * pMM + Xp[c - 1] * pGAPM) + Yp[c - 1] * pGAPM) * prior;             \
        Xc[c] = Mp[c] * pMX + Xp[c] * pXX;                                                       \
        Yc[c] = Mc[c - 1] * pMY + Yc[c - 1] * pYY;                                               \
      }                                                                                          \
      NUM* t;                                                                                    \
      t = Mp; Mp = Mc; Mc = t;  t = Xp; Xp = Xc; Xc = t;  t = Yp; Yp = Yc; Yc = t;               \
    }                                                                                            \
    for (int c = 1; c < C; c++) { sumM += Mp[c]; sumX += Xp[c]; }                                \
    return sumM + sumX;                                                                          \
  }

DEFINE_FULL_PROB(full_prob_f, float, ph2pr_f, mm_f, init_const_f)
DEFINE_FULL_PROB(full_prob_d, double, ph2pr_d, mm_d, init_const_d)

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* One pair, IntelPairHmm.cc:154-167.  *fell_back reports whether the fp64 rerun was taken. */
static double one_pair(int rslen, int haplen, const uint8_t* rs, const uint8_t* q, const uint8_t* ig,
                       const uint8_t* dg, const uint8_t* cg, const uint8_t* hap, int use_double,
                       void* work, int* fell_back) {
  float rf = use_double ? 0.0f : full_prob_f(rslen, haplen, rs, q, ig, dg, cg, hap, (float*)work);
  if (rf < MIN_ACCEPTED) {
    double rd = full_prob_d(rslen, haplen, rs, q, ig, dg, cg, hap, (double*)work);
    if (fell_back) *fell_back = 1;
    return log10(rd) - log10_init_d;
  }
  if (fell_back) *fell_back = 0;
  return (double)(log10f(rf) - log10_init_f);
}

/* Same flat batch layout as include/gklb_pairhmm.h; out[r * n_haps + h] (JavaData.h:94-105).
 * fallback (optional, may be NULL) receives 1 for every pair that took the fp64 rerun. */
int gklport_pairhmm(int n_reads, int n_haps, const int64_t* read_off, const uint8_t* read_bases,
                    const uint8_t* read_quals, const uint8_t* ins_gop, const uint8_t* del_gop,
                    const uint8_t* gcp, const int64_t* hap_off, const uint8_t* hap_bases,
                    int use_double, int n_threads, double* out, uint8_t* fallback, double* seconds) {
  gklport_init_tables();
  int max_hap = 0;
  for (int h = 0; h < n_haps; h++) {
    int l = (int)(hap_off[h + 1] - hap_off[h]);
    if (l > max_hap) max_hap = l;
  }
  const int threads = n_threads < 1 ? 1 : n_threads; /* honoured even under OMP_NUM_THREADS=1 (torchrun) */
  const long n = (long)n_reads * (long)n_haps;
  double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
#if defined(__SSE__)
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON); /* IntelPairHmm.cc:93-96 */
#endif
    void* work = malloc(sizeof(double) * 6 * (size_t)(max_hap + 1));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (long i = 0; i < n; i++) {
      const int r = (int)(i / n_haps), h = (int)(i % n_haps);
      const int64_t ro = read_off[r], ho = hap_off[h];
      int fb = 0;
      out[i] = one_pair((int)(read_off[r + 1] - ro), (int)(hap_off[h + 1] - ho), read_bases + ro,
                        read_quals + ro, ins_gop + ro, del_gop + ro, gcp + ro, hap_bases + ho,
                        use_double, work, &fb);
      if (fallback) fallback[i] = (uint8_t)fb;
    }
    free(work);
  }
  if (seconds) *seconds = now_s() - t0;
  return 0;
}

/* Table accessors so tests can pin the engine's host-built tables bit-for-bit. */
const float* gklport_ph2pr_f(void) { gklport_init_tables(); return ph2pr_f; }
const float* gklport_mm_f(void) { gklport_init_tables(); return mm_f; }
const double* gklport_ph2pr_d(void) { gklport_init_tables(); return ph2pr_d; }
const double* gklport_mm_d(void) { gklport_init_tables(); return mm_d; }
int gklport_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
