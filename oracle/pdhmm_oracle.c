/*
 * TEST INFRASTRUCTURE ONLY -- a CPU restatement ("port") of GKL's partially-determined-haplotype PairHMM
 * (PDHMM), used solely as the checker in tests/, __graft_entry__.smoke() and bench.py's CPU arm.
 *
 * Plain scalar C following the reference's serial path (the one GATK's Java generated the goldens with):
 *
 *   tables        qualToErrorProb, Jacobian log table, matchToMatchProb (exact 1/ln10)
 *                                                   reference pdhmm/MathUtils.cc:31-109, pdhmm-common.h:149-195
 *   transitions   [mm, indelToMatch, mi, ii, md, dd], negative ins/del/gcp -> INPUT_DATA_ERROR
 *                                                   reference pdhmm/pdhmm-serial.cc:157-225
 *   prior         x==y | x=='N' | y=='N' | (pd&SNP && pd&bit(x)), raw byte compare
 *                                                   reference pdhmm/pdhmm-serial.cc:228-277
 *   recurrence    M/I/D + branch matrices with the NORMAL / INSIDE_DEL / AFTER_DEL column state machine
 *                                                   reference pdhmm/pdhmm-serial.cc:279-404
 *   result        log10(sum_j M[R][j] + I[R][j]) - log10(2^1020)   reference pdhmm/pdhmm-serial.cc:406-411
 *
 * carry_state != 0 reproduces the serial path exactly: `currentState` is declared outside the row loop
 * (pdhmm-serial.cc:306), so the state reached at the end of a row carries into the next one.  carry_state == 0
 * resets it to NORMAL at every row start, which is what the AVX paths do (pdhmm.h:507-511,736-737).
 *
 * Parity pin: tests/test_oracle_pdhmm.py checks this file against the reference's golden files
 * (1e-4 absolute, IntelPDHMMUnitTest.java:33) and against GKL's own compiled serial/AVX code (oracle/_ref).
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
#define MM_SIZE (((MAX_QUAL + 1) * (MAX_QUAL + 2)) >> 1)
#define JAC_SIZE 80001
#define PD_SNP 1
#define PD_DEL_START 2
#define PD_DEL_END 4
#define PD_A 8
#define PD_C 16
#define PD_G 32
#define PD_T 64
enum { ST_NORMAL = 0, ST_INSIDE = 1, ST_AFTER = 2 };
enum { PDHMM_SUCCESS = 0, PDHMM_INPUT_DATA_ERROR = 2, PDHMM_FAILURE = 3 };

static double q2err[MAX_QUAL + 1], mm_tab[MM_SIZE], jac[JAC_SIZE];
static double init_cond, init_cond_log10;
static int ready = 0;

static double approx_log10_sum(double a, double b) { /* MathUtils.cc:89-109 */
  if (a > b) { double t = a; a = b; b = t; }
  if (a == -1e10) return b;
  double diff = b - a;
  if (diff < 8.0) {
    double v = diff * (1.0 / 0.0001);
    int idx = (v > 0.0) ? (int)(v + 0.5) : (int)(v - 0.5);
    return b + jac[idx];
  }
  return b;
}

void gklport_pdhmm_init(void) {
  if (ready) return;
  const double inv_ln10 = 1.0 / log(10);
  for (int k = 0; k < JAC_SIZE; k++) jac[k] = log10(1.0 + pow(10.0, -k * 0.0001));
  for (int i = 0, off = 0; i <= MAX_QUAL; off += ++i)
    for (int j = 0; j <= i; j++) {
      double ls = approx_log10_sum(-0.1 * i, -0.1 * j);
      double l = log1p(-fmin(1.0, pow(10, ls))) * inv_ln10;
      mm_tab[off + j] = pow(10, l);
    }
  for (int i = 0; i <= MAX_QUAL; i++) q2err[i] = pow(10.0, ((double)i) / -10.0);
  init_cond = pow(2, 1020);
  init_cond_log10 = log10(init_cond);
  ready = 1;
}

const double* gklport_pdhmm_q2err(void) { gklport_pdhmm_init(); return q2err; }
const double* gklport_pdhmm_mm(void) { gklport_pdhmm_init(); return mm_tab; }

static int pd_match(int8_t x, int8_t pd) { /* pdhmm-serial.cc:228-252, without the error exit */
  if (!(pd & PD_SNP)) return 0;
  switch (x) {
    case 'A': case 'a': return (pd & PD_A) != 0;
    case 'C': case 'c': return (pd & PD_C) != 0;
    case 'T': case 't': return (pd & PD_T) != 0;
    case 'G': case 'g': return (pd & PD_G) != 0;
    default: return -1;
  }
}

/* One pair.  work: 6 * (H + 1) doubles. */
static double one_pair(const int8_t* hap, const int8_t* pd, int H, const int8_t* rd, const int8_t* q, const int8_t* iq,
                       const int8_t* dq, const int8_t* gc, int R, int carry_state, double* work, int* status) {
  double *M = work, *I = M + (H + 1), *D = I + (H + 1), *bM = D + (H + 1), *bI = bM + (H + 1), *bD = bI + (H + 1);
  memset(work, 0, sizeof(double) * 6 * (size_t)(H + 1));
  const double init = init_cond / H;
  for (int j = 0; j <= H; j++) D[j] = init;
  int state = ST_NORMAL;
  for (int i = 1; i <= R; i++) {
    const int8_t ins = iq[i - 1], del = dq[i - 1], g = gc[i - 1];
    if (ins < 0 || del < 0 || g < 0) { *status = PDHMM_INPUT_DATA_ERROR; }
    const int qi = ins & 0xFF, qd = del & 0xFF;
    const int mn = qi <= qd ? qi : qd, mx = qi <= qd ? qd : qi;
    const double tMM = (MAX_QUAL < mx) ? 1.0 - pow(10, approx_log10_sum(-0.1 * mn, -0.1 * mx)) : mm_tab[((mx * (mx + 1)) >> 1) + mn];
    const double tMI = q2err[ins & 0xFF], tMD = q2err[del & 0xFF];
    const double tIM = 1.0 - q2err[g & 0xFF], tII = q2err[g & 0xFF], tDD = tII;
    const double p_match = 1.0 - q2err[q[i - 1] & 0xFF], p_mis = q2err[q[i - 1] & 0xFF] / 3.0;
    const int8_t x = rd[i - 1];
    double bmmL = 0, bmmDg = 0, bimL = 0, bimDg = 0, bdmL = 0, bdmDg = 0, mmL = 0, mmDg = 0, imL = 0, imDg = 0, dmL = 0, dmDg = 0;
    if (i == 1) dmDg = D[0];
    if (!carry_state) state = ST_NORMAL;
    for (int j = 1; j <= H; j++) {
      const double bmmT = bM[j], bimT = bI[j], bdmT = bD[j], mmT = M[j], imT = I[j], dmT = D[j];
      if (state == ST_NORMAL) { bM[j] = mmL; bD[j] = dmL; bI[j] = imL; }
      else if (state == ST_INSIDE) { bM[j] = bmmL; bD[j] = bdmL; bI[j] = bimL; }
      else {
        bM[j] = fmax(bmmL, mmL); bD[j] = fmax(bdmL, dmL); bI[j] = fmax(bimL, imL);
        mmDg = fmax(mmDg, bmmDg); imDg = fmax(imDg, bimDg); dmDg = fmax(dmDg, bdmDg);
        mmL = fmax(mmL, bmmL); dmL = fmax(dmL, bdmL);
      }
      const int8_t y = hap[j - 1], p = pd[j - 1];
      int pm = 0;
      if (!(x == y || x == 'N' || y == 'N')) {
        pm = pd_match(x, p);
        if (pm < 0) { *status = PDHMM_INPUT_DATA_ERROR; pm = 0; }
      }
      const double prior = (x == y || x == 'N' || y == 'N' || pm) ? p_match : p_mis;
      M[j] = prior * (mmDg * tMM + imDg * tIM + dmDg * tIM);
      D[j] = mmL * tMD + dmL * tDD;
      if (p & PD_DEL_END) I[j] = fmax(bmmT, mmT) * tMI + fmax(bimT, imT) * tII;
      else I[j] = mmT * tMI + imT * tII;
      if (state == ST_AFTER) state = ST_NORMAL;
      if (p & PD_DEL_START) state = ST_INSIDE;
      if (p & PD_DEL_END) state = ST_AFTER;
      bmmDg = bmmT; bimDg = bimT; bdmDg = bdmT; mmDg = mmT; imDg = imT; dmDg = dmT;
      bmmL = bM[j]; bimL = bI[j]; bdmL = bD[j]; mmL = M[j]; imL = I[j]; dmL = D[j];
    }
  }
  double sum = 0.0;
  for (int j = 1; j <= H; j++) sum += M[j] + I[j];
  return log10(sum) - init_cond_log10;
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Flat layout of IntelPDHMM.computePDHMM (IntelPDHMM.java:163-204): pair k uses hap[k * max_hap ...],
 * read[k * max_read ...].  Returns the worst status seen. */
int gklport_pdhmm(const int8_t* hap_bases, const int8_t* hap_pdbases, const int8_t* read_bases, const int8_t* read_qual,
                  const int8_t* read_ins_qual, const int8_t* read_del_qual, const int8_t* gcp, double* result, int64_t n,
                  const int64_t* hap_lengths, const int64_t* read_lengths, int max_read, int max_hap, int carry_state,
                  int n_threads, double* seconds) {
  gklport_pdhmm_init();
  int worst = PDHMM_SUCCESS;
  const int threads = n_threads < 1 ? 1 : n_threads;
  (void)threads;
  const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
  {
#if defined(__SSE__)
    /* The PDHMM never turns flush-to-zero on (only GKL's PairHMM init does, IntelPairHmm.cc:93-96) and its
     * lowest likelihoods live in the fp64 denormal range: make sure a PairHMM run earlier in this process did
     * not leave FTZ set on this (pooled) thread. */
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_OFF);
#endif
    double* work = (double*)malloc(sizeof(double) * 6 * (size_t)(max_hap + 1));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
    for (int64_t k = 0; k < n; k++) {
      int st = PDHMM_SUCCESS;
      const int64_t ho = k * (int64_t)max_hap, ro = k * (int64_t)max_read;
      result[k] = one_pair(hap_bases + ho, hap_pdbases + ho, (int)hap_lengths[k], read_bases + ro, read_qual + ro,
                           read_ins_qual + ro, read_del_qual + ro, gcp + ro, (int)read_lengths[k], carry_state, work, &st);
      if (!(result[k] <= 0.0)) st = PDHMM_FAILURE; /* pdhmm-serial.cc:432-441: above 0 or not a number */
      if (st != PDHMM_SUCCESS) {
#ifdef _OPENMP
#pragma omp critical
#endif
        worst = st;
      }
    }
    free(work);
  }
  if (seconds) *seconds = now_s() - t0;
  return worst;
}
