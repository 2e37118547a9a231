// pdhmm_device.cuh -- sm_100a device code of the partially-determined-haplotype PairHMM (PDHMM, config 5).
//
// Reference semantics (/root/reference/src/main/native/pdhmm): the scalar path pdhmm-serial.cc:279-411 --
// M/I/D plus the three "branch" matrices, a NORMAL / INSIDE_DEL / AFTER_DEL state per haplotype column driven by
// the PD flag bytes (DEL_START=2, DEL_END=4), SNP alleles (SNP=1, A=8, C=16, G=32, T=64) widening the match test,
// fp64 throughout, INITIAL_CONDITION = 2^1020, result log10(sum_j M[R][j] + I[R][j]) - log10(2^1020).
// `currentState` is declared outside the row loop there (pdhmm-serial.cc:306), so the state after the last column
// of a row is the state the next row starts in; that is reproduced (carry_state = 1).  carry_state = 0 resets
// it per row like the reference's AVX paths (pdhmm.h:507-511,736-737).
//
// Mapping: the same systolic sweep as the PairHMM kernels -- a group of G lanes owns G*K read rows (K per lane,
// padded at the TOP with rows that reproduce row 0: M = I = 0, D = init), lane t processes column s - t at step
// s, the bottom row of lane t-1 arrives by warp shuffle (six doubles: M, I, D and their branch twins), reads
// longer than G*K rows take several passes with the bottom row carried through a global scratch line.
// The column state does not depend on the data, only on the PD bytes and on the state the row started in, so it
// is precomputed per pair: one byte per column holds the state for each of the three possible row-start states
// plus the DEL_END flag, and each row's start state follows from the 3-entry end-state map.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace gklb {

constexpr int kPdMargin = 40;  // zero columns on both sides of a haplotype in shared memory

struct PdhmmParams {
  // flat layout (IntelPDHMM.computePDHMM): pair k at k * max_hap / k * max_read; or cross layout
  // (IntelPDHMM.computeLikelihoods): pair k = r * n_haps + h with reads/haps stored once
  const int8_t* hap_bases;
  const int8_t* hap_pdbases;
  const int8_t* read_bases;
  const int8_t* read_qual;
  const int8_t* read_ins_qual;
  const int8_t* read_del_qual;
  const int8_t* gcp;
  const int64_t* hap_lengths;   // [n] flat, [n_haps] cross
  const int64_t* read_lengths;  // [n] flat, [n_reads] cross
  long long n;                  // pairs
  int n_haps;                   // cross layout: haplotypes per read; 0 = flat layout
  int max_hap, max_read;
  const double* q2err;          // [255]   10^(-q/10)
  const double* mm;             // [32640] matchToMatchProb
  double init_cond;             // 2^1020
  double log10_init;
  double* out;                  // [n]
  unsigned int* counter;        // work counter
  unsigned int* error_flag;     // bit 0: a negative insertion/deletion/gcp quality; bit 1: a result above 0 or NaN
  double* carry;                // per warp: GPW * 2 * 6 * (max_hap + 2) doubles
  size_t carry_stride;          // doubles per warp
  int carry_state;
};

// The three cell updates with a fixed evaluation order (the source order of pdhmm-serial.cc:354-365, products fused
// into the following sum), so that every kernel and code path rounds identically.
__device__ __forceinline__ double pd_match(double prior, double dM, double dI, double dD, double tMM, double tIM) {
  return __dmul_rn(prior, fma(dD, tIM, fma(dI, tIM, __dmul_rn(dM, tMM))));
}
__device__ __forceinline__ double pd_gap(double a, double ca, double b, double cb) {  // a * ca + b * cb
  return fma(b, cb, __dmul_rn(a, ca));
}

// pdhmm-serial.cc:411,432-441: log10(sum) - log10(initial condition); a result above 0 or a NaN is PDHMM_FAILURE
__device__ __forceinline__ void pd_store_result(const PdhmmParams& p, long long item, double sum, int shift = 0) {
  // shift: the kernel ran with the initial condition lowered by that many bits (0.30102999566398120 = log10 2)
  const double r = log10(sum) - (p.log10_init - (double)shift * 0.30102999566398120);
  if (!(r <= 0.0)) atomicOr(p.error_flag, 2u);
  p.out[item] = r;
}

__device__ __forceinline__ double shfl_up_d(double v, int width) { return __shfl_up_sync(0xffffffffu, v, 1, width); }

// Read bytes fall into ten classes for the prior's match test (pdhmm-serial.cc:228-277: raw byte equality, 'N' on
// either side, or an SNP allele of the column matching the read letter case-insensitively): A C G T a c g t N other.
// cmask[c] bit k tells whether a read byte of class k matches column c; "other" bytes match only an 'N' column or
// an identical byte (the latter is tested explicitly).
__device__ __forceinline__ uint32_t read_class(uint32_t x) {
  switch (x) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
    case 'a': return 4; case 'c': return 5; case 'g': return 6; case 't': return 7;
    case 'N': return 8; default: return 9;
  }
}
__device__ __forceinline__ uint32_t column_mask(uint32_t y, uint32_t al) {
  if (y == 'N') return 0x3FFu;
  uint32_t m = 0x100u;  // a read 'N' matches every column
  const uint32_t letters[4] = {'A', 'C', 'G', 'T'};
  const uint32_t abit[4] = {8u, 16u, 32u, 64u};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (y == letters[k] || (al & abit[k])) m |= 1u << k;
    if (y == (letters[k] | 0x20u) || (al & abit[k])) m |= 1u << (k + 4);
  }
  return m;
}

// Column tables of one haplotype in shared memory, built by the G lanes of a group (lane t of G): the haplotype byte,
// the SNP allele bits, the class mask of the match test, the state code of every column for each of the three
// possible row-start states (+ DEL_END in bit 6, "twins written here are read later" in bit 7) and, for every
// column, the first column at or after it that needs the state machine.  Returns the packed end states (the state
// after the last column for each start state; pdhmm-serial.cc:370-385), valid on every lane.
__device__ __forceinline__ int pd_build_column_tables(int t, int G, int group_leader_lane, int H, int max_hap,
                                                      const int8_t* hap, const int8_t* pd, int carry_state, uint8_t* ys,
                                                      uint8_t* infos, uint8_t* alleles, uint16_t* nspec, uint16_t* cmask) {
  __syncwarp();
  for (int c = t - kPdMargin; c < max_hap + kPdMargin; c += G) {
    uint8_t y = 0, al = 0;
    if (c >= 1 && c <= H) {
      y = (uint8_t)hap[c - 1];
      const uint8_t f = (uint8_t)pd[c - 1];
      al = (f & 1) ? (f & 0x78) : 0;
    }
    ys[c] = y;
    alleles[c] = al;
    infos[c] = 0;
    cmask[c] = (uint16_t)((c >= 1 && c <= H) ? column_mask(y, al) : 0u);
  }
  __syncwarp();
  int end_state = 0;  // packed 2-bit end states for the three start states
  if (t == 0) {
    int s0 = 0, s1 = 1, s2 = 2;  // state entering column c when the row started NORMAL / INSIDE / AFTER
    for (int c = 1; c <= H; c++) {
      const uint8_t f = (uint8_t)pd[c - 1];
      infos[c] = (uint8_t)(s0 | (s1 << 2) | (s2 << 4) | ((f & 4) ? 0x40 : 0) | ((s0 != 1 && (f & 6)) ? 0x80 : 0));
      // pdhmm-serial.cc:370-385
      if (s0 == 2) s0 = 0;
      if (s1 == 2) s1 = 0;
      if (s2 == 2) s2 = 0;
      if (f & 2) s0 = s1 = s2 = 1;
      if (f & 4) s0 = s1 = s2 = 2;
    }
    end_state = s0 | (s1 << 2) | (s2 << 4);
    // which row-start states can occur at all: NORMAL, then the end state of the previous row, ...
    int reach = 1, st = 0;
    if (carry_state)
      for (int r = 0; r < 3; r++) { st = (end_state >> (2 * st)) & 3; reach |= 1 << st; }
    const uint32_t rmask = ((reach & 1) ? 0x03u : 0u) | ((reach & 2) ? 0x0Cu : 0u) | ((reach & 4) ? 0x30u : 0u) | 0x40u;
    // a column is "plain" when every reachable start state sees it NORMAL and it does not close a deletion;
    // nspec[c] = the first column >= c that is not plain (0xFFFF if none)
    uint32_t nxt = 0xFFFFu;
    for (int c = max_hap + kPdMargin - 1; c >= -kPdMargin; c--) {
      if (c >= 1 && c <= H && (infos[c] & rmask)) nxt = (uint32_t)c;
      nspec[c] = (uint16_t)nxt;
    }
  }
  end_state = __shfl_sync(0xffffffffu, end_state, group_leader_lane);
  __syncwarp();
  return end_state;
}

// Smem per group: y[c], info[c], allele[c] (bytes) and nspec[c] (uint16: first column >= c that needs the state
// machine) for c in [-kPdMargin, max_hap + kPdMargin), plus cmask[c] (uint16, see read_class): 7 bytes per column
// MULTI = false: the host guarantees max_read <= G*K, so every pair takes one pass and the carry line does not exist.
template <int G, int K, int WARPS, bool MULTI>
__global__ void __launch_bounds__(WARPS * 32, 1) k_pdhmm(const PdhmmParams p) {
  constexpr int GPW = 32 / G;
  constexpr int CAP = G * K;
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane % G, g = lane / G;
  const int col_pitch = (p.max_hap + 2 * kPdMargin + 1) & ~1;  // even: the uint16 prefix array follows three byte arrays
  uint8_t* gs = smem + (size_t)(warp * GPW + g) * 7 * col_pitch;
  uint8_t* ys = gs + kPdMargin;                  // haplotype byte of column c at ys[c] (c = 1..H)
  uint8_t* infos = gs + col_pitch + kPdMargin;   // state bits + DEL_END
  uint8_t* alleles = gs + 2 * col_pitch + kPdMargin;
  uint16_t* nspec = reinterpret_cast<uint16_t*>(gs + 3 * col_pitch) + kPdMargin;  // col_pitch is even
  uint16_t* cmask = reinterpret_cast<uint16_t*>(gs + 5 * col_pitch) + kPdMargin;  // which read-byte classes match column c
  const int carry_pitch = p.max_hap + 2;
  double* carry_g = p.carry + ((size_t)blockIdx.x * WARPS + warp) * p.carry_stride + (size_t)g * 12 * carry_pitch;
  const unsigned long long n_warp_items = ((unsigned long long)p.n + GPW - 1) / GPW;

  for (;;) {
    unsigned int wi = 0;
    if (lane == 0) wi = atomicAdd(p.counter, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= n_warp_items) break;
    const long long item = (long long)wi * GPW + g;
    const bool mine = item < p.n;
    long long hi = 0, ri = 0;
    if (mine) {
      if (p.n_haps > 0) { ri = item / p.n_haps; hi = item - ri * p.n_haps; }
      else { ri = item; hi = item; }
    }
    const int H = mine ? (int)p.hap_lengths[hi] : 0;
    const int R = mine ? (int)p.read_lengths[ri] : 0;
    const int8_t* hap = p.hap_bases + hi * p.max_hap;
    const int8_t* pd = p.hap_pdbases + hi * p.max_hap;
    const int64_t ro = ri * (int64_t)p.max_read;

    // ---- per-pair column tables in shared memory ----
    const int end_state = pd_build_column_tables(t, G, g * G, H, p.max_hap, hap, pd, p.carry_state, ys, infos, alleles,
                                                 nspec, cmask);

    const int n_pass = MULTI ? max(1, (R + CAP - 1) / CAP) : 1;
    int n_pass_w = n_pass;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) n_pass_w = max(n_pass_w, __shfl_xor_sync(0xffffffffu, n_pass_w, o));
    const int n_pad = n_pass * CAP - R;
    const double init = p.init_cond / (double)max(H, 1);
    int n_steps = mine ? H + G - 1 : 0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) n_steps = max(n_steps, __shfl_xor_sync(0xffffffffu, n_steps, o));
    double sum = 0.0;

    for (int pass = 0; pass < n_pass_w; pass++) {
      const bool live = mine && pass < n_pass;
      // ---- per-row constants ----
      double tMM[K], tIM[K], tMI[K], tII[K], tMD[K], pMa[K], pMi[K];
      uint32_t xb[K], xbit[K], shift[K];  // read byte, its allele bit, 2 * row-start state
      uint32_t rcls[K], xeq[K];           // read-byte class; the byte itself when the class is "other" (else no byte)
      bool padrow[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int row = pass * CAP + t * K + j - n_pad;  // 0-based read row; < 0: top padding
        padrow[j] = !(live && row >= 0);
        tMM[j] = tIM[j] = tMI[j] = tMD[j] = 0.0;
        tII[j] = 1.0;
        pMa[j] = pMi[j] = 0.0;
        xb[j] = 0x100;  // matches nothing
        xbit[j] = 0;
        shift[j] = 0;
        rcls[j] = 15;   // no class bit is ever set there
        xeq[j] = 0x100;
        if (!padrow[j]) {
          const int8_t iq = p.read_ins_qual[ro + row], dq = p.read_del_qual[ro + row], gq = p.gcp[ro + row];
          if (iq < 0 || dq < 0 || gq < 0) atomicOr(p.error_flag, 1u);  // pdhmm-serial.cc:184-198
          const int qi = iq & 0xFF, qd = dq & 0xFF, qg = gq & 0xFF, qq = p.read_qual[ro + row] & 0xFF;
          const int mn = min(qi, qd), mx = max(qi, qd);
          // MAX_QUAL (254) < maxQual only for 255: 1 - 10^(approximate log10 sum) -- evaluated directly
          tMM[j] = (mx > 254) ? 1.0 - (pow(10.0, -0.1 * mn) + pow(10.0, -0.1 * mx))
                              : __ldg(p.mm + ((mx * (mx + 1)) >> 1) + mn);
          tMI[j] = __ldg(p.q2err + min(qi, 254));
          tMD[j] = __ldg(p.q2err + min(qd, 254));
          const double eg = __ldg(p.q2err + min(qg, 254));
          tIM[j] = 1.0 - eg;
          tII[j] = eg;
          const double eq = __ldg(p.q2err + min(qq, 254));
          pMa[j] = 1.0 - eq;
          pMi[j] = eq / 3.0;
          const uint32_t x = (uint8_t)p.read_bases[ro + row];
          xb[j] = x;
          rcls[j] = read_class(x);
          xeq[j] = (rcls[j] == 9) ? x : 0x100u;
          const uint32_t u = x & 0xDF;  // upper case
          xbit[j] = (u == 'A') ? 8u : (u == 'C') ? 16u : (u == 'G') ? 32u : (u == 'T') ? 64u : 0u;
          // state the row starts in: NORMAL for the first row, then the end state of the previous row
          int st = 0;
          if (p.carry_state)
            for (int r = 0; r < row; r++) st = (end_state >> (2 * st)) & 3;
          shift[j] = 2u * (uint32_t)st;
        }
      }
      // ---- state of the previous column ----
      double M[K], I[K], D[K], bM[K], bI[K], bD[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        M[j] = I[j] = bM[j] = bI[j] = bD[j] = 0.0;
        D[j] = (padrow[j] && live) ? init : 0.0;
      }
      const bool first = (t == 0);
      const bool from_carry = MULTI && first && pass > 0;
      const double* cin = carry_g + (size_t)(pass & 1) * 6 * carry_pitch;
      double* cout = carry_g + (size_t)((pass + 1) & 1) * 6 * carry_pitch;
      const bool write_carry = MULTI && live && (t == G - 1) && (pass + 1 < n_pass);
      // diagonal inputs of the lane's first row at its first column: column 0 of the row above
      double gM = 0, gI = 0, gD = (first && pass == 0) ? init : 0.0, gbM = 0, gbI = 0, gbD = 0;
      if (from_carry && live) {
        gM = cin[0]; gI = cin[carry_pitch]; gD = cin[2 * carry_pitch];
        gbM = cin[3 * carry_pitch]; gbI = cin[4 * carry_pitch]; gbD = cin[5 * carry_pitch];
      }
      if (write_carry) {
#pragma unroll
        for (int q = 0; q < 6; q++) cout[q * carry_pitch] = 0.0;
        cout[2 * carry_pitch] = D[K - 1];  // column 0 of a padding bottom row is init, of a real row 0
      }
      int c = 1 - t;
      // values of the lane above at this step's column (shuffled at the end of the previous step)
      double uM, uI, uD, ubM = 0.0, ubI = 0.0, ubD = 0.0;
      auto fetch = [&](bool with_branch) {
        uM = shfl_up_d(M[K - 1], G); uI = shfl_up_d(I[K - 1], G); uD = shfl_up_d(D[K - 1], G);
        if (with_branch) { ubM = shfl_up_d(bM[K - 1], G); ubI = shfl_up_d(bI[K - 1], G); ubD = shfl_up_d(bD[K - 1], G); }
        if (first) {
          if (!MULTI || pass == 0) { uM = uI = ubM = ubI = ubD = 0.0; uD = init; }  // row 0: D = init, everything else 0
          else if (live) {
            const int cc = min(max(c, 0), H + 1);
            uM = cin[cc]; uI = cin[carry_pitch + cc]; uD = cin[2 * carry_pitch + cc];
            ubM = cin[3 * carry_pitch + cc]; ubI = cin[4 * carry_pitch + cc]; ubD = cin[5 * carry_pitch + cc];
          }
        }
      };
      auto finish_step = [&]() {
        if (pass == n_pass - 1) sum += M[K - 1] + I[K - 1];
        if (write_carry) {
          cout[c] = M[K - 1]; cout[carry_pitch + c] = I[K - 1]; cout[2 * carry_pitch + c] = D[K - 1];
          cout[3 * carry_pitch + c] = bM[K - 1]; cout[4 * carry_pitch + c] = bI[K - 1];
          cout[5 * carry_pitch + c] = bD[K - 1];
        }
      };
      fetch(true);
      int s = 1;
      while (s <= n_steps) {
        // Lane 0 is at column s, lane G-1 at s-G+1.  Steps are "fast" while none of the columns the warp touches,
        // nor the one lane 0 reaches next, needs the state machine: the window [s-G, s+1] holds no special column.
        // (Warp-uniform only when the warp carries one pair, i.e. G = 32.)
        int n_fast = 0;
        if (G == 32 && live && s >= G) {
          const int first_special = nspec[max(s - G, -kPdMargin)];
          // fast steps also run without the per-lane activity test: every lane must be inside its haplotype
          n_fast = min(first_special - (s + 1), H - s + 1) & ~1;  // pairs of steps
        }
        n_fast = (int)__reduce_max_sync(0xffffffffu, (unsigned)max(n_fast, 0));  // identical on all lanes; tells the compiler so
        if (n_fast > 0) {
          // ---- fast steps: M/I/D only.  Two steps per iteration ping-pong between two register sets, so the
          // branch twins (which on plain columns are just the previous column's M/I/D) cost nothing: after an
          // even number of steps they are the second set. ----
          double M2[K], I2[K], D2[K];
          auto half = [&](const double (&Mi)[K], const double (&Ii)[K], const double (&Di)[K], double (&Mo)[K],
                          double (&Io)[K], double (&Do)[K]) {
            const uint32_t y = ys[c], cm = cmask[c];
            double tM = uM, tI = uI;
            double dM = gM, dI = gI, dD = gD;
#pragma unroll
            for (int j = 0; j < K; j++) {
              const bool match = ((((cm >> rcls[j]) & 1u) != 0u) | (xeq[j] == y));
              const double prior = match ? pMa[j] : pMi[j];
              const double nM = pd_match(prior, dM, dI, dD, tMM[j], tIM[j]);
              const double nD = pd_gap(Mi[j], tMD[j], Di[j], tII[j]);
              const double nI = pd_gap(tM, tMI[j], tI, tII[j]);
              dM = Mi[j]; dI = Ii[j]; dD = Di[j];
              Mo[j] = nM; Io[j] = nI; Do[j] = nD;
              tM = nM; tI = nI;
            }
            if (!MULTI || pass == n_pass - 1) sum += Mo[K - 1] + Io[K - 1];
            if (write_carry) {
              cout[c] = Mo[K - 1]; cout[carry_pitch + c] = Io[K - 1]; cout[2 * carry_pitch + c] = Do[K - 1];
              cout[3 * carry_pitch + c] = Mi[K - 1]; cout[4 * carry_pitch + c] = Ii[K - 1];
              cout[5 * carry_pitch + c] = Di[K - 1];
            }
            gM = uM; gI = uI; gD = uD;
            c++;
            uM = shfl_up_d(Mo[K - 1], G); uI = shfl_up_d(Io[K - 1], G); uD = shfl_up_d(Do[K - 1], G);
            if (first) {
              if (!MULTI || pass == 0) { uM = uI = 0.0; uD = init; }
              else {
                const int cc = min(max(c, 0), H + 1);
                uM = cin[cc]; uI = cin[carry_pitch + cc]; uD = cin[2 * carry_pitch + cc];
              }
            }
          };
          for (int k = 0; k < n_fast; k += 2) {
            half(M, I, D, M2, I2, D2);
            half(M2, I2, D2, M, I, D);
          }
#pragma unroll
          for (int j = 0; j < K; j++) { bM[j] = M2[j]; bI[j] = I2[j]; bD[j] = D2[j]; }
          s += n_fast;
          // the branch twins of the lane above were not exchanged during the fast steps; the next (slow) step
          // needs them as its top values -- its diagonal twins are only read on special columns, which by
          // construction of the window are at least one slow step away
          ubM = shfl_up_d(bM[K - 1], G); ubI = shfl_up_d(bI[K - 1], G); ubD = shfl_up_d(bD[K - 1], G);
          if (first) {
            if (!MULTI || pass == 0) { ubM = ubI = ubD = 0.0; }
            else {
              const int cc = min(max(c, 0), H + 1);
              ubM = cin[3 * carry_pitch + cc]; ubI = cin[4 * carry_pitch + cc]; ubD = cin[5 * carry_pitch + cc];
            }
          }
          gbM = gbI = gbD = 0.0;
          continue;
        }
        // ---- one slow step ----
        // Most columns that keep a window out of the fast path are INSIDE a deletion span: there only the branch
        // twins behave differently (they stay put).  The expensive merges (max of twin and main values) happen on
        // the column that closes a deletion (DEL_END: insertion inputs) and the one after it (AFTER_DEL); they are
        // executed only when some lane of the warp sits on such a column.
        const bool inrange = live && (unsigned)(c - 1) < (unsigned)H;
        const uint32_t info = inrange ? infos[c] : 0u;
        bool merge_lane = (info & 0x40u) != 0;
#pragma unroll
        for (int j = 0; j < K; j++) merge_lane |= ((info >> shift[j]) & 3u) == 2u;
        const bool merge = __any_sync(0xffffffffu, merge_lane);
        if (inrange) {
          const uint32_t y = ys[c], cm = cmask[c];
          if (!merge) {
            double tM = uM, tI = uI;
            double dM = gM, dI = gI, dD = gD;
#pragma unroll
            for (int j = 0; j < K; j++) {
              const bool inside = ((info >> shift[j]) & 3u) == 1u;
              const double lM = M[j], lI = I[j], lD = D[j];
              const bool match = ((((cm >> rcls[j]) & 1u) != 0u) | (xeq[j] == y));
              const double prior = match ? pMa[j] : pMi[j];
              const double nM = pd_match(prior, dM, dI, dD, tMM[j], tIM[j]);
              const double nD = pd_gap(lM, tMD[j], lD, tII[j]);
              const double nI = pd_gap(tM, tMI[j], tI, tII[j]);
              dM = lM; dI = lI; dD = lD;
              bM[j] = inside ? bM[j] : lM; bI[j] = inside ? bI[j] : lI; bD[j] = inside ? bD[j] : lD;
              M[j] = nM; I[j] = nI; D[j] = nD;
              tM = nM; tI = nI;
            }
          } else {
            const bool del_end = (info & 0x40u) != 0;
            double tM = uM, tI = uI, tbM = ubM, tbI = ubI;                               // top of row j
            double dM = gM, dI = gI, dD = gD, dbM = gbM, dbI = gbI, dbD = gbD;          // diagonal of row j
#pragma unroll
            for (int j = 0; j < K; j++) {
              const uint32_t st = (info >> shift[j]) & 3u;
              const bool inside = st == 1u, after = st == 2u;
              const double lM = M[j], lI = I[j], lD = D[j];
              const double lbM = bM[j], lbI = bI[j], lbD = bD[j];
              const double mxM = fmax(lbM, lM), mxI = fmax(lbI, lI), mxD = fmax(lbD, lD);
              const double nbM = after ? mxM : (inside ? lbM : lM);
              const double nbI = after ? mxI : (inside ? lbI : lI);
              const double nbD = after ? mxD : (inside ? lbD : lD);
              const double eM = after ? fmax(dM, dbM) : dM, eI = after ? fmax(dI, dbI) : dI, eD = after ? fmax(dD, dbD) : dD;
              const double leftM = after ? mxM : lM, leftD = after ? mxD : lD;
              const double topM = del_end ? fmax(tbM, tM) : tM, topI = del_end ? fmax(tbI, tI) : tI;
              const bool match = ((((cm >> rcls[j]) & 1u) != 0u) | (xeq[j] == y));
              const double prior = match ? pMa[j] : pMi[j];
              const double nM = pd_match(prior, eM, eI, eD, tMM[j], tIM[j]);
              const double nD = pd_gap(leftM, tMD[j], leftD, tII[j]);  // deletionToDeletion == insertionToInsertion
              const double nI = pd_gap(topM, tMI[j], topI, tII[j]);
              // the next row's diagonal is this row's previous column, its top this row's new column
              dM = lM; dI = lI; dD = lD; dbM = lbM; dbI = lbI; dbD = lbD;
              M[j] = nM; I[j] = nI; D[j] = nD; bM[j] = nbM; bI[j] = nbI; bD[j] = nbD;
              tM = nM; tI = nI; tbM = nbM; tbI = nbI;
            }
          }
          finish_step();
        }
        gM = uM; gI = uI; gD = uD; gbM = ubM; gbI = ubI; gbD = ubD;
        c++;
        s++;
        fetch(true);
      }
      __syncwarp();
    }
    if (mine && t == G - 1) pd_store_result(p, item, sum);
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_pdhmm2 -- the single-pass kernel (reads of at most 32*K rows), organised around the haplotype.
//   * a task is one haplotype x a block of reads (cross layout: pair r * n_haps + h; flat layout: one pair): the
//     column tables (bytes, allele bits, state codes, class masks, next-special index) are built once per task and
//     reused by every read of the block -- the reference rebuilds its per-pair prior matrix for each pair;
//   * the state a row starts in follows from the orbit of the 3-entry end-state map (pdhmm-serial.cc:306,370-385):
//     its pre-period is at most 2 and its period divides 6, so eight entries describe every row;
//   * steps whose whole window of columns is plain run without the state machine: the fill (s < 32) and drain
//     (lane 0 past the last column) phases as guarded single steps, the steady phase as ping-pong pairs;
//   * everything else (columns inside or just after a deletion span) takes the general step of k_pdhmm.
// ------------------------------------------------------------------------------------------------------------
// The reference merges with std::max on non-negative finite doubles (pdhmm-serial.cc:330-365): a compare and a select,
// without fmax's NaN handling.
__device__ __forceinline__ double dmax(double a, double b) { return (a < b) ? b : a; }
// The prior of one cell: the read byte's class bit against the column's class mask, or an identical "other" byte
// (pdhmm-serial.cc:228-277).  One predicate and one 64-bit select (the compiler nests two selects otherwise).
__device__ __forceinline__ double pick_prior(uint32_t cm, uint32_t rbit, uint32_t xeq, uint32_t y, double p_match,
                                             double p_mismatch) {
  double r;
  asm("{\n\t.reg .pred p, q;\n\t.reg .b32 t;\n\tand.b32 t, %1, %2;\n\tsetp.ne.u32 p, t, 0;\n\t"
      "setp.eq.or.u32 q, %3, %4, p;\n\tselp.f64 %0, %5, %6, q;\n\t}"
      : "=d"(r) : "r"(cm), "r"(rbit), "r"(xeq), "r"(y), "d"(p_match), "d"(p_mismatch));
  return r;
}
// (cond && a < b) ? b : a with the condition folded into the compare (one DSETP + one 64-bit select)
__device__ __forceinline__ double dmax_if(double a, double b, bool cond) {
  double r;
  asm("{\n\t.reg .pred p, c;\n\tsetp.ne.u32 c, %3, 0;\n\tsetp.lt.and.f64 p, %1, %2, c;\n\tselp.f64 %0, %2, %1, p;\n\t}"
      : "=d"(r) : "d"(a), "d"(b), "r"((unsigned)cond));
  return r;
}

template <int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_pdhmm2(const PdhmmParams p, int read_block, int n_blocks,
                                                          unsigned int n_tasks, const uint8_t* __restrict__ only) {
  constexpr int G = 32;
  constexpr int CAP = G * K;
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane;
  const int col_pitch = (p.max_hap + 2 * kPdMargin + 1) & ~1;
  uint8_t* gs = smem + (size_t)warp * 7 * col_pitch;
  uint8_t* ys = gs + kPdMargin;
  uint8_t* infos = gs + col_pitch + kPdMargin;
  uint8_t* alleles = gs + 2 * col_pitch + kPdMargin;
  uint16_t* nspec = reinterpret_cast<uint16_t*>(gs + 3 * col_pitch) + kPdMargin;
  uint16_t* cmask = reinterpret_cast<uint16_t*>(gs + 5 * col_pitch) + kPdMargin;
  const bool cross = p.n_haps > 0;
  const bool first = (t == 0);

  unsigned int next_static = blockIdx.x * WARPS + warp;   // `only`: few haplotypes of many, strided instead of queued
  for (;;) {
    unsigned int task = 0;
    if (only) {
      task = next_static;
      next_static += gridDim.x * WARPS;
    } else {
      if (lane == 0) task = atomicAdd(p.counter, 1u);
      task = __shfl_sync(0xffffffffu, task, 0);
    }
    if (task >= n_tasks) break;
    long long hi, r_begin, r_end;
    if (cross) {
      hi = task / (unsigned)n_blocks;
      if (only && !only[hi]) continue;   // the other haplotypes were taken by k_pdhmm3
      const long long blk = task - hi * (unsigned)n_blocks;
      const long long n_reads = p.n / p.n_haps;
      r_begin = blk * read_block;
      r_end = min(n_reads, r_begin + read_block);
    } else {
      hi = task;
      r_begin = task;
      r_end = r_begin + 1;
    }
    const int H = (int)p.hap_lengths[hi];
    const int8_t* hap = p.hap_bases + hi * p.max_hap;
    const int8_t* pd = p.hap_pdbases + hi * p.max_hap;

    // ---- column tables of the haplotype (once per task) ----
    const int end_state = pd_build_column_tables(t, G, 0, H, p.max_hap, hap, pd, p.carry_state, ys, infos, alleles, nspec,
                                                 cmask);
    // orbit[k] = state row k starts in for k < 8; rows >= 2 repeat with a period that divides 6
    uint32_t orbit = 0;
    if (p.carry_state) {
      int st = 0;
#pragma unroll
      for (int k = 1; k < 8; k++) {
        st = (end_state >> (2 * st)) & 3;
        orbit |= (uint32_t)st << (2 * k);
      }
    }
    const double init = p.init_cond / (double)H;
    const int n_steps = H + G - 1;

    for (long long ri = r_begin; ri < r_end; ri++) {
      const long long item = cross ? ri * p.n_haps + hi : ri;
      const int R = (int)p.read_lengths[ri];
      const int64_t ro = ri * (int64_t)p.max_read;
      const int n_pad = CAP - R;
      // ---- per-row constants ----
      double tMM[K], tIM[K], tMI[K], tII[K], tMD[K], pMa[K], pMi[K];
      uint32_t shift[K], rbit[K], xeq[K];   // rbit: the read byte's class as a bit of the column masks
      bool padrow[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int row = t * K + j - n_pad;
        padrow[j] = row < 0;
        tMM[j] = tIM[j] = tMI[j] = tMD[j] = 0.0;
        tII[j] = 1.0;
        pMa[j] = pMi[j] = 0.0;
        shift[j] = 0;
        rbit[j] = 0;       // padding rows match nothing
        xeq[j] = 0x100;
        if (!padrow[j]) {
          const int8_t iq = p.read_ins_qual[ro + row], dq = p.read_del_qual[ro + row], gq = p.gcp[ro + row];
          if (iq < 0 || dq < 0 || gq < 0) atomicOr(p.error_flag, 1u);
          const int qi = iq & 0xFF, qd = dq & 0xFF, qg = gq & 0xFF, qq = p.read_qual[ro + row] & 0xFF;
          const int mn = min(qi, qd), mx = max(qi, qd);
          tMM[j] = (mx > 254) ? 1.0 - (pow(10.0, -0.1 * mn) + pow(10.0, -0.1 * mx))
                              : __ldg(p.mm + ((mx * (mx + 1)) >> 1) + mn);
          tMI[j] = __ldg(p.q2err + min(qi, 254));
          tMD[j] = __ldg(p.q2err + min(qd, 254));
          const double eg = __ldg(p.q2err + min(qg, 254));
          tIM[j] = 1.0 - eg;
          tII[j] = eg;
          const double eq = __ldg(p.q2err + min(qq, 254));
          pMa[j] = 1.0 - eq;
          pMi[j] = eq / 3.0;
          const uint32_t x = (uint8_t)p.read_bases[ro + row];
          const uint32_t rc = read_class(x);
          rbit[j] = 1u << rc;
          xeq[j] = (rc == 9) ? x : 0x100u;
          const int k = (row < 8) ? row : 2 + ((row - 2) % 6);
          shift[j] = 2u * ((orbit >> (2 * k)) & 3u);
        }
      }
      double M[K], I[K], D[K], bM[K], bI[K], bD[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        M[j] = I[j] = bM[j] = bI[j] = bD[j] = 0.0;
        D[j] = padrow[j] ? init : 0.0;
      }
      double gM = 0, gI = 0, gD = first ? init : 0.0, gbM = 0, gbI = 0, gbD = 0;
      double sum = 0.0;
      int c = 1 - t;
      double uM, uI, uD, ubM = 0.0, ubI = 0.0, ubD = 0.0;
      // Lane 0 holds only padding rows here (the host sends reads of more than 32 K - K rows to k_pdhmm), and a
      // padding row IS row 0: M = I = 0, D = init at every column, twins likewise.  shfl_up hands lane 0 its own
      // bottom row back, which is therefore exactly the row above it -- no first-lane fix-up anywhere.
      auto fetch_main = [&]() {
        uM = shfl_up_d(M[K - 1], G); uI = shfl_up_d(I[K - 1], G); uD = shfl_up_d(D[K - 1], G);
      };
      auto fetch_twins = [&]() {
        ubM = shfl_up_d(bM[K - 1], G); ubI = shfl_up_d(bI[K - 1], G); ubD = shfl_up_d(bD[K - 1], G);
      };
      fetch_main();
      fetch_twins();
      int s = 1;
      while (s <= n_steps) {
        // lane 0 is at column s, lane 31 at s - 31.  A run of steps needs no state machine while the window
        // [s - 32, s + 1] holds no special column (the +1 keeps one general step between a run and the next
        // special column, which re-establishes the diagonal twins).
        const int first_special = nspec[max(s - G, -kPdMargin)];
        int n_run = min(first_special - (s + 1), n_steps - s + 1);
        n_run = (int)__reduce_max_sync(0xffffffffu, (unsigned)max(n_run, 0));  // identical on all lanes
        if (n_run > 0) {
          const int s_end = s + n_run;
          // one guarded plain step, in place.  The twins of a plain column are the previous column's values; nothing
          // reads them during a run, so only the last step of a run stores them.
          auto single = [&](auto save_twins) {
            if ((unsigned)(c - 1) < (unsigned)H) {
              const uint32_t y = ys[c], cm = cmask[c];
              double tM = uM, tI = uI;
              double dM = gM, dI = gI, dD = gD;
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double lM = M[j], lI = I[j], lD = D[j];
                const double prior = pick_prior(cm, rbit[j], xeq[j], y, pMa[j], pMi[j]);
                const double nM = pd_match(prior, dM, dI, dD, tMM[j], tIM[j]);
                const double nD = pd_gap(lM, tMD[j], lD, tII[j]);
                const double nI = pd_gap(tM, tMI[j], tI, tII[j]);
                dM = lM; dI = lI; dD = lD;
                if constexpr (decltype(save_twins)::value) { bM[j] = lM; bI[j] = lI; bD[j] = lD; }
                M[j] = nM; I[j] = nI; D[j] = nD;
                tM = nM; tI = nI;
              }
              sum += M[K - 1] + I[K - 1];
            }
            gM = uM; gI = uI; gD = uD;
            c++;
            s++;
            fetch_main();
          };
          while (s < s_end && s < G) {                                // fill: lanes enter one by one
            if (s + 1 == s_end) single(std::true_type{});
            else single(std::false_type{});
          }
          // steady and drain: every lane has reached column 1; lanes past the last column compute values nobody
          // reads (their sums are masked, the column tables have margins)
          const int n_fast = max(0, s_end - s) & ~1;
          if (n_fast > 0) {
            double M2[K], I2[K], D2[K];
            auto half = [&](const double (&Mi)[K], const double (&Ii)[K], const double (&Di)[K], double (&Mo)[K],
                            double (&Io)[K], double (&Do)[K]) {
              const uint32_t y = ys[c], cm = cmask[c];
              double tM = uM, tI = uI;
              double dM = gM, dI = gI, dD = gD;
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double prior = pick_prior(cm, rbit[j], xeq[j], y, pMa[j], pMi[j]);
                const double nM = pd_match(prior, dM, dI, dD, tMM[j], tIM[j]);
                const double nD = pd_gap(Mi[j], tMD[j], Di[j], tII[j]);
                const double nI = pd_gap(tM, tMI[j], tI, tII[j]);
                dM = Mi[j]; dI = Ii[j]; dD = Di[j];
                Mo[j] = nM; Io[j] = nI; Do[j] = nD;
                tM = nM; tI = nI;
              }
              const double add = Mo[K - 1] + Io[K - 1];
              sum += (c <= H) ? add : 0.0;
              gM = uM; gI = uI; gD = uD;
              c++;
              uM = shfl_up_d(Mo[K - 1], G); uI = shfl_up_d(Io[K - 1], G); uD = shfl_up_d(Do[K - 1], G);
            };
            for (int k = 0; k < n_fast; k += 2) {
              half(M, I, D, M2, I2, D2);
              half(M2, I2, D2, M, I, D);
            }
#pragma unroll
            for (int j = 0; j < K; j++) { bM[j] = M2[j]; bI[j] = I2[j]; bD[j] = D2[j]; }
            s += n_fast;
          }
          while (s < s_end) single(std::true_type{});                            // the odd step of a run
          // the twins of the lane above were not exchanged during the run; the next (general) step needs them
          // as its top values -- its diagonal twins are only read on special columns, which by construction
          // of the window are at least one general step away
          fetch_twins();
          gbM = gbI = gbD = 0.0;
          continue;
        }
        const bool inrange = (unsigned)(c - 1) < (unsigned)H;
        const uint32_t info = inrange ? infos[c] : 0u;
        if (orbit == 0) {
          // ---- one special-window step when every row starts NORMAL (always, unless the haplotype ends inside a
          // deletion): the column state is the same for all rows, so the state machine reduces to the plain update
          // plus three per-lane corrections.  AFTER_DEL: the left and diagonal inputs are first merged with their
          // twins (pdhmm-serial.cc:330-352), after which the column behaves like a NORMAL one.  INSIDE_DEL: the
          // twins stay frozen.  DEL_END: the insertion chain is redone with max(twin, value) tops (:362-365).  The
          // twins are only stored where a later column reads them (bit 7 of the column code). ----
          const uint32_t st = info & 3u;
          const bool after = st == 2u, del_end = (info & 0x40u) != 0, capture = (info & 0x80u) != 0;
          if (__any_sync(0xffffffffu, after)) {
#pragma unroll
            for (int j = 0; j < K; j++) {
              M[j] = dmax_if(M[j], bM[j], after);
              I[j] = dmax_if(I[j], bI[j], after);
              D[j] = dmax_if(D[j], bD[j], after);
            }
            gM = dmax_if(gM, gbM, after);
            gI = dmax_if(gI, gbI, after);
            gD = dmax_if(gD, gbD, after);
          }
          // the plain update; once every lane has started (s >= 32) nothing has to be protected: lanes past the last
          // column compute values nobody reads
          auto update = [&](auto guarded) {
            if (!decltype(guarded)::value || inrange) {
              const uint32_t y = ys[c], cm = cmask[c];
              double tM = uM, tI = uI;
              double dM = gM, dI = gI, dD = gD;
#pragma unroll
              for (int j = 0; j < K; j++) {
                const double lM = M[j], lI = I[j], lD = D[j];
                const double prior = pick_prior(cm, rbit[j], xeq[j], y, pMa[j], pMi[j]);
                const double nM = pd_match(prior, dM, dI, dD, tMM[j], tIM[j]);
                const double nD = pd_gap(lM, tMD[j], lD, tII[j]);
                const double nI = pd_gap(tM, tMI[j], tI, tII[j]);
                dM = lM; dI = lI; dD = lD;
                if (capture) { bM[j] = lM; bI[j] = lI; bD[j] = lD; }
                M[j] = nM; I[j] = nI; D[j] = nD;
                tM = nM; tI = nI;
              }
            }
          };
          if (s >= G) update(std::false_type{});
          else update(std::true_type{});
          if (__any_sync(0xffffffffu, del_end)) {
            double tM = uM, tI = uI, tbM = ubM, tbI = ubI;
#pragma unroll
            for (int j = 0; j < K; j++) {
              const double nI = pd_gap(dmax(tbM, tM), tMI[j], dmax(tbI, tI), tII[j]);
              I[j] = del_end ? nI : I[j];
              tM = M[j]; tI = I[j]; tbM = bM[j]; tbI = bI[j];
            }
          }
          {
            const double add = M[K - 1] + I[K - 1];
            sum += inrange ? add : 0.0;
          }
          gM = uM; gI = uI; gD = uD; gbM = ubM; gbI = ubI; gbD = ubD;
          c++;
          s++;
          fetch_main();
          fetch_twins();
          continue;
        }
        // ---- one general step (as in k_pdhmm): rows of one lane may sit in different states ----
        bool merge_lane = (info & 0x40u) != 0;
#pragma unroll
        for (int j = 0; j < K; j++) merge_lane |= ((info >> shift[j]) & 3u) == 2u;
        const bool merge = __any_sync(0xffffffffu, merge_lane);
        if (inrange) {
          const uint32_t y = ys[c], cm = cmask[c];
          if (!merge) {
            double tM = uM, tI = uI;
            double dM = gM, dI = gI, dD = gD;
#pragma unroll
            for (int j = 0; j < K; j++) {
              const bool inside = ((info >> shift[j]) & 3u) == 1u;
              const double lM = M[j], lI = I[j], lD = D[j];
              const double prior = pick_prior(cm, rbit[j], xeq[j], y, pMa[j], pMi[j]);
              const double nM = pd_match(prior, dM, dI, dD, tMM[j], tIM[j]);
              const double nD = pd_gap(lM, tMD[j], lD, tII[j]);
              const double nI = pd_gap(tM, tMI[j], tI, tII[j]);
              dM = lM; dI = lI; dD = lD;
              bM[j] = inside ? bM[j] : lM; bI[j] = inside ? bI[j] : lI; bD[j] = inside ? bD[j] : lD;
              M[j] = nM; I[j] = nI; D[j] = nD;
              tM = nM; tI = nI;
            }
          } else {
            const bool del_end = (info & 0x40u) != 0;
            double tM = uM, tI = uI, tbM = ubM, tbI = ubI;
            double dM = gM, dI = gI, dD = gD, dbM = gbM, dbI = gbI, dbD = gbD;
#pragma unroll
            for (int j = 0; j < K; j++) {
              const uint32_t st = (info >> shift[j]) & 3u;
              const bool inside = st == 1u, after = st == 2u;
              const double lM = M[j], lI = I[j], lD = D[j];
              const double lbM = bM[j], lbI = bI[j], lbD = bD[j];
              const double mxM = dmax(lbM, lM), mxI = dmax(lbI, lI), mxD = dmax(lbD, lD);
              const double nbM = after ? mxM : (inside ? lbM : lM);
              const double nbI = after ? mxI : (inside ? lbI : lI);
              const double nbD = after ? mxD : (inside ? lbD : lD);
              const double eM = after ? dmax(dM, dbM) : dM, eI = after ? dmax(dI, dbI) : dI, eD = after ? dmax(dD, dbD) : dD;
              const double leftM = after ? mxM : lM, leftD = after ? mxD : lD;
              const double topM = del_end ? dmax(tbM, tM) : tM, topI = del_end ? dmax(tbI, tI) : tI;
              const double prior = pick_prior(cm, rbit[j], xeq[j], y, pMa[j], pMi[j]);
              const double nM = pd_match(prior, eM, eI, eD, tMM[j], tIM[j]);
              const double nD = pd_gap(leftM, tMD[j], leftD, tII[j]);
              const double nI = pd_gap(topM, tMI[j], topI, tII[j]);
              dM = lM; dI = lI; dD = lD; dbM = lbM; dbI = lbI; dbD = lbD;
              M[j] = nM; I[j] = nI; D[j] = nD; bM[j] = nbM; bI[j] = nbI; bD[j] = nbD;
              tM = nM; tI = nI; tbM = nbM; tbI = nbI;
            }
          }
          sum += M[K - 1] + I[K - 1];
        }
        gM = uM; gI = uI; gD = uD; gbM = ubM; gbI = ubI; gbD = ubD;
        c++;
        s++;
        fetch_main();
        fetch_twins();
      }
      if (t == G - 1) pd_store_result(p, item, sum);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// k_pdhmm3 -- two reads per warp (16 lanes x K rows each) against one haplotype.
//   The recurrence of k_pdhmm2 with the deletion state folded (pd_match_y; results agree to ~1e-13); the shape differs:
//   * 101-row reads fill 101 of 16 x 7 = 112 rows (k_pdhmm2: 128), the fill/drain is 15 steps, a special column keeps
//     the warp in the state-machine steps for 16 + 2 steps instead of 32 + 2, and the per-step overhead (shuffles, table
//     loads, loop) is spread over 7 cells per lane instead of 4;
//   * both reads of a warp meet the same haplotype, so the column tables, the step count and the special windows are
//     shared and every branch of the step loop stays warp-uniform;
//   * the branch twins live in shared memory ([3 K][32 lanes] doubles per warp): they are written on the columns that
//     capture them and read on the two columns that merge them, nowhere else -- the registers they occupied in k_pdhmm2
//     pay for K = 7;
//   * haplotypes whose rows do not all start in the NORMAL state (they end inside or right after a deletion and the
//     state is carried to the next row, pdhmm-serial.cc:306) are left to k_pdhmm2: `deferred[h]` is set by the host.
// ------------------------------------------------------------------------------------------------------------
// Columns of one haplotype grouped by what decides the prior: the class mask and, for a byte outside ACGTacgtN, the
// byte itself (it matches an identical read byte).  keys[i] = mask | (0x100 | byte) << 16; colid[c] = index of column
// c's key, max_ids outside the haplotype.  Returns the number of distinct keys; columns beyond `max_ids` keys get the
// last index (the caller gives such a haplotype to k_pdhmm2).  Warp-cooperative: 32 columns at a time, new keys appended in column order.
__device__ __forceinline__ bool pd_other_byte(uint32_t y) {
  const uint32_t u = y & 0xDFu;
  return !(u == 'A' || u == 'C' || u == 'G' || u == 'T' || y == 'N');
}
__device__ __forceinline__ int pd_assign_column_ids(int lane, int H, int max_hap, const uint8_t* ys, const uint16_t* cmask,
                                                    uint8_t* colid, uint32_t* keys, int max_ids) {
  for (int c = lane - kPdMargin; c < max_hap + kPdMargin; c += 32) colid[c] = (uint8_t)max_ids;   // "no column"
  __syncwarp();
  int n = 0;
  for (int base = 1; base <= H; base += 32) {
    const int c = base + lane;
    const bool valid = c <= H;
    uint32_t key = 0xFFFFFFFFu;
    if (valid) {
      const uint32_t y = ys[c];
      key = (uint32_t)cmask[c] | (pd_other_byte(y) ? (0x100u | y) << 16 : 0u);
    }
    int id = -1;
    for (int i = 0; i < min(n, max_ids); i++)
      if (key == keys[i]) id = i;
    __syncwarp();   // every lane has compared against the keys found so far
    unsigned pending = __ballot_sync(0xffffffffu, valid && id < 0);
    while (pending) {
      const uint32_t kk = __shfl_sync(0xffffffffu, key, __ffs(pending) - 1);
      if (lane == 0 && n < max_ids) keys[n] = kk;
      if (key == kk) id = n;
      n++;
      pending = __ballot_sync(0xffffffffu, valid && id < 0);
    }
    __syncwarp();
    if (valid) colid[c] = (uint8_t)min(id, max_ids - 1);
  }
  __syncwarp();
  return n;
}

// k_pdhmm3 keeps the deletion state divided by its row's match-to-deletion probability (Y = D / tMD): the update
// becomes one fused multiply-add, Y' = M + Y * tDD, and the match state reads it back through bD = tMD(row above) * tIM.
// Y never exceeds the sum of its row's match values, which the total probability mass bounds by the initial condition.
//
// WFOLD: the insertion state is folded the same way, W = I / tMI(row): W' = M(up) + W(up) * kap with
// kap = tII * tMI(row above) / tMI(row), and the match state reads it through bI = tMI(row above) * tIM -- six fp64
// operations per cell instead of seven.  W's scale crosses rows: W(j) <= (max tMI / min tMI over the read) * sum of the
// match values above, so the host only selects this variant when every read's insertion qualities span at most 30 dB
// (a factor 2^10), and the kernel then runs with the initial condition lowered by kPdWFoldShift bits, which keeps W
// below 2^1013 for any haplotype and read length; the result subtracts the same shift.  Values that the reference
// still holds as fp64 denormals would be lost 24 bits earlier: the host also requires the likelihood's lower bound
// (one match, one gap open, gap extensions) to stay far above that range (pdhmm_engine.cu: wfold_ok).
constexpr int kPdWFoldShift = 24;
__device__ __forceinline__ double pd_match_y(double prior, double dM, double dI, double dY, double tMM, double cI,
                                             double bD) {   // cI: tIM, or bI when the insertion state is folded
  return __dmul_rn(prior, fma(dY, bD, fma(dI, cI, __dmul_rn(dM, tMM))));
}
// the insertion update: I' = M(up) * tMI + I(up) * tII, or folded W' = M(up) + W(up) * kap
template <bool WFOLD>
__device__ __forceinline__ double pd_ins(double tM, double tI, double tMI, double c) {   // c: tII, or kap
  if constexpr (WFOLD) return fma(tI, c, tM);
  else return pd_gap(tM, tMI, tI, c);
}

template <int G, int K, int WARPS, int NID, bool WFOLD>
__global__ void __launch_bounds__(WARPS * 32, 1) k_pdhmm3(const PdhmmParams p, int read_block, int n_blocks,
                                                          unsigned int n_tasks, uint8_t* deferred) {
  static_assert(G == 16 || G == 32, "one or two reads per warp");
  constexpr int GPW = 32 / G;
  constexpr int CAP = G * K;
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = lane & (G - 1), g = lane / G;
  const int col_pitch = (p.max_hap + 2 * kPdMargin + 1) & ~1;
  const size_t table_bytes = ((size_t)7 * col_pitch + 15) & ~(size_t)15;
  constexpr size_t kDoubles = (size_t)(3 + NID) * K * 32;
  uint8_t* gs = smem + (size_t)warp * (table_bytes + kDoubles * sizeof(double) + 64);
  uint8_t* ys = gs + kPdMargin;
  uint8_t* infos = gs + col_pitch + kPdMargin;
  uint8_t* alleles = gs + 2 * col_pitch + kPdMargin;
  uint16_t* nspec = reinterpret_cast<uint16_t*>(gs + 3 * col_pitch) + kPdMargin;
  uint16_t* cmask = reinterpret_cast<uint16_t*>(gs + 5 * col_pitch) + kPdMargin;
  uint8_t* colid = alleles;   // which prior-table block column c reads (the allele bits are only used to build cmask)
  double* tw = reinterpret_cast<double*>(gs + table_bytes);   // twins: tw[(3 j + q) * 32 + lane]
  double* tw_me = tw + lane;
  // priors of this lane's rows for every kind of column: tab[(id * K + j) * 32 + lane]
  double* tab_me = tw + 3 * K * 32 + lane;
  uint32_t* keys = reinterpret_cast<uint32_t*>(gs + table_bytes + kDoubles * sizeof(double));
  const double* tw_up = tw + 3 * (K - 1) * 32 + (t == 0 ? lane : lane - 1);  // bottom-row twins of the lane above
  const long long n_reads = p.n / p.n_haps;
#pragma unroll
  for (int j = 0; j < K; j++) tab_me[((NID - 1) * K + j) * 32] = 0.0;

  for (;;) {
    unsigned int task = 0;
    if (lane == 0) task = atomicAdd(p.counter, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= n_tasks) break;
    const long long hi = task / (unsigned)n_blocks;
    if (deferred[hi]) continue;
    const long long blk = task - hi * (unsigned)n_blocks;
    const long long r_begin = blk * read_block;
    const long long r_end = min(n_reads, r_begin + read_block);
    const int H = (int)p.hap_lengths[hi];
    const int8_t* hap = p.hap_bases + hi * p.max_hap;
    const int8_t* pd = p.hap_pdbases + hi * p.max_hap;
    pd_build_column_tables(lane, 32, 0, H, p.max_hap, hap, pd, p.carry_state, ys, infos, alleles, nspec, cmask);
    // kind NID - 1 is "no column" (prior 0): the margins on both sides of the haplotype
    const int n_ids = pd_assign_column_ids(lane, H, p.max_hap, ys, cmask, colid, keys, NID - 1);
    if (n_ids > NID - 1) {   // more kinds of columns than the prior table holds: k_pdhmm2 takes this haplotype
      if (lane == 0) deferred[hi] = 1;
      continue;
    }
    const double init = (WFOLD ? scalbn(p.init_cond, -kPdWFoldShift) : p.init_cond) / (double)H;
    const int n_steps = H + G - 1;

    for (long long r0 = r_begin; r0 < r_end; r0 += GPW) {
      const bool mine = r0 + g < r_end;
      const long long ri = mine ? r0 + g : r0;     // an odd last read is computed twice, stored once
      const long long item = ri * p.n_haps + hi;
      const int R = (int)p.read_lengths[ri];
      const int64_t ro = ri * (int64_t)p.max_read;
      const int n_pad = CAP - R;
      // ---- per-row constants ----
      double tMM[K], tIM[K], tMI[K], tII[K], tMD[K], pMa[K], pMi[K];   // tMD: 1 on padding rows (Y = D there)
      uint32_t rbit[K], xeq[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int row = t * K + j - n_pad;
        tMM[j] = tIM[j] = 0.0;
        tMI[j] = WFOLD ? 1.0 : 0.0;   // folded: the scale of a padding row's (zero) insertion state
        tII[j] = tMD[j] = 1.0;
        pMa[j] = pMi[j] = 0.0;
        rbit[j] = 0;
        xeq[j] = 0x200;
        if (row >= 0) {
          const int8_t iq = p.read_ins_qual[ro + row], dq = p.read_del_qual[ro + row], gq = p.gcp[ro + row];
          if (iq < 0 || dq < 0 || gq < 0) atomicOr(p.error_flag, 1u);
          const int qi = iq & 0xFF, qd = dq & 0xFF, qg = gq & 0xFF, qq = p.read_qual[ro + row] & 0xFF;
          const int mn = min(qi, qd), mx = max(qi, qd);
          tMM[j] = (mx > 254) ? 1.0 - (pow(10.0, -0.1 * mn) + pow(10.0, -0.1 * mx))
                              : __ldg(p.mm + ((mx * (mx + 1)) >> 1) + mn);
          tMI[j] = __ldg(p.q2err + min(qi, 254));
          tMD[j] = __ldg(p.q2err + min(qd, 254));
          const double eg = __ldg(p.q2err + min(qg, 254));
          tIM[j] = 1.0 - eg;
          tII[j] = eg;
          const double eq = __ldg(p.q2err + min(qq, 254));
          pMa[j] = 1.0 - eq;
          pMi[j] = eq / 3.0;
          const uint32_t x = (uint8_t)p.read_bases[ro + row];
          const uint32_t rc = read_class(x);
          rbit[j] = 1u << rc;
          xeq[j] = (rc == 9) ? (0x100u | x) : 0x200u;
        }
      }
      double bD[K];   // what the match state multiplies the Y of the row above with
      {
        double above = shfl_up_d(tMD[K - 1], G);   // lane 0 of a half: its own padding row, 1
#pragma unroll
        for (int j = 0; j < K; j++) {
          bD[j] = __dmul_rn(above, tIM[j]);
          above = tMD[j];
        }
      }
      // WFOLD: cI = bI (what the match state multiplies the W of the row above with), cW = kap; else tIM and tII
      double cI[K], cW[K];
      const double tMI_last = tMI[K - 1];
      if constexpr (WFOLD) {
        double above = shfl_up_d(tMI[K - 1], G);
#pragma unroll
        for (int j = 0; j < K; j++) {
          cI[j] = __dmul_rn(above, tIM[j]);
          cW[j] = __ddiv_rn(__dmul_rn(tII[j], above), tMI[j]);
          above = tMI[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < K; j++) { cI[j] = tIM[j]; cW[j] = tII[j]; }
      }
      double M[K], I[K], D[K];   // D holds Y
#pragma unroll
      for (int j = 0; j < K; j++) {
        M[j] = I[j] = 0.0;
        D[j] = (t * K + j < n_pad) ? init : 0.0;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 3 * K; q++) tw_me[q * 32] = 0.0;
      // the prior of row j on a column of kind id (pdhmm-serial.cc:228-277), once per pair instead of once per cell
      for (int id = 0; id < n_ids; id++) {
        const uint32_t key = keys[id];
#pragma unroll
        for (int j = 0; j < K; j++)
          tab_me[(id * K + j) * 32] = (((key & rbit[j]) != 0u) | ((key >> 16) == xeq[j])) ? pMa[j] : pMi[j];
      }
      __syncwarp();
      double gM = 0, gI = 0, gD = (t == 0) ? init : 0.0, gbM = 0, gbI = 0, gbD = 0;
      double sum = 0.0;
      int c = 1 - t;
      double uM, uI, uD, ubM = 0.0, ubI = 0.0, ubD = 0.0;
      // lane 0 of each half holds only padding rows (reads of more than G K - K rows go to the other kernels), and a
      // padding row IS row 0, so the half-wide shuffle hands it exactly the row above it.
      // Lanes that have not reached column 1 yet run the same update: the margin columns are of the kind with prior 0,
      // which keeps a lane in the state of column 0 (M = I = 0, Y unchanged: 0 on real rows, init on padding rows).
      // Lanes past the last column compute values nobody reads; only their sums are masked.
      double M2[K], I2[K], D2[K];
      auto fetch_twins = [&]() {
        __syncwarp();
        ubM = tw_up[0]; ubI = tw_up[32]; ubD = tw_up[64];
      };
      // one plain column, from one register set into the other (`masked`: some lane may be past the last column)
      auto half = [&](auto masked, const double (&Mi)[K], const double (&Ii)[K], const double (&Di)[K],
                      double (&Mo)[K], double (&Io)[K], double (&Do)[K]) {
        const double* pt = tab_me + (int)colid[c] * (K * 32);
        double tM = uM, tI = uI;
        double dM = gM, dI = gI, dD = gD;
#pragma unroll
        for (int j = 0; j < K; j++) {
          const double prior = pt[j * 32];
          const double nM = pd_match_y(prior, dM, dI, dD, tMM[j], cI[j], bD[j]);
          const double nD = fma(Di[j], tII[j], Mi[j]);
          const double nI = pd_ins<WFOLD>(tM, tI, tMI[j], cW[j]);
          dM = Mi[j]; dI = Ii[j]; dD = Di[j];
          Mo[j] = nM; Io[j] = nI; Do[j] = nD;
          tM = nM; tI = nI;
        }
        const double add = WFOLD ? fma(Io[K - 1], tMI_last, Mo[K - 1]) : Mo[K - 1] + Io[K - 1];
        if constexpr (decltype(masked)::value) sum += (c <= H) ? add : 0.0;
        else sum += add;
        gM = uM; gI = uI; gD = uD;
        c++;
        uM = shfl_up_d(Mo[K - 1], G); uI = shfl_up_d(Io[K - 1], G); uD = shfl_up_d(Do[K - 1], G);
      };
      // one column of a special window (every row starts NORMAL, see k_pdhmm2): AFTER_DEL merges the left and diagonal
      // inputs with their twins, the columns flagged 0x80 capture the twins, DEL_END redoes the insertion chain
      auto window = [&](double (&Mi)[K], double (&Ii)[K], double (&Di)[K], double (&Mo)[K], double (&Io)[K],
                        double (&Do)[K]) {
        const bool inrange = (unsigned)(c - 1) < (unsigned)H;
        const uint32_t info = inrange ? infos[c] : 0u;
        const bool after = (info & 3u) == 2u, del_end = (info & 0x40u) != 0, capture = (info & 0x80u) != 0;
        if (__any_sync(0xffffffffu, after)) {
#pragma unroll
          for (int j = 0; j < K; j++) {
            Mi[j] = dmax_if(Mi[j], tw_me[(3 * j) * 32], after);
            Ii[j] = dmax_if(Ii[j], tw_me[(3 * j + 1) * 32], after);
            Di[j] = dmax_if(Di[j], tw_me[(3 * j + 2) * 32], after);
          }
          gM = dmax_if(gM, gbM, after);
          gI = dmax_if(gI, gbI, after);
          gD = dmax_if(gD, gbD, after);
        }
        if (__any_sync(0xffffffffu, capture)) {
          __syncwarp();   // the lane below has read last step's twins
          if (capture) {
#pragma unroll
            for (int j = 0; j < K; j++) {
              tw_me[(3 * j) * 32] = Mi[j];
              tw_me[(3 * j + 1) * 32] = Ii[j];
              tw_me[(3 * j + 2) * 32] = Di[j];
            }
          }
        }
        {
          const double* pt = tab_me + (int)colid[c] * (K * 32);
          double tM = uM, tI = uI;
          double dM = gM, dI = gI, dD = gD;
#pragma unroll
          for (int j = 0; j < K; j++) {
            const double prior = pt[j * 32];
            const double nM = pd_match_y(prior, dM, dI, dD, tMM[j], cI[j], bD[j]);
            const double nD = fma(Di[j], tII[j], Mi[j]);
            const double nI = pd_ins<WFOLD>(tM, tI, tMI[j], cW[j]);
            dM = Mi[j]; dI = Ii[j]; dD = Di[j];
            Mo[j] = nM; Io[j] = nI; Do[j] = nD;
            tM = nM; tI = nI;
          }
        }
        if (__any_sync(0xffffffffu, del_end)) {
          double tM = uM, tI = uI, tbM = ubM, tbI = ubI;
#pragma unroll
          for (int j = 0; j < K; j++) {
            const double nI = pd_ins<WFOLD>(dmax(tbM, tM), dmax(tbI, tI), tMI[j], cW[j]);
            Io[j] = del_end ? nI : Io[j];
            tM = Mo[j]; tI = Io[j];
            tbM = tw_me[(3 * j) * 32]; tbI = tw_me[(3 * j + 1) * 32];
          }
        }
        {
          const double add = WFOLD ? fma(Io[K - 1], tMI_last, Mo[K - 1]) : Mo[K - 1] + Io[K - 1];
          sum += inrange ? add : 0.0;
        }
        gM = uM; gI = uI; gD = uD; gbM = ubM; gbI = ubI; gbD = ubD;
        c++;
        uM = shfl_up_d(Mo[K - 1], G); uI = shfl_up_d(Io[K - 1], G); uD = shfl_up_d(Do[K - 1], G);
        fetch_twins();
      };
      uM = shfl_up_d(M[K - 1], G); uI = shfl_up_d(I[K - 1], G); uD = shfl_up_d(D[K - 1], G);
      int s = 1;
      while (s <= n_steps) {
        // lane 0 is at column s, lane G-1 at s - G + 1.  A run of columns needs no state machine while the window
        // [s - G, s + 1] holds no special column (the +1 keeps one window step between a run and the next special
        // column, which re-establishes the diagonal twins).  Runs and windows both advance two columns at a time,
        // ping-ponging between the register sets; a window step is valid on any column.
        const int first_special = nspec[max(s - G, -kPdMargin)];
        int n_run = min(first_special - (s + 1), n_steps - s + 1);
        n_run = (int)__reduce_max_sync(0xffffffffu, (unsigned)max(n_run, 0)) & ~1;   // identical on all lanes
        if (n_run > 0) {
          // while lane 0 has not passed the last column every lane sits on a real column or on the left margin
          const int n_steady = max(0, min(n_run, H + 1 - s)) & ~1;
          int k = 0;
          for (; k + 8 <= n_steady; k += 8) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
              half(std::false_type{}, M, I, D, M2, I2, D2);
              half(std::false_type{}, M2, I2, D2, M, I, D);
            }
          }
          for (; k < n_steady; k += 2) {
            half(std::false_type{}, M, I, D, M2, I2, D2);
            half(std::false_type{}, M2, I2, D2, M, I, D);
          }
          for (; k < n_run; k += 2) {
            half(std::true_type{}, M, I, D, M2, I2, D2);
            half(std::true_type{}, M2, I2, D2, M, I, D);
          }
          s += n_run;
          // no twin is read or written during a run; the next window step takes its top twins from shared memory,
          // its diagonal twins are only read on special columns, at least one window step away
          fetch_twins();
          gbM = gbI = gbD = 0.0;
          continue;
        }
        window(M, I, D, M2, I2, D2);
        s++;
        if (s <= n_steps) {
          window(M2, I2, D2, M, I, D);
          s++;
        }
      }
      if (t == G - 1 && mine) pd_store_result(p, item, sum, WFOLD ? kPdWFoldShift : 0);
    }
  }
}


}  // namespace gklb
