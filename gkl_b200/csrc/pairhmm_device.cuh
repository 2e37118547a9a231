// pairhmm_device.cuh -- sm_100a device code of the PairHMM engine.
//
// What is computed (reference semantics, /root/reference/src/main/native/pairhmm):
//   the three-state forward recurrence of computeMXY (avx-pairhmm-template.h:208-223) in its
//   cell form
//       M[r][c] = prior(r,c) * (M[r-1][c-1]*pMM_r + (X[r-1][c-1] + Y[r-1][c-1])*pGAPM_r)
//       X[r][c] = M[r-1][c]*pMX_r + X[r-1][c]*pXX_r
//       Y[r][c] = M[r][c-1]*pMY_r + Y[r][c-1]*pYY_r
//   with row 0 = {M=0, X=0, Y=INITIAL_CONSTANT/haplen} (:114-121,160-202), column 0 = 0 for
//   rows >= 1, result = sum_c (M[R][c] + X[R][c]) (:325-371), per-row probabilities from
//   quals & 127 (:106-152) and the fp32 -> fp64 rerun / log10 of IntelPairHmm.cc:150-169.
//
// How it is mapped to the machine (nothing like the reference's striped AVX code):
//   * a GROUP of G lanes (G = 8, 16 or 32; 32/G groups per warp) sweeps one read (two reads
//     when packed) against one haplotype as a systolic array: lane t owns K consecutive read
//     rows, holds their per-row constants and the state of the previous column in registers,
//     and at step s processes column c = s - t, top row to bottom row;
//   * the bottom row of lane t-1 reaches lane t with three __shfl_up_sync per step; the
//     diagonal inputs are simply the previous step's shuffle results;
//   * packed fp32 (VF2): each lane carries TWO reads in the halves of 64-bit register pairs and
//     all arithmetic is FFMA2/FMUL2/FADD2 (sm_100 fma.rn.ftz.f32x2) -- half the issue slots of
//     scalar FFMA;
//   * the kernel is bound by register-file operand bandwidth, so the recurrence is rewritten to
//     read fewer operands per cell than the reference's 8 mul + 4 add (template parameter VAR,
//     see LaneRows): Y is kept as Y/pMY and the diagonal state as pGAPM(next row)*(X+Y) (VAR 2);
//     the prior of every (row, haplotype symbol) comes from a per-warp shared-memory table with
//     one LDS per cell instead of a LOP3 + FSEL (VAR 3); X is kept as X/pMX(row) (VAR 4/5, fp32
//     only, with a non-finite-sum guard that falls back to the fp64 kernel).  VAR 0/1 are the
//     plain forms, kept for measurement;
//   * reads shorter than G*K rows are padded at the TOP with rows that reproduce row 0
//     (A=G=pMX=pMY=0, pXX=pYY=1, Y(c=0)=init): the last read row is then always the last row
//     of the last lane, so the running sum lives in fixed registers;
//   * reads longer than G*K rows take several passes; the bottom row of a pass is carried to
//     the next pass through a per-warp global scratch line (the analogue of the reference's
//     shiftOut arrays, avx-pairhmm-template.h:249,313-317);
//   * the haplotype panel lives in shared memory as one image fetched with a single TMA bulk
//     copy (cp.async.bulk + mbarrier) per CTA; the packed read records of each task are
//     fetched the same way into a per-warp slot;
//   * persistent CTAs (one per SM) pull tasks (a block of reads x a chunk of haplotypes) from
//     an atomic counter -- the analogue of the reference's schedule(dynamic,1); batches with
//     several length classes are served by one multi-class launch (k_mega_*).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gklb {

constexpr int kHapLeftMargin = 32;   // bytes of zero symbols before column 1 of each haplotype
constexpr int kHapRightMargin = 48;  // and after its last column (covers c = s - t overshoot + prefetch)
constexpr int kMmTableSize = (128 * 129) / 2;  // reachable part of matchToMatchProb (quals & 127)
constexpr float kMinAccepted = 1e-28f;         // pairhmm_common.h:39

// One-hot base nibbles.  ConvertChar (pairhmm_common.h:49-67): A C T G N, everything else == 'A'.
// N matches everything (precompute_masks, avx-pairhmm-template.h:26-58).
__host__ __device__ inline uint8_t base_nibble(uint8_t ch) {
  switch (ch) {
    case 'C': return 2;
    case 'T': return 4;
    case 'G': return 8;
    case 'N': return 15;
    default: return 1;  // 'A' and every other byte
  }
}

// Symbol index of a haplotype base (ConvertChar order: A C T G N) and its byte in the panel image:
// low 3 bits = index (VAR 3 prior-table lookup), high nibble = one-hot code (VAR 0-2 match test).
__host__ __device__ inline uint8_t base_index(uint8_t ch) {
  switch (ch) {
    case 'C': return 1;
    case 'T': return 2;
    case 'G': return 3;
    case 'N': return 4;
    default: return 0;
  }
}
__host__ __device__ inline uint8_t panel_byte(uint8_t ch) { return (uint8_t)(base_index(ch) | (base_nibble(ch) << 4)); }
constexpr int kPriorSyms = 5;

// ------------------------------------------------------------------------------------------
// Lane arithmetic policies.  V is what one lane holds per matrix cell slot (64 bits for the
// two product policies), S the scalar type, NR the number of reads carried per lane.
// ------------------------------------------------------------------------------------------
struct VF2 {  // two reads per lane, packed fp32
  typedef float2 V;
  typedef float S;
  static constexpr int NR = 2;
  static constexpr bool kDouble = false;
  __device__ static __forceinline__ V mul(V a, V b) { return __fmul2_rn(a, b); }
  __device__ static __forceinline__ V fma(V a, V b, V c) { return __ffma2_rn(a, b, c); }
  __device__ static __forceinline__ V add(V a, V b) { return __fadd2_rn(a, b); }
  __device__ static __forceinline__ V splat(S s) { return make_float2(s, s); }
  __device__ static __forceinline__ S get(const V& v, int x) { return x == 0 ? v.x : v.y; }
  __device__ static __forceinline__ void set(V& v, int x, S s) {
    if (x == 0) v.x = s; else v.y = s;
  }
  // per-half select on the nibble test of each read's match word
  __device__ static __forceinline__ V sel(const uint32_t* mw, uint32_t bit, V a, V b) {
    return make_float2((mw[0] & bit) ? a.x : b.x, (mw[1] & bit) ? a.y : b.y);
  }
  __device__ static __forceinline__ V shfl_up(V v, int width) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1, width), __shfl_up_sync(0xffffffffu, v.y, 1, width));
  }
};

struct VF1 {  // one read per lane, scalar fp32 (kept for measurement against VF2)
  typedef float V;
  typedef float S;
  static constexpr int NR = 1;
  static constexpr bool kDouble = false;
  __device__ static __forceinline__ V mul(V a, V b) { return a * b; }
  __device__ static __forceinline__ V fma(V a, V b, V c) { return fmaf(a, b, c); }
  __device__ static __forceinline__ V add(V a, V b) { return a + b; }
  __device__ static __forceinline__ V splat(S s) { return s; }
  __device__ static __forceinline__ S get(const V& v, int) { return v; }
  __device__ static __forceinline__ void set(V& v, int, S s) { v = s; }
  __device__ static __forceinline__ V sel(const uint32_t* mw, uint32_t bit, V a, V b) {
    return (mw[0] & bit) ? a : b;
  }
  __device__ static __forceinline__ V shfl_up(V v, int width) { return __shfl_up_sync(0xffffffffu, v, 1, width); }
};

struct VD1 {  // one read per lane, fp64 (useDoublePrecision and the fallback rerun)
  typedef double V;
  typedef double S;
  static constexpr int NR = 1;
  static constexpr bool kDouble = true;
  __device__ static __forceinline__ V mul(V a, V b) { return a * b; }
  __device__ static __forceinline__ V fma(V a, V b, V c) { return ::fma(a, b, c); }
  __device__ static __forceinline__ V add(V a, V b) { return a + b; }
  __device__ static __forceinline__ V splat(S s) { return s; }
  __device__ static __forceinline__ S get(const V& v, int) { return v; }
  __device__ static __forceinline__ void set(V& v, int, S s) { v = s; }
  __device__ static __forceinline__ V sel(const uint32_t* mw, uint32_t bit, V a, V b) {
    return (mw[0] & bit) ? a : b;
  }
  __device__ static __forceinline__ V shfl_up(V v, int width) { return __shfl_up_sync(0xffffffffu, v, 1, width); }
};

// ------------------------------------------------------------------------------------------
// Kernel parameters
// ------------------------------------------------------------------------------------------
struct PanelRef {          // one haplotype tile, as an image that is bulk-copied to shared memory:
  const uint8_t* image;    //   int32 hpos[n]  (byte offset of column 0 inside the image)
  uint32_t bytes;          //   int32 hlen[n]
  int n_haps;              //   pad to 16, then per haplotype: left margin | nibbles | right margin
  int hap0;                // index of the tile's first haplotype in the batch
  int n_haps_total;        // H of the tile's region (row pitch of its output)
  int max_hap_len;         // longest haplotype in the tile
  uint32_t smem_off;       // where this image sits inside the CTA's resident image (multi-panel launches; else 0)
};

struct ClassRef {          // the reads of one length class, packed by k_pack_reads
  const uint8_t* records;  // [n_rec][5 planes][stride]   planes: base nibble, qual, insGOP, delGOP, GCP (all & 127)
  const int32_t* rec_rid;  // [n_rec] read index in the batch, -1 for filler records
  const int32_t* rec_len;  // [n_rec] read length (0 for filler)
  int n_rec;               // multiple of the reads-per-warp of the class kernel
  int rows;                // rows per record = n_pass * G * K (reads are padded at the top up to this)
  int stride;              // bytes per plane = rows rounded up to 16
  int n_pass;              // passes of G*K rows
};

struct SweepParams {
  PanelRef panel;
  ClassRef cls;
  const void* ph2pr;            // S[128]
  const void* mm;               // S[kMmTableSize]
  double* out;                  // [n_reads][n_haps_total] log10 likelihoods
  // task mode
  int hap_chunk;                // haplotypes per task
  int n_chunks;                 // chunks per tile
  int n_tasks;                  // (n_rec / RPW) * n_chunks
  unsigned int* task_counter;
  // list mode (fp64 rerun of flagged pairs): items are (record, haplotype-in-batch)
  const uint2* list_items;
  const unsigned int* list_count;
  // where the fp32 kernel appends pairs whose scaled sum is < 1e-28f (IntelPairHmm.cc:159)
  uint2* fb_items;
  unsigned int* fb_count;
  // multi-pass carry scratch: per warp, per group, 3 * (max_hap_len + 2) V's
  void* carry;
  size_t carry_stride_bytes;    // bytes per warp
  double init_const;            // 2^120 (fp32) or 2^1020 (fp64)   Context.h:142,183
  double log10_init;            // log10f(2^120) widened, or log10(2^1020)
  uint32_t slot_bytes;          // per-warp shared-memory slot: packed records (task mode) + prior table (VAR 3)
};

// ------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk TMA
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// Per-lane constants of K rows (x NR reads)
// ------------------------------------------------------------------------------------------
template <class P, int K>
struct LaneRows {
  // VAR 0 (premultiplied): Am=(1-e)*pMM, Ax=(e/3)*pMM, Gm=(1-e)*pGAPM, Gx=(e/3)*pGAPM
  // VAR 1 (plain):         Am=pMM, Ax=pGAPM, Gm=1-e, Gx=e/3
  // VAR 2 (folded):        Am=pMM, Ax=pGAPM(next row), Gm=1-e, Gx=e/3, pMY holds pMY*pGAPM(next row):
  //                        the Y state is Y/pMY and the diagonal state is pGAPM(next row)*(X+Y), which
  //                        takes one multiply out of the Y update and one out of the M update
  // VAR 3 (table):         as VAR 2, but the prior (1-e | e/3) of every (row, haplotype symbol) sits in a per-warp
  //                        shared-memory table and arrives by LDS: no LOP3/FSEL and no Gm/Gx registers
  // VAR 4 (fp32 only):     as VAR 3, and the X state is kept as W = X / pMX(row): W = M(up) + kappa * W(up) with
  //                        kappa = pXX * pMX(row-1) / pMX(row) in pMX[], Ax = pGAPM(next row) * pMX(row); the final
  //                        sum is sumM + pMX(last row) * sumW.  One FMUL less per cell.  W can overflow fp32 for
  //                        adversarial gap-open patterns: a non-finite sum sends the pair to the fp64 rerun.
  typename P::V Am[K], Ax[K], Gm[K], Gx[K];
  typename P::V pMX[K], pXX[K], pMY[K];      // pYY == pXX (both ph2pr[gcp], avx-pairhmm-template.h:142-146)
  typename P::V pXXtop;                      // pXX[0] as used by the X update (zeroed on the first lane of pass 0)
  typename P::V gTop;                        // VAR >= 2: pGAPM of the lane's first row (0 if it is padding)
  typename P::V xlast;                       // VAR 4: pMX of the lane's last row (scales sumW on the last lane)
  uint32_t rbm[P::NR][(K + 7) / 8];          // read-base one-hot nibbles, row j at bits 4*(j%8) of word j/8
  uint32_t padmask[P::NR];                   // bit j set: row j is a top-padding row
};

__device__ __forceinline__ int mm_index(int ins_q, int del_q) {  // Context.h:156-167,197-209
  int mx = max(ins_q, del_q), mn = min(ins_q, del_q);
  return ((mx * (mx + 1)) >> 1) + mn;
}

// rec: one read's 5 planes (generic pointer: shared slot or global), rows [row0, row0+K).
// first_pad: number of top-padding rows of the whole read; top_is_row0: lane 0 of pass 0.
template <class P, int K, int VAR>
__device__ __forceinline__ void load_lane_rows(LaneRows<P, K>& L, int x, const uint8_t* rec, int stride, int n_rows,
                                               int row0, int n_pad, bool top_is_row0, const typename P::S* __restrict__ ph2pr,
                                               const typename P::S* __restrict__ mm, typename P::S* tbl = nullptr,
                                               int shift = 0) {
  // shift: the record holds n_rows - shift rows (it was packed for a class with fewer rows than this kernel
  // sweeps); kernel row `row` is record row `row - shift`, rows below n_pad >= shift are padding either way
  typedef typename P::S S;
  rec -= shift;
  uint32_t pm = 0;
#pragma unroll
  for (int w = 0; w < (K + 7) / 8; w++) L.rbm[x][w] = 0;
#pragma unroll
  for (int j = 0; j < K; j++) {
    const int row = row0 + j;
    const bool pad = row < n_pad;
    const int ri = max(row, shift);  // padding rows read a valid byte and ignore it
    const uint32_t nib = rec[ri];
    const int q = rec[stride + ri], ig = rec[2 * stride + ri], dg = rec[3 * stride + ri], cg = rec[4 * stride + ri];
    S e = ph2pr[q];
    S om = (S)1 - e;    // stripeINITIALIZATION: _1_distm = 1 - distm
    S th = e / (S)3;    //                       distm = distm / 3
    S pmm = __ldg(mm + mm_index(ig, dg));
    S pc = ph2pr[cg];
    S pgap = (S)1 - pc;
    S am, ax, gm, gx;
    if (VAR == 0) { am = om * pmm; ax = th * pmm; gm = om * pgap; gx = th * pgap; }
    else { am = pmm; ax = pgap; gm = om; gx = th; }
    S mx = ph2pr[ig], my = ph2pr[dg], xx = pc;
    if (pad) { am = ax = gm = gx = (S)0; mx = my = (S)0; xx = (S)1; pm |= 1u << j; }
    if (VAR >= 3) {
      // prior table: tbl is this lane's column; entry (symbol, row j) of read x at ((symbol * K + j) * 32) * NR + x
      const uint32_t symnib[kPriorSyms] = {1u, 2u, 4u, 8u, 15u};
#pragma unroll
      for (int sy = 0; sy < kPriorSyms; sy++)
        tbl[(size_t)((sy * K + j) * 32) * P::NR + x] = pad ? (S)0 : ((nib & symnib[sy]) ? om : th);
    }
    if (VAR >= 2) {
      // pGAPM of the row below (0 past the last row and for padding rows, whose M must stay 0)
      S gnext = (S)0;
      if (row + 1 < n_rows && row + 1 >= n_pad) gnext = (S)1 - ph2pr[rec[4 * stride + row + 1]];
      ax = gnext;
      my = (pad ? (S)1 : my) * gnext;
      if (j == 0) P::set(L.gTop, x, pad ? (S)0 : pgap);
    }
    if (VAR >= 4) {
      // W = X / pMX(row):  W = M(up) + kappa * W(up);  the row above a first real row has X = 0 (kappa = 0)
      const S pmx = ph2pr[ig];
      S kappa = (S)0;
      if (!pad && row - 1 >= n_pad) kappa = pc * ph2pr[rec[2 * stride + row - 1]] / pmx;
      ax = pad ? (S)0 : ax * pmx;   // pGAPM(next) * pMX(row) multiplies W in the diagonal state
      mx = kappa;
      if (j == K - 1) P::set(L.xlast, x, pad ? (S)0 : pmx);
    }
    S xxtop = xx;
    if (j == 0 && top_is_row0) {  // row 0 above: M = X = 0, so the M-diagonal and X inputs are killed
      am = (S)0;
      if (VAR == 0) ax = (S)0;
      mx = (S)0;
      xxtop = (S)0;
    }
    P::set(L.Am[j], x, am); P::set(L.Ax[j], x, ax); P::set(L.Gm[j], x, gm); P::set(L.Gx[j], x, gx);
    P::set(L.pMX[j], x, mx); P::set(L.pXX[j], x, xx); P::set(L.pMY[j], x, my);
    if (j == 0) P::set(L.pXXtop, x, xxtop);
    L.rbm[x][j / 8] |= (pad ? 0u : nib) << (4 * (j % 8));
  }
  L.padmask[x] = pm;
}

// ------------------------------------------------------------------------------------------
// One sweep of a group over one haplotype (one pass of G*K rows).
//   hap       shared-memory pointer to column 0 of this lane's haplotype (margins on both sides)
//   haplen    its length
//   steady_end  warp-uniform: every lane of the warp has 1 <= c <= haplen for steps G..steady_end
//   n_steps   warp-uniform loop bound >= haplen + G - 1
//   carry_in / carry_out: bottom row of the previous / this pass (nullptr when unused),
//   layout [3][carry_pitch] V, index c
// Steps G..steady_end run without the per-lane activity test: one basic block in which the
// shuffles for the next step are issued last, so that their latency is covered by the M block of
// the next step (which depends only on the previous step's registers).
// Returns the running sum of M+X over the lane's bottom row (meaningful on the last lane).
// ------------------------------------------------------------------------------------------
template <class P, int G, int K, bool MULTI, int VAR>
struct Sweeper {
  typedef typename P::V V;
  const LaneRows<P, K>& L;
  V Ml[K], Yl[K], Zl[K];  // previous column: M, Y (VAR 2: Y / pMY), X+Y (VAR 2: pGAPM_next * (X+Y))
  V botX, sum, sumW;
  V uM, uX, uZ;           // bottom row of the lane above at this step's column
  V dMp, dZp;             // ... and at the previous column
  V uM1, uX1, uZ1;        // VAR 6: bottom row of the lane above at the step's second column
  V aM, aX, aZ;           // VAR 6: this lane's bottom row at the step's first column (the second one is the state)
  uint32_t hb1;           // VAR 6: symbol byte of the step's second column
  V inj;                  // what row 0 feeds the first row's M update
  const uint8_t* hap;
  int haplen, c;
  uint32_t hb;            // VAR 0-2: symbol byte of column c;  VAR 3: symbol byte of column c + 1
  const V* tb;            // VAR 3: this lane's column of the warp's prior table
  V pr[K];                // VAR 3: priors of column c
  bool first, row0_above, last;
  const V* carry_in;
  V* carry_out;
  int carry_pitch;

  __device__ __forceinline__ Sweeper(const LaneRows<P, K>& L_) : L(L_) {}

  __device__ __forceinline__ void fetch_up() {
    uM = P::shfl_up(Ml[K - 1], G);
    uX = P::shfl_up(botX, G);
    uZ = P::shfl_up(Zl[K - 1], G);
    if (row0_above) {  // row 0: M = X = 0 (killed by the zeroed top constants), Y = init
      uZ = inj;
      if (VAR >= 4) uM = P::splat(0);  // the W update adds M(up) without a coefficient
    }
    if (MULTI) {
      if (first && carry_in != nullptr) {
        const int cc = min(max(c, 0), haplen + 1);
        uM = carry_in[cc];
        uX = carry_in[carry_pitch + cc];
        uZ = carry_in[2 * carry_pitch + cc];
      }
    }
  }

  __device__ __forceinline__ void cells(uint32_t hrep) {
    uint32_t mw[P::NR][(K + 7) / 8];
#pragma unroll
    for (int x = 0; x < P::NR; x++)
#pragma unroll
      for (int w = 0; w < (K + 7) / 8; w++) mw[x][w] = (VAR >= 3) ? 0u : (L.rbm[x][w] & hrep);
    V dM = dMp, dZ = dZp, upM = uM, upX = uX;
#pragma unroll
    for (int j = 0; j < K; j++) {
      uint32_t m2[P::NR];
#pragma unroll
      for (int x = 0; x < P::NR; x++) m2[x] = mw[x][j / 8];
      const uint32_t bit = 0xFu << (4 * (j % 8));
      V Mn, Yn, Zn;
      const V Xn = (VAR >= 4) ? P::fma(L.pMX[j], upX, upM)  // W form: kappa * W(up) + M(up)
                              : P::fma(j == 0 ? L.pXXtop : L.pXX[j], upX, P::mul(L.pMX[j], upM));
      if (VAR >= 3) {
        Mn = P::mul(pr[j], P::fma(L.Am[j], dM, dZ));
        Yn = P::fma(L.pXX[j], Yl[j], Ml[j]);
        Zn = P::fma(L.pMY[j], Yn, P::mul(L.Ax[j], Xn));
      } else if (VAR == 0) {
        const V A = P::sel(m2, bit, L.Am[j], L.Ax[j]);
        const V Gs = P::sel(m2, bit, L.Gm[j], L.Gx[j]);
        Mn = P::fma(A, dM, P::mul(Gs, dZ));
        Yn = P::fma(L.pXX[j], Yl[j], P::mul(L.pMY[j], Ml[j]));
        Zn = P::add(Xn, Yn);
      } else if (VAR == 1) {
        const V prior = P::sel(m2, bit, L.Gm[j], L.Gx[j]);
        Mn = P::mul(prior, P::fma(L.Am[j], dM, P::mul(L.Ax[j], dZ)));
        Yn = P::fma(L.pXX[j], Yl[j], P::mul(L.pMY[j], Ml[j]));
        Zn = P::add(Xn, Yn);
      } else {
        const V prior = P::sel(m2, bit, L.Gm[j], L.Gx[j]);
        Mn = P::mul(prior, P::fma(L.Am[j], dM, dZ));       // dZ already carries this row's pGAPM
        Yn = P::fma(L.pXX[j], Yl[j], Ml[j]);               // Y / pMY
        Zn = P::fma(L.pMY[j], Yn, P::mul(L.Ax[j], Xn));    // pGAPM_next * (X + pMY * Y/pMY)
      }
      dM = Ml[j];
      dZ = Zl[j];
      Ml[j] = Mn;
      Yl[j] = Yn;
      Zl[j] = Zn;
      upM = Mn;
      upX = Xn;
    }
    botX = upX;
    if (VAR >= 4) {
      sum = P::add(sum, upM);
      sumW = P::add(sumW, upX);
    } else {
      sum = P::add(sum, P::add(upM, upX));
    }
    if (MULTI) {
      if (carry_out != nullptr && last) {
        carry_out[c] = upM;
        carry_out[carry_pitch + c] = upX;
        carry_out[2 * carry_pitch + c] = Zl[K - 1];
      }
    }
  }


  // ---- VAR 6: two columns per step (fp32 W form, no prior prefetch) ------------------------------------
  // Lane t processes columns c and c + 1 (c = 2 (s - t) - 1) row by row; the second column's cells read the
  // same per-row constants as the first one's in the same operand slots, which lets the operand-reuse cache
  // take a third of the register reads away, and the two columns give two dependency chains per warp.
  __device__ __forceinline__ void fetch_up2() {
    uM = P::shfl_up(aM, G);
    uX = P::shfl_up(aX, G);
    uZ = P::shfl_up(aZ, G);
    uM1 = P::shfl_up(Ml[K - 1], G);
    uX1 = P::shfl_up(botX, G);
    uZ1 = P::shfl_up(Zl[K - 1], G);
    if (row0_above) {
      uZ = inj;
      uZ1 = inj;
      uM = P::splat(0);
      uM1 = P::splat(0);
    }
  }

  template <bool GUARD>
  __device__ __forceinline__ void cells2(const V* t0, const V* t1) {
    V dM0 = dMp, dZ0 = dZp;       // row above, column c - 1
    V upM0 = uM, upX0 = uX;       // row above, column c
    V dZ1 = uZ;                   // ... its diagonal state (feeds column c + 1)
    V upM1 = uM1, upX1 = uX1;     // row above, column c + 1
#pragma unroll
    for (int j = 0; j < K; j++) {
      const V p0 = t0[j * 32], p1 = t1[j * 32];
      const V Xa = P::fma(L.pMX[j], upX0, upM0);
      const V Xb = P::fma(L.pMX[j], upX1, upM1);
      const V Ma = P::mul(p0, P::fma(L.Am[j], dM0, dZ0));
      const V Mb = P::mul(p1, P::fma(L.Am[j], upM0, dZ1));
      const V Ya = P::fma(L.pXX[j], Yl[j], Ml[j]);
      const V Yb = P::fma(L.pXX[j], Ya, Ma);
      const V Za = P::fma(L.pMY[j], Ya, P::mul(L.Ax[j], Xa));
      const V Zb = P::fma(L.pMY[j], Yb, P::mul(L.Ax[j], Xb));
      dM0 = Ml[j];
      dZ0 = Zl[j];
      upM0 = Ma;
      upX0 = Xa;
      dZ1 = Za;
      upM1 = Mb;
      upX1 = Xb;
      Ml[j] = Mb;
      Yl[j] = Yb;
      Zl[j] = Zb;
    }
    aM = upM0;
    aX = upX0;
    aZ = dZ1;
    botX = upX1;
    sum = P::add(sum, upM0);
    sumW = P::add(sumW, upX0);
    if (!GUARD || c + 1 <= haplen) {  // an odd haplotype length ends on a first column
      sum = P::add(sum, upM1);
      sumW = P::add(sumW, upX1);
    }
  }

  template <bool GUARD>
  __device__ __forceinline__ void step2() {
    const V* t0 = tb + (size_t)(hb & 7u) * (K * 32);
    const V* t1 = tb + (size_t)(hb1 & 7u) * (K * 32);
    const bool act = !GUARD || (unsigned)(c - 1) < (unsigned)haplen;
    if (GUARD) {
      hb = hap[min(max(c + 2, -kHapLeftMargin + 1), haplen + 1)];
      hb1 = hap[min(max(c + 3, -kHapLeftMargin + 1), haplen + 1)];
    } else {
      hb = hap[c + 2];   // c + 3 <= haplen + 2: inside the right margin
      hb1 = hap[c + 3];
    }
    if (act) cells2<GUARD>(t0, t1);
    dMp = uM1;
    dZp = uZ1;
    c += 2;
    fetch_up2();
  }

  __device__ __forceinline__ V run2(const uint8_t* hap_, int haplen_, int t, typename P::S initY, const V* tb_) {
    tb = tb_;
    const V zero = P::splat(0);
    hap = hap_;
    haplen = haplen_;
    first = (t == 0);
    last = (t == G - 1);
    row0_above = first;
    carry_in = nullptr;
    carry_out = nullptr;
    carry_pitch = 0;
    const V initYv = P::splat(initY);
    inj = P::mul(L.gTop, initYv);
#pragma unroll
    for (int j = 0; j < K; j++) {
      Ml[j] = zero;
      V y0 = zero;
#pragma unroll
      for (int x = 0; x < P::NR; x++)
        if (L.padmask[x] & (1u << j)) P::set(y0, x, initY);
      Yl[j] = y0;
      Zl[j] = P::mul(L.pMY[j], y0);
    }
    botX = zero;
    sum = zero;
    sumW = zero;
    aM = zero;   // column 0 of the bottom row, as both of the "previous step's" columns
    aX = zero;
    aZ = Zl[K - 1];
    dMp = zero;
    dZp = row0_above ? inj : zero;
    c = 1 - 2 * t;
    hb = hap[max(c, -kHapLeftMargin + 1)];
    hb1 = hap[max(c + 1, -kHapLeftMargin + 1)];
    fetch_up2();
    const int n_steps = (haplen + 1) / 2 + G - 1;
    const int steady_end = haplen / 2;   // lane 0 (the most advanced) still has both columns inside
    int s = 1;
    const int pre_end = min(G - 1, n_steps);
    for (; s <= pre_end; s++) step2<true>();
    for (; s <= steady_end; s++) step2<false>();
    for (; s <= n_steps; s++) step2<true>();
    return P::fma(L.xlast, sumW, sum);
  }

  template <bool GUARD>
  __device__ __forceinline__ void step() {
    if (VAR == 5) {  // measurement variant: VAR 4 without the one-column-ahead prior prefetch (fewer registers)
      const V* tn = tb + (size_t)(hb & 7u) * (K * 32);
#pragma unroll
      for (int j = 0; j < K; j++) pr[j] = tn[j * 32];
      hb = hap[min(c + 1, haplen + 1)];
      if (!GUARD || (unsigned)(c - 1) < (unsigned)haplen) cells(0u);
    } else if (VAR >= 3) {
      V prn[K];  // priors of column c + 1, in flight while column c is computed
      const V* tn = tb + (size_t)(hb & 7u) * (K * 32);
#pragma unroll
      for (int j = 0; j < K; j++) prn[j] = tn[j * 32];
      hb = hap[min(c + 2, haplen + 1)];
      if (!GUARD || (unsigned)(c - 1) < (unsigned)haplen) cells(0u);
#pragma unroll
      for (int j = 0; j < K; j++) pr[j] = prn[j];
    } else {
      const uint32_t hrep = (hb >> 4) * 0x11111111u;
      hb = hap[min(c + 1, haplen + 1)];  // prefetch the next column's symbol
      if (!GUARD || (unsigned)(c - 1) < (unsigned)haplen) cells(hrep);
    }
    dMp = uM;
    dZp = uZ;
    c++;
    fetch_up();
  }

  __device__ __forceinline__ V run(const uint8_t* hap_, int haplen_, int steady_end, int n_steps, int t,
                                   typename P::S initY, bool pass0, const V* carry_in_, V* carry_out_,
                                   int carry_pitch_, const V* tb_) {
    tb = tb_;
    const V zero = P::splat(0);
    hap = hap_;
    haplen = haplen_;
    first = (t == 0);
    last = (t == G - 1);
    row0_above = first && pass0;
    carry_in = (MULTI && !pass0) ? carry_in_ : nullptr;
    carry_out = carry_out_;
    carry_pitch = carry_pitch_;
    const V initYv = P::splat(initY);
    inj = (VAR >= 2) ? P::mul(L.gTop, initYv) : initYv;
#pragma unroll
    for (int j = 0; j < K; j++) {
      Ml[j] = zero;
      V y0 = zero;
#pragma unroll
      for (int x = 0; x < P::NR; x++)
        if (L.padmask[x] & (1u << j)) P::set(y0, x, initY);
      Yl[j] = y0;
      Zl[j] = (VAR >= 2) ? P::mul(L.pMY[j], y0) : y0;
    }
    botX = zero;
    sum = zero;
    sumW = zero;
    // column 0 of the row above the lane's first row: row 0 (M = 0, Y = init) for the first lane
    // of pass 0, the previous pass's bottom row for the first lane of later passes
    dMp = zero;
    dZp = row0_above ? inj : zero;
    c = 1 - t;
    if (MULTI) {
      if (first && carry_in != nullptr) {
        dMp = carry_in[0];
        dZp = carry_in[2 * carry_pitch];
      }
      if (carry_out != nullptr && last) {  // column 0 of this pass's bottom row
        carry_out[0] = zero;
        carry_out[carry_pitch] = zero;
        carry_out[2 * carry_pitch] = Zl[K - 1];
      }
    }
    hb = hap[max(c, -kHapLeftMargin + 1)];
    if (VAR == 3 || VAR == 4) {
      const V* t0 = tb + (size_t)(hb & 7u) * (K * 32);
#pragma unroll
      for (int j = 0; j < K; j++) pr[j] = t0[j * 32];
      hb = hap[min(max(c + 1, -kHapLeftMargin + 1), haplen + 1)];
    }
    fetch_up();
    int s = 1;
    const int pre_end = min(G - 1, n_steps);
    for (; s <= pre_end; s++) step<true>();
#pragma unroll 2
    for (; s <= steady_end; s++) step<false>();
    for (; s <= n_steps; s++) step<true>();
    return (VAR >= 4) ? P::fma(L.xlast, sumW, sum) : sum;
  }
};

template <class P, int G, int K, bool MULTI, int VAR>
__device__ __forceinline__ typename P::V sweep(const LaneRows<P, K>& L, const uint8_t* hap, int haplen,
                                               int steady_end, int n_steps, int t, typename P::S initY, bool pass0,
                                               const typename P::V* carry_in, typename P::V* carry_out,
                                               int carry_pitch, const typename P::V* tb = nullptr) {
  Sweeper<P, G, K, MULTI, VAR> sw(L);
  if constexpr (VAR == 6 && !MULTI && !P::kDouble) {  // the two-column sweep is single-pass fp32 only
    return sw.run2(hap, haplen, t, initY, tb);
  }
  return sw.run(hap, haplen, steady_end, n_steps, t, initY, pass0, carry_in, carry_out, carry_pitch, tb);
}

// Result of one pair (IntelPairHmm.cc:157-167).  Returns false when the pair must be rerun in fp64.
template <class P>
__device__ __forceinline__ bool finish_pair(typename P::S sum, double log10_init, double* out) {
  if (P::kDouble) {
    *out = log10((double)sum) - log10_init;
    return true;
  } else {
    // below GKL's threshold -> fp64 rerun; so does a non-finite sum (the W form of VAR 4 can overflow)
    if (!((float)sum >= kMinAccepted) || (float)sum > 3.0e38f) return false;
    // log10f(result_float) - log10f(2^120), evaluated in float, widened
    const float lg = (float)log10((double)sum);
    *out = (double)(lg - (float)log10_init);
    return true;
  }
}

// Shared memory layout of the sweep kernels.
struct SmemLayout {
  uint32_t bars;     // offset of mbarriers: [0] panel, [1 + w] slot of warp w
  uint32_t ends;     // int[n_ends]: exclusive task prefix of the classes of a multi-class launch
  uint32_t ph2pr;    // S[128]
  uint32_t panel;    // panel image(s)
  uint32_t slots;    // per-warp record slots
  uint32_t slot_bytes;
  uint32_t total;
};
__host__ __device__ inline SmemLayout smem_layout(int warps, uint32_t panel_bytes, uint32_t slot_bytes,
                                                  uint32_t scalar_bytes, uint32_t n_ends = 0) {
  SmemLayout l;
  l.bars = 0;
  l.ends = (uint32_t)((8 * (1 + warps) + 15) / 16 * 16);
  l.ph2pr = (uint32_t)((l.ends + 4 * n_ends + 127) / 128 * 128);
  l.panel = l.ph2pr + 128 * scalar_bytes;
  l.slots = (l.panel + panel_bytes + 127) / 128 * 128;
  l.slot_bytes = (slot_bytes + 127) / 128 * 128;
  l.total = l.slots + (uint32_t)warps * l.slot_bytes;
  return l;
}

// ------------------------------------------------------------------------------------------
// Per-CTA setup shared by all sweep kernels: mbarriers, ph2pr table, panel image (one TMA bulk copy).
// ------------------------------------------------------------------------------------------
// The context holds OFFSETS into the dynamic shared memory, not pointers: device functions rebuild their pointers
// from the `extern __shared__` symbol so that the compiler keeps emitting LDS/STS (a generic pointer that crosses
// the call boundary of the multi-class kernels' task functions turns every table and symbol load into a generic LD).
template <class S>
struct WarpCtx {
  uint32_t slot_bar;       // offset of this warp's record-slot mbarrier
  uint32_t slot_parity;
  uint32_t slot;           // offset of this warp's record slot
  uint32_t ph2pr;          // offset of S[128]
  uint32_t image;          // offset of the CTA's resident image (one panel, or the panels of a multi-region launch)
  uint32_t ends;           // offset of int[n_ends]: exclusive task prefix per class (multi-class launches)
  uint32_t panel;          // offset of the panel the current task works on
  int warp, lane;
  size_t warp_global;      // index of this warp in the grid
};

// Point a warp at one panel of the resident image.
template <class S>
__device__ __forceinline__ void bind_panel(WarpCtx<S>& c, const PanelRef& panel) {
  c.panel = c.image + panel.smem_off;
}

template <class S>
__device__ __forceinline__ WarpCtx<S> setup_cta(const uint8_t* image, uint32_t image_bytes, const void* ph2pr, int warps,
                                                uint32_t slot_bytes, uint32_t n_ends = 0) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout lay = smem_layout(warps, image_bytes, slot_bytes, sizeof(S), n_ends);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  S* ph2pr_s = reinterpret_cast<S*>(smem + lay.ph2pr);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 1 + warps; i++) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) ph2pr_s[i] = reinterpret_cast<const S*>(ph2pr)[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], image_bytes);
    tma_bulk_g2s(smem + lay.panel, image, image_bytes, &bars[0]);
  }
  mbar_wait(&bars[0], 0);
  WarpCtx<S> c;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.slot_bar = lay.bars + 8u * (1 + c.warp);
  c.slot_parity = 0;
  c.slot = lay.slots + (uint32_t)c.warp * lay.slot_bytes;
  c.ph2pr = lay.ph2pr;
  c.image = lay.panel;
  c.ends = lay.ends;
  c.panel = lay.panel;
  c.warp_global = (size_t)blockIdx.x * warps + c.warp;
  return c;
}

// Class of a task in a multi-class queue: first c with task < ends[c] (ends is an exclusive prefix in shared memory).
__device__ __forceinline__ int class_of_task(const int* ends, int n, unsigned int task) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (task < (unsigned)ends[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------
// One task = (block of RPW records) x (chunk of haplotypes), executed by one warp.
// ------------------------------------------------------------------------------------------
template <class P, int G, int K, bool MULTI, int VAR>
__device__ __forceinline__ void run_task(const SweepParams& p, unsigned int task, WarpCtx<typename P::S>& ctx) {
  typedef typename P::V V;
  typedef typename P::S S;
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int GPW = 32 / G;           // groups per warp
  constexpr int RPW = GPW * P::NR;      // records per warp-task
  uint8_t* const slot = smem + ctx.slot;
  uint64_t* const slot_bar = reinterpret_cast<uint64_t*>(smem + ctx.slot_bar);
  const S* const ph2pr_s = reinterpret_cast<const S*>(smem + ctx.ph2pr);
  const uint8_t* const panel_s = smem + ctx.panel;
  const int32_t* const hpos = reinterpret_cast<const int32_t*>(panel_s);
  const int32_t* const hlen = hpos + p.panel.n_haps;
  const int lane = ctx.lane;
  const int t = lane % G, g = lane / G;
  const uint32_t rec_bytes = 5u * (uint32_t)p.cls.stride;
  const int cap = G * K;
  const int carry_pitch = p.panel.max_hap_len + 2;
  V* carry = MULTI ? reinterpret_cast<V*>(reinterpret_cast<uint8_t*>(p.carry) + ctx.warp_global * p.carry_stride_bytes)
                   : nullptr;
  const int blk = task / p.n_chunks, chunk = task - blk * p.n_chunks;
  const int rec0 = blk * RPW;
  // VAR >= 3: the warp's prior table sits behind the record slot; lane-private column.  Multi-pass classes
  // (reads of thousands of rows) do not stage their records: the rows are read from global memory pass by pass.
  V* tbv = reinterpret_cast<V*>(slot + (MULTI ? 0u : ((RPW * rec_bytes + 127u) & ~127u))) + lane;
  S* tbs = reinterpret_cast<S*>(tbv);
  if (!MULTI) {
    // stage the block's packed records into this warp's slot
    __syncwarp();
    if (lane == 0) {
      fence_proxy_async();
      mbar_expect_tx(slot_bar, RPW * rec_bytes);
      tma_bulk_g2s(slot, p.cls.records + (size_t)rec0 * rec_bytes, RPW * rec_bytes, slot_bar);
    }
    mbar_wait(slot_bar, ctx.slot_parity);
    ctx.slot_parity ^= 1;
  }

  const int krows = MULTI ? p.cls.rows : cap;    // rows this kernel sweeps; the records may hold fewer (see load_lane_rows)
  const int shift = krows - p.cls.rows;
  int rid[P::NR], npad[P::NR];
#pragma unroll
  for (int x = 0; x < P::NR; x++) {
    const int rec = rec0 + g * P::NR + x;
    rid[x] = p.cls.rec_rid[rec];
    npad[x] = krows - p.cls.rec_len[rec];
  }
  const int h_begin = chunk * p.hap_chunk, h_end = min(p.panel.n_haps, h_begin + p.hap_chunk);

  LaneRows<P, K> L;
  if (!MULTI) {
#pragma unroll
    for (int x = 0; x < P::NR; x++)
      load_lane_rows<P, K, VAR>(L, x, slot + (size_t)(g * P::NR + x) * rec_bytes, p.cls.stride, krows, t * K,
                                npad[x], t == 0, ph2pr_s, reinterpret_cast<const S*>(p.mm), tbs, shift);
  }
  for (int h = h_begin; h < h_end; h++) {
    const int haplen = hlen[h];
    const uint8_t* hap = panel_s + hpos[h];
    const S initY = (S)p.init_const / (S)haplen;
    const int n_steps = haplen + G - 1;
    const int steady_end = haplen;
    V sum = P::splat(0);
    if (!MULTI) {
      sum = sweep<P, G, K, false, VAR>(L, hap, haplen, steady_end, n_steps, t, initY, true, nullptr, nullptr, 0, tbv);
    } else {
      V* cg = carry + (size_t)g * 6 * carry_pitch;  // two buffers of 3 lines, ping-pong
      for (int pass = 0; pass < p.cls.n_pass; pass++) {
#pragma unroll
        for (int x = 0; x < P::NR; x++)
          load_lane_rows<P, K, VAR>(L, x, p.cls.records + (size_t)(rec0 + g * P::NR + x) * rec_bytes, p.cls.stride,
                                    p.cls.rows, pass * cap + t * K, npad[x], t == 0 && pass == 0, ph2pr_s,
                                    reinterpret_cast<const S*>(p.mm), tbs);
        V* cin = cg + (size_t)(pass & 1) * 3 * carry_pitch;
        V* cout = cg + (size_t)((pass + 1) & 1) * 3 * carry_pitch;
        sum = sweep<P, G, K, true, VAR>(L, hap, haplen, steady_end, n_steps, t, initY, pass == 0, cin,
                                        pass + 1 < p.cls.n_pass ? cout : nullptr, carry_pitch, tbv);
        __syncwarp();  // carry written by lane G-1 is read by lane 0 of the next pass
      }
    }
    if (t == G - 1) {
#pragma unroll
      for (int x = 0; x < P::NR; x++) {
        if (rid[x] >= 0) {
          double* o = p.out + (size_t)rid[x] * p.panel.n_haps_total + (p.panel.hap0 + h);
          if (!finish_pair<P>(P::get(sum, x), p.log10_init, o)) {
            *o = __longlong_as_double(0x7ff8000000000000LL);  // overwritten by the fp64 rerun
            const unsigned int k = atomicAdd(p.fb_count, 1u);
            p.fb_items[k] = make_uint2((unsigned)(rec0 + g * P::NR + x), (unsigned)(p.panel.hap0 + h));
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// One warp-item of the rerun list: every group takes one (record, haplotype) pair.  NR must be 1.
// ------------------------------------------------------------------------------------------
template <class P, int G, int K, bool MULTI, int VAR>
__device__ __forceinline__ void run_list_item(const SweepParams& p, unsigned int wi, unsigned int n_items,
                                              WarpCtx<typename P::S>& ctx) {
  typedef typename P::V V;
  typedef typename P::S S;
  static_assert(P::NR == 1, "list items carry one read per lane");
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int GPW = 32 / G;
  const S* const ph2pr_s = reinterpret_cast<const S*>(smem + ctx.ph2pr);
  const uint8_t* const panel_s = smem + ctx.panel;
  const int32_t* const hpos = reinterpret_cast<const int32_t*>(panel_s);
  const int32_t* const hlen = hpos + p.panel.n_haps;
  const int lane = ctx.lane;
  const int t = lane % G, g = lane / G;
  const uint32_t rec_bytes = 5u * (uint32_t)p.cls.stride;
  const int cap = G * K;
  const int carry_pitch = p.panel.max_hap_len + 2;
  V* carry = MULTI ? reinterpret_cast<V*>(reinterpret_cast<uint8_t*>(p.carry) + ctx.warp_global * p.carry_stride_bytes)
                   : nullptr;
  V* tbv = reinterpret_cast<V*>(smem + ctx.slot) + lane;  // VAR 3: the slot holds only the prior table here
  S* tbs = reinterpret_cast<S*>(tbv);
  const unsigned int item = wi * GPW + g;
  const bool valid = item < n_items;
  const uint2 it = valid ? p.list_items[item] : make_uint2(0u, 0u);
  const int h = (int)it.y - p.panel.hap0;
  const bool mine = valid && h >= 0 && h < p.panel.n_haps;  // items of other tiles are skipped
  const int rec = (int)it.x;
  const int haplen = mine ? hlen[h] : 0;
  const uint8_t* hap = panel_s + (mine ? hpos[h] : hpos[0]);
  const int krows = MULTI ? p.cls.rows : cap;
  const int shift = krows - p.cls.rows;
  const int rid = mine ? p.cls.rec_rid[rec] : -1;
  const int npad = mine ? krows - p.cls.rec_len[rec] : 0;
  const S initY = (S)p.init_const / (S)max(haplen, 1);
  int n_steps = mine ? haplen + G - 1 : 0;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) n_steps = max(n_steps, __shfl_xor_sync(0xffffffffu, n_steps, o));
  int steady_end = haplen;  // steps G..min(haplen) are unguarded; a warp with an idle group has none
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) steady_end = min(steady_end, __shfl_xor_sync(0xffffffffu, steady_end, o));
  if (n_steps == 0) return;
  const uint8_t* recp = p.cls.records + (size_t)rec * rec_bytes;
  LaneRows<P, K> L;
  V sum = P::splat(0);
  if (!MULTI) {
    load_lane_rows<P, K, VAR>(L, 0, recp, p.cls.stride, krows, t * K, npad, t == 0, ph2pr_s,
                              reinterpret_cast<const S*>(p.mm), tbs, shift);
    sum = sweep<P, G, K, false, VAR>(L, hap, haplen, steady_end, n_steps, t, initY, true, nullptr, nullptr, 0, tbv);
  } else {
    V* cg = carry + (size_t)g * 6 * carry_pitch;
    for (int pass = 0; pass < p.cls.n_pass; pass++) {
      load_lane_rows<P, K, VAR>(L, 0, recp, p.cls.stride, p.cls.rows, pass * cap + t * K, npad, t == 0 && pass == 0,
                                ph2pr_s, reinterpret_cast<const S*>(p.mm), tbs);
      V* cin = cg + (size_t)(pass & 1) * 3 * carry_pitch;
      V* cout = cg + (size_t)((pass + 1) & 1) * 3 * carry_pitch;
      sum = sweep<P, G, K, true, VAR>(L, hap, haplen, steady_end, n_steps, t, initY, pass == 0, cin,
                                      pass + 1 < p.cls.n_pass ? cout : nullptr, carry_pitch, tbv);
      __syncwarp();
    }
  }
  if (t == G - 1 && mine && rid >= 0) {
    double* o = p.out + (size_t)rid * p.panel.n_haps_total + (p.panel.hap0 + h);
    finish_pair<P>(P::get(sum, 0), p.log10_init, o);
  }
}

// ------------------------------------------------------------------------------------------
// Single-class kernels: persistent CTAs, tasks / list items pulled from an atomic counter.
// ------------------------------------------------------------------------------------------
template <class P, int G, int K, int WARPS, bool MULTI, int VAR>
__global__ void __launch_bounds__(WARPS * 32, 1) k_sweep_tasks(const SweepParams p) {
  typedef typename P::S S;
  WarpCtx<S> ctx = setup_cta<S>(p.panel.image, p.panel.bytes, p.ph2pr, WARPS, p.slot_bytes);
  bind_panel(ctx, p.panel);
  for (;;) {
    unsigned int task = 0;
    if (ctx.lane == 0) task = atomicAdd(p.task_counter, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= (unsigned)p.n_tasks) break;
    run_task<P, G, K, MULTI, VAR>(p, task, ctx);
  }
}

template <class P, int G, int K, int WARPS, bool MULTI, int VAR>
__global__ void __launch_bounds__(WARPS * 32, 1) k_sweep_list(const SweepParams p) {
  typedef typename P::S S;
  constexpr int GPW = 32 / G;
  const unsigned int n_items = *p.list_count;
  if (n_items == 0) return;
  WarpCtx<S> ctx = setup_cta<S>(p.panel.image, p.panel.bytes, p.ph2pr, WARPS, p.slot_bytes);
  bind_panel(ctx, p.panel);
  const unsigned int n_warp_items = (n_items + GPW - 1) / GPW;
  for (;;) {
    unsigned int wi = 0;
    if (ctx.lane == 0) wi = atomicAdd(p.task_counter, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= n_warp_items) break;
    run_list_item<P, G, K, MULTI, VAR>(p, wi, n_items, ctx);
  }
}

// ------------------------------------------------------------------------------------------
// Multi-class ("mega") kernels: one launch covers every length class of a batch.  Each warp pulls from
// ONE queue that concatenates the classes' tasks (longest class first) and dispatches on the class
// configuration, so a HaplotypeCaller-shaped batch with a dozen length classes of ~25 reads each
// fills the GPU with a single launch instead of a dozen serialised under-filled ones.
// ------------------------------------------------------------------------------------------
constexpr int kMaxMegaClasses = 1024;  // (class, panel) entries of one multi-class launch
constexpr int kCfgMulti = 17;  // configuration index of the multi-pass class (32 x 8 rows per pass)

struct MegaParams {
  int n_classes;
  const int* cfg;                  // [n_classes] index into the class table (G, K), kCfgMulti for multi-pass   (device)
  const int* task_end;             // [n_classes] exclusive prefix of tasks in the unified queue (task mode)    (device)
  unsigned int* queue;             // the unified work counter
  const SweepParams* cls;          // [n_classes] (device memory, uploaded with the batch's meta block)
  const uint8_t* image;            // the panels of the launch, contiguous; each class's panel.smem_off points into it
  uint32_t image_bytes;
};

// product variants: fp32 uses the W form, fp64 must not (2^1020 leaves no headroom for X / pMX).  With 12 warps
// per CTA (168 registers) the fp32 kernels drop the one-column-ahead prior prefetch (VAR 5); with 8 they keep it.
template <class P, int WARPS> struct ProductVar { static constexpr int value = P::kDouble ? 3 : (WARPS >= 12 ? 5 : 4); };
// The task functions of the multi-class kernels take the context by value (no pointer into the caller's frame) and
// return the slot barrier's phase.
template <class P, int G, int K, bool MULTI, int WARPS>
__device__ __noinline__ uint32_t mega_task(const SweepParams& p, unsigned int task, WarpCtx<typename P::S> ctx) {
  run_task<P, G, K, MULTI, ProductVar<P, WARPS>::value>(p, task, ctx);
  return ctx.slot_parity;
}
template <class P, int G, int K, bool MULTI, int WARPS>
__device__ __noinline__ uint32_t mega_item(const SweepParams& p, unsigned int wi, unsigned int n_items,
                                           WarpCtx<typename P::S> ctx) {
  run_list_item<P, G, K, MULTI, ProductVar<P, WARPS>::value>(p, wi, n_items, ctx);
  return ctx.slot_parity;
}

#define GKLB_MEGA_DISPATCH(FN, ...)                                   \
  switch (cfg) {                                                      \
    case 0: ctx.slot_parity = FN<P, 8, 4, false, WARPS>(__VA_ARGS__); break;                   \
    case 1: ctx.slot_parity = FN<P, 8, 5, false, WARPS>(__VA_ARGS__); break;                   \
    case 2: ctx.slot_parity = FN<P, 8, 6, false, WARPS>(__VA_ARGS__); break;                   \
    case 3: ctx.slot_parity = FN<P, 8, 7, false, WARPS>(__VA_ARGS__); break;                   \
    case 4: ctx.slot_parity = FN<P, 8, 8, false, WARPS>(__VA_ARGS__); break;                   \
    case 5: ctx.slot_parity = FN<P, 16, 5, false, WARPS>(__VA_ARGS__); break;                  \
    case 6: ctx.slot_parity = FN<P, 16, 6, false, WARPS>(__VA_ARGS__); break;                  \
    case 7: ctx.slot_parity = FN<P, 16, 7, false, WARPS>(__VA_ARGS__); break;                  \
    case 8: ctx.slot_parity = FN<P, 16, 8, false, WARPS>(__VA_ARGS__); break;                  \
    case 9: ctx.slot_parity = FN<P, 16, 9, false, WARPS>(__VA_ARGS__); break;                  \
    case 10: ctx.slot_parity = FN<P, 16, 10, false, WARPS>(__VA_ARGS__); break;                \
    case 11: ctx.slot_parity = FN<P, 32, 5, false, WARPS>(__VA_ARGS__); break;                 \
    case 12: ctx.slot_parity = FN<P, 32, 6, false, WARPS>(__VA_ARGS__); break;                 \
    case 13: ctx.slot_parity = FN<P, 32, 7, false, WARPS>(__VA_ARGS__); break;                 \
    case 14: ctx.slot_parity = FN<P, 32, 8, false, WARPS>(__VA_ARGS__); break;                 \
    case 15: ctx.slot_parity = FN<P, 32, 9, false, WARPS>(__VA_ARGS__); break;                 \
    case 16: ctx.slot_parity = FN<P, 32, 10, false, WARPS>(__VA_ARGS__); break;                \
    default: ctx.slot_parity = FN<P, 32, 8, true, WARPS>(__VA_ARGS__); break;                  \
  }

template <class P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_mega_tasks(const __grid_constant__ MegaParams m, uint32_t slot_bytes) {
  typedef typename P::S S;
  extern __shared__ __align__(128) uint8_t smem[];
  WarpCtx<S> ctx = setup_cta<S>(m.image, m.image_bytes, m.cls[0].ph2pr, WARPS, slot_bytes, (uint32_t)m.n_classes);
  int* const ends_s = reinterpret_cast<int*>(smem + ctx.ends);
  for (int i = threadIdx.x; i < m.n_classes; i += blockDim.x) ends_s[i] = m.task_end[i];
  __syncthreads();
  const unsigned int total = (unsigned)ends_s[m.n_classes - 1];
  for (;;) {
    unsigned int task = 0;
    if (ctx.lane == 0) task = atomicAdd(m.queue, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= total) break;
    const int c = class_of_task(ends_s, m.n_classes, task);
    const unsigned int local = task - (c ? (unsigned)ends_s[c - 1] : 0u);
    const int cfg = m.cfg[c];
    const SweepParams& p = m.cls[c];
    bind_panel(ctx, p.panel);
    GKLB_MEGA_DISPATCH(mega_task, p, local, ctx)
  }
}

template <class P, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_mega_list(const __grid_constant__ MegaParams m, uint32_t slot_bytes) {
  typedef typename P::S S;
  extern __shared__ __align__(128) uint8_t smem[];
  // warp-items per class follow from the list lengths the fp32 kernel left in device memory
  unsigned int total;
  {
    const SmemLayout lay = smem_layout(WARPS, m.image_bytes, slot_bytes, sizeof(S), (uint32_t)m.n_classes);
    int* ends = reinterpret_cast<int*>(smem + lay.ends);
    for (int c = threadIdx.x; c < m.n_classes; c += blockDim.x) {
      const unsigned int gpw = (m.cfg[c] <= 4) ? 4u : (m.cfg[c] <= 8) ? 2u : 1u;
      ends[c] = (int)((*m.cls[c].list_count + gpw - 1) / gpw);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = 0;
      for (int c = 0; c < m.n_classes; c++) { t += (unsigned)ends[c]; ends[c] = (int)t; }
    }
    __syncthreads();
    total = (unsigned)ends[m.n_classes - 1];
  }
  if (total == 0) return;
  WarpCtx<S> ctx = setup_cta<S>(m.image, m.image_bytes, m.cls[0].ph2pr, WARPS, slot_bytes, (uint32_t)m.n_classes);
  const int* const ends_s = reinterpret_cast<const int*>(smem + ctx.ends);
  for (;;) {
    unsigned int wi = 0;
    if (ctx.lane == 0) wi = atomicAdd(m.queue, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= total) break;
    const int c = class_of_task(ends_s, m.n_classes, wi);
    const unsigned int local = wi - (c ? (unsigned)ends_s[c - 1] : 0u);
    const int cfg = m.cfg[c];
    const SweepParams& p = m.cls[c];
    const unsigned int n_items = *p.list_count;
    bind_panel(ctx, p.panel);
    GKLB_MEGA_DISPATCH(mega_item, p, local, n_items, ctx)
  }
}

// ------------------------------------------------------------------------------------------
struct PackClass {
  uint8_t* records;         // [n_rec][5][stride]
  const int32_t* rec_rid;   // [n_rec]
  int rec_begin;            // index of the class's first record in the launch's flat record numbering
  int n_rec;
  int rows;
  int stride;
};

struct PackParams {         // one launch packs the records of every class of a batch
  const int64_t* read_off;
  const uint8_t* bases;
  const uint8_t* quals;
  const uint8_t* ins;
  const uint8_t* del;
  const uint8_t* gcp;
  int n_classes;
  int n_rec_total;
  PackClass cls[32];
};

}  // namespace gklb
