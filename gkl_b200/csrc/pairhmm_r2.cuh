// pairhmm_r2.cuh -- range-extended fp32 rerun of the pairs whose plain fp32 sum fell under GKL's threshold
// (SURVEY.md 8(f) N3), sm_100a.
//
// Reference semantics (IntelPairHmm.cc:157-165, pairhmm_common.h:39): when the fp32 forward sum, scaled by 2^120,
// is below 1e-28f the pair is recomputed from scratch in fp64 with a 2^1020 scale and the result is
// log10(sum) - log10(2^1020), i.e. the true log10 likelihood.  What makes fp32 fail there is its RANGE (the
// likelihood of a poorly matching read is 1e-64 .. 1e-600), not its precision: 1e-5 relative on a log10 value of -64
// or less is more than 1e-3 relative on the likelihood itself.  So the rerun stays in packed fp32 (the H2 sweep of
// pairhmm_h2.cuh, twice the rate of the fp64 pipe and half its register cost) and extends the range:
//
//   * every lane (K consecutive read rows) keeps its state scaled by its own power of two, 2^e, per haplotype half;
//     the bottom row travels to the next lane together with e (one more shuffle) and is converted with two exact
//     multiplications by powers of two;
//   * a lane lowers its e when converted inputs come near the top of the fp32 range (checked every step) and
//     raises it when an exact maximum over its whole state (every 4th step) has decayed below 2^64; all rescaling is
//     by powers of two, hence exact;
//   * the result is log10(sum) - (e + 120) log10(2) in double.
//
// What is lost: values more than ~49 decades below the largest value a lane holds at that moment.  Mass that small
// cannot matter if the large value's mass can reach the same cells at a bounded cost, which gap transitions
// guarantee when gap and mismatch penalties are moderate.  Reads outside that regime (any quality byte above the
// limits of `r2_unsafe`), results too close to where the reference's own fp64 run underflows, and non-finite sums
// are passed on to the fp64 kernel (k_sweep_list<VD1>), which remains the exact restatement of the reference.
#pragma once

#include "pairhmm_h2.cuh"

namespace gklb {

// rerun item of the H2 sweep: x = record, y = pair index in the tile | half mask << 30 (1: haplotype A, 2: B)
struct R2Class {
  const uint8_t* records;
  const int32_t* rec_rid;
  const int32_t* rec_len;
  const uint2* items;            // appended by the H2 sweep
  const unsigned int* n_items;
  uint2* fb_items;               // (record, haplotype) for the fp64 kernel
  unsigned int* fb_count;
  int rows, stride;
  uint32_t panel_off;
  int n_pairs;
  int n_haps_total;
  double* out;
  int force_fp64;                // measurement (GKLB_R2=0): pass every item on to the fp64 kernel
  int debug_flags;               // GKLB_R2_DEBUG: 1 no first-input normalisation, 2 no raising, 4 no lowering
};

struct R2Params {                // single-class launch
  H2Common com;
  R2Class cls;
  unsigned int* item_counter;
};

struct R2MegaParams {
  H2Common com;
  int n_classes;
  unsigned int* queue;
  const int* cfg;
  const R2Class* cls;
};

constexpr float kR2Hi = 3.3230699e35f;    // 2^118: converted inputs above this lower the lane's exponent
constexpr float kR2Lo = 1.8446744e19f;    // 2^64: a lane whose whole state is below this raises its exponent
constexpr int kR2RaiseBits = 32, kR2LowerBits = 64, kR2CheckEvery = 4;
constexpr double kR2MinLog10 = -550.0;    // below this the reference's fp64 run is close to its own underflow

// 2^d for d in [-126, 127]
__device__ __forceinline__ float pow2i(int d) { return __int_as_float((d + 127) << 23); }

// Qualities for which the range-extension argument does not hold: the fp64 kernel takes the pair.
__device__ __forceinline__ bool r2_unsafe(int q, int ig, int dg, int cg) { return q > 60 || ig > 60 || dg > 60 || cg > 20; }

template <int G, int K>
struct SweeperR2 {
  LaneRowsH2<K>& L;
  float2 Ml[K], Yl[K], Zl[K];
  float2 botX, sum, sumW;
  float2 uM, uX, uZ, dMp, dZp, inj;
  const uint8_t* hap;
  const float* tb;
  int lenA, lenB, c;
  int eA, eB;            // the lane's state is (true value) * 2^e, per haplotype half
  uint32_t hb;
  int dbg;
  bool row0_above, last;

  __device__ __forceinline__ SweeperR2(LaneRowsH2<K>& L_) : L(L_) {}

  __device__ __forceinline__ void scale_state(float2 g) {
#pragma unroll
    for (int j = 0; j < K; j++) {
      Ml[j] = __fmul2_rn(Ml[j], g);
      Yl[j] = __fmul2_rn(Yl[j], g);
      Zl[j] = __fmul2_rn(Zl[j], g);
    }
    botX = __fmul2_rn(botX, g);
    sum = __fmul2_rn(sum, g);
    sumW = __fmul2_rn(sumW, g);
    dMp = __fmul2_rn(dMp, g);
    dZp = __fmul2_rn(dZp, g);
    inj = __fmul2_rn(inj, g);
  }

  // bottom row of the lane above (with its exponents), converted to this lane's scale
  __device__ __forceinline__ void fetch_up() {
    const float2 rM = make_float2(__shfl_up_sync(0xffffffffu, Ml[K - 1].x, 1, G), __shfl_up_sync(0xffffffffu, Ml[K - 1].y, 1, G));
    const float2 rX = make_float2(__shfl_up_sync(0xffffffffu, botX.x, 1, G), __shfl_up_sync(0xffffffffu, botX.y, 1, G));
    const float2 rZ = make_float2(__shfl_up_sync(0xffffffffu, Zl[K - 1].x, 1, G), __shfl_up_sync(0xffffffffu, Zl[K - 1].y, 1, G));
    const int ep = __shfl_up_sync(0xffffffffu, (eA & 0xffff) | (eB << 16), 1, G);
    if (row0_above) {  // row 0: M = X = 0, Y = init (in this lane's scale)
      uZ = inj;
      uM = make_float2(0.0f, 0.0f);
      uX = make_float2(0.0f, 0.0f);
      return;
    }
    const int upA = (int)(short)(ep & 0xffff), upB = ep >> 16;
    if (c == 1 && L.padmask == 0u && !(dbg & 1)) {
      // First column of a lane without padding rows: its matrices are still zero, so its scale is nearly free (only
      // the column-0 diagonal inputs saved from the previous step may be non-zero, below a padding lane).  Choose it so
      // that the first inputs land in [2^64, 2^96): the bottom row of the lane above may sit far below that lane's own
      // maximum.  (Padding rows hold Y = init: a lane with padding, and every lane above it, stays at e = 0.)
      const float mxA = fmaxf(fmaxf(rM.x, rX.x), rZ.x), mxB = fmaxf(fmaxf(rM.y, rX.y), rZ.y);
      const int xA = ((__float_as_int(mxA) >> 23) & 0xff) - 127, xB = ((__float_as_int(mxB) >> 23) & 0xff) - 127;
      const int nA = upA + (mxA > 0.0f ? ((64 - xA + 31) >> 5) * 32 : 0), nB = upB + (mxB > 0.0f ? ((64 - xB + 31) >> 5) * 32 : 0);
      int sA = max(-512, min(512, nA - eA)), sB = max(-512, min(512, nB - eB));
      while (sA != 0 || sB != 0) {  // exact, in steps a single power of two can express
        const int a = max(-126, min(127, sA)), b = max(-126, min(127, sB));
        scale_state(make_float2(pow2i(a), pow2i(b)));
        sA -= a;
        sB -= b;
      }
      eA += max(-512, min(512, nA - eA));
      eB += max(-512, min(512, nB - eB));
    }
    // runs once unless the converted inputs come too close to the top of the range; bounded: a non-finite input
    // (overflow in the lane above) can never be brought down, the pair then ends non-finite and goes to the fp64 kernel
    for (int tries = 0;; tries++) {
      const int dA = eA - upA, dB = eB - upB;
      const int a1 = max(-126, min(127, dA)), b1 = max(-126, min(127, dB));
      const int a2 = max(-126, min(127, dA - a1)), b2 = max(-126, min(127, dB - b1));
      const float2 f1 = make_float2(pow2i(a1), pow2i(b1)), f2 = make_float2(pow2i(a2), pow2i(b2));
      uM = __fmul2_rn(__fmul2_rn(rM, f1), f2);
      uX = __fmul2_rn(__fmul2_rn(rX, f1), f2);
      uZ = __fmul2_rn(__fmul2_rn(rZ, f1), f2);
      const bool hiA = fmaxf(fmaxf(uM.x, uX.x), uZ.x) > kR2Hi, hiB = fmaxf(fmaxf(uM.y, uX.y), uZ.y) > kR2Hi;
      if (!(hiA || hiB) || (dbg & 4) || tries >= 6) break;
      const float down = pow2i(-kR2LowerBits);
      scale_state(make_float2(hiA ? down : 1.0f, hiB ? down : 1.0f));
      if (hiA) eA -= kR2LowerBits;
      if (hiB) eB -= kR2LowerBits;
    }
  }

  // exact maximum over the lane's state; raise the exponent of a half whose state has decayed
  __device__ __forceinline__ void maintain() {
    // (the running sums only mean something on the last lane: elsewhere they must not hold the scale down)
    float mA = fmaxf(botX.x, fmaxf(uM.x, fmaxf(uX.x, fmaxf(uZ.x, fmaxf(dMp.x, dZp.x)))));
    float mB = fmaxf(botX.y, fmaxf(uM.y, fmaxf(uX.y, fmaxf(uZ.y, fmaxf(dMp.y, dZp.y)))));
    if (last) {
      mA = fmaxf(mA, fmaxf(sum.x, sumW.x));
      mB = fmaxf(mB, fmaxf(sum.y, sumW.y));
    }
#pragma unroll
    for (int j = 0; j < K; j++) {
      mA = fmaxf(mA, fmaxf(Ml[j].x, fmaxf(Yl[j].x, Zl[j].x)));
      mB = fmaxf(mB, fmaxf(Ml[j].y, fmaxf(Yl[j].y, Zl[j].y)));
    }
    const bool loA = mA < kR2Lo && mA > 0.0f, loB = mB < kR2Lo && mB > 0.0f;
    if (!(loA || loB) || row0_above || (dbg & 2)) return;
    float gA = 1.0f, gB = 1.0f;
    const float up = pow2i(kR2RaiseBits);
    if (loA) do { gA *= up; mA *= up; eA += kR2RaiseBits; } while (mA < kR2Lo && gA < 1e27f);
    if (loB) do { gB *= up; mB *= up; eB += kR2RaiseBits; } while (mB < kR2Lo && gB < 1e27f);
    const float2 g = make_float2(gA, gB);
    scale_state(g);
    uM = __fmul2_rn(uM, g);
    uX = __fmul2_rn(uX, g);
    uZ = __fmul2_rn(uZ, g);
  }

  template <bool GUARD>
  __device__ __forceinline__ void cells(const float* tA, const float* tB) {
    float2 dM = dMp, dZ = dZp, upM = uM, upX = uX;
#pragma unroll
    for (int j = 0; j < K; j++) {
      const float2 pr = make_float2(tA[j * 32], tB[j * 32]);
      const float2 Xn = __ffma2_rn(bc2(L.kap[j]), upX, upM);
      const float2 Mn = __fmul2_rn(pr, __ffma2_rn(bc2(L.Am[j]), dM, dZ));
      const float2 Yn = __ffma2_rn(bc2(L.pXX[j]), Yl[j], Ml[j]);
      const float2 Zn = __ffma2_rn(bc2(L.pMY[j]), Yn, __fmul2_rn(bc2(L.Ax[j]), Xn));
      dM = Ml[j];
      dZ = Zl[j];
      Ml[j] = Mn;
      Yl[j] = Yn;
      Zl[j] = Zn;
      upM = Mn;
      upX = Xn;
    }
    botX = upX;
    if (!last) return;  // only the last lane's bottom row is the read's last row
    if (GUARD) {
      const bool b = c <= lenB;
      sum = make_float2(sum.x + upM.x, sum.y + (b ? upM.y : 0.0f));
      sumW = make_float2(sumW.x + upX.x, sumW.y + (b ? upX.y : 0.0f));
    } else {
      sum = __fadd2_rn(sum, upM);
      sumW = __fadd2_rn(sumW, upX);
    }
  }

  template <bool GUARD>
  __device__ __forceinline__ void step() {
    const float* tA = tb + (hb & 7u) * (K * 32);
    const float* tB = tb + ((hb >> 3) & 7u) * (K * 32);
    hb = hap[max(-kHapLeftMargin + 1, min(c + 1, lenA + 1))];
    if (!GUARD || (unsigned)(c - 1) < (unsigned)lenA) cells<GUARD>(tA, tB);
    dMp = uM;
    dZp = uZ;
    c++;
    fetch_up();
    if ((c & (kR2CheckEvery - 1)) == 0) maintain();
  }

  // On the last lane of the group: (sum A, sum B) in the scale 2^eA / 2^eB.
  __device__ __forceinline__ float2 run(const uint8_t* hap_, int lenA_, int lenB_, int steady_end, int n_steps, int t,
                                        float2 initY, const float* tb_, int dbg_) {
    tb = tb_;
    hap = hap_;
    lenA = lenA_;
    lenB = lenB_;
    dbg = dbg_;
    row0_above = (t == 0);
    last = (t == G - 1);
    eA = eB = 0;
    const float2 zero = make_float2(0.0f, 0.0f);
    inj = __fmul2_rn(bc2(L.gTop), initY);
#pragma unroll
    for (int j = 0; j < K; j++) {
      Ml[j] = zero;
      const float2 y0 = (L.padmask & (1u << j)) ? initY : zero;
      Yl[j] = y0;
      Zl[j] = __fmul2_rn(bc2(L.pMY[j]), y0);
    }
    botX = zero;
    sum = zero;
    sumW = zero;
    dMp = zero;
    dZp = row0_above ? inj : zero;
    uM = uX = uZ = zero;
    c = 1 - t;
    hb = hap[max(c, -kHapLeftMargin + 1)];
    fetch_up();
    int s = 1;
    const int pre_end = min(G - 1, n_steps);
    for (; s <= pre_end; s++) step<true>();
    for (; s <= steady_end; s++) step<false>();
    for (; s <= n_steps; s++) step<true>();
    return __ffma2_rn(bc2(L.xlast), sumW, sum);
  }
};

// One warp-item: every group of G lanes takes one rerun item (record, haplotype pair, half mask).
template <int G, int K>
__device__ __forceinline__ void run_item_r2(const H2Common& p, const R2Class& cls, unsigned int wi, unsigned int n_items,
                                            const WarpCtxH2& ctx) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int GPW = 32 / G;
  const int lane = ctx.lane;
  const int t = lane % G, g = lane / G;
  const uint32_t rec_bytes = 5u * (uint32_t)cls.stride;
  const float* ph2pr_s = reinterpret_cast<const float*>(smem + ctx.ph2pr);
  float* tbs = reinterpret_cast<float*>(smem + ctx.slot) + lane;  // the slot holds only the prior table here
  const uint8_t* panel_s = smem + ctx.image + cls.panel_off;
  const int32_t* ppos = reinterpret_cast<const int32_t*>(panel_s);
  const int32_t* plenA = ppos + cls.n_pairs;
  const int32_t* plenB = plenA + cls.n_pairs;
  const int32_t* pidxA = plenB + cls.n_pairs;
  const int32_t* pidxB = pidxA + cls.n_pairs;
  const unsigned int item = wi * GPW + g;
  const bool valid = item < n_items;
  const uint2 it = valid ? cls.items[item] : make_uint2(0u, 0u);
  const int rec = (int)it.x, q = (int)(it.y & 0x3fffffffu);
  const unsigned int mask = valid ? (it.y >> 30) : 0u;
  const int lenA = valid ? plenA[q] : 0, lenB = valid ? plenB[q] : 0;
  const uint8_t* hap = panel_s + ppos[valid ? q : 0];
  const int rid = valid ? cls.rec_rid[rec] : -1;
  const int npad = valid ? cls.rows - cls.rec_len[rec] : 0;
  int n_steps = valid ? lenA + G - 1 : 0, steady_end = valid ? min(lenA, lenB) : 0;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    n_steps = max(n_steps, __shfl_xor_sync(0xffffffffu, n_steps, o));
    steady_end = min(steady_end, __shfl_xor_sync(0xffffffffu, steady_end, o));
  }
  if (n_steps == 0) return;
  const uint8_t* recp = cls.records + (size_t)rec * rec_bytes;
  LaneRowsH2<K> L;
  __syncwarp();  // the previous item's table reads are done
  load_lane_rows_h2<K>(L, recp, cls.stride, cls.rows, t * K, npad, t == 0, ph2pr_s, p.mm, tbs);
  // qualities outside the regime of the range-extension argument -> fp64
  bool unsafe = false;
#pragma unroll
  for (int j = 0; j < K; j++) {
    const int row = t * K + j;
    if (row >= npad)
      unsafe |= r2_unsafe(recp[cls.stride + row], recp[2 * cls.stride + row], recp[3 * cls.stride + row],
                          recp[4 * cls.stride + row]);
  }
  const unsigned int grp_mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (g * G));
  unsafe = (__ballot_sync(0xffffffffu, unsafe) & grp_mask) != 0u;
  __syncwarp();
  const float2 initY = make_float2(p.init_const / (float)max(lenA, 1), p.init_const / (float)max(lenB, 1));
  SweeperR2<G, K> sw(L);
  const float2 s2 = sw.run(hap, lenA, lenB, steady_end, n_steps, t, initY, tbs, cls.debug_flags);
  if (t == G - 1 && rid >= 0) {
#pragma unroll
    for (int x = 0; x < 2; x++) {
      if (!(mask & (1u << x))) continue;
      const int h = x == 0 ? pidxA[q] : pidxB[q];
      if (h < 0) continue;
      const float s = x == 0 ? s2.x : s2.y;
      const int e = x == 0 ? sw.eA : sw.eB;
      // log10(true sum) - log10(2^120) = log10(s) - (e + 120) log10(2), evaluated from the mantissa of s so that the
      // value does not depend on which power of two the lane happened to end with (the other groups of the warp
      // decide how many idle steps, and therefore rescaling checks, follow a pair's last column)
      int x2 = 0;
      const double mant = frexp((double)s, &x2);
      const double lg = log10(mant) + (double)(x2 - e - 120) * 0.30102999566398119521;
      const bool ok = !unsafe && !cls.force_fp64 && s > 0.0f && s < 3.0e38f && lg > kR2MinLog10;
      if (ok) {
        cls.out[(size_t)rid * cls.n_haps_total + h] = lg;
      } else {
        const unsigned int k = atomicAdd(cls.fb_count, 1u);
        cls.fb_items[k] = make_uint2((unsigned)rec, (unsigned)h);
      }
    }
  }
}

template <int G, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_r2_list(const __grid_constant__ R2Params p) {
  constexpr int GPW = 32 / G;
  const unsigned int n_items = *p.cls.n_items;
  if (n_items == 0) return;
  WarpCtxH2 ctx = setup_cta_h2(p.com, WARPS);
  const unsigned int n_warp_items = (n_items + GPW - 1) / GPW;
  for (;;) {
    unsigned int wi = 0;
    if (ctx.lane == 0) wi = atomicAdd(p.item_counter, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= n_warp_items) break;
    run_item_r2<G, K>(p.com, p.cls, wi, n_items, ctx);
  }
}

template <int G, int K>
__device__ __noinline__ void mega_item_r2(const R2MegaParams& m, int c, unsigned int wi, unsigned int n_items, WarpCtxH2 ctx) {
  run_item_r2<G, K>(m.com, m.cls[c], wi, n_items, ctx);
}

#define GKLB_R2_ROW(G, B)                                                  \
  case B + 0: mega_item_r2<G, 8>(m, c, local, n_items, ctx); break;        \
  case B + 1: mega_item_r2<G, 9>(m, c, local, n_items, ctx); break;        \
  case B + 2: mega_item_r2<G, 10>(m, c, local, n_items, ctx); break;       \
  case B + 3: mega_item_r2<G, 11>(m, c, local, n_items, ctx); break;       \
  case B + 4: mega_item_r2<G, 12>(m, c, local, n_items, ctx); break;       \
  case B + 5: mega_item_r2<G, 13>(m, c, local, n_items, ctx); break;       \
  case B + 6: mega_item_r2<G, 14>(m, c, local, n_items, ctx); break;       \
  case B + 7: mega_item_r2<G, 15>(m, c, local, n_items, ctx); break;       \
  case B + 8: mega_item_r2<G, 16>(m, c, local, n_items, ctx); break;

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_r2_mega(const __grid_constant__ R2MegaParams m) {
  extern __shared__ __align__(128) uint8_t smem[];
  // warp-items per entry follow from the list lengths the sweep left in device memory
  unsigned int total;
  {
    const SmemLayout lay = smem_layout(WARPS, m.com.image_bytes, m.com.slot_bytes, sizeof(float), (uint32_t)m.n_classes);
    int* ends = reinterpret_cast<int*>(smem + lay.ends);
    for (int c = threadIdx.x; c < m.n_classes; c += blockDim.x) {
      const unsigned int gpw = 8u >> (m.cfg[c] / 9);  // G = 4 << (cfg / 9)
      ends[c] = (int)((*m.cls[c].n_items + gpw - 1) / gpw);
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
  WarpCtxH2 ctx = setup_cta_h2(m.com, WARPS, (uint32_t)m.n_classes);
  const int* ends_s = reinterpret_cast<const int*>(smem + ctx.ends);
  for (;;) {
    unsigned int wi = 0;
    if (ctx.lane == 0) wi = atomicAdd(m.queue, 1u);
    wi = __shfl_sync(0xffffffffu, wi, 0);
    if (wi >= total) break;
    const int c = class_of_task(ends_s, m.n_classes, wi);
    const unsigned int local = wi - (c ? (unsigned)ends_s[c - 1] : 0u);
    const unsigned int n_items = *m.cls[c].n_items;
    switch (m.cfg[c]) {
      GKLB_R2_ROW(4, 0)
      GKLB_R2_ROW(8, 9)
      GKLB_R2_ROW(16, 18)
      default: break;
    }
  }
}

}  // namespace gklb
