// pairhmm_h2.cuh -- the "two haplotypes per lane" forward sweep (policy H2), sm_100a.
//
// Same recurrence and the same systolic mapping as pairhmm_device.cuh (lane t of a group of G lanes
// owns K consecutive read rows, the bottom row moves to lane t+1 by shuffle), but the two halves of
// every packed fp32 register pair hold the SAME read against TWO haplotypes of similar length
// instead of two reads against one haplotype.  What that buys on sm_100:
//   * every per-row constant (pMM, kappa, pXX, pMY', Ax) is now a 32-bit scalar that FFMA2/FMUL2 read
//     through their broadcast operand form (SASS `R.F32`): 5 K registers instead of 10 K, and one
//     32-bit register-file read instead of a 64-bit one for a third of the operands -- the packed-read
//     kernel is bound by register operand bandwidth;
//   * with the registers that frees, a lane can own up to 13 rows: 101-row reads run as 8 lanes x 13
//     rows (3 padding rows, 7-step fill) instead of 16 x 7 (11 padding rows, 15-step fill), and the
//     per-step overhead (shuffles, symbol fetch, loop) is amortised over twice as many cells;
//   * constants and the prior table are set up for one read per lane instead of two.
// The price is the prior: the two haplotypes show different symbols, so it is two LDS.32 per cell pair
// instead of one LDS.64 (the shared-memory bandwidth used is the same).
//
// The panel is a PAIR image (built on the host by engine.cu: haplotypes of a tile sorted by length and
// paired): one byte per column = symbol of haplotype A | symbol of haplotype B << 3.
// Reference semantics: avx-pairhmm-template.h:106-223,325-371 and IntelPairHmm.cc:150-169, as in
// pairhmm_device.cuh (W form of the insertion state, folded Y and diagonal states).
#pragma once

#include "pairhmm_device.cuh"

namespace gklb {

// One haplotype tile as a pair image:
//   int32 ppos[n]            byte offset of column 0 of pair i inside the image
//   int32 lenA[n], lenB[n]   (lenA >= lenB)
//   int32 idxA[n], idxB[n]   haplotype index in its region (idxB -1: B repeats A, drop it)
//   pad to 16, then per pair: left margin | bytes | right margin
// A launch keeps one image resident in shared memory: one tile, or the tiles of several regions back to back.

struct H2Class {           // the reads of one length class of one region against one tile of its haplotypes
  const uint8_t* records;  // [n_rec][5 planes][stride]   (rows = G * K of the class kernel, packed by k_pack_reads)
  const int32_t* rec_rid;  // [n_rec] read index in the batch, -1 for filler records
  const int32_t* rec_len;  // [n_rec]
  uint2* r2_items;         // rerun items: (record, pair index | half mask << 30) of pairs whose scaled sum is
  unsigned int* r2_count;  //   < 1e-28f or not finite (consumed by the range-extended rerun, pairhmm_r2.cuh)
  unsigned int* fb_pairs;  // how many pairs that is (statistics)
  uint2* fb_items;         // use_r2 == 0: the flagged pairs go straight to the fp64 kernel as (record, haplotype)
  unsigned int* fb_count;
  int use_r2;
  int n_rec;               // multiple of 32 / G
  int rows;                // G * K
  int stride;              // bytes per plane
  int pair_chunk;          // haplotype pairs per task
  int n_chunks;
  int n_tasks;             // (n_rec / (32 / G)) * n_chunks
  uint32_t panel_off;      // the tile's pair image inside the resident image
  int n_pairs;
  int n_haps_total;        // H of the region (row pitch of its output)
  double* out;             // the region's likelihoods, biased so that out[rid * n_haps_total + h] is read rid's row
};

struct H2Common {
  const uint8_t* image;    // resident image (bulk-copied to shared memory once per CTA)
  uint32_t image_bytes;
  const float* ph2pr;
  const float* mm;
  float init_const;        // 2^120
  float log10_init;
  uint32_t slot_bytes;
};

struct H2Params {          // single-class launch
  H2Common com;
  H2Class cls;
  unsigned int* task_counter;
};

// Multi-class launch: one queue that concatenates the tasks of every (class, tile) entry, longest class first (see
// k_mega_tasks in pairhmm_device.cuh for why); the entries may belong to different regions (several active regions
// in one launch).  cfg = gi * 9 + (K - 8) with G = 4 << gi.  The arrays live in device memory (meta block).
struct H2MegaParams {
  H2Common com;
  int n_classes;
  unsigned int* queue;
  const int* cfg;
  const int* task_end;
  const H2Class* cls;
};

__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

template <int K>
struct LaneRowsH2 {
  float Am[K], kap[K], pXX[K], pMY[K], Ax[K];
  float gTop, xlast;
  uint32_t padmask;
};

// One read's rows [row0, row0 + K) -> scalar constants + this lane's column of the prior table.
// Same construction as load_lane_rows<.., VAR 5> (pairhmm_device.cuh), one read per lane.
template <int K>
__device__ __forceinline__ void load_lane_rows_h2(LaneRowsH2<K>& L, const uint8_t* rec, int stride, int n_rows, int row0,
                                                  int n_pad, bool top_is_row0, const float* __restrict__ ph2pr,
                                                  const float* __restrict__ mm, float* tbl) {
  uint32_t pm = 0;
#pragma unroll
  for (int j = 0; j < K; j++) {
    const int row = row0 + j;
    const bool pad = row < n_pad;
    const uint32_t nib = rec[row];
    const int q = rec[stride + row], ig = rec[2 * stride + row], dg = rec[3 * stride + row], cg = rec[4 * stride + row];
    const float e = ph2pr[q];
    const float om = 1.0f - e;   // stripeINITIALIZATION: _1_distm = 1 - distm
    const float th = e / 3.0f;   //                       distm = distm / 3
    const float pmm = __ldg(mm + mm_index(ig, dg));
    const float pc = ph2pr[cg];
    const float pmx = ph2pr[ig];
    float am = pmm, my = ph2pr[dg], xx = pc;
    if (pad) { am = 0.0f; my = 1.0f; xx = 1.0f; pm |= 1u << j; }
    const uint32_t symnib[kPriorSyms] = {1u, 2u, 4u, 8u, 15u};
#pragma unroll
    for (int sy = 0; sy < kPriorSyms; sy++) tbl[(sy * K + j) * 32] = pad ? 0.0f : ((nib & symnib[sy]) ? om : th);
    // pGAPM of the row below (0 past the last row and for padding rows, whose M must stay 0)
    float gnext = 0.0f;
    if (row + 1 < n_rows && row + 1 >= n_pad) gnext = 1.0f - ph2pr[rec[4 * stride + row + 1]];
    my *= gnext;
    if (j == 0) L.gTop = pad ? 0.0f : 1.0f - pc;
    // W = X / pMX(row):  W = M(up) + kappa * W(up);  the row above a first real row has X = 0 (kappa = 0)
    float kappa = 0.0f;
    if (!pad && row - 1 >= n_pad) kappa = pc * ph2pr[rec[2 * stride + row - 1]] / pmx;
    const float ax = pad ? 0.0f : gnext * pmx;
    if (j == K - 1) L.xlast = pad ? 0.0f : pmx;
    if (j == 0 && top_is_row0) { am = 0.0f; kappa = 0.0f; }
    L.Am[j] = am; L.kap[j] = kappa; L.pXX[j] = xx; L.pMY[j] = my; L.Ax[j] = ax;
  }
  L.padmask = pm;
}

template <int G, int K>
struct SweeperH2 {
  const LaneRowsH2<K>& L;
  float2 Ml[K], Yl[K], Zl[K];
  float2 botX, sum, sumW;
  float2 uM, uX, uZ, dMp, dZp, inj;
  const uint8_t* hap;
  const float* tb;
  int lenA, lenB, c;
  uint32_t hb;
  bool row0_above;

  __device__ __forceinline__ SweeperH2(const LaneRowsH2<K>& L_) : L(L_) {}

  __device__ __forceinline__ void fetch_up() {
    uM = make_float2(__shfl_up_sync(0xffffffffu, Ml[K - 1].x, 1, G), __shfl_up_sync(0xffffffffu, Ml[K - 1].y, 1, G));
    uX = make_float2(__shfl_up_sync(0xffffffffu, botX.x, 1, G), __shfl_up_sync(0xffffffffu, botX.y, 1, G));
    uZ = make_float2(__shfl_up_sync(0xffffffffu, Zl[K - 1].x, 1, G), __shfl_up_sync(0xffffffffu, Zl[K - 1].y, 1, G));
    if (row0_above) {  // row 0: M = X = 0, Y = init
      uZ = inj;
      uM = make_float2(0.0f, 0.0f);
    }
  }

  template <bool GUARD>
  __device__ __forceinline__ void cells(const float* tA, const float* tB) {
    float2 dM = dMp, dZ = dZp, upM = uM, upX = uX;
#pragma unroll
    for (int j = 0; j < K; j++) {
      const float2 pr = make_float2(tA[j * 32], tB[j * 32]);
      const float2 Xn = __ffma2_rn(bc2(L.kap[j]), upX, upM);              // W form: kappa * W(up) + M(up)
      const float2 Mn = __fmul2_rn(pr, __ffma2_rn(bc2(L.Am[j]), dM, dZ));
      const float2 Yn = __ffma2_rn(bc2(L.pXX[j]), Yl[j], Ml[j]);           // Y / pMY
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
    if (GUARD) {  // past the end of the shorter haplotype only the longer one still accumulates
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
    hb = hap[min(c + 1, lenA + 1)];
    if (!GUARD || (unsigned)(c - 1) < (unsigned)lenA) cells<GUARD>(tA, tB);
    dMp = uM;
    dZp = uZ;
    c++;
    fetch_up();
  }

  // returns (sum of haplotype A, sum of haplotype B) on the last lane of the group
  __device__ __forceinline__ float2 run(const uint8_t* hap_, int lenA_, int lenB_, int steady_end, int n_steps, int t,
                                        float2 initY, const float* tb_) {
    tb = tb_;
    hap = hap_;
    lenA = lenA_;
    lenB = lenB_;
    row0_above = (t == 0);
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
    c = 1 - t;
    hb = hap[max(c, -kHapLeftMargin + 1)];
    fetch_up();
    int s = 1;
    const int pre_end = min(G - 1, n_steps);
    for (; s <= pre_end; s++) step<true>();
#pragma unroll 2
    for (; s <= steady_end; s++) step<false>();
    for (; s <= n_steps; s++) step<true>();
    return __ffma2_rn(bc2(L.xlast), sumW, sum);
  }
};

// Per-CTA setup for the H2 kernels: mbarriers, ph2pr table, resident image (one TMA bulk copy).  The context holds
// OFFSETS into the dynamic shared memory, not pointers: device functions rebuild their pointers from the
// `extern __shared__` symbol, so that the compiler keeps emitting LDS/STS (a generic pointer that crosses the call
// boundary of the multi-class kernel's task functions turns every table load into a generic LD).
struct WarpCtxH2 {
  uint32_t slot_bar;     // offset of this warp's record-slot mbarrier
  uint32_t slot_parity;
  uint32_t slot;         // offset of this warp's slot (records, then the prior table)
  uint32_t ph2pr;        // offset of float[128]
  uint32_t image;        // offset of the resident image
  uint32_t ends;         // offset of int[n_ends] (multi-class launches)
  int warp, lane;
};

__device__ __forceinline__ WarpCtxH2 setup_cta_h2(const H2Common& com, int warps, uint32_t n_ends = 0) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SmemLayout lay = smem_layout(warps, com.image_bytes, com.slot_bytes, sizeof(float), n_ends);
  const float* ph2pr = com.ph2pr;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  float* ph2pr_s = reinterpret_cast<float*>(smem + lay.ph2pr);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 1 + warps; i++) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) ph2pr_s[i] = ph2pr[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], com.image_bytes);
    tma_bulk_g2s(smem + lay.panel, com.image, com.image_bytes, &bars[0]);
  }
  mbar_wait(&bars[0], 0);
  WarpCtxH2 c;
  c.warp = threadIdx.x >> 5;
  c.lane = threadIdx.x & 31;
  c.slot_bar = lay.bars + 8u * (1 + c.warp);
  c.slot_parity = 0;
  c.slot = lay.slots + (uint32_t)c.warp * lay.slot_bytes;
  c.ph2pr = lay.ph2pr;
  c.image = lay.panel;
  c.ends = lay.ends;
  return c;
}

// One task = (block of 32/G records) x (chunk of haplotype pairs), executed by one warp.  Flips ctx.slot_parity.
template <int G, int K>
__device__ __forceinline__ void run_task_h2(const H2Common& p, const H2Class& cls, unsigned int task, WarpCtxH2& ctx) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int GPW = 32 / G;
  const int lane = ctx.lane;
  const int t = lane % G, g = lane / G;
  const uint32_t rec_bytes = 5u * (uint32_t)cls.stride;
  const int blk = task / cls.n_chunks, chunk = task - blk * cls.n_chunks;
  const int rec0 = blk * GPW;
  uint8_t* slot = smem + ctx.slot;
  uint64_t* slot_bar = reinterpret_cast<uint64_t*>(smem + ctx.slot_bar);
  const float* ph2pr_s = reinterpret_cast<const float*>(smem + ctx.ph2pr);
  float* tbs = reinterpret_cast<float*>(slot + ((GPW * rec_bytes + 127u) & ~127u)) + lane;
  __syncwarp();
  if (lane == 0) {
    fence_proxy_async();
    mbar_expect_tx(slot_bar, GPW * rec_bytes);
    tma_bulk_g2s(slot, cls.records + (size_t)rec0 * rec_bytes, GPW * rec_bytes, slot_bar);
  }
  mbar_wait(slot_bar, ctx.slot_parity);
  ctx.slot_parity ^= 1;

  const int rec = rec0 + g;
  const int rid = cls.rec_rid[rec];
  const int npad = cls.rows - cls.rec_len[rec];
  LaneRowsH2<K> L;
  load_lane_rows_h2<K>(L, slot + (size_t)g * rec_bytes, cls.stride, cls.rows, t * K, npad, t == 0, ph2pr_s, p.mm, tbs);
  const uint8_t* panel_s = smem + ctx.image + cls.panel_off;
  const int32_t* ppos = reinterpret_cast<const int32_t*>(panel_s);
  const int32_t* plenA = ppos + cls.n_pairs;
  const int32_t* plenB = plenA + cls.n_pairs;
  const int32_t* pidxA = plenB + cls.n_pairs;
  const int32_t* pidxB = pidxA + cls.n_pairs;
  const int q_begin = chunk * cls.pair_chunk, q_end = min(cls.n_pairs, q_begin + cls.pair_chunk);
  double* const out = cls.out;
  const int n_haps_total = cls.n_haps_total;
  for (int q = q_begin; q < q_end; q++) {
    const int lenA = plenA[q], lenB = plenB[q];
    const uint8_t* hap = panel_s + ppos[q];
    const float2 initY = make_float2(p.init_const / (float)lenA, p.init_const / (float)max(lenB, 1));
    SweeperH2<G, K> sw(L);
    // steps G..min(lenA, lenB) need no guard: every lane is inside both haplotypes
    const float2 sum = sw.run(hap, lenA, lenB, min(lenA, lenB), lenA + G - 1, t, initY, tbs);
    if (t == G - 1 && rid >= 0) {
      unsigned int mask = 0;
#pragma unroll
      for (int x = 0; x < 2; x++) {
        const int h = x == 0 ? pidxA[q] : pidxB[q];
        if (h < 0) continue;   // an odd haplotype out is paired with itself; its second result is dropped
        double* o = out + (size_t)rid * n_haps_total + h;
        if (!finish_pair<VF1>(x == 0 ? sum.x : sum.y, (double)p.log10_init, o)) {
          *o = __longlong_as_double(0x7ff8000000000000LL);  // overwritten by the rerun
          mask |= 1u << x;
        }
      }
      if (mask) {
        atomicAdd(cls.fb_pairs, (unsigned)__popc(mask));
        if (cls.use_r2) {
          const unsigned int k = atomicAdd(cls.r2_count, 1u);
          cls.r2_items[k] = make_uint2((unsigned)rec, (unsigned)q | (mask << 30));
        } else {
#pragma unroll
          for (int x = 0; x < 2; x++)
            if (mask & (1u << x)) {
              const unsigned int k = atomicAdd(cls.fb_count, 1u);
              cls.fb_items[k] = make_uint2((unsigned)rec, (unsigned)(x == 0 ? pidxA[q] : pidxB[q]));
            }
        }
      }
    }
  }
}

template <int G, int K, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_h2_tasks(const __grid_constant__ H2Params p) {
  WarpCtxH2 ctx = setup_cta_h2(p.com, WARPS);
  for (;;) {
    unsigned int task = 0;
    if (ctx.lane == 0) task = atomicAdd(p.task_counter, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= (unsigned)p.cls.n_tasks) break;
    run_task_h2<G, K>(p.com, p.cls, task, ctx);
  }
}

// The task functions of the multi-class kernel take everything by value (no pointer to the caller's frame).
template <int G, int K>
__device__ __noinline__ void mega_task_h2(const H2MegaParams& m, int c, unsigned int task, WarpCtxH2 ctx) {
  run_task_h2<G, K>(m.com, m.cls[c], task, ctx);
}

#define GKLB_H2_ROW(G, B)                                           \
  case B + 0: mega_task_h2<G, 8>(m, c, local, ctx); break;          \
  case B + 1: mega_task_h2<G, 9>(m, c, local, ctx); break;          \
  case B + 2: mega_task_h2<G, 10>(m, c, local, ctx); break;         \
  case B + 3: mega_task_h2<G, 11>(m, c, local, ctx); break;         \
  case B + 4: mega_task_h2<G, 12>(m, c, local, ctx); break;         \
  case B + 5: mega_task_h2<G, 13>(m, c, local, ctx); break;         \
  case B + 6: mega_task_h2<G, 14>(m, c, local, ctx); break;         \
  case B + 7: mega_task_h2<G, 15>(m, c, local, ctx); break;         \
  case B + 8: mega_task_h2<G, 16>(m, c, local, ctx); break;

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k_h2_mega(const __grid_constant__ H2MegaParams m) {
  extern __shared__ __align__(128) uint8_t smem[];
  WarpCtxH2 ctx = setup_cta_h2(m.com, WARPS, (uint32_t)m.n_classes);
  int* ends_s = reinterpret_cast<int*>(smem + ctx.ends);
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
    switch (m.cfg[c]) {
      GKLB_H2_ROW(4, 0)
      GKLB_H2_ROW(8, 9)
      GKLB_H2_ROW(16, 18)
      case 28: mega_task_h2<32, 9>(m, c, local, ctx); break;    // 32 lanes: 288 and 320 rows only
      case 29: mega_task_h2<32, 10>(m, c, local, ctx); break;
      default: break;
    }
    ctx.slot_parity ^= 1;  // every task waits once on the warp's slot barrier
  }
}

}  // namespace gklb
