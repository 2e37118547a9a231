// sw_device.cuh -- sm_100a device code of the Smith-Waterman aligner with backtrack.
//
// Reference semantics (/root/reference/src/main/native/smithwaterman/PairWiseSW.h): affine-gap local/global hybrid
// over int32 scores,
//     E(i,j) = max(H(i,j-1) + open, E(i,j-1) + extend)          gap in seq1 ("INSERT", consumes seq2)
//     F(i,j) = max(F(i-1,j) + extend, H(i-1,j) + open)          gap in seq2 ("DELETE", consumes seq1)
//     H(i,j) = max(max(H(i-1,j-1) + s(i,j), MATRIX_MIN_CUTOFF), E, F)   with ties resolved MATCH > INSERT > DELETE
// (MAIN_CODE, :27-62), edges H(i,0) = H(0,j) = 0 or open + (k-1) * extend for the INDEL strategies (:212-221), a
// 4-bit backtrack code per cell (direction | INSERT_EXT | DELETE_EXT), the best cell of the last row / column picked
// in anti-diagonal order with the reference's tie rules (:225-251) and the CIGAR walked back from there (:269-437).
//
// Mapping: one warp per pair, pulled longest-first from an atomic queue.  The warp is a systolic array over seq1:
// lane t owns K consecutive rows (K = 4, 8, 12 or 16, the smallest that covers seq1 in one pass of 32 K rows), keeps
// H and E of the previous column in registers and at step s processes column s - t; the bottom row of lane t-1
// (H, F) arrives by warp shuffle, its previous value is the diagonal.  The backtrack nibbles of a lane's column are
// one or two 32-bit words, stored to a per-warp scratch area in global memory that stays L2-resident; sequences
// longer than 512 rows take several passes with the bottom row of a pass carried through a scratch line.  The choice among equal maxima and the walk back are done by the same warp
// right after the fill: the candidates are scanned 32 at a time with ballots, and the walk reads the backtrack words
// through a 32-column x 16-row register window refreshed by coalesced loads.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gklb {

// Rows per lane: the smallest of 4 / 8 / 12 / 16 whose 32-lane pass covers seq1, else 16 with several passes.
__host__ __device__ inline int sw_rows_per_lane(int nrow) {
  return nrow <= 128 ? 4 : nrow <= 256 ? 8 : nrow <= 384 ? 12 : 16;
}
// 32-bit backtrack words per lane and column (eight 4-bit codes each)
__host__ __device__ inline int sw_words_per_lane(int k) { return (k + 7) / 8; }
// backtrack words one pair needs
__host__ __device__ inline size_t sw_bt_words(int nrow, int ncol) {
  const int k = sw_rows_per_lane(nrow);
  const size_t passes = (size_t)((nrow + 32 * k - 1) / (32 * k));
  return passes * 32 * (size_t)sw_words_per_lane(k) * (size_t)(ncol + 1);
}
constexpr int kSwCutoff = -100000000;            // MATRIX_MIN_CUTOFF, smithwaterman_common.h:81
constexpr int kSwLow = INT32_MIN / 2;            // LOW_INIT_VALUE, :82
// run element: op in the low 4 bits (0 M, 1 I, 2 D, 9 S -- smithwaterman_common.h:42-47), length above
__host__ __device__ inline uint32_t sw_run(int op, int len) { return (uint32_t)op | ((uint32_t)len << 4); }

struct SwParams {
  const uint8_t* seq1;
  const int64_t* off1;
  const uint8_t* seq2;
  const int64_t* off2;
  const int32_t* order;        // pairs, longest first
  int n;
  int match, mismatch, open, extend, strategy;
  // per-warp scratch (global): backtrack words, bottom-row carry, last row / last column, run list
  uint32_t* bt;
  size_t bt_stride;            // words per warp
  int32_t* lines;              // per warp: carry H[2][W], carry F[2][W], lastrow[W], lastcol[R]
  size_t lines_stride;         // int32 per warp
  int line_w;                  // W = max ncol + 1 (padded)
  int line_r;                  // R = max nrow + 1 (padded)
  uint32_t* runs_scratch;      // per warp: line_w + line_r + 4 elements
  // outputs
  uint32_t* runs;              // compact arena
  unsigned int* cursor;        // elements used in `runs`
  int32_t* run_start;          // [n]
  int32_t* run_count;          // [n]
  int32_t* offsets;            // [n]
  unsigned int* queue;
};

__device__ __forceinline__ int sw_edge(bool indel_edges, int k, int open, int extend) {  // H(k,0) = H(0,k), k >= 1
  return indel_edges ? open + (k - 1) * extend : 0;
}

struct SwWarpScratch {
  uint32_t* bt;
  int32_t *carryH, *carryF, *lastrow, *lastcol;
  uint32_t* myruns;
};

// One pair, K rows per lane.  Inlined into the kernel's switch so that the scratch pointers keep their global address
// space (a call boundary turns them into generic pointers).
template <int K>
__device__ __forceinline__ void sw_pair(const SwParams& p, const SwWarpScratch& w, int pair, int lane) {
  const int w_match = p.match, w_mismatch = p.mismatch, w_open = p.open, w_extend = p.extend, strategy = p.strategy;
  const int line_w = p.line_w;
  constexpr int PASS_ROWS = 32 * K;
  constexpr int KW = (K + 7) / 8;
  uint32_t* bt = w.bt;
  int32_t* carryH = w.carryH;
  int32_t* carryF = w.carryF;
  int32_t* lastrow = w.lastrow;
  int32_t* lastcol = w.lastcol;
  uint32_t* myruns = w.myruns;
  const bool indel_edges = (strategy == 10) || (strategy == 11);
  const bool row_candidates = (strategy == 9) || (strategy == 12);
  {
    const uint8_t* s1 = p.seq1 + p.off1[pair];
    const uint8_t* s2 = p.seq2 + p.off2[pair];
    const int nrow = (int)(p.off1[pair + 1] - p.off1[pair]);
    const int ncol = (int)(p.off2[pair + 1] - p.off2[pair]);
    const int n_pass = (nrow + PASS_ROWS - 1) / PASS_ROWS;
    const int btw = ncol + 1;  // words per (pass, lane) line; column c at index c
    __syncwarp();

    // ------------------------------------------------------------------ fill
    for (int pass = 0; pass < n_pass; pass++) {
      const int i0 = pass * PASS_ROWS + lane * K;  // this lane's rows are i0+1 .. i0+K (1-based)
      int Hl[K], El[K];   // H and E of the previous column
      uint32_t r1[K];
#pragma unroll
      for (int j = 0; j < K; j++) {
        const int i = i0 + j + 1;
        Hl[j] = sw_edge(indel_edges, i, w_open, w_extend);     // column 0
        El[j] = kSwLow;                                         // PairWiseSW.h:90-93,223
        r1[j] = (i <= nrow) ? (uint32_t)s1[i - 1] : 0x100u;     // rows past the end: values nobody reads
      }
      const int32_t* cinH = carryH + (size_t)((pass + 1) & 1) * line_w;   // bottom row of the previous pass
      const int32_t* cinF = carryF + (size_t)((pass + 1) & 1) * line_w;
      int32_t* coutH = carryH + (size_t)(pass & 1) * line_w;
      int32_t* coutF = carryF + (size_t)(pass & 1) * line_w;
      const bool writes_carry = (lane == 31) && (pass + 1 < n_pass);
      if (writes_carry) {  // column 0 of the pass's bottom row
        coutH[0] = sw_edge(indel_edges, i0 + K, w_open, w_extend);
        coutF[0] = kSwLow;
      }
      // Row i0, the row above the first lane's rows: row 0 of the matrix (edge values, F = lowInitValue,
      // :86-89,212-222) in the first pass, the carried bottom row afterwards.
      auto row_above_first_lane = [&](int c, int& h, int& f) {
        if (pass == 0) {
          h = (c == 0) ? 0 : sw_edge(indel_edges, c, w_open, w_extend);
          f = kSwLow;
        } else {
          h = __ldcg(cinH + c);
          f = __ldcg(cinF + c);
        }
      };
      int dH = 0;            // H of the row above at the previous column (the diagonal of the lane's first row)
      int tH = 0, tF = 0;    // first lane: the row above at the next column, fetched one step ahead
      if (lane == 0) {
        int f;
        row_above_first_lane(0, dH, f);
        row_above_first_lane(1, tH, tF);
      }
      int fbot = kSwLow;     // F of the lane's bottom row at the column it processed last
      uint32_t* btline = bt + ((size_t)pass * 32 + lane) * KW * btw;   // KW lines of btw words
      const int last_local = nrow - 1 - i0;  // index of row nrow among this lane's rows when in 0..K-1
      const int n_steps = ncol + 31;
      int c = 1 - lane;
      for (int s = 1; s <= n_steps; s++, c++) {
        // bottom row of the lane above at column c: it processed that column one step ago (or, before its first
        // column, still holds column 0 -- which is exactly the diagonal the first column needs)
        int uH = __shfl_up_sync(0xffffffffu, Hl[K - 1], 1);
        int uF = __shfl_up_sync(0xffffffffu, fbot, 1);
        const bool active = (c >= 1) && (c <= ncol);
        if (lane == 0) {
          uH = tH;
          uF = tF;
          if (c + 1 <= ncol) row_above_first_lane(c + 1, tH, tF);
        }
        if (active) {
          const uint32_t b2 = s2[c - 1];
          int hd = dH, hu = uH, fu = uF;
          uint32_t word[KW];
#pragma unroll
          for (int q = 0; q < KW; q++) word[q] = 0;
#pragma unroll
          for (int j = 0; j < K; j++) {
            // MAIN_CODE, PairWiseSW.h:27-62
            const int ext_h = El[j] + w_extend, open_h = Hl[j] + w_open;
            const int e = max(open_h, ext_h);
            uint32_t code = (open_h > ext_h) ? 0u : 4u;           // INSERT_EXT unless opening is strictly better
            const int ext_v = fu + w_extend, open_v = hu + w_open;
            const int f = max(ext_v, open_v);
            code |= (open_v > ext_v) ? 0u : 8u;                   // DELETE_EXT
            int h = max(hd + ((r1[j] == b2) ? w_match : w_mismatch), kSwCutoff);
            if (e > h) { code |= 1u; h = e; }                     // INSERT
            if (f > h) { code = (code & 12u) | 2u; h = f; }       // DELETE
            word[j / 8] |= code << (4 * (j % 8));
            hd = Hl[j];   // H(i, c-1): the diagonal of the row below
            Hl[j] = h;
            El[j] = e;
            hu = h;
            fu = f;
          }
          fbot = fu;
#pragma unroll
          for (int q = 0; q < KW; q++) btline[(size_t)q * btw + c] = word[q];
          if ((unsigned)last_local < (unsigned)K) {
            int v = Hl[0];
#pragma unroll
            for (int j = 1; j < K; j++) v = (j == last_local) ? Hl[j] : v;
            lastrow[c] = v;
          }
          if (writes_carry) { coutH[c] = Hl[K - 1]; coutF[c] = fbot; }
        }
        if (lane != 0 || active) dH = uH;  // this step's top is the next step's diagonal
      }
      // every lane's last active step was column ncol: its registers hold the last column
#pragma unroll
      for (int j = 0; j < K; j++)
        if (i0 + j + 1 <= nrow) lastcol[i0 + j + 1] = Hl[j];
      __syncwarp();
    }

    // ------------------------------------------------------------------ best cell (PairWiseSW.h:225-251)
    int max_i = 0, max_j = 0;
    {
      int best = INT32_MIN;
      for (int k = lane + 1; k <= nrow; k += 32) best = max(best, __ldcg(lastcol + k));
      if (row_candidates)
        for (int k = lane + 1; k <= ncol; k += 32) best = max(best, __ldcg(lastrow + k));
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
      // candidates equal to the maximum, in anti-diagonal order; on one anti-diagonal the last-row cell comes first
      bool have = false;
      for (int ad0 = 1; ad0 <= nrow + ncol; ad0 += 32) {
        const int ad = ad0 + lane;
        const bool inr = ad <= nrow + ncol;
        const bool a_ok = inr && row_candidates && ad >= nrow + 1 && __ldcg(lastrow + (ad - nrow)) == best;
        const bool b_ok = inr && ad >= ncol + 1 && __ldcg(lastcol + (ad - ncol)) == best;
        uint32_t am = __ballot_sync(0xffffffffu, a_ok), bm = __ballot_sync(0xffffffffu, b_ok);
        uint32_t any = am | bm;
        while (any) {
          const int l = __ffs(any) - 1;
          any &= any - 1;
          const int d = ad0 + l;
          if ((am >> l) & 1u) {
            const int i = nrow, j = d - nrow;
            if (!have || abs(i - j) < abs(max_i - max_j)) { max_i = i; max_j = j; have = true; }
          }
          if ((bm >> l) & 1u) {
            const int i = d - ncol, j = ncol;
            if (!have || max_j == ncol || abs(i - j) <= abs(max_i - max_j)) { max_i = i; max_j = j; have = true; }
          }
        }
      }
    }

    // ------------------------------------------------------------------ walk back (getCIGAR, :269-437)
    int i, j;
    if (strategy == 10) { i = nrow; j = ncol; }
    else if (strategy == 11) { i = max_i; j = ncol; }
    else { i = max_i; j = max_j; }
    int n_runs = 0, cur_op = -1, cur_len = 0;
    auto push = [&](int op, int len) {  // adjacent equal elements merge (:392-409)
      if (op == cur_op) { cur_len += len; return; }
      if (cur_op >= 0) { if (lane == 0) myruns[n_runs] = sw_run(cur_op, cur_len); n_runs++; }
      cur_op = op;
      cur_len = len;
    };
    if (j < ncol) push(9, ncol - j);
    int state = 0;
    // register window: lane l holds the backtrack words of column wc0 - l for the lines wb (w0) and wb - 1 (w1)
    int wc0 = -1, wb = -1;
    uint32_t w0 = 0, w1 = 0;
    while (i > 0 && j > 0) {
      // row i sits in pass g / (32 K), lane (g % (32 K)) / K, row r = g % K of that lane, word r / 8, nibble r % 8;
      // lines are numbered (pass * 32 + lane) * KW + word, which grows with the row
      const int g = i - 1;
      const int r = g % K;
      const int blk = (g / K) * KW + (r >> 3);
      if (!(blk == wb || blk == wb - 1) || j > wc0 || j < wc0 - 31 || wc0 < 0) {
        wb = blk;
        wc0 = j;
        const int col = j - lane;
        w0 = (col >= 1) ? __ldcg(bt + (size_t)blk * btw + col) : 0u;
        w1 = (col >= 1 && blk >= 1) ? __ldcg(bt + (size_t)(blk - 1) * btw + col) : 0u;
      }
      const uint32_t word = __shfl_sync(0xffffffffu, (blk == wb) ? w0 : w1, wc0 - j);
      const int btr = (int)((word >> (4 * (r & 7))) & 15u);
      if (state == 4) { j--; cur_len++; state = btr & 4; }
      else if (state == 8) { i--; cur_len++; state = btr & 8; }
      else {
        switch (btr & 3) {
          case 0: i--; j--; push(0, 1); state = 0; break;
          case 1: j--; push(1, 1); state = btr & 4; break;
          default: i--; push(2, 1); state = btr & 8; break;
        }
      }
    }
    int offset;
    if (strategy == 9) {
      if (j > 0) push(9, j);
      offset = i;
    } else if (strategy == 12) {
      if (j > 0) cur_len += j;  // an element of the last element's type (:371-377), merged with it
      offset = (int)(int16_t)(i - j);
    } else {
      if (i > 0) push(2, i);
      else if (j > 0) push(1, j);
      offset = 0;
    }
    if (cur_op >= 0) { if (lane == 0) myruns[n_runs] = sw_run(cur_op, cur_len); n_runs++; }
    __syncwarp();
    unsigned int base = 0;
    if (lane == 0) base = atomicAdd(p.cursor, (unsigned)n_runs);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int k = lane; k < n_runs; k += 32) p.runs[base + k] = __ldcg(myruns + k);
    if (lane == 0) {
      p.run_start[pair] = (int32_t)base;
      p.run_count[pair] = n_runs;
      p.offsets[pair] = offset;
    }
  }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_smith_waterman(const SwParams p) {
  const int lane = threadIdx.x & 31;
  const size_t wg = (size_t)blockIdx.x * WARPS + (threadIdx.x >> 5);
  int32_t* lines = p.lines + wg * p.lines_stride;
  SwWarpScratch w;
  w.bt = p.bt + wg * p.bt_stride;
  w.carryH = lines;                              // [2][W]
  w.carryF = lines + 2 * (size_t)p.line_w;       // [2][W]
  w.lastrow = lines + 4 * (size_t)p.line_w;
  w.lastcol = lines + 5 * (size_t)p.line_w;
  w.myruns = p.runs_scratch + wg * (size_t)(p.line_w + p.line_r + 4);
  for (;;) {
    unsigned int q = 0;
    if (lane == 0) q = atomicAdd(p.queue, 1u);
    q = __shfl_sync(0xffffffffu, q, 0);
    if (q >= (unsigned)p.n) break;
    const int pair = p.order[q];
    const int nrow = (int)(p.off1[pair + 1] - p.off1[pair]);
    switch (sw_rows_per_lane(nrow)) {
      case 4: sw_pair<4>(p, w, pair, lane); break;
      case 8: sw_pair<8>(p, w, pair, lane); break;
      case 12: sw_pair<12>(p, w, pair, lane); break;
      default: sw_pair<16>(p, w, pair, lane); break;
    }
  }
}

}  // namespace gklb
