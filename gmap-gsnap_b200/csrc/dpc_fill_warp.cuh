/* dpc_fill_warp.cuh -- banded Gotoh fill, one warp per matrix, cells in registers.
 *
 * Lane l owns rows l+1, l+33, l+65, ... and walks each of them left to right, one cell per
 * anti-diagonal step d = r + c.  Its left neighbour is its own previous cell; the upper and the
 * diagonal neighbours belong to row r-1, i.e. to lane l-1, which computed them one and two steps
 * earlier -- they arrive through __shfl_sync, never through memory.  A lane is done with row r
 * (d > r + min(L2, r + rband)) before row r+32 starts (d = r + 32 + max(1, r + 32 - lband)) as
 * long as lband + rband < 64, which holds for every band the reference's callers use
 * (extraband 3 / 7 / 30 plus the length difference); wider bands take dpc_fill_generic.
 *
 * Writes the same outputs as dpc_fill_generic: direction nibbles (8 per 32-bit word, one row per
 * lane, accumulated in a register and stored once per word), the nogap band for the bridges, and
 * the lane-local end-point candidate.
 *
 * Reference recurrence: compute_scores_lookup_fwd/_rev/_fwd_12/_rev_12, dynprog.c:1424-2044.
 */
#ifndef DPC_FILL_WARP_CUH
#define DPC_FILL_WARP_CUH

#include "dpc_core.h"

__device__ __forceinline__ void dpc_fill_warp(const Mat &m, const int8_t *score, EndSearch &es, const Lanes &ln) {
  const int L1 = m.L1, L2 = m.L2, lband = m.lband, rband = m.rband;
  const int open = m.open, extend = m.extend, late = m.late;
  const unsigned up_lane = (unsigned)((ln.lane + 31) & 31);
  int r = ln.lane + 1;                 /* current row of this lane */
  int clo = 1, chi = 0;                /* its column range */
  uint32_t prof = 0;                   /* query rows: 6 signed 4-bit scores of this row's query character */
  int rowg = 0;                        /* genome rows: this row's genome code */
  int N = DPC_NEG, G1 = DPC_NEG, G2 = DPC_NEG;
  int dgN = DPC_NEG, dgG1 = DPC_NEG, dgG2 = DPC_NEG;
  uint32_t acc = 0;
  const int wstride8 = m.wstride << 3;

#define DPC_LOAD_ROW()                                                                              \
  do {                                                                                              \
    clo = r - lband < 1 ? 1 : r - lband;                                                            \
    chi = r + rband > L2 ? L2 : r + rband;                                                          \
    N = DPC_NEG; G1 = DPC_NEG; /* column 0 (1477-1488) or the forced cell left of the band (1507-1513) */ \
    acc = 0;                                                                                        \
    if (m.query_rows) {                                                                             \
      const int8_t *s_ = score + (m.rowch[r - 1] & 127) * 8;                                        \
      prof = 0;                                                                                     \
      for (int g_ = 0; g_ < 6; g_++) prof |= ((uint32_t)s_[g_] & 15u) << (4 * g_);                  \
    } else rowg = m.rowch[r - 1];                                                                   \
  } while (0)

  if (r <= L1) DPC_LOAD_ROW();
  for (int d = 2; d <= L1 + L2; d++) {
    /* (1) row r-1's newest cell, computed by the upper lane in the previous step */
    const int tN = __shfl_sync(0xffffffffu, N, up_lane);
    const int tG1 = __shfl_sync(0xffffffffu, G1, up_lane);
    const int tG2 = __shfl_sync(0xffffffffu, G2, up_lane);
    /* (2) next row once this one is finished */
    while (r <= L1 && d > r + chi) {
      r += 32;
      if (r <= L1) DPC_LOAD_ROW();
    }
    /* (3) the cell of this step */
    const int c = d - r;
    if (r <= L1 && c >= clo && c <= chi) {
      const int k = c - r;
      int Nu, G2u, Nd, G1d, G2d;
      if (r > 1 && k < rband) { Nu = tN; G2u = tG2; }
      else { Nu = DPC_NEG; G2u = DPC_NEG; }        /* row 0 (1464-1475) or forced above the band (1501-1506) */
      if (r == 1) {
        if (c == 1) { Nd = 0; G1d = DPC_NEG; G2d = DPC_NEG; }
        else { Nd = DPC_NEG; G1d = open + (c - 1) * extend; G2d = DPC_NEG; }
      } else if (c == 1) { Nd = DPC_NEG; G1d = DPC_NEG; G2d = open + (r - 1) * extend; }
      else { Nd = dgN; G1d = dgG1; G2d = dgG2; }
      int P;
      if (m.query_rows) P = ((int)(prof << (28 - 4 * m.colch[c - 1]))) >> 28;
      else P = score[(m.colch[c - 1] & 127) * 8 + rowg];
      int nN, nG1, nG2, nib;
      dpc_cell(N, G1, Nu, G2u, Nd, G1d, G2d, P, open, extend, late, &nN, &nG1, &nG2, &nib);
      N = nN; G1 = nG1; G2 = nG2;
      const int kk = k + lband;
      acc |= (uint32_t)nib << ((kk & 7) << 2);
      if ((kk & 7) == 7 || c == chi) { m.dir[((r - 1) * wstride8 + kk) >> 3] = acc; acc = 0; }
      if (m.nband) m.nband[(r - 1) * m.W + kk] = N;
      if (es.mode == 1) {
        if (k >= -es.eb && k <= es.eb && dpc_better(N, r * (L2 + 1) + c, es.best, late)) { es.best.score = N; es.best.key = r * (L2 + 1) + c; }
      } else if (es.mode == 2) {
        if (r == L1 && dpc_better(N, r * (L2 + 1) + c, es.best, late)) { es.best.score = N; es.best.key = r * (L2 + 1) + c; }
      } else if (es.mode == 3) {
        if (r == L1 && c == L2) { es.best.score = N; es.best.key = r * (L2 + 1) + c; }
      }
    }
    dgN = tN; dgG1 = tG1; dgG2 = tG2;
  }
#undef DPC_LOAD_ROW
  __syncwarp();
}

struct WarpFill {
  enum { needs_state = 2 };            /* arena carries anti-diagonal state only for bands >= 64 */
  __device__ __forceinline__ void operator()(const Mat &m, int32_t *st, const int8_t *score, EndSearch &es, const Lanes &ln) const {
    if (m.lband + m.rband < 64) dpc_fill_warp(m, score, es, ln);
    else dpc_fill_generic(m, st, score, es, ln);
  }
};
#endif
