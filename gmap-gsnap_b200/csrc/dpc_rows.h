/* dpc_rows.h -- banded Gotoh fill as a ROW sweep, one warp per matrix, and the lane-parallel traceback walk.
 *
 * Lane l owns the CPL adjacent diagonals k = CPL*l + j (k = c - r + lband) for the whole matrix.  Going from row r-1
 * to row r along a diagonal is the recurrence's diagonal move, so
 *   nogap[r][c]  needs the lane's OWN three values of the previous row            (no communication),
 *   gap2 [r][c]  needs nogap/gap2 of (r-1, c)   = diagonal k+1 of the previous row (one shuffle down),
 *   gap1 [r][c]  needs nogap/gap1 of (r, c-1)   = diagonal k-1 of the SAME row: the serial chain
 *                gap1[k] = max(nogap[k-1] + open, gap1[k-1]) + extend unrolls to
 *                gap1[k] = k*extend + max_{j<k} (nogap[j] + open - j*extend),
 *                an exclusive prefix maximum over the lanes: five shuffle+max steps per 32 diagonals.
 * Every lane works on every row (no wavefront ramp, no idle half-warp, no per-lane row bookkeeping) and the
 * control flow is uniform.  Direction bits leave the warp as four ballots per owned diagonal (bit planes) -- 4 bits
 * per cell, written by one 16-byte store per row and owned diagonal.
 *
 * Exactness: max is exact, so the prefix maximum yields exactly the reference's gap1 values; the direction
 * of gap1[k+1] is better(gap1[k], nogap[k] + open), which lane k can evaluate by itself (plane 2 therefore
 * holds the bit of cell k+1 at position k).  Cells left of column 1 / outside the matrix feed the scan with
 * NEG - k*extend, which reproduces the reference's forced NEG cell (1507-1513) and column 0 (1477-1488).
 *
 * Written against dpc_vec.h, so tests/emul runs the same source as a lock-step 32-lane simulation.
 * Reference recurrence: compute_scores_lookup_fwd/_rev/_fwd_12/_rev_12, dynprog.c:1424-2044.
 */
#ifndef DPC_ROWS_H
#define DPC_ROWS_H

#include "dpc_core.h"
#include "dpc_vec.h"

#ifdef __CUDACC__
#define DPC_VFN __device__ __forceinline__
#else
#define DPC_VFN static inline
#endif

/* CPL diagonals per lane (k = CPL*lane + j), LATE = jump_late_p of this matrix, EP = end-point search over all
 * rows (find_best_endpoint), NBAND = keep the nogap band for a bridge, QROWS = rows are the query.
 *
 * No per-cell boundary handling.  A valid cell only reads cells to its left and above, so:
 *  - cells right of column L2 are simply computed (garbage that can only flow right and down, never back into
 *    the matrix; their direction bits and band entries are never read);
 *  - cells left of column 1 come out of the recurrence itself: row 0 starts with N(0,0) = 0 and NEG everywhere
 *    left of it, so N and gap1 stay "NEG-ish" (NEG plus a bounded number of score/penalty terms, always far
 *    below any real score) for every c <= 0, and gap2 of column 0 is max(N(r-1,0) + open, gap2(r-1,0)) + extend
 *    = open + r*extend exactly as 1477-1488 set it (N(0,0) + open for r = 1, the gap2 chain afterwards);
 *  - diagonals beyond the band (k >= W, they exist because a lane owns CPL of them) take NEG instead of k*extend
 *    as their gap1 term, so gap1 -- and through it nogap -- stays NEG-ish there: "NEG above the band" (1501-1507)
 *    for the cells that read them;
 *  - the ends of the warp: lane 31 has no lane above and lane 0 none to its left.  Lane 31 reads its own value
 *    through a shuffle that keeps and pays NEG in its copy of `extend` (RowState::eU); lane 0 takes lane 31's scan
 *    total, which nobody else needs, pushed down by NEG (RowState::send).  No select in the loop.
 * NEG-ish values never win a max against a real one and never tie with one; the direction bits produced from
 * NEG-ish operands belong to cells the traceback cannot reach (it follows real-valued chains) and the bridges
 * only read in-band cells.  Everything a reachable cell can observe is exact. */
template <int CPL> struct RowState {
  vec::VI Np[CPL], G1p[CPL], G2p[CPL], sh[CPL], kE[CPL], okE[CPL], cm1[CPL];
  vec::VM kok[CPL], ebok[CPL];
  vec::VI bs, bk;
  vec::VP colp;                   /* QROWS: colp[r] = code of the column that enters the lane's last diagonal after row r */
  vec::VI eU;                     /* extend for the gap2 move into the lane's last diagonal: + NEG on lane 31 */
  vec::VI send;                   /* NEG on lane 31, else 0: what lane 31 hands to lane 0 in the prefix scan must lose */
  uint32_t b0;                    /* plane 0 (the nogap state did not come from nogap) of the row just swept, first diagonal set */
  int rowq;                       /* score-table row of the NEXT matrix row's character (loaded one row ahead) */
};

/* offset of the score-table row of matrix row r (1-based): query rows index score[q][.], genome rows score[.][g] */
template <bool QROWS> DPC_VFN int dpc_rows_rowq(const Mat &m, int r) {
  const int ch = (int)m.rowch[r - 1];          /* query rows are staged masked to 7 bits */
  return QROWS ? ch << 3 : ch;
}

template <int CPL, bool QROWS>
DPC_VFN void dpc_rows_init(RowState<CPL> &s, const Mat &m, const EndSearch &es) {
  using namespace vec;
  const int L2 = m.L2, lband = m.lband, W = m.W, open = m.open, extend = m.extend;
  const VI lane = lane_index();
  s.bs = splat(es.best.score); s.bk = splat(es.best.key);
  s.rowq = dpc_rows_rowq<QROWS>(m, 1);
  s.colp = vptr(m.colch, lane * CPL + (CPL - 1 - lband));
  s.send = vsel(lane == 31, DPC_NEG, 0);
  /* (r-1, c) of lane 31's last diagonal lies beyond everything the warp owns: NEG, as 1501-1507 force it.  Lane 31
     reads its own value there (a shuffle that keeps) and pays NEG on the way in -- NEG-ish either way */
  s.eU = vsel(lane == 31, extend + DPC_NEG, extend);
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    const VI k = lane * CPL + j;
    s.kok[j] = k < W;
    /* diagonals past the band (k >= W; they exist because a lane owns CPL of them) get NEG here: their gap1 then
       stays NEG-ish instead of inheriting the real prefix maximum of the band, and with it their nogap value */
    s.kE[j] = vsel(s.kok[j], k * extend, DPC_NEG);
    s.okE[j] = open - s.kE[j];
    s.ebok[j] = vand(k >= lband - es.eb, k <= lband + es.eb);
    const VI c0 = k - lband;                              /* column of this diagonal in row 0 */
    s.cm1[j] = vsel(s.kok[j], c0 - 1, 1 << 24);           /* c - 1 = r + cm1; diagonals past the band never become valid */
    /* row 0 (1460-1475): (0,0) nogap 0; (0,c) gap1 = open + c*extend for 1 <= c <= min(rband, L2) */
    s.Np[j] = vsel(c0 == 0, 0, DPC_NEG);
    s.G1p[j] = vsel(vand(vand(c0 >= 1, c0 <= L2), s.kok[j]), open + c0 * extend, DPC_NEG);
    s.G2p[j] = splat(DPC_NEG);
    /* column characters of row 1 (a window that slides one column per row, so it is filled for every
       diagonal, in or out of the band): genome code, or score-table row of the query character */
    const VI ch = load_u8(m.colch, vsel(vlt_u(c0, L2), c0, 0));
    s.sh[j] = QROWS ? ch : ((ch & 127) << 3);
  }
  if (CPL == 1) s.okE[0] = s.okE[0] + s.send;             /* one diagonal per lane: the scan feed is the lane total */
}

/* one row of one matrix */
template <int CPL, bool LATE, bool EP, bool NBAND, bool QROWS>
DPC_VFN void dpc_rows_step(RowState<CPL> &s, const Mat &m, const int8_t *score, const int r) {
  using namespace vec;
  const int L1 = m.L1, L2 = m.L2, lband = m.lband, W = m.W, open = m.open, extend = m.extend;
  const VI lane = lane_index();
  /* this row's scores: one byte load per cell from the 8-byte table row of (row character, column code) -- the
     load/store pipe has room, the integer pipe that a shift-and-mask extract would use does not */
  const int8_t *srow = score + s.rowq;
  s.rowq = dpc_rows_rowq<QROWS>(m, r + 1);                /* rowch has two bytes of padding past the last row */
  (void)L1;
  /* (r-1, c) of a lane's last diagonal is the first diagonal of the lane above.  That lane evaluates the gap2 move
     (1532-1542) from its own registers and hands over the maximum -- one shuffle instead of two; its ballot carries
     the direction bit one lane too high */
  VM geU;
  const VI xU = LATE ? vmax_ge(s.G2p[0], s.Np[0] + open, geU) : vmax_ge(s.Np[0] + open, s.G2p[0], geU);
  const VI upX = shfl_down1_keep(xU);
  const uint32_t b3U = vballot(LATE ? geU : vnot(geU)) >> 1;
  VI Nn[CPL], G2n[CPL], sv[CPL], li[CPL];
  VM p1[CPL], p2[CPL], pv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    /* nogap, 1545-1561: one three-way maximum; the direction follows from which operands attain it.  Without
       jump_late_p the earlier candidate wins a tie (DIAG, then HORIZ), with it the later one (VERT, then HORIZ).
       Plane 0 = the nogap state did NOT come from nogap (the traceback turns here), plane 1 = it came from gap2. */
    const VI best = vmax3(s.Np[j], s.G1p[j], s.G2p[j]);
    if (LATE) { p2[j] = s.G2p[j] == best; p1[j] = vor(p2[j], s.G1p[j] == best); }
    else { p1[j] = s.Np[j] != best; p2[j] = vand(p1[j], s.G1p[j] != best); }
    Nn[j] = best + load_i8(srow, s.sh[j]);
    /* gap2, 1532-1542 */
    if (j + 1 < CPL) {
      const VI Nu = s.Np[j + 1 < CPL ? j + 1 : j], G2u = s.G2p[j + 1 < CPL ? j + 1 : j];
      const VI a = Nu + open;
      VM ge;
      VI g2m;
      if (LATE) { g2m = vmax_ge(G2u, a, ge); pv[j] = ge; } else { g2m = vmax_ge(a, G2u, ge); pv[j] = vnot(ge); }
      G2n[j] = g2m + extend;
    } else {
      G2n[j] = upX + s.eU;
      pv[j] = geU;                                       /* not used: plane 3 of this diagonal is b3U */
    }
    /* gap1 feed: nogap + open - k*extend, running maximum inside the lane */
    sv[j] = Nn[j] + s.okE[j];
    li[j] = j == 0 ? sv[j] : vmax(li[j > 0 ? j - 1 : 0], sv[j]);
  }
  /* exclusive prefix maximum of the lane totals (1519-1529 unrolled along the row).  Lane 0 has nothing to its left:
     it takes lane 31's total, which nobody needs, pushed down by NEG (the forced NEG cell of 1507-1513) */
  VI t = shfl_rot1(CPL == 1 ? li[0] : li[CPL - 1] + s.send);
  t = vmax(t, shfl_up_keep(t, 1));
  t = vmax(t, shfl_up_keep(t, 2));
  t = vmax(t, shfl_up_keep(t, 4));
  t = vmax(t, shfl_up_keep(t, 8));
  t = vmax(t, shfl_up_keep(t, 16));
  /* next row's column characters: every diagonal moves one column to the right */
  if (QROWS) {
    /* every lane reads the code that enters its last diagonal; the staged columns carry filler on both sides
       (MatDims::padL / padR), so no index is clamped */
#pragma unroll
    for (int j = 0; j + 1 < CPL; j++) s.sh[j] = s.sh[j + 1];
    s.sh[CPL - 1] = load_u8p(s.colp, r);
  } else {
    /* column index entering at the last diagonal: r + 32*CPL - lband - 1 >= 1 because lband < W <= 32*CPL;
       past the matrix the staged sentinel colch[L2] is read (those diagonals are out of the matrix anyway) */
    const int gi = r + 32 * CPL - lband - 1;
    const int chn = (int)m.colch[gi < L2 ? gi : L2];
    const VI nxt = shfl_down1(s.sh[0], (chn & 127) << 3);
#pragma unroll
    for (int j = 0; j + 1 < CPL; j++) s.sh[j] = s.sh[j + 1];
    s.sh[CPL - 1] = nxt;
  }
#pragma unroll
  for (int j = 0; j < CPL; j++) {
    /* gap1 = k*extend + prefix maximum; it came from gap1 (HORIZ) when that beats nogap + open of the same cell:
       compared before k*extend is added to both sides */
    const VI tm = j == 0 ? t : vmax(t, li[j > 0 ? j - 1 : 0]);
    const VI G1n = tm + s.kE[j];
    const VM h = LATE ? (tm >= sv[j]) : (tm > sv[j]);
    /* directions: four ballots, one 16-byte store */
    const uint32_t b0 = vballot(p1[j]), b1 = vballot(p2[j]), b2 = vballot(h), b3 = j + 1 < CPL ? vballot(pv[j]) : b3U;
    store4_lane0(m.dir + ((r - 1) * CPL + j) * 4, b0, b1, b2, b3);
    if (j == 0) s.b0 = b0;
    if (NBAND) store_i16(m.nband, lane * CPL + ((r - 1) * W + j), vmax(Nn[j], -32768), s.kok[j]);
    if (EP) {
      /* find_best_endpoint (2235-2290) scans rows upwards and columns upwards; a lane meets the cells of its
         diagonals in that order, so "first best" is a strict comparison and, with jump_late_p, "last best" a >=.
         Ties between lanes are settled by the scan-order key in dpc_rows_finish. */
      const VI x = s.cm1[j] + r;
      const VM take = vand(vand(vlt_u(x, L2), s.ebok[j]), LATE ? (Nn[j] >= s.bs) : (Nn[j] > s.bs));
      s.bs = vsel(take, Nn[j], s.bs);
      s.bk = vsel(take, x + (r * (L2 + 1) + 1), s.bk);
    }
    s.Np[j] = Nn[j];
    s.G1p[j] = G1n;
    s.G2p[j] = G2n[j];
  }
}

template <int CPL, bool LATE>
DPC_VFN void dpc_rows_finish(RowState<CPL> &s, const Mat &m, EndSearch &es) {
  using namespace vec;
  const int L1 = m.L1, L2 = m.L2;
  if (es.mode == 3) {
    /* the corner (L1, L2) (4541): one diagonal, owned by one lane -- no search, no reduction */
    const int k = L2 - L1 + m.lband;
    VI v = s.Np[0];
#pragma unroll
    for (int j = 1; j < CPL; j++) v = vsel(splat(k % CPL) == j, s.Np[j], v);
    es.best.score = extract(v, k / CPL);
    es.best.key = L1 * (L2 + 1) + L2;
    return;
  }
  if (es.mode == 2) {
    /* last row: best of the band (2293-2355) */
#pragma unroll
    for (int j = 0; j < CPL; j++) {
      const VI xl = s.cm1[j] + L1;
      keep_better(s.bs, s.bk, s.Np[j], xl + (L1 * (L2 + 1) + 1), vlt_u(xl, L2), LATE);
    }
  }
  reduce_better(s.bs, s.bk, LATE, &es.best.score, &es.best.key);
}

template <int CPL, bool LATE, bool EP, bool NBAND, bool QROWS>
DPC_VFN void dpc_fill_rows(const Mat &m, const int8_t *score, EndSearch &es) {
  RowState<CPL> s;
  dpc_rows_init<CPL, QROWS>(s, m, es);
  int r = 1;
  /* two rows per trip: the loop-carried state then stays in place instead of being copied at the end of every row */
  for (; r < m.L1; r += 2) {
    dpc_rows_step<CPL, LATE, EP, NBAND, QROWS>(s, m, score, r);
    dpc_rows_step<CPL, LATE, EP, NBAND, QROWS>(s, m, score, r + 1);
  }
  if (r == m.L1) dpc_rows_step<CPL, LATE, EP, NBAND, QROWS>(s, m, score, r);
  dpc_rows_finish<CPL, LATE>(s, m, es);
  vec::sync();
}

/* The two matrices of a genome / cDNA gap have the same rows: sweeping them in ONE loop gives the scheduler two
 * independent dependency chains per iteration (the fill is latency-bound at the 16 warps per SM this kernel has).
 * mA runs with LATE, mB with !LATE (4965-4987, 4683-4694). */
template <int CPL, bool LATE, bool QROWS>
DPC_VFN void dpc_fill_rows2(const Mat &mA, const Mat &mB, const int8_t *score, EndSearch &es) {
  RowState<CPL> a, b;
  dpc_rows_init<CPL, QROWS>(a, mA, es);
  dpc_rows_init<CPL, QROWS>(b, mB, es);
  for (int r = 1; r <= mA.L1; r++) {
    dpc_rows_step<CPL, LATE, false, true, QROWS>(a, mA, score, r);
    dpc_rows_step<CPL, !LATE, false, true, QROWS>(b, mB, score, r);
  }
  vec::sync();
}

/* Genome gap, common case (ArenaLayout::fused): L is swept first and keeps its nogap band; then R is swept and the
 * intron bridge (bridge_intron_gap, default mode, dynprog.c:3698-3827) is evaluated row by row inside that sweep --
 * row rR of R against row rL = length1 - rR of the stored L band -- so R's band never exists in memory and the bridge
 * has no pass of its own.  Lane = diagonal on both sides: one left-scan and one right-scan candidate per lane and
 * row pair.  The reference walks rL upwards and keeps the first best; here rL comes downwards, so every candidate
 * carries its scan-order key rL * 8192 + position and ties go to the smaller key.  mL runs with LATE, mR with !LATE. */
template <bool LATE>
DPC_VFN void dpc_fill_rows_bridge(const Mat &mL, const Mat &mR, const int8_t *score, const DevProb &p,
                                  const uint8_t *ldi, const uint8_t *rdi, const int8_t *itab, Best &best) {
  using namespace vec;
  EndSearch none; none.mode = 0; none.eb = 0; none.best.score = 0; none.best.key = 0;
  dpc_fill_rows<1, LATE, false, true, true>(mL, score, none);
  RowState<1> s;
  dpc_rows_init<1, true>(s, mR, none);
  const int L1 = mL.L1, L2L = mL.L2, L2R = mR.L2, eb = p.extraband, gap = p.gap;
  const int rbandL = L2L - L1 + eb, lbandL = eb, lbandR = eb;      /* 3545-3549 */
  const VI lane = lane_index();
  /* lanes that own a diagonal of the band: c <= r + rband, and c >= r - lband holds for every lane.  c - 1 = x + r;
     a lane without a diagonal gets an x that never passes the range test below */
  const VI xL = vsel(lane < mL.W, lane - (mL.lband + 1), 1 << 24), xR = vsel(lane < mR.W, lane - (mR.lband + 1), 1 << 24);
  /* per-lane views that slide one position per row: the dinucleotide codes of column c = lane + r - lband (the
     arrays carry the same filler as the staged columns, MatDims::padL / padR) and the lane's diagonal of L's band */
  const VP ldiP = vptr(ldi, lane - mL.lband), rdiP = vptr(rdi, lane - mR.lband);
  const VPS bandL = vptr16(mL.nband, vmin(lane, mL.W - 1));
  /* Per lane the candidates come in DESCENDING scan order -- rL downwards, right scan before left scan -- so "the
     first best in scan order" is "the last one that is at least as good": one >= per candidate, no key compare.
     A lane remembers where its best came from as rL * 2 + (1 for the right scan); the scan-order key is made
     from that once, at the end. */
  VI bs = splat(best.score), bw = splat(-1);
  int offL = (L1 - 1) * mL.W;                  /* row rL of L's band starts at (rL - 1) * W */
  for (int rR = 1; rR < L1; rR++) {            /* row length1 of either matrix is never looked at (3700, 5013) */
    const int rL = L1 - rR;
    offL -= mL.W;
    /* the stored row of L first, so that its latency hides behind the sweep of R's row */
    const VI vL = load_i16p(bandL, offL);
    const int dL = mL.nband[offL + mL.lband];                     /* (rL, rL) */
    const uint32_t hL = mL.dir[(rL - 1) * 4];
    const int diR = rdi[rR], diL = ldi[rL];
    dpc_rows_step<1, !LATE, false, false, true>(s, mR, score, rR);
    const VI vR = vmax(s.Np[0], -32768);
    const int dR = extract(vR, mR.lband);                         /* (rR, rR) */
    const uint32_t hR = s.b0;
    /* 1 <= c <= length2 - 1 and c < rightoffset - leftoffset - (the other side's column): one unsigned compare */
    int upL = gap - rR - 1, upR = gap - rL - 1;
    upL = upL < L2L - 1 ? upL : L2L - 1; upR = upR < L2R - 1 ? upR : L2R - 1;
    upL = upL < 0 ? 0 : upL; upR = upR < 0 ? 0 : upR;
    {   /* right scan (3768-3816): cR over the band of row rR, cL = rL; -1 when R's cell was entered through a gap */
      const VM ok = vlt_u(xR + rR, upR);
      const VI di = load_u8p(rdiP, rR) & diL;
      const VI sc = vR - ((splat((int)hR) >> lane) & 1) + load_i8(itab, di) + dL;
      const VM take = vand(ok, sc >= bs);
      bs = vsel(take, sc, bs); bw = vsel(take, rL * 2 + 1, bw);
    }
    {   /* left scan (3700-3766): cL over the band of row rL, cR = rR */
      const VM ok = vlt_u(xL + rL, upL);
      const VI di = load_u8p(ldiP, rL) & diR;
      const VI sc = vL - ((splat((int)hL) >> lane) & 1) + load_i8(itab, di) + dR;
      const VM take = vand(ok, sc >= bs);
      bs = vsel(take, sc, bs); bw = vsel(take, rL * 2, bw);
    }
  }
  {
    /* the key of each lane's best: rL * 8192 + position in the row pair's scan (left scan first) */
    const VI rL = bw >> 1, rR = L1 - rL;
    const VI cloL = vmax(rL - lbandL, 1), chighL = vmin(rL + rbandL, L2L - 1), cloR = vmax(rR - lbandR, 1);
    const VI nL = vmax(chighL - cloL + 1, 0);
    const VI cL = lane + (rL - mL.lband), cR = lane + (rR - mR.lband);
    const VI key = vsel((bw & 1) != 0, rL * 8192 + nL + (cR - cloR), rL * 8192 + (cL - cloL));
    reduce_better(bs, vsel(bw < 0, best.key, key), 0, &best.score, &best.key);
  }
  vec::sync();
}

/* Lane-parallel traceback walk over bit-plane directions: 32 cells of the current diagonal (or of the current
 * gap run) are probed at once and a ballot finds where the path turns.  Same ops as dpc_walk_serial. */
DPC_VFN int dpc_walk_planes(const Mat &m, int r0, int c0, int revp, int cdna_direction, uint16_t *ops) {
  using namespace vec;
  const VI lane = lane_index();
  const int cpl4 = m.cpl * 4, cmask = m.cpl - 1, csh = m.cplsh;
  int r = r0, c = c0, run = 0, nops = 0;
  while (dpc_inband(m, r, c)) {
    const int k = c - r + m.lband;
    const VI rr = r - lane;
    const VM inb = vand(rr >= 1, (c - lane) >= 1);
    const VI widx = (vsel(inb, rr, 1) - 1) * cpl4 + ((k & cmask) << 2);
    const VI w0 = load_u32(m.dir, widx), w1 = load_u32(m.dir, widx + 1);
    const VI turn = (w0 >> (k >> csh)) & 1;
    const uint32_t b = vballot(vor(vnot(inb), turn != 0));
    if (b == 0) { run += 32; r -= 32; c -= 32; continue; }
    const int f = first_set(b);
    if (r - f < 1 || c - f < 1) { run += f; r -= f; c -= f; break; }     /* ran into row 0 / column 0: STOP */
    const int d = ((uint32_t)extract(w1, f) >> (k >> csh)) & 1u ? DPC_VERT : DPC_HORIZ;
    run += f + 1; r -= f + 1; c -= f + 1;
    store_u16_lane0(&ops[nops++], (run << 2) | DPC_OP_M); run = 0;
    int dist = 1;
    if (d == DPC_HORIZ) {
      /* gap1 chain along row r, leftwards from column c (2672-2679) */
      for (;;) {
        VM hz;
        if (r == 0) hz = vand(vand((c - lane) >= 2, (c - lane) <= m.rband), (c - lane) <= m.L2);
        else {
          const VI cc = c - lane, kk = cc - r + m.lband;
          const VM in = vand(vand(cc >= 1, kk >= 0), vand(cc <= m.L2, kk < m.W));
          const VI k1 = vsel(in, vmax(kk - 1, 0), 0);
          const VI w = load_u32(m.dir, ((r - 1) * m.cpl + (k1 & cmask)) * 4 + 2);
          hz = vand(in, vor(vor(cc == 1, kk == 0), ((w >> (k1 >> csh)) & 1) != 0));
        }
        const uint32_t nb = vballot(vnot(hz));
        if (nb == 0) { dist += 32; c -= 32; continue; }
        const int g = first_set(nb);
        dist += g; c -= g;
        break;
      }
      c--;
    } else {
      /* gap2 chain up column c from row r (2693-2700) */
      for (;;) {
        VM vt;
        const VI rr2 = r - lane;
        if (c == 0) vt = vand(vand(rr2 >= 2, rr2 <= m.lband), rr2 <= m.L1);
        else {
          const VI kk = c - rr2 + m.lband;
          const VM in = vand(vand(rr2 >= 1, rr2 <= m.L1), vand(kk >= 0, kk < m.W));
          const VI k1 = vsel(in, kk, 0);
          const VI w = load_u32(m.dir, ((vsel(in, rr2, 1) - 1) * m.cpl + (k1 & cmask)) * 4 + 3);
          vt = vand(in, ((w >> (k1 >> csh)) & 1) != 0);
        }
        const uint32_t nb = vballot(vnot(vt));
        if (nb == 0) { dist += 32; r -= 32; continue; }
        const int g = first_set(nb);
        dist += g; r -= g;
        break;
      }
      r--;
    }
    store_u16_lane0(&ops[nops++], (dist << 2) | dpc_run_op(m, d, dist, r, c, revp, cdna_direction));
  }
  if (run) store_u16_lane0(&ops[nops++], (run << 2) | DPC_OP_M);
  sync();
  return nops;
}

/* The product's fill policy: row sweep for bands of up to 128 diagonals, memory-state fill beyond.
 * MAXCPL = 2 compiles only the 1- and 2-diagonals-per-lane sweeps (bands of up to 64 diagonals, which is every
 * band the reference's callers produce and most of the band-30 bench): the host puts wider problems into a launch
 * of the MAXCPL = 4 instantiation, so the common kernel carries neither the 4-diagonal loops nor the memory-state
 * fill and fits a smaller register budget. */
template <int MAXCPL, int KG>
struct RowFillT {
  enum { fillmode = 2, maxcpl = MAXCPL };
  /* KG (kind group of the kernel: 0 single gap, 1 genome gap, 2 cDNA gap, 3 end gaps, -1 any) prunes the variants a
     kernel cannot meet: the single-gap kernel carries one loop per (diagonals per lane, jump_late_p) and nothing else */
  template <int CPL, bool LATE>
  DPC_HDM void go(const Mat &m, const int8_t *score, EndSearch &es) const {
    if ((KG == 2 || KG == -1) && !m.query_rows) dpc_fill_rows<CPL, LATE, false, true, false>(m, score, es);          /* cDNA gap */
    else if ((KG == 1 || KG == -1) && m.nband) dpc_fill_rows<CPL, LATE, false, true, true>(m, score, es);            /* genome gap */
    else if (KG == 1 || KG == 2) return;
    else if (KG != 0 && es.mode == 1) dpc_fill_rows<CPL, LATE, true, false, true>(m, score, es);       /* end gap, best end point */
    else dpc_fill_rows<CPL, LATE, false, false, true>(m, score, es);                        /* single gap, end to query end */
  }
  DPC_HDM void operator()(const Mat &m, int32_t *st, const int8_t *score, EndSearch &es, const Lanes &ln) const {
    if (MAXCPL > 2 && !m.planes) { dpc_fill_generic(m, st, score, es, ln); dpc_warp_best(es.best, m.late, ln); return; }
    if (m.late) {
      if (m.cpl == 1) go<1, true>(m, score, es);
      else if (MAXCPL <= 2 || m.cpl == 2) go<2, true>(m, score, es);
      else go<MAXCPL, true>(m, score, es);
    } else {
      if (m.cpl == 1) go<1, false>(m, score, es);
      else if (MAXCPL <= 2 || m.cpl == 2) go<2, false>(m, score, es);
      else go<MAXCPL, false>(m, score, es);
    }
  }
  /* both matrices of a genome / cDNA gap (mA.late == !mB.late, same rows) */
  DPC_HDM void pair(const Mat &mA, const Mat &mB, int32_t *st, const int8_t *score, EndSearch &es, const Lanes &ln) const {
    if (MAXCPL <= 2 || (mA.planes && mB.planes && mA.cpl == mB.cpl && mA.cpl <= 2 && mA.L1 == mB.L1)) {
      const bool q = mA.query_rows != 0;
      if (mA.cpl == 1 && mB.cpl == 1) {
        if (mA.late) { if (q) dpc_fill_rows2<1, true, true>(mA, mB, score, es); else dpc_fill_rows2<1, true, false>(mA, mB, score, es); }
        else { if (q) dpc_fill_rows2<1, false, true>(mA, mB, score, es); else dpc_fill_rows2<1, false, false>(mA, mB, score, es); }
        return;
      } else if (mA.cpl == 2 && mB.cpl == 2) {
        if (mA.late) { if (q) dpc_fill_rows2<2, true, true>(mA, mB, score, es); else dpc_fill_rows2<2, true, false>(mA, mB, score, es); }
        else { if (q) dpc_fill_rows2<2, false, true>(mA, mB, score, es); else dpc_fill_rows2<2, false, false>(mA, mB, score, es); }
        return;
      }
    }
    (*this)(mA, st, score, es, ln);
    (*this)(mB, st, score, es, ln);
  }
  /* genome gap with ArenaLayout::fused: both sweeps and the intron bridge; leaves the winning candidate in `best` */
  DPC_HDM void pair_bridge(const Mat &mL, const Mat &mR, const int8_t *score, const DevProb &p,
                           const uint8_t *ldi, const uint8_t *rdi, const int8_t *itab, Best &best, const Lanes &ln) const {
    (void)ln;
    if (KG == 1 || KG == -1) {
      if (mL.late) dpc_fill_rows_bridge<true>(mL, mR, score, p, ldi, rdi, itab, best);
      else dpc_fill_rows_bridge<false>(mL, mR, score, p, ldi, rdi, itab, best);
    }
  }
  DPC_HDM int walk(const Mat &m, int r, int c, int revp, int cdna_direction, uint16_t *ops, const Lanes &ln) const {
    if (MAXCPL > 2 && !m.planes) return dpc_walk_serial(m, r, c, revp, cdna_direction, ops, ln);
    return dpc_walk_planes(m, r, c, revp, cdna_direction, ops);
  }
};
typedef RowFillT<DPC_MAX_CPL, -1> RowFill;
/* true when every matrix of the problem runs in the 1- or 2-diagonals-per-lane sweep (RowFillT<2>) */
DPC_HB bool dpc_narrow(const ArenaLayout &a) {
  for (int i = 0; i < a.nmat; i++) if (!a.d[i].planes || a.d[i].cpl > 2) return false;
  return true;
}
#endif /* DPC_ROWS_H */
