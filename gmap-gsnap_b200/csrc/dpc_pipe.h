/* dpc_pipe.h -- the per-problem routines of the DEVICE pipeline of dpc_solve: everything the host half used to do
 * per problem -- argument checks and early returns, descriptor packing, arena classification, result finalisation
 * and the expansion of traceback ops into Pair records -- restated as device routines, so that a bulk call moves
 * whole arrays over PCIe and the host touches no problem individually.
 *
 *   dpc_prepare_one   dpc_problem_t -> DevProb + launch class       (mirrors Batch::add_impl, dpc_host.h)
 *   dpc_finish_one    DevRes -> dpc_result_t, number of pairs        (mirrors Batch::finalize)
 *   dpc_expand_one    ops + characters -> dpc_pair_t records         (mirrors Batch::rebuild / replay_t)
 *
 * Reference semantics: the entry points of src/dynprog.c (Dynprog_single_gap 4450-4572, Dynprog_genome_gap
 * 4798-5061, Dynprog_cdna_gap 4577-4793, Dynprog_end5_gap 5094-5284, Dynprog_end3_gap 5556-5741) and its traceback
 * (2372-2712); line numbers are cited where a rule is applied.
 *
 * Written in the style of dpc_core.h (a `Lanes` handle instead of warp intrinsics), so tests/emul runs the same source
 * on the CPU.  Problems that need a host hook (known splice sites, probability mode) or the splice-junction solvers
 * do not take this pipeline; dpc_solve routes such calls through the host half (dpc_host.h).
 */
#ifndef DPC_PIPE_H
#define DPC_PIPE_H

#include "../../include/dynprog_cuda.h"
#include "dpc_core.h"
#include "dpc_rows.h"

/* what the prepare step needs to know about the library state */
struct PrepEnv {
  int maxlength1, maxlength2;
  uint64_t genome_nbases;
  int novelsplicingp;
  int fillmode;                    /* 1 memory-state fill everywhere (test hook), 2 row sweep */
  uint32_t class_bytes[6];         /* shared-memory arena of each launch class (dynprog_cuda.cu) */
  uint64_t qbase;                  /* host address of the first byte of the query buffer */
  uint64_t qbytes;
};

struct PrepOut {
  int cls;                         /* launch class * DPC_NKG + kind group */
  int bucket;                      /* work bucket inside the class, 0 = most work */
  uint64_t scratch;                /* HBM scratch bytes */
  uint32_t gout;                   /* bytes of the staged genome characters */
  uint32_t ovf;                    /* worst-case traceback ops beyond the inline slots */
};

#define DPC_NBUCKET 64
#define DPC_NKG 4                  /* kind groups = kernel instantiations: single gap, genome gap, cDNA gap, end gaps */
DPC_HB int dpc_kind_group(int kind) { return kind == DPC_GENOME_GAP ? 1 : kind == DPC_CDNA_GAP ? 2 : kind == DPC_SINGLE_GAP ? 0 : 3; }
#define DPC_PREP_HOST 0
#define DPC_PREP_DEVICE 1

DPC_HB void dpc_result_init(dpc_result_t &r, const dpc_problem_t &p) {
  r.null_list = 1; r.dynprogindex_out = p.dynprogindex;
  r.finalscore = r.nmatches = r.nmismatches = r.nopens = r.nindels = DPC_UNSET;
  r.new_leftgenomepos = r.new_rightgenomepos = r.exonhead = r.introntype = DPC_UNSET;
  r.incompletep = DPC_UNSET; r.npairs = 0; r.reserved = 0;
  r.left_prob = r.right_prob = -1.0;
}
DPC_HB int dpc_bump(int idx) { return idx + (idx > 0 ? 1 : -1); }
DPC_HB int dpc_quality(double defect_rate) { return defect_rate < 0.003 ? 0 : defect_rate < 0.014 ? 1 : 2; }   /* dynprog.h:27-28 */
DPC_HB bool dpc_allstar(const dpc_problem_t &p) {                       /* dynprog.c:415-419 */
  uint32_t pos = p.chroffset + p.chrpos;
  return pos < p.chroffset || pos >= p.chrhigh;
}

/* Where problem `d` runs: launch class, work bucket, HBM scratch, overflow ops (the same rules for the host half's
 * flush and the device pipeline). */
DPC_HB void dpc_classify(const DevProb &d, int fillmode, const uint32_t *class_bytes, PrepOut &o) {
  int k = 0, bucket = DPC_NBUCKET - 1;
  o.scratch = 0; o.ovf = 0;
  if (!((d.kind == DPC_END5_GAP || d.kind == DPC_END3_GAP) && d.endalign == DPC_QUERYEND_NOGAPS)) {
    ArenaLayout a;
    dpc_layout(d, a, fillmode);
    const int wide = dpc_narrow(a) ? 0 : 2;
    if (a.total <= class_bytes[0]) k = wide;
    else if (a.total <= class_bytes[1]) k = wide + 1;
    else if (a.small <= class_bytes[4]) { k = 4; o.scratch = a.bulk; }
    else { k = 5; o.scratch = a.total; }
    uint64_t worst = 0, work = 0;
    for (int m = 0; m < a.nmat; m++) {
      worst += (uint64_t)(a.d[m].rows + a.d[m].cols + 2);
      work += (uint64_t)a.d[m].rows * (uint64_t)a.d[m].cpl;           /* row-sweep iterations x diagonals per lane */
    }
    if (worst > DPC_INLINE_OPS) o.ovf = (uint32_t)worst;
    /* bucket 0 = most work: 8 lane-rows per bucket up to 504, everything longer in bucket 0 */
    int wb = (int)(work >> 3);
    if (wb > DPC_NBUCKET - 1) wb = DPC_NBUCKET - 1;
    bucket = DPC_NBUCKET - 1 - wb;
  }
  o.cls = k * DPC_NKG + dpc_kind_group(d.kind);
  o.bucket = bucket;
}

/* The argument checks and early returns of the five entry points, without hooks.  Returns DPC_PREP_DEVICE (d and o
 * are filled, except for the scratch / gout offsets the caller hands out), DPC_PREP_HOST (r is the complete
 * result), or a negative error code.  qpool: the device copy of the query buffer (for the alphabet check). */
DPC_HD int dpc_prepare_one(const dpc_problem_t &p, const PrepEnv &env, const uint8_t *qpool, DevProb &d, dpc_result_t &r, PrepOut &o) {
  dpc_result_init(r, p);
  d.q0 = d.q1 = 0; d.gbase = p.chroffset + p.chrpos; d.glen = p.genomiclength; d.aux = 0;
  d.L1 = d.L1R = d.L2 = d.L2R = 0; d.off2 = d.off2R = 0; d.gap = 0;
  d.score_threshold = p.score_threshold; d.extraband = p.extraband;
  d.scratch_lo = d.scratch_hi = 0;
  d.open = d.extend = d.reward = 0;
  d.cdna_direction = (int8_t)(p.cdna_direction > 0 ? 1 : p.cdna_direction < 0 ? -1 : 0);
  d.kind = (uint8_t)p.kind; d.endalign = (uint8_t)p.endalign; d.type = 0; d.pad = 0;
  d.flags = (p.watsonp ? DPC_F_WATSON : 0) | (p.jump_late_p ? DPC_F_LATE : 0) | (p.widebandp ? DPC_F_WIDEBAND : 0) |
            (p.halfp ? DPC_F_HALFP : 0) | (p.finalp ? DPC_F_FINALP : 0) | (dpc_allstar(p) ? DPC_F_ALLSTAR : 0) |
            (env.novelsplicingp ? DPC_F_NOVEL : 0);
  d.gout = DPC_NO_GOUT;
  o.cls = 0; o.bucket = 0; o.scratch = 0; o.gout = 0; o.ovf = 0;
  if (p.extraband < 0 || p.extraband > 4000) return DPC_ERR_ARG;
  uint64_t qaddr = 0; int qlen = 0;          /* the query span: host address of its first byte, length */
  switch (p.kind) {
  case DPC_SINGLE_GAP: {                                               /* 4450-4572 */
    const int L1 = p.length1, L2 = p.length2;
    int lband, rband;
    if (L1 > env.maxlength1 || L2 > env.maxlength2) {                 /* 4509-4519 */
      r.finalscore = -10000; r.nmatches = r.nmismatches = r.nopens = r.nindels = 0;
      r.dynprogindex_out = dpc_bump(p.dynprogindex);
      return DPC_PREP_HOST;
    }
    if (L1 <= 0 || L2 <= 0) return DPC_ERR_ARG;
    dpc_bands(L1, L2, p.extraband, p.widebandp, &lband, &rband);
    if (L2 - L1 > rband || L1 - L2 > lband) {                         /* unwidened band: the corner was never filled */
      if (L2 - L1 > rband + 1) return DPC_ERR_ARG;
      r.finalscore = (L2 - L1 == rband + 1) ? DPC_NEG_INFINITY : 0;
      r.nmatches = r.nmismatches = r.nopens = r.nindels = 0;
      r.dynprogindex_out = dpc_bump(p.dynprogindex);
      return DPC_PREP_HOST;
    }
    d.type = (uint8_t)dpc_quality(p.defect_rate); d.open = -10; d.extend = -3;
    d.L1 = L1; d.L2 = L2; d.off2 = p.offset2;
    qaddr = (uint64_t)(uintptr_t)p.seq1; qlen = L1;
    break;
  }
  case DPC_END5_GAP: case DPC_END3_GAP: {                              /* 5094-5284, 5556-5741 */
    const bool five = p.kind == DPC_END5_GAP;
    int L1 = p.length1, L2 = p.length2;
    const int ea = p.endalign;
    if (ea < 0 || ea > 3) return DPC_ERR_ARG;
    if (L1 <= 0 || L2 <= 0) {                                          /* 5140-5157 */
      r.nmatches = r.nmismatches = r.nopens = r.nindels = 0; r.finalscore = 0;
      return DPC_PREP_HOST;
    }
    if (ea != DPC_QUERYEND_NOGAPS) {
      if (L1 > env.maxlength1) L1 = env.maxlength1;
      if (L2 > env.maxlength2) L2 = env.maxlength2;
    } else {
      L1 = L2 = (L1 < L2 ? L1 : L2);                                   /* 2358-2369 */
      if (L1 > DPC_INLINE_OPS * DPC_OP_MAXLEN) return DPC_ERR_UNSUPPORTED;
    }
    d.type = 3; d.open = -12; d.extend = -1;                           /* ENDQ, END penalties */
    d.flags |= DPC_F_WIDEBAND;
    d.L1 = L1; d.L2 = L2; d.off2 = p.offset2;
    qaddr = (uint64_t)(uintptr_t)p.seq1 - (five ? (uint64_t)(L1 - 1) : 0u); qlen = L1;
    break;
  }
  case DPC_GENOME_GAP: {                                               /* 4798-5061 */
    const int L1 = p.length1, L2L = p.length2, L2R = p.length2R;
    r.nmatches = r.nmismatches = r.nopens = r.nindels = 0;
    r.left_prob = r.right_prob = 0.0;
    if (L1 <= 1) { r.finalscore = DPC_NEG_INFINITY; return DPC_PREP_HOST; }      /* 4855-4858 */
    if (L1 > env.maxlength1 || L2L > env.maxlength2 || L2R > env.maxlength2) {     /* 4922-4954 */
      r.new_leftgenomepos = p.offset2 - 1; r.new_rightgenomepos = p.offset2R + 1; r.exonhead = p.offset1 + L1 - 1;
      r.dynprogindex_out = dpc_bump(p.dynprogindex); r.finalscore = DPC_NEG_INFINITY;
      return DPC_PREP_HOST;
    }
    if (L2L <= 0 || L2R <= 0 || L2L < L1 - 1 || L2R < L1 - 1) return DPC_ERR_ARG;
    if (p.use_probabilities_p) return DPC_ERR_STATE;                   /* needs the host half (hooks) */
    d.type = (uint8_t)dpc_quality(p.defect_rate);
    if (L1 > p.maxpeelback * 4) { d.open = -10; d.extend = -3; } else { d.open = -18; d.extend = -3; }   /* 4862-4870 */
    d.reward = (int8_t)(!p.splicingp ? 0 : (p.finalp ? 30 : 10) + 6 * d.type);
    d.L1 = L1; d.L2 = L2L; d.L2R = L2R; d.off2 = p.offset2; d.off2R = p.offset2R;
    d.gap = p.offset2R - p.offset2;
    qaddr = (uint64_t)(uintptr_t)p.seq1; qlen = L1;
    break;
  }
  case DPC_CDNA_GAP: {                                                 /* 4577-4793 */
    const int L1L = p.length1, L1R = p.length1R, L2 = p.length2;
    if (L2 <= 1) return DPC_PREP_HOST;                                 /* 4605-4607: nothing is written */
    if (L2 > env.maxlength1 || L1R > env.maxlength2 || L1L > env.maxlength2) {     /* 4648-4670 */
      r.dynprogindex_out = dpc_bump(p.dynprogindex);
      return DPC_PREP_HOST;
    }
    if (L1L <= 0 || L1R <= 0) return DPC_ERR_ARG;
    const int span = p.offset1R - p.offset1 + 1;
    if (span < L1L || span < L1R || span > 100000 || p.seq1R != p.seq1 + (span - 1)) return DPC_ERR_ARG;
    d.type = (uint8_t)dpc_quality(p.defect_rate); d.open = -10; d.extend = -7;
    d.L1 = L1L; d.L1R = L1R; d.L2 = L2; d.off2 = p.offset2;
    d.gap = p.offset1R - p.offset1;
    qaddr = (uint64_t)(uintptr_t)p.seq1; qlen = span;
    break;
  }
  default:
    return DPC_ERR_ARG;
  }
  /* the query bytes must lie in the buffer the caller handed over; the alphabet check reads the device copy */
  if (qaddr < env.qbase || qaddr - env.qbase + (uint64_t)qlen > env.qbytes) return DPC_ERR_ARG;
  d.q0 = (uint32_t)(qaddr - env.qbase);
  if (p.kind == DPC_CDNA_GAP) d.q1 = d.q0 + (uint32_t)(qlen - 1);
  {
    unsigned acc = 0;
    for (int i = 0; i < qlen; i++) acc |= qpool[d.q0 + (uint32_t)i];
    if (acc >= 128) return DPC_ERR_ALPHABET;
  }
  if (!dpc_allstar(p) && (uint64_t)(uint32_t)(p.chroffset + p.chrpos) + p.genomiclength > env.genome_nbases) return DPC_ERR_ARG;
  if (p.kind != DPC_CDNA_GAP)
    o.gout = dpc_gout_span(d.L2) + (p.kind == DPC_GENOME_GAP ? dpc_gout_span(d.L2R) : 0u) + 8u;
  dpc_classify(d, env.fillmode, env.class_bytes, o);
  return DPC_PREP_DEVICE;
}

/* ---- expansion of the traceback ops into Pair records (dynprog.c:2372-2712 and each entry point's assembly) ---- */

/* one side's characters in matrix order: query bytes from the pool, genome characters as the solve kernel staged
 * them (or decoded again when the problem has no staged span) */
struct QSrc { const uint8_t *q; int start, step; };
DPC_HD int dpc_qat(const QSrc &s, int k) { return s.q[s.start + s.step * k]; }
struct GSrc { const uint8_t *staged; const DevProb *p; const uint32_t *blocks; int start, step; };
DPC_HD int dpc_gat(const GSrc &g, int k) {
  if (g.staged) return g.staged[k];
  return dpc_code_char(dpc_genomic_code(*g.p, g.blocks, g.start + g.step * k));
}

/* where push number s of a list goes: dst[base + dir * s], pushes before `first` are dropped (the leading indel
 * pairs an end gap strips, 5265-5268) */
struct Emit { dpc_pair_t *dst; int base, dir, first; };
DPC_HD void dpc_put(const Emit &e, int s, int qpos, int gpos, int cdna, int comp, int genome, int idx, int gapp) {
  if (s < e.first) return;
  dpc_pair_t *o = e.dst + (e.base + e.dir * s);
#ifdef __CUDACC__
  *reinterpret_cast<uint4 *>(o) = make_uint4((unsigned)qpos, (unsigned)gpos, (unsigned)idx,
                                             (unsigned)(cdna & 255) | ((unsigned)(comp & 255) << 8) | ((unsigned)(genome & 255) << 16) | ((unsigned)gapp << 24));
#else
  o->querypos = qpos; o->genomepos = gpos; o->dynprogindex = idx;
  o->cdna = (char)cdna; o->comp = (char)comp; o->genome = (char)genome; o->gapp = (uint8_t)gapp;
#endif
}

/* One matrix: replays the ops from (r,c).  With e.dst == NULL only counts.  Returns the number of pushes; *lead
 * receives the number of pushes before the first one that is not an indel pair.  REV / GROWS as in the host half's
 * replay_t: reversed coordinates (traceback of a `rev` matrix), genome on the rows (cDNA gap). */
DPC_HD int dpc_replay(const Emit &e, const uint16_t *ops, int nops, int r, int c, const QSrc &qs, const GSrc &gs,
                      int q0, int g0, bool REV, bool GROWS, int idx, bool nostar, const DevTables *tb, int *lead, const Lanes &ln) {
  const int step = REV ? -1 : 1;
  const bool count_only = e.dst == 0;
  int s = 0, leading = 0, seen = 0;       /* seen: a push that is not an indel pair has happened */
  for (int i = 0; i < nops; i++) {
    const int op = ops[i] & 3, len = ops[i] >> 2;
    if (op == DPC_OP_M) {
      const int qi = (GROWS ? c : r) - 1, gi = (GROWS ? r : c) - 1;
      if (GROWS || nostar) {
        if (!count_only)
          for (int j = ln.lane; j < len; j += ln.n) {
            const int c1 = dpc_qat(qs, qi - j), c2 = dpc_gat(gs, gi - j);
            int comp = '*';
            if (c1 != c2 && dpc_query_uc(c1) != c2) {
              int code = c2 == 'A' ? 0 : c2 == 'C' ? 1 : c2 == 'G' ? 2 : c2 == 'T' ? 3 : c2 == 'N' ? 4 : 5;
              const int consistent = GROWS ? (tb->consT[c1 & 127] >> code) & 1 : (tb->cons[c1 & 127] >> code) & 1;   /* 2654 vs 2752 */
              comp = consistent ? ':' : ' ';
            }
            dpc_put(e, s + j, q0 + step * (qi - j), g0 + step * (gi - j), c1, comp, c2, idx, 0);
          }
        if (len > 0) seen = 1;
        s += len;
      } else {
        /* columns off the genomic segment ('*') push nothing (2644): rare, one lane walks the run */
        int pushed = 0;
        for (int j = 0; j < len; j++) {
          const int c2 = dpc_gat(gs, gi - j);
          if (c2 == '*') continue;
          if (!count_only && ln.lane == 0) {
            const int c1 = dpc_qat(qs, qi - j);
            int comp = '*';
            if (c1 != c2 && dpc_query_uc(c1) != c2) {
              int code = c2 == 'A' ? 0 : c2 == 'C' ? 1 : c2 == 'G' ? 2 : c2 == 'T' ? 3 : c2 == 'N' ? 4 : 5;
              comp = ((tb->cons[c1 & 127] >> code) & 1) ? ':' : ' ';
            }
            dpc_put(e, s + pushed, q0 + step * (qi - j), g0 + step * (gi - j), c1, comp, c2, idx, 0);
          }
          pushed++;
        }
        if (pushed > 0) seen = 1;
        s += pushed;
      }
      r -= len; c -= len;
      continue;
    }
    const bool along_cols = (op == DPC_OP_QSKIP) ? GROWS : !GROWS;
    if (along_cols) c -= len; else r -= len;
    if (op == DPC_OP_GAPHOLDER) {                                      /* 2507 */
      if (!count_only && ln.lane == 0) dpc_put(e, s, -1, -1, ' ', ' ', ' ', 0, 1);
      seen = 1; s += 1;
      continue;
    }
    if (!seen) leading += len;
    if (!count_only) {
      if (op == DPC_OP_GSKIP) {                                        /* add_genomeskip dashes, 2444-2505 */
        const int lo = GROWS ? r : c, qi2 = GROWS ? c - 1 : r - 1;
        const int qpos = REV ? q0 - qi2 : q0 + qi2 + 1;
        for (int j = ln.lane; j < len; j += ln.n) {
          const int gi2 = lo + len - 1 - j;
          dpc_put(e, s + j, qpos, g0 + step * gi2, ' ', '-', dpc_gat(gs, gi2), idx, 0);
        }
      } else {                                                         /* add_queryskip, 2372-2413 */
        const int lo = GROWS ? c : r, gi2 = GROWS ? r - 1 : c - 1;
        const int gpos = REV ? g0 - gi2 : g0 + gi2 + 1;
        for (int j = ln.lane; j < len; j += ln.n) {
          const int qi2 = lo + len - 1 - j;
          dpc_put(e, s + j, q0 + step * qi2, gpos, dpc_qat(qs, qi2), '-', ' ', idx, 0);
        }
      }
    }
    s += len;
  }
  if (lead) *lead = leading;
  return s;
}

/* The pairs of one problem, in the order of the List_T the reference returns (head first), written to dst (NULL:
 * count only).  Returns their number.  gout: the staged genome characters of the batch (device copy). */
DPC_HD int dpc_expand_one(const dpc_problem_t &p, const DevProb &d, const DevRes &dr, const uint16_t *ops,
                          const uint8_t *pool, const uint8_t *gout, const uint32_t *blocks, const DevTables *tb,
                          dpc_pair_t *dst, const Lanes &ln) {
  const bool nostar = !(dr.status & DPC_ST_STAR);
  const uint8_t *q = pool + d.q0;
  const uint8_t *staged = (gout && d.gout != DPC_NO_GOUT) ? gout + d.gout : (const uint8_t *)0;
  const Emit none = { (dpc_pair_t *)0, 0, 1, 0 };
  switch (p.kind) {
  case DPC_SINGLE_GAP: {
    /* List_reverse of the pushed list (4571) = push order */
    const QSrc qs = { q, 0, 1 };
    const GSrc gs = { staged, &d, blocks, d.off2, 1 };
    const Emit e = { dst, 0, 1, 0 };
    return dpc_replay(dst ? e : none, ops, dr.nopsL, dr.bestrL, dr.bestcL, qs, gs, p.offset1, p.offset2, false, false, p.dynprogindex, nostar, tb, 0, ln);
  }
  case DPC_END5_GAP: case DPC_END3_GAP: {
    const bool five = p.kind == DPC_END5_GAP;
    if ((p.endalign == DPC_QUERYEND_GAP || p.endalign == DPC_BEST_LOCAL) && dr.nmatches + 1 < dr.nmismatches) return 0;   /* 5259 */
    const QSrc qs = { q, five ? d.L1 - 1 : 0, five ? -1 : 1 };
    const GSrc gs = { staged, &d, blocks, d.off2, five ? -1 : 1 };
    int lead = 0, n;
    if (nostar) {
      /* without '*' columns the first push is an aligned column: nothing to strip, and the count is the sum of the runs */
      n = 0;
      for (int k = 0; k < dr.nopsL; k++) n += (ops[k] & 3) == DPC_OP_GAPHOLDER ? 1 : ops[k] >> 2;
    } else {
      n = dpc_replay(none, ops, dr.nopsL, dr.bestrL, dr.bestcL, qs, gs, p.offset1, p.offset2, five, false, p.dynprogindex, nostar, tb, &lead, ln);
    }
    if (dst) {
      /* end5: List_reverse again (5283); end3: as is (5740); leading indel pairs dropped (5265-5268) */
      const Emit e = { dst, five ? n - 1 : -lead, five ? -1 : 1, lead };
      dpc_replay(e, ops, dr.nopsL, dr.bestrL, dr.bestcL, qs, gs, p.offset1, p.offset2, five, false, p.dynprogindex, nostar, tb, 0, ln);
    }
    return n - lead;
  }
  case DPC_GENOME_GAP: {
    if (!(dr.status & DPC_ST_OK)) return 0;
    const int L1 = p.length1, revoffset1 = p.offset1 + L1 - 1;
    const QSrc qf = { q, 0, 1 }, qb = { q, L1 - 1, -1 };
    const GSrc ga = { staged, &d, blocks, d.off2, 1 };
    const GSrc gb = { staged ? staged + dpc_gout_span(d.L2) : (const uint8_t *)0, &d, blocks, d.off2R, -1 };
    const uint16_t *opsR = ops + dr.nopsL;
    const int nR = dpc_replay(none, opsR, dr.nopsR, dr.bestrR, dr.bestcR, qb, gb, revoffset1, p.offset2R, true, false, p.dynprogindex, nostar, tb, 0, ln);
    const int nL = dpc_replay(none, ops, dr.nopsL, dr.bestrL, dr.bestcL, qf, ga, p.offset1, p.offset2, false, false, p.dynprogindex, nostar, tb, 0, ln);
    if (nR + nL == 0) return 0;                                        /* List_length == 1 -> NULL, 5051 */
    if (dst) {
      const Emit eR = { dst, nR - 1, -1, 0 }, eL = { dst, nR + 1, 1, 0 };
      dpc_replay(eR, opsR, dr.nopsR, dr.bestrR, dr.bestcR, qb, gb, revoffset1, p.offset2R, true, false, p.dynprogindex, nostar, tb, 0, ln);
      if (ln.lane == 0) { const Emit eg = { dst, nR, 1, 0 }; dpc_put(eg, 0, -1, -1, ' ', ' ', ' ', 0, 1); }
      dpc_replay(eL, ops, dr.nopsL, dr.bestrL, dr.bestcL, qf, ga, p.offset1, p.offset2, false, false, p.dynprogindex, nostar, tb, 0, ln);
    }
    return nR + 1 + nL;
  }
  case DPC_CDNA_GAP: {
    if (!(dr.status & DPC_ST_OK)) return 0;
    const int L2 = p.length2, revoffset2 = p.offset2 + L2 - 1, span = p.offset1R - p.offset1 + 1;
    const QSrc qf = { q, 0, 1 }, qb = { q, span - 1, -1 };
    const GSrc ga = { (const uint8_t *)0, &d, blocks, d.off2, 1 }, gb = { (const uint8_t *)0, &d, blocks, d.off2 + L2 - 1, -1 };
    const uint16_t *opsR = ops + dr.nopsL;
    const int nR = dpc_replay(none, opsR, dr.nopsR, dr.bestrR, dr.bestcR, qb, gb, p.offset1R, revoffset2, true, true, p.dynprogindex, true, tb, 0, ln);
    const int nL = dpc_replay(none, ops, dr.nopsL, dr.bestrL, dr.bestcL, qf, ga, p.offset1, p.offset2, false, true, p.dynprogindex, true, tb, 0, ln);
    const int queryjump = (p.offset1R - dr.bestcR) - (p.offset1 + dr.bestcL) + 1;      /* 4725-4726 */
    const int genomejump = (revoffset2 - dr.bestrR) - (p.offset2 + dr.bestrL) + 1;
    const int insert = queryjump == 9 && genomejump == 9;
    const int nmid = insert ? 18 : 1;
    if (nR + nmid + nL == 1) return 0;                                 /* 4784-4787 */
    if (dst) {
      const Emit eR = { dst, nR - 1, -1, 0 }, eM = { dst, nR, 1, 0 }, eL = { dst, nR + nmid, 1, 0 };
      dpc_replay(eR, opsR, dr.nopsR, dr.bestrR, dr.bestcR, qb, gb, p.offset1R, revoffset2, true, true, p.dynprogindex, true, tb, 0, ln);
      if (insert) {                                                    /* INSERT_PAIRS, 4730-4751 */
        for (int j = ln.lane; j < 9; j += ln.n) {
          const int kq = p.offset1R - dr.bestcR - j;
          dpc_put(eM, j, kq, revoffset2 - dr.bestrR + 1, q[kq - p.offset1], '~', ' ', p.dynprogindex, 0);
          const int kg = revoffset2 - dr.bestrR - j;
          dpc_put(eM, 9 + j, p.offset1 + dr.bestcL, kg, ' ', '~', dpc_gat(ga, kg - p.offset2), p.dynprogindex, 0);
        }
      } else if (ln.lane == 0) dpc_put(eM, 0, -1, -1, ' ', ' ', ' ', 0, 1);
      dpc_replay(eL, ops, dr.nopsL, dr.bestrL, dr.bestcL, qf, ga, p.offset1, p.offset2, false, true, p.dynprogindex, true, tb, 0, ln);
    }
    return nR + nmid + nL;
  }
  default:
    return 0;
  }
}

/* Turns the device record of a problem into the reference's output parameters (Batch::finalize without hooks:
 * left_prob / right_prob of a final genome gap are left for the host, which owns the MaxEnt hook; the return value
 * says so).  r must have been initialised by dpc_prepare_one.  npairs comes from dpc_expand_one(count only). */
DPC_HD int dpc_finish_one(const dpc_problem_t &p, const DevRes &dr, int npairs, dpc_result_t &r) {
  int needs_probs = 0;
  switch (p.kind) {
  case DPC_SINGLE_GAP:
    r.finalscore = dr.finalscore;
    r.nmatches = dr.nmatches; r.nmismatches = dr.nmismatches; r.nopens = dr.nopens; r.nindels = dr.nindels;
    r.dynprogindex_out = dpc_bump(p.dynprogindex);
    break;
  case DPC_END5_GAP: case DPC_END3_GAP:
    r.finalscore = dr.finalscore;
    r.nmatches = dr.nmatches; r.nmismatches = dr.nmismatches; r.nopens = dr.nopens; r.nindels = dr.nindels;
    r.dynprogindex_out = dpc_bump(p.dynprogindex);
    if ((p.endalign == DPC_QUERYEND_GAP || p.endalign == DPC_BEST_LOCAL) && dr.nmatches + 1 < dr.nmismatches) r.finalscore = 0;   /* 5259-5262 */
    break;
  case DPC_GENOME_GAP:
    r.finalscore = dr.finalscore;
    r.introntype = (dr.status & DPC_ST_HAVE) ? dr.introntype : DPC_UNSET;
    if (dr.status & DPC_ST_OK) {
      needs_probs = p.finalp != 0;                                     /* 4104-4108 */
      r.new_leftgenomepos = p.offset2 + (dr.bestcL - 1);               /* 5000-5004 */
      r.new_rightgenomepos = p.offset2R - (dr.bestcR - 1);
      r.exonhead = (p.offset1 + p.length1 - 1) - (dr.bestrR - 1);
      r.nmatches = dr.nmatches; r.nmismatches = dr.nmismatches; r.nopens = dr.nopens; r.nindels = dr.nindels;
      r.dynprogindex_out = dpc_bump(p.dynprogindex);
      /* the host evaluates get_splicesite_probs for (bestcL, bestcR): pass them through the probability fields */
      if (needs_probs) { r.left_prob = (double)dr.bestcL; r.right_prob = (double)dr.bestcR; }
    }
    break;
  case DPC_CDNA_GAP:
    r.finalscore = dr.finalscore;
    if (dr.status & DPC_ST_OK) {
      const int revoffset2 = p.offset2 + p.length2 - 1;
      const int queryjump = (p.offset1R - dr.bestcR) - (p.offset1 + dr.bestcL) + 1;
      const int genomejump = (revoffset2 - dr.bestrR) - (p.offset2 + dr.bestrL) + 1;
      if (!(queryjump == 9 && genomejump == 9)) r.incompletep = 1;
      r.dynprogindex_out = dpc_bump(p.dynprogindex);
    }
    break;
  default: break;
  }
  r.npairs = npairs;
  r.null_list = npairs == 0;
  return needs_probs;
}

#endif /* DPC_PIPE_H */
