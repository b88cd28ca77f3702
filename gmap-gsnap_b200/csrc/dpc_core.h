/* dpc_core.h -- the gap-fill algorithms of libdynprog_cuda, as warp-level routines.
 *
 * One warp solves one problem: it stages the query and genomic characters in its
 * arena (shared memory, or HBM scratch for oversize problems), fills one or two
 * banded 3-state Gotoh matrices, searches the end point / intron bridge, and
 * walks the traceback, leaving run-length ops for the host to rebuild Pair
 * records from.  Semantics follow the reference's src/dynprog.c (GMAP/GSNAP
 * 2012-07-03); each routine cites the lines it answers to.
 *
 * The routines are written against a `Lanes` handle (lane id, lane count):
 *   - under nvcc they run with 32 lanes and __syncwarp / __shfl_sync;
 *   - tests/emul compiles the same header with g++ and a single lane, so the
 *     index arithmetic, tie-breaks and boundary rules can be unit-tested on a
 *     machine without a GPU.  That build is test scaffolding only; the product
 *     library (dynprog_cuda.cu) has no CPU path.
 */
#ifndef DPC_CORE_H
#define DPC_CORE_H

#include <stdint.h>

#ifdef __CUDACC__
#define DPC_HD __device__ __forceinline__
#define DPC_HDM __device__ __forceinline__
#define DPC_HB __host__ __device__ __forceinline__   /* also called by the host packing code */
#define DPC_SYNC() __syncwarp()
#else
#define DPC_HD static inline
#define DPC_HDM inline
#define DPC_HB static inline
#define DPC_SYNC() ((void)0)
#endif

#define DPC_NEG (-1000000)                 /* NEG_INFINITY, dynprog.c:119 */
#define DPC_BRIDGE_FLOOR (-100000)         /* initial bestscore of the bridges, dynprog.c:3072, 3310 */
#define DPC_MICROINTRON 9                  /* MICROINTRON_LENGTH, dynprog.c:139 */

/* direction nibble of one cell: bits 0-1 nogap direction, bit 2 gap1 came from gap1 (HORIZ),
 * bit 3 gap2 came from gap2 (VERT).  (struct Direction3_T, dynprog.c:717, needs 3 bytes.) */
enum { DPC_DIAG = 0, DPC_HORIZ = 1, DPC_VERT = 2 };

/* traceback ops: (length << 2) | type, in traceback order */
enum { DPC_OP_M = 0,        /* n aligned columns (match / mismatch decided from the characters) */
       DPC_OP_GSKIP = 1,    /* genome-only run kept as dashes (add_genomeskip, dynprog.c:2416) */
       DPC_OP_QSKIP = 2,    /* query-only run (add_queryskip, dynprog.c:2372) */
       DPC_OP_GAPHOLDER = 3 /* genome-only run of >= 9 with intron dinucleotides: one gapholder (2507) */ };

/* genome codes: A C G T N and '*' (outside the genomic segment, dynprog.c:415-419) */
enum { DPC_GN = 4, DPC_GSTAR = 5 };

/* problem flags */
enum {
  DPC_F_WATSON = 1, DPC_F_LATE = 2, DPC_F_WIDEBAND = 4, DPC_F_HALFP = 8, DPC_F_FINALP = 16,
  DPC_F_PROBMODE = 32, DPC_F_ALLSTAR = 64, DPC_F_KNOWN = 128, DPC_F_NOVEL = 256,
  DPC_F_INTRONS = 1024,     /* bridge constrained to the given introns listed after the known flags (dynprog.c:3552-3696) */
  DPC_F_SEQ2 = 512          /* the columns' genome codes are in the byte pool at q1, in matrix order (splice-junction solvers) */
};

/* device-side problem descriptor (host packs it from dpc_problem_t) */
struct DevProb {
  uint32_t q0, q1;          /* byte-pool offsets of the first / second query span (forward order) */
  uint32_t gbase, glen;     /* chroffset + chrpos, genomiclength */
  uint32_t aux;             /* byte-pool offset (8-aligned) of known flags / probabilities, genome gaps only */
  int32_t L1, L1R, L2, L2R; /* lengths after the reference's clipping */
  int32_t off2, off2R;      /* segment-relative genomic start of each matrix */
  int32_t gap;              /* rightoffset - leftoffset of the bridge constraint */
  int32_t score_threshold;
  int32_t extraband;
  uint32_t scratch_lo, scratch_hi; /* HBM scratch offset (bytes) for problems that do not fit shared memory */
  int8_t open, extend, reward, cdna_direction;
  uint8_t kind, endalign, type, pad;
  uint32_t flags;
  uint32_t gout;            /* where the staged genome characters of this problem go in the output byte stream
                               (the host rebuilds the Pair records from them instead of decoding the genome again) */
};
#define DPC_NO_GOUT 0xffffffffu
DPC_HB uint32_t dpc_gout_span(int len) { return ((uint32_t)len + 7u) & ~7u; }     /* second span of a genome gap starts here */

#define DPC_INLINE_OPS 38
#define DPC_OP_MAXLEN 16383                 /* (length << 2) | type in 16 bits */
/* device-side result record, 128 bytes */
struct DevRes {
  int32_t finalscore;
  int32_t nmatches, nmismatches, nopens, nindels;
  int32_t bestrL, bestcL, bestrR, bestcR;
  int32_t introntype;
  uint32_t status;          /* DPC_ST_* */
  uint16_t nopsL, nopsR;
  uint32_t ovf;             /* first op in the overflow arena when nopsL + nopsR > DPC_INLINE_OPS */
  uint16_t ops[DPC_INLINE_OPS];
};
enum { DPC_ST_DONE = 1, DPC_ST_HAVE = 2, DPC_ST_OK = 4, DPC_ST_STAR = 8, DPC_ST_OVF_LOST = 16 };

/* score and consistency tables reduced to the 6 genome codes.
 * score[type][q][g] = pairdistance_array[type][q][g] (dynprog.c:1127-1226);
 * cons[q] bit g = consistent_array[q][g]; consT[q] bit g = consistent_array[g][q]. */
struct DevTables {
  int8_t score[4][128][8];
  uint8_t cons[128];
  uint8_t consT[128];
};

struct Lanes { int lane, n; };

/* ---- genome access: get_genomic_nt, dynprog.c:403-441; uncompress_one_char, genome.c:9325 ---- */
DPC_HB int dpc_genome_code(const uint32_t *blocks, uint32_t pos) {
  const uint32_t *b = blocks + (uint64_t)(pos >> 5) * 3;
  int bit = (int)(pos & 31);
  if ((b[2] >> bit) & 1U) return DPC_GN;
  return (int)((bit < 16 ? b[1] >> (2 * bit) : b[0] >> (2 * bit - 32)) & 3U);
}
DPC_HD int dpc_genomic_code(const DevProb &p, const uint32_t *blocks, int genomicpos) {
  if (genomicpos < 0 || (uint32_t)genomicpos >= p.glen || (p.flags & DPC_F_ALLSTAR)) return DPC_GSTAR;
  if (p.flags & DPC_F_WATSON) return dpc_genome_code(blocks, p.gbase + (uint32_t)genomicpos);
  int c = dpc_genome_code(blocks, p.gbase + (p.glen - 1) - (uint32_t)genomicpos);
  return c < 4 ? (c ^ 3) : c;                      /* complCode, complement.h:31 */
}
DPC_HB int dpc_code_char(int code) { return "ACGTN*"[code]; }
DPC_HB int dpc_query_uc(int c) {                   /* UPPERCASE_U2T, complement.h:36 */
  if (c >= 'a' && c <= 'z') c -= 32;
  return c == 'U' ? 'T' : c;
}

/* One exact-match scan of Dynprog_microexon_int (dynprog.c:7305-7312: BoyerMoore_nt of the query's middle piece over
 * the intron, boyer-moore.c:384): does pat[0..len) occur at segment position textleft + j, for j in [0, npos)?
 * The characters come from the 2-bit genome with the semantics of boyer-moore.c:357-381 (Watson as stored, Crick
 * complemented from the far end; an N never matches).  Hits go to hits[hits_off ..], their number to count[query]. */
struct ScanQuery {
  uint32_t gbase, glen;          /* chroffset + chrpos, genomiclength */
  int32_t textleft, npos;
  uint32_t pat;                  /* 2 bits per base, base i at bits 2i.. (at most 16 bases) */
  uint8_t len, watson, pad[2];
  uint32_t hits_off;
};
DPC_HB bool dpc_scan_match(const ScanQuery &q, const uint32_t *blocks, uint64_t nbases, int j) {
  for (int i = 0; i < q.len; i++) {
    const int pos = q.textleft + j + i;
    const uint64_t g = q.watson ? (uint64_t)q.gbase + (uint32_t)pos : (uint64_t)q.gbase + (q.glen - 1) - (uint32_t)pos;
    if (g >= nbases) return false;
    int code = dpc_genome_code(blocks, (uint32_t)g);
    if (code > 3) return false;
    if (!q.watson) code ^= 3;
    if (code != (int)((q.pat >> (2 * i)) & 3u)) return false;
  }
  return true;
}

/* ---- one banded matrix ----------------------------------------------------------------- */
struct Mat {
  int L1, L2;               /* rows, columns */
  int lband, rband, W;      /* W = lband + rband + 1 diagonals */
  int wstride;              /* 32-bit words of direction nibbles per row */
  int open, extend, late;
  int query_rows;           /* 1: rows = query, columns = genome (compute_scores_lookup_fwd/_rev, 1424-1736);
                               0: rows = genome, columns = query (_fwd_12/_rev_12, 1741-2044) */
  uint8_t *rowch, *colch;   /* characters in matrix order: raw query bytes / genome codes */
  int planes, cpl, cplsh;   /* planes != 0: directions are bit planes; every lane owns cpl = 1 << cplsh adjacent diagonals */
  uint32_t *dir;            /* nibbles: rows 1..L1, nibble (r-1)*wstride*8 + (c-r+lband);
                               planes:  word ((r-1)*cpl + k%cpl)*4 + p, bit k/cpl, k = c-r+lband, with plane
                               p = 0 nogap did not come from nogap (HORIZ or VERT), 1 nogap came from gap2 (VERT),
                                   2 gap1 of cell k+1 came from gap1 (HORIZ), 3 gap2 came from gap2 (VERT) */
  int16_t *nband;           /* nogap score of rows 1..L1, (r-1)*W + (c-r+lband), 16 bit (real scores are within
                               +-28000 for any size dpc_init accepts); NULL when no bridge follows */
};

DPC_HB void dpc_bands(int L1, int L2, int extraband, int widebandp, int *lband, int *rband) {
  /* dynprog.c:1442-1454 */
  if (!widebandp) { *lband = *rband = extraband; }
  else if (L2 >= L1) { *rband = L2 - L1 + extraband; *lband = extraband; }
  else { *lband = L1 - L2 + extraband; *rband = extraband; }
}
DPC_HB int dpc_wstride(int W) { return (W + 7) >> 3; }
DPC_HD bool dpc_inband(const Mat &m, int r, int c) {
  int k = c - r;
  return r >= 1 && c >= 1 && r <= m.L1 && c <= m.L2 && k >= -m.lband && k <= m.rband;
}
DPC_HD int dpc_nib(const Mat &m, int r, int c) {
  int idx = (r - 1) * (m.wstride << 3) + (c - r + m.lband);
  return (int)((m.dir[idx >> 3] >> ((idx & 7) << 2)) & 15U);
}
DPC_HD int dpc_plane_bit(const Mat &m, int r, int k, int p) {
  return (int)((m.dir[((r - 1) * m.cpl + (k & (m.cpl - 1))) * 4 + p] >> (k >> m.cplsh)) & 1U);
}
/* directions as the reference's matrices hold them, including row 0 / column 0 (1460-1488)
 * and the memset STOP everywhere else (724-751): returns -1 for STOP */
DPC_HD int dpc_dirN(const Mat &m, int r, int c) {
  if (!dpc_inband(m, r, c)) return -1;
  if (!m.planes) return dpc_nib(m, r, c) & 3;
  int k = c - r + m.lband;
  if (!dpc_plane_bit(m, r, k, 0)) return DPC_DIAG;
  return dpc_plane_bit(m, r, k, 1) ? DPC_VERT : DPC_HORIZ;
}
/* 1 when the nogap direction of an IN-BAND cell is HORIZ or VERT (the bridges' -1, dynprog.c:3724) */
DPC_HD int dpc_nondiag(const Mat &m, int r, int c) {
  const int k = c - r + m.lband;
  if (!m.planes) return (dpc_nib(m, r, c) & 3) != DPC_DIAG;
  return dpc_plane_bit(m, r, k, 0);
}
DPC_HD bool dpc_g1_horiz(const Mat &m, int r, int c) {
  if (r == 0) return c >= 2 && c <= m.rband && c <= m.L2;
  if (!dpc_inband(m, r, c)) return false;
  if (!m.planes) return (dpc_nib(m, r, c) >> 2) & 1;
  int k = c - r + m.lband;
  if (c == 1 || k == 0) return true;     /* left neighbour is column 0 or the forced cell: NEG beats NEG + open */
  return dpc_plane_bit(m, r, k - 1, 2);
}
DPC_HD bool dpc_g2_vert(const Mat &m, int r, int c) {
  if (c == 0) return r >= 2 && r <= m.lband && r <= m.L1;
  if (!dpc_inband(m, r, c)) return false;
  if (!m.planes) return (dpc_nib(m, r, c) >> 3) & 1;
  return dpc_plane_bit(m, r, c - r + m.lband, 3);
}
DPC_HD int dpc_nscore(const Mat &m, int r, int c) {
  /* nogap score as the bridges read it; row 0 inside the band is NEG (1464-1475) */
  if (dpc_inband(m, r, c)) return m.nband[(r - 1) * m.W + (c - r + m.lband)];
  return DPC_NEG;
}

/* argmax with an explicit scan-order key: the reference's loops keep the FIRST best (`>`)
 * or, with jump_late_p, the LAST best (`>=`) in their own scan order (2251-2285). */
struct Best { int score; int key; };
DPC_HD bool dpc_better(int s, int k, const Best &b, int late) {
  return s > b.score || (s == b.score && (late ? k > b.key : k < b.key));
}
DPC_HD void dpc_warp_best(Best &b, int late, const Lanes &ln) {
#ifdef __CUDACC__
  for (int o = 16; o > 0; o >>= 1) {
    int s = __shfl_xor_sync(0xffffffffu, b.score, o), k = __shfl_xor_sync(0xffffffffu, b.key, o);
    if (dpc_better(s, k, b, late)) { b.score = s; b.key = k; }
  }
#else
  (void)b; (void)late; (void)ln;
#endif
}
DPC_HD int dpc_warp_sum(int v, const Lanes &ln) {
#ifdef __CUDACC__
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#else
  (void)ln;
#endif
  return v;
}
DPC_HD int dpc_bcast(int v, const Lanes &ln) {
#ifdef __CUDACC__
  return __shfl_sync(0xffffffffu, v, 0);
#else
  (void)ln; return v;
#endif
}

/* end-point search folded into the fill.
 * mode 1: find_best_endpoint (2235-2290): rows 1..L1, |c - r| <= extraband (the UNwidened band), start 0 at (0,0)
 * mode 2: find_best_endpoint_to_queryend_indels (2293-2355): last row, widened band, start NEG at (L1,0) */
struct EndSearch { int mode, eb; Best best; };   /* mode 3: the corner (L1,L2), Dynprog_single_gap 4541 */

DPC_HD void dpc_cell(int Nl, int G1l, int Nu, int G2u, int Nd, int G1d, int G2d, int P,
                     int open, int extend, int late, int *N, int *G1, int *G2, int *nib) {
  /* dynprog.c:1519-1561 */
  int best, s, d1 = 0, d2 = 0, dn = DPC_DIAG;
  best = Nl + open; s = G1l;
  if (s > best || (s == best && late)) { best = s; d1 = 1; }
  *G1 = best + extend;
  best = Nu + open; s = G2u;
  if (s > best || (s == best && late)) { best = s; d2 = 1; }
  *G2 = best + extend;
  best = Nd;
  if (G1d > best || (G1d == best && late)) { best = G1d; dn = DPC_HORIZ; }
  if (G2d > best || (G2d == best && late)) { best = G2d; dn = DPC_VERT; }
  *N = best + P;
  *nib = dn | (d1 << 2) | (d2 << 3);
}

/* Reference fill, any band width: anti-diagonal sweep with the last two anti-diagonals kept in
 * `st` (3 rotating buffers of N, G1, G2 indexed by row; 9*(L1+1) ints).  compute_scores_lookup_*,
 * dynprog.c:1424-2044: all four variants leave the same matrix (see oracle/dynprog_port.c:fill). */
DPC_HD void dpc_fill_generic(const Mat &m, int32_t *st, const int8_t *score /* [128][8] of the mismatch type */,
                             EndSearch &es, const Lanes &ln) {
  const int R = m.L1 + 1, L1 = m.L1, L2 = m.L2;
  for (int d = 2; d <= L1 + L2; d++) {
    int32_t *cur = st + (d % 3) * 3 * R, *p1 = st + ((d + 2) % 3) * 3 * R, *p2 = st + ((d + 1) % 3) * 3 * R;
    int rlo = (d - m.rband + 1) >> 1, rhi = (d + m.lband) >> 1;
    if (rlo < 1) rlo = 1;
    if (rlo < d - L2) rlo = d - L2;
    if (rhi > L1) rhi = L1;
    if (rhi > d - 1) rhi = d - 1;
    for (int r = rlo + ln.lane; r <= rhi; r += ln.n) {
      int c = d - r, k = c - r;
      int Nl, G1l, Nu, G2u, Nd, G1d, G2d;
      /* left neighbour (r, c-1) */
      if (c - 1 == 0) { Nl = DPC_NEG; G1l = DPC_NEG; }
      else if (k - 1 < -m.lband) { Nl = DPC_NEG; G1l = DPC_NEG; }          /* forced, 1507-1513 */
      else { Nl = p1[r]; G1l = p1[R + r]; }
      /* upper neighbour (r-1, c) */
      if (r - 1 == 0) { Nu = DPC_NEG; G2u = DPC_NEG; }
      else if (k + 1 > m.rband) { Nu = DPC_NEG; G2u = DPC_NEG; }           /* forced, 1501-1506 */
      else { Nu = p1[r - 1]; G2u = p1[2 * R + r - 1]; }
      /* diagonal neighbour (r-1, c-1): row 0 / column 0 per 1460-1488 */
      if (r - 1 == 0) {
        if (c - 1 == 0) { Nd = 0; G1d = DPC_NEG; G2d = DPC_NEG; }
        else { Nd = DPC_NEG; G1d = m.open + (c - 1) * m.extend; G2d = DPC_NEG; }
      } else if (c - 1 == 0) { Nd = DPC_NEG; G1d = DPC_NEG; G2d = m.open + (r - 1) * m.extend; }
      else { Nd = p2[r - 1]; G1d = p2[R + r - 1]; G2d = p2[2 * R + r - 1]; }
      int q = m.query_rows ? m.rowch[r - 1] : m.colch[c - 1];
      int g = m.query_rows ? m.colch[c - 1] : m.rowch[r - 1];
      int N, G1, G2, nib;
      dpc_cell(Nl, G1l, Nu, G2u, Nd, G1d, G2d, score[(q & 127) * 8 + g], m.open, m.extend, m.late, &N, &G1, &G2, &nib);
      cur[r] = N; cur[R + r] = G1; cur[2 * R + r] = G2;
      int idx = (r - 1) * (m.wstride << 3) + (k + m.lband), sh = (idx & 7) << 2;
      m.dir[idx >> 3] = (m.dir[idx >> 3] & ~(15U << sh)) | ((uint32_t)nib << sh);
      if (m.nband) m.nband[(r - 1) * m.W + (k + m.lband)] = (int16_t)(N < -32768 ? -32768 : N);
      if (es.mode == 1) {
        if (k >= -es.eb && k <= es.eb && dpc_better(N, r * (L2 + 1) + c, es.best, m.late)) { es.best.score = N; es.best.key = r * (L2 + 1) + c; }
      } else if (es.mode == 2) {
        if (r == L1 && dpc_better(N, r * (L2 + 1) + c, es.best, m.late)) { es.best.score = N; es.best.key = r * (L2 + 1) + c; }
      } else if (es.mode == 3) {
        if (r == L1 && c == L2) { es.best.score = N; es.best.key = r * (L2 + 1) + c; }
      }
    }
    DPC_SYNC();
  }
}


/* ---- traceback (2611-2712), traceback_cdna (2715-2810), add_genomeskip (2416-2601) ------- */
DPC_HD int dpc_intron_type(int l1, int l2, int r2, int r1, int cdna_direction) {
  /* Intron_type, intron.c:17-190, on genome codes (A0 C1 G2 T3) */
  int left, right, t;
  if (l1 == 2 && l2 == 3) left = 0x21; else if (l1 == 2 && l2 == 1) left = 0x10;
  else if (l1 == 0 && l2 == 3) left = 0x08; else if (l1 == 1 && l2 == 3) left = 0x06; else return 0;
  if (r2 == 0 && r1 == 2) right = 0x30; else if (r2 == 0 && r1 == 1) right = 0x0C;
  else if (r2 == 2 && r1 == 1) right = 0x02; else if (r2 == 0 && r1 == 3) right = 0x01; else return 0;
  t = left & right;
  if (t == 0) return 0;
  if (cdna_direction > 0) return t < 0x08 ? 0 : t;
  if (cdna_direction < 0) return t > 0x04 ? 0 : t;
  return 0;
}

struct Counts { int nmatches, nmismatches, nopens, nindels, star; };

/* One gap run found by the walk at the cell (r,c) it hangs off: which op it becomes (2416-2601). */
DPC_HD int dpc_run_op(const Mat &m, int d, int dist, int r, int c, int revp, int cdna_direction) {
  const int genome_rows = !m.query_rows;
  int genome_run = (d == DPC_HORIZ) ? !genome_rows : genome_rows;
  int op = genome_run ? DPC_OP_GSKIP : DPC_OP_QSKIP;
  if (genome_run && dist >= DPC_MICROINTRON) {
    const uint8_t *g = genome_rows ? m.rowch : m.colch;
    int lo = genome_rows ? r : c, a, b, y, z;          /* 0-based first genome index of the run */
    if (!revp) { a = g[lo]; b = g[lo + 1]; y = g[lo + dist - 2]; z = g[lo + dist - 1]; }
    else { a = g[lo + dist - 1]; b = g[lo + dist - 2]; y = g[lo + 1]; z = g[lo]; }
    if (dpc_intron_type(a, b, y, z, cdna_direction) != 0) op = DPC_OP_GAPHOLDER;
  }
  return op;
}

/* Serial walk (lane 0) from (r0,c0): writes run-length ops in traceback order, returns their number. */
DPC_HD int dpc_walk_serial(const Mat &m, int r0, int c0, int revp, int cdna_direction, uint16_t *ops, const Lanes &ln) {
  int nops = 0;
  if (ln.lane == 0) {
    int r = r0, c = c0, run = 0;
    while (dpc_inband(m, r, c)) {
      int d = dpc_dirN(m, r, c);
      run++;
      r--; c--;
      if (d == DPC_DIAG) continue;
      ops[nops++] = (uint16_t)((run << 2) | DPC_OP_M); run = 0;
      int dist = 1;
      if (d == DPC_HORIZ) { while (dpc_g1_horiz(m, r, c)) { dist++; c--; } c--; }
      else { while (dpc_g2_vert(m, r, c)) { dist++; r--; } r--; }
      ops[nops++] = (uint16_t)((dist << 2) | dpc_run_op(m, d, dist, r, c, revp, cdna_direction));
    }
    if (run) ops[nops++] = (uint16_t)((run << 2) | DPC_OP_M);
  }
  DPC_SYNC();
  return dpc_bcast(nops, ln);
}

/* All lanes classify the cells of the M runs (2644-2667) and add up the indel counts. */
DPC_HD void dpc_count_ops(const Mat &m, int r0, int c0, const uint16_t *ops, int nops, Counts &ct,
                          const DevTables *tb, const Lanes &ln) {
  const int genome_rows = !m.query_rows;
  int r = r0, c = c0, nm = 0, nmm = 0, star = 0, opens = 0, indels = 0;
  for (int i = 0; i < nops; i++) {
    int op = ops[i] & 3, len = ops[i] >> 2;
    if (op == DPC_OP_M) {
      for (int j = ln.lane; j < len; j += ln.n) {
        int q = genome_rows ? m.colch[c - 1 - j] : m.rowch[r - 1 - j];
        int g = genome_rows ? m.rowch[r - 1 - j] : m.colch[c - 1 - j];
        if (!genome_rows && g == DPC_GSTAR) star++;
        else if (dpc_query_uc(q) == dpc_code_char(g)) nm++;
        else if (((genome_rows ? tb->consT[q & 127] : tb->cons[q & 127]) >> g) & 1) nm++;
        else nmm++;
      }
      r -= len; c -= len;
    } else {
      int along_cols = (op == DPC_OP_QSKIP) ? genome_rows : !genome_rows;
      if (along_cols) c -= len; else r -= len;
      if (op != DPC_OP_GAPHOLDER) { opens++; indels += len; }
    }
  }
  ct.nopens += opens; ct.nindels += indels;
  ct.nmatches += dpc_warp_sum(nm, ln);
  ct.nmismatches += dpc_warp_sum(nmm, ln);
  ct.star += dpc_warp_sum(star, ln);
}

/* ---- intron bridge: intron_score 3148-3192, bridge_intron_gap 3290-4122 ----------------- */
DPC_HD int dpc_leftdi(int a, int b) {   /* 3331-3352 */
  return (a == 2 && b == 3) ? 0x21 : (a == 2 && b == 1) ? 0x10 : (a == 0 && b == 3) ? 0x08 : (a == 1 && b == 3) ? 0x06 : 0;
}
DPC_HD int dpc_rightdi(int r2, int r1) { /* 3354-3373 */
  return (r2 == 0 && r1 == 2) ? 0x30 : (r2 == 0 && r1 == 1) ? 0x0C : (r2 == 2 && r1 == 1) ? 0x02 : (r2 == 0 && r1 == 3) ? 0x01 : 0;
}
DPC_HD int dpc_intron_score(int *introntype, int leftdi, int rightdi, int cdna_direction, int reward, int finalp) {
  int t = leftdi & rightdi, fwd, s;
  *introntype = 0;
  if (t == 0) return 0;
  fwd = t >= 0x08;
  if ((cdna_direction > 0 && !fwd) || (cdna_direction < 0 && fwd)) return 0;
  switch (t) {
  case 0x20: case 0x04: s = reward; break;
  case 0x10: case 0x02: s = finalp ? 20 : 15; break;
  case 0x08: case 0x01: s = 12; break;
  default: return 0;
  }
  *introntype = t;
  return s;
}

struct Bridge { int have, finalscore, rL, cL, rR, cR, introntype; };

/* mL: rows = query forward, columns = genome from offset2L; mR: rows = query backward, columns =
 * genome backward from revoffset2R.  lknown / rknown: 1 where a known splice site sits (NULL =
 * none); lp / rp: per-position probabilities for use_probabilities_p (3829-4081).
 * ldi / rdi (length2L / length2R bytes) and itab (64 bytes) are scratch: the dinucleotide codes of every column
 * (3331-3373) and intron_score for every (leftdi & rightdi) value are tabulated once, so one candidate costs two
 * byte loads, an AND and a table load instead of two compare chains and a switch.  Candidates are in band by
 * construction (their column ranges are clipped to the band, 3545-3549), so the nogap band is read directly. */
/* the tables of the bridge: dinucleotide code of every column (3331-3373) and intron_score for every
 * (leftdi & rightdi) value, so that one candidate costs two byte loads, an AND and a table load */
DPC_HD void dpc_bridge_tables(const Mat &mL, const Mat &mR, const DevProb &p, uint8_t *ldi, uint8_t *rdi, int8_t *itab, const Lanes &ln) {
  const int L2L = mL.L2, L2R = mR.L2, finalp = (p.flags & DPC_F_FINALP) != 0;
  const uint8_t *gL = mL.colch, *gR = mR.colch;
  int it = 0; (void)it;
  for (int c = ln.lane; c < L2L; c += ln.n) ldi[c] = (uint8_t)dpc_leftdi(gL[c], gL[c + 1]);
  for (int c = ln.lane; c < L2R; c += ln.n) rdi[c] = (uint8_t)dpc_rightdi(gR[c + 1], gR[c]);
  for (int t = ln.lane; t < 64; t += ln.n) itab[t] = (int8_t)dpc_intron_score(&it, t, t, p.cdna_direction, p.reward, finalp);
  DPC_SYNC();
}

/* from the winning candidate (scan-order key) to the reference's outputs: bestrL/cL/rR/cR, finalscore, introntype */
DPC_HD void dpc_bridge_finish(Bridge &br, Best best, double bestprob, int probkey, const Mat &mL, const Mat &mR, const DevProb &p,
                              const uint8_t *lknown, const uint8_t *rknown, const uint8_t *introns, const Lanes &ln) {
  const int L1 = mL.L1, L2L = mL.L2, L2R = mR.L2, eb = p.extraband;
  const int rbandL = L2L - L1 + eb, lbandL = eb, lbandR = eb;
  const int finalp = (p.flags & DPC_F_FINALP) != 0, halfp = (p.flags & DPC_F_HALFP) != 0;
  const int probmode = (p.flags & DPC_F_PROBMODE) != 0;
  const uint8_t *gL = mL.colch, *gR = mR.colch;
  int it = 0; (void)it;
  (void)L2R; (void)bestprob;
  if (probmode) {
#ifdef __CUDACC__
    for (int o = 16; o > 0; o >>= 1) {
      double q = __shfl_xor_sync(0xffffffffu, bestprob, o); int k = __shfl_xor_sync(0xffffffffu, probkey, o);
      if (q > bestprob || (q == bestprob && k < probkey)) { bestprob = q; probkey = k; }
    }
#endif
    best.key = probkey;
    br.have = probkey != 0x7fffffff;
  } else {
    dpc_warp_best(best, 0, ln);
    br.have = best.key != 0x7fffffff;
  }
  br.introntype = 0;
  if (!br.have) {
    br.finalscore = (probmode || introns) ? DPC_BRIDGE_FLOOR : (halfp ? DPC_BRIDGE_FLOOR - DPC_BRIDGE_FLOOR / 2 : DPC_BRIDGE_FLOOR);
    br.rL = br.cL = br.rR = br.cR = 0;
    return;
  }
  {
    int rL = best.key >> 13, j = best.key & 8191, rR = L1 - rL;
    int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
    int cloR = rR - lbandR < 1 ? 1 : rR - lbandR;
    int nL = chighL - cloL + 1;
    if (nL < 0) nL = 0;
    int left = j < nL, cL = left ? cloL + j : rL, cR = left ? rR : cloR + (j - nL);
    int sI = dpc_intron_score(&it, dpc_leftdi(gL[cL], gL[cL + 1]), dpc_rightdi(gR[cR + 1], gR[cR]), p.cdna_direction, p.reward, finalp);
    br.rL = rL; br.cL = cL; br.rR = rR; br.cR = cR;
    if (introns) {                                                      /* 3694-3695 */
      br.finalscore = best.score;
      br.introntype = 0;
    } else if (probmode) {                                                     /* 4055-4080: -1 on both sides */
      int sL = dpc_nscore(mL, rL, cL) + (lknown && lknown[cL] ? 20 : 0) - (dpc_dirN(mL, rL, cL) > 0 ? 1 : 0);
      int sR = dpc_nscore(mR, rR, cR) + (rknown && rknown[cR] ? 20 : 0) - (dpc_dirN(mR, rR, cR) > 0 ? 1 : 0);
      br.finalscore = halfp ? sL + sI + sR - sI / 2 : sL + sI + sR;
      br.introntype = -1;                                               /* *introntype left untouched, 4071 */
    } else {
      br.finalscore = halfp ? best.score - sI / 2 : best.score;        /* 3823-3827 */
      br.introntype = it;
    }
  }
}

DPC_HD void dpc_bridge_intron(Bridge &br, const Mat &mL, const Mat &mR, const DevProb &p,
                              const uint8_t *lknown, const uint8_t *rknown, const double *lp, const double *rp,
                              const uint8_t *introns, uint8_t *ldi, uint8_t *rdi, int8_t *itab, const Lanes &ln) {
  const int L1 = mL.L1, L2L = mL.L2, L2R = mR.L2, eb = p.extraband;
  const int rbandL = L2L - L1 + eb, lbandL = eb, rbandR = L2R - L1 + eb, lbandR = eb;   /* 3545-3549 */
  const int finalp = (p.flags & DPC_F_FINALP) != 0, halfp = (p.flags & DPC_F_HALFP) != 0;
  const int probmode = (p.flags & DPC_F_PROBMODE) != 0;
  const uint8_t *gL = mL.colch, *gR = mR.colch;
  Best best; best.score = DPC_BRIDGE_FLOOR; best.key = 0x7fffffff;
  double bestprob = 0.0; int probkey = 0x7fffffff;
  int it = 0; (void)it;
  dpc_bridge_tables(mL, mR, p, ldi, rdi, itab, ln);
  const bool fast = mL.planes && mR.planes && mL.cpl == 1 && mR.cpl == 1 && !lknown && !rknown && !probmode;
  if (introns) {
    /* novelsplicingp == false with an intron-level IIT, 3552-3696: only (cL, cR) pairs that are given introns (the
       host looked them up), both sites known, no intron score and no known-site reward; the pair (cL, cR) can be
       met by the left scan of row pair (length1 - cR, cR) and by the right scan of row pair (cL, length1 - cL) */
    const uint32_t n = *(const uint32_t *)introns;
    const uint16_t *pr = (const uint16_t *)(introns + 4);
    for (uint32_t i = (uint32_t)ln.lane; i < n; i += (uint32_t)ln.n) {
      const int cL = pr[2 * i], cR = pr[2 * i + 1];
      {
        const int rR = cR, rL = L1 - rR;
        const int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
        if (rL >= 1 && rL < L1 && cL >= cloL && cL <= chighL && cR < p.gap - cL) {
          const int s = dpc_nscore(mL, rL, cL) - dpc_nondiag(mL, rL, cL) + dpc_nscore(mR, rR, cR);
          const int key = rL * 8192 + (cL - cloL);
          if (dpc_better(s, key, best, 0)) { best.score = s; best.key = key; }
        }
      }
      {
        const int rL = cL, rR = L1 - rL;
        const int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
        const int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L2R - 1 ? L2R - 1 : rR + rbandR;
        const int nL = chighL >= cloL ? chighL - cloL + 1 : 0;
        if (rL >= 1 && rL < L1 && cR >= cloR && cR <= chighR && cL < p.gap - cR) {
          const int s = dpc_nscore(mR, rR, cR) - dpc_nondiag(mR, rR, cR) + dpc_nscore(mL, rL, cL);
          const int key = rL * 8192 + nL + (cR - cloR);
          if (dpc_better(s, key, best, 0)) { best.score = s; best.key = key; }
        }
      }
    }
  }
  /* common case (bands of at most 32 diagonals, no known sites, integer mode): lane = diagonal, exactly like the
     fill -- one candidate per lane and side, rows of the nogap band and of the direction planes read directly.
     Four row pairs at a time, loads first: for long gaps the bands and planes sit in HBM scratch (L2), and one
     row pair at a time is a chain of dependent load latencies (ncu: half of the long-gap launch was spent here). */
  enum { BR = 4 };
  for (int rL0 = 1; rL0 < L1 && fast && !introns; rL0 += BR) {
    for (int k = ln.lane; k < 32; k += ln.n) {                   /* one trip per lane on the GPU */
      const int kL = k < mL.W ? k : mL.W - 1, kR = k < mR.W ? k : mR.W - 1;
      int vL[BR], vR[BR], dL[BR], dR[BR];
      uint32_t hL[BR], hR[BR];
#pragma unroll
      for (int u = 0; u < BR; u++) {
        const int rL = rL0 + u < L1 ? rL0 + u : L1 - 1, rR = L1 - rL;
        const int16_t *rowL = mL.nband + (rL - 1) * mL.W, *rowR = mR.nband + (rR - 1) * mR.W;
        const uint32_t *wL = mL.dir + (rL - 1) * 4, *wR = mR.dir + (rR - 1) * 4;
        vL[u] = rowL[kL]; vR[u] = rowR[kR];
        dL[u] = rowL[mL.lband]; dR[u] = rowR[mR.lband];          /* the main-diagonal cells (rL,rL) and (rR,rR) */
        hL[u] = wL[0]; hR[u] = wR[0];
      }
#pragma unroll
      for (int u = 0; u < BR; u++) {
        const int rL = rL0 + u, rR = L1 - rL;
        if (rL >= L1) break;
        const int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
        const int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L2R - 1 ? L2R - 1 : rR + rbandR;
        const int nL = chighL >= cloL ? chighL - cloL + 1 : 0;
        const int diR = rdi[rR], diL = ldi[rL];
        const int cL = rL - mL.lband + k, cR = rR - mR.lband + k;
        if (cL >= cloL && cL <= chighL && rR < p.gap - cL) {
          const int s = vL[u] - (int)((hL[u] >> k) & 1U) + itab[ldi[cL] & diR] + dR[u];
          /* with one diagonal per lane (ln.n == 32, the GPU) a lane meets its candidates in ascending key order --
             rows ascending, left scan before right scan -- so "first best" inside the lane is a strict comparison;
             ties between lanes are settled by key afterwards.  A single lane walking all diagonals needs the key. */
          const int key = rL * 8192 + (cL - cloL);
          if (ln.n == 32 ? s > best.score : dpc_better(s, key, best, 0)) { best.score = s; best.key = key; }
        }
        if (cR >= cloR && cR <= chighR && rL < p.gap - cR) {
          const int s = vR[u] - (int)((hR[u] >> k) & 1U) + itab[diL & rdi[cR]] + dL[u];
          const int key = rL * 8192 + nL + (cR - cloR);
          if (ln.n == 32 ? s > best.score : dpc_better(s, key, best, 0)) { best.score = s; best.key = key; }
        }
      }
    }
  }
  for (int rL = 1; rL < L1 && !fast && !introns; rL++) {
    const int rR = L1 - rL;
    int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
    int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L2R - 1 ? L2R - 1 : rR + rbandR;
    int nL = chighL - cloL + 1, nR = chighR - cloR + 1;
    if (nL < 0) nL = 0;
    if (nR < 0) nR = 0;
    /* the two main-diagonal cells every candidate of this row pair is combined with */
    const int dR = dpc_nscore(mR, rR, rR) + (rknown && rknown[rR] ? 20 : 0);
    const int dL = dpc_nscore(mL, rL, rL) + (lknown && lknown[rL] ? 20 : 0);
    const int diR = rdi[rR], diL = ldi[rL];
    const int16_t *rowL = mL.nband + (rL - 1) * mL.W + (mL.lband - rL);     /* rowL[cL] = nogap[rL][cL] */
    const int16_t *rowR = mR.nband + (rR - 1) * mR.W + (mR.lband - rR);
    /* left scan (3700-3766): cL over the band of row rL, cR = rR */
    for (int j = ln.lane; j < nL; j += ln.n) {
      const int cL = cloL + j;
      if (!(rR < p.gap - cL)) continue;
      if (probmode && lp[cL] + rp[rR] <= bestprob) continue;            /* 3918 */
      const int s = rowL[cL] + (lknown && lknown[cL] ? 20 : 0) - dpc_nondiag(mL, rL, cL)                /* 3724-3727 */
                    + itab[ldi[cL] & diR] + dR;
      const int key = rL * 8192 + j;
      if (probmode) { if (s >= p.score_threshold) { bestprob = lp[cL] + rp[rR]; probkey = key; } }
      else if (dpc_better(s, key, best, 0)) { best.score = s; best.key = key; }
    }
    /* right scan (3768-3816): cR over the band of row rR, cL = rL */
    for (int j = ln.lane; j < nR; j += ln.n) {
      const int cR = cloR + j;
      if (!(rL < p.gap - cR)) continue;
      if (probmode && lp[rL] + rp[cR] <= bestprob) continue;            /* 3971 */
      const int s = rowR[cR] + (rknown && rknown[cR] ? 20 : 0) - dpc_nondiag(mR, rR, cR)                /* 3774-3777 */
                    + itab[diL & rdi[cR]] + dL;
      const int key = rL * 8192 + nL + j;
      if (probmode) { if (s >= p.score_threshold) { bestprob = lp[rL] + rp[cR]; probkey = key; } }
      else if (dpc_better(s, key, best, 0)) { best.score = s; best.key = key; }
    }
  }
  dpc_bridge_finish(br, best, bestprob, probkey, mL, mR, p, lknown, rknown, introns, ln);
}

/* bridge_cdna_gap, 3066-3146: rows = genome, the gap is in the cDNA. */
DPC_HD void dpc_bridge_cdna(Bridge &br, const Mat &mL, const Mat &mR, const DevProb &p, const Lanes &ln) {
  const int L2 = mL.L1, L1L = mL.L2, L1R = mR.L2, eb = p.extraband;
  const int rbandL = L1L - L2 + eb, lbandL = eb, rbandR = L1R - L2 + eb, lbandR = eb;
  int bs = DPC_BRIDGE_FLOOR; long long bk = 0x7fffffffffffffffLL;
  int brL = 0, brR = 0, bcL = 0, bcR = 0;
  long long order = 0;
  for (int rL = 1; rL < L2; rL++) {
    int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L1L - 1 ? L1L - 1 : rL + rbandL;
    int nL = chighL - cloL + 1;
    if (nL < 0) nL = 0;
    for (int rR = L2 - rL; rR >= 0; rR--) {
      int pen = (rR == L2 - rL) ? 0 : p.open;                            /* 3092, 3136 */
      int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L1R - 1 ? L1R - 1 : rR + rbandR;
      int nR = chighR - cloR + 1;
      if (nR < 0) nR = 0;
      for (int j = ln.lane; j < nL * nR; j += ln.n) {
        int cL = cloL + j / nR, cR = cloR + j % nR;
        if (!(cR < p.gap - cL)) continue;
        int s = dpc_nscore(mL, rL, cL) + dpc_nscore(mR, rR, cR) + pen;
        long long key = order + j;
        if (s > bs || (s == bs && key < bk)) { bs = s; bk = key; brL = rL; brR = rR; bcL = cL; bcR = cR; }
      }
      order += (long long)nL * nR;
    }
  }
#ifdef __CUDACC__
  for (int o = 16; o > 0; o >>= 1) {
    int s = __shfl_xor_sync(0xffffffffu, bs, o); long long k = __shfl_xor_sync(0xffffffffu, bk, o);
    int a = __shfl_xor_sync(0xffffffffu, brL, o), b = __shfl_xor_sync(0xffffffffu, brR, o);
    int c = __shfl_xor_sync(0xffffffffu, bcL, o), d = __shfl_xor_sync(0xffffffffu, bcR, o);
    if (s > bs || (s == bs && k < bk)) { bs = s; bk = k; brL = a; brR = b; bcL = c; bcR = d; }
  }
#endif
  br.have = bk != 0x7fffffffffffffffLL;
  br.finalscore = bs; br.rL = brL; br.rR = brR; br.cL = bcL; br.cR = bcR; br.introntype = 0;
}

/* ---- arena layout (shared by host sizing and the kernel) --------------------------------- */
struct MatDims {
  int rows, cols, lband, rband, W, wstride, planes, cpl;
  int padL, padR;           /* bytes of valid filler before colch[0] / after the sentinel colch[cols]: the row sweep reads the
                               column code of EVERY diagonal a lane owns, in or out of the matrix (dpc_rows.h) */
};
struct ArenaLayout {
  int nmat;
  int fused;                             /* genome gap whose bridge runs inside the R sweep (no stored R band) */
  MatDims d[2];
  /* two regions: `small` (characters, profiles, bridge tables: read on the fill's critical path, always in shared
   * memory when the problem runs in the shared-memory class) and `bulk` (direction bits, nogap bands, op strings,
   * fallback-fill state: written once per row, read by bridge and traceback; goes to HBM scratch when it does not fit) */
  uint32_t rowch[2], colch[2], di[2], itab, small;
  uint32_t dir[2], nband[2], ops[2], state, bulk;
  uint32_t total;                        /* small + bulk */
};
#define DPC_MAX_CPL 4                       /* row-sweep fill: 1, 2 or 4 diagonals per lane, bands of up to 128 */
DPC_HB uint32_t dpc_al(uint32_t x, uint32_t a) { return (x + a - 1) & ~(a - 1); }

/* kind codes as in include/dynprog_cuda.h: 0 single, 1 genome, 2 cdna, 3 end5, 4 end3.
 * fillmode 0: sizes only (stats); 1: every matrix through the memory-state fill (nibble directions +
 * anti-diagonal state); 2: row-sweep fill (bit planes) wherever the band has at most 128 diagonals. */
DPC_HB void dpc_layout(const DevProb &p, ArenaLayout &a, int fillmode) {
  uint32_t so = 0, bo = 0;
  int maxrows = 0, need_state = fillmode == 1;
  a.nmat = (p.kind == 1 || p.kind == 2) ? 2 : 1;
  /* constant trip counts: with a runtime bound the compiler indexes `a` dynamically and keeps it in local memory */
#pragma unroll
  for (int i = 0; i < 2; i++) {
    if (i >= a.nmat) break;
    MatDims &d = a.d[i];
    if (p.kind == 2) { d.rows = p.L2; d.cols = i ? p.L1R : p.L1; }
    else { d.rows = p.L1; d.cols = i ? p.L2R : p.L2; }
    dpc_bands(d.rows, d.cols, p.extraband, (p.flags & DPC_F_WIDEBAND) != 0, &d.lband, &d.rband);
    d.W = d.lband + d.rband + 1;
    d.wstride = dpc_wstride(d.W);
    d.cpl = d.W <= 32 ? 1 : d.W <= 64 ? 2 : 4;
    d.planes = fillmode == 2 && d.W <= 32 * DPC_MAX_CPL;
    if (fillmode == 2 && !d.planes) need_state = 1;
    if (d.rows > maxrows) maxrows = d.rows;
    /* row sweep with the query as rows: in row r lane l reads column index r + cpl*(l+1) - lband - 1 (0-based) for the
       NEXT row, r = 0..rows: from cpl - lband - 1 up to rows + 32*cpl - lband - 1 */
    d.padL = d.padR = 0;
    if (d.planes && p.kind != 2) {
      d.padL = d.lband + 1 - d.cpl > 0 ? d.lband + 1 - d.cpl : 0;
      const int last = d.rows + 32 * d.cpl - d.lband - 1;
      d.padR = last > d.cols ? last - d.cols : 0;
    }
  }
  /* genome gap, the common case (bands of at most 32 diagonals, integer mode, no known sites): the intron bridge
     runs inside the sweep of the R matrix against the stored band of L, so R's band is never stored */
  a.fused = p.kind == 1 && a.d[0].planes && a.d[1].planes && a.d[0].cpl == 1 && a.d[1].cpl == 1 &&
            !(p.flags & (DPC_F_PROBMODE | DPC_F_KNOWN | DPC_F_INTRONS));
#pragma unroll
  for (int i = 0; i < 2; i++) {
    if (i >= a.nmat) break;
    const MatDims &d = a.d[i];
    a.rowch[i] = so; so = dpc_al(so + (uint32_t)d.rows + 2, 4);
    a.colch[i] = so + (uint32_t)d.padL; so = dpc_al(so + (uint32_t)(d.padL + d.cols + 2 + d.padR), 4);
    a.di[i] = so + (uint32_t)d.padL;
    if (p.kind == 1) so = dpc_al(so + (uint32_t)(d.padL + d.cols + 2 + d.padR), 4);   /* dinucleotide code per column (intron bridge), filler like colch */
    a.dir[i] = bo;
    if (d.planes) bo += (uint32_t)d.rows * (uint32_t)d.cpl * 16; else bo += (uint32_t)d.rows * (uint32_t)d.wstride * 4;
    if (a.nmat == 2 && !(a.fused && i == 1)) { a.nband[i] = bo; bo = dpc_al(bo + (uint32_t)d.rows * (uint32_t)d.W * 2, 16); } else a.nband[i] = 0;
    a.ops[i] = bo; bo = dpc_al(bo + 2 * (uint32_t)(d.rows + d.cols + 2), 16);
  }
  a.state = bo;
  if (need_state) bo += 9 * (uint32_t)(maxrows + 1) * 4;
  a.itab = so;
  if (p.kind == 1) so += 64;
  a.small = dpc_al(so, 16);
  a.bulk = dpc_al(bo, 16);
  a.total = a.small + a.bulk;
}

DPC_HD void dpc_make_mat(Mat &m, const ArenaLayout &a, int i, uint8_t *small, uint8_t *bulk, const DevProb &p, int late, int query_rows) {
  const MatDims &d = a.d[i];
  m.L1 = d.rows; m.L2 = d.cols; m.lband = d.lband; m.rband = d.rband; m.W = d.W; m.wstride = d.wstride;
  m.open = p.open; m.extend = p.extend; m.late = late; m.query_rows = query_rows;
  m.rowch = small + a.rowch[i]; m.colch = small + a.colch[i];
  m.planes = d.planes; m.cpl = d.cpl; m.cplsh = d.cpl == 1 ? 0 : d.cpl == 2 ? 1 : 2;
  m.dir = (uint32_t *)(bulk + a.dir[i]);
  m.nband = (a.nmat == 2 && !(a.fused && i == 1)) ? (int16_t *)(bulk + a.nband[i]) : (int16_t *)0;
}

/* filler around the staged columns of matrix i (MatDims::padL / padR): any valid genome code */
DPC_HD void dpc_stage_pads(const Mat &m, const MatDims &d, const Lanes &ln) {
  for (int i = ln.lane; i < d.padL; i += ln.n) m.colch[-1 - i] = 7;
  for (int i = ln.lane; i < d.padR; i += ln.n) m.colch[d.cols + 1 + i] = 7;
}
/* the same around the dinucleotide codes of the intron bridge: everything outside [0, cols) */
DPC_HD void dpc_stage_dipads(uint8_t *di, const MatDims &d, const Lanes &ln) {
  for (int i = ln.lane; i < d.padL; i += ln.n) di[-1 - i] = 0;
  for (int i = ln.lane; i < d.padR + 2; i += ln.n) di[d.cols + i] = 0;
}

/* ---- one problem ------------------------------------------------------------------------------ */
struct OvfArena { uint16_t *ops; unsigned int *used; unsigned int cap; };

DPC_HD void dpc_emit_ops(DevRes *res, const uint16_t *opsL, int nL, const uint16_t *opsR, int nR,
                         const OvfArena &ovf, uint32_t &status, const Lanes &ln) {
  int total = nL + nR;
  uint16_t *dst = res->ops;
  uint32_t at = 0;
  if (total > DPC_INLINE_OPS) {
    if (ln.lane == 0) {
#ifdef __CUDACC__
      at = atomicAdd(ovf.used, (unsigned int)total);
#else
      at = *ovf.used; *ovf.used += (unsigned int)total;
#endif
    }
    at = (uint32_t)dpc_bcast((int)at, ln);
    if (at + (uint32_t)total > ovf.cap) { status |= DPC_ST_OVF_LOST; total = 0; nL = nR = 0; }
    dst = ovf.ops + at;
  }
  for (int i = ln.lane; i < total; i += ln.n) dst[i] = i < nL ? opsL[i] : opsR[i - nL];
  if (ln.lane == 0) { res->nopsL = (uint16_t)nL; res->nopsR = (uint16_t)nR; res->ovf = at; }
}

/* Solves problem `p` with the fill routine `FILL` (generic here; the CUDA build also has the
 * register/shuffle fill for narrow bands).  `arena` must hold dpc_layout(p).total bytes. */
/* KG selects what is compiled in: 0 single gap, 1 genome gap, 2 cDNA gap, 3 end gaps (and the splice-junction
 * solvers, which run as end gaps), -1 everything (the CPU simulation).  The CUDA build instantiates one kernel per group so that each carries only
 * its own code and register needs. */
/* BULK says where the bulk region lives: 1 in the warp's arena right after the small region, 0 in the problem's HBM
 * scratch, -1 decided per problem (the CPU simulation).  The CUDA build makes it a launch property, so that every
 * access to direction planes, bands and op strings has a known address space (STS/LDS instead of generic ST/LD). */
template <class FILL, int KG, int BULK>
DPC_HD void dpc_solve_problem(const DevProb &p, const uint8_t *pool, const uint32_t *blocks, const DevTables *tb,
                              uint8_t *arena, uint32_t arena_bytes, uint8_t *scratch, DevRes *res, const OvfArena &ovf,
                              uint8_t *gout, FILL &fill, const Lanes &ln) {
  /* gout: output byte stream for the genome characters this problem stages (NULL or p.gout == DPC_NO_GOUT: none) */
  uint8_t *const gch = (gout && p.gout != DPC_NO_GOUT) ? gout + p.gout : (uint8_t *)0;
  /* arena: this warp's shared-memory (or HBM) arena of arena_bytes; scratch: this problem's HBM scratch, used for
   * the bulk region when small + bulk does not fit the arena (the host sized both with the same dpc_layout) */
  /* FILL provides: fillmode (layout), operator() = the matrix fill, walk() = the traceback walk */
  uint32_t status = DPC_ST_DONE;
  Counts ct; ct.nmatches = ct.nmismatches = ct.nopens = ct.nindels = ct.star = 0;
  int finalscore = 0, brL = 0, bcL = 0, brR = 0, bcR = 0, introntype = 0, nopsL = 0, nopsR = 0;
  const int late = (p.flags & DPC_F_LATE) != 0;
  const int8_t *score = &tb->score[p.type][0][0];
  const uint16_t *opsL = 0, *opsR = 0;

  if ((KG == 3 || KG == -1) && (p.kind == 3 || p.kind == 4) && p.endalign == 2) {
    /* QUERYEND_NOGAPS: find_best_endpoint_to_queryend_nogaps 2358-2369 + traceback_nogaps 2815-2872 */
    int n = p.L1 < p.L2 ? p.L1 : p.L2, nm = 0, nmm = 0, star = 0, five = p.kind == 3;
    for (int i = ln.lane; i < n; i += ln.n) {
      int q = pool[five ? p.q0 + (uint32_t)(p.L1 - 1 - i) : p.q0 + (uint32_t)i];
      int g = dpc_genomic_code(p, blocks, five ? p.off2 - i : p.off2 + i);
      if (gch) gch[i] = (uint8_t)dpc_code_char(g);
      if (g == DPC_GSTAR) star++;
      else if (dpc_query_uc(q) == dpc_code_char(g)) nm++;
      else if ((tb->cons[q & 127] >> g) & 1) nm++;
      else nmm++;
    }
    ct.nmatches = dpc_warp_sum(nm, ln); ct.nmismatches = dpc_warp_sum(nmm, ln); ct.star = dpc_warp_sum(star, ln);
    finalscore = 3 * ct.nmatches - 5 * ct.nmismatches;                    /* 5243, 5700 */
    brL = bcL = n;
    /* one run of n aligned columns; an op carries 14 bits of length, so a long unaligned end (QUERYEND_NOGAPS is not
       clipped to maxlength1/2) becomes several M ops -- the host accepts at most DPC_INLINE_OPS of them */
    const int nrun = (n + DPC_OP_MAXLEN - 1) / DPC_OP_MAXLEN;
    for (int k = ln.lane; k < nrun; k += ln.n) {
      const int len = n - k * DPC_OP_MAXLEN < DPC_OP_MAXLEN ? n - k * DPC_OP_MAXLEN : DPC_OP_MAXLEN;
      res->ops[k] = (uint16_t)((len << 2) | DPC_OP_M);
    }
    if (ln.lane == 0) { res->nopsL = (uint16_t)nrun; res->nopsR = 0; res->ovf = 0; }
    status |= DPC_ST_HAVE | DPC_ST_OK;
  } else {
    ArenaLayout a;
    dpc_layout(p, a, FILL::fillmode);
    Mat m0, m1;
    uint8_t *bulk = BULK == 1 ? arena + a.small : BULK == 0 ? scratch : (a.total <= arena_bytes ? arena + a.small : scratch);
    int32_t *st = (int32_t *)(bulk + a.state);
    if (KG == 0 || KG == 3 || (KG == -1 && (p.kind == 0 || p.kind == 3 || p.kind == 4))) {
      /* Dynprog_single_gap 4450-4572, Dynprog_end5_gap 5094-5284, Dynprog_end3_gap 5556-5741 */
      const int kind = KG == 0 ? 0 : p.kind;              /* the single-gap kernel carries no end-gap code */
      const int five = kind == 3;
      dpc_make_mat(m0, a, 0, arena, bulk, p, five ? !late : late, 1);
      for (int i = ln.lane; i < p.L1; i += ln.n) {
        int q = pool[five ? p.q0 + (uint32_t)(p.L1 - 1 - i) : p.q0 + (uint32_t)i];
        m0.rowch[i] = (uint8_t)(q & 127);            /* query bytes are < 128 (checked when the problem is accepted) */
      }
      if (p.flags & DPC_F_SEQ2) {
        /* Dynprog_end5/3_splicejunction 5411-5552, 5869-6012: use_genomicseg_p, sequence2 = the splice-junction string */
        for (int i = ln.lane; i < p.L2; i += ln.n) m0.colch[i] = pool[p.q1 + (uint32_t)i];
      } else {
        for (int i = ln.lane; i < p.L2; i += ln.n) {
          const int g = dpc_genomic_code(p, blocks, five ? p.off2 - i : p.off2 + i);
          m0.colch[i] = (uint8_t)g;
          if (gch) gch[i] = (uint8_t)dpc_code_char(g);
        }
      }
      if (ln.lane == 0) m0.colch[p.L2] = 7;                               /* sentinel one past the last column */
      dpc_stage_pads(m0, a.d[0], ln);
      DPC_SYNC();
      EndSearch es; es.eb = p.extraband;
      if (kind == 0) { es.mode = 3; es.best.score = -2147483647; es.best.key = 0; }
      else if (p.endalign == 1) { es.mode = 2; es.best.score = DPC_NEG; es.best.key = p.L1 * (p.L2 + 1); }
      else { es.mode = 1; es.best.score = 0; es.best.key = 0; }
      fill(m0, st, score, es, ln);                                      /* leaves es.best reduced over the lanes */
      finalscore = es.best.score;
      brL = es.best.key / (p.L2 + 1); bcL = es.best.key % (p.L2 + 1);
      uint16_t *ops = (uint16_t *)(bulk + a.ops[0]);
      nopsL = fill.walk(m0, brL, bcL, five, p.cdna_direction, ops, ln);
      dpc_count_ops(m0, brL, bcL, ops, nopsL, ct, tb, ln);
      opsL = ops;
      status |= DPC_ST_HAVE | DPC_ST_OK;
    } else {
      const int cdna = KG == 2 || (KG == -1 && p.kind == 2);
      dpc_make_mat(m0, a, 0, arena, bulk, p, late, !cdna);
      dpc_make_mat(m1, a, 1, arena, bulk, p, !late, !cdna);
      if (!cdna) {
        /* Dynprog_genome_gap 4798-5061: L = fwd(query, genome @ offset2L), R = rev(query, genome @ revoffset2R) */
        for (int i = ln.lane; i < p.L1; i += ln.n) {
          int qf = pool[p.q0 + (uint32_t)i], qr = pool[p.q0 + (uint32_t)(p.L1 - 1 - i)];
          m0.rowch[i] = (uint8_t)(qf & 127); m1.rowch[i] = (uint8_t)(qr & 127);
        }
        for (int i = ln.lane; i < p.L2; i += ln.n) {
          const int g = dpc_genomic_code(p, blocks, p.off2 + i);
          m0.colch[i] = (uint8_t)g;
          if (gch) gch[i] = (uint8_t)dpc_code_char(g);
        }
        for (int i = ln.lane; i < p.L2R; i += ln.n) {
          const int g = dpc_genomic_code(p, blocks, p.off2R - i);
          m1.colch[i] = (uint8_t)g;
          if (gch) gch[dpc_gout_span(p.L2) + (uint32_t)i] = (uint8_t)dpc_code_char(g);
        }
        if (ln.lane == 0) { m0.colch[p.L2] = 7; m1.colch[p.L2R] = 7; }    /* leftdi[length2L-1] = rightdi[length2R-1] = 0, 3354, 3376 */
        dpc_stage_pads(m0, a.d[0], ln); dpc_stage_pads(m1, a.d[1], ln);
      } else {
        /* Dynprog_cdna_gap 4577-4793: rows = genome (fwd from offset2 / rev from offset2+length2-1) */
        for (int i = ln.lane; i < p.L2; i += ln.n) {
          m0.rowch[i] = (uint8_t)dpc_genomic_code(p, blocks, p.off2 + i);
          m1.rowch[i] = (uint8_t)dpc_genomic_code(p, blocks, p.off2 + p.L2 - 1 - i);
        }
        for (int i = ln.lane; i < p.L1; i += ln.n) m0.colch[i] = pool[p.q0 + (uint32_t)i];
        for (int i = ln.lane; i < p.L1R; i += ln.n) m1.colch[i] = pool[p.q1 - (uint32_t)i];
        if (ln.lane == 0) { m0.colch[p.L1] = 0; m1.colch[p.L1R] = 0; }   /* sentinel one past the last column */
      }
      DPC_SYNC();
      EndSearch es; es.mode = 0; es.eb = 0; es.best.score = 0; es.best.key = 0;
      Bridge br;
      int ok;
      if (!cdna && a.fused) {
        /* both sweeps and the bridge in one pass (no stored R band) */
        uint8_t *ldi = arena + a.di[0], *rdi = arena + a.di[1];
        int8_t *itab = (int8_t *)(arena + a.itab);
        /* filler on both sides of the dinucleotide codes (the fused bridge reads one per lane and row, in or out of
           the matrix): everything outside [0, length2) */
        dpc_stage_dipads(ldi, a.d[0], ln); dpc_stage_dipads(rdi, a.d[1], ln);
        dpc_bridge_tables(m0, m1, p, ldi, rdi, itab, ln);
        Best best; best.score = DPC_BRIDGE_FLOOR; best.key = 0x7fffffff;
        fill.pair_bridge(m0, m1, score, p, ldi, rdi, itab, best, ln);
        dpc_bridge_finish(br, best, 0.0, 0x7fffffff, m0, m1, p, (const uint8_t *)0, (const uint8_t *)0, (const uint8_t *)0, ln);
        ok = br.have && br.finalscore >= 0;                               /* 4083-4101 */
      } else if (!cdna) {
        fill.pair(m0, m1, st, score, es, ln);
        const uint8_t *aux = pool + p.aux, *lknown = 0, *rknown = 0;
        const double *lp = 0, *rp = 0;
        const uint8_t *introns = 0;
        if (p.flags & DPC_F_PROBMODE) { lp = (const double *)aux; rp = lp + p.L2 + 1; aux += 8 * (uint32_t)(p.L2 + p.L2R + 2); }
        if (p.flags & DPC_F_KNOWN) { lknown = aux; rknown = aux + p.L2 + 1; aux += (uint32_t)(p.L2 + p.L2R + 2); }
        if (p.flags & DPC_F_INTRONS) introns = pool + (((uint32_t)(aux - pool) + 3u) & ~3u);
        dpc_bridge_intron(br, m0, m1, p, lknown, rknown, lp, rp, introns, arena + a.di[0], arena + a.di[1], (int8_t *)(arena + a.itab), ln);
        ok = br.have && br.finalscore >= 0;                               /* 4083-4101 */
        if (ok && !(p.flags & (DPC_F_NOVEL | DPC_F_INTRONS)) && (p.flags & DPC_F_KNOWN) && (!lknown[br.cL] || !rknown[br.cR])) ok = 0;
      } else {
        fill.pair(m0, m1, st, score, es, ln);
        dpc_bridge_cdna(br, m0, m1, p, ln);
        ok = br.have;
      }
      finalscore = br.finalscore; brL = br.rL; bcL = br.cL; brR = br.rR; bcR = br.cR; introntype = br.introntype;
      if (br.have) status |= DPC_ST_HAVE;
      if (ok) {
        status |= DPC_ST_OK;
        uint16_t *oR = (uint16_t *)(bulk + a.ops[1]), *oL = (uint16_t *)(bulk + a.ops[0]);
        nopsR = fill.walk(m1, brR, bcR, 1, p.cdna_direction, oR, ln);
        dpc_count_ops(m1, brR, bcR, oR, nopsR, ct, tb, ln);
        nopsL = fill.walk(m0, brL, bcL, 0, p.cdna_direction, oL, ln);
        dpc_count_ops(m0, brL, bcL, oL, nopsL, ct, tb, ln);
        opsL = oL; opsR = oR;
      }
    }
    dpc_emit_ops(res, opsL, nopsL, opsR, nopsR, ovf, status, ln);
  }
  if (ct.star) status |= DPC_ST_STAR;
  if (ln.lane == 0) {
    res->finalscore = finalscore;
    res->nmatches = ct.nmatches; res->nmismatches = ct.nmismatches; res->nopens = ct.nopens; res->nindels = ct.nindels;
    res->bestrL = brL; res->bestcL = bcL; res->bestrR = brR; res->bestcR = bcR;
    res->introntype = introntype; res->status = status;
  }
}

struct GenericFill {
  enum { fillmode = 1 };
  DPC_HDM void operator()(const Mat &m, int32_t *st, const int8_t *score, EndSearch &es, const Lanes &ln) const {
    dpc_fill_generic(m, st, score, es, ln);
    dpc_warp_best(es.best, m.late, ln);
  }
  DPC_HDM void pair(const Mat &mA, const Mat &mB, int32_t *st, const int8_t *score, EndSearch &es, const Lanes &ln) const {
    dpc_fill_generic(mA, st, score, es, ln);
    dpc_fill_generic(mB, st, score, es, ln);
  }
  DPC_HDM void pair_bridge(const Mat &, const Mat &, const int8_t *, const DevProb &, const uint8_t *, const uint8_t *, const int8_t *,
                           Best &, const Lanes &) const {}      /* never asked for: fused layouts need the row sweep */
  DPC_HDM int walk(const Mat &m, int r, int c, int revp, int cdna_direction, uint16_t *ops, const Lanes &ln) const {
    return dpc_walk_serial(m, r, c, revp, cdna_direction, ops, ln);
  }
};
#endif /* DPC_CORE_H */
