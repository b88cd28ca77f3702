/* oracle/dynprog_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the gap-fill path of the
 * reference's src/dynprog.c (GMAP/GSNAP 2012-07-03), written against the same
 * problem/result structs as the C ABI in include/dynprog_cuda.h so that one
 * set of inputs can be pushed through (1) the compiled reference
 * (oracle/_ref/libdynprog_ref.so, built from /root/reference by
 * oracle/Makefile), (2) this restatement and (3) the CUDA library.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this file's library.  Parity is PINNED: tests/test_oracle_vs_ref.py
 * checks this restatement against the compiled reference on every kind of
 * problem, and tests/golden/ holds vectors generated from the compiled
 * reference by tests/golden/make_golden.py.
 *
 * Every function cites the reference lines it restates (all in
 * /root/reference/src/dynprog.c unless another file is named).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ctype.h>
#include "../include/dynprog_cuda.h"

#define NEG DPC_NEG_INFINITY
enum { D_STOP = 0, D_DIAG = 1, D_HORIZ = 2, D_VERT = 4 };   /* dynprog.c:308-312 */
enum { HIGHQ = 0, MEDQ = 1, LOWQ = 2, ENDQ = 3 };           /* dynprog.c:150 */

/* ---- tables: pairdistance_init, dynprog.c:1127-1226 ---------------------- */
static int P[4][128][128];
static unsigned char CONS[128][128];
static int g_maxlength1 = 611, g_maxlength2 = 2000;
static dpc_setup_t g_setup;
static int g_have_setup = 0;

static void both_cases(int a, int b, int score, int oneway) {
  /* permute_cases / permute_cases_oneway, dynprog.c:1053-1124 */
  int A[2] = { a, tolower(a) }, B[2] = { b, tolower(b) };
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) {
      CONS[A[i]][B[j]] = 1;
      for (int t = 0; t < 4; t++) P[t][A[i]][B[j]] = score;
      if (!oneway) {
        CONS[B[j]][A[i]] = 1;
        for (int t = 0; t < 4; t++) P[t][B[j]][A[i]] = score;
      }
    }
}

static void build_tables(int mode) {
  static const int mm[4] = { -3, -2, -1, -5 };   /* dynprog.c:169-179 */
  memset(P, 0, sizeof P);
  memset(CONS, 0, sizeof CONS);
  for (int c1 = 'A'; c1 <= 'z'; c1++)            /* dynprog.c:1150-1157: note c2 < 'z' */
    for (int c2 = 'A'; c2 < 'z'; c2++)
      for (int t = 0; t < 4; t++) P[t][c1][c2] = mm[t];
  both_cases('U', 'T', 3, 0);                     /* 1169 */
  static const char *half[] = { "RA", "RG", "YT", "YC", "WA", "WT", "SG", "SC", "MA", "MC", "KG", "KT", 0 };
  for (int i = 0; half[i]; i++) both_cases(half[i][0], half[i][1], 1, 0);          /* 1171-1187 */
  static const char *amb[] = { "HA", "HT", "HC", "BG", "BC", "BT", "VG", "VA", "VC", "DG", "DA", "DT",
                               "NT", "NC", "NA", "NG", "XT", "XC", "XA", "XG", 0 };
  for (int i = 0; amb[i]; i++) both_cases(amb[i][0], amb[i][1], -1, 0);            /* 1189-1213 */
  if (mode == DPC_MODE_CMET_STRANDED || mode == DPC_MODE_CMET_NONSTRANDED) {      /* 1215-1219 */
    both_cases('T', 'C', 3, 1);
    both_cases('A', 'G', 3, 1);
  }
  for (int c = 'A'; c < 'Z'; c++) both_cases(c, c, 3, 0);                           /* 1221-1223 */
}

int port_init(int maxlookback, int extraquerygap, int maxpeelback,
              int extramaterial_end, int extramaterial_paired, int mode) {
  /* compute_maxlengths, dynprog.c:831-852 */
  int m1 = maxlookback + maxpeelback;
  if (m1 < 500) m1 = 500;
  int m2 = m1 + extraquerygap + (extramaterial_end > extramaterial_paired ? extramaterial_end : extramaterial_paired);
  if (m2 < 2000) m2 = 2000;
  g_maxlength1 = m1;
  g_maxlength2 = m2;
  build_tables(mode);
  return 0;
}

int port_setup(const dpc_setup_t *s) { g_setup = *s; g_have_setup = 1; return 0; }
int port_pairdistance(int type, int c1, int c2) { return P[type][c1 & 127][c2 & 127]; }
int port_consistent(int c1, int c2) { return CONS[c1 & 127][c2 & 127]; }

/* ---- genome access ------------------------------------------------------- */
static char genome_char(uint32_t pos) {
  /* uncompress_one_char, genome.c:9325-9362 (little endian) */
  const uint32_t *b = g_setup.genome_blocks + (uint64_t)(pos / 32U) * 3;
  int bit = pos % 32;
  if (b[2] & (1U << bit)) return 'N';
  uint32_t w = bit < 16 ? b[1] >> (2 * bit) : b[0] >> (2 * bit - 32);
  return "ACGT"[w & 3];
}

static char compl_nt(char c) {
  /* complCode = COMPLEMENT_LC restricted to what genome_char can return, complement.h:31 */
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

static char genomic_nt(int genomicpos, const dpc_problem_t *p) {
  /* get_genomic_nt, dynprog.c:403-441 */
  uint32_t pos;
  if (genomicpos < 0) return '*';
  if ((uint32_t)genomicpos >= p->genomiclength) return '*';   /* int >= Genomicpos_T compares unsigned */
  pos = p->chroffset + p->chrpos;
  if (pos < p->chroffset) return '*';
  if (pos >= p->chrhigh) return '*';
  if (p->watsonp) return genome_char(p->chroffset + p->chrpos + (uint32_t)genomicpos);
  return compl_nt(genome_char(p->chroffset + p->chrpos + (p->genomiclength - 1) - (uint32_t)genomicpos));
}

static char query_uc(char c) {
  /* sequenceuc = UPPERCASE_U2T of the query, sequence.c:815,962 / complement.h:36 */
  int u = toupper((unsigned char)c);
  return (char)(u == 'U' ? 'T' : u);
}

/* ---- matrices -------------------------------------------------------------- */
typedef struct {
  int L1, L2;               /* rows, cols */
  int *N, *G1, *G2;         /* nogap, gap1, gap2  (struct Int3_T, dynprog.c:482) */
  unsigned char *dN, *dG1, *dG2;
} Mat;
#define AT(m, r, c) ((size_t)(r) * ((m)->L2 + 1) + (c))

static void mat_free(Mat *m) {
  free(m->N); free(m->G1); free(m->G2); free(m->dN); free(m->dG1); free(m->dG2);
  memset(m, 0, sizeof *m);
}

static void bands(int L1, int L2, int extraband, int widebandp, int *lband, int *rband) {
  /* dynprog.c:1442-1454 */
  if (!widebandp) { *lband = *rband = extraband; }
  else if (L2 >= L1) { *rband = L2 - L1 + extraband; *lband = extraband; }
  else { *lband = L1 - L2 + extraband; *rband = extraband; }
}

/* One banded 3-state Gotoh fill.  rowch/colch are in matrix order (row r <->
 * rowch[r-1]).  query_rows != 0: rows are the query (compute_scores_lookup_fwd /
 * _rev, dynprog.c:1424-1736); else rows are the genome (_fwd_12 / _rev_12,
 * 1741-2044).  The score is always pairdistance[query][genome] (1560, 1872).
 * All four reference variants leave the same matrix: Matrix3_alloc memset
 * (489-521), row/column 0 (1460-1488), the forced -inf cells just outside the
 * band (1501-1513 / 1813-1825) and the in-band recurrence (1516-1562). */
static void fill(Mat *m, int L1, int L2, const char *rowch, const char *colch, int query_rows,
                 int type, int open, int extend, int extraband, int widebandp, int late) {
  int lband, rband;
  size_t n = (size_t)(L1 + 1) * (L2 + 1);
  m->L1 = L1; m->L2 = L2;
  m->N = calloc(n, sizeof(int)); m->G1 = calloc(n, sizeof(int)); m->G2 = calloc(n, sizeof(int));
  m->dN = calloc(n, 1); m->dG1 = calloc(n, 1); m->dG2 = calloc(n, 1);
  bands(L1, L2, extraband, widebandp, &lband, &rband);

  m->N[0] = 0; m->G1[0] = m->G2[0] = NEG;
  for (int c = 1; c <= rband && c <= L2; c++) {
    m->N[AT(m, 0, c)] = NEG; m->G2[AT(m, 0, c)] = NEG;
    m->G1[AT(m, 0, c)] = open + c * extend;
    m->dG1[AT(m, 0, c)] = D_HORIZ;
  }
  m->dG1[AT(m, 0, 1)] = D_STOP;
  for (int r = 1; r <= lband && r <= L1; r++) {
    m->N[AT(m, r, 0)] = NEG; m->G1[AT(m, r, 0)] = NEG;
    m->G2[AT(m, r, 0)] = open + r * extend;
    m->dG2[AT(m, r, 0)] = D_VERT;
  }
  m->dG2[AT(m, 1, 0)] = D_STOP;

  /* cells one step outside the band that the recurrence reads */
  for (int c = 1; c <= L2; c++) {
    int rlo = c - rband, rhigh = c + lband;
    if (rlo >= 1 && rlo - 1 <= L1) { m->G2[AT(m, rlo - 1, c)] = NEG; m->N[AT(m, rlo - 1, c)] = NEG; }
    if (rhigh <= L1) { m->G1[AT(m, rhigh, c - 1)] = NEG; m->N[AT(m, rhigh, c - 1)] = NEG; }
  }

  for (int c = 1; c <= L2; c++) {
    int rlo = c - rband < 1 ? 1 : c - rband;
    int rhigh = c + lband > L1 ? L1 : c + lband;
    for (int r = rlo; r <= rhigh; r++) {
      int best, s; unsigned char dir;
      int q = query_rows ? rowch[r - 1] : colch[c - 1];
      int g = query_rows ? colch[c - 1] : rowch[r - 1];

      best = m->N[AT(m, r, c - 1)] + open; dir = D_DIAG;               /* gap1, 1519-1529 */
      s = m->G1[AT(m, r, c - 1)];
      if (s > best || (s == best && late)) { best = s; dir = D_HORIZ; }
      m->G1[AT(m, r, c)] = best + extend; m->dG1[AT(m, r, c)] = dir;

      best = m->N[AT(m, r - 1, c)] + open; dir = D_DIAG;               /* gap2, 1532-1542 */
      s = m->G2[AT(m, r - 1, c)];
      if (s > best || (s == best && late)) { best = s; dir = D_VERT; }
      m->G2[AT(m, r, c)] = best + extend; m->dG2[AT(m, r, c)] = dir;

      best = m->N[AT(m, r - 1, c - 1)]; dir = D_DIAG;                  /* nogap, 1545-1561 */
      s = m->G1[AT(m, r - 1, c - 1)];
      if (s > best || (s == best && late)) { best = s; dir = D_HORIZ; }
      s = m->G2[AT(m, r - 1, c - 1)];
      if (s > best || (s == best && late)) { best = s; dir = D_VERT; }
      m->N[AT(m, r, c)] = best + P[type][q & 127][g & 127]; m->dN[AT(m, r, c)] = dir;
    }
  }
}

/* ---- pair stack (push order) ---------------------------------------------- */
typedef struct { dpc_pair_t *v; int n, cap; } Stack;
static void push(Stack *s, int qpos, int gpos, char cdna, char comp, char genome, int idx, int gapp) {
  if (s->n == s->cap) { s->cap = s->cap ? 2 * s->cap : 256; s->v = realloc(s->v, s->cap * sizeof *s->v); }
  dpc_pair_t *p = &s->v[s->n++];
  p->querypos = qpos; p->genomepos = gpos; p->dynprogindex = idx;
  p->cdna = cdna; p->comp = comp; p->genome = genome; p->gapp = (uint8_t)gapp;
}
static void push_gapholder(Stack *s) { push(s, -1, -1, ' ', ' ', ' ', 0, 1); }   /* pairpool.c:352-401 */

static int intron_type(char l1, char l2, char r2, char r1, int cdna_direction) {
  /* Intron_type, intron.c:17-190 (non-PMAP, no INTRON_HELP) */
  int left, right, t;
  if (l1 == 'G' && l2 == 'T') left = 0x21; else if (l1 == 'G' && l2 == 'C') left = 0x10;
  else if (l1 == 'A' && l2 == 'T') left = 0x08; else if (l1 == 'C' && l2 == 'T') left = 0x06; else return 0;
  if (r2 == 'A' && r1 == 'G') right = 0x30; else if (r2 == 'A' && r1 == 'C') right = 0x0C;
  else if (r2 == 'G' && r1 == 'C') right = 0x02; else if (r2 == 'A' && r1 == 'T') right = 0x01; else return 0;
  t = left & right;
  if (t == 0) return 0;
  if (cdna_direction > 0) return t < 0x08 ? 0 : t;
  if (cdna_direction < 0) return t > 0x04 ? 0 : t;
  return 0;
}

typedef struct { int nmatches, nmismatches, nopens, nindels; } Counts;

/* traceback (2611-2712), traceback_cdna (2715-2810) and the add_*skip helpers
 * (2372-2601) in one routine.  Matrix coordinates -> sequence coordinates:
 * index i (0-based, matrix order) of the query sits at q0 + qs*i, of the
 * genome at g0 + gs*i, with qs = gs = -1 for the "rev" callers.
 * genome_rows != 0 is the cDNA layout (rows = genome). */
/* endc < 0: traceback / traceback_cdna, which stop on a STOP direction.  endc >= 0: traceback_local (2875-2969),
 * which runs while c > endc, has no '*' test, and hands the reached (r,c) back through rp / cp. */
static void traceback_from(Stack *st, Counts *ct, const Mat *m, int *rp, int *cp, int endc,
                           const char *qch, const char *gch, int q0, int g0, int revp, int genome_rows,
                           int cdna_direction, int idx) {
  int step = revp ? -1 : 1, r = *rp, c = *cp;
  const int local = endc >= 0;
  while (local ? c > endc : m->dN[AT(m, r, c)] != D_STOP) {
    int qi = genome_rows ? c - 1 : r - 1, gi = genome_rows ? r - 1 : c - 1;
    char c1 = qch[qi], c2 = gch[gi];
    int consistent = genome_rows ? CONS[c2 & 127][c1 & 127] : CONS[c1 & 127][c2 & 127];   /* 2654 vs 2752 */
    if (!genome_rows && !local && c2 == '*') {
      /* 2644: no pair past the end of the chromosome (traceback_cdna has no such test) */
    } else if (query_uc(c1) == c2) { ct->nmatches++; push(st, q0 + step * qi, g0 + step * gi, c1, '*', c2, idx, 0); }
    else if (consistent) { ct->nmatches++; push(st, q0 + step * qi, g0 + step * gi, c1, ':', c2, idx, 0); }
    else { ct->nmismatches++; push(st, q0 + step * qi, g0 + step * gi, c1, ' ', c2, idx, 0); }

    unsigned char d = m->dN[AT(m, r, c)];
    if (d == D_DIAG) { r--; c--; continue; }
    int dist = 1, genome_run;
    r--; c--;
    if (d == D_HORIZ) { while (m->dG1[AT(m, r, c)] == D_HORIZ) { dist++; c--; } c--; genome_run = !genome_rows; }
    else { while (m->dG2[AT(m, r, c)] == D_VERT) { dist++; r--; } r--; genome_run = genome_rows; }
    /* After the moves (r,c) is the nogap cell the run hangs off.  The run covers
     * `dist` residues of one sequence starting at matrix index (run axis) lo..lo+dist-1. */
    if (genome_run) {
      /* add_genomeskip 2416-2512 / add_genomeskip_cdna 2515-2601 */
      int lo = genome_rows ? r : c;                 /* 0-based first genome index of the run */
      int qi2 = genome_rows ? c - 1 : r - 1;        /* query index of the anchoring cell */
      int qpos = revp ? q0 - qi2 : q0 + qi2 + 1;    /* 2432-2442: fwd advances to next querypos */
      int dashes = 1;
      if (dist >= 9) {                              /* MICROINTRON_LENGTH, 2444 */
        /* dinucleotides at the two ends of the run, in genomic (left-to-right) order */
        char a, b, y, z;
        if (!revp) { a = gch[lo]; b = gch[lo + 1]; y = gch[lo + dist - 2]; z = gch[lo + dist - 1]; }
        else { a = gch[lo + dist - 1]; b = gch[lo + dist - 2]; y = gch[lo + 1]; z = gch[lo]; }
        dashes = intron_type(a, b, y, z, cdna_direction) == 0;
      }
      if (dashes) {
        /* fwd: from the right end backwards; rev: matrix index descending too (genomic coord ascending) */
        for (int j = 0; j < dist; j++) {
          int gi2 = lo + dist - 1 - j;
          push(st, qpos, g0 + step * gi2, ' ', '-', gch[gi2], idx, 0);
        }
        ct->nopens++; ct->nindels += dist;
      } else {
        push_gapholder(st);                        /* 2507 */
      }
    } else {
      /* add_queryskip 2372-2413 */
      int lo = genome_rows ? c : r;                 /* 0-based first query index of the run */
      int gi2 = genome_rows ? r - 1 : c - 1;
      int gpos = revp ? g0 - gi2 : g0 + gi2 + 1;    /* 2387-2395: fwd advances to next genomepos */
      for (int j = 0; j < dist; j++) {
        int qi2 = lo + dist - 1 - j;
        push(st, q0 + step * qi2, gpos, qch[qi2], '-', ' ', idx, 0);
      }
      ct->nopens++; ct->nindels += dist;
    }
  }
  *rp = r; *cp = c;
}

static void traceback(Stack *st, Counts *ct, const Mat *m, int r, int c,
                      const char *qch, const char *gch, int q0, int g0, int revp, int genome_rows,
                      int cdna_direction, int idx) {
  traceback_from(st, ct, m, &r, &c, -1, qch, gch, q0, g0, revp, genome_rows, cdna_direction, idx);
}

static void traceback_nogaps(Stack *st, Counts *ct, int r, int c, const char *qch, const char *gch,
                             int q0, int g0, int revp, int idx) {
  /* traceback_nogaps, 2815-2872 */
  int step = revp ? -1 : 1;
  while (r > 0 && c > 0) {
    char c1 = qch[r - 1], c2 = gch[c - 1];
    if (c2 == '*') { }
    else if (query_uc(c1) == c2) { ct->nmatches++; push(st, q0 + step * (r - 1), g0 + step * (c - 1), c1, '*', c2, idx, 0); }
    else if (CONS[c1 & 127][c2 & 127]) { ct->nmatches++; push(st, q0 + step * (r - 1), g0 + step * (c - 1), c1, ':', c2, idx, 0); }
    else { ct->nmismatches++; push(st, q0 + step * (r - 1), g0 + step * (c - 1), c1, ' ', c2, idx, 0); }
    r--; c--;
  }
}

/* ---- helpers --------------------------------------------------------------- */
static void quality(double defect_rate, int *type) {
  *type = defect_rate < 0.003 ? HIGHQ : defect_rate < 0.014 ? MEDQ : LOWQ;   /* dynprog.h:27-28 */
}
static int bump(int idx) { return idx + (idx > 0 ? 1 : -1); }   /* e.g. 4570 */

static void result_init(dpc_result_t *r, const dpc_problem_t *p) {
  r->null_list = 1; r->dynprogindex_out = p->dynprogindex;
  r->finalscore = r->nmatches = r->nmismatches = r->nopens = r->nindels = DPC_UNSET;
  r->new_leftgenomepos = r->new_rightgenomepos = r->exonhead = r->introntype = DPC_UNSET;
  r->incompletep = DPC_UNSET; r->npairs = 0; r->reserved = 0;
  r->left_prob = r->right_prob = -1.0;
}

static char *gather_query(const char *seq, int len, int rev) {
  char *b = malloc(len + 1);
  for (int i = 0; i < len; i++) b[i] = rev ? seq[-i] : seq[i];
  b[len] = 0;
  return b;
}
static char *gather_genome(const dpc_problem_t *p, int start, int len, int rev) {
  char *b = malloc(len + 1);
  for (int i = 0; i < len; i++) b[i] = genomic_nt(rev ? start - i : start + i, p);
  b[len] = 0;
  return b;
}

static void emit(Stack *out, const dpc_pair_t *v, int n, int reversed) {
  for (int i = 0; i < n; i++) {
    const dpc_pair_t *p = &v[reversed ? n - 1 - i : i];
    push(out, p->querypos, p->genomepos, p->cdna, p->comp, p->genome, p->dynprogindex, p->gapp);
  }
}

/* ---- Dynprog_microexon_int, 7127-7429 (+ make_microexon_pairs_double 6941-7053, BoyerMoore_nt boyer-moore.c:384) ---- */
static char bm_nt(const dpc_problem_t *p, int pos) {
  /* boyer-moore.c:357-381: its own get_genomic_nt, without the segment bounds of dynprog.c:415-419 */
  if (p->watsonp) return genome_char(p->chroffset + p->chrpos + (uint32_t)pos);
  return compl_nt(genome_char(p->chroffset + p->chrpos + (p->genomiclength - 1) - (uint32_t)pos));
}
static void micro_segment(Stack *st, const dpc_problem_t *p, int off1, int off2, int len, int idx) {
  for (int i = 0; i < len; i++) {
    char c1 = p->seq1[off1 - p->offset1 + i], c2 = genomic_nt(off2 + i, p);
    char comp = query_uc(c1) == c2 ? '*' : CONS[c1 & 127][c2 & 127] ? ':' : ' ';
    push(st, off1 + i, off2 + i, c1, comp, c2, idx, 0);
  }
}
static int solve_microexon(const dpc_problem_t *p, dpc_result_t *res, Stack *out) {
  const int L1 = p->length1, lo = p->offset2, ro = p->offset2R;
  const double pvalue = p->defect_rate < 0.003 ? 0.01 : p->defect_rate < 0.014 ? 0.001 : 0.0001;     /* 128-130, 7160-7166 */
  char i1, i2, i3, i4, gapchar;
  int bestcL = -1, bestcR = -1, bestmid = 0, candidate = 0;
  double bestprob = 0.0;
  res->left_prob = res->right_prob = 0.0;
  if (p->cdna_direction > 0) { i1 = 'G'; i2 = 'T'; i3 = 'A'; i4 = 'G'; gapchar = '>'; res->introntype = 0x20; }
  else if (p->cdna_direction < 0) { i1 = 'C'; i2 = 'T'; i3 = 'A'; i4 = 'C'; gapchar = '<'; res->introntype = 0x04; }
  else return DPC_ERR_ARG;                                                  /* abort(), 7192 */
  const int span = ro - lo;
  if (span <= 0 || L1 <= 0) return DPC_ERR_ARG;                             /* abort(), 7215 */
  int minlen = (int)ceil(-log(1.0 - pow(1.0 - pvalue, 1.0 / (double)span)) / log(4));
  minlen -= 8;
  if (minlen > 12) { res->introntype = 0; return 0; }                       /* MAX_MICROEXON_LENGTH, 7222-7227 */
  if (minlen < 3) minlen = 3;
  int leftbound = 0, rightbound = 0, nmm = 0, i;
  while (leftbound < L1 - 1 && nmm <= 1) { if (query_uc(p->seq1[leftbound]) != genomic_nt(lo + leftbound, p)) nmm++; leftbound++; }
  leftbound--;
  i = L1 - 1; nmm = 0;
  while (i >= 0 && nmm <= 1) { if (query_uc(p->seq1[i]) != genomic_nt(ro - rightbound, p)) nmm++; rightbound++; i--; }
  rightbound--;
  for (int cL = 1; cL <= leftbound; cL++) {
    if (!(genomic_nt(lo + cL, p) == i1 && genomic_nt(lo + cL + 1, p) == i2)) continue;
    int mincR = L1 - 12 - cL, maxcR = L1 - minlen - cL;
    if (mincR < 1) mincR = 1;
    if (maxcR > rightbound) maxcR = rightbound;
    for (int cR = mincR; cR <= maxcR; cR++) {
      if (!(genomic_nt(ro - cR - 1, p) == i3 && genomic_nt(ro - cR, p) == i4)) continue;
      const int mid = L1 - cL - cR, textleft = lo + cL + 9, textright = ro - cR - 9, textlen = textright - textleft;
      int okay = 1;
      for (int k = 0; k < mid; k++) { char c = query_uc(p->seq1[cL + k]); if (c != 'A' && c != 'C' && c != 'G' && c != 'T') okay = 0; }
      if (!okay) continue;
      /* BoyerMoore_nt pushes its hits on a list: they come back highest position first */
      for (int j = textlen - mid; j >= 0; j--) {
        int k = 0;
        while (k < mid && query_uc(p->seq1[cL + k]) == bm_nt(p, textleft + j + k)) k++;
        if (k < mid) continue;
        candidate = textleft + j;                                           /* sic: assigned for every hit, 7326 */
        if (genomic_nt(candidate - 2, p) == i3 && genomic_nt(candidate - 1, p) == i4 &&
            genomic_nt(candidate + mid, p) == i1 && genomic_nt(candidate + mid + 1, p) == i2) {
          uint32_t s2, s3; int w2, w3;
          if (p->watsonp) {
            s2 = p->chrpos + (uint32_t)(candidate - 1) + 1; s3 = p->chrpos + (uint32_t)(candidate + mid);
            if (p->cdna_direction > 0) { w2 = 1; w3 = 0; } else { w2 = 2; w3 = 3; }
          } else {
            s2 = p->chrpos + (p->genomiclength - 1) - (uint32_t)(candidate - 1); s3 = p->chrpos + (p->genomiclength - 1) - (uint32_t)(candidate + mid) + 1;
            if (p->cdna_direction > 0) { w2 = 3; w3 = 2; } else { w2 = 0; w3 = 1; }
          }
          const double prob2 = g_setup.splice_prob(w2, p->chroffset + s2, p->chroffset, g_setup.user);
          const double prob3 = g_setup.splice_prob(w3, p->chroffset + s3, p->chroffset, g_setup.user);
          if (prob2 + prob3 > bestprob) { bestcL = cL; bestcR = cR; bestmid = mid; res->left_prob = prob2; res->right_prob = prob3; bestprob = prob2 + prob3; }
        }
      }
    }
  }
  if (bestcL < 0 || bestcR < 0) { res->introntype = 0; return 0; }
  Stack st = { 0, 0, 0 };
  /* make_microexon_pairs_double with offset2M = candidate: the LAST hit looked at, not the best one (7401) */
  micro_segment(&st, p, p->offset1, lo, bestcL, p->dynprogindex);
  push(&st, -1, -1, ' ', gapchar, ' ', 0, 1);
  micro_segment(&st, p, p->offset1 + bestcL, candidate, bestmid, p->dynprogindex);
  push(&st, -1, -1, ' ', gapchar, ' ', 0, 1);
  micro_segment(&st, p, p->offset1 + bestcL + bestmid, ro - bestcR + 1, bestcR, p->dynprogindex);
  emit(out, st.v, st.n, 1);                                                 /* the list is returned as pushed: last pair first */
  res->npairs = st.n; res->null_list = 0;
  res->dynprogindex_out = bump(p->dynprogindex);
  free(st.v);
  return 0;
}

/* ---- Dynprog_single_gap, 4450-4572 ------------------------------------------ */
static int solve_single(const dpc_problem_t *p, dpc_result_t *res, Stack *out) {
  int type, L1 = p->length1, L2 = p->length2;
  quality(p->defect_rate, &type);
  if (L1 > g_maxlength1 || L2 > g_maxlength2) {           /* 4509-4519 */
    res->finalscore = -10000; res->nmatches = res->nmismatches = res->nopens = res->nindels = 0;
    res->dynprogindex_out = bump(p->dynprogindex);
    return 0;
  }
  if (L1 <= 0 || L2 <= 0) return DPC_ERR_ARG;             /* Matrix3_alloc abort, 495-498 */
  char *q = gather_query(p->seq1, L1, 0), *g = gather_genome(p, p->offset2, L2, 0);
  Mat m; Counts ct = { 0, 0, 0, 0 }; Stack st = { 0, 0, 0 };
  fill(&m, L1, L2, q, g, 1, type, -10, -3, p->extraband, p->widebandp, p->jump_late_p);
  res->finalscore = m.N[AT(&m, L1, L2)];
  traceback(&st, &ct, &m, L1, L2, q, g, p->offset1, p->offset2, 0, 0, p->cdna_direction, p->dynprogindex);
  res->nmatches = ct.nmatches; res->nmismatches = ct.nmismatches; res->nopens = ct.nopens; res->nindels = ct.nindels;
  res->dynprogindex_out = bump(p->dynprogindex);
  emit(out, st.v, st.n, 0);                               /* List_reverse of the pushed list */
  res->npairs = st.n; res->null_list = st.n == 0;
  mat_free(&m); free(q); free(g); free(st.v);
  return 0;
}

/* ---- end gaps: find_best_endpoint* 2235-2369, Dynprog_end5_gap 5094-5284,
 *      Dynprog_end3_gap 5556-5741 ------------------------------------------------ */
static void best_endpoint(int *score, int *br, int *bc, const Mat *m, int L1, int L2, int eb, int late) {
  int best = 0; *br = *bc = 0;
  for (int r = 1; r <= L1; r++) {
    int clo = r - eb < 1 ? 1 : r - eb, chigh = r + eb > L2 ? L2 : r + eb;
    for (int c = clo; c <= chigh; c++) {
      int v = m->N[AT(m, r, c)];
      if (late ? v >= best : v > best) { best = v; *br = r; *bc = c; }
    }
  }
  *score = best;
}
static void best_endpoint_queryend(int *score, int *br, int *bc, const Mat *m, int L1, int L2, int eb, int late) {
  int best = NEG, lband, rband, r = L1;
  bands(L1, L2, eb, 1, &lband, &rband);
  *br = L1; *bc = 0;
  int clo = r - lband < 1 ? 1 : r - lband, chigh = r + rband > L2 ? L2 : r + rband;
  for (int c = clo; c <= chigh; c++) {
    int v = m->N[AT(m, r, c)];
    if (late ? v >= best : v > best) { best = v; *br = r; *bc = c; }
  }
  *score = best;
}

static int solve_end(const dpc_problem_t *p, dpc_result_t *res, Stack *out, int five) {
  int L1 = p->length1, L2 = p->length2, ea = p->endalign, br, bc;
  int late = five ? !p->jump_late_p : p->jump_late_p;     /* 5190-5192 vs 5648-5650 */
  if (ea < 0 || ea > 3) return DPC_ERR_ARG;               /* abort(), 5215 */
  if (L1 <= 0) { res->nmatches = res->nmismatches = res->nopens = res->nindels = 0; res->finalscore = 0; return 0; }
  if (ea != DPC_QUERYEND_NOGAPS && L1 > g_maxlength1) L1 = g_maxlength1;
  if (L2 <= 0) { res->nmatches = res->nmismatches = res->nopens = res->nindels = 0; res->finalscore = 0; return 0; }
  if (ea != DPC_QUERYEND_NOGAPS && L2 > g_maxlength2) L2 = g_maxlength2;

  Counts ct = { 0, 0, 0, 0 }; Stack st = { 0, 0, 0 }; Mat m; memset(&m, 0, sizeof m);
  int score = DPC_UNSET;
  if (ea == DPC_QUERYEND_NOGAPS) {
    br = bc = L2 < L1 ? L2 : L1;                          /* 2358-2369 */
    char *q = gather_query(p->seq1, br, five), *g = gather_genome(p, p->offset2, bc, five);
    traceback_nogaps(&st, &ct, br, bc, q, g, p->offset1, p->offset2, five, p->dynprogindex);
    score = ct.nmatches * 3 + ct.nmismatches * -5;        /* 5243 */
    free(q); free(g);
  } else {
    char *q = gather_query(p->seq1, L1, five), *g = gather_genome(p, p->offset2, L2, five);
    fill(&m, L1, L2, q, g, 1, ENDQ, -12, -1, p->extraband, 1, late);
    if (ea == DPC_QUERYEND_INDELS) best_endpoint_queryend(&score, &br, &bc, &m, L1, L2, p->extraband, late);
    else best_endpoint(&score, &br, &bc, &m, L1, L2, p->extraband, late);
    traceback(&st, &ct, &m, br, bc, q, g, p->offset1, p->offset2, five, 0, p->cdna_direction, p->dynprogindex);
    mat_free(&m); free(q); free(g);
  }
  res->finalscore = score;
  res->nmatches = ct.nmatches; res->nmismatches = ct.nmismatches; res->nopens = ct.nopens; res->nindels = ct.nindels;
  res->dynprogindex_out = bump(p->dynprogindex);
  if ((ea == DPC_QUERYEND_GAP || ea == DPC_BEST_LOCAL) && ct.nmatches + 1 < ct.nmismatches) {
    res->finalscore = 0;                                  /* 5259-5262 */
  } else {
    int first = 0;                                        /* 5265-5268: drop leading '-' of the reversed list */
    while (first < st.n && st.v[first].comp == '-') first++;
    if (five) emit(out, st.v + first, st.n - first, 1);   /* List_reverse again, 5283 */
    else emit(out, st.v + first, st.n - first, 0);        /* returned as is, 5740 */
    res->npairs = st.n - first;
  }
  res->null_list = res->npairs == 0;
  free(st.v);
  return 0;
}

/* Dynprog_end5_splicejunction 5411-5552 / Dynprog_end3_splicejunction 5869-6012: an end of the query against the
 * splice-junction string (seq1R, use_genomicseg_p); end point on the last row; traceback_local down to contlength
 * (length2R) with the far offset (offset2R), the known gapholder, traceback_local to column 0 with the anchor
 * offset (offset2); the score is recomputed from the counts. */
static int solve_splicejunction(const dpc_problem_t *p, dpc_result_t *res, Stack *out, int five) {
  int L1 = p->length1, L2 = p->length2, br, bc, score;
  int late = five ? !p->jump_late_p : p->jump_late_p;     /* 5504 vs 5964 */
  if (L1 <= 0 || L1 > g_maxlength1 || L2 <= 0 || L2 > g_maxlength2) {    /* 5452-5465 */
    res->nmatches = res->nmismatches = res->nopens = res->nindels = 0; res->finalscore = 0; return 0;
  }
  Counts ct = { 0, 0, 0, 0 }; Stack st = { 0, 0, 0 }; Mat m; memset(&m, 0, sizeof m);
  char *q = gather_query(p->seq1, L1, five), *g = gather_query(p->seq1R, L2, five);
  fill(&m, L1, L2, q, g, 1, ENDQ, -12, -1, p->extraband, 1, late);
  best_endpoint_queryend(&score, &br, &bc, &m, L1, L2, p->extraband, late);
  traceback_from(&st, &ct, &m, &br, &bc, p->length2R, q, g, p->offset1, p->offset2R, five, 0, p->cdna_direction, p->dynprogindex);
  push(&st, 0, five ? p->offset2 - p->offset2R : p->offset2R - p->offset2, ' ', ' ', ' ', 0, 2);   /* 5518, 5977 */
  traceback_from(&st, &ct, &m, &br, &bc, 0, q, g, p->offset1, p->offset2, five, 0, p->cdna_direction, p->dynprogindex);
  mat_free(&m); free(q); free(g);
  res->finalscore = ct.nmatches * 3 + ct.nmismatches * -5 + ct.nopens * -12 + ct.nindels * -1;      /* 5537, 5996 */
  res->nmatches = ct.nmatches; res->nmismatches = ct.nmismatches; res->nopens = ct.nopens; res->nindels = ct.nindels;
  res->dynprogindex_out = bump(p->dynprogindex);
  int first = 0;                                          /* 5541-5544 */
  while (first < st.n && st.v[first].comp == '-') first++;
  emit(out, st.v + first, st.n - first, five);            /* 5551 vs 6011 */
  res->npairs = st.n - first;
  res->null_list = res->npairs == 0;
  free(st.v);
  return 0;
}

/* ---- genome gap: intron_score 3148-3192, bridge_intron_gap 3290-4122,
 *      Dynprog_genome_gap 4798-5061 ---------------------------------------------- */
static int intron_score(int *introntype, int leftdi, int rightdi, int cdna_direction, int reward, int finalp) {
  int t = leftdi & rightdi, fwd, s;
  if (t == 0) { *introntype = 0; return 0; }
  fwd = t >= 0x08;
  if ((cdna_direction > 0 && !fwd) || (cdna_direction < 0 && fwd)) { *introntype = 0; return 0; }
  switch (t) {
  case 0x20: case 0x04: s = reward; break;
  case 0x10: case 0x02: s = finalp ? 20 : 15; break;
  case 0x08: case 0x01: s = 12; break;
  default: *introntype = 0; return 0;
  }
  *introntype = t;
  return s;
}

static double site_prob(const dpc_problem_t *p, int left, int c, int lo, int ro, int known) {
  /* get_splicesite_probs 3195-3287 and the per-position arrays 3856-3903 */
  uint32_t pos; int which;
  if (known) return 1.0;
  if (left) {
    if (p->watsonp) { pos = p->chrpos + lo + c; which = p->cdna_direction > 0 ? 0 : 3; }
    else { pos = p->chrpos + (p->genomiclength - 1) - lo - c + 1; which = p->cdna_direction > 0 ? 2 : 1; }
  } else {
    if (p->watsonp) { pos = p->chrpos + ro - c + 1; which = p->cdna_direction > 0 ? 1 : 2; }
    else { pos = p->chrpos + (p->genomiclength - 1) - ro + c; which = p->cdna_direction > 0 ? 3 : 0; }
  }
  return g_setup.splice_prob(which, p->chroffset + pos, p->chroffset, g_setup.user);
}

static int site_known(const dpc_problem_t *p, int left, int c, int lo, int ro) {
  /* splice-site level lookups, 3377-3458 (the intron-level flavour 3460-3542 maps onto the same hook) */
  uint32_t pos; int which, sign;
  if (!g_setup.splice_known) return 0;
  if (left) {
    if (p->watsonp) { pos = p->chrpos + lo + c; which = p->cdna_direction > 0 ? 0 : 3; sign = p->cdna_direction > 0 ? +1 : -1; }
    else { pos = p->chrpos + (p->genomiclength - 1) - lo - c + 1; which = p->cdna_direction > 0 ? 2 : 1; sign = p->cdna_direction > 0 ? -1 : +1; }
  } else {
    if (p->watsonp) { pos = p->chrpos + ro - c + 1; which = p->cdna_direction > 0 ? 1 : 2; sign = p->cdna_direction > 0 ? +1 : -1; }
    else { pos = p->chrpos + (p->genomiclength - 1) - ro + c; which = p->cdna_direction > 0 ? 3 : 0; sign = p->cdna_direction > 0 ? -1 : +1; }
  }
  return g_setup.splice_known(which, p->chrnum, pos, sign, g_setup.user) ? 20 : 0;   /* KNOWN_SPLICESITE_REWARD */
}

static int solve_genome(const dpc_problem_t *p, dpc_result_t *res, Stack *out) {
  int type, open, extend, reward, L1 = p->length1, L2L = p->length2, L2R = p->length2R;
  int lo = p->offset2, ro = p->offset2R, eb = p->extraband;
  res->nmatches = res->nmismatches = res->nopens = res->nindels = 0;
  res->left_prob = res->right_prob = 0.0;
  if (L1 <= 1) { res->finalscore = NEG; return 0; }                     /* 4855-4858 */
  quality(p->defect_rate, &type);
  if (L1 > p->maxpeelback * 4) { open = -10; extend = -3; } else { open = -18; extend = -3; }   /* 4862-4870 */
  if (!p->splicingp) reward = 0;
  else reward = (p->finalp ? 30 : 10) + 6 * type;                       /* 277-283 */
  if (L1 > g_maxlength1 || L2L > g_maxlength2 || L2R > g_maxlength2) {  /* 4922-4954 */
    res->new_leftgenomepos = lo - 1; res->new_rightgenomepos = ro + 1; res->exonhead = p->offset1 + L1 - 1;
    res->dynprogindex_out = bump(p->dynprogindex); res->finalscore = NEG;
    return 0;
  }
  if (L2L <= 0 || L2R <= 0) return DPC_ERR_ARG;
  if (L2L < L1 - 1 || L2R < L1 - 1) return DPC_ERR_ARG;   /* the reference would index rightdi/leftdi out of bounds */
  /* 3552: with novelsplicingp == false and an intron-level IIT the scan is constrained to the given introns */
  const int constrained = !g_setup.novelsplicingp && g_setup.splice_known && g_setup.intron_level;
  if (constrained && !g_setup.splice_intron) return DPC_ERR_STATE;

  char *q = gather_query(p->seq1, L1, 0), *qr = gather_query(p->seq1 + (L1 - 1), L1, 1);
  char *gL = gather_genome(p, lo, L2L, 0), *gR = gather_genome(p, ro, L2R, 1);
  Mat mL, mR;
  fill(&mL, L1, L2L, q, gL, 1, type, open, extend, eb, 1, p->jump_late_p);
  fill(&mR, L1, L2R, qr, gR, 1, type, open, extend, eb, 1, !p->jump_late_p);

  /* dinucleotides, 3331-3373 */
  int *leftdi = calloc(L2L + 1, sizeof(int)), *rightdi = calloc(L2R + 1, sizeof(int));
  int *lknown = calloc(L2L + 1, sizeof(int)), *rknown = calloc(L2R + 1, sizeof(int));
  for (int c = 0; c < L2L - 1; c++) {
    char a = gL[c], b = gL[c + 1];
    leftdi[c] = (a == 'G' && b == 'T') ? 0x21 : (a == 'G' && b == 'C') ? 0x10 : (a == 'A' && b == 'T') ? 0x08 : (a == 'C' && b == 'T') ? 0x06 : 0;
    lknown[c] = site_known(p, 1, c, lo, ro);
  }
  for (int c = 0; c < L2R - 1; c++) {
    char r2 = gR[c + 1], r1 = gR[c];
    rightdi[c] = (r2 == 'A' && r1 == 'G') ? 0x30 : (r2 == 'A' && r1 == 'C') ? 0x0C : (r2 == 'G' && r1 == 'C') ? 0x02 : (r2 == 'A' && r1 == 'T') ? 0x01 : 0;
    rknown[c] = site_known(p, 0, c, lo, ro);
  }

  int rbandL = L2L - L1 + eb, lbandL = eb, rbandR = L2R - L1 + eb, lbandR = eb;   /* 3545-3549 */
  int bestrL = -1, bestrR = -1, bestcL = -1, bestcR = -1, have = 0;
  int finalscore, introntype = DPC_UNSET, it;
  #define NONDIAG(m, r, c) ((m).dN[AT(&(m), r, c)] == D_HORIZ || (m).dN[AT(&(m), r, c)] == D_VERT)

  if (constrained) {
    int best = -100000;                                                 /* 3552-3696 */
    for (int rL = 1, rR = L1 - 1; rL < L1; rL++, rR--) {
      int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
      int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L2R - 1 ? L2R - 1 : rR + rbandR;
      for (int side = 0; side < 2; side++) {
        int lo_c = side == 0 ? cloL : cloR, hi_c = side == 0 ? chighL : chighR;
        for (int cc = lo_c; cc <= hi_c; cc++) {
          int cL = side == 0 ? cc : rL, cR = side == 0 ? rR : cc;
          /* 3573-3583 / 3635-3650: the scanned side must be known, then the offset test, then the other side */
          if (side == 0 ? lknown[cL] <= 0 : rknown[cR] <= 0) continue;
          if (!(side == 0 ? cR < ro - lo - cL : cL < ro - lo - cR)) continue;
          if (side == 0 ? rknown[cR] <= 0 : lknown[cL] <= 0) continue;
          int sL = mL.N[AT(&mL, rL, cL)], sR = mR.N[AT(&mR, rR, cR)];
          if (side == 0 ? NONDIAG(mL, rL, cL) : NONDIAG(mR, rR, cR)) { if (side == 0) sL -= 1; else sR -= 1; }
          if (sL + sR > best) {
            int yes;
            if (p->watsonp) {
              uint32_t pos1 = p->chrpos + lo + cL, pos2 = p->chrpos + ro - cR + 1;
              yes = g_setup.splice_intron(p->chrnum, pos1, pos2 + 1U, p->cdna_direction, g_setup.user);
            } else {
              uint32_t pos1 = p->chrpos + (p->genomiclength - 1) - lo - cL + 1, pos2 = p->chrpos + (p->genomiclength - 1) - ro + cR;
              yes = g_setup.splice_intron(p->chrnum, pos2, pos1 + 1U, -p->cdna_direction, g_setup.user);
            }
            if (yes) { best = sL + sR; bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; have = 1; }
          }
        }
      }
    }
    finalscore = best;                                                  /* 3694-3695 */
    introntype = 0;
  } else if (!p->use_probabilities_p) {
    int best = -100000, bestI = -100000;                                /* 3698-3827 */
    for (int rL = 1, rR = L1 - 1; rL < L1; rL++, rR--) {
      int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
      int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L2R - 1 ? L2R - 1 : rR + rbandR;
      for (int cL = cloL; cL <= chighL; cL++) {
        int cR = rR;
        int sL = mL.N[AT(&mL, rL, cL)] + lknown[cL] - (NONDIAG(mL, rL, cL) ? 1 : 0);
        if (cR < ro - lo - cL) {
          int sR = mR.N[AT(&mR, rR, cR)] + rknown[cR];
          int sI = intron_score(&it, leftdi[cL], rightdi[cR], p->cdna_direction, reward, p->finalp);
          if (sL + sI + sR > best) { best = sL + sI + sR; bestI = sI; bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; introntype = it; have = 1; }
        }
      }
      for (int cR = cloR; cR <= chighR; cR++) {
        int cL = rL;
        int sR = mR.N[AT(&mR, rR, cR)] + rknown[cR] - (NONDIAG(mR, rR, cR) ? 1 : 0);
        if (cL < ro - lo - cR) {
          int sL = mL.N[AT(&mL, rL, cL)] + lknown[cL];
          int sI = intron_score(&it, leftdi[cL], rightdi[cR], p->cdna_direction, reward, p->finalp);
          if (sL + sI + sR > best) { best = sL + sI + sR; bestI = sI; bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; introntype = it; have = 1; }
        }
      }
    }
    finalscore = p->halfp ? best - bestI / 2 : best;                    /* 3823-3827 */
  } else {
    double bestprob = 0.0;                                              /* 3829-4081 */
    double *lp = calloc(L2L + 1, sizeof(double)), *rp = calloc(L2R + 1, sizeof(double));
    for (int c = 0; c < L2L - 1; c++) lp[c] = site_prob(p, 1, c, lo, ro, lknown[c]);
    for (int c = 0; c < L2R - 1; c++) rp[c] = site_prob(p, 0, c, lo, ro, rknown[c]);
    for (int rL = 1, rR = L1 - 1; rL < L1; rL++, rR--) {
      int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L2L - 1 ? L2L - 1 : rL + rbandL;
      int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L2R - 1 ? L2R - 1 : rR + rbandR;
      for (int cL = cloL; cL <= chighL; cL++) {
        int cR = rR;
        if (cR < ro - lo - cL && !(lp[cL] + rp[cR] <= bestprob)) {
          int sL = mL.N[AT(&mL, rL, cL)] + lknown[cL] - (NONDIAG(mL, rL, cL) ? 1 : 0);
          int sR = mR.N[AT(&mR, rR, cR)] + rknown[cR];
          int sI = intron_score(&it, leftdi[cL], rightdi[cR], p->cdna_direction, reward, p->finalp);
          if (sL + sI + sR >= p->score_threshold) { bestprob = lp[cL] + rp[cR]; bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; have = 1; }
        }
      }
      for (int cR = cloR; cR <= chighR; cR++) {
        int cL = rL;
        if (cL < ro - lo - cR && !(lp[cL] + rp[cR] <= bestprob)) {
          int sL = mL.N[AT(&mL, rL, cL)] + lknown[cL];
          int sR = mR.N[AT(&mR, rR, cR)] + rknown[cR] - (NONDIAG(mR, rR, cR) ? 1 : 0);
          int sI = intron_score(&it, leftdi[cL], rightdi[cR], p->cdna_direction, reward, p->finalp);
          if (sL + sI + sR >= p->score_threshold) { bestprob = lp[cL] + rp[cR]; bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; have = 1; }
        }
      }
    }
    free(lp); free(rp);
    if (!have) {
      /* The reference reads uninitialised bestr and bestc here (4055): undefined. Reported as "no solution". */
      finalscore = -100000;
    } else {
      int sL = mL.N[AT(&mL, bestrL, bestcL)] + lknown[bestcL] - (NONDIAG(mL, bestrL, bestcL) ? 1 : 0);
      int sR = mR.N[AT(&mR, bestrR, bestcR)] + rknown[bestcR] - (NONDIAG(mR, bestrR, bestcR) ? 1 : 0);
      int sI = intron_score(&it, leftdi[bestcL], rightdi[bestcR], p->cdna_direction, reward, p->finalp);
      finalscore = p->halfp ? sL + sI + sR - sI / 2 : sL + sI + sR;     /* 4055-4080; introntype stays untouched */
    }
  }
  res->finalscore = finalscore;
  res->introntype = introntype;

  int ok = finalscore >= 0;                                             /* 4083-4101 */
  if (ok && !g_setup.novelsplicingp && g_setup.splice_known && !g_setup.intron_level && (lknown[bestcL] == 0 || rknown[bestcR] == 0)) ok = 0;
  if (ok && p->finalp) {                                                /* 4104-4108 */
    res->left_prob = site_prob(p, 1, bestcL, lo, ro, lknown[bestcL] > 0);
    res->right_prob = site_prob(p, 0, bestcR, lo, ro, rknown[bestcR] > 0);
  }
  if (ok) {
    Counts ct = { 0, 0, 0, 0 }; Stack sR = { 0, 0, 0 }, sL = { 0, 0, 0 };
    int revoffset1 = p->offset1 + L1 - 1;
    res->new_leftgenomepos = lo + (bestcL - 1);                         /* 5000-5004 */
    res->new_rightgenomepos = ro - (bestcR - 1);
    res->exonhead = revoffset1 - (bestrR - 1);
    traceback(&sR, &ct, &mR, bestrR, bestcR, qr, gR, revoffset1, ro, 1, 0, p->cdna_direction, p->dynprogindex);
    traceback(&sL, &ct, &mL, bestrL, bestcL, q, gL, p->offset1, lo, 0, 0, p->cdna_direction, p->dynprogindex);
    res->nmatches = ct.nmatches; res->nmismatches = ct.nmismatches; res->nopens = ct.nopens; res->nindels = ct.nindels;
    res->dynprogindex_out = bump(p->dynprogindex);
    if (sR.n + sL.n > 0) {                                              /* List_length == 1 -> NULL, 5051 */
      emit(out, sR.v, sR.n, 1);
      push_gapholder(out);
      emit(out, sL.v, sL.n, 0);
      res->npairs = sR.n + sL.n + 1;
    }
    free(sR.v); free(sL.v);
  }
  res->null_list = res->npairs == 0;
  mat_free(&mL); mat_free(&mR);
  free(q); free(qr); free(gL); free(gR); free(leftdi); free(rightdi); free(lknown); free(rknown);
  return 0;
}

/* ---- cDNA gap: bridge_cdna_gap 3066-3146, Dynprog_cdna_gap 4577-4793 ----------- */
static int solve_cdna(const dpc_problem_t *p, dpc_result_t *res, Stack *out) {
  int type, L1L = p->length1, L1R = p->length1R, L2 = p->length2, eb = p->extraband;
  int open = -10, extend = -7;
  if (L2 <= 1) return 0;                                                /* 4605-4607: nothing written */
  quality(p->defect_rate, &type);
  if (L2 > g_maxlength1 || L1R > g_maxlength2 || L1L > g_maxlength2) {  /* 4648-4670 */
    res->dynprogindex_out = bump(p->dynprogindex);
    return 0;
  }
  if (L1L <= 0 || L1R <= 0) return DPC_ERR_ARG;
  int revoffset2 = p->offset2 + L2 - 1;
  char *qL = gather_query(p->seq1, L1L, 0), *qR = gather_query(p->seq1R, L1R, 1);
  char *gF = gather_genome(p, p->offset2, L2, 0), *gB = gather_genome(p, revoffset2, L2, 1);
  Mat mL, mR;
  fill(&mR, L2, L1R, gB, qR, 0, type, open, extend, eb, 1, !p->jump_late_p);
  fill(&mL, L2, L1L, gF, qL, 0, type, open, extend, eb, 1, p->jump_late_p);

  int best = -100000, bestrL = 0, bestrR = 0, bestcL = 0, bestcR = 0;
  int rbandL = L1L - L2 + eb, lbandL = eb, rbandR = L1R - L2 + eb, lbandR = eb;
  int lo = p->offset1, ro = p->offset1R;                                /* leftoffset, rightoffset = offset1L, revoffset1R */
  for (int rL = 1; rL < L2; rL++) {
    int pen = 0;
    for (int rR = L2 - rL; rR >= 0; rR--, pen += extend) {
      int cloL = rL - lbandL < 1 ? 1 : rL - lbandL, chighL = rL + rbandL > L1L - 1 ? L1L - 1 : rL + rbandL;
      int cloR = rR - lbandR < 1 ? 1 : rR - lbandR, chighR = rR + rbandR > L1R - 1 ? L1R - 1 : rR + rbandR;
      for (int cL = cloL; cL <= chighL; cL++) {
        int sL = mL.N[AT(&mL, rL, cL)];
        for (int cR = cloR; cR <= chighR && cR < ro - lo - cL; cR++) {
          int sR = mR.N[AT(&mR, rR, cR)];
          if (sL + sR + pen > best) { best = sL + sR + pen; bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; }
        }
      }
      pen = open - extend;                                              /* 3136 */
    }
  }
  res->finalscore = best;
  if (best == -100000) {
    /* bestr and bestc are read uninitialised by the reference (4714); treated as an argument error */
    mat_free(&mL); mat_free(&mR); free(qL); free(qR); free(gF); free(gB);
    return DPC_ERR_ARG;
  }

  Counts ct = { 0, 0, 0, 0 }; Stack sR = { 0, 0, 0 }, sL = { 0, 0, 0 }, mid = { 0, 0, 0 };
  traceback(&sR, &ct, &mR, bestrR, bestcR, qR, gB, p->offset1R, revoffset2, 1, 1, p->cdna_direction, p->dynprogindex);
  int queryjump = (p->offset1R - bestcR) - (p->offset1 + bestcL) + 1;   /* 4725-4726 */
  int genomejump = (revoffset2 - bestrR) - (p->offset2 + bestrL) + 1;
  if (queryjump == 9 && genomejump == 9) {                              /* INSERT_PAIRS, 4730-4751 */
    for (int k = p->offset1R - bestcR; k >= p->offset1 + bestcL; k--)
      push(&mid, k, revoffset2 - bestrR + 1, p->seq1[k - p->offset1], '~', ' ', p->dynprogindex, 0);
    for (int k = revoffset2 - bestrR; k >= p->offset2 + bestrL; k--)
      push(&mid, p->offset1 + bestcL, k, ' ', '~', gF[k - p->offset2], p->dynprogindex, 0);
  } else {
    push_gapholder(&mid);
    res->incompletep = 1;
  }
  traceback(&sL, &ct, &mL, bestrL, bestcL, qL, gF, p->offset1, p->offset2, 0, 1, p->cdna_direction, p->dynprogindex);
  res->dynprogindex_out = bump(p->dynprogindex);
  if (sR.n + mid.n + sL.n != 1) {                                       /* 4784-4787 */
    emit(out, sR.v, sR.n, 1);
    emit(out, mid.v, mid.n, 0);
    emit(out, sL.v, sL.n, 0);
    res->npairs = sR.n + mid.n + sL.n;
  }
  res->null_list = res->npairs == 0;
  free(sR.v); free(sL.v); free(mid.v);
  mat_free(&mL); mat_free(&mR); free(qL); free(qR); free(gF); free(gB);
  return 0;
}

/* ---- batch driver, same shape as dpc_solve ----------------------------------- */
int port_solve(const dpc_problem_t *problems, int n, dpc_result_t *results,
               dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off) {
  int64_t used = 0;
  if (!g_have_setup) return DPC_ERR_STATE;
  for (int i = 0; i < n; i++) {
    const dpc_problem_t *p = &problems[i];
    Stack out = { 0, 0, 0 };
    int rc;
    result_init(&results[i], p);
    switch (p->kind) {
    case DPC_SINGLE_GAP: rc = solve_single(p, &results[i], &out); break;
    case DPC_GENOME_GAP: rc = solve_genome(p, &results[i], &out); break;
    case DPC_CDNA_GAP: rc = solve_cdna(p, &results[i], &out); break;
    case DPC_END5_GAP: rc = solve_end(p, &results[i], &out, 1); break;
    case DPC_END3_GAP: rc = solve_end(p, &results[i], &out, 0); break;
    case DPC_END5_SPLICEJUNCTION: rc = solve_splicejunction(p, &results[i], &out, 1); break;
    case DPC_END3_SPLICEJUNCTION: rc = solve_splicejunction(p, &results[i], &out, 0); break;
    case DPC_MICROEXON_INT: rc = solve_microexon(p, &results[i], &out); break;
    default: rc = DPC_ERR_ARG;
    }
    if (rc != 0) { free(out.v); return rc; }
    if (pair_off) pair_off[i] = used;
    if (pairs) {
      if (used + out.n > pair_cap) { free(out.v); return DPC_ERR_NOMEM; }
      if (out.n) memcpy(pairs + used, out.v, out.n * sizeof(dpc_pair_t));
    }
    used += out.n;
    free(out.v);
  }
  if (pair_off) pair_off[n] = used;
  return 0;
}
