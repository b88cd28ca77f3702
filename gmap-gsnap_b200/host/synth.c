/* synth.c -- synthetic workload generator for the gap-fill path (tests + bench.py).
 *
 * Builds a random genome in the reference's 3-word block format (genome.c:9325)
 * and arrays of dpc_problem_t for the configurations SURVEY.md 8(d) names:
 *   config 2  Dynprog_single_gap   10-100 bp gaps, substitutions + indels
 *   config 3  Dynprog_genome_gap   cDNA pieces across planted GT-AG / GC-AG / AT-AC introns
 *   config 4  Dynprog_end5/3_gap   read-end extension
 *   (extra)   Dynprog_cdna_gap     cDNA insertions, parameterised as stage3.c:5602-5607
 * Deterministic: xorshift64* seeded by the caller.  Host-only C; no CUDA.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/dynprog_cuda.h"

typedef struct { uint64_t s; } Rng;
static uint64_t rnd(Rng *r) {
  r->s ^= r->s >> 12; r->s ^= r->s << 25; r->s ^= r->s >> 27;
  return r->s * 2685821657736338717ULL;
}
static uint32_t below(Rng *r, uint32_t n) { return (uint32_t)((rnd(r) >> 11) % n); }
static double unif(Rng *r) { return (double)(rnd(r) >> 11) / 9007199254740992.0; }
static int range(Rng *r, int lo, int hi) { return lo + (int)below(r, (uint32_t)(hi - lo + 1)); }

/* ---- genome blocks ---------------------------------------------------------- */
static const char NT[4] = { 'A', 'C', 'G', 'T' };
static int code_of(char c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3; }

static char blk_get(const uint32_t *blocks, uint32_t pos) {
  const uint32_t *b = blocks + (uint64_t)(pos / 32U) * 3;
  int bit = pos % 32;
  if (b[2] & (1U << bit)) return 'N';
  return NT[(bit < 16 ? b[1] >> (2 * bit) : b[0] >> (2 * bit - 32)) & 3];
}
static void blk_set(uint32_t *blocks, uint32_t pos, char c) {
  uint32_t *b = blocks + (uint64_t)(pos / 32U) * 3;
  int bit = pos % 32;
  if (c == 'N') { b[2] |= 1U << bit; return; }
  b[2] &= ~(1U << bit);
  if (bit < 16) b[1] = (b[1] & ~(3U << (2 * bit))) | ((uint32_t)code_of(c) << (2 * bit));
  else b[0] = (b[0] & ~(3U << (2 * bit - 32))) | ((uint32_t)code_of(c) << (2 * bit - 32));
}

/* number of UINT4 needed for nbases (two spare blocks so MaxEnt's ptr+4 reads stay inside) */
uint64_t synth_genome_nwords(uint64_t nbases) { return ((nbases + 31) / 32 + 2) * 3; }

void synth_genome(uint32_t *blocks, uint64_t nbases, uint64_t seed, double n_frac) {
  Rng r = { seed ? seed : 1 };
  uint64_t nblocks = (nbases + 31) / 32 + 2;
  for (uint64_t i = 0; i < nblocks; i++) {
    uint64_t v = rnd(&r);
    blocks[3 * i] = (uint32_t)(v >> 32);
    blocks[3 * i + 1] = (uint32_t)v;
    blocks[3 * i + 2] = 0;
  }
  if (n_frac > 0) {
    uint64_t nn = (uint64_t)(n_frac * (double)nbases);
    for (uint64_t i = 0; i < nn; i++) {
      uint32_t pos = (uint32_t)(rnd(&r) % nbases);
      blocks[(uint64_t)(pos / 32U) * 3 + 2] |= 1U << (pos % 32);
    }
  }
}

void synth_genome_from_ascii(uint32_t *blocks, const char *seq, uint64_t nbases) {
  memset(blocks, 0, synth_genome_nwords(nbases) * sizeof(uint32_t));
  for (uint64_t i = 0; i < nbases; i++) {
    char c = seq[i];
    if (c >= 'a' && c <= 'z') c -= 32;
    blk_set(blocks, (uint32_t)i, (c == 'A' || c == 'C' || c == 'G' || c == 'T') ? c : 'N');
  }
}
char synth_genome_char(const uint32_t *blocks, uint32_t pos) { return blk_get(blocks, pos); }

/* ---- segment view: S[i] = get_genomic_nt(i) of dynprog.c:403-441 ------------- */
typedef struct { uint32_t *blocks; uint32_t base, glen; int watson; } Seg;
static char compl_nt(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
static char seg_get(const Seg *s, int i) {
  if (i < 0 || (uint32_t)i >= s->glen) return '*';
  return s->watson ? blk_get(s->blocks, s->base + (uint32_t)i) : compl_nt(blk_get(s->blocks, s->base + s->glen - 1 - (uint32_t)i));
}
static void seg_set(const Seg *s, int i, char c) {
  if (i < 0 || (uint32_t)i >= s->glen) return;
  if (s->watson) blk_set(s->blocks, s->base + (uint32_t)i, c);
  else blk_set(s->blocks, s->base + s->glen - 1 - (uint32_t)i, compl_nt(c));
}

/* copy S[from..from+len) into dst applying substitutions / deletions / insertions; returns the new length */
static int mutate(Rng *r, const Seg *s, int from, int len, char *dst, int cap,
                  double p_sub, double p_del, double p_ins, int lower) {
  int n = 0;
  for (int i = 0; i < len && n < cap - 2; i++) {
    double u = unif(r);
    char c = seg_get(s, from + i);
    if (c == '*' || c == 'N') c = NT[below(r, 4)];
    if (u < p_del) continue;
    if (u < p_del + p_ins) dst[n++] = NT[below(r, 4)];
    if (u >= p_del + p_ins && u < p_del + p_ins + p_sub) c = NT[(code_of(c) + 1 + below(r, 3)) & 3];
    if (lower && below(r, 16) == 0) c = (char)(c + 32);
    dst[n++] = c;
  }
  return n;
}

typedef struct synth_params {
  uint64_t seed;
  uint64_t genome_nbases;
  int32_t nchr;             /* the genome is cut into nchr equal chromosomes */
  int32_t extraband;        /* extraband_single / _paired / _end for the kind generated */
  int32_t len_lo, len_hi;   /* kind-specific length range (see each generator) */
  int32_t intron_lo, intron_hi;
  double p_sub, p_del, p_ins;
  double long_frac;         /* genome gaps: fraction with length1 up to long_hi */
  int32_t long_hi;
  int32_t edge_frac_pm;     /* per mille of problems pushed against a segment edge so '*' shows up */
  int32_t finalp_mode;      /* genome gaps: 0 never, 1 always, 2 mixed */
  int32_t prob_mode_pm;     /* genome gaps: per mille with use_probabilities_p */
  int32_t lower_case;       /* sprinkle lower-case query letters */
  int32_t iupac_pm;         /* per mille of query letters replaced by an IUPAC code / N */
  int32_t reserved;
} synth_params_t;

static void common(dpc_problem_t *p, Rng *r, const synth_params_t *sp, Seg *seg, uint32_t *blocks) {
  uint32_t chrlen = (uint32_t)(sp->genome_nbases / (uint64_t)(sp->nchr > 0 ? sp->nchr : 1));
  uint32_t chr = below(r, (uint32_t)(sp->nchr > 0 ? sp->nchr : 1));
  uint32_t glen = 30000 + below(r, 30000);
  if (glen > chrlen) glen = chrlen;
  memset(p, 0, sizeof *p);
  p->chrnum = (int32_t)chr + 1;
  p->chroffset = chr * chrlen;
  p->chrhigh = p->chroffset + chrlen;
  p->chrpos = below(r, chrlen - glen + 1);
  p->genomiclength = glen;
  p->watsonp = (uint8_t)(rnd(r) & 1);
  p->jump_late_p = (uint8_t)!p->watsonp;          /* gmap.c:787 */
  p->cdna_direction = (rnd(r) & 1) ? +1 : -1;
  p->extraband = sp->extraband;
  p->maxpeelback = 11;
  p->widebandp = 1;
  p->splicingp = 1;
  p->defect_rate = (below(r, 8) == 0) ? 0.01 : (below(r, 8) == 0 ? 0.05 : 0.0);
  p->dynprogindex = (rnd(r) & 1) ? 1 + (int)below(r, 50) : -1 - (int)below(r, 50);
  seg->blocks = blocks; seg->base = p->chroffset + p->chrpos; seg->glen = glen; seg->watson = p->watsonp;
}

static void iupac(Rng *r, const synth_params_t *sp, char *q, int n) {
  static const char codes[] = "NRYWSMKHBVDXUnry";
  if (sp->iupac_pm <= 0) return;
  for (int i = 0; i < n; i++) if ((int)below(r, 1000) < sp->iupac_pm) q[i] = codes[below(r, sizeof codes - 1)];
}

/* config 2: len_lo..len_hi is the genomic gap length (10..100) */
int64_t synth_single_gaps(const synth_params_t *sp, uint32_t *blocks, int n, dpc_problem_t *out, char *qbuf, int64_t qcap) {
  Rng r = { sp->seed ^ 0x9E3779B97F4A7C15ULL ^ 2 };
  int64_t used = 0;
  for (int i = 0; i < n; i++) {
    Seg seg; dpc_problem_t *p = &out[i];
    common(p, &r, sp, &seg, blocks);
    int L2 = range(&r, sp->len_lo, sp->len_hi);
    int off2 = 200 + (int)below(&r, seg.glen - 400 - (uint32_t)L2);
    if ((int)below(&r, 1000) < sp->edge_frac_pm) off2 = (rnd(&r) & 1) ? -(int)below(&r, 5) : (int)seg.glen - L2 + (int)below(&r, 5);
    if (used + 2 * L2 + 16 > qcap) return -1;
    int L1 = mutate(&r, &seg, off2, L2, qbuf + used, 2 * L2 + 16, sp->p_sub, sp->p_del, sp->p_ins, sp->lower_case);
    if (L1 < 1) { qbuf[used] = 'A'; L1 = 1; }
    iupac(&r, sp, qbuf + used, L1);
    p->kind = DPC_SINGLE_GAP;
    p->seq1 = qbuf + used; p->length1 = L1; p->length2 = L2;
    p->offset1 = 50 + (int)below(&r, 2000); p->offset2 = off2;
    used += L1;
  }
  return used;
}

/* config 4: len_lo..len_hi is the unaligned tail length (1..40); 11 peeled bases are added */
int64_t synth_end_gaps(const synth_params_t *sp, uint32_t *blocks, int n, dpc_problem_t *out, char *qbuf, int64_t qcap) {
  Rng r = { sp->seed ^ 0x9E3779B97F4A7C15ULL ^ 4 };
  int64_t used = 0;
  for (int i = 0; i < n; i++) {
    Seg seg; dpc_problem_t *p = &out[i];
    common(p, &r, sp, &seg, blocks);
    int five = (int)(rnd(&r) & 1);
    int want = range(&r, sp->len_lo, sp->len_hi) + 11;
    int off = 200 + (int)below(&r, seg.glen - 400 - (uint32_t)want - 10);
    if ((int)below(&r, 1000) < sp->edge_frac_pm) off = five ? want - 1 - (int)below(&r, 6) : (int)seg.glen - want - 10 + (int)below(&r, 16);
    if (used + 2 * want + 16 > qcap) return -1;
    char *q = qbuf + used;
    int L1;
    uint32_t mix = below(&r, 10);
    p->endalign = mix < 5 ? DPC_QUERYEND_GAP : mix < 9 ? DPC_QUERYEND_NOGAPS : DPC_BEST_LOCAL;
    if (below(&r, 50) == 0) p->endalign = DPC_QUERYEND_INDELS;
    if (!five) {
      /* 3' end: query continues S[off..]; the far part of the tail degrades (mimics a clipped adapter) */
      L1 = mutate(&r, &seg, off, want, q, 2 * want + 16, sp->p_sub, sp->p_del, sp->p_ins, sp->lower_case);
      if (L1 < 1) { q[0] = 'A'; L1 = 1; }
      if (below(&r, 4) == 0) for (int k = L1 - (int)below(&r, (uint32_t)L1 / 2 + 1); k < L1; k++) q[k] = NT[below(&r, 4)];
      iupac(&r, sp, q, L1);
      p->kind = DPC_END3_GAP; p->seq1 = q; p->length1 = L1; p->length2 = L1 + 10;
      p->offset1 = 250 - L1; p->offset2 = off;
    } else {
      /* 5' end: the piece ends at S[off] and runs leftwards */
      L1 = mutate(&r, &seg, off - want + 1, want, q, 2 * want + 16, sp->p_sub, sp->p_del, sp->p_ins, sp->lower_case);
      if (L1 < 1) { q[0] = 'A'; L1 = 1; }
      if (below(&r, 4) == 0) for (int k = (int)below(&r, (uint32_t)L1 / 2 + 1); k >= 0; k--) q[k] = NT[below(&r, 4)];
      iupac(&r, sp, q, L1);
      p->kind = DPC_END5_GAP; p->seq1 = q + (L1 - 1); p->length1 = L1; p->length2 = L1 + 10;
      p->offset1 = L1 - 1; p->offset2 = off;
    }
    used += L1;
  }
  return used;
}

/* config 3: len_lo..len_hi is length1 (22..60), long_frac of them up to long_hi (600) */
int64_t synth_genome_gaps(const synth_params_t *sp, uint32_t *blocks, int n, dpc_problem_t *out, char *qbuf, int64_t qcap) {
  Rng r = { sp->seed ^ 0x9E3779B97F4A7C15ULL ^ 3 };
  int64_t used = 0;
  /* pass 1: choose geometry and plant the dinucleotides; pass 2 derives the queries from the final genome */
  int *geo = malloc(sizeof(int) * 4 * (size_t)n);
  for (int i = 0; i < n; i++) {
    Seg seg; dpc_problem_t *p = &out[i];
    common(p, &r, sp, &seg, blocks);
    int want = (unif(&r) < sp->long_frac) ? range(&r, sp->len_hi, sp->long_hi) : range(&r, sp->len_lo, sp->len_hi);
    int a = 1 + (int)below(&r, (uint32_t)want - 1), b = want - a;
    double lg = unif(&r);
    int intron = (int)((double)sp->intron_lo * __builtin_pow((double)sp->intron_hi / (double)sp->intron_lo, lg));
    int span = a + intron + b + 40;
    if ((uint32_t)span + 400 > seg.glen) { intron = (int)seg.glen - 440 - a - b; if (intron < 20) intron = 20; span = a + intron + b + 40; }
    int x0 = 200 + (int)below(&r, seg.glen - 400 - (uint32_t)span + 1);   /* first left-exon base */
    int is = x0 + a, ie = is + intron - 1;                                  /* intron = S[is..ie] */
    uint32_t kind = below(&r, 100);
    const char *don = 0, *acc = 0;
    if (kind < 60) { don = "GT"; acc = "AG"; } else if (kind < 70) { don = "GC"; acc = "AG"; } else if (kind < 75) { don = "AT"; acc = "AC"; }
    if (don) {
      if (p->cdna_direction > 0) { seg_set(&seg, is, don[0]); seg_set(&seg, is + 1, don[1]); seg_set(&seg, ie - 1, acc[0]); seg_set(&seg, ie, acc[1]); }
      else {   /* reverse complement of donor..acceptor as seen on S: revcomp(acc) ... revcomp(don) */
        seg_set(&seg, is, compl_nt(acc[1])); seg_set(&seg, is + 1, compl_nt(acc[0]));
        seg_set(&seg, ie - 1, compl_nt(don[1])); seg_set(&seg, ie, compl_nt(don[0]));
      }
    }
    geo[4 * i] = x0; geo[4 * i + 1] = a; geo[4 * i + 2] = intron; geo[4 * i + 3] = b;
  }
  for (int i = 0; i < n; i++) {
    dpc_problem_t *p = &out[i];
    Seg seg = { blocks, p->chroffset + p->chrpos, p->genomiclength, p->watsonp };
    int x0 = geo[4 * i], a = geo[4 * i + 1], intron = geo[4 * i + 2], b = geo[4 * i + 3];
    if (used + 2 * (a + b) + 32 > qcap) { free(geo); return -1; }
    char *q = qbuf + used;
    int la = mutate(&r, &seg, x0, a, q, 2 * a + 16, sp->p_sub, sp->p_del, sp->p_ins, sp->lower_case);
    int lb = mutate(&r, &seg, x0 + a + intron, b, q + la, 2 * b + 16, sp->p_sub, sp->p_del, sp->p_ins, sp->lower_case);
    int L1 = la + lb;
    if (L1 < 2) { q[0] = 'A'; q[1] = 'C'; L1 = 2; }
    iupac(&r, sp, q, L1);
    p->kind = DPC_GENOME_GAP;
    p->seq1 = q; p->length1 = L1; p->length2 = L1 + 8; p->length2R = L1 + 8;   /* stage3.c:5793 */
    p->offset1 = 50 + (int)below(&r, 2000);
    p->offset2 = x0; p->offset2R = x0 + a + intron + b - 1;
    p->extraband = sp->extraband;
    p->halfp = (uint8_t)(below(&r, 4) == 0);
    p->finalp = (uint8_t)(sp->finalp_mode == 1 ? 1 : sp->finalp_mode == 2 ? (rnd(&r) & 1) : 0);
    p->use_probabilities_p = 0;
    p->score_threshold = 0;
    if ((int)below(&r, 1000) < sp->prob_mode_pm) p->use_probabilities_p = 2;   /* 2 = "fill in threshold from a first pass" (tests) */
    if (below(&r, 64) == 0) p->cdna_direction = 0;
    used += L1;
  }
  free(geo);
  return used;
}

/* cDNA gaps: len_lo..len_hi is the genomic gap (genomejump); the query carries an insertion of intron_lo..intron_hi bases */
int64_t synth_cdna_gaps(const synth_params_t *sp, uint32_t *blocks, int n, dpc_problem_t *out, char *qbuf, int64_t qcap) {
  Rng r = { sp->seed ^ 0x9E3779B97F4A7C15ULL ^ 5 };
  int64_t used = 0;
  for (int i = 0; i < n; i++) {
    Seg seg; dpc_problem_t *p = &out[i];
    common(p, &r, sp, &seg, blocks);
    int L2 = range(&r, sp->len_lo, sp->len_hi);
    int ins = range(&r, sp->intron_lo, sp->intron_hi);
    int off2 = 200 + (int)below(&r, seg.glen - 400 - (uint32_t)L2);
    int a = 1 + (int)below(&r, (uint32_t)L2 - 1);
    if (used + 2 * L2 + ins + 64 > qcap) return -1;
    char *q = qbuf + used;
    int la = mutate(&r, &seg, off2, a, q, 2 * a + 16, sp->p_sub, sp->p_del, sp->p_ins, sp->lower_case);
    for (int k = 0; k < ins; k++) q[la + k] = NT[below(&r, 4)];
    int lb = mutate(&r, &seg, off2 + a, L2 - a, q + la + ins, 2 * (L2 - a) + 16, sp->p_sub, sp->p_del, sp->p_ins, sp->lower_case);
    int qlen = la + ins + lb;
    int len1 = L2 + 8;                                 /* stage3.c:5602 */
    if (len1 > qlen) len1 = qlen;
    p->kind = DPC_CDNA_GAP;
    p->seq1 = q; p->seq1R = q + (qlen - 1);
    p->length1 = len1; p->length1R = len1; p->length2 = L2;
    p->offset1 = 100; p->offset1R = 100 + qlen - 1; p->offset2 = off2;
    used += qlen;
  }
  return used;
}
