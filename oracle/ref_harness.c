/* oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin driver around the UNMODIFIED reference gap-fill solvers.  It is compiled
 * by oracle/Makefile together with the reference's own dynprog.c, list.c, mem.c,
 * except.c, assert.c, intlist.c, pairpool.c, intron.c, maxent.c, maxent_hr.c,
 * boyer-moore.c and splicetrie.c, read in place from /root/reference/src, into
 * oracle/_ref/libdynprog_ref.so (git-ignored).  No reference source is copied
 * into this repository; this file only calls the reference's public functions
 * (src/dynprog.h) and flattens the returned List_T of Pair_T.
 *
 * It speaks the problem/result structs of include/dynprog_cuda.h so that the
 * same inputs can be pushed through the reference, the restatement
 * (oracle/dynprog_port.c) and the CUDA library.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <time.h>
#include <pthread.h>

#include "bool.h"
#include "types.h"
#include "genomicpos.h"
#include "list.h"
#include "pairdef.h"
#include "pairpool.h"
#include "dynprog.h"
#include "maxent_hr.h"

#include "../include/dynprog_cuda.h"

/* ---- state shared with the stubs below ----------------------------------- */
static const UINT4 *ref_blocks_;
static dpc_setup_t ref_setup_;
static int ref_crosstable_[4096];
static int P_maxlookback = 600, P_extraquerygap = 10, P_maxpeelback = 11, P_end = 10, P_paired = 8;
enum { DONOR_TYPEINT = 0, ACCEPTOR_TYPEINT = 1 };

/* ---- stubs for the symbols the reference objects leave undefined ----------- */
static char block_char(Genomicpos_T pos) {
  /* same decoding as genome.c:9325-9362 (little endian) */
  const UINT4 *b = ref_blocks_ + (size_t)(pos / 32U) * 3;
  int bit = pos % 32;
  if (b[2] & (1U << bit)) return 'N';
  return "ACGT"[(bit < 16 ? b[1] >> (2 * bit) : b[0] >> (2 * bit - 32)) & 3];
}
char Genome_get_char_blocks(Genomicpos_T left) { return block_char(left); }
char Genome_get_char(Genome_T genome, Genomicpos_T left) { (void)genome; return block_char(left); }
void Genome_fill_buffer_blocks_noterm(Genomicpos_T left, Genomicpos_T length, char *gbuffer1) {
  for (Genomicpos_T i = 0; i < length; i++) gbuffer1[i] = block_char(left + i);
}
static int known_hook(int divno, unsigned int x, int type, int sign) {
  int which = type == DONOR_TYPEINT ? (sign > 0 ? 0 : 2) : (sign > 0 ? 1 : 3);
  if (!ref_setup_.splice_known) return 0;
  return ref_setup_.splice_known(which, divno, x, sign, ref_setup_.user) != 0;
}
bool IIT_exists_with_divno_typed_signed(IIT_T iit, int divno, unsigned int x, unsigned int y, int type, int sign) {
  (void)iit; (void)y; return (bool)known_hook(divno, x, type, sign);
}
/* intron-level IIT (donor_typeint < 0, dynprog.c:3460-3542): low ends are donors / antiacceptors, high ends (asked
 * with splicesitepos + 1) acceptors / antidonors; the hook gets splicesitepos in both flavours */
bool IIT_exists_with_divno_signed(IIT_T iit, int divno, unsigned int x, unsigned int y, int sign) {
  (void)iit;
  if (!ref_setup_.splice_intron) return false;
  return ref_setup_.splice_intron(divno, x, y, sign, ref_setup_.user) != 0;
}
bool IIT_low_exists_signed_p(IIT_T iit, int divno, unsigned int x, int sign) {
  (void)iit;
  if (!ref_setup_.splice_known) return false;
  return ref_setup_.splice_known(sign > 0 ? 0 : 3, divno, x, sign, ref_setup_.user) != 0;
}
bool IIT_high_exists_signed_p(IIT_T iit, int divno, unsigned int x, int sign) {
  (void)iit;
  if (!ref_setup_.splice_known) return false;
  return ref_setup_.splice_known(sign > 0 ? 1 : 2, divno, x - 1U, sign, ref_setup_.user) != 0;
}
List_T Pair_protect(List_T pairs) { return pairs; }

/* ---- lifecycle ---------------------------------------------------------------- */
int ref_init(int maxlookback, int extraquerygap, int maxpeelback,
             int extramaterial_end, int extramaterial_paired, int mode) {
  P_maxlookback = maxlookback; P_extraquerygap = extraquerygap; P_maxpeelback = maxpeelback;
  P_end = extramaterial_end; P_paired = extramaterial_paired;
  Dynprog_init(maxlookback, extraquerygap, maxpeelback, extramaterial_end, extramaterial_paired, (Mode_T)mode);
  for (int i = 0; i < 4096; i++) ref_crosstable_[i] = i;
  return 0;
}

int ref_setup(const dpc_setup_t *s) {
  ref_setup_ = *s;
  ref_blocks_ = (const UINT4 *)s->genome_blocks;
  Maxent_hr_setup((UINT4 *)s->genome_blocks);
  Dynprog_setup((bool)(s->novelsplicingp != 0),
                s->splice_known ? (IIT_T)ref_crosstable_ : (IIT_T)NULL, ref_crosstable_,
                s->intron_level ? -1 : DONOR_TYPEINT, s->intron_level ? -1 : ACCEPTOR_TYPEINT,
                NULL, NULL, NULL, 0, NULL, NULL, NULL, NULL, /*genome*/ (Genome_T)NULL);
  return 0;
}

int ref_pairdistance(int c1, int c2) { return Dynprog_pairdistance(c1, c2); }

/* The reference's MaxEnt models, exported so tests can hand them to the other
 * two implementations as the dpc_splice_prob_fn hook. */
double ref_splice_prob(int which, uint32_t splice_pos, uint32_t chroffset, void *user) {
  (void)user;
  switch (which) {
  case 0: return Maxent_hr_donor_prob(splice_pos, chroffset);
  case 1: return Maxent_hr_acceptor_prob(splice_pos, chroffset);
  case 2: return Maxent_hr_antidonor_prob(splice_pos, chroffset);
  default: return Maxent_hr_antiacceptor_prob(splice_pos, chroffset);
  }
}

/* ---- one worker = one Dynprog_T triple + Pairpool, like gmap.c:2267-2276 ------- */
typedef struct {
  Dynprog_T L, M, R;
  Pairpool_T pool;
  char *uc, *uc2, *gseg;
  int uccap;
} Worker;

static void worker_open(Worker *w) {
  w->L = Dynprog_new(P_maxlookback, P_extraquerygap, P_maxpeelback, P_end, P_paired);
  w->M = Dynprog_new(P_maxlookback, P_extraquerygap, P_maxpeelback, P_end, P_paired);
  w->R = Dynprog_new(P_maxlookback, P_extraquerygap, P_maxpeelback, P_end, P_paired);
  w->pool = Pairpool_new();
  w->uccap = 4096;
  w->uc = malloc(w->uccap); w->uc2 = malloc(w->uccap); w->gseg = malloc(w->uccap);
}
static void worker_close(Worker *w) {
  Dynprog_free(&w->L); Dynprog_free(&w->M); Dynprog_free(&w->R);
  Pairpool_free(&w->pool);
  free(w->uc); free(w->uc2); free(w->gseg);
}
static void need(Worker *w, int n) {
  if (n + 1 > w->uccap) {
    w->uccap = 2 * (n + 1);
    w->uc = realloc(w->uc, w->uccap); w->uc2 = realloc(w->uc2, w->uccap); w->gseg = realloc(w->gseg, w->uccap);
  }
}
static char uc_u2t(char c) { int u = toupper((unsigned char)c); return (char)(u == 'U' ? 'T' : u); }   /* complement.h:36 */

/* upper-case twin of a forward span / of a span addressed by its last char */
static char *uc_fwd(char *dst, const char *seq, int len) { for (int i = 0; i < len; i++) dst[i] = uc_u2t(seq[i]); return dst; }
static char *uc_rev(char *dst, const char *revseq, int len) {
  for (int i = 0; i < len; i++) dst[len - 1 - i] = uc_u2t(revseq[-i]);
  return dst + (len - 1);
}

static List_T call_one(Worker *w, const dpc_problem_t *p, dpc_result_t *r) {
  int idx = p->dynprogindex;
  int finalscore = DPC_UNSET, nmatches = DPC_UNSET, nmismatches = DPC_UNSET, nopens = DPC_UNSET, nindels = DPC_UNSET;
  int newleft = DPC_UNSET, newright = DPC_UNSET, exonhead = DPC_UNSET, introntype = DPC_UNSET;
  double lprob = -1.0, rprob = -1.0;
  bool incomplete = false;
  List_T pairs = NULL;
  int len1 = p->length1 > 0 ? p->length1 : 0, len1R = p->length1R > 0 ? p->length1R : 0;
  need(w, len1 > len1R ? len1 : len1R);
  if (p->kind == DPC_CDNA_GAP && p->length2 > 0) need(w, p->length2);

  switch (p->kind) {
  case DPC_SINGLE_GAP:
    pairs = Dynprog_single_gap(&idx, &finalscore, &nmatches, &nmismatches, &nopens, &nindels, w->M,
                               (char *)p->seq1, uc_fwd(w->uc, p->seq1, len1), NULL, NULL,
                               p->length1, p->length2, p->offset1, p->offset2,
                               p->chroffset, p->chrhigh, p->chrpos, p->genomiclength,
                               p->cdna_direction, p->watsonp, p->jump_late_p, w->pool,
                               p->extraband, p->defect_rate, /*close_indels_mode*/ +1, p->widebandp);
    break;
  case DPC_GENOME_GAP:
    pairs = Dynprog_genome_gap(&idx, &finalscore, &newleft, &newright, &lprob, &rprob,
                               &nmatches, &nmismatches, &nopens, &nindels, &exonhead, &introntype, w->L, w->R,
                               (char *)p->seq1, uc_fwd(w->uc, p->seq1, len1), NULL, NULL, NULL, NULL,
                               p->length1, p->length2, p->length2R, p->offset1, p->offset2, p->offset2R,
                               p->chrnum, p->chroffset, p->chrhigh, p->chrpos, p->genomiclength,
                               /*genomicuc_ptr*/ NULL, /*use_genomicseg_p*/ false,
                               p->cdna_direction, p->watsonp, p->jump_late_p, w->pool, p->extraband,
                               p->defect_rate, p->maxpeelback, p->halfp, p->finalp, p->use_probabilities_p,
                               p->score_threshold, p->splicingp);
    break;
  case DPC_CDNA_GAP: {
    /* sequence2 is dereferenced only by the SHORTGAP insertion (dynprog.c:4748); give it the real segment */
    char *g = NULL;
    if (p->length2 > 0) {
      for (int i = 0; i < p->length2; i++) {
        int gp = p->offset2 + i;
        char ch = '*';
        Genomicpos_T pos = p->chroffset + p->chrpos;
        if (gp >= 0 && (Genomicpos_T)gp < p->genomiclength && !(pos < p->chroffset) && pos < p->chrhigh) {
          if (p->watsonp) ch = block_char(pos + gp);
          else {
            ch = block_char(pos + (p->genomiclength - 1) - gp);
            ch = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : ch;
          }
        }
        w->gseg[i] = ch;
      }
      g = w->gseg;
    }
    pairs = Dynprog_cdna_gap(&idx, &finalscore, &incomplete, w->L, w->R,
                             (char *)p->seq1, uc_fwd(w->uc, p->seq1, len1),
                             (char *)p->seq1R, uc_rev(w->uc2, p->seq1R, len1R),
                             g, g, p->length1, p->length1R, p->length2,
                             p->offset1, p->offset1R, p->offset2,
                             p->chroffset, p->chrhigh, p->chrpos, p->genomiclength,
                             p->cdna_direction, p->watsonp, p->jump_late_p, w->pool,
                             p->extraband, p->defect_rate);
    break;
  }
  case DPC_END5_GAP:
    pairs = Dynprog_end5_gap(&idx, &finalscore, &nmatches, &nmismatches, &nopens, &nindels, w->M,
                             (char *)p->seq1, uc_rev(w->uc, p->seq1, len1), NULL, NULL,
                             p->length1, p->length2, p->offset1, p->offset2,
                             p->chroffset, p->chrhigh, p->chrpos, p->genomiclength,
                             p->cdna_direction, p->watsonp, p->jump_late_p, w->pool,
                             p->extraband, p->defect_rate, (Endalign_T)p->endalign, /*use_genomicseg_p*/ false);
    break;
  case DPC_END3_GAP:
    pairs = Dynprog_end3_gap(&idx, &finalscore, &nmatches, &nmismatches, &nopens, &nindels, w->M,
                             (char *)p->seq1, uc_fwd(w->uc, p->seq1, len1), NULL, NULL,
                             p->length1, p->length2, p->offset1, p->offset2,
                             p->chroffset, p->chrhigh, p->chrpos, p->genomiclength,
                             p->cdna_direction, p->watsonp, p->jump_late_p, w->pool,
                             p->extraband, p->defect_rate, (Endalign_T)p->endalign, /*use_genomicseg_p*/ false);
    break;
  case DPC_END5_SPLICEJUNCTION:       /* seq1R = revsequence2 (already upper case), offset2R = revoffset2_far, length2R = contlength */
    pairs = Dynprog_end5_splicejunction(&idx, &finalscore, &nmatches, &nmismatches, &nopens, &nindels, w->M,
                                        (char *)p->seq1, uc_rev(w->uc, p->seq1, len1), (char *)p->seq1R, (char *)p->seq1R,
                                        p->length1, p->length2, p->offset1, p->offset2, p->offset2R,
                                        p->chroffset, p->chrhigh, p->chrpos, p->genomiclength,
                                        p->cdna_direction, p->watsonp, p->jump_late_p, w->pool,
                                        p->extraband, p->defect_rate, /*contlength*/ p->length2R);
    break;
  case DPC_END3_SPLICEJUNCTION:
    pairs = Dynprog_end3_splicejunction(&idx, &finalscore, &nmatches, &nmismatches, &nopens, &nindels, w->M,
                                        (char *)p->seq1, uc_fwd(w->uc, p->seq1, len1), (char *)p->seq1R, (char *)p->seq1R,
                                        p->length1, p->length2, p->offset1, p->offset2, p->offset2R,
                                        p->chroffset, p->chrhigh, p->chrpos, p->genomiclength,
                                        p->cdna_direction, p->watsonp, p->jump_late_p, w->pool,
                                        p->extraband, p->defect_rate, /*contlength*/ p->length2R);
    break;
  case DPC_MICROEXON_INT: {
    /* queryseq / queryuc are addressed with absolute query offsets (offset1 + i): rebase the piece */
    char *uc = uc_fwd(w->uc, p->seq1, len1);
    lprob = rprob = 0.0;
    pairs = Dynprog_microexon_int(&lprob, &rprob, &idx, &introntype, (char *)p->seq1, uc, NULL, NULL, NULL, NULL,
                                  p->length1, p->length2, p->length2R, p->offset1, p->offset2, p->offset2R, p->cdna_direction,
                                  (char *)p->seq1 - p->offset1, uc - p->offset1, NULL, NULL,
                                  p->chroffset, p->chrhigh, p->chrpos, p->genomiclength, p->watsonp,
                                  /*use_genomicseg_p*/ false, w->pool, p->defect_rate);
    break;
  }
  default:
    break;
  }
  r->null_list = pairs == NULL;
  r->dynprogindex_out = idx;
  r->finalscore = finalscore;
  r->nmatches = nmatches; r->nmismatches = nmismatches; r->nopens = nopens; r->nindels = nindels;
  r->new_leftgenomepos = newleft; r->new_rightgenomepos = newright;
  r->exonhead = exonhead; r->introntype = introntype;
  r->incompletep = incomplete ? 1 : DPC_UNSET;
  r->left_prob = lprob; r->right_prob = rprob;
  r->npairs = List_length(pairs);
  r->reserved = 0;
  return pairs;
}

int ref_solve(const dpc_problem_t *problems, int n, dpc_result_t *results,
              dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off) {
  Worker w;
  int64_t used = 0;
  worker_open(&w);
  for (int i = 0; i < n; i++) {
    Pairpool_reset(w.pool);
    List_T l = call_one(&w, &problems[i], &results[i]);
    if (pair_off) pair_off[i] = used;
    for (; l != NULL; l = List_next(l)) {
      Pair_T pr = (Pair_T)List_head(l);
      if (pairs) {
        if (used >= pair_cap) { worker_close(&w); return DPC_ERR_NOMEM; }
        dpc_pair_t *o = &pairs[used];
        o->querypos = pr->querypos; o->genomepos = (int32_t)pr->genomepos; o->dynprogindex = pr->dynprogindex;
        o->cdna = pr->cdna; o->comp = pr->comp; o->genome = pr->genome; o->gapp = pr->gapp;
        if (pr->gapp && pr->knowngapp) {      /* the known gapholder of the splice-junction solvers (dynprog.c:5518) */
          o->gapp = 2; o->querypos = pr->queryjump; o->genomepos = pr->genomejump;
        }
      }
      used++;
    }
  }
  if (pair_off) pair_off[n] = used;
  worker_close(&w);
  return 0;
}

/* ---- CPU baseline: N worker threads over contiguous slices, like gmap -t N ---- */
typedef struct { const dpc_problem_t *p; dpc_result_t *r; int lo, hi; pthread_barrier_t *ready; } Slice;
static void *slice_run(void *arg) {
  Slice *s = (Slice *)arg;
  Worker w;
  worker_open(&w);                       /* Dynprog_new x 3 + Pairpool_new: per worker thread, once, like gmap.c:2267-2276 */
  pthread_barrier_wait(s->ready);        /* the clock starts when every worker is set up */
  for (int i = s->lo; i < s->hi; i++) {
    Pairpool_reset(w.pool);
    (void)call_one(&w, &s->p[i], &s->r[i]);
  }
  pthread_barrier_wait(s->ready);        /* ... and stops when the last one has solved its slice */
  worker_close(&w);
  return NULL;
}

/* Returns the wall seconds the N workers spend solving; thread start and the allocation of the per-thread
 * Dynprog_T scratch (3 x 18 MB, touched by calloc) are outside the timed region. */
double ref_solve_mt(const dpc_problem_t *problems, int n, dpc_result_t *results, int nthreads) {
  struct timespec t0, t1;
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  Slice *sl = malloc(sizeof(Slice) * nthreads);
  pthread_barrier_t ready;
  pthread_barrier_init(&ready, NULL, (unsigned)nthreads + 1);
  for (int t = 0; t < nthreads; t++) {
    sl[t].p = problems; sl[t].r = results; sl[t].ready = &ready;
    sl[t].lo = (int)((int64_t)n * t / nthreads); sl[t].hi = (int)((int64_t)n * (t + 1) / nthreads);
    pthread_create(&th[t], NULL, slice_run, &sl[t]);
  }
  pthread_barrier_wait(&ready);
  clock_gettime(CLOCK_MONOTONIC, &t0);
  pthread_barrier_wait(&ready);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  pthread_barrier_destroy(&ready);
  free(th); free(sl);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
