/* dynprog_dropin.c -- the reference's five gap-fill entry points, verbatim signatures, on top of libdynprog_cuda.
 *
 * Compiled INSIDE a GMAP/GSNAP source tree (it includes the tree's own headers): see INTEGRATION.md and
 * oracle/build_gmap.sh.  It replaces these symbols of src/dynprog.c (file:line of the reference):
 *
 *   Dynprog_init         dynprog.c:1338      also initialises the reference's own tables (other solvers still use them)
 *   Dynprog_setup        dynprog.c:349       also registers the genome blocks and the splice hooks with the library
 *   Dynprog_term         dynprog.c:1348
 *   Dynprog_single_gap   dynprog.c:4450
 *   Dynprog_cdna_gap     dynprog.c:4577
 *   Dynprog_genome_gap   dynprog.c:4798
 *   Dynprog_end5_gap     dynprog.c:5094
 *   Dynprog_end3_gap     dynprog.c:5556
 *
 * The originals stay linked under the names <name>_cpu (objcopy --redefine-sym, or eight #defines on top of
 * dynprog.c) because Dynprog_init/_setup must still run for the solvers this library does not replace
 * (Dynprog_end5_known, Dynprog_microexon_*, ...).  The five solvers below never call their _cpu twins.
 *
 * Each call here is "add 1 + flush + wait + pairs": correct, and as slow as one kernel launch per gap.  The
 * batched use (a modified stage3.c that collects the gaps of many alignments) goes
 * through the same dpc_add / dpc_flush / dpc_result / dpc_pairs calls with more than one problem per flush.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bool.h"
#include "types.h"
#include "genomicpos.h"
#include "chrnum.h"
#include "mode.h"
#include "list.h"
#include "pairdef.h"
#include "pairpool.h"
#include "genome.h"
#include "iit-read.h"
#include "splicetrie_build.h"
#include "dynprog.h"
#include "maxent_hr.h"

#include "dynprog_cuda.h"

/* the reference's own functions, renamed in dynprog.o */
extern void Dynprog_init_cpu (int maxlookback, int extraquerygap, int maxpeelback,
			      int extramaterial_end, int extramaterial_paired, Mode_T mode);
extern void Dynprog_term_cpu (void);
extern void Dynprog_setup_cpu (bool novelsplicingp_in,
			       IIT_T splicesites_iit_in, int *splicesites_divint_crosstable_in,
			       int donor_typeint_in, int acceptor_typeint_in,
			       Genomicpos_T *splicesites_in, Splicetype_T *splicetypes_in,
			       Genomicpos_T *splicedists_in, int nsplicesites_in,
			       unsigned int *trieoffsets_obs_in, unsigned int *triecontents_obs_in,
			       unsigned int *trieoffsets_max_in, unsigned int *triecontents_max_in,
			       Genome_T genome_in);

static IIT_T dropin_iit;
static int *dropin_crosstable;
static int dropin_donor_typeint = -1, dropin_acceptor_typeint = -1;
static const UINT4 *dropin_blocks;
static uint64_t dropin_nwords;
static __thread dpc_ctx_t *dropin_ctx;         /* one context per worker thread, like the Dynprog_T triple of gmap.c:2270 */
static __thread dpc_pair_t *dropin_pairs;
static __thread int dropin_pairs_cap;

static void
dropin_fatal (const char *what, int code) {
  fprintf(stderr,"libdynprog_cuda drop-in: %s: %s\n",what,dpc_strerror(code));
  exit(9);			/* the reference's convention for unrecoverable errors, gmap.c:2287-2308 */
}

/* User-segment runs (gmap -g) keep their genome blocks in a static of gmap.c; the one added line there calls this. */
void
Dynprog_cuda_register_blocks (UINT4 *blocks, unsigned int nwords) {
  dropin_blocks = blocks;
  dropin_nwords = nwords;
}

static double
dropin_splice_prob (int which, uint32_t splice_pos, uint32_t chroffset, void *user) {
  (void) user;
  switch (which) {		/* maxent_hr.c:27217-27340 */
  case 0: return Maxent_hr_donor_prob(splice_pos,chroffset);
  case 1: return Maxent_hr_acceptor_prob(splice_pos,chroffset);
  case 2: return Maxent_hr_antidonor_prob(splice_pos,chroffset);
  default: return Maxent_hr_antiacceptor_prob(splice_pos,chroffset);
  }
}

static int
dropin_splice_known (int which, int chrnum, uint32_t splicesitepos, int sign, void *user) {
  int type = (which == 0 || which == 2) ? dropin_donor_typeint : dropin_acceptor_typeint;
  (void) user;
  /* dynprog.c:3377-3458 */
  return IIT_exists_with_divno_typed_signed(dropin_iit,dropin_crosstable[chrnum],
					    splicesitepos,splicesitepos+1U,type,sign) == true;
}

void
Dynprog_init (int maxlookback, int extraquerygap, int maxpeelback,
	      int extramaterial_end, int extramaterial_paired, Mode_T mode) {
  int rc;
  Dynprog_init_cpu(maxlookback,extraquerygap,maxpeelback,extramaterial_end,extramaterial_paired,mode);
  if ((rc = dpc_init(maxlookback,extraquerygap,maxpeelback,extramaterial_end,extramaterial_paired,(int) mode)) != DPC_OK) {
    dropin_fatal("dpc_init",rc);
  }
}

void
Dynprog_term (void) {
  Dynprog_term_cpu();
  dpc_term();
}

void
Dynprog_setup (bool novelsplicingp_in,
	       IIT_T splicesites_iit_in, int *splicesites_divint_crosstable_in,
	       int donor_typeint_in, int acceptor_typeint_in,
	       Genomicpos_T *splicesites_in, Splicetype_T *splicetypes_in,
	       Genomicpos_T *splicedists_in, int nsplicesites_in,
	       unsigned int *trieoffsets_obs_in, unsigned int *triecontents_obs_in,
	       unsigned int *trieoffsets_max_in, unsigned int *triecontents_max_in,
	       Genome_T genome_in) {
  dpc_setup_t s;
  int rc;

  Dynprog_setup_cpu(novelsplicingp_in,splicesites_iit_in,splicesites_divint_crosstable_in,
		    donor_typeint_in,acceptor_typeint_in,splicesites_in,splicetypes_in,splicedists_in,nsplicesites_in,
		    trieoffsets_obs_in,triecontents_obs_in,trieoffsets_max_in,triecontents_max_in,genome_in);
  dropin_iit = splicesites_iit_in;
  dropin_crosstable = splicesites_divint_crosstable_in;
  dropin_donor_typeint = donor_typeint_in;
  dropin_acceptor_typeint = acceptor_typeint_in;
  if (splicesites_iit_in != NULL && (donor_typeint_in < 0 || acceptor_typeint_in < 0)) {
    /* intron-level known splicing (dynprog.c:3460-3542, 3552-3696) needs IIT pair lookups the library does not model */
    dropin_fatal("Dynprog_setup: intron-level splicing IIT",DPC_ERR_UNSUPPORTED);
  }
  if (genome_in != NULL) {
    dropin_blocks = Genome_blocks(genome_in);
    dropin_nwords = (uint64_t) (Genome_totallength(genome_in)/32U + 1)*3;
  }
  if (dropin_blocks == NULL) {
    dropin_fatal("Dynprog_setup: no genome blocks registered (call Dynprog_cuda_register_blocks for a user segment)",DPC_ERR_STATE);
  }
  memset(&s,0,sizeof(s));
  s.genome_blocks = (const uint32_t *) dropin_blocks;
  s.genome_nwords = dropin_nwords;
  s.novelsplicingp = novelsplicingp_in == true;
  s.splice_prob = dropin_splice_prob;
  s.splice_known = splicesites_iit_in != NULL ? dropin_splice_known : NULL;
  if ((rc = dpc_setup(&s)) != DPC_OK) {
    dropin_fatal("dpc_setup",rc);
  }
}

/* ---- one problem through the library ------------------------------------------------------------ */
static int
dropin_device (void) {
  const char *e = getenv("DPC_DEVICE");
  return e != NULL ? atoi(e) : 0;
}

static List_T
dropin_solve (dpc_result_t *r, const dpc_problem_t *p, Pairpool_T pairpool) {
  List_T pairs = NULL;
  int rc, ticket, n, i;

  if (dropin_ctx == NULL && (dropin_ctx = dpc_ctx_new(dropin_device())) == NULL) {
    dropin_fatal("dpc_ctx_new",DPC_ERR_CUDA);
  }
  if ((rc = dpc_reset(dropin_ctx)) < 0) dropin_fatal("dpc_reset",rc);
  if ((ticket = dpc_add(dropin_ctx,p)) < 0) dropin_fatal("dpc_add",ticket);
  if ((rc = dpc_flush(dropin_ctx)) < 0) dropin_fatal("dpc_flush",rc);
  if ((rc = dpc_wait(dropin_ctx)) < 0) dropin_fatal("dpc_wait",rc);
  if ((rc = dpc_result(dropin_ctx,ticket,r)) < 0) dropin_fatal("dpc_result",rc);
  if (r->npairs > dropin_pairs_cap) {
    dropin_pairs_cap = 2*r->npairs + 256;
    dropin_pairs = (dpc_pair_t *) realloc(dropin_pairs,dropin_pairs_cap*sizeof(dpc_pair_t));
  }
  if ((n = dpc_pairs(dropin_ctx,ticket,dropin_pairs,dropin_pairs_cap)) < 0) dropin_fatal("dpc_pairs",n);

  /* dpc_pairs lists the records head first; Pairpool_push prepends (pairpool.c:169-215) */
  for (i = n - 1; i >= 0; i--) {
    const dpc_pair_t *q = &dropin_pairs[i];
    if (q->gapp) {
      pairs = Pairpool_push_gapholder(pairs,pairpool,/*queryjump*/UNKNOWNJUMP,/*genomejump*/UNKNOWNJUMP,/*knownp*/false);
    } else {
      pairs = Pairpool_push(pairs,pairpool,q->querypos,q->genomepos,q->cdna,q->comp,q->genome,q->dynprogindex);
    }
  }
  return pairs;
}

static void
dropin_common (dpc_problem_t *p, int kind, int dynprogindex,
	       Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength,
	       int cdna_direction, bool watsonp, bool jump_late_p, int extraband, double defect_rate) {
  memset(p,0,sizeof(*p));
  p->kind = kind;
  p->dynprogindex = dynprogindex;
  p->chroffset = chroffset; p->chrhigh = chrhigh; p->chrpos = chrpos; p->genomiclength = genomiclength;
  p->cdna_direction = cdna_direction;
  p->watsonp = watsonp == true; p->jump_late_p = jump_late_p == true;
  p->extraband = extraband;
  p->defect_rate = defect_rate;
  p->widebandp = 1;
  p->splicingp = 1;
  p->maxpeelback = 11;
}

#define OUT(dst,src) do { if ((src) != DPC_UNSET) *(dst) = (src); } while (0)

List_T
Dynprog_single_gap (int *dynprogindex, int *finalscore,
		    int *nmatches, int *nmismatches, int *nopens, int *nindels,
		    Dynprog_T dynprog, char *sequence1, char *sequenceuc1, char *sequence2, char *sequenceuc2,
		    int length1, int length2, int offset1, int offset2,
		    Genomicpos_T chroffset, Genomicpos_T chrhigh,
		    Genomicpos_T chrpos, Genomicpos_T genomiclength,
		    int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
		    int extraband_single, double defect_rate, int close_indels_mode, bool widebandp) {
  dpc_problem_t p;
  dpc_result_t r;
  List_T pairs;
  (void) dynprog; (void) sequenceuc1; (void) sequence2; (void) sequenceuc2; (void) close_indels_mode;

  dropin_common(&p,DPC_SINGLE_GAP,*dynprogindex,chroffset,chrhigh,chrpos,genomiclength,
		cdna_direction,watsonp,jump_late_p,extraband_single,defect_rate);
  p.seq1 = sequence1; p.length1 = length1; p.length2 = length2; p.offset1 = offset1; p.offset2 = offset2;
  p.widebandp = widebandp == true;
  pairs = dropin_solve(&r,&p,pairpool);
  OUT(finalscore,r.finalscore);
  OUT(nmatches,r.nmatches); OUT(nmismatches,r.nmismatches); OUT(nopens,r.nopens); OUT(nindels,r.nindels);
  *dynprogindex = r.dynprogindex_out;
  return pairs;
}

List_T
Dynprog_cdna_gap (int *dynprogindex, int *finalscore, bool *incompletep,
		  Dynprog_T dynprogL, Dynprog_T dynprogR, char *sequence1L, char *sequenceuc1L,
		  char *revsequence1R, char *revsequenceuc1R,
		  char *sequence2, char *sequenceuc2,
		  int length1L, int length1R, int length2,
		  int offset1L, int revoffset1R, int offset2,
		  Genomicpos_T chroffset, Genomicpos_T chrhigh,
		  Genomicpos_T chrpos, Genomicpos_T genomiclength,
		  int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
		  int extraband_paired, double defect_rate) {
  dpc_problem_t p;
  dpc_result_t r;
  List_T pairs;
  (void) dynprogL; (void) dynprogR; (void) sequenceuc1L; (void) revsequenceuc1R; (void) sequence2; (void) sequenceuc2;

  dropin_common(&p,DPC_CDNA_GAP,*dynprogindex,chroffset,chrhigh,chrpos,genomiclength,
		cdna_direction,watsonp,jump_late_p,extraband_paired,defect_rate);
  p.seq1 = sequence1L; p.seq1R = revsequence1R;
  p.length1 = length1L; p.length1R = length1R; p.length2 = length2;
  p.offset1 = offset1L; p.offset1R = revoffset1R; p.offset2 = offset2;
  pairs = dropin_solve(&r,&p,pairpool);
  OUT(finalscore,r.finalscore);
  if (r.incompletep != DPC_UNSET) *incompletep = true;
  *dynprogindex = r.dynprogindex_out;
  return pairs;
}

List_T
Dynprog_genome_gap (int *dynprogindex, int *finalscore, int *new_leftgenomepos, int *new_rightgenomepos,
		    double *left_prob, double *right_prob,
		    int *nmatches, int *nmismatches, int *nopens, int *nindels,
		    int *exonhead, int *introntype, Dynprog_T dynprogL, Dynprog_T dynprogR,
		    char *sequence1, char *sequenceuc1,
		    char *sequence2L, char *sequenceuc2L,
		    char *revsequence2R, char *revsequenceuc2R,
		    int length1, int length2L, int length2R,
		    int offset1, int offset2L, int revoffset2R,
		    Chrnum_T chrnum, Genomicpos_T chroffset, Genomicpos_T chrhigh,
		    Genomicpos_T chrpos, Genomicpos_T genomiclength,
		    char *genomicuc_ptr, bool use_genomicseg_p,
		    int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_paired,
		    double defect_rate, int maxpeelback, bool halfp, bool finalp, bool use_probabilities_p,
		    int score_threshold, bool splicingp) {
  dpc_problem_t p;
  dpc_result_t r;
  List_T pairs;
  (void) dynprogL; (void) dynprogR; (void) sequenceuc1; (void) sequence2L; (void) sequenceuc2L;
  (void) revsequence2R; (void) revsequenceuc2R; (void) genomicuc_ptr; (void) use_genomicseg_p;

  dropin_common(&p,DPC_GENOME_GAP,*dynprogindex,chroffset,chrhigh,chrpos,genomiclength,
		cdna_direction,watsonp,jump_late_p,extraband_paired,defect_rate);
  p.seq1 = sequence1; p.length1 = length1; p.length2 = length2L; p.length2R = length2R;
  p.offset1 = offset1; p.offset2 = offset2L; p.offset2R = revoffset2R;
  p.chrnum = chrnum; p.maxpeelback = maxpeelback;
  p.halfp = halfp == true; p.finalp = finalp == true; p.use_probabilities_p = use_probabilities_p == true;
  p.score_threshold = score_threshold; p.splicingp = splicingp == true;
  pairs = dropin_solve(&r,&p,pairpool);
  OUT(finalscore,r.finalscore);
  OUT(new_leftgenomepos,r.new_leftgenomepos); OUT(new_rightgenomepos,r.new_rightgenomepos);
  OUT(nmatches,r.nmatches); OUT(nmismatches,r.nmismatches); OUT(nopens,r.nopens); OUT(nindels,r.nindels);
  OUT(exonhead,r.exonhead); OUT(introntype,r.introntype);
  if (r.left_prob >= 0.0) *left_prob = r.left_prob;
  if (r.right_prob >= 0.0) *right_prob = r.right_prob;
  *dynprogindex = r.dynprogindex_out;
  return pairs;
}

static List_T
dropin_end_gap (int kind, int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches,
		int *nopens, int *nindels, char *sequence1,
		int length1, int length2, int offset1, int offset2,
		Genomicpos_T chroffset, Genomicpos_T chrhigh,
		Genomicpos_T chrpos, Genomicpos_T genomiclength,
		int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
		int extraband_end, double defect_rate, Endalign_T endalign) {
  dpc_problem_t p;
  dpc_result_t r;
  List_T pairs;

  dropin_common(&p,kind,*dynprogindex,chroffset,chrhigh,chrpos,genomiclength,
		cdna_direction,watsonp,jump_late_p,extraband_end,defect_rate);
  p.seq1 = sequence1; p.length1 = length1; p.length2 = length2; p.offset1 = offset1; p.offset2 = offset2;
  p.endalign = (int) endalign;
  pairs = dropin_solve(&r,&p,pairpool);
  OUT(finalscore,r.finalscore);
  OUT(nmatches,r.nmatches); OUT(nmismatches,r.nmismatches); OUT(nopens,r.nopens); OUT(nindels,r.nindels);
  *dynprogindex = r.dynprogindex_out;
  return pairs;
}

List_T
Dynprog_end5_gap (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches,
		  int *nopens, int *nindels, Dynprog_T dynprog,
		  char *revsequence1, char *revsequenceuc1,
		  char *revsequence2, char *revsequenceuc2,
		  int length1, int length2, int revoffset1, int revoffset2,
		  Genomicpos_T chroffset, Genomicpos_T chrhigh,
		  Genomicpos_T chrpos, Genomicpos_T genomiclength,
		  int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
		  int extraband_end, double defect_rate, Endalign_T endalign,
		  bool use_genomicseg_p) {
  (void) dynprog; (void) revsequenceuc1; (void) revsequence2; (void) revsequenceuc2; (void) use_genomicseg_p;
  return dropin_end_gap(DPC_END5_GAP,dynprogindex,finalscore,nmatches,nmismatches,nopens,nindels,revsequence1,
			length1,length2,revoffset1,revoffset2,chroffset,chrhigh,chrpos,genomiclength,
			cdna_direction,watsonp,jump_late_p,pairpool,extraband_end,defect_rate,endalign);
}

List_T
Dynprog_end3_gap (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches,
		  int *nopens, int *nindels, Dynprog_T dynprog,
		  char *sequence1, char *sequenceuc1,
		  char *sequence2, char *sequenceuc2,
		  int length1, int length2, int offset1, int offset2,
		  Genomicpos_T chroffset, Genomicpos_T chrhigh,
		  Genomicpos_T chrpos, Genomicpos_T genomiclength,
		  int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
		  int extraband_end, double defect_rate, Endalign_T endalign,
		  bool use_genomicseg_p) {
  (void) dynprog; (void) sequenceuc1; (void) sequence2; (void) sequenceuc2; (void) use_genomicseg_p;
  return dropin_end_gap(DPC_END3_GAP,dynprogindex,finalscore,nmatches,nmismatches,nopens,nindels,sequence1,
			length1,length2,offset1,offset2,chroffset,chrhigh,chrpos,genomiclength,
			cdna_direction,watsonp,jump_late_p,pairpool,extraband_end,defect_rate,endalign);
}
