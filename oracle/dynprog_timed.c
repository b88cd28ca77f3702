/* oracle/dynprog_timed.c -- ORACLE / MEASUREMENT SCAFFOLDING, NOT PRODUCT CODE (generated once from the signatures
 * of src/dynprog.h:71-161).  Wraps the reference's five gap-fill solvers (renamed *_cpu with objcopy, see
 * oracle/build_gmap.sh) with a per-thread timer, so that `gmap_ref_timed` reports how much of a GMAP run the
 * reference spends inside the hot path (the Amdahl bound of the whole-program bench).  The solvers are unmodified. */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <pthread.h>
#include "bool.h"
#include "types.h"
#include "genomicpos.h"
#include "chrnum.h"
#include "list.h"
#include "pairpool.h"
#include "dynprog.h"

static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static double total_s; static long total_calls; static int registered;
static __thread double my_s; static __thread long my_calls; static __thread int my_key_set;
static pthread_key_t key;
static double now (void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC,&t); return t.tv_sec + 1e-9*t.tv_nsec; }
static void report (void) { fprintf(stderr,"dynprog_timed: %ld solver calls, %.3f thread-seconds inside the five gap-fill solvers\n",total_calls,total_s); }
static void fold (void *p) { (void) p; pthread_mutex_lock(&mu); total_s += my_s; total_calls += my_calls; pthread_mutex_unlock(&mu); }
static void enter (void) {
  if (!my_key_set) {
    pthread_mutex_lock(&mu);
    if (!registered) { pthread_key_create(&key,fold); atexit(report); registered = 1; }
    pthread_mutex_unlock(&mu);
    pthread_setspecific(key,(void *) 1); my_key_set = 1;
  }
}

extern List_T Dynprog_single_gap_cpu (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches, int *nopens, int *nindels, Dynprog_T dynprog, char *sequence1, char *sequenceuc1, char *sequence2, char *sequenceuc2, int length1, int length2, int offset1, int offset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_single, double defect_rate, int close_indels_mode, bool widebandp);
List_T
Dynprog_single_gap (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches, int *nopens, int *nindels, Dynprog_T dynprog, char *sequence1, char *sequenceuc1, char *sequence2, char *sequenceuc2, int length1, int length2, int offset1, int offset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_single, double defect_rate, int close_indels_mode, bool widebandp) {
  List_T r; double t0;
  enter(); t0 = now();
  r = Dynprog_single_gap_cpu(dynprogindex,finalscore,nmatches,nmismatches,nopens,nindels,dynprog,sequence1,sequenceuc1,sequence2,sequenceuc2,length1,length2,offset1,offset2,chroffset,chrhigh,chrpos,genomiclength,cdna_direction,watsonp,jump_late_p,pairpool,extraband_single,defect_rate,close_indels_mode,widebandp);
  my_s += now() - t0; my_calls++;
  return r;
}

extern List_T Dynprog_cdna_gap_cpu (int *dynprogindex, int *finalscore, bool *incompletep, Dynprog_T dynprogL, Dynprog_T dynprogR, char *sequence1L, char *sequenceuc1L, char *revsequence1R, char *revsequenceuc1R, char *sequence2, char *sequenceuc2, int length1L, int length1R, int length2, int offset1L, int revoffset1R, int offset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_paired, double defect_rate);
List_T
Dynprog_cdna_gap (int *dynprogindex, int *finalscore, bool *incompletep, Dynprog_T dynprogL, Dynprog_T dynprogR, char *sequence1L, char *sequenceuc1L, char *revsequence1R, char *revsequenceuc1R, char *sequence2, char *sequenceuc2, int length1L, int length1R, int length2, int offset1L, int revoffset1R, int offset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_paired, double defect_rate) {
  List_T r; double t0;
  enter(); t0 = now();
  r = Dynprog_cdna_gap_cpu(dynprogindex,finalscore,incompletep,dynprogL,dynprogR,sequence1L,sequenceuc1L,revsequence1R,revsequenceuc1R,sequence2,sequenceuc2,length1L,length1R,length2,offset1L,revoffset1R,offset2,chroffset,chrhigh,chrpos,genomiclength,cdna_direction,watsonp,jump_late_p,pairpool,extraband_paired,defect_rate);
  my_s += now() - t0; my_calls++;
  return r;
}

extern List_T Dynprog_genome_gap_cpu (int *dynprogindex, int *finalscore, int *new_leftgenomepos, int *new_rightgenomepos, double *left_prob, double *right_prob, int *nmatches, int *nmismatches, int *nopens, int *nindels, int *exonhead, int *introntype, Dynprog_T dynprogL, Dynprog_T dynprogR, char *sequence1, char *sequenceuc1, char *sequence2L, char *sequenceuc2L, char *revsequence2R, char *revsequenceuc2R, int length1, int length2L, int length2R, int offset1, int offset2L, int revoffset2R, Chrnum_T chrnum, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, char *genomicuc_ptr, bool use_genomicseg_p, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_paired, double defect_rate, int maxpeelback, bool halfp, bool finalp, bool use_probabilities_p, int score_threshold, bool splicingp);
List_T
Dynprog_genome_gap (int *dynprogindex, int *finalscore, int *new_leftgenomepos, int *new_rightgenomepos, double *left_prob, double *right_prob, int *nmatches, int *nmismatches, int *nopens, int *nindels, int *exonhead, int *introntype, Dynprog_T dynprogL, Dynprog_T dynprogR, char *sequence1, char *sequenceuc1, char *sequence2L, char *sequenceuc2L, char *revsequence2R, char *revsequenceuc2R, int length1, int length2L, int length2R, int offset1, int offset2L, int revoffset2R, Chrnum_T chrnum, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, char *genomicuc_ptr, bool use_genomicseg_p, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_paired, double defect_rate, int maxpeelback, bool halfp, bool finalp, bool use_probabilities_p, int score_threshold, bool splicingp) {
  List_T r; double t0;
  enter(); t0 = now();
  r = Dynprog_genome_gap_cpu(dynprogindex,finalscore,new_leftgenomepos,new_rightgenomepos,left_prob,right_prob,nmatches,nmismatches,nopens,nindels,exonhead,introntype,dynprogL,dynprogR,sequence1,sequenceuc1,sequence2L,sequenceuc2L,revsequence2R,revsequenceuc2R,length1,length2L,length2R,offset1,offset2L,revoffset2R,chrnum,chroffset,chrhigh,chrpos,genomiclength,genomicuc_ptr,use_genomicseg_p,cdna_direction,watsonp,jump_late_p,pairpool,extraband_paired,defect_rate,maxpeelback,halfp,finalp,use_probabilities_p,score_threshold,splicingp);
  my_s += now() - t0; my_calls++;
  return r;
}

extern List_T Dynprog_end5_gap_cpu (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches, int *nopens, int *nindels, Dynprog_T dynprog, char *revsequence1, char *revsequenceuc1, char *revsequence2, char *revsequenceuc2, int length1, int length2, int revoffset1, int revoffset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_end, double defect_rate, Endalign_T endalign, bool use_genomicseg_p);
List_T
Dynprog_end5_gap (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches, int *nopens, int *nindels, Dynprog_T dynprog, char *revsequence1, char *revsequenceuc1, char *revsequence2, char *revsequenceuc2, int length1, int length2, int revoffset1, int revoffset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_end, double defect_rate, Endalign_T endalign, bool use_genomicseg_p) {
  List_T r; double t0;
  enter(); t0 = now();
  r = Dynprog_end5_gap_cpu(dynprogindex,finalscore,nmatches,nmismatches,nopens,nindels,dynprog,revsequence1,revsequenceuc1,revsequence2,revsequenceuc2,length1,length2,revoffset1,revoffset2,chroffset,chrhigh,chrpos,genomiclength,cdna_direction,watsonp,jump_late_p,pairpool,extraband_end,defect_rate,endalign,use_genomicseg_p);
  my_s += now() - t0; my_calls++;
  return r;
}

extern List_T Dynprog_end3_gap_cpu (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches, int *nopens, int *nindels, Dynprog_T dynprog, char *sequence1, char *sequenceuc1, char *sequence2, char *sequenceuc2, int length1, int length2, int offset1, int offset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_end, double defect_rate, Endalign_T endalign, bool use_genomicseg_p);
List_T
Dynprog_end3_gap (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches, int *nopens, int *nindels, Dynprog_T dynprog, char *sequence1, char *sequenceuc1, char *sequence2, char *sequenceuc2, int length1, int length2, int offset1, int offset2, Genomicpos_T chroffset, Genomicpos_T chrhigh, Genomicpos_T chrpos, Genomicpos_T genomiclength, int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool, int extraband_end, double defect_rate, Endalign_T endalign, bool use_genomicseg_p) {
  List_T r; double t0;
  enter(); t0 = now();
  r = Dynprog_end3_gap_cpu(dynprogindex,finalscore,nmatches,nmismatches,nopens,nindels,dynprog,sequence1,sequenceuc1,sequence2,sequenceuc2,length1,length2,offset1,offset2,chroffset,chrhigh,chrpos,genomiclength,cdna_direction,watsonp,jump_late_p,pairpool,extraband_end,defect_rate,endalign,use_genomicseg_p);
  my_s += now() - t0; my_calls++;
  return r;
}
