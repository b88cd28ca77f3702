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
 *   Dynprog_end5_splicejunction  dynprog.c:5411   (called by Splicetrie_solve_end5, splicetrie.c:352, 447, once per
 *   Dynprog_end3_splicejunction  dynprog.c:5869    known far splice site: runs with a known-splicing file, gmap -s)
 *
 * In dynprog.c the DEFINITIONS of these functions are renamed <name>_cpu (oracle/build_gmap.sh does it with sed on a
 * scratch copy; a maintainer would delete the seven solver bodies).  Dynprog_init/_setup/_term_cpu are still called
 * from here, because the reference's own tables must exist for the functions this library does not replace
 * (Dynprog_end5_known, Dynprog_microexon_*, ...); the renamed solver bodies are dead code: every call, including
 * the ones Dynprog_end5_known / Dynprog_end3_known make from inside dynprog.c (6474-6900), lands here.
 *
 * Batching.  The gaps of ONE alignment depend on each other (peel-back can eat the pairs of the previous fill,
 * stage3.c:5546), so the unit of parallelism is the alignment.  Instead of rewriting stage3.c's path traversal
 * as an explicit state machine, every worker thread runs DPC_FIBERS (default 16) copies of the reference's
 * worker loop (gmap.c:2254 worker_thread) as cooperative fibers, each with its own stack, Pairpool and request
 * in flight: a fiber that reaches one of the five solvers below ENQUEUES its gap (dpc_add) and yields; when
 * every fiber of the thread is parked on a gap (or finished) the scheduler flushes the collected batch to the
 * device, and the fibers resume with their results (dpc_result / dpc_pairs -> Pairpool_push).  stage3.c's
 * traversal thus fills device batches instead of solving gaps one at a time, with its decision logic untouched.
 * Two batch contexts alternate per thread (fibers read round k-1's results while they enqueue round k).
 * gmap.c changes by one call: the pthread_create of its workers becomes Dynprog_cuda_worker_create
 * (INTEGRATION.md section 3).  Called outside a fiber (single_thread(), DPC_FIBERS=1) a solver is
 * "add 1 + flush + wait + pairs": correct, and as slow as one kernel launch per gap.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <time.h>
#include <pthread.h>
#include <sys/mman.h>
#include <malloc.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

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
#include "except.h"

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
  if (dropin_donor_typeint < 0 || dropin_acceptor_typeint < 0) {
    /* intron-level IIT, dynprog.c:3460-3542: donors and antiacceptors are low ends of introns, acceptors and
       antidonors high ends (looked up one past the splice site) */
    if (which == 0 || which == 3) {
      return IIT_low_exists_signed_p(dropin_iit,dropin_crosstable[chrnum],splicesitepos,sign) == true;
    } else {
      return IIT_high_exists_signed_p(dropin_iit,dropin_crosstable[chrnum],splicesitepos+1U,sign) == true;
    }
  }
  /* dynprog.c:3377-3458 */
  return IIT_exists_with_divno_typed_signed(dropin_iit,dropin_crosstable[chrnum],
					    splicesitepos,splicesitepos+1U,type,sign) == true;
}

static int
dropin_splice_intron (int chrnum, uint32_t pos1, uint32_t pos2, int sign, void *user) {
  (void) user;
  /* the given-introns test of the constrained bridge, dynprog.c:3600-3615 */
  return IIT_exists_with_divno_signed(dropin_iit,dropin_crosstable[chrnum],pos1,pos2,sign) == true;
}

static int dropin_device (void);

static void *
dropin_warmup_thread (void *data) {
  (void) data;
  dpc_warmup(dropin_device());	/* errors surface later, in dpc_ctx_new */
  return NULL;
}

void
Dynprog_init (int maxlookback, int extraquerygap, int maxpeelback,
	      int extramaterial_end, int extramaterial_paired, Mode_T mode) {
  pthread_t warm;
  int rc;
  /* gmap.c:3456 calls this before it loads the genome and its index (3510-3620): start the CUDA context now */
  if (pthread_create(&warm,NULL,dropin_warmup_thread,NULL) == 0) pthread_detach(warm);
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
  s.splice_intron = splicesites_iit_in != NULL ? dropin_splice_intron : NULL;
  s.intron_level = splicesites_iit_in != NULL && (donor_typeint_in < 0 || acceptor_typeint_in < 0);
  if ((rc = dpc_setup(&s)) != DPC_OK) {
    dropin_fatal("dpc_setup",rc);
  }
}


/* ---- fibers: many copies of the reference's worker loop per OS thread ---------------------------------- */
#define FIBER_STACK_BYTES (8UL << 20)	/* what a default pthread gets; reserved, touched lazily */
enum { FIBER_RUNNABLE = 0, FIBER_PARKED = 1, FIBER_DONE = 2 };

#if defined(__x86_64__)
/* dropin_switch(&save_sp, load_sp): push the System V callee-saved registers, swap stacks, pop, return.  No
   signal-mask system call (swapcontext makes one per switch). */
__asm__ (".text\n"
	 ".type dropin_switch,@function\n"
	 "dropin_switch:\n"
	 "	pushq %rbp\n	pushq %rbx\n	pushq %r12\n	pushq %r13\n	pushq %r14\n	pushq %r15\n"
	 "	movq %rsp,(%rdi)\n"
	 "	movq %rsi,%rsp\n"
	 "	popq %r15\n	popq %r14\n	popq %r13\n	popq %r12\n	popq %rbx\n	popq %rbp\n"
	 "	ret\n"
	 ".size dropin_switch,.-dropin_switch\n");
extern void dropin_switch (void **save_sp, void *load_sp);
typedef void *fiber_regs_t;
#else
typedef ucontext_t fiber_regs_t;
#endif

typedef struct dropin_fiber {
  fiber_regs_t regs;
  char *stack;
  int state;
} dropin_fiber_t;

typedef struct dropin_sched {
  fiber_regs_t regs;		/* the scheduler's own context */
  dropin_fiber_t *fibers;
  int nfibers;
  dropin_fiber_t *cur;		/* fiber running right now, NULL while the scheduler runs */
  dpc_ctx_t *ctx[2];		/* ctx[fill] collects this round's gaps, ctx[fill^1] holds last round's results */
  int fill;
  int npending;
  void *(*start_routine) (void *);
  void *arg;
  int nexcept;			/* Except_stack_create calls outstanding on this OS thread */
  long nrounds, nproblems, njunction;	/* njunction: splice-junction solver calls among nproblems */
  double t_device, t_start;	/* seconds blocked in flush + wait; start of the scheduler (DPC_FIBER_STATS) */
  double t_add, t_post;		/* seconds in dpc_add; in dpc_result + dpc_pairs + Pairpool_push (DPC_FIBER_STATS) */
  int stats;
} dropin_sched_t;

static __thread dropin_sched_t *dropin_sched;
static __thread long dropin_memo_hits, dropin_memo_misses;	/* memo of device results, below */

static double
dropin_now (void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC,&t);
  return (double) t.tv_sec + 1e-9*(double) t.tv_nsec;
}

static void
dropin_to_scheduler (dropin_sched_t *s, dropin_fiber_t *f) {
#if defined(__x86_64__)
  dropin_switch(&f->regs,s->regs);
#else
  swapcontext(&f->regs,&s->regs);
#endif
}

static void
dropin_to_fiber (dropin_sched_t *s, dropin_fiber_t *f) {
  s->cur = f;
#if defined(__x86_64__)
  dropin_switch(&s->regs,f->regs);
#else
  swapcontext(&s->regs,&f->regs);
#endif
  s->cur = NULL;
}

static void
dropin_fiber_main (void) {
  dropin_sched_t *s = dropin_sched;
  dropin_fiber_t *f = s->cur;
  s->start_routine(s->arg);	/* the reference's worker_thread: loops over Inbuffer_get_request until input ends */
  f->state = FIBER_DONE;
  dropin_to_scheduler(s,f);
  abort();			/* a finished fiber is never resumed */
}

static void
dropin_fiber_init (dropin_fiber_t *f) {
  f->stack = (char *) mmap(NULL,FIBER_STACK_BYTES,PROT_READ|PROT_WRITE,MAP_PRIVATE|MAP_ANONYMOUS|MAP_NORESERVE|MAP_STACK,-1,0);
  if (f->stack == (char *) MAP_FAILED) {
    fprintf(stderr,"libdynprog_cuda drop-in: cannot reserve a fiber stack\n");
    exit(9);
  }
  mprotect(f->stack,4096,PROT_NONE);	/* guard page */
  f->state = FIBER_RUNNABLE;
#if defined(__x86_64__)
  {
    void **sp = (void **) (f->stack + FIBER_STACK_BYTES);	/* 16-byte aligned */
    *--sp = NULL;			/* return address slot of dropin_fiber_main (never used): entry rsp = 8 mod 16 */
    *--sp = (void *) dropin_fiber_main;	/* popped by the ret of dropin_switch */
    sp -= 6;				/* rbp rbx r12 r13 r14 r15 */
    memset(sp,0,6*sizeof(void *));
    f->regs = (void *) sp;
  }
#else
  getcontext(&f->regs);
  f->regs.uc_stack.ss_sp = f->stack;
  f->regs.uc_stack.ss_size = FIBER_STACK_BYTES;
  f->regs.uc_link = NULL;
  makecontext(&f->regs,dropin_fiber_main,0);
#endif
}

static int dropin_device (void);

static void *
dropin_scheduler (void *data) {
  dropin_sched_t *s = (dropin_sched_t *) data;
  int i, live, rc;
  const int stats = getenv("DPC_FIBER_STATS") != NULL;

  dropin_sched = s;
  s->stats = stats;
  s->t_start = dropin_now();
  for (i = 0; i < 2; i++) {
    if ((s->ctx[i] = dpc_ctx_new(dropin_device())) == NULL) dropin_fatal("dpc_ctx_new",DPC_ERR_CUDA);
  }
  for (i = 0; i < s->nfibers; i++) dropin_fiber_init(&s->fibers[i]);
  do {
    live = 0;
    for (i = 0; i < s->nfibers; i++) {
      dropin_fiber_t *f = &s->fibers[i];
      if (f->state == FIBER_DONE) continue;
      f->state = FIBER_RUNNABLE;
      dropin_to_fiber(s,f);	/* picks up its result of the last round, runs on to its next gap (or to the end) */
      if (f->state != FIBER_DONE) live++;
    }
    /* every live fiber is parked on a gap in ctx[fill]; all results of ctx[fill^1] have been consumed */
    if (s->npending > 0) {
      double t0 = stats ? dropin_now() : 0.0;
      if ((rc = dpc_flush(s->ctx[s->fill])) < 0) dropin_fatal("dpc_flush",rc);
      if ((rc = dpc_wait(s->ctx[s->fill])) < 0) dropin_fatal("dpc_wait",rc);
      s->nrounds++; s->nproblems += s->npending;
      if (stats) s->t_device += dropin_now() - t0;
    }
    s->fill ^= 1;
    if ((rc = dpc_reset(s->ctx[s->fill])) < 0) dropin_fatal("dpc_reset",rc);
    s->npending = 0;
  } while (live > 0);

  if (stats) {
    fprintf(stderr,"libdynprog_cuda drop-in: %d fibers, %ld device batches, %ld gaps (%.1f per batch), %.2f s in flush+wait, %.2f s in add, %.2f s in result+pairs+push, of %.2f s; memo %ld hits, %ld misses; %ld splice-junction solves\n",
	    s->nfibers,s->nrounds,s->nproblems,s->nrounds ? (double) s->nproblems/(double) s->nrounds : 0.0,
	    s->t_device,s->t_add,s->t_post,dropin_now() - s->t_start,dropin_memo_hits,dropin_memo_misses,s->njunction);
  }
  for (i = 0; i < s->nfibers; i++) munmap(s->fibers[i].stack,FIBER_STACK_BYTES);
  dpc_ctx_free(s->ctx[0]); dpc_ctx_free(s->ctx[1]);
  dropin_sched = NULL;
  free(s->fibers);
  free(s);
  return NULL;
}

/* Replaces pthread_create for the workers of gmap.c:3920 / gsnap.c: the new OS thread runs `start_routine` (the
   reference's worker_thread) DPC_FIBERS times as fibers.  DPC_FIBERS=1 gives the plain thread back. */
/* Process-wide glibc malloc policy (INTEGRATION.md section 3 says so prominently): no mmap for large blocks, no
   trimming, a 256 MB top pad per arena.  Applied once, from the first fibered worker; DPC_MALLOC_TUNING=0 skips it. */
static void
dropin_malloc_tuning (void) {
  const char *m = getenv("DPC_MALLOC_TUNING");
  if (m == NULL || atoi(m) != 0) {
    mallopt(M_MMAP_MAX,0);
    mallopt(M_TRIM_THRESHOLD,INT_MAX);
    mallopt(M_TOP_PAD,256 << 20);
  }
}

int
Dynprog_cuda_worker_create (pthread_t *thread, const pthread_attr_t *attr, void *(*start_routine) (void *), void *arg) {
  const char *e = getenv("DPC_FIBERS");
  int nfibers = e != NULL ? atoi(e) : 16;
  dropin_sched_t *s;

  if (nfibers <= 1) return pthread_create(thread,attr,start_routine,arg);
  {
    /* Many requests in flight per thread interleave their large, short-lived allocations (stage 2's position
       tables, diagonals, pair pools); with glibc's defaults every one of them is mapped, faulted in and unmapped
       again under the process-wide mmap lock.  Keep freed memory in the arenas instead (measured on the
       whole-program bench: 9.2 M -> 4.5 M page faults, 29 s -> 8 s of system time).  DPC_MALLOC_TUNING=0 skips it. */
    static pthread_once_t tuned = PTHREAD_ONCE_INIT;
    pthread_once(&tuned,dropin_malloc_tuning);
  }
  s = (dropin_sched_t *) calloc(1,sizeof(*s));
  s->fibers = (dropin_fiber_t *) calloc(nfibers,sizeof(dropin_fiber_t));
  s->nfibers = nfibers;
  s->start_routine = start_routine;
  s->arg = arg;
  return pthread_create(thread,attr,dropin_scheduler,(void *) s);
}

/* worker_thread creates and destroys the exception-frame stack of its OS thread (gmap.c:2277, 2313; except.c:46);
   with several worker loops per thread only the first creates and the last destroys it.  gmap.c is compiled
   with -DExcept_stack_create=Dynprog_cuda_except_stack_create (and _destroy). */
void
Dynprog_cuda_except_stack_create (void) {
  dropin_sched_t *s = dropin_sched;
  if (s == NULL || s->nexcept++ == 0) Except_stack_create();
}

void
Dynprog_cuda_except_stack_destroy (void) {
  dropin_sched_t *s = dropin_sched;
  if (s == NULL || --s->nexcept == 0) Except_stack_destroy();
}

/* ---- one problem through the library ------------------------------------------------------------ */
static int
dropin_device (void) {
  const char *e = getenv("DPC_DEVICE");
  return e != NULL ? atoi(e) : 0;
}

/* ---- memo of device results --------------------------------------------------------------------------- */
/* Stage 3 solves the same gap again and again: every pass over a path (stage3.c build_pairs_singles /
   _introns / _end5 / _end3, repeated by path_compute until nothing changes) peels the same pairs back and calls
   the solver with the same arguments.  A solver's outputs are a pure function of its arguments, the query bytes
   and the (constant) genome, so each worker thread keeps the device results of its recent problems in a
   direct-mapped table and replays them -- result fields and Pair records, re-stamped with the caller's current
   dynprogindex -- instead of sending the gap to the device again.  Exact by construction: a hit requires every
   scalar argument and every query byte to be equal.  DPC_MEMO=0 turns it off; DPC_MEMO=<n> sets the table size. */
typedef struct dropin_memo {
  uint64_t hash;
  dpc_problem_t key;		/* pointers cleared, dynprogindex reduced to its sign */
  dpc_result_t res;
  char *q; int qlen, qcap;
  dpc_pair_t *pairs; int npairs, pcap;
  int in_index;
  int cdna_direction;		/* of the stored problem (single / end gaps keep it out of the key) */
  int any_direction;		/* single / end gaps: the result does not depend on cdna_direction */
  int valid;
} dropin_memo_t;

static __thread dropin_memo_t *dropin_memo_tab;
static __thread int dropin_memo_n = -1;		/* -1: not configured yet; 0: off */

static int
dropin_memo_spans (const dpc_problem_t *p, const char **a, int *na, const char **b, int *nb) {
  *a = *b = NULL; *na = *nb = 0;
  if (p->seq1 == NULL || p->length1 <= 0 || p->length1 > 4096) return 0;
  if (p->kind == DPC_END5_SPLICEJUNCTION || p->kind == DPC_END3_SPLICEJUNCTION) return 0;   /* every call has its own junction string */
  if (p->kind == DPC_MICROEXON_INT) return 0;		/* rare, and its answer depends on the whole intron */
  if (p->kind == DPC_END5_GAP) { *a = p->seq1 - (p->length1 - 1); *na = p->length1; }
  else { *a = p->seq1; *na = p->length1; }
  if (p->kind == DPC_CDNA_GAP) {
    if (p->seq1R == NULL || p->length1R <= 0 || p->length1R > 4096) return 0;
    *b = p->seq1R - (p->length1R - 1); *nb = p->length1R;
  }
  return 1;
}

static uint64_t
dropin_fnv (uint64_t h, const void *data, size_t n) {
  const unsigned char *c = (const unsigned char *) data;
  size_t i;
  for (i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ULL; }
  return h;
}

/* Returns the table slot for p (filling *key, *hash); *hit tells whether it already holds p's result. */
static dropin_memo_t *
dropin_memo_find (const dpc_problem_t *p, dpc_problem_t *key, uint64_t *hash, int *hit) {
  const char *a, *b;
  int na, nb, k;
  dropin_memo_t *e;

  *hit = 0;
  if (dropin_memo_n < 0) {
    const char *env = getenv("DPC_MEMO");
    dropin_memo_n = env != NULL ? atoi(env) : 8192;
    if (dropin_memo_n > 0) dropin_memo_tab = (dropin_memo_t *) calloc((size_t) dropin_memo_n,sizeof(dropin_memo_t));
    if (dropin_memo_tab == NULL) dropin_memo_n = 0;
  }
  if (dropin_memo_n == 0 || !dropin_memo_spans(p,&a,&na,&b,&nb)) return NULL;
  *key = *p;
  key->seq1 = key->seq1R = NULL;
  key->dynprogindex = p->dynprogindex > 0 ? 1 : -1;
  /* the solvers only look at the quality class of defect_rate (dynprog.h:27-28, dynprog.c:4480-4500) */
  key->defect_rate = p->defect_rate < DEFECT_HIGHQ ? 0.0 : p->defect_rate < DEFECT_MEDQ ? 1.0 : 2.0;
  /* score_threshold is only read in probability mode (dynprog.c:3829-4081) */
  if (!p->use_probabilities_p) key->score_threshold = 0;
  /* single and end gaps see cdna_direction only in add_genomeskip's intron test of genome runs of 9 or more
     (dynprog.c:2416-2512): GMAP solves every gap once per direction, and a result without such a run serves both */
  if (p->kind != DPC_GENOME_GAP && p->kind != DPC_CDNA_GAP) key->cdna_direction = 0;
  /* genomiclength only bounds the segment (get_genomic_nt returns '*' outside it, dynprog.c:415) and, on the Crick
     strand, places it: position = chrpos + genomiclength - 1 - genomicpos (428).  When every position the solver can
     touch lies inside the segment, two calls that differ only in how far the segment extends are the same
     problem: Watson keys drop genomiclength, Crick keys carry chrpos + genomiclength instead.  (get_splicesite_probs
     uses the same two forms, 3219-3243.) */
  {
    int lo, hi;		/* segment-relative positions the solver reads: [lo, hi] */
    switch (p->kind) {
    case DPC_END5_GAP: lo = p->offset2 - (p->length2 - 1); hi = p->offset2; break;
    case DPC_GENOME_GAP:
      lo = p->offset2 < p->offset2R - (p->length2R - 1) ? p->offset2 : p->offset2R - (p->length2R - 1);
      hi = p->offset2 + p->length2 - 1 > p->offset2R ? p->offset2 + p->length2 - 1 : p->offset2R;
      break;
    default: lo = p->offset2; hi = p->offset2 + p->length2 - 1; break;
    }
    /* ... provided the segment itself starts inside the chromosome: otherwise every position reads '*' (415-419),
       which depends on chroffset + chrpos alone and must stay visible in the key */
    if (lo >= 0 && hi >= lo && (Genomicpos_T) hi < p->genomiclength &&
	p->chroffset + p->chrpos >= p->chroffset && p->chroffset + p->chrpos < p->chrhigh) {
      if (!p->watsonp) key->chrpos = p->chrpos + p->genomiclength;
      key->genomiclength = 0;
    }
  }
  *hash = dropin_fnv(dropin_fnv(dropin_fnv(1469598103934665603ULL,key,sizeof(*key)),a,(size_t) na),b,(size_t) nb);
  /* direction-independent results live in the slot of the hash; a single / end gap whose result does depend on
     cdna_direction (rare) lives one or two slots further, by direction, so the two directions do not evict each other */
  for (k = 0; k < 2; k++) {
    e = &dropin_memo_tab[(*hash + (uint64_t) (k == 0 ? 0 : p->cdna_direction > 0 ? 2 : 1)) % (uint64_t) dropin_memo_n];
    if (e->valid && e->hash == *hash && e->qlen == na + nb && memcmp(&e->key,key,sizeof(*key)) == 0 &&
	memcmp(e->q,a,(size_t) na) == 0 && memcmp(e->q + na,b,(size_t) nb) == 0 &&
	(e->any_direction || e->cdna_direction == p->cdna_direction)) {
      *hit = 1;
      return e;
    }
  }
  e = &dropin_memo_tab[*hash % (uint64_t) dropin_memo_n];
  return e;
}

static void
dropin_memo_store (dropin_memo_t *e, const dpc_problem_t *p, const dpc_problem_t *key, uint64_t hash,
		   const dpc_result_t *r, const dpc_pair_t *pairs, int npairs) {
  const char *a, *b;
  int na, nb, i, any_direction;
  if (!dropin_memo_spans(p,&a,&na,&b,&nb)) return;
  /* a genome run of MICROINTRON_LENGTH (9, dynprog.c:139) or more shows up as a gapholder record or, as dashes,
     in nindels (counted before end gaps strip leading indel pairs, dynprog.c:5265) */
  any_direction = r->nindels != DPC_UNSET && r->nindels < 9;
  for (i = 0; i < npairs; i++) if (pairs[i].gapp) any_direction = 0;
  /* an end gap the reference nullified (nmatches + 1 < nmismatches, dynprog.c:5259-5262) returns no pairs, so a
     genome run that became a gapholder is invisible here: keep such a result for its own direction only */
  if ((p->kind == DPC_END5_GAP || p->kind == DPC_END3_GAP) && npairs == 0 &&
      r->nmatches != DPC_UNSET && r->nmatches + 1 < r->nmismatches) any_direction = 0;
  if (!any_direction && key->cdna_direction == 0 && p->cdna_direction != 0) {
    e = &dropin_memo_tab[(hash + (uint64_t) (p->cdna_direction > 0 ? 2 : 1)) % (uint64_t) dropin_memo_n];
  }
  if (na + nb > e->qcap) { e->qcap = 2*(na + nb) + 64; e->q = (char *) realloc(e->q,(size_t) e->qcap); }
  if (npairs > e->pcap) { e->pcap = 2*npairs + 64; e->pairs = (dpc_pair_t *) realloc(e->pairs,(size_t) e->pcap*sizeof(dpc_pair_t)); }
  if (e->q == NULL || (npairs > 0 && e->pairs == NULL)) { e->valid = 0; e->qcap = e->pcap = 0; return; }
  memcpy(e->q,a,(size_t) na);
  if (nb > 0) memcpy(e->q + na,b,(size_t) nb);
  e->qlen = na + nb;
  if (npairs > 0) memcpy(e->pairs,pairs,(size_t) npairs*sizeof(dpc_pair_t));
  e->npairs = npairs;
  e->key = *key; e->hash = hash; e->res = *r; e->in_index = p->dynprogindex;
  e->cdna_direction = p->cdna_direction;
  e->any_direction = any_direction;
  e->valid = 1;
}

/* dpc_pairs lists the records head first; Pairpool_push prepends (pairpool.c:169-215) */
static List_T
dropin_push_pairs (const dpc_pair_t *rec, int n, int dynprogindex, Pairpool_T pairpool) {
  List_T pairs = NULL;
  int i;
  for (i = n - 1; i >= 0; i--) {
    const dpc_pair_t *q = &rec[i];
    if (q->gapp == 2) {		/* the known gapholder of the splice-junction solvers, dynprog.c:5518, 5977 */
      pairs = Pairpool_push_gapholder(pairs,pairpool,/*queryjump*/q->querypos,/*genomejump*/q->genomepos,/*knownp*/true);
    } else if (q->gapp) {
      pairs = Pairpool_push_gapholder(pairs,pairpool,/*queryjump*/UNKNOWNJUMP,/*genomejump*/UNKNOWNJUMP,/*knownp*/false);
      if (q->comp != ' ') {	/* the intron gapholders of a microexon carry their gap character, dynprog.c:6979-6982 */
	((Pair_T) List_head(pairs))->comp = q->comp;
      }
    } else {
      pairs = Pairpool_push(pairs,pairpool,q->querypos,q->genomepos,q->cdna,q->comp,q->genome,dynprogindex);
    }
  }
  return pairs;
}

static List_T
dropin_solve (dpc_result_t *r, const dpc_problem_t *p, Pairpool_T pairpool) {
  List_T pairs = NULL;
  dropin_sched_t *s = dropin_sched;
  dpc_ctx_t *ctx;
  dpc_problem_t key;
  dropin_memo_t *memo;
  uint64_t hash = 0;
  double tpost = 0.0;
  int rc, ticket, n, hit;

  memo = dropin_memo_find(p,&key,&hash,&hit);
  if (hit) {
    *r = memo->res;
    r->dynprogindex_out = p->dynprogindex + (memo->res.dynprogindex_out - memo->in_index);
    dropin_memo_hits++;
    return dropin_push_pairs(memo->pairs,memo->npairs,p->dynprogindex,pairpool);
  }
  dropin_memo_misses++;
#ifdef DPC_MEMO_DEBUG
  fprintf(stderr,"MISS %d %d | %d %d %d %d | %d %d %d %d | %u %u | %d %d %d %d %d | %d%d%d%d%d%d%d | %g | %llx\n",p->kind,p->endalign,p->length1,p->length1R,p->length2,p->length2R,
	  p->offset1,p->offset1R,p->offset2,p->offset2R,p->chrpos,p->genomiclength,p->chrnum,p->cdna_direction,p->extraband,p->maxpeelback,p->score_threshold,
	  p->watsonp,p->jump_late_p,p->widebandp,p->halfp,p->finalp,p->use_probabilities_p,p->splicingp,key.defect_rate,(unsigned long long) dropin_fnv(1469598103934665603ULL,p->seq1 ? (p->kind == DPC_END5_GAP ? p->seq1 - (p->length1-1) : p->seq1) : "",p->length1 > 0 ? p->length1 : 0));
#endif

  if (s != NULL && s->cur != NULL) {
    /* inside a fiber: enqueue into this round's batch and park until the scheduler has run it on the device */
    dropin_fiber_t *f = s->cur;
    double t0 = s->stats ? dropin_now() : 0.0;
    ctx = s->ctx[s->fill];
    if ((ticket = dpc_add(ctx,p)) < 0) dropin_fatal("dpc_add",ticket);
    s->npending++;
    if (p->kind == DPC_END5_SPLICEJUNCTION || p->kind == DPC_END3_SPLICEJUNCTION) s->njunction++;
    f->state = FIBER_PARKED;
    if (s->stats) s->t_add += dropin_now() - t0;
    dropin_to_scheduler(s,f);
    if (s->stats) tpost = dropin_now();
  } else {
    if (dropin_ctx == NULL && (dropin_ctx = dpc_ctx_new(dropin_device())) == NULL) {
      dropin_fatal("dpc_ctx_new",DPC_ERR_CUDA);
    }
    ctx = dropin_ctx;
    if ((rc = dpc_reset(ctx)) < 0) dropin_fatal("dpc_reset",rc);
    if ((ticket = dpc_add(ctx,p)) < 0) dropin_fatal("dpc_add",ticket);
    if ((rc = dpc_flush(ctx)) < 0) dropin_fatal("dpc_flush",rc);
    if ((rc = dpc_wait(ctx)) < 0) dropin_fatal("dpc_wait",rc);
  }
  if ((rc = dpc_result(ctx,ticket,r)) < 0) dropin_fatal("dpc_result",rc);
  if (r->npairs > dropin_pairs_cap) {
    dropin_pairs_cap = 2*r->npairs + 256;
    dropin_pairs = (dpc_pair_t *) realloc(dropin_pairs,dropin_pairs_cap*sizeof(dpc_pair_t));
  }
  if ((n = dpc_pairs(ctx,ticket,dropin_pairs,dropin_pairs_cap)) < 0) dropin_fatal("dpc_pairs",n);
  if (memo != NULL) dropin_memo_store(memo,p,&key,hash,r,dropin_pairs,n);
  pairs = dropin_push_pairs(dropin_pairs,n,p->dynprogindex,pairpool);
  if (tpost != 0.0) s->t_post += dropin_now() - tpost;
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


static List_T
dropin_splicejunction (int kind, int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches,
		       int *nopens, int *nindels, char *sequence1, char *sequence2,
		       int length1, int length2, int offset1, int offset2_anchor, int offset2_far,
		       Genomicpos_T chroffset, Genomicpos_T chrhigh,
		       Genomicpos_T chrpos, Genomicpos_T genomiclength,
		       int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
		       int extraband_end, double defect_rate, int contlength) {
  dpc_problem_t p;
  dpc_result_t r;
  List_T pairs;

  dropin_common(&p,kind,*dynprogindex,chroffset,chrhigh,chrpos,genomiclength,
		cdna_direction,watsonp,jump_late_p,extraband_end,defect_rate);
  p.seq1 = sequence1; p.seq1R = sequence2;	/* the splice-junction string travels as characters */
  p.length1 = length1; p.length2 = length2; p.length2R = contlength;
  p.offset1 = offset1; p.offset2 = offset2_anchor; p.offset2R = offset2_far;
  pairs = dropin_solve(&r,&p,pairpool);
  OUT(finalscore,r.finalscore);
  OUT(nmatches,r.nmatches); OUT(nmismatches,r.nmismatches); OUT(nopens,r.nopens); OUT(nindels,r.nindels);
  *dynprogindex = r.dynprogindex_out;
  return pairs;
}

List_T
Dynprog_end5_splicejunction (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches,
			     int *nopens, int *nindels, Dynprog_T dynprog,
			     char *revsequence1, char *revsequenceuc1,
			     char *revsequence2, char *revsequenceuc2,
			     int length1, int length2, int revoffset1, int revoffset2_anchor, int revoffset2_far,
			     Genomicpos_T chroffset, Genomicpos_T chrhigh,
			     Genomicpos_T chrpos, Genomicpos_T genomiclength,
			     int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
			     int extraband_end, double defect_rate, int contlength) {
  (void) dynprog; (void) revsequenceuc1; (void) revsequenceuc2;
  return dropin_splicejunction(DPC_END5_SPLICEJUNCTION,dynprogindex,finalscore,nmatches,nmismatches,nopens,nindels,
			       revsequence1,revsequence2,length1,length2,revoffset1,revoffset2_anchor,revoffset2_far,
			       chroffset,chrhigh,chrpos,genomiclength,cdna_direction,watsonp,jump_late_p,pairpool,
			       extraband_end,defect_rate,contlength);
}

List_T
Dynprog_end3_splicejunction (int *dynprogindex, int *finalscore, int *nmatches, int *nmismatches,
			     int *nopens, int *nindels, Dynprog_T dynprog,
			     char *sequence1, char *sequenceuc1,
			     char *sequence2, char *sequenceuc2,
			     int length1, int length2, int offset1, int offset2_anchor, int offset2_far,
			     Genomicpos_T chroffset, Genomicpos_T chrhigh,
			     Genomicpos_T chrpos, Genomicpos_T genomiclength,
			     int cdna_direction, bool watsonp, bool jump_late_p, Pairpool_T pairpool,
			     int extraband_end, double defect_rate, int contlength) {
  (void) dynprog; (void) sequenceuc1; (void) sequenceuc2;
  return dropin_splicejunction(DPC_END3_SPLICEJUNCTION,dynprogindex,finalscore,nmatches,nmismatches,nopens,nindels,
			       sequence1,sequence2,length1,length2,offset1,offset2_anchor,offset2_far,
			       chroffset,chrhigh,chrpos,genomiclength,cdna_direction,watsonp,jump_late_p,pairpool,
			       extraband_end,defect_rate,contlength);
}


/* Dynprog_microexon_int, dynprog.c:7127-7429 (called by traverse_genome_gap, stage3.c:5915, when the intron solution
   is poor): the exact-match scans of the microexon candidates run on the device, the MaxEnt probabilities of the
   hits come through the hook. */
List_T
Dynprog_microexon_int (double *bestprob2, double *bestprob3, int *dynprogindex, int *microintrontype,
		       char *sequence1, char *sequenceuc1,
		       char *sequence2L, char *sequenceuc2L,
		       char *revsequence2R, char *revsequenceuc2R,
		       int length1, int length2L, int length2R,
		       int offset1, int offset2L, int revoffset2R, int cdna_direction,
		       char *queryseq, char *queryuc, char *genomicseg, char *genomicuc,
		       Genomicpos_T chroffset, Genomicpos_T chrhigh,
		       Genomicpos_T chrpos, Genomicpos_T genomiclength, bool watsonp,
		       bool use_genomicseg_p, Pairpool_T pairpool, double defect_rate) {
  dpc_problem_t p;
  dpc_result_t r;
  List_T pairs;

  (void) sequenceuc1; (void) sequence2L; (void) sequenceuc2L; (void) revsequence2R; (void) revsequenceuc2R;
  (void) queryseq; (void) queryuc; (void) genomicseg; (void) genomicuc; (void) use_genomicseg_p;
  dropin_common(&p,DPC_MICROEXON_INT,*dynprogindex,chroffset,chrhigh,chrpos,genomiclength,
		cdna_direction,watsonp,/*jump_late_p*/false,/*extraband*/0,defect_rate);
  p.seq1 = sequence1;
  p.length1 = length1; p.length2 = length2L; p.length2R = length2R;
  p.offset1 = offset1; p.offset2 = offset2L; p.offset2R = revoffset2R;
  pairs = dropin_solve(&r,&p,pairpool);
  *bestprob2 = r.left_prob; *bestprob3 = r.right_prob;
  *microintrontype = r.introntype;
  *dynprogindex = r.dynprogindex_out;
  return pairs;
}
