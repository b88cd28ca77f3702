/* libdynprog_cuda -- C ABI of the B200-native banded gap-fill DP for GMAP/GSNAP.
 *
 * This header is the drop-in boundary.  It replaces, for the five gap-fill
 * solvers only, the interface declared in the reference's src/dynprog.h:
 *
 *   reference (src/dynprog.h)                    this library
 *   -------------------------------------------  ---------------------------------
 *   Dynprog_init          dynprog.h:67-69        dpc_init
 *   Dynprog_setup         dynprog.h:37-45        dpc_setup
 *   Dynprog_term          dynprog.h:65-66        dpc_term
 *   Dynprog_new/_free     dynprog.h:51-55        dpc_ctx_new / dpc_ctx_free
 *   Dynprog_single_gap    dynprog.h:71-82        dpc_add(kind=DPC_SINGLE_GAP) ...
 *   Dynprog_cdna_gap      dynprog.h:84-97        dpc_add(kind=DPC_CDNA_GAP)
 *   Dynprog_genome_gap    dynprog.h:99-117       dpc_add(kind=DPC_GENOME_GAP)
 *   Dynprog_end5_gap      dynprog.h:119-132      dpc_add(kind=DPC_END5_GAP)
 *   Dynprog_end3_gap      dynprog.h:148-161      dpc_add(kind=DPC_END3_GAP)
 *   Dynprog_end5_splicejunction dynprog.h:134-146   dpc_add(kind=DPC_END5_SPLICEJUNCTION)   
 *   Dynprog_end3_splicejunction dynprog.h:163-175   dpc_add(kind=DPC_END3_SPLICEJUNCTION)
 *   Dynprog_microexon_int       dynprog.h:177-191   dpc_add(kind=DPC_MICROEXON_INT)         
 *   (List_T of Pair_T via Pairpool_push)         dpc_pairs (flat dpc_pair_t records,
 *                                                in the order of the returned List_T)
 *
 * The reference solves one problem per call; this library collects problems
 * (dpc_add / dpc_add_bulk), solves a whole batch on the GPU (dpc_flush +
 * dpc_wait) and hands back scores plus device-computed traceback runs, from
 * which dpc_pairs rebuilds the Pair records on the host.  The reference-named
 * functions with their exact signatures live in
 * gmap-gsnap_b200/host/dynprog_dropin.c (compiled inside a GMAP tree, see
 * INTEGRATION.md) and are "add 1 + flush + wait + pairs".
 *
 * Plain C, plain pointers and sizes.  No CPU fallback: every solver entry
 * point returns DPC_ERR_CUDA when no device/driver is usable.
 */
#ifndef DYNPROG_CUDA_H
#define DYNPROG_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants mirrored from the reference ---------------------------- */
#define DPC_UNKNOWNJUMP   (-1000000)   /* dynprog.h:30  UNKNOWNJUMP   */
#define DPC_NEG_INFINITY  (-1000000)   /* dynprog.c:119 NEG_INFINITY  */
#define DPC_UNSET         (-2000000001)/* "output parameter left untouched by the reference" */

/* Endalign_T, dynprog.h:7 */
enum { DPC_QUERYEND_GAP = 0, DPC_QUERYEND_INDELS = 1, DPC_QUERYEND_NOGAPS = 2, DPC_BEST_LOCAL = 3 };

/* Mode_T, mode.h (only reaches the DP through pairdistance_init, dynprog.c:1215) */
enum { DPC_MODE_STANDARD = 0, DPC_MODE_CMET_STRANDED = 1, DPC_MODE_CMET_NONSTRANDED = 2,
       DPC_MODE_ATOI_STRANDED = 3, DPC_MODE_ATOI_NONSTRANDED = 4 };

/* which solver a problem is for */
enum { DPC_SINGLE_GAP = 0, DPC_GENOME_GAP = 1, DPC_CDNA_GAP = 2, DPC_END5_GAP = 3, DPC_END3_GAP = 4,
       DPC_END5_SPLICEJUNCTION = 5, DPC_END3_SPLICEJUNCTION = 6, DPC_MICROEXON_INT = 7 };

/* error codes (negative return values) */
enum {
  DPC_OK = 0,
  DPC_ERR_CUDA = -1,          /* no device, driver error, kernel fault: fatal, like exit(9) in gmap.c:2287 */
  DPC_ERR_ARG = -2,           /* malformed problem (negative length where the reference abort()s, ...) */
  DPC_ERR_ALPHABET = -3,      /* query byte >= 128, or genome byte outside ACGTNX* */
  DPC_ERR_UNSUPPORTED = -4,   /* a run length the 16-bit traceback ops cannot carry (a QUERYEND_NOGAPS end of 16384+ columns) */
  DPC_ERR_STATE = -5,         /* called before dpc_init/dpc_setup, ticket out of range, ... */
  DPC_ERR_NOMEM = -6
};

/* ---- one gap-fill problem ---------------------------------------------
 * Field meaning per kind (names on the right are the reference's parameters):
 *
 *              SINGLE        GENOME            CDNA              END5            END3
 *  seq1        sequence1     sequence1         sequence1L        revsequence1    sequence1
 *  seq1R       -             -                 revsequence1R     -               -
 *  length1     length1       length1           length1L          length1         length1
 *  length1R    -             -                 length1R          -               -
 *  length2     length2       length2L          length2           length2         length2
 *  length2R    -             length2R          -                 -               -
 *  offset1     offset1       offset1           offset1L          revoffset1      offset1
 *  offset1R    -             -                 revoffset1R       -               -
 *  offset2     offset2       offset2L          offset2           revoffset2      offset2
 *  offset2R    -             revoffset2R       -                 -               -
 *  extraband   _single       _paired           _paired           _end            _end
 *
 * The two splice-junction solvers (Dynprog_end5/3_splicejunction, dynprog.c:5411-5552, 5869-6012: an end of the
 * query against the string Splicetrie builds from the anchor exon and one known far splice site) use the END5 /
 * END3 columns with three differences: the genomic side IS passed as characters, because it is not a contiguous
 * piece of the genome -- seq1R = (rev)sequence2, length2 its length, upper-case A C G T N only; offset2 =
 * (rev)offset2_anchor, offset2R = (rev)offset2_far; length2R = contlength.  endalign is ignored (the reference
 * always runs find_best_endpoint_to_queryend_indels here).
 *
 * Dynprog_microexon_int (dynprog.c:7127-7429: a microexon between two introns, found by an exact scan of the
 * query's middle piece over the intron, BoyerMoore_nt boyer-moore.c:384) uses the GENOME column: seq1 = sequence1,
 * length1, offset1, offset2 = offset2L, offset2R = revoffset2R, cdna_direction, defect_rate (length2 / length2R
 * are not looked at, as in the reference).  Results: left_prob / right_prob = *bestprob2 / *bestprob3, introntype =
 * *microintrontype; the two gapholders of the returned list carry comp '>' or '<' (dynprog.c:6982, 7021).  The scan
 * runs on the device; the MaxEnt probabilities of the hits come through the splice_prob hook.
 *
 * "rev" pointers follow the reference convention: they point at the LAST
 * character and are indexed with non-positive offsets (dynprog.c:1674).
 * The upper-case twin (sequenceuc1) is derived as toupper(seq1[i]).
 * The genomic side is never passed as characters: like the reference with
 * use_genomicseg_p == false (stage3.c:9864) it is fetched from the 2-bit
 * genome registered with dpc_setup, through get_genomic_nt semantics
 * (dynprog.c:403-441) evaluated on the device.
 * Dead reference parameters (onesidegapp, close_indels_mode, maxhoriz/vertjump,
 * sequence2*, genomicuc_ptr, use_genomicseg_p) have no field.
 */
typedef struct dpc_problem {
  const char *seq1;
  const char *seq1R;
  int32_t kind;
  int32_t endalign;
  int32_t length1, length1R, length2, length2R;
  int32_t offset1, offset1R, offset2, offset2R;
  uint32_t chroffset, chrhigh, chrpos, genomiclength;
  int32_t chrnum;
  int32_t cdna_direction;
  int32_t extraband;
  int32_t maxpeelback;
  int32_t score_threshold;
  int32_t dynprogindex;       /* value of *dynprogindex at call time (stamped on the pairs) */
  uint8_t watsonp, jump_late_p, widebandp, halfp;
  uint8_t finalp, use_probabilities_p, splicingp, reserved;
  double defect_rate;
} dpc_problem_t;

/* ---- one result ---------------------------------------------------------
 * Every field starts as DPC_UNSET (doubles: -1.0) and is written exactly when
 * the reference writes the corresponding output parameter, so "left untouched"
 * (e.g. *introntype in probability mode, dynprog.c:4071) is observable.
 * null_list != 0 means the reference returns (List_T) NULL.
 * dynprogindex_out is *dynprogindex after the call (unchanged on the early
 * returns listed in SURVEY.md 8a).
 */
typedef struct dpc_result {
  int32_t null_list;
  int32_t dynprogindex_out;
  int32_t finalscore;
  int32_t nmatches, nmismatches, nopens, nindels;
  int32_t new_leftgenomepos, new_rightgenomepos;
  int32_t exonhead, introntype;
  int32_t incompletep;          /* cdna gap: 1 when the reference sets *incompletep = true */
  int32_t npairs;               /* number of dpc_pair_t records dpc_pairs will produce */
  int32_t reserved;
  double left_prob, right_prob;
} dpc_result_t;

/* One rebuilt Pair (pairdef.h:10-47), only the fields Pairpool_push /
 * Pairpool_push_gapholder set to something other than their constants
 * (pairpool.c:169-215, 352-401).  gapp == 1 is a gapholder: querypos =
 * genomepos = -1, chars ' ', queryjump = genomejump = DPC_UNKNOWNJUMP,
 * dynprogindex 0.  gapp == 2 is the KNOWN gapholder of the splice-junction solvers
 * (Pairpool_push_gapholder(..., queryjump, genomejump, knownp = true), dynprog.c:5518): querypos carries
 * queryjump (0) and genomepos carries genomejump. */
typedef struct dpc_pair {
  int32_t querypos;
  int32_t genomepos;
  int32_t dynprogindex;
  char cdna, comp, genome;
  uint8_t gapp;
} dpc_pair_t;

/* Host hooks that stay on the CPU side of the boundary.
 * splice_prob: the reference's Maxent_hr_{donor,acceptor,antidonor,antiacceptor}_prob
 *   (maxent_hr.c:27217-27340); which = 0 donor, 1 acceptor, 2 antidonor, 3 antiacceptor.
 *   Needed for finalp (dynprog.c:3195-3287) and use_probabilities_p (3829-3903).
 * splice_known: known-splice-site lookup (dynprog.c:3377-3542); returns nonzero when
 *   the position is a known site.  which as above; sign as the reference passes it.
 *   NULL means splicing_iit == NULL.
 * With an intron-level splicing IIT (donor_typeint < 0 || acceptor_typeint < 0 in Dynprog_setup; set
 *   dpc_setup_t.intron_level) the same hook answers dynprog.c:3460-3542 instead: the host program maps which
 *   0 / 3 (donor, antiacceptor) to IIT_low_exists_signed_p(splicesitepos, sign) and which 1 / 2 (acceptor,
 *   antidonor) to IIT_high_exists_signed_p(splicesitepos + 1, sign) -- the library passes the same splicesitepos
 *   in both flavours.
 * splice_intron: IIT_exists_with_divno_signed(iit, crosstable[chrnum], pos1, pos2, sign) (dynprog.c:3600-3615):
 *   nonzero when (pos1, pos2) is a known intron.  Needed only for intron_level && !novelsplicingp, where the bridge
 *   is constrained to the given introns (dynprog.c:3552-3696). */
typedef double (*dpc_splice_prob_fn)(int which, uint32_t splice_pos, uint32_t chroffset, void *user);
typedef int (*dpc_splice_known_fn)(int which, int chrnum, uint32_t splicesitepos, int sign, void *user);
typedef int (*dpc_splice_intron_fn)(int chrnum, uint32_t pos1, uint32_t pos2, int sign, void *user);

typedef struct dpc_setup {
  const uint32_t *genome_blocks;   /* 3 x UINT4 per 32 nt: high, low, flags (genome.c:9325-9362) */
  uint64_t genome_nwords;          /* number of UINT4 in genome_blocks */
  int32_t novelsplicingp;          /* Dynprog_setup novelsplicingp_in */
  int32_t intron_level;            /* nonzero: the splicing IIT holds introns, not typed splice sites
                                      (donor_typeint < 0 || acceptor_typeint < 0, dynprog.c:3459) */
  dpc_splice_prob_fn splice_prob;
  dpc_splice_known_fn splice_known;
  void *user;
  dpc_splice_intron_fn splice_intron;
} dpc_setup_t;

typedef struct dpc_ctx dpc_ctx_t;

/* ---- lifecycle ---------------------------------------------------------- */
/* Dynprog_init (dynprog.c:1338): builds the four pairdistance tables and the
 * consistent table for `mode`, and the length caps of compute_maxlengths
 * (dynprog.c:831-852).  GMAP passes (nullgap=600, 10, maxpeelback=11, 10, 8). */
int dpc_init(int maxlookback, int extraquerygap, int maxpeelback,
             int extramaterial_end, int extramaterial_paired, int mode);
/* Dynprog_setup (dynprog.c:349): registers the genome and the splicing hooks;
 * the genome blocks are mirrored into HBM on every device that has a context. */
int dpc_setup(const dpc_setup_t *setup);
void dpc_term(void);
/* Dynprog_pairdistance (dynprog.c:1048) for any of the 4 mismatch types (0 HIGHQ .. 3 ENDQ). */
int dpc_pairdistance(int mismatchtype, int c1, int c2);
/* maxlength1 / maxlength2 of struct Dynprog_T (dynprog.c:822-829). */
void dpc_maxlengths(int *maxlength1, int *maxlength2);
const char *dpc_strerror(int code);
/* Number of usable CUDA devices (0 when there is no driver/GPU). */
int dpc_device_count(void);
/* Optional: creates the CUDA context of `device` ahead of time (callable from any thread, before dpc_setup), so
 * that a host program can overlap the ~0.5 s of driver/context start-up with its own initialisation (the
 * drop-in does it from Dynprog_init, which gmap.c:3456 calls before it loads the genome index). */
int dpc_warmup(int device);

/* ---- per-thread batch context (one Dynprog_T triple + stream) ---------- */
dpc_ctx_t *dpc_ctx_new(int device);
/* A context whose bulk call (dpc_solve) deals its chunks to several devices of this box -- "shard by read across
 * the GPUs, gather on the host": chunk j of the problem array runs on devices[j mod ndevices], results and pairs
 * come back in input order.  The ticket API of such a context runs on devices[0].  NULL when a device is unusable. */
dpc_ctx_t *dpc_ctx_new_multi(const int *devices, int ndevices);
void dpc_ctx_free(dpc_ctx_t *ctx);

/* Enqueue.  Sequences are copied at enqueue time.  Returns the ticket (>= 0,
 * consecutive) of the first problem added, or a negative error code. */
int dpc_add(dpc_ctx_t *ctx, const dpc_problem_t *problem);
int dpc_add_bulk(dpc_ctx_t *ctx, const dpc_problem_t *problems, int n);
/* Asynchronous: H2D, fill/bridge/traceback kernels, D2H on the context's stream. */
int dpc_flush(dpc_ctx_t *ctx);
/* Blocks until the batch is back; afterwards dpc_result / dpc_pairs are valid. */
int dpc_wait(dpc_ctx_t *ctx);
int dpc_result(dpc_ctx_t *ctx, int ticket, dpc_result_t *out);
/* Rebuilds the Pair records of one problem in the order of the List_T the
 * reference returns (head first).  Returns the count, or a negative code;
 * cap too small -> DPC_ERR_ARG. */
int dpc_pairs(dpc_ctx_t *ctx, int ticket, dpc_pair_t *out, int cap);
/* Forget the batch (tickets restart at 0). */
int dpc_reset(dpc_ctx_t *ctx);

/* Host threads the bulk call (dpc_solve) may use for packing, finalising and rebuilding pairs;
 * default: all cores, or the environment variable DPC_HOST_THREADS. */
int dpc_set_threads(dpc_ctx_t *ctx, int nthreads);

/* Whole batch in one call: add_bulk + flush + wait + results (+ pairs when
 * pairs != NULL; pair_off[i] = index of problem i's first record, pair_off has
 * n+1 entries).  Returns 0 or a negative code; DPC_ERR_NOMEM if pair_cap is too small. */
/* The batch is cut into chunks that run concurrently on their own streams; outputs are in input order.  The
 * problems' sequences must stay valid during the call.  A chunk takes the DEVICE PIPELINE -- the problem records
 * and the query bytes they point to are copied as they are, checks / packing / result finalisation / Pair-record
 * expansion are kernels, and results and pairs are copied straight into the caller's arrays -- unless it needs a
 * host hook per problem (known splice sites, use_probabilities_p) or a splice-junction solver; such chunks are
 * packed, finalised and rebuilt by the host threads.  Fastest when the queries of a chunk lie close together in one
 * buffer (one copy instead of a gather) and `pairs` is page-locked (dpc_host_register: the copy engine then writes
 * the records in place; otherwise they pass through a page-locked bounce buffer and a memcpy). */
int dpc_solve(dpc_ctx_t *ctx, const dpc_problem_t *problems, int n,
              dpc_result_t *results, dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off);

/* Page-locks / releases a caller-owned array (cudaHostRegister) so that dpc_solve can copy into it directly. */
int dpc_host_register(void *p, uint64_t bytes);
int dpc_host_unregister(void *p);

/* ---- measurement hooks (bench.py) -------------------------------------- */
/* After dpc_flush+dpc_wait the batch stays resident in HBM; dpc_relaunch runs
 * the device work again (kernels only, no copies) on the context's stream and
 * returns the number of kernel launches it made.  stream_out receives the
 * cudaStream_t so the caller can bracket it with events. */
int dpc_relaunch(dpc_ctx_t *ctx);
void *dpc_stream(dpc_ctx_t *ctx);
/* Waits for the work queued by dpc_relaunch and refreshes dpc_last_kernel_ms. */
int dpc_sync(dpc_ctx_t *ctx);
/* Test hook: nonzero routes every matrix through the memory-state anti-diagonal fill
 * (the routine used for bands of 64+ diagonals) instead of the register/shuffle fill.
 * Both are device code; the environment variable DPC_FORCE_GENERIC_FILL=1 does the same. */
int dpc_set_fill(int force_generic);
/* Test hook: which half serves dpc_solve -- 0 per chunk as described above, 1 host half for every chunk, 2 device
 * pipeline or DPC_ERR_STATE; 3 and 4 are 2 with the Pair records of every chunk expanded by the device / by the host
 * threads (by default a pipeline chunk takes the host route when the PCIe backlog would outlast the rebuild).
 * All of them return identical outputs. */
int dpc_set_path(int path);
/* Calibration of the roofline denominators on `device`: giga lane-operations per second of the integer ALU pipe
 * (independent VIMNMX + LOP3 chains: the pipe the fill saturates) and of the fill's add / max / select mix (adds may
 * issue on the FMA pipe).  Runs for a few milliseconds. */
int dpc_measure_int_peak(int device, double *alu_gops, double *mix_gops);
/* Device time in ms of the last dpc_flush / dpc_relaunch (CUDA events recorded on the context's stream around the
 * kernels; fill, bridge and traceback are one fused kernel per launch class). */
int dpc_last_kernel_ms(dpc_ctx_t *ctx, float *ms);
/* Work counters of the current batch (after dpc_solve: of the whole call): in-band DP cells (SURVEY.md 8d
 * definition), matrices filled, algorithmic HBM bytes of the solve kernels (descriptors, sequences, 2-bit genome in;
 * result records and staged genome characters out), bytes copied H2D / D2H. */
typedef struct dpc_stats {
  int64_t nproblems, nmatrices, cells;
  int64_t fill_bytes;
  int64_t pipeline_chunks;      /* dpc_solve: chunks that took the device pipeline ... */
  int64_t h2d_bytes, d2h_bytes;
  int32_t launches;             /* kernel launches of the last flush / relaunch / solve */
  int32_t host_chunks;          /* ... and chunks served by the host half */
} dpc_stats_t;
int dpc_get_stats(dpc_ctx_t *ctx, dpc_stats_t *out);

#ifdef __cplusplus
}
#endif
#endif /* DYNPROG_CUDA_H */
