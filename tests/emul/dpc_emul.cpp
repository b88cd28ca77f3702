/* tests/emul/dpc_emul.cpp -- TEST SCAFFOLDING, NOT PRODUCT CODE.
 *
 * Compiles the warp-level routines of gmap-gsnap_b200/csrc/dpc_core.h with g++ and a single
 * lane, plus the real host side (dpc_host.h: packing, finalisation, pair rebuild), so that the
 * CPU test-suite can check index arithmetic, tie-breaks and boundary rules against the oracle on
 * a machine without a GPU.  libdynprog_cuda never links or loads this file; it has no CPU path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../gmap-gsnap_b200/csrc/dpc_host.h"
#include "../../gmap-gsnap_b200/csrc/dpc_rows.h"
#include "../../gmap-gsnap_b200/csrc/dpc_pipe.h"

extern "C" int emul_init(int maxlookback, int extraquerygap, int maxpeelback, int end, int paired, int mode) {
  return dpc::host_init(maxlookback, extraquerygap, maxpeelback, end, paired, mode);
}
extern "C" int emul_setup(const dpc_setup_t *s) {
  dpc::Globals &g = dpc::G();
  g.setup = *s;
  g.genome_nbases = s->genome_nwords / 3 * 32;
  g.setup_done = true;
  return 0;
}
static int g_force_generic = 0;
static int g_no_gout = 0;
extern "C" int emul_set_fill(int force_generic) { g_force_generic = force_generic & 1; g_no_gout = (force_generic >> 1) & 1; return 0; }
extern "C" int emul_pairdistance(int type, int c1, int c2) { return dpc::G().P[type & 3][c1 & 127][c2 & 127]; }

/* Arena memory for one problem.  Under AddressSanitizer (tools: profiles/r2_sanitizer_host.log) every block has exactly
 * the size dpc_layout computed, so that a device routine stepping one byte outside its arena, its HBM scratch or its
 * staged-genome span is reported; otherwise a vector with slack. */
struct Block {
  std::vector<uint8_t> v;
  void *exact;
  Block() : exact(NULL) {}
  ~Block() { free(exact); }
  uint8_t *get(size_t bytes, int fill) {
#if defined(__SANITIZE_ADDRESS__)
    free(exact);
    if (posix_memalign(&exact, 16, bytes ? bytes : 1) != 0) abort();
    memset(exact, fill, bytes);
    return (uint8_t *)exact;
#else
    v.assign(bytes + 64, (uint8_t)fill);
    return v.data() + ((16 - ((uintptr_t)v.data() & 15)) & 15);
#endif
  }
};

static int g_pipe = 0;
extern "C" int emul_set_path(int pipe) { g_pipe = pipe; return 0; }

/* The DEVICE PIPELINE of dpc_solve (dpc_pipe.h: prepare / finish / expand per problem) run on the CPU: queries
 * gathered into one pool, descriptors from dpc_prepare_one, results from dpc_finish_one, pairs from dpc_expand_one
 * with 1 lane (even problems) or 32 round-robin "lanes" (odd problems: every lane's share, one after the other). */
static int emul_solve_pipe(const dpc_problem_t *problems, int n, dpc_result_t *results,
                           dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off) {
  const dpc::Globals &g = dpc::G();
  static const uint32_t class_bytes[6] = { 6 << 10, 13 << 10, 6 << 10, 13 << 10, 6 << 10, 0 };
  std::vector<uint8_t> pool;
  std::vector<dpc_problem_t> hp(problems, problems + n);
  for (int i = 0; i < n; i++) {          /* gather (the library copies a contiguous range instead when it can) */
    dpc_problem_t &q = hp[i];
    const int len = q.kind == DPC_CDNA_GAP ? q.offset1R - q.offset1 + 1 : q.length1;
    if (len <= 0 || !q.seq1) continue;
    const char *first = q.seq1 - (q.kind == DPC_END5_GAP ? len - 1 : 0);
    const size_t at = pool.size();
    pool.insert(pool.end(), (const uint8_t *)first, (const uint8_t *)first + len);
    const int64_t delta = (int64_t)at - (int64_t)(uintptr_t)first;
    q.seq1 = (const char *)(uintptr_t)((uint64_t)(uintptr_t)q.seq1 + (uint64_t)delta);
    if (q.seq1R) q.seq1R = (const char *)(uintptr_t)((uint64_t)(uintptr_t)q.seq1R + (uint64_t)delta);
  }
  pool.resize(pool.size() + 64);
  PrepEnv env;
  env.maxlength1 = g.maxlength1; env.maxlength2 = g.maxlength2; env.genome_nbases = g.genome_nbases;
  env.novelsplicingp = g.setup.novelsplicingp; env.fillmode = g_force_generic ? 1 : 2;
  for (int k = 0; k < 6; k++) env.class_bytes[k] = class_bytes[k];
  env.qbase = 0; env.qbytes = pool.size();
  std::vector<uint16_t> ovfbuf(1 << 22);
  unsigned int used = 0;
  OvfArena ovf; ovf.ops = ovfbuf.data(); ovf.used = &used; ovf.cap = (unsigned int)ovfbuf.size();
  Lanes one; one.lane = 0; one.n = 1;
  GenericFill gfill; RowFill rfill;
  Block arena, goutb;
  int64_t out = 0;
  std::vector<dpc_pair_t> st;
  for (int i = 0; i < n; i++) {
    DevProb d; PrepOut o; DevRes dr;
    if (g.setup.splice_known || problems[i].use_probabilities_p || problems[i].kind > DPC_END3_GAP) return DPC_ERR_STATE;
    const int rc = dpc_prepare_one(hp[i], env, pool.data(), d, results[i], o);
    if (rc < 0) { if (getenv("EMUL_DEBUG")) fprintf(stderr, "emul pipe: problem %d kind %d rc %d seq1 %p len %d qbytes %llu\n", i, problems[i].kind, rc, (void *)hp[i].seq1, problems[i].length1, (unsigned long long)env.qbytes); return rc; }
    if (pair_off) pair_off[i] = out;
    if (rc == DPC_PREP_HOST) continue;
    uint8_t *gout = goutb.get((size_t)o.gout, 0xEE);        /* exactly the span dpc_prepare_one asked for */
    if (o.gout) d.gout = 0;
    ArenaLayout a;
    dpc_layout(d, a, env.fillmode);
    memset(&dr, 0, sizeof dr);
    uint8_t *base = arena.get(a.total, 0xAB);
    if (g_force_generic) dpc_solve_problem<GenericFill, -1, -1>(d, pool.data(), g.setup.genome_blocks, &g.tables, base, a.total, base, &dr, ovf, gout, gfill, one);
    else dpc_solve_problem<RowFill, -1, -1>(d, pool.data(), g.setup.genome_blocks, &g.tables, base, a.total, base, &dr, ovf, gout, rfill, one);
    if (dr.status & DPC_ST_OVF_LOST) return DPC_ERR_NOMEM;
    const uint16_t *ops = (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? ovfbuf.data() + dr.ovf : dr.ops;
    const uint8_t *staged = (i & 2) ? gout : NULL;              /* also exercise the decode-again path */
    if (!staged) d.gout = DPC_NO_GOUT;
    const int np = dpc_expand_one(problems[i], d, dr, ops, pool.data(), staged, g.setup.genome_blocks, &g.tables, NULL, one);
    if (dpc_finish_one(problems[i], dr, np, results[i])) {
      const int cL = (int)results[i].left_prob, cR = (int)results[i].right_prob;
      results[i].left_prob = dpc::Batch::site_prob(problems[i], true, cL, false);
      results[i].right_prob = dpc::Batch::site_prob(problems[i], false, cR, false);
    }
    if (pairs && np) {
      if (out + np > pair_cap) return DPC_ERR_NOMEM;
      st.assign((size_t)np + 2, dpc_pair_t());
      memset(st.data(), 0x5A, st.size() * sizeof(dpc_pair_t));
      if (i & 1) {
        for (int l = 0; l < 32; l++) { Lanes ln; ln.lane = l; ln.n = 32; if (dpc_expand_one(problems[i], d, dr, ops, pool.data(), staged, g.setup.genome_blocks, &g.tables, st.data(), ln) != np) return -101; }
      } else if (dpc_expand_one(problems[i], d, dr, ops, pool.data(), staged, g.setup.genome_blocks, &g.tables, st.data(), one) != np) return -101;
      if ((unsigned char)st[(size_t)np].comp != 0x5A) return -102;          /* wrote past its block */
      memcpy(pairs + out, st.data(), (size_t)np * sizeof(dpc_pair_t));
    }
    out += np;
  }
  if (pair_off) pair_off[n] = out;
  return 0;
}

extern "C" int emul_solve(const dpc_problem_t *problems, int n, dpc_result_t *results,
                          dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off) {
  if (g_pipe) return emul_solve_pipe(problems, n, results, pairs, pair_cap, pair_off);
  dpc::Batch b;
  for (int i = 0; i < n; i++) {
    int t = b.add(problems[i]);
    if (t < 0) return t;
  }
  std::vector<DevRes> dres(b.dprobs.size());
  std::vector<uint16_t> ovfbuf(1 << 22);
  unsigned int used = 0;
  OvfArena ovf; ovf.ops = ovfbuf.data(); ovf.used = &used; ovf.cap = (unsigned int)ovfbuf.size();
  Lanes ln; ln.lane = 0; ln.n = 1;
  GenericFill gfill;     /* single lane: memory-state fill + serial walk */
  RowFill rfill;         /* 32 simulated lanes: row-sweep fill + lane-parallel walk (dpc_vec.h host build) */
  Block arena, scratchb;
  std::vector<uint8_t> gout((size_t)b.gout_total + 64, 0xEE);     /* the staged genome characters the device returns */
  b.pool_align(16);
  for (size_t k = 0; k < b.dprobs.size(); k++) {
    ArenaLayout a;
    dpc_layout(b.dprobs[k], a, g_force_generic ? 1 : 2);
    memset(&dres[k], 0, sizeof(DevRes));
    /* odd problems keep everything in one arena, even ones put the bulk region in a separate "HBM scratch" */
    const uint32_t arena_bytes = (k & 1) ? a.total : a.small;
    uint8_t *base = arena.get(arena_bytes, 0xAB);
    uint8_t *sbase = scratchb.get(a.bulk, 0xCD);
    if (g_force_generic)
      dpc_solve_problem<GenericFill, -1, -1>(b.dprobs[k], b.pool.data(), dpc::G().setup.genome_blocks, &dpc::G().tables, base, arena_bytes, sbase, &dres[k], ovf, gout.data(), gfill, ln);
    else
      dpc_solve_problem<RowFill, -1, -1>(b.dprobs[k], b.pool.data(), dpc::G().setup.genome_blocks, &dpc::G().tables, base, arena_bytes, sbase, &dres[k], ovf, gout.data(), rfill, ln);
  }
  b.gout_host = g_no_gout ? NULL : gout.data();     /* NULL: the rebuild decodes the 2-bit genome itself (ticket users without the stream) */
  if (!b.micro.empty()) {
    /* Dynprog_microexon_int: the device's exact-match scans, hits written back to front to show that their order is free */
    std::vector<uint32_t> hits((size_t)b.hits_total + 1), count(b.scans.size(), 0);
    for (size_t q = 0; q < b.scans.size(); q++)
      for (int j = b.scans[q].npos - 1; j >= 0; j--)
        if (dpc_scan_match(b.scans[q], dpc::G().setup.genome_blocks, dpc::G().genome_nbases, j)) hits[b.scans[q].hits_off + count[q]++] = (uint32_t)j;
    for (size_t k = 0; k < b.micro.size(); k++) b.finalize_micro(b.micro[k], hits.data(), count.data());
  }
  int64_t out = 0;
  dpc::Scratch sc;
  std::vector<dpc_pair_t> st;
  for (int i = 0; i < n; i++) {
    dpc::HostProb &h = b.probs[i];
    if (pair_off) pair_off[i] = out;
    if (h.micro >= 0) {
      const std::vector<dpc_pair_t> &v = b.micro[(size_t)h.micro].pairs;
      if (pairs && !v.empty()) {
        if (out + (int64_t)v.size() > pair_cap) return DPC_ERR_NOMEM;
        memcpy(pairs + out, v.data(), v.size() * sizeof(dpc_pair_t));
      }
      out += (int64_t)v.size();
    }
    if (h.dev >= 0) {
      const DevRes &dr = dres[h.dev];
      const uint16_t *ops = (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? ovfbuf.data() + dr.ovf : dr.ops;
      if (dr.status & DPC_ST_OVF_LOST) return DPC_ERR_NOMEM;
      b.finalize(i, dr, ops, sc);
      st.resize((size_t)b.max_pairs(i));
      int k = b.rebuild(i, dr, ops, st.data(), sc);
      if (k != b.R(i).npairs) return -100;
      if (pairs) {
        if (out + k > pair_cap) return DPC_ERR_NOMEM;
        if (k) memcpy(pairs + out, st.data(), (size_t)k * sizeof(dpc_pair_t));
      }
      out += k;
    }
    results[i] = b.R(i);
  }
  if (pair_off) pair_off[n] = out;
  return 0;
}

/* Host-side micro-benchmark aid (no GPU needed): solves once with the simulated device routines, then times
 * `reps` passes of the Pair rebuild alone.  Returns nanoseconds per problem of one pass. */
#include <time.h>
/* Host-side micro-benchmark aid: nanoseconds per problem of packing (Batch::add_ext) and of finalisation. */
extern "C" void emul_time_pack_finalize(const dpc_problem_t *problems, int n, int reps, double out[2]) {
  std::vector<dpc_result_t> results((size_t)n);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int rep = 0; rep < reps; rep++) { dpc::Batch b; b.add_ext(problems, results.data(), n); }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  out[0] = ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / reps / n;
  dpc::Batch b;
  if (b.add_ext(problems, results.data(), n) < 0) { out[1] = -1; return; }
  std::vector<DevRes> dres(b.dprobs.size());
  std::vector<uint16_t> ovfbuf(1 << 22);
  unsigned int used = 0;
  OvfArena ovf; ovf.ops = ovfbuf.data(); ovf.used = &used; ovf.cap = (unsigned int)ovfbuf.size();
  Lanes ln; ln.lane = 0; ln.n = 1;
  RowFill rfill;
  std::vector<uint8_t> arena, gout((size_t)b.gout_total + 64, 0);
  b.pool_align(16);
  for (size_t k = 0; k < b.dprobs.size(); k++) {
    ArenaLayout a;
    dpc_layout(b.dprobs[k], a, 2);
    arena.assign(a.total + 64, 0);
    uint8_t *base = arena.data() + ((16 - ((uintptr_t)arena.data() & 15)) & 15);
    memset(&dres[k], 0, sizeof(DevRes));
    dpc_solve_problem<RowFill, -1, -1>(b.dprobs[k], b.pool.data(), dpc::G().setup.genome_blocks, &dpc::G().tables, base, a.total, base, &dres[k], ovf, gout.data(), rfill, ln);
  }
  b.gout_host = gout.data();
  dpc::Scratch sc;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int rep = 0; rep < reps; rep++)
    for (int i = 0; i < n; i++) if (b.probs[i].dev >= 0) { const DevRes &dr = dres[b.probs[i].dev]; b.finalize(i, dr, dr.nopsL + dr.nopsR > DPC_INLINE_OPS ? ovfbuf.data() + dr.ovf : dr.ops, sc); }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  out[1] = ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / reps / n;
}

extern "C" double emul_time_rebuild(const dpc_problem_t *problems, int n, int reps, int64_t *npairs_out) {
  dpc::Batch b;
  std::vector<dpc_result_t> results((size_t)n);
  if (b.add_ext(problems, results.data(), n) < 0) return -1.0;
  std::vector<DevRes> dres(b.dprobs.size());
  std::vector<uint16_t> ovfbuf(1 << 22);
  unsigned int used = 0;
  OvfArena ovf; ovf.ops = ovfbuf.data(); ovf.used = &used; ovf.cap = (unsigned int)ovfbuf.size();
  Lanes ln; ln.lane = 0; ln.n = 1;
  RowFill rfill;
  std::vector<uint8_t> arena;
  std::vector<uint8_t> gout((size_t)b.gout_total + 64, 0xEE);
  b.pool_align(16);
  for (size_t k = 0; k < b.dprobs.size(); k++) {
    ArenaLayout a;
    dpc_layout(b.dprobs[k], a, 2);
    arena.assign(a.total + 64, 0);
    uint8_t *base = arena.data() + ((16 - ((uintptr_t)arena.data() & 15)) & 15);
    memset(&dres[k], 0, sizeof(DevRes));
    dpc_solve_problem<RowFill, -1, -1>(b.dprobs[k], b.pool.data(), dpc::G().setup.genome_blocks, &dpc::G().tables, base, a.total, base, &dres[k], ovf, gout.data(), rfill, ln);
  }
  b.gout_host = getenv("EMUL_NO_GOUT") ? NULL : gout.data();
  dpc::Scratch sc;
  int64_t total = 0;
  for (int i = 0; i < n; i++) if (b.probs[i].dev >= 0) { const DevRes &dr = dres[b.probs[i].dev]; b.finalize(i, dr, dr.nopsL + dr.nopsR > DPC_INLINE_OPS ? ovfbuf.data() + dr.ovf : dr.ops, sc); total += results[i].npairs; }
  std::vector<dpc_pair_t> out((size_t)total + 4096);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int rep = 0; rep < reps; rep++) {
    int64_t at = 0;
    for (int i = 0; i < n; i++) {
      if (b.probs[i].dev < 0) continue;
      if (i + 8 < n) b.prefetch_genome(i + 8);
      const DevRes &dr = dres[b.probs[i].dev];
      at += b.rebuild(i, dr, dr.nopsL + dr.nopsR > DPC_INLINE_OPS ? ovfbuf.data() + dr.ovf : dr.ops, out.data() + at, sc, getenv("EMUL_NOSTREAM") == NULL);
    }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (npairs_out) *npairs_out = total;
  return ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / reps / n;
}
