/* dpc_pipeline.cuh -- the device pipeline of dpc_solve (included by dynprog_cuda.cu).
 *
 * One chunk of the caller's problem array goes through
 *   H2D   the dpc_problem_t records as they are + the byte range of the query buffer they point into
 *   K1    dpc_prepare_kernel   checks, early returns, descriptors, launch class        (dpc_prepare_one)
 *   K2    dpc_sort_kernel      counting sort of the problems by (class, work bucket): the launch lists
 *   K3    dpc_solve_kernel<>   fill + bridge + traceback, as for the ticket API
 *   K4    dpc_finish_kernel    results as the reference reports them, pairs per problem  (dpc_finish_one)
 *   K5    dpc_scan_kernel      pair offsets (exclusive prefix sum in input order)
 *   K6    dpc_expand_kernel    traceback ops -> dpc_pair_t records, one warp per problem (dpc_expand_one)
 *   D2H   results and pair records straight into the caller's arrays
 * so the host does no per-problem work (except the MaxEnt hook for final genome gaps).  Chunks run concurrently on
 * one stream per host thread; the only ordered step is the hand-over of the running pair offset from chunk to chunk.
 */
#ifndef DPC_PIPELINE_CUH
#define DPC_PIPELINE_CUH

#include "dpc_pipe.h"

#define PIPE_NBIN (NCLASS * DPC_NKG * DPC_NBUCKET)
struct PipeCounters {
  unsigned int bin[PIPE_NBIN];            /* histogram of (class, kind group, bucket) */
  unsigned int class_count[NCLASS * DPC_NKG], class_off[NCLASS * DPC_NKG];
  unsigned int work[NCLASS * DPC_NKG];          /* work-claim counters of the solve launches */
  unsigned long long scratch_total, ovf_total;
  long long pair_total;
  unsigned int gout_total, npatch, ovf_used, ndevice;
  int err;
  int pad;
};
#define PIPE_NOCLASS 0xffffffffu

__global__ void __launch_bounds__(256) dpc_prepare_kernel(const dpc_problem_t *hp, int n, const PrepEnv env, const uint8_t *qpool,
                                                          const uint32_t *qoff, DevProb *dprobs, dpc_result_t *res, uint32_t *bin_of, uint32_t *goff, PipeCounters *pc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const dpc_problem_t p = hp[i];
  DevProb d;
  dpc_result_t r;
  PrepOut o;
  const PrepEnv &e = env;
  int rc;
  if (qoff) {
    /* scattered queries: the host gathered them into a pool (env.qbase == 0) where problem i's span starts at
       qoff[i]; rebase the pointers so that the span arithmetic of dpc_prepare_one applies */
    dpc_problem_t q = p;
    const uint64_t first = (p.kind == DPC_END5_GAP) ? (uint64_t)(uintptr_t)p.seq1 - (uint64_t)(p.length1 > 0 ? (p.length1 - 1) : 0) : (uint64_t)(uintptr_t)p.seq1;
    const int64_t delta = (int64_t)qoff[i] - (int64_t)first;
    q.seq1 = (const char *)(uintptr_t)((uint64_t)(uintptr_t)p.seq1 + (uint64_t)delta);
    if (p.seq1R) q.seq1R = (const char *)(uintptr_t)((uint64_t)(uintptr_t)p.seq1R + (uint64_t)delta);
    rc = dpc_prepare_one(q, e, qpool, d, r, o);
  } else {
    rc = dpc_prepare_one(p, e, qpool, d, r, o);
  }
  uint32_t bin = PIPE_NOCLASS;
  if (rc < 0) atomicMin(&pc->err, rc);
  else if (rc == DPC_PREP_DEVICE) {
    if (o.scratch) {
      const unsigned long long at = atomicAdd(&pc->scratch_total, (unsigned long long)((o.scratch + 15) & ~15ull));
      d.scratch_lo = (uint32_t)at; d.scratch_hi = (uint32_t)(at >> 32);
    }
    if (o.gout) d.gout = atomicAdd(&pc->gout_total, o.gout);
    if (o.ovf) atomicAdd(&pc->ovf_total, (unsigned long long)o.ovf);
    bin = (uint32_t)(o.cls * DPC_NBUCKET + o.bucket);
    atomicAdd(&pc->bin[bin], 1u);
    dprobs[i] = d;
  }
  bin_of[i] = bin;
  goff[i] = d.gout;              /* for chunks whose pairs the host rebuilds: where the staged genome characters are */
  res[i] = r;
}

/* single block: bin offsets, class extents, scatter into the launch lists */
__global__ void __launch_bounds__(1024) dpc_sort_kernel(const uint32_t *bin_of, int n, uint32_t *list, PipeCounters *pc) {
  __shared__ unsigned int cur[PIPE_NBIN];
  __shared__ unsigned int ctot[NCLASS * DPC_NKG];
  /* per class: exclusive scan of its buckets (one thread per class; 64 buckets each), then the class offsets */
  for (int b = threadIdx.x; b < PIPE_NBIN; b += blockDim.x) cur[b] = pc->bin[b];
  __syncthreads();
  if (threadIdx.x < NCLASS * DPC_NKG) {
    unsigned int at = 0;
    for (int q = 0; q < DPC_NBUCKET; q++) { const unsigned int v = cur[threadIdx.x * DPC_NBUCKET + q]; cur[threadIdx.x * DPC_NBUCKET + q] = at; at += v; }
    ctot[threadIdx.x] = at;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int at = 0;
    for (int k = 0; k < NCLASS * DPC_NKG; k++) { pc->class_off[k] = at; pc->class_count[k] = ctot[k]; const unsigned int v = ctot[k]; ctot[k] = at; at += v; }
    pc->ndevice = at;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < PIPE_NBIN; b += blockDim.x) cur[b] += ctot[b / DPC_NBUCKET];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const uint32_t b = bin_of[i];
    if (b != PIPE_NOCLASS) list[atomicAdd(&cur[b], 1u)] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256) dpc_finish_kernel(const dpc_problem_t *hp, int n, const uint32_t *bin_of, const DevProb *dprobs,
                                                         const DevRes *dres, const uint16_t *ovf, const uint8_t *pool, const uint8_t *gout,
                                                         const uint32_t *blocks, const DevTables *tables, dpc_result_t *res,
                                                         uint32_t *patch, PipeCounters *pc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (bin_of[i] == PIPE_NOCLASS) return;        /* resolved by dpc_prepare_one: res[i] is final, no pairs */
  const DevRes &dr = dres[i];
  if (!(dr.status & DPC_ST_DONE) || (dr.status & DPC_ST_OVF_LOST)) { atomicMin(&pc->err, (int)DPC_ERR_CUDA); return; }
  const uint16_t *ops = (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? ovf + dr.ovf : dr.ops;
  Lanes one; one.lane = 0; one.n = 1;
  const int np = dpc_expand_one(hp[i], dprobs[i], dr, ops, pool, gout, blocks, tables, (dpc_pair_t *)0, one);
  dpc_result_t r = res[i];
  if (dpc_finish_one(hp[i], dr, np, r)) patch[atomicAdd(&pc->npatch, 1u)] = (uint32_t)i;
  res[i] = r;
}

/* single block: exclusive prefix sum of npairs in input order */
__global__ void __launch_bounds__(1024) dpc_scan_kernel(const dpc_result_t *res, int n, long long *off, PipeCounters *pc) {
  __shared__ long long part[1024];
  const int per = (n + 1023) / 1024, lo = threadIdx.x * per, hi = min(n, lo + per);
  long long s = 0;
  for (int i = lo; i < hi; i++) s += res[i].npairs;
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long at = 0;
    for (int t = 0; t < 1024; t++) { const long long v = part[t]; part[t] = at; at += v; }
    pc->pair_total = at;
    off[n] = at;
  }
  __syncthreads();
  long long at = part[threadIdx.x];
  for (int i = lo; i < hi; i++) { off[i] = at; at += res[i].npairs; }
}

__global__ void __launch_bounds__(256) dpc_expand_kernel(const dpc_problem_t *hp, int n, const uint32_t *bin_of, const DevProb *dprobs,
                                                         const DevRes *dres, const uint16_t *ovf, const uint8_t *pool, const uint8_t *gout,
                                                         const uint32_t *blocks, const DevTables *tables, const dpc_result_t *res,
                                                         const long long *off, dpc_pair_t *pairs) {
  __shared__ DevTables s_tables;
  {
    const uint32_t *src = (const uint32_t *)tables;
    uint32_t *dst = (uint32_t *)&s_tables;
    for (int k = threadIdx.x; k < (int)(sizeof(DevTables) / 4); k += blockDim.x) dst[k] = src[k];
  }
  __syncthreads();
  Lanes ln; ln.lane = threadIdx.x & 31; ln.n = 32;
  const int nwarp = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarp) {
    if (bin_of[i] == PIPE_NOCLASS || res[i].npairs == 0) continue;
    const DevRes &dr = dres[i];
    const uint16_t *ops = (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? ovf + dr.ovf : dr.ops;
    dpc_expand_one(hp[i], dprobs[i], dr, ops, pool, gout, blocks, &s_tables, pairs + off[i], ln);
  }
}

/* ---- one host thread's pipeline state --------------------------------------------------------------------- */
struct Pipe {
  int device;
  cudaStream_t stream;
  cudaEvent_t ev;
  bool live;
  DBuf<dpc_problem_t> d_hp;
  DBuf<uint8_t> d_q, d_scratch, d_gout;
  DBuf<uint32_t> d_qoff, d_bin, d_list, d_patch, d_goff;
  DBuf<DevProb> d_probs;
  DBuf<DevRes> d_dres;
  DBuf<dpc_result_t> d_res;
  DBuf<uint16_t> d_ovf;
  DBuf<long long> d_off;
  DBuf<dpc_pair_t> d_pairs;
  DBuf<PipeCounters> d_pc;
  PBuf<PipeCounters> h_pc;
  PBuf<long long> h_off;
  PBuf<uint32_t> h_patch, h_qoff;
  PBuf<uint8_t> h_q;
  PBuf<dpc_pair_t> h_pairs;
  PBuf<dpc_problem_t> h_hp;
  PBuf<dpc_result_t> h_res;
  PBuf<DevRes> h_dres;            /* host route: device records, staged genome characters, overflow ops */
  PBuf<uint32_t> h_goff;
  PBuf<uint8_t> h_gout;
  PBuf<uint16_t> h_ovf;
  size_t ovf_cap;
  dpc_result_t *res_dst;          /* where the staged results go when the caller's array is pageable */
  int64_t h2d_bytes, d2h_bytes;
  int nlaunch;

  Pipe() : device(0), stream(0), ev(0), live(false), h2d_bytes(0), d2h_bytes(0), nlaunch(0), res_dst(NULL), ovf_cap(0) {}
  int open(int dev) {
    device = dev;
    CK(cudaSetDevice(dev));
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    h_pc.set_alloc(pinned()); h_off.set_alloc(pinned()); h_patch.set_alloc(pinned()); h_qoff.set_alloc(pinned());
    h_q.set_alloc(pinned()); h_pairs.set_alloc(pinned()); h_hp.set_alloc(pinned()); h_res.set_alloc(pinned());
    h_dres.set_alloc(pinned()); h_goff.set_alloc(pinned()); h_gout.set_alloc(pinned()); h_ovf.set_alloc(pinned());
    live = true;
    return DPC_OK;
  }
  void close() {
    if (!live) return;
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
    d_hp.release(); d_q.release(); d_scratch.release(); d_gout.release(); d_qoff.release(); d_bin.release(); d_list.release();
    d_patch.release(); d_goff.release(); d_probs.release(); d_dres.release(); d_res.release(); d_ovf.release(); d_off.release(); d_pairs.release();
    d_pc.release();
    cudaEventDestroy(ev); cudaStreamDestroy(stream);
    live = false;
  }
  ~Pipe() { close(); }
};

static bool host_pinned(const void *p, size_t bytes) {
  if (!p || !bytes) return false;
  cudaPointerAttributes a, b;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaPointerGetAttributes(&b, (const char *)p + bytes - 1) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && b.type == cudaMemoryTypeHost;
}

/* what a chunk needs from the caller's problems before it can take the device pipeline */
struct ChunkScan { bool eligible; uint64_t qmin, qmax; uint64_t qsum; };
static ChunkScan scan_chunk(const dpc_problem_t *p, int n) {
  const Globals &g = G();
  ChunkScan s; s.eligible = g.setup.splice_known == NULL; s.qmin = ~0ull; s.qmax = 0; s.qsum = 0;
  for (int i = 0; i < n && s.eligible; i++) {
    const dpc_problem_t &q = p[i];
    if (q.kind < 0 || q.kind > DPC_END3_GAP || q.use_probabilities_p || (q.finalp && g.setup.splice_prob == NULL)) { s.eligible = false; break; }
    int len = q.kind == DPC_CDNA_GAP ? q.offset1R - q.offset1 + 1 : q.length1;
    if (len <= 0 || !q.seq1) continue;
    if (len > (1 << 24)) { s.eligible = false; break; }
    const uint64_t first = (uint64_t)(uintptr_t)q.seq1 - (q.kind == DPC_END5_GAP ? (uint64_t)(len - 1) : 0u);
    if (first < s.qmin) s.qmin = first;
    if (first + (uint64_t)len > s.qmax) s.qmax = first + (uint64_t)len;
    s.qsum += (uint64_t)len;
  }
  return s;
}

/* Runs problems[0..n) through the device pipeline on `pp`.  On return results[] are final (including the MaxEnt
 * probabilities of final genome gaps), h_off holds the chunk-local pair offsets, *total the number of pair records,
 * which wait in pp.d_pairs for copy_pairs().  want_pairs == false skips the expansion. */
static double pipe_now() {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

/* blocks of a solve launch (cached: the occupancy query is a driver call) */
static int solve_grid(const DeviceState &dv, int variant, int kg, bool gen, size_t smem, int nproblems) {
  static std::atomic<int> cache[MAXDEV][NVARIANT][DPC_NKG][2][NCLASS];
  int cls = 0;
  for (int k = 0; k < NCLASS; k++) if ((size_t)8 * k_class_bytes[k] == smem && k_class_variant[k] == variant) cls = k;
  const int dev = (int)(&dv - g_dev);
  int per_sm = cache[dev][variant][kg][gen ? 1 : 0][cls].load();
  if (per_sm <= 0) {
    per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kernel_of(variant, kg, gen), 256, smem) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    if (per_sm < 1) per_sm = 1;
    cache[dev][variant][kg][gen ? 1 : 0][cls].store(per_sm);
  }
  int grid = (nproblems + 7) / 8;
  if (grid > dv.sm_count * per_sm) grid = dv.sm_count * per_sm;
  return grid;
}

/* STAGE 1 (queued, not waited for): copy-in, prepare, sort, the counters back.
 * hp_pinned: the caller's problem array is page-locked (dpc_host_register), so the copy engine reads it in place;
 * pageable memory is staged through page-locked buffers with a memcpy (a cudaMemcpyAsync on pageable memory would
 * serialise inside the driver).  The same goes for the query bytes. */
static int pipe_stage1(Pipe &pp, const dpc_problem_t *problems, int n, const ChunkScan &cs, int fill_gen, bool hp_pinned) {
  const Globals &g = G();
  CK(cudaSetDevice(pp.device));
  pp.h2d_bytes = pp.d2h_bytes = 0; pp.nlaunch = 0;
  int rc;
  /* queries: one contiguous range of the caller's buffer when the problems point into one (the usual case: a
     read batch), else gathered here */
  uint64_t qbytes = cs.qmax > cs.qmin ? cs.qmax - cs.qmin : 0;
  const bool gather = qbytes > 4 * cs.qsum + (1u << 16) || qbytes >= 0xf0000000ull;
  PrepEnv env;
  env.maxlength1 = g.maxlength1; env.maxlength2 = g.maxlength2; env.genome_nbases = g.genome_nbases;
  env.novelsplicingp = g.setup.novelsplicingp; env.fillmode = fill_gen ? 1 : 2;
  for (int k = 0; k < NCLASS; k++) env.class_bytes[k] = k_class_bytes[k];
  if ((rc = pp.d_hp.need((size_t)n)) || (rc = pp.d_probs.need((size_t)n)) || (rc = pp.d_dres.need((size_t)n)) ||
      (rc = pp.d_res.need((size_t)n)) || (rc = pp.d_bin.need((size_t)n)) || (rc = pp.d_list.need((size_t)n)) ||
      (rc = pp.d_patch.need((size_t)n)) || (rc = pp.d_goff.need((size_t)n)) || (rc = pp.d_off.need((size_t)n + 1)) || (rc = pp.d_pc.need(1)))
    return rc;
  const uint32_t *d_qoff = NULL;
  if (gather) {
    if (cs.qsum >= 0xf0000000ull) return DPC_ERR_NOMEM;
    pp.h_q.clear(); pp.h_q.grow((size_t)cs.qsum + 16);
    pp.h_qoff.clear(); pp.h_qoff.grow((size_t)n);
    uint32_t at = 0;
    for (int i = 0; i < n; i++) {
      const dpc_problem_t &q = problems[i];
      const int len = q.kind == DPC_CDNA_GAP ? q.offset1R - q.offset1 + 1 : q.length1;
      pp.h_qoff[(size_t)i] = at;
      if (len <= 0 || !q.seq1) continue;
      memcpy(pp.h_q.data() + at, q.seq1 - (q.kind == DPC_END5_GAP ? len - 1 : 0), (size_t)len);
      at += (uint32_t)len;
    }
    qbytes = at;
    if ((rc = pp.d_q.need((size_t)qbytes + 16)) || (rc = pp.d_qoff.need((size_t)n))) return rc;
    CK(cudaMemcpyAsync(pp.d_q.p, pp.h_q.data(), (size_t)qbytes, cudaMemcpyHostToDevice, pp.stream));
    CK(cudaMemcpyAsync(pp.d_qoff.p, pp.h_qoff.data(), (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, pp.stream));
    d_qoff = pp.d_qoff.p;
    env.qbase = 0; env.qbytes = qbytes;
    pp.h2d_bytes += (int64_t)n * 4;
  } else {
    if ((rc = pp.d_q.need((size_t)qbytes + 16))) return rc;
    const void *src = (const void *)(uintptr_t)cs.qmin;
    if (qbytes && !host_pinned(src, (size_t)qbytes)) {
      pp.h_q.clear(); pp.h_q.grow((size_t)qbytes);
      memcpy(pp.h_q.data(), src, (size_t)qbytes);
      src = pp.h_q.data();
    }
    if (qbytes) CK(cudaMemcpyAsync(pp.d_q.p, src, (size_t)qbytes, cudaMemcpyHostToDevice, pp.stream));
    env.qbase = cs.qmin; env.qbytes = qbytes;
  }
  {
    const dpc_problem_t *src = problems;
    if (!hp_pinned) {
      pp.h_hp.clear(); pp.h_hp.grow((size_t)n);
      memcpy(pp.h_hp.data(), problems, (size_t)n * sizeof(dpc_problem_t));
      src = pp.h_hp.data();
    }
    CK(cudaMemcpyAsync(pp.d_hp.p, src, (size_t)n * sizeof(dpc_problem_t), cudaMemcpyHostToDevice, pp.stream));
  }
  pp.h2d_bytes += (int64_t)((size_t)n * sizeof(dpc_problem_t) + qbytes);
  CK(cudaMemsetAsync(pp.d_pc.p, 0, sizeof(PipeCounters), pp.stream));
  dpc_prepare_kernel<<<(n + 255) / 256, 256, 0, pp.stream>>>(pp.d_hp.p, n, env, pp.d_q.p, d_qoff, pp.d_probs.p, pp.d_res.p, pp.d_bin.p, pp.d_goff.p, pp.d_pc.p);
  dpc_sort_kernel<<<1, 1024, 0, pp.stream>>>(pp.d_bin.p, n, pp.d_list.p, pp.d_pc.p);
  CK(cudaGetLastError());
  pp.nlaunch += 2;
  pp.h_pc.clear(); pp.h_pc.grow(1);
  CK(cudaMemcpyAsync(pp.h_pc.data(), pp.d_pc.p, sizeof(PipeCounters), cudaMemcpyDeviceToHost, pp.stream));
  CK(cudaEventRecord(pp.ev, pp.stream));
  return DPC_OK;
}

/* STAGE 2 (after stage 1 is back): the solve kernels, one launch per non-empty class -- the same kernels as the
 * ticket API -- then results and pair offsets */
static int pipe_stage2(Pipe &pp, int n, int fill_gen) {
  DeviceState &dv = g_dev[pp.device];
  int rc;
  const PipeCounters &pc = pp.h_pc[0];
  if (pc.err < 0) return pc.err;
  if (pc.scratch_total > SCRATCH_BUDGET) return DPC_ERR_NOMEM;
  const size_t ovf_cap = (size_t)(pc.ovf_total > (1ull << 30) ? (1ull << 30) : pc.ovf_total) + 64;
  pp.ovf_cap = ovf_cap;
  if ((rc = pp.d_scratch.need((size_t)pc.scratch_total + 16)) || (rc = pp.d_gout.need((size_t)pc.gout_total + 64)) ||
      (rc = pp.d_ovf.need(ovf_cap)))
    return rc;
  for (int k = 0; k < NCLASS * DPC_NKG; k++) {
    if (!pc.class_count[k]) continue;
    KernelArgs a;
    a.probs = pp.d_probs.p; a.list = pp.d_list.p + pc.class_off[k]; a.n = (int)pc.class_count[k];
    a.pool = pp.d_q.p; a.blocks = dv.d_blocks; a.tables = dv.d_tables; a.res = pp.d_dres.p;
    a.ovf.ops = pp.d_ovf.p; a.ovf.used = &pp.d_pc.p->ovf_used; a.ovf.cap = (unsigned int)ovf_cap;
    a.scratch = pp.d_scratch.p; a.gout = pp.d_gout.p; a.arena_bytes = k_class_bytes[k / DPC_NKG]; a.counter = &pp.d_pc.p->work[k];
    a.claim = claim_of(pc.bin + k * DPC_NBUCKET, DPC_NBUCKET);
    const int variant = k_class_variant[k / DPC_NKG];
    const size_t smem = variant != V_HBM ? (size_t)8 * a.arena_bytes : 0;
    const int grid = solve_grid(dv, variant, k % DPC_NKG, fill_gen != 0, smem, a.n);
    kernel_of(variant, k % DPC_NKG, fill_gen != 0, a.claim)<<<grid, 256, smem, pp.stream>>>(a);
    pp.nlaunch++;
  }
  dpc_finish_kernel<<<(n + 255) / 256, 256, 0, pp.stream>>>(pp.d_hp.p, n, pp.d_bin.p, pp.d_probs.p, pp.d_dres.p, pp.d_ovf.p, pp.d_q.p, pp.d_gout.p,
                                                          dv.d_blocks, dv.d_tables, pp.d_res.p, pp.d_patch.p, pp.d_pc.p);
  dpc_scan_kernel<<<1, 1024, 0, pp.stream>>>(pp.d_res.p, n, pp.d_off.p, pp.d_pc.p);
  CK(cudaGetLastError());
  pp.nlaunch += 2;
  CK(cudaMemcpyAsync(pp.h_pc.data(), pp.d_pc.p, sizeof(PipeCounters), cudaMemcpyDeviceToHost, pp.stream));
  pp.h_off.clear(); pp.h_off.grow((size_t)n + 1);
  CK(cudaMemcpyAsync(pp.h_off.data(), pp.d_off.p, ((size_t)n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, pp.stream));
  pp.h_patch.clear(); pp.h_patch.grow((size_t)n);
  CK(cudaEventRecord(pp.ev, pp.stream));
  return DPC_OK;
}

/* results into the caller's array (staged when it is pageable) + the indices that still need the MaxEnt hook */
static int pipe_queue_results(Pipe &pp, int n, dpc_result_t *results, bool out_pinned) {
  const PipeCounters &pc = pp.h_pc[0];
  pp.res_dst = NULL;
  if (out_pinned) {
    CK(cudaMemcpyAsync(results, pp.d_res.p, (size_t)n * sizeof(dpc_result_t), cudaMemcpyDeviceToHost, pp.stream));
  } else {
    pp.h_res.clear(); pp.h_res.grow((size_t)n);
    CK(cudaMemcpyAsync(pp.h_res.data(), pp.d_res.p, (size_t)n * sizeof(dpc_result_t), cudaMemcpyDeviceToHost, pp.stream));
    pp.res_dst = results;
  }
  if (pc.npatch) CK(cudaMemcpyAsync(pp.h_patch.data(), pp.d_patch.p, (size_t)pc.npatch * sizeof(uint32_t), cudaMemcpyDeviceToHost, pp.stream));
  pp.d2h_bytes += (int64_t)((size_t)n * sizeof(dpc_result_t) + ((size_t)n + 1) * 8 + 2 * sizeof(PipeCounters) + (size_t)pc.npatch * 4);
  return DPC_OK;
}
/* host work after the results are back: the staged copy, get_splicesite_probs (4104-4108) through the MaxEnt hook */
static void pipe_apply_hooks(Pipe &pp, const dpc_problem_t *problems, int n, dpc_result_t *results) {
  if (pp.res_dst) memcpy(pp.res_dst, pp.h_res.data(), (size_t)n * sizeof(dpc_result_t));
  const PipeCounters &pc = pp.h_pc[0];
  for (unsigned int k = 0; k < pc.npatch; k++) {
    const uint32_t i = pp.h_patch[k];
    dpc_result_t &r = results[i];
    const int cL = (int)r.left_prob, cR = (int)r.right_prob;
    r.left_prob = Batch::site_prob(problems[i], true, cL, false);
    r.right_prob = Batch::site_prob(problems[i], false, cR, false);
  }
}

/* DEVICE ROUTE, queued when the chunk's pair count is known: expand on the device, results on their way */
static int pipe_device_route_start(Pipe &pp, int n, dpc_result_t *results, bool out_pinned, bool want_pairs, int64_t total) {
  DeviceState &dv = g_dev[pp.device];
  int rc;
  if (want_pairs && total > 0) {
    if ((rc = pp.d_pairs.need((size_t)total))) return rc;
    int grid = (n + 7) / 8;
    if (grid > dv.sm_count * 8) grid = dv.sm_count * 8;
    dpc_expand_kernel<<<grid, 256, 0, pp.stream>>>(pp.d_hp.p, n, pp.d_bin.p, pp.d_probs.p, pp.d_dres.p, pp.d_ovf.p, pp.d_q.p, pp.d_gout.p,
                                                  dv.d_blocks, dv.d_tables, pp.d_res.p, pp.d_off.p, pp.d_pairs.p);
    CK(cudaGetLastError());
    pp.nlaunch++;
  }
  return pipe_queue_results(pp, n, results, out_pinned);
}
/* ... and when the caller's pair block is known: the records go straight into it (through the bounce buffer when
 * it is pageable) */
static int pipe_device_route_copy(Pipe &pp, dpc_pair_t *pairs_at, bool direct, int64_t total) {
  if (pairs_at && total > 0) {
    if (direct) {
      CK(cudaMemcpyAsync(pairs_at, pp.d_pairs.p, (size_t)total * sizeof(dpc_pair_t), cudaMemcpyDeviceToHost, pp.stream));
    } else {
      pp.h_pairs.clear(); pp.h_pairs.grow((size_t)total);
      CK(cudaMemcpyAsync(pp.h_pairs.data(), pp.d_pairs.p, (size_t)total * sizeof(dpc_pair_t), cudaMemcpyDeviceToHost, pp.stream));
    }
    pp.d2h_bytes += total * (int64_t)sizeof(dpc_pair_t);
  }
  CK(cudaEventRecord(pp.ev, pp.stream));
  return DPC_OK;
}
static void pipe_device_route_final(Pipe &pp, const dpc_problem_t *problems, int n, dpc_result_t *results, dpc_pair_t *pairs_at,
                                    bool direct, int64_t total) {
  if (pairs_at && total > 0 && !direct) memcpy(pairs_at, pp.h_pairs.data(), (size_t)total * sizeof(dpc_pair_t));
  pipe_apply_hooks(pp, problems, n, results);
}

/* HOST ROUTE (the link is the bottleneck and a host thread has time): the compact device records come back --
 * 128 B of result + ops and the staged genome characters per problem instead of 16 B per Pair -- and a host
 * thread expands them with the host half's rebuild.  Same records either way. */
static int pipe_host_route_start(Pipe &pp, int n, dpc_result_t *results, bool out_pinned) {
  const PipeCounters &pc = pp.h_pc[0];
  int rc;
  pp.h_dres.clear(); pp.h_dres.grow((size_t)n);
  pp.h_goff.clear(); pp.h_goff.grow((size_t)n);
  pp.h_gout.clear(); pp.h_gout.grow((size_t)pc.gout_total + 64);
  CK(cudaMemcpyAsync(pp.h_dres.data(), pp.d_dres.p, (size_t)n * sizeof(DevRes), cudaMemcpyDeviceToHost, pp.stream));
  CK(cudaMemcpyAsync(pp.h_goff.data(), pp.d_goff.p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, pp.stream));
  if (pc.gout_total) CK(cudaMemcpyAsync(pp.h_gout.data(), pp.d_gout.p, (size_t)pc.gout_total, cudaMemcpyDeviceToHost, pp.stream));
  pp.h_ovf.clear();
  if (pc.ovf_used) {
    if (pc.ovf_used > pp.ovf_cap) return DPC_ERR_NOMEM;
    pp.h_ovf.grow((size_t)pc.ovf_used);
    CK(cudaMemcpyAsync(pp.h_ovf.data(), pp.d_ovf.p, (size_t)pc.ovf_used * sizeof(uint16_t), cudaMemcpyDeviceToHost, pp.stream));
  }
  pp.d2h_bytes += (int64_t)((size_t)n * (sizeof(DevRes) + 4) + pc.gout_total + (size_t)pc.ovf_used * 2);
  if ((rc = pipe_queue_results(pp, n, results, out_pinned)) != DPC_OK) return rc;
  CK(cudaEventRecord(pp.ev, pp.stream));
  return DPC_OK;
}
static int pipe_host_route_final(Pipe &pp, const dpc_problem_t *problems, int n, dpc_result_t *results, dpc_pair_t *pairs_at,
                                 int64_t total, bool stream_stores, Scratch &scratch) {
  pipe_apply_hooks(pp, problems, n, results);
  int64_t at = 0;
  for (int i = 0; i < n; i++) {
    const int np = results[i].npairs;
    if (np == 0) continue;
    const dpc_problem_t &p = problems[i];
    const DevRes &dr = pp.h_dres[(size_t)i];
    const uint16_t *ops = (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? pp.h_ovf.data() + dr.ovf : dr.ops;
    int L1c, L2c;
    Batch::clipped_lengths(p, &L1c, &L2c);
    const char *q = p.kind == DPC_END5_GAP ? p.seq1 - (L1c - 1) : p.seq1;
    const uint32_t go = pp.h_goff[(size_t)i];
    const char *staged = go != DPC_NO_GOUT ? (const char *)pp.h_gout.data() + go : NULL;
    const int k = Batch::rebuild_core(p, q, NULL, L1c, L2c, staged, dr, ops, pairs_at + at, scratch, stream_stores);
    if (k != np) return DPC_ERR_STATE;
    at += k;
  }
  if (at != total) return DPC_ERR_STATE;
  return DPC_OK;
}

#endif /* DPC_PIPELINE_CUH */
