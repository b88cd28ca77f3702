/* dynprog_cuda.cu -- libdynprog_cuda: C ABI (include/dynprog_cuda.h) + sm_100a kernels.
 *
 * Replaces the five gap-fill solvers of the reference's src/dynprog.c (Dynprog_single_gap 4450,
 * Dynprog_cdna_gap 4577, Dynprog_genome_gap 4798, Dynprog_end5_gap 5094, Dynprog_end3_gap 5556)
 * with batched device work: one warp per problem, matrices and direction nibbles in shared
 * memory (HBM scratch only for problems that do not fit), run-length traceback ops back to the
 * host.  There is no CPU solver in this library: without a usable device every entry point that
 * needs one returns DPC_ERR_CUDA.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <mutex>
#include <vector>

#include "dpc_host.h"
#include "dpc_fill_warp.cuh"

using namespace dpc;

/* ---- kernels ------------------------------------------------------------------------------ */
struct KernelArgs {
  const DevProb *probs;
  const uint32_t *list;        /* problem indices of this launch */
  int n;
  const uint8_t *pool;
  const uint32_t *blocks;
  const DevTables *tables;
  DevRes *res;
  OvfArena ovf;
  uint8_t *scratch;            /* HBM arenas (SMEM == false) */
  uint32_t arena_bytes;        /* per-warp shared-memory arena (SMEM == true) */
  unsigned int *counter;       /* dynamic work distribution */
  int force_generic;
};

template <bool SMEM>
__global__ void __launch_bounds__(256) dpc_solve_kernel(const KernelArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ DevTables s_tables;
  {
    const uint32_t *src = (const uint32_t *)a.tables;
    uint32_t *dst = (uint32_t *)&s_tables;
    for (int i = threadIdx.x; i < (int)(sizeof(DevTables) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  Lanes ln; ln.lane = threadIdx.x & 31; ln.n = 32;
  for (;;) {
    int i = 0;
    if (ln.lane == 0) i = (int)atomicAdd(a.counter, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= a.n) break;
    const uint32_t pi = a.list[i];
    const DevProb p = a.probs[pi];
    uint8_t *arena = SMEM ? smem + (size_t)warp * a.arena_bytes
                          : a.scratch + (((uint64_t)p.scratch_hi << 32) | p.scratch_lo);
    if (a.force_generic) {
      GenericFill fill;
      dpc_solve_problem(p, a.pool, a.blocks, &s_tables, arena, &a.res[pi], a.ovf, fill, ln);
    } else {
      WarpFill fill;
      dpc_solve_problem(p, a.pool, a.blocks, &s_tables, arena, &a.res[pi], a.ovf, fill, ln);
    }
    __syncwarp();
  }
}

/* ---- process-wide device state ------------------------------------------------------------ */
#define MAXDEV 16
struct DeviceState {
  bool ready;
  uint64_t version;
  uint32_t *d_blocks;
  DevTables *d_tables;
  int sm_count;
  int max_smem;
};
static DeviceState g_dev[MAXDEV];
static std::mutex g_mu;
static uint64_t g_version = 0;
static int g_force_generic = 0;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
  fprintf(stderr, "libdynprog_cuda: %s failed at %s:%d: %s\n", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
  return DPC_ERR_CUDA; } } while (0)

static int ensure_device(int dev) {
  std::lock_guard<std::mutex> lock(g_mu);
  Globals &g = G();
  if (dev < 0 || dev >= MAXDEV) return DPC_ERR_ARG;
  if (!g.inited || !g.setup_done) return DPC_ERR_STATE;
  DeviceState &d = g_dev[dev];
  CK(cudaSetDevice(dev));
  if (d.ready && d.version == g_version) return DPC_OK;
  if (d.ready) { cudaFree(d.d_blocks); cudaFree(d.d_tables); d.ready = false; }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) {
    fprintf(stderr, "libdynprog_cuda: device %d is sm_%d%d; this library is built for sm_100a only\n", dev, prop.major, prop.minor);
    return DPC_ERR_CUDA;
  }
  d.sm_count = prop.multiProcessorCount;
  d.max_smem = (int)prop.sharedMemPerBlockOptin;
  CK(cudaMalloc(&d.d_blocks, g.setup.genome_nwords * sizeof(uint32_t)));
  CK(cudaMemcpy(d.d_blocks, g.setup.genome_blocks, g.setup.genome_nwords * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d.d_tables, sizeof(DevTables)));
  CK(cudaMemcpy(d.d_tables, &g.tables, sizeof(DevTables), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(dpc_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.max_smem - (int)sizeof(DevTables) - 1024));
  d.version = g_version;
  d.ready = true;
  return DPC_OK;
}

/* ---- context -------------------------------------------------------------------------------- */
template <class T> struct DBuf {
  T *p; size_t cap;
  DBuf() : p(NULL), cap(0) {}
  int need(size_t n) {
    if (n <= cap) return DPC_OK;
    if (p) cudaFree(p);
    p = NULL; cap = 0;
    size_t want = n + n / 4 + 1024;
    if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return DPC_ERR_NOMEM; }
    cap = want;
    return DPC_OK;
  }
  void release() { if (p) cudaFree(p); p = NULL; cap = 0; }
};
template <class T> struct HBuf {   /* pinned */
  T *p; size_t cap;
  HBuf() : p(NULL), cap(0) {}
  int need(size_t n) {
    if (n <= cap) return DPC_OK;
    if (p) cudaFreeHost(p);
    p = NULL; cap = 0;
    size_t want = n + n / 4 + 1024;
    if (cudaMallocHost(&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return DPC_ERR_NOMEM; }
    cap = want;
    return DPC_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = NULL; cap = 0; }
};

struct ClassLaunch {
  bool smem;
  uint32_t arena_bytes;
  int wpb;
  size_t list_off;
  int n;
};

#define NCLASS 8
static const uint32_t k_class_bytes[NCLASS] = { 3 << 10, 6 << 10, 12 << 10, 24 << 10, 48 << 10, 96 << 10, 192 << 10, 0 };
#define SCRATCH_BUDGET (6ull << 30)

struct dpc_ctx {
  int device;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  Batch batch;
  DBuf<DevProb> d_probs;
  DBuf<uint8_t> d_pool, d_scratch;
  DBuf<DevRes> d_res;
  DBuf<uint32_t> d_list;
  DBuf<uint16_t> d_ovf;
  DBuf<unsigned int> d_counters;
  HBuf<DevRes> h_res;
  HBuf<uint16_t> h_ovf;
  HBuf<unsigned int> h_counters;
  std::vector<uint32_t> list;
  std::vector<ClassLaunch> launches;
  size_t ovf_cap;
  bool flushed, waited;
  int nlaunch;
  float ms_total;
  int64_t h2d_bytes, d2h_bytes;
  int err;
};

static int launch_all(dpc_ctx *c) {
  DeviceState &d = g_dev[c->device];
  const size_t ncnt = c->launches.size() + 1;
  CK(cudaMemsetAsync(c->d_counters.p, 0, ncnt * sizeof(unsigned int), c->stream));
  CK(cudaEventRecord(c->ev0, c->stream));
  c->nlaunch = 0;
  for (size_t k = 0; k < c->launches.size(); k++) {
    const ClassLaunch &L = c->launches[k];
    KernelArgs a;
    a.probs = c->d_probs.p; a.list = c->d_list.p + L.list_off; a.n = L.n;
    a.pool = c->d_pool.p; a.blocks = d.d_blocks; a.tables = d.d_tables; a.res = c->d_res.p;
    a.ovf.ops = c->d_ovf.p; a.ovf.used = c->d_counters.p; a.ovf.cap = (unsigned int)c->ovf_cap;
    a.scratch = c->d_scratch.p; a.arena_bytes = L.arena_bytes; a.counter = c->d_counters.p + 1 + k;
    a.force_generic = g_force_generic;
    const int threads = L.wpb * 32;
    const size_t smem = L.smem ? (size_t)L.wpb * L.arena_bytes : 0;
    int per_sm = 1;
    if (L.smem) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dpc_solve_kernel<true>, threads, smem)); }
    else { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dpc_solve_kernel<false>, threads, 0)); }
    if (per_sm < 1) per_sm = 1;
    int grid = (L.n + L.wpb - 1) / L.wpb;
    if (grid > d.sm_count * per_sm) grid = d.sm_count * per_sm;
    if (L.smem) dpc_solve_kernel<true><<<grid, threads, smem, c->stream>>>(a);
    else dpc_solve_kernel<false><<<grid, threads, 0, c->stream>>>(a);
    CK(cudaGetLastError());
    c->nlaunch++;
  }
  CK(cudaEventRecord(c->ev1, c->stream));
  return DPC_OK;
}

/* ---- C ABI ---------------------------------------------------------------------------------- */
extern "C" {

int dpc_init(int maxlookback, int extraquerygap, int maxpeelback, int extramaterial_end, int extramaterial_paired, int mode) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = host_init(maxlookback, extraquerygap, maxpeelback, extramaterial_end, extramaterial_paired, mode);
  g_version++;
  const char *e = getenv("DPC_FORCE_GENERIC_FILL");
  g_force_generic = (e && *e && *e != '0') ? 1 : 0;
  return rc;
}

int dpc_setup(const dpc_setup_t *setup) {
  std::lock_guard<std::mutex> lock(g_mu);
  Globals &g = G();
  if (!g.inited) return DPC_ERR_STATE;
  if (!setup || !setup->genome_blocks || setup->genome_nwords < 3) return DPC_ERR_ARG;
  g.setup = *setup;
  g.genome_nbases = setup->genome_nwords / 3 * 32;
  g.setup_done = true;
  g_version++;
  return DPC_OK;
}

void dpc_term(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  for (int i = 0; i < MAXDEV; i++)
    if (g_dev[i].ready) {
      cudaSetDevice(i);
      cudaFree(g_dev[i].d_blocks); cudaFree(g_dev[i].d_tables);
      g_dev[i].ready = false;
    }
  G().setup_done = false;
}

int dpc_set_fill(int force_generic) { g_force_generic = force_generic ? 1 : 0; return DPC_OK; }

int dpc_pairdistance(int mismatchtype, int c1, int c2) { return G().P[mismatchtype & 3][c1 & 127][c2 & 127]; }
void dpc_maxlengths(int *maxlength1, int *maxlength2) { *maxlength1 = G().maxlength1; *maxlength2 = G().maxlength2; }
const char *dpc_strerror(int code) { return strerror_(code); }

int dpc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

dpc_ctx_t *dpc_ctx_new(int device) {
  if (device < 0 || device >= dpc_device_count()) return NULL;
  if (ensure_device(device) != DPC_OK) return NULL;
  dpc_ctx *c = new dpc_ctx();
  c->device = device; c->flushed = c->waited = false; c->err = 0; c->nlaunch = 0; c->ms_total = 0;
  c->ovf_cap = 0; c->h2d_bytes = c->d2h_bytes = 0;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) { delete c; return NULL; }
  return c;
}

void dpc_ctx_free(dpc_ctx_t *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->d_probs.release(); c->d_pool.release(); c->d_scratch.release(); c->d_res.release(); c->d_list.release();
  c->d_ovf.release(); c->d_counters.release(); c->h_res.release(); c->h_ovf.release(); c->h_counters.release();
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaStreamDestroy(c->stream);
  delete c;
}

int dpc_add(dpc_ctx_t *c, const dpc_problem_t *problem) {
  if (!c || !problem) return DPC_ERR_ARG;
  if (c->flushed) return DPC_ERR_STATE;
  return c->batch.add(*problem);
}

int dpc_add_bulk(dpc_ctx_t *c, const dpc_problem_t *problems, int n) {
  if (!c || (!problems && n > 0) || n < 0) return DPC_ERR_ARG;
  if (c->flushed) return DPC_ERR_STATE;
  int first = (int)c->batch.probs.size();
  c->batch.probs.reserve(c->batch.probs.size() + (size_t)n);
  c->batch.dprobs.reserve(c->batch.dprobs.size() + (size_t)n);
  for (int i = 0; i < n; i++) {
    int t = c->batch.add(problems[i]);
    if (t < 0) return t;
  }
  return first;
}

int dpc_reset(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  cudaSetDevice(c->device);
  if (c->flushed && !c->waited) cudaStreamSynchronize(c->stream);
  c->batch.clear();
  c->flushed = c->waited = false;
  return DPC_OK;
}

int dpc_flush(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  if (c->flushed) return DPC_ERR_STATE;
  int rc = ensure_device(c->device);
  if (rc != DPC_OK) return rc;
  DeviceState &d = g_dev[c->device];
  Batch &b = c->batch;
  const size_t n = b.dprobs.size();
  c->flushed = true; c->waited = false;
  c->launches.clear();
  c->h2d_bytes = c->d2h_bytes = 0;
  if (n == 0) return DPC_OK;

  /* bin the problems by arena size; oversize ones get HBM scratch */
  const int with_state = g_force_generic ? 1 : 2;   /* must match FILL::needs_state of the kernel */
  const uint32_t smem_limit = (uint32_t)(d.max_smem - (int)sizeof(DevTables) - 2048);
  std::vector<uint8_t> cls(n);
  size_t count[NCLASS] = { 0 };
  uint64_t scratch_total = 0, ovf_worst = 0;
  for (size_t i = 0; i < n; i++) {
    DevProb &p = b.dprobs[i];
    int k = NCLASS - 1;
    if (!((p.kind == DPC_END5_GAP || p.kind == DPC_END3_GAP) && p.endalign == DPC_QUERYEND_NOGAPS)) {
      ArenaLayout a;
      dpc_layout(p, a, with_state);
      for (k = 0; k < NCLASS - 1; k++) if (a.total <= k_class_bytes[k] && k_class_bytes[k] <= smem_limit) break;
      if (k == NCLASS - 1) {
        if (scratch_total + a.total > SCRATCH_BUDGET) return DPC_ERR_NOMEM;
        p.scratch_lo = (uint32_t)scratch_total; p.scratch_hi = (uint32_t)(scratch_total >> 32);
        scratch_total += a.total;
      }
      uint64_t worst = 0;
      for (int m = 0; m < a.nmat; m++) worst += (uint64_t)(a.d[m].rows + a.d[m].cols + 2);
      if (worst > DPC_INLINE_OPS) ovf_worst += worst;
    } else k = 0;
    cls[i] = (uint8_t)k;
    count[k]++;
  }
  c->list.resize(n);
  size_t off[NCLASS], at = 0;
  for (int k = 0; k < NCLASS; k++) { off[k] = at; at += count[k]; }
  {
    size_t cur[NCLASS];
    for (int k = 0; k < NCLASS; k++) cur[k] = off[k];
    for (size_t i = 0; i < n; i++) c->list[cur[cls[i]]++] = (uint32_t)i;
  }
  for (int k = 0; k < NCLASS; k++) {
    if (!count[k]) continue;
    ClassLaunch L;
    L.smem = k < NCLASS - 1;
    L.arena_bytes = k_class_bytes[k];
    L.wpb = 8;
    if (L.smem) { while (L.wpb > 1 && (uint64_t)L.wpb * L.arena_bytes > smem_limit) L.wpb >>= 1; }
    else L.wpb = 4;
    L.list_off = off[k]; L.n = (int)count[k];
    c->launches.push_back(L);
  }
  /* ops overflow arena: problems whose worst case exceeds the inline slots (bounded) */
  if (ovf_worst > (1ull << 30)) ovf_worst = 1ull << 30;
  c->ovf_cap = (size_t)ovf_worst + 64;

  b.pool_align(16);
  if ((rc = c->d_probs.need(n)) || (rc = c->d_pool.need(b.pool.size())) || (rc = c->d_res.need(n)) ||
      (rc = c->d_list.need(n)) || (rc = c->d_ovf.need(c->ovf_cap)) || (rc = c->d_counters.need(NCLASS + 2)) ||
      (rc = c->d_scratch.need((size_t)scratch_total + 16)) || (rc = c->h_res.need(n)) || (rc = c->h_counters.need(NCLASS + 2)))
    return rc;
  CK(cudaMemcpyAsync(c->d_probs.p, b.dprobs.data(), n * sizeof(DevProb), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_pool.p, b.pool.data(), b.pool.size(), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_list.p, c->list.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  c->h2d_bytes = (int64_t)(n * sizeof(DevProb) + b.pool.size() + n * sizeof(uint32_t));
  if ((rc = launch_all(c)) != DPC_OK) return rc;
  CK(cudaMemcpyAsync(c->h_res.p, c->d_res.p, n * sizeof(DevRes), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(c->h_counters.p, c->d_counters.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
  c->d2h_bytes = (int64_t)(n * sizeof(DevRes) + sizeof(unsigned int));
  return DPC_OK;
}

int dpc_wait(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  if (!c->flushed) return DPC_ERR_STATE;
  if (c->waited) return DPC_OK;
  Batch &b = c->batch;
  const size_t n = b.dprobs.size();
  if (n > 0) {
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventElapsedTime(&c->ms_total, c->ev0, c->ev1));
    unsigned int used = c->h_counters.p[0];
    if (used > 0) {
      if (used > c->ovf_cap) return DPC_ERR_NOMEM;
      int rc = c->h_ovf.need(used);
      if (rc) return rc;
      CK(cudaMemcpy(c->h_ovf.p, c->d_ovf.p, used * sizeof(uint16_t), cudaMemcpyDeviceToHost));
      c->d2h_bytes += (int64_t)used * 2;
    }
    for (size_t k = 0; k < n; k++) {
      const DevRes &dr = c->h_res.p[k];
      if (!(dr.status & DPC_ST_DONE) || (dr.status & DPC_ST_OVF_LOST)) return DPC_ERR_CUDA;
      const uint16_t *ops = (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? c->h_ovf.p + dr.ovf : dr.ops;
      b.finalize((int)b.dev2host[k], dr, ops);
    }
  }
  c->waited = true;
  return DPC_OK;
}

int dpc_result(dpc_ctx_t *c, int ticket, dpc_result_t *out) {
  if (!c || !out) return DPC_ERR_ARG;
  if (!c->waited || ticket < 0 || ticket >= (int)c->batch.probs.size()) return DPC_ERR_STATE;
  *out = c->batch.probs[ticket].res;
  return DPC_OK;
}

static int pairs_of(dpc_ctx *c, int ticket, Batch::Stack &st) {
  const HostProb &h = c->batch.probs[ticket];
  st.clear();
  if (h.dev < 0) return 0;
  const DevRes &dr = c->h_res.p[h.dev];
  const uint16_t *ops = (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? c->h_ovf.p + dr.ovf : dr.ops;
  c->batch.rebuild(ticket, dr, ops, st);
  return (int)st.size();
}

int dpc_pairs(dpc_ctx_t *c, int ticket, dpc_pair_t *out, int cap) {
  if (!c) return DPC_ERR_ARG;
  if (!c->waited || ticket < 0 || ticket >= (int)c->batch.probs.size()) return DPC_ERR_STATE;
  Batch::Stack st;
  int n = pairs_of(c, ticket, st);
  if (n > cap || (n > 0 && !out)) return DPC_ERR_ARG;
  if (n) memcpy(out, st.data(), (size_t)n * sizeof(dpc_pair_t));
  return n;
}

int dpc_solve(dpc_ctx_t *c, const dpc_problem_t *problems, int n, dpc_result_t *results,
              dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off) {
  int rc;
  if (!c || n < 0 || (n > 0 && (!problems || !results))) return DPC_ERR_ARG;
  if ((rc = dpc_reset(c)) < 0) return rc;
  if ((rc = dpc_add_bulk(c, problems, n)) < 0) return rc;
  if ((rc = dpc_flush(c)) < 0) return rc;
  if ((rc = dpc_wait(c)) < 0) return rc;
  int64_t used = 0;
  Batch::Stack st;
  for (int i = 0; i < n; i++) {
    results[i] = c->batch.probs[i].res;
    if (pair_off) pair_off[i] = used;
    if (pairs) {
      int k = pairs_of(c, i, st);
      if (used + k > pair_cap) return DPC_ERR_NOMEM;
      if (k) memcpy(pairs + used, st.data(), (size_t)k * sizeof(dpc_pair_t));
      used += k;
    } else used += results[i].npairs;
  }
  if (pair_off) pair_off[n] = used;
  return DPC_OK;
}

int dpc_relaunch(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  if (!c->flushed || !c->waited) return DPC_ERR_STATE;
  if (c->batch.dprobs.empty()) return 0;
  CK(cudaSetDevice(c->device));
  int rc = launch_all(c);
  if (rc != DPC_OK) return rc;
  return c->nlaunch;
}

void *dpc_stream(dpc_ctx_t *c) { return c ? (void *)c->stream : NULL; }

int dpc_sync(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaEventElapsedTime(&c->ms_total, c->ev0, c->ev1));
  return DPC_OK;
}

int dpc_last_kernel_ms(dpc_ctx_t *c, float ms[3]) {
  if (!c || !ms) return DPC_ERR_ARG;
  if (!c->flushed) return DPC_ERR_STATE;
  ms[0] = c->ms_total; ms[1] = 0.0f; ms[2] = c->ms_total;   /* fill, bridge and traceback are one fused kernel */
  return DPC_OK;
}

int dpc_get_stats(dpc_ctx_t *c, dpc_stats_t *out) {
  if (!c || !out) return DPC_ERR_ARG;
  memset(out, 0, sizeof *out);
  const Batch &b = c->batch;
  out->nproblems = (int64_t)b.probs.size();
  for (size_t i = 0; i < b.dprobs.size(); i++) {
    const DevProb &p = b.dprobs[i];
    if ((p.kind == DPC_END5_GAP || p.kind == DPC_END3_GAP) && p.endalign == DPC_QUERYEND_NOGAPS) {
      out->fill_bytes += (int64_t)sizeof(DevProb) + p.L1 + (p.L2 + 3) / 4 + (int64_t)sizeof(DevRes);
      continue;
    }
    ArenaLayout a;
    dpc_layout(p, a, 0);
    for (int m = 0; m < a.nmat; m++) {
      const MatDims &d = a.d[m];
      out->nmatrices++;
      for (int cc = 1; cc <= d.cols; cc++) {               /* in-band cells, SURVEY.md 8(d) */
        int lo = cc - d.rband < 1 ? 1 : cc - d.rband, hi = cc + d.lband > d.rows ? d.rows : cc + d.lband;
        if (hi >= lo) out->cells += hi - lo + 1;
      }
      /* algorithmic HBM bytes: query bytes + 2-bit genome in */
      out->fill_bytes += (p.kind == DPC_CDNA_GAP) ? d.cols + (d.rows + 3) / 4 : d.rows + (d.cols + 3) / 4;
    }
    out->fill_bytes += (int64_t)sizeof(DevProb) + 4 + (int64_t)sizeof(DevRes);   /* descriptor + list entry in, result out */
  }
  out->traceback_bytes = 0;
  out->h2d_bytes = c->h2d_bytes; out->d2h_bytes = c->d2h_bytes;
  out->launches = c->nlaunch;
  return DPC_OK;
}

}  /* extern "C" */
