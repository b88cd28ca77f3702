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
#include <time.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "dpc_host.h"
#include "dpc_rows.h"
#include "dpc_pipe.h"

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
  uint8_t *gout;               /* staged genome characters, returned to the host with the results */
  uint32_t arena_bytes;        /* per-warp shared-memory arena (SMEM == true) */
  unsigned int *counter;       /* dynamic work distribution */
  int claim;                   /* work items per claim: 4 for launches of light problems (end gaps), else 1 */
};
/* light = fewer than this many lane-rows per problem on average (an end gap: ~30; a band-30 single gap: ~110) */
#define DPC_LIGHT_WORK 64
/* from a class's histogram over the work buckets (bucket b holds problems of about (63 - b) * 8 lane-rows) */
template <class COUNT> static int claim_of(const COUNT *bucket_count, int nbucket) {
  uint64_t n = 0, work = 0;
  for (int b = 0; b < nbucket; b++) { n += (uint64_t)bucket_count[b]; work += (uint64_t)bucket_count[b] * (uint64_t)(nbucket - 1 - b) * 8u; }
  return (n > 0 && work < (uint64_t)DPC_LIGHT_WORK * n) ? 4 : 1;
}

/* resident blocks per SM the compiler must leave room for: the one-matrix kernels fit DPC_MIN_BLOCKS_1M x 256
 * threads (3 -> 80 registers); their narrow-band instantiation (1 or 2 diagonals per lane only) is asked for
 * DPC_MIN_BLOCKS_NARROW; the two-matrix kernels keep 2 (128 registers) */
#ifndef DPC_MIN_BLOCKS_1M
#define DPC_MIN_BLOCKS_1M 3
#endif
#ifndef DPC_MIN_BLOCKS_NARROW
#define DPC_MIN_BLOCKS_NARROW 4
#endif
#ifndef DPC_MIN_BLOCKS_GENOME
#define DPC_MIN_BLOCKS_GENOME 3
#endif
/* Launch variants (one kernel instantiation each, so that each carries only its own code and address spaces):
 *   V_NARROW  arenas in shared memory, bulk region in the arena, bands of at most 64 diagonals (RowFillT<2>)
 *   V_WIDE    arenas in shared memory, bulk region in the arena, any band
 *   V_SPILL   small region in shared memory, bulk region (direction planes, nogap bands, ops) in HBM scratch
 *   V_HBM     everything in HBM scratch
 * KG: kind group (0 single gap, 1 genome gap, 2 cDNA gap, 3 end gaps); GEN: every matrix through the memory-state fill
 * (test hook; not instantiated for V_NARROW). */
enum { V_NARROW = 0, V_WIDE = 1, V_SPILL = 2, V_HBM = 3, NVARIANT = 4 };
template <int V, int KG, bool GEN, int CLAIM>
__global__ void __launch_bounds__(256, ((KG == 0 || KG == 3) ? (V == V_NARROW ? DPC_MIN_BLOCKS_NARROW : DPC_MIN_BLOCKS_1M) : (KG == 1 && (V == V_NARROW || V == V_SPILL) ? DPC_MIN_BLOCKS_GENOME : 2))) dpc_solve_kernel(const KernelArgs a) {
  constexpr bool SMEM = V != V_HBM;
  constexpr int BULK = (V == V_NARROW || V == V_WIDE) ? 1 : 0;
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
  /* work items are claimed CLAIM at a time and one claim ahead, so neither the atomic nor the next problem's
     descriptor is waited for (the list is ordered by work, not by address: adjacent items are alike) */
  int nxt = 0;
  if (ln.lane == 0) nxt = (int)atomicAdd(a.counter, (unsigned int)CLAIM);
  nxt = __shfl_sync(0xffffffffu, nxt, 0);
  for (;;) {
    const int base = nxt;
    if (base >= a.n) break;
    if (ln.lane == 0) nxt = (int)atomicAdd(a.counter, (unsigned int)CLAIM);
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    for (int u = 0; u < CLAIM; u++) {
    const int i = base + u;
    if (i >= a.n) break;
    const uint32_t pi = a.list[i];
    {
      const int j = u + 1 < CLAIM ? i + 1 : nxt;       /* the item after this one */
      if (j < a.n) {
        const uint32_t pn = a.list[j];
        const char *d = (const char *)&a.probs[pn];
        if (ln.lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(d + 64 * ln.lane));
      }
    }
    const DevProb p = a.probs[pi];
    uint8_t *scratch = a.scratch + (((uint64_t)p.scratch_hi << 32) | p.scratch_lo);
    uint8_t *arena = SMEM ? smem + (size_t)warp * a.arena_bytes : scratch;
    const uint32_t arena_bytes = SMEM ? a.arena_bytes : 0xffffffffu;
    if (GEN) {
      GenericFill fill;
      dpc_solve_problem<GenericFill, KG, BULK>(p, a.pool, a.blocks, &s_tables, arena, arena_bytes, scratch, &a.res[pi], a.ovf, a.gout, fill, ln);
    } else if (V == V_NARROW) {
      RowFillT<2, KG> fill;
      dpc_solve_problem<RowFillT<2, KG>, KG, BULK>(p, a.pool, a.blocks, &s_tables, arena, arena_bytes, scratch, &a.res[pi], a.ovf, a.gout, fill, ln);
    } else {
      RowFillT<DPC_MAX_CPL, KG> fill;
      dpc_solve_problem<RowFillT<DPC_MAX_CPL, KG>, KG, BULK>(p, a.pool, a.blocks, &s_tables, arena, arena_bytes, scratch, &a.res[pi], a.ovf, a.gout, fill, ln);
    }
    __syncwarp();
    }
  }
}

/* the exact-match scans of Dynprog_microexon_int (one block per scan; dpc_core.h: ScanQuery) */
__global__ void __launch_bounds__(256) dpc_microexon_scan_kernel(const ScanQuery *queries, const uint32_t *blocks, uint64_t nbases,
                                                                 uint32_t *hits, unsigned int *count) {
  const ScanQuery q = queries[blockIdx.x];
  for (int j = threadIdx.x; j < q.npos; j += blockDim.x)
    if (dpc_scan_match(q, blocks, nbases, j)) hits[q.hits_off + atomicAdd(&count[blockIdx.x], 1u)] = (uint32_t)j;
}

typedef void (*kernel_fn)(const KernelArgs);
/* [variant][kind group][generic]; the generic test hook runs the narrow class in the V_WIDE instantiation */
static kernel_fn kernel_of(int v, int kg, bool gen, int claim = 1) {
  /* launches of light one-matrix problems (end gaps) claim four work items at a time */
  if (v == V_NARROW && kg == 3 && !gen && claim > 1) return dpc_solve_kernel<V_NARROW, 3, false, 4>;
  static const kernel_fn tab[NVARIANT][DPC_NKG][2] = {
    { { dpc_solve_kernel<V_NARROW, 0, false, 1>, dpc_solve_kernel<V_WIDE, 0, true, 1> },
      { dpc_solve_kernel<V_NARROW, 1, false, 1>, dpc_solve_kernel<V_WIDE, 1, true, 1> },
      { dpc_solve_kernel<V_NARROW, 2, false, 1>, dpc_solve_kernel<V_WIDE, 2, true, 1> },
      { dpc_solve_kernel<V_NARROW, 3, false, 1>, dpc_solve_kernel<V_WIDE, 3, true, 1> } },
    { { dpc_solve_kernel<V_WIDE, 0, false, 1>, dpc_solve_kernel<V_WIDE, 0, true, 1> },
      { dpc_solve_kernel<V_WIDE, 1, false, 1>, dpc_solve_kernel<V_WIDE, 1, true, 1> },
      { dpc_solve_kernel<V_WIDE, 2, false, 1>, dpc_solve_kernel<V_WIDE, 2, true, 1> },
      { dpc_solve_kernel<V_WIDE, 3, false, 1>, dpc_solve_kernel<V_WIDE, 3, true, 1> } },
    { { dpc_solve_kernel<V_SPILL, 0, false, 1>, dpc_solve_kernel<V_SPILL, 0, true, 1> },
      { dpc_solve_kernel<V_SPILL, 1, false, 1>, dpc_solve_kernel<V_SPILL, 1, true, 1> },
      { dpc_solve_kernel<V_SPILL, 2, false, 1>, dpc_solve_kernel<V_SPILL, 2, true, 1> },
      { dpc_solve_kernel<V_SPILL, 3, false, 1>, dpc_solve_kernel<V_SPILL, 3, true, 1> } },
    { { dpc_solve_kernel<V_HBM, 0, false, 1>, dpc_solve_kernel<V_HBM, 0, true, 1> },
      { dpc_solve_kernel<V_HBM, 1, false, 1>, dpc_solve_kernel<V_HBM, 1, true, 1> },
      { dpc_solve_kernel<V_HBM, 2, false, 1>, dpc_solve_kernel<V_HBM, 2, true, 1> },
      { dpc_solve_kernel<V_HBM, 3, false, 1>, dpc_solve_kernel<V_HBM, 3, true, 1> } } };
  return tab[v][kg][gen ? 1 : 0];
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

#define CK_VOID(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) fprintf(stderr, "libdynprog_cuda: %s failed: %s\n", #call, cudaGetErrorString(e_)); } while (0)
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
  if (!d.ready) {
    /* how a host thread waits in dpc_wait: spin (CUDA's default with few threads), yield, or sleep on an interrupt.
       Many worker threads that share cores with other work (gmap -t N) are better off sleeping: DPC_SYNC=block. */
    const char *e = getenv("DPC_SYNC");
    if (e && !strcmp(e, "block")) cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync);
    else if (e && !strcmp(e, "yield")) cudaSetDeviceFlags(cudaDeviceScheduleYield);
    else if (e && !strcmp(e, "spin")) cudaSetDeviceFlags(cudaDeviceScheduleSpin);
    cudaGetLastError();
  }
  if (d.ready) {
    /* dpc_init / dpc_setup were called again: the genome mirror and the tables are replaced.  Work queued by live
       contexts must have drained (re-registration while solver calls are in flight is a caller error; waiting for
       the device at least keeps queued kernels off freed memory). */
    cudaDeviceSynchronize();
    cudaFree(d.d_blocks); cudaFree(d.d_tables); d.ready = false;
  }
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
  for (int v = 0; v < V_HBM; v++)
    for (int kg = 0; kg < DPC_NKG; kg++)
      for (int gen = 0; gen < 2; gen++)
        for (int claim = 1; claim <= 4; claim += 3)
          CK(cudaFuncSetAttribute((const void *)kernel_of(v, kg, gen != 0, claim), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  d.max_smem - (int)sizeof(DevTables) - 1024));
  d.version = g_version;
  d.ready = true;
  return DPC_OK;
}

/* ---- buffers -------------------------------------------------------------------------------- */
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
static void *pinned_alloc(size_t n) {
  void *p = NULL;
  if (cudaMallocHost(&p, n) != cudaSuccess) { cudaGetLastError(); return NULL; }
  return p;
}
static void pinned_release(void *p) { cudaFreeHost(p); }
static Alloc pinned() { Alloc a = { pinned_alloc, pinned_release }; return a; }

static int solve_grid(const DeviceState &dv, int variant, int kg, bool gen, size_t smem, int nproblems);

struct ClassLaunch {
  int variant;                     /* V_* */
  int kg;
  uint32_t arena_bytes;
  int wpb;
  size_t list_off;
  int n;
  int claim;
};

/* Launch classes.  Two shared-memory arena sizes: 6 KB per warp leaves most of the 228 KB to L1 (descriptors,
 * genome blocks and query bytes are read through it), 13 KB takes what is bigger (the register file allows at most
 * 2-4 blocks of 8 warps per SM, so 13 KB per warp never limits occupancy below that of the two-matrix kernels).  A
 * problem whose bulk region does not fit keeps it in HBM scratch (V_SPILL: only the small region -- characters, bridge
 * tables, under 4 KB at the largest sizes dpc_init accepts -- sits in shared memory, so the 6 KB arena does and the
 * launch keeps three blocks per SM); a problem whose small region alone does not fit runs entirely from HBM scratch
 * (V_HBM).  Narrow problems (every band <= 64 diagonals) get the narrow
 * instantiation. */
#define NCLASS 6
static const int k_class_variant[NCLASS] = { V_NARROW, V_NARROW, V_WIDE, V_WIDE, V_SPILL, V_HBM };
static uint32_t k_class_bytes[NCLASS] = { 6 << 10, 13 << 10, 6 << 10, 13 << 10, 6 << 10, 0 };
#define NBUCKET 64                       /* work buckets for longest-first scheduling */
#define SCRATCH_BUDGET (16ull << 30)

/* ---- engine: one stream, its device buffers and the batch in flight on it -------------------- */
struct Engine {
  int device;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  bool live;
  Batch batch;
  DBuf<DevProb> d_probs;
  DBuf<uint8_t> d_pool, d_scratch, d_gout;
  DBuf<DevRes> d_res;
  DBuf<uint32_t> d_list;
  DBuf<uint16_t> d_ovf;
  DBuf<unsigned int> d_counters;
  PBuf<DevRes> h_res;
  PBuf<uint16_t> h_ovf;
  PBuf<uint8_t> h_gout;
  PBuf<unsigned int> h_counters;
  PBuf<uint32_t> list;
  DBuf<ScanQuery> d_scans;           /* microexon searches */
  DBuf<uint32_t> d_hits;
  DBuf<unsigned int> d_hitcount;
  PBuf<uint32_t> h_hits;
  PBuf<unsigned int> h_hitcount;
  std::vector<uint16_t> cls;
  std::vector<ClassLaunch> launches;
  Scratch scratch;
  size_t ovf_cap;
  bool flushed, waited;
  bool dirty;                      /* something may be queued on the stream (set when a flush starts, even one that fails) */
  int flush_err;                   /* why the last flush failed (dpc_wait reports it) */
  int fill_gen;                    /* g_force_generic as it was when this batch was laid out */
  int nlaunch;
  float ms_total;
  int64_t h2d_bytes, d2h_bytes;
  double t_finalize;

  Engine() : device(0), stream(0), ev0(0), ev1(0), live(false), ovf_cap(0), flushed(false), waited(false), dirty(false), flush_err(0), fill_gen(0), nlaunch(0),
             ms_total(0), h2d_bytes(0), d2h_bytes(0), t_finalize(0) {}

  int open(int dev) {
    device = dev;
    CK(cudaSetDevice(dev));
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    batch.pool.set_alloc(pinned()); batch.dprobs.set_alloc(pinned());
    h_res.set_alloc(pinned()); h_ovf.set_alloc(pinned()); h_gout.set_alloc(pinned()); h_counters.set_alloc(pinned()); list.set_alloc(pinned());
    h_hits.set_alloc(pinned()); h_hitcount.set_alloc(pinned()); batch.scans.set_alloc(pinned());
    live = true;
    return DPC_OK;
  }
  void close() {
    if (!live) return;
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
    d_probs.release(); d_pool.release(); d_scratch.release(); d_gout.release(); d_res.release(); d_list.release();
    d_ovf.release(); d_counters.release(); d_scans.release(); d_hits.release(); d_hitcount.release();
    cudaEventDestroy(ev0); cudaEventDestroy(ev1); cudaStreamDestroy(stream);
    live = false;
  }
  ~Engine() { close(); }

  int reset() {
    if (dirty) { cudaSetDevice(device); cudaStreamSynchronize(stream); dirty = false; }
    batch.clear();
    flushed = waited = false;
    flush_err = 0;
    return DPC_OK;
  }

  int launch_all() {
    DeviceState &d = g_dev[device];
    const size_t ncnt = launches.size() + 1;
    CK(cudaMemsetAsync(d_counters.p, 0, ncnt * sizeof(unsigned int), stream));
    CK(cudaEventRecord(ev0, stream));
    nlaunch = 0;
    for (size_t k = 0; k < launches.size(); k++) {
      const ClassLaunch &L = launches[k];
      KernelArgs a;
      a.probs = d_probs.p; a.list = d_list.p + L.list_off; a.n = L.n;
      a.pool = d_pool.p; a.blocks = d.d_blocks; a.tables = d.d_tables; a.res = d_res.p;
      a.ovf.ops = d_ovf.p; a.ovf.used = d_counters.p; a.ovf.cap = (unsigned int)ovf_cap;
      a.scratch = d_scratch.p; a.gout = d_gout.p; a.arena_bytes = L.arena_bytes; a.counter = d_counters.p + 1 + k;
      a.claim = L.claim;
      const int threads = L.wpb * 32;
      const size_t smem = L.variant != V_HBM ? (size_t)L.wpb * L.arena_bytes : 0;
      const kernel_fn fn = kernel_of(L.variant, L.kg, fill_gen != 0, L.claim);
      const int grid = solve_grid(d, L.variant, L.kg, fill_gen != 0, smem, L.n);       /* occupancy query cached per kernel */
      fn<<<grid, threads, smem, stream>>>(a);
      CK(cudaGetLastError());
      nlaunch++;
    }
    CK(cudaEventRecord(ev1, stream));
    return DPC_OK;
  }

  /* H2D + kernels + D2H, all asynchronous on this engine's stream */
  int flush() {
    if (flushed) return DPC_ERR_STATE;
    /* `flushed` is only set once everything is queued: a failed flush leaves the context open for dpc_reset (or a
       retry after the caller freed memory), and dpc_wait reports the stored reason instead of a stale event */
    dirty = true; waited = false;
    flush_err = flush_impl();
    if (flush_err == DPC_OK) flushed = true;
    return flush_err;
  }
  int flush_impl() {
    CK(cudaSetDevice(device));
    DeviceState &d = g_dev[device];
    Batch &b = batch;
    const size_t n = b.dprobs.size();
    launches.clear();
    h2d_bytes = d2h_bytes = 0;
    if (!b.scans.empty()) {
      /* Dynprog_microexon_int: exact-match scans of the microexon candidates over their introns */
      const size_t nq = b.scans.size();
      int rc;
      if ((rc = d_scans.need(nq)) || (rc = d_hits.need((size_t)b.hits_total + 16)) || (rc = d_hitcount.need(nq))) return rc;
      h_hits.clear(); h_hits.grow((size_t)b.hits_total + 16);
      h_hitcount.clear(); h_hitcount.grow(nq);
      CK(cudaMemcpyAsync(d_scans.p, b.scans.data(), nq * sizeof(ScanQuery), cudaMemcpyHostToDevice, stream));
      CK(cudaMemsetAsync(d_hitcount.p, 0, nq * sizeof(unsigned int), stream));
      dpc_microexon_scan_kernel<<<(unsigned int)nq, 256, 0, stream>>>(d_scans.p, d.d_blocks, G().genome_nbases, d_hits.p, d_hitcount.p);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(h_hitcount.data(), d_hitcount.p, nq * sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
      CK(cudaMemcpyAsync(h_hits.data(), d_hits.p, (size_t)b.hits_total * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      h2d_bytes += (int64_t)(nq * sizeof(ScanQuery)); d2h_bytes += (int64_t)(nq * 4 + b.hits_total * 4);
    }
    if (n == 0) return DPC_OK;

    fill_gen = g_force_generic;                       /* one snapshot per batch: layout and kernels must agree */
    const int with_state = fill_gen ? 1 : 2;          /* must match FILL::fillmode of the kernel */
    /* class (shared-memory arena or HBM only) and a work bucket per problem: within a class the list is ordered
       by descending work so that the long problems start first and the tail of the launch is made of short ones */
    cls.resize(n);
    size_t count[NCLASS * DPC_NKG * NBUCKET] = { 0 };
    uint64_t scratch_total = 0, ovf_worst = 0;
    for (size_t i = 0; i < n; i++) {
      DevProb &p = b.dprobs[i];
      int k = 0, bucket = NBUCKET - 1;
      if (!((p.kind == DPC_END5_GAP || p.kind == DPC_END3_GAP) && p.endalign == DPC_QUERYEND_NOGAPS)) {
        ArenaLayout a;
        dpc_layout(p, a, with_state);
        uint64_t need = 0;                 /* HBM scratch of this problem */
        const int wide = dpc_narrow(a) ? 0 : 2;
        if (a.total <= k_class_bytes[0]) k = wide;
        else if (a.total <= k_class_bytes[1]) k = wide + 1;
        else if (a.small <= k_class_bytes[4]) { k = 4; need = a.bulk; }
        else { k = 5; need = a.total; }
        if (need) {
          if (scratch_total + need > SCRATCH_BUDGET) return DPC_ERR_NOMEM;
          p.scratch_lo = (uint32_t)scratch_total; p.scratch_hi = (uint32_t)(scratch_total >> 32);
          scratch_total += need;
        }
        uint64_t worst = 0, work = 0;
        for (int m = 0; m < a.nmat; m++) {
          worst += (uint64_t)(a.d[m].rows + a.d[m].cols + 2);
          work += (uint64_t)a.d[m].rows * (uint64_t)a.d[m].cpl;             /* row-sweep iterations x diagonals per lane */
        }
        if (worst > DPC_INLINE_OPS) ovf_worst += worst;
        /* bucket 0 = most work: 8 rows per bucket up to 504 lane-rows, everything longer in bucket 0 */
        int wb = (int)(work >> 3);
        if (wb > NBUCKET - 1) wb = NBUCKET - 1;
        bucket = NBUCKET - 1 - wb;
        static const bool nosort = getenv("DPC_NO_SORT") != NULL;
        if (nosort) bucket = 0;
      }
      cls[i] = (uint16_t)((k * DPC_NKG + dpc_kind_group(p.kind)) * NBUCKET + bucket);
      count[cls[i]]++;
    }
    list.clear();
    list.grow(n);
    size_t off[NCLASS * DPC_NKG * NBUCKET], at = 0;
    for (int k = 0; k < NCLASS * DPC_NKG * NBUCKET; k++) { off[k] = at; at += count[k]; }
    {
      size_t cur[NCLASS * DPC_NKG * NBUCKET];
      for (int k = 0; k < NCLASS * DPC_NKG * NBUCKET; k++) cur[k] = off[k];
      for (size_t i = 0; i < n; i++) list[cur[cls[i]]++] = (uint32_t)i;
    }
    for (int k = 0; k < NCLASS * DPC_NKG; k++) {       /* one launch per (arena class, kind group) that has work */
      size_t cnt = 0;
      for (int q = 0; q < NBUCKET; q++) cnt += count[k * NBUCKET + q];
      if (!cnt) continue;
      ClassLaunch L;
      L.variant = k_class_variant[k / DPC_NKG];
      L.kg = k % DPC_NKG;
      L.arena_bytes = k_class_bytes[k / DPC_NKG];
      L.wpb = 8;
      L.list_off = off[k * NBUCKET]; L.n = (int)cnt;
      L.claim = claim_of(count + k * NBUCKET, NBUCKET);
      launches.push_back(L);
    }
    /* ops overflow arena: problems whose worst case exceeds the inline slots (bounded) */
    if (ovf_worst > (1ull << 30)) ovf_worst = 1ull << 30;
    ovf_cap = (size_t)ovf_worst + 64;

    b.pool_align(16);
    int rc;
    if ((rc = d_probs.need(n)) || (rc = d_pool.need(b.pool.size())) || (rc = d_res.need(n)) ||
        (rc = d_list.need(n)) || (rc = d_ovf.need(ovf_cap)) || (rc = d_counters.need(NCLASS * DPC_NKG + 2)) ||
        (rc = d_scratch.need((size_t)scratch_total + 16)) || (rc = d_gout.need((size_t)b.gout_total + 64)))
      return rc;
    h_gout.clear(); h_gout.grow((size_t)b.gout_total + 64);
    b.gout_host = NULL;
    h_res.clear(); h_res.grow(n);
    h_counters.clear(); h_counters.grow(NCLASS * DPC_NKG + 2);
    CK(cudaMemcpyAsync(d_probs.p, b.dprobs.data(), n * sizeof(DevProb), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_pool.p, b.pool.data(), b.pool.size(), cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(d_list.p, list.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
    h2d_bytes = (int64_t)(n * sizeof(DevProb) + b.pool.size() + n * sizeof(uint32_t));
    if ((rc = launch_all()) != DPC_OK) return rc;
    CK(cudaMemcpyAsync(h_res.data(), d_res.p, n * sizeof(DevRes), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_counters.data(), d_counters.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(h_gout.data(), d_gout.p, b.gout_total, cudaMemcpyDeviceToHost, stream));
    d2h_bytes = (int64_t)(n * sizeof(DevRes) + sizeof(unsigned int) + b.gout_total);
    return DPC_OK;
  }

  const uint16_t *ops_of(const DevRes &dr) const {
    return (dr.nopsL + dr.nopsR > DPC_INLINE_OPS) ? h_ovf.data() + dr.ovf : dr.ops;
  }

  /* blocks until the batch is back, then turns every device record into the reference's outputs */
  int wait() {
    if (!flushed) return flush_err ? flush_err : DPC_ERR_STATE;
    if (waited) return DPC_OK;
    Batch &b = batch;
    const size_t n = b.dprobs.size();
    if (!b.micro.empty()) {
      CK(cudaSetDevice(device));
      CK(cudaStreamSynchronize(stream));
      dirty = false;
      for (size_t k = 0; k < b.micro.size(); k++) b.finalize_micro(b.micro[k], h_hits.data(), h_hitcount.data());
    }
    if (n > 0) {
      CK(cudaSetDevice(device));
      CK(cudaStreamSynchronize(stream));
      dirty = false;
      CK(cudaEventElapsedTime(&ms_total, ev0, ev1));
      static const bool no_gout = getenv("DPC_NO_GOUT") != NULL;      /* measurement aid: rebuild from the host's 2-bit genome */
      b.gout_host = no_gout ? NULL : h_gout.data();
      unsigned int used = h_counters[0];
      if (used > 0) {
        if (used > ovf_cap) return DPC_ERR_NOMEM;
        h_ovf.clear(); h_ovf.grow(used);
        CK(cudaMemcpyAsync(h_ovf.data(), d_ovf.p, used * sizeof(uint16_t), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        d2h_bytes += (int64_t)used * 2;
      }
      struct timespec ta, tb;
      clock_gettime(CLOCK_MONOTONIC, &ta);
      for (size_t k = 0; k < n; k++) {
        const DevRes &dr = h_res[k];
        if (!(dr.status & DPC_ST_DONE) || (dr.status & DPC_ST_OVF_LOST)) return DPC_ERR_CUDA;
        b.finalize((int)b.dev2host[k], dr, ops_of(dr), scratch);
      }
      clock_gettime(CLOCK_MONOTONIC, &tb);
      t_finalize = (tb.tv_sec - ta.tv_sec) + 1e-9 * (tb.tv_nsec - ta.tv_nsec);
    }
    waited = true;
    return DPC_OK;
  }

  int pairs_into(int ticket, dpc_pair_t *dst, bool stream_dst = false) {
    const HostProb &h = batch.probs[ticket];
    if (h.micro >= 0) {
      const std::vector<dpc_pair_t> &v = batch.micro[(size_t)h.micro].pairs;
      if (!v.empty()) memcpy(dst, v.data(), v.size() * sizeof(dpc_pair_t));
      return (int)v.size();
    }
    if (h.dev < 0) return 0;
    const DevRes &dr = h_res[h.dev];
    return batch.rebuild(ticket, dr, ops_of(dr), dst, scratch, stream_dst);
  }
};

/* ---- host worker threads of the bulk call -------------------------------------------------------- */
struct Workers {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::function<void(int)> fn;
  std::atomic<int> next;
  int njobs, active;
  uint64_t gen;
  bool stop;
  explicit Workers(int n) : next(0), njobs(0), active(0), gen(0), stop(false) {
    for (int t = 0; t < n; t++) th.emplace_back([this] { loop(); });
  }
  ~Workers() {
    { std::lock_guard<std::mutex> l(mu); stop = true; }
    cv.notify_all();
    for (auto &t : th) t.join();
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [&] { return stop || gen != seen; });
        if (stop) return;
        seen = gen;
      }
      for (int j; (j = next.fetch_add(1)) < njobs;) fn(j);
      {
        std::lock_guard<std::mutex> l(mu);
        if (--active == 0) cv_done.notify_one();
      }
    }
  }
  void run(int jobs, const std::function<void(int)> &f) {
    std::unique_lock<std::mutex> l(mu);
    fn = f; njobs = jobs; next = 0; active = (int)th.size(); gen++;
    cv.notify_all();
    cv_done.wait(l, [&] { return active == 0; });
  }
};

#include "dpc_pipeline.cuh"

struct dpc_ctx {
  Engine main;                       /* ticket API, dpc_relaunch */
  std::vector<int> devices;          /* bulk API: the devices the chunks are dealt to (main.device first) */
  std::vector<std::vector<Pipe *> > pipes;     /* bulk API: pipeline slots per device */
  std::vector<std::vector<Engine *> > engines; /* bulk API: engines for chunks that need the host half, per device */
  Workers *workers;
  int nthreads;
  std::atomic<int64_t> solve_h2d, solve_d2h, solve_launches, solve_problems, solve_pipe_chunks, solve_host_chunks;   /* totals of the last dpc_solve */
  dpc_ctx() : workers(NULL), nthreads(1), solve_h2d(0), solve_d2h(0), solve_launches(0), solve_problems(0), solve_pipe_chunks(0), solve_host_chunks(0) {}
};
static int g_path = 0;               /* dpc_set_path: 0 auto, 1 host half only, 2 device pipeline or fail, 3 / 4 = 2 with the
                                        Pair records expanded on the device / by the host */

/* ---- calibration of the roofline denominators (SURVEY.md 8d asks for a measured INT32 peak, not an assumed one) ---- */
#define PEAK_ITER 4096
template <int MODE> __global__ void __launch_bounds__(256) dpc_int_peak_kernel(int *out, int a, int b) {
  int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 8
  for (int i = 0; i < PEAK_ITER; i++) {
    if (MODE == 0) {          /* VIMNMX + LOP3: both on the integer ALU pipe (the pipe the fill saturates) */
      x0 = max(x0, a) ^ b; x1 = max(x1, a) ^ b; x2 = max(x2, a) ^ b; x3 = max(x3, a) ^ b;
      x4 = max(x4, a) ^ b; x5 = max(x5, a) ^ b; x6 = max(x6, a) ^ b; x7 = max(x7, a) ^ b;
    } else {                  /* the fill's mix -- add, max, compare/select -- where the adds may go to the FMA pipe as IMAD */
      x0 = max(x0 + a, x1); x1 = x1 > x2 ? x1 + b : x2; x2 = max(x2 + a, x3); x3 = x3 > x4 ? x3 + b : x4;
      x4 = max(x4 + a, x5); x5 = x5 > x6 ? x5 + b : x6; x6 = max(x6 + a, x7); x7 = x7 > x0 ? x7 + b : x0;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}
template <int MODE> static int int_peak_run(int *d, int grid, double *gops) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; w++) dpc_int_peak_kernel<MODE><<<grid, 256>>>(d, 3, 5);
  CK(cudaEventRecord(e0));
  const int reps = 20;
  for (int r = 0; r < reps; r++) dpc_int_peak_kernel<MODE><<<grid, 256>>>(d, 3, 5);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *gops = (double)grid * 256 * PEAK_ITER * 16.0 * reps / (ms * 1e-3) / 1e9;      /* 8 chains x 2 lane-operations per step */
  return DPC_OK;
}

/* ---- C ABI ---------------------------------------------------------------------------------- */
extern "C" {

int dpc_init(int maxlookback, int extraquerygap, int maxpeelback, int extramaterial_end, int extramaterial_paired, int mode) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = host_init(maxlookback, extraquerygap, maxpeelback, extramaterial_end, extramaterial_paired, mode);
  g_version++;
  const char *kb = getenv("DPC_CLASS0_KB");          /* tuning aid: size of the small shared-memory class */
  if (kb && atoi(kb) >= 1 && atoi(kb) <= 13) k_class_bytes[0] = k_class_bytes[2] = (uint32_t)atoi(kb) << 10;
  kb = getenv("DPC_CLASS1_KB");                      /* ... and of the big one */
  if (kb && atoi(kb) >= 1 && atoi(kb) <= 27) k_class_bytes[1] = k_class_bytes[3] = (uint32_t)atoi(kb) << 10;
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

int dpc_warmup(int device) {
  if (device < 0 || device >= dpc_device_count()) return DPC_ERR_CUDA;
  CK(cudaSetDevice(device));
  CK(cudaFree(0));
  return DPC_OK;
}

dpc_ctx_t *dpc_ctx_new(int device) {
  if (device < 0 || device >= dpc_device_count()) return NULL;
  if (ensure_device(device) != DPC_OK) return NULL;
  dpc_ctx *c = new dpc_ctx();
  if (c->main.open(device) != DPC_OK) { delete c; return NULL; }
  int t = (int)std::thread::hardware_concurrency();
  const char *e = getenv("DPC_HOST_THREADS");
  if (e && atoi(e) > 0) t = atoi(e);
  if (t < 1) t = 1;
  if (t > 64) t = 64;
  c->nthreads = t;
  c->devices.push_back(device);
  return c;
}

dpc_ctx_t *dpc_ctx_new_multi(const int *devices, int ndevices) {
  if (!devices || ndevices < 1 || ndevices > MAXDEV) return NULL;
  for (int i = 0; i < ndevices; i++)
    for (int k = 0; k < i; k++) if (devices[k] == devices[i]) return NULL;
  dpc_ctx *c = dpc_ctx_new(devices[0]);
  if (!c) return NULL;
  for (int i = 1; i < ndevices; i++) {
    if (devices[i] < 0 || devices[i] >= dpc_device_count() || ensure_device(devices[i]) != DPC_OK) { dpc_ctx_free(c); return NULL; }
    c->devices.push_back(devices[i]);
  }
  return c;
}

int dpc_host_register(void *p, uint64_t bytes) {
  if (!p || !bytes) return DPC_ERR_ARG;
  if (cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return DPC_ERR_CUDA; }
  return DPC_OK;
}
int dpc_host_unregister(void *p) {
  if (!p) return DPC_ERR_ARG;
  if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return DPC_ERR_CUDA; }
  return DPC_OK;
}
int dpc_set_path(int path) {
  if (path < 0 || path > 4) return DPC_ERR_ARG;
  g_path = path;
  return DPC_OK;
}

void dpc_ctx_free(dpc_ctx_t *c) {
  if (!c) return;
  delete c->workers;
  for (size_t d = 0; d < c->pipes.size(); d++) for (size_t i = 0; i < c->pipes[d].size(); i++) delete c->pipes[d][i];
  for (size_t d = 0; d < c->engines.size(); d++) for (size_t i = 0; i < c->engines[d].size(); i++) delete c->engines[d][i];
  delete c;
}

int dpc_set_threads(dpc_ctx_t *c, int nthreads) {
  if (!c || nthreads < 1 || nthreads > 64) return DPC_ERR_ARG;
  if (c->workers && nthreads != c->nthreads) { delete c->workers; c->workers = NULL; }
  c->nthreads = nthreads;
  return DPC_OK;
}

#define GUARD(expr) do { try { expr; } catch (const std::bad_alloc &) { return DPC_ERR_NOMEM; } } while (0)

int dpc_add(dpc_ctx_t *c, const dpc_problem_t *problem) {
  if (!c || !problem) return DPC_ERR_ARG;
  if (c->main.flushed) return DPC_ERR_STATE;
  int rc = 0;
  GUARD(rc = c->main.batch.add(*problem));
  return rc;
}

int dpc_add_bulk(dpc_ctx_t *c, const dpc_problem_t *problems, int n) {
  if (!c || (!problems && n > 0) || n < 0) return DPC_ERR_ARG;
  if (c->main.flushed) return DPC_ERR_STATE;
  Batch &b = c->main.batch;
  int first = (int)b.probs.size();
  try {
    b.probs.reserve(b.probs.size() + (size_t)n);
    b.dprobs.reserve(b.dprobs.size() + (size_t)n);
    for (int i = 0; i < n; i++) {
      int t = b.add(problems[i]);
      if (t < 0) return t;
    }
  } catch (const std::bad_alloc &) { return DPC_ERR_NOMEM; }
  return first;
}

int dpc_reset(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  return c->main.reset();
}

int dpc_flush(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  int rc = ensure_device(c->main.device);
  if (rc != DPC_OK) return rc;
  GUARD(rc = c->main.flush());
  return rc;
}

int dpc_wait(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  int rc = 0;
  GUARD(rc = c->main.wait());
  return rc;
}

int dpc_result(dpc_ctx_t *c, int ticket, dpc_result_t *out) {
  if (!c || !out) return DPC_ERR_ARG;
  if (!c->main.waited || ticket < 0 || ticket >= (int)c->main.batch.probs.size()) return DPC_ERR_STATE;
  *out = c->main.batch.R(ticket);
  return DPC_OK;
}

int dpc_pairs(dpc_ctx_t *c, int ticket, dpc_pair_t *out, int cap) {
  if (!c) return DPC_ERR_ARG;
  Engine &e = c->main;
  if (!e.waited || ticket < 0 || ticket >= (int)e.batch.probs.size()) return DPC_ERR_STATE;
  int n = e.batch.R(ticket).npairs;
  if (n > cap || (n > 0 && !out)) return DPC_ERR_ARG;
  if (n == 0) return 0;
  int k = 0;
  try {
    dpc_pair_t *tmp = Batch::fit(e.scratch.out, e.batch.max_pairs(ticket));
    k = e.pairs_into(ticket, tmp);
    memcpy(out, tmp, (size_t)k * sizeof(dpc_pair_t));
  } catch (const std::bad_alloc &) { return DPC_ERR_NOMEM; }
  return k;
}

/* Bulk call.  The problems are cut into chunks; chunk j belongs to device j mod ndevices (dpc_ctx_new_multi).  Per
 * device ONE driver thread walks its chunks through the device pipeline (dpc_pipeline.cuh) as a state machine over a
 * handful of pipeline slots: it only queues copies and kernels and polls events, so a dozen chunks are in flight per
 * device whatever the number of host cores.  The other host threads are workers: they scan chunks ahead of the
 * driver (does a chunk need a host hook? where do its query bytes lie?), expand the Pair records of the chunks the
 * driver routes to the host, apply the MaxEnt hook, and serve whole chunks that need the host half (known splice
 * sites, probability mode, splice-junction solvers) with an Engine, like the ticket API.  Results and pairs come
 * out in input order; the only ordered step is the running pair offset handed from chunk to chunk. */
static double now_s() {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

enum { JS_NEW = 0, JS_WAIT1, JS_WAIT2, JS_CHAIN, JS_WAIT3, JS_WORKER, JS_DONE };
enum { WK_FINAL = 0, WK_REBUILD, WK_LEGACY_A, WK_LEGACY_B };
struct SolveJob {
  int lo, cnt;
  std::atomic<int> scanned, mine_known, at_known, worker_done;
  ChunkScan cs;
  int state, route;
  bool piped;
  Pipe *pp;
  Engine *eng;
  int64_t mine, at, queued;
  SolveJob() : lo(0), cnt(0), scanned(0), mine_known(0), at_known(0), worker_done(0), state(JS_NEW), route(0), piped(false),
               pp(NULL), eng(NULL), mine(0), at(0), queued(0) {}
};

int dpc_solve(dpc_ctx_t *c, const dpc_problem_t *problems, int n, dpc_result_t *results,
              dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off) {
  if (!c || n < 0 || (n > 0 && (!problems || !results))) return DPC_ERR_ARG;
  const int ndev = (int)c->devices.size();
  for (int d = 0; d < ndev; d++) { int rc = ensure_device(c->devices[(size_t)d]); if (rc != DPC_OK) return rc; }
  if (pair_off && n == 0) pair_off[0] = 0;
  if (n == 0) return DPC_OK;
  /* threads: drivers (two per device when the budget of dpc_set_threads allows: queueing a chunk's ~25 copies and
     launches costs ~0.1 ms of a thread) + workers (at least one) */
  static const int env_drivers = getenv("DPC_DRIVERS") ? std::max(1, atoi(getenv("DPC_DRIVERS"))) : 0;
  const int dpd = env_drivers ? env_drivers : (c->nthreads >= 6 * ndev ? 2 : 1);       /* drivers per device */
  const int ndrv = dpd * ndev;
  const int nworkers = std::max(1, c->nthreads - ndrv);
  const int T = ndrv + nworkers;
  if (c->workers && (int)c->workers->th.size() != T) { delete c->workers; c->workers = NULL; }
  if (!c->workers) c->workers = new Workers(T);
  static const int nslots = getenv("DPC_PIPE_SLOTS") ? std::max(2, atoi(getenv("DPC_PIPE_SLOTS"))) : 12;
  c->pipes.resize((size_t)ndrv); c->engines.resize((size_t)ndrv);
  /* chunk size: small enough that the workers' last rebuilds do not trail the GPU by much, large enough that a
     driver's ~0.1 ms of queueing per chunk stays off the critical path */
  int chunk = n / (3 * nslots * ndrv) + 1;
  if (chunk < 2048) chunk = 2048;
  if (chunk > (dpd >= 2 ? 8192 : 16384)) chunk = dpd >= 2 ? 8192 : 16384;
  { static const int forced = getenv("DPC_CHUNK") ? atoi(getenv("DPC_CHUNK")) : 0; if (forced >= 256) chunk = forced; }   /* tuning aid */
  const int nchunks = (n + chunk - 1) / chunk;
  static const bool timing = getenv("DPC_TIMING") != NULL;
  static const bool stream_pairs = getenv("DPC_NO_STREAM") == NULL;     /* non-temporal stores for rebuilt pair records */
  static const double host_depth = getenv("DPC_HOST_DEPTH") ? atof(getenv("DPC_HOST_DEPTH")) : 2.0;
  static const int env_route = getenv("DPC_ROUTE") ? (!strcmp(getenv("DPC_ROUTE"), "host") ? 1 : !strcmp(getenv("DPC_ROUTE"), "device") ? 0 : -1) : -1;
  const int forced_route = g_path == 3 ? 0 : g_path == 4 ? 1 : env_route;
  /* tuning aid: DPC_ROUTE_MIX=p sends p percent of the chunks through the host route, evenly spread */
  static const int env_mix = getenv("DPC_ROUTE_MIX") ? atoi(getenv("DPC_ROUTE_MIX")) : -1;
  std::atomic<int> mix_acc(0);
  const bool direct = pairs != NULL && host_pinned(pairs, (size_t)pair_cap * sizeof(dpc_pair_t));
  const bool hp_pinned = host_pinned(problems, (size_t)n * sizeof(dpc_problem_t));
  const bool res_pinned = host_pinned(results, (size_t)n * sizeof(dpc_result_t));
  const int fill_gen = g_force_generic;
  const int path = g_path;

  std::vector<SolveJob> jobs((size_t)nchunks);
  for (int j = 0; j < nchunks; j++) { jobs[(size_t)j].lo = j * chunk; jobs[(size_t)j].cnt = std::min(chunk, n - j * chunk); }
  std::atomic<int> err(0), scan_next(0), drivers_left(ndrv), host_routed(0), rebuild_queued(0);
  std::atomic<int64_t> t_scan(0), t_work(0), t_legacy(0), d2h_backlog(0);
  std::mutex qmu, chain_mu;
  std::deque<std::pair<int, int> > queue;          /* (work kind, chunk) */
  int chain_next = 0; int64_t chain_total = 0;
  std::vector<Scratch> scratch((size_t)T);
  const double t0 = now_s();
  c->solve_h2d = 0; c->solve_d2h = 0; c->solve_launches = 0; c->solve_problems = n; c->solve_pipe_chunks = 0; c->solve_host_chunks = 0;
  auto fail = [&](int code) { int z = 0; err.compare_exchange_strong(z, code); };
  auto push = [&](int kind, int j) { std::lock_guard<std::mutex> l(qmu); queue.push_back(std::make_pair(kind, j)); };

  /* hands out the running pair offset in chunk order; host-half chunks get their second half queued here */
  auto advance_chain = [&]() {
    std::lock_guard<std::mutex> l(chain_mu);
    while (chain_next < nchunks && jobs[(size_t)chain_next].mine_known.load(std::memory_order_acquire)) {
      SolveJob &jb = jobs[(size_t)chain_next];
      jb.at = chain_total;
      chain_total += jb.mine;
      if (pairs && chain_total > pair_cap) fail(DPC_ERR_NOMEM);
      jb.at_known.store(1, std::memory_order_release);
      if (!jb.piped && !err.load()) push(WK_LEGACY_B, chain_next);
      chain_next++;
    }
  };

  auto fill_pair_off = [&](const SolveJob &jb, const long long *off) {
    if (pair_off) for (int i = 0; i < jb.cnt; i++) pair_off[jb.lo + i] = jb.at + off[i];
  };

  /* ---- worker: queued work first, else scan ahead ---- */
  auto worker = [&](int w) {
    Scratch &sc = scratch[(size_t)w];
    int idle = 0;
    for (;;) {
      std::pair<int, int> it(-1, -1);
      { std::lock_guard<std::mutex> l(qmu); if (!queue.empty()) { it = queue.front(); queue.pop_front(); } }
      if (it.first >= 0) {
        idle = 0;
        SolveJob &jb = jobs[(size_t)it.second];
        const dpc_problem_t *P = problems + jb.lo;
        dpc_result_t *R = results + jb.lo;
        const double a0 = timing ? now_s() : 0;
        int r = DPC_OK;
        try {
          switch (it.first) {
          case WK_FINAL:
            pipe_device_route_final(*jb.pp, P, jb.cnt, R, pairs ? pairs + jb.at : NULL, direct, jb.mine);
            fill_pair_off(jb, jb.pp->h_off.data());
            jb.worker_done.store(1, std::memory_order_release);
            break;
          case WK_REBUILD:
            if (!err.load()) r = pipe_host_route_final(*jb.pp, P, jb.cnt, R, pairs + jb.at, jb.mine, stream_pairs, sc);
            fill_pair_off(jb, jb.pp->h_off.data());
            rebuild_queued -= 1;
            jb.worker_done.store(1, std::memory_order_release);
            break;
          case WK_LEGACY_A: {
            Engine &e = *jb.eng;
            if (!err.load()) {
              e.reset();
              r = e.batch.add_ext(P, R, jb.cnt);
              if (r >= 0) r = e.flush();
              if (r >= 0) r = e.wait();
              if (r >= 0) for (int i = 0; i < jb.cnt; i++) jb.mine += R[i].npairs;
              else jb.mine = 0;
            }
            jb.mine_known.store(1, std::memory_order_release);
            if (timing) t_legacy += (int64_t)((now_s() - a0) * 1e9);
            break;
          }
          case WK_LEGACY_B: {
            Engine &e = *jb.eng;
            int64_t at = jb.at;
            for (int i = 0; i < jb.cnt && (pairs || pair_off) && !err.load(); i++) {
              const int np = R[i].npairs;
              if (pairs && i + 8 < jb.cnt) e.batch.prefetch_genome(i + 8);
              if (pair_off) pair_off[jb.lo + i] = at;
              if (!pairs || np == 0) { at += np; continue; }
              const int k = e.pairs_into(i, pairs + at, stream_pairs);
              if (k != np) { r = DPC_ERR_STATE; break; }
              at += k;
            }
            c->solve_h2d += e.h2d_bytes; c->solve_d2h += e.d2h_bytes; c->solve_launches += e.nlaunch; c->solve_host_chunks += 1;
            jb.worker_done.store(1, std::memory_order_release);
            break;
          }
          }
        } catch (const std::bad_alloc &) { r = DPC_ERR_NOMEM; }
        if (r < 0) {
          fail(r);
          if (it.first == WK_LEGACY_A) { jb.mine = 0; jb.mine_known.store(1, std::memory_order_release); }
          else jb.worker_done.store(1, std::memory_order_release);
        }
        if (timing && it.first != WK_LEGACY_A) t_work += (int64_t)((now_s() - a0) * 1e9);
        continue;
      }
      int sidx = scan_next.load();
      if (sidx < nchunks) {
        if (scan_next.compare_exchange_strong(sidx, sidx + 1)) {
          const double a0 = timing ? now_s() : 0;
          SolveJob &jb = jobs[(size_t)sidx];
          jb.cs.eligible = false;
          if (path != 1) jb.cs = scan_chunk(problems + jb.lo, jb.cnt);
          jb.scanned.store(1, std::memory_order_release);
          if (timing) t_scan += (int64_t)((now_s() - a0) * 1e9);
        }
        idle = 0;
        continue;
      }
      if (drivers_left.load() == 0) { std::lock_guard<std::mutex> l(qmu); if (queue.empty()) break; else continue; }
      if (++idle < 64) std::this_thread::yield(); else { struct timespec ts = { 0, 20000 }; nanosleep(&ts, NULL); }
    }
  };

  /* ---- driver number dslot: chunks dslot, dslot + ndrv, ... on device dslot mod ndev ---- */
  auto driver = [&](int dslot) {
    const int device = c->devices[(size_t)(dslot % ndev)];
    cudaSetDevice(device);
    std::vector<Pipe *> &pool = c->pipes[(size_t)dslot];
    std::vector<Engine *> &engs = c->engines[(size_t)dslot];
    std::vector<Pipe *> free_pipes;
    std::vector<Engine *> free_engs;
    std::vector<int> active, legacy_active;
    for (size_t i = 0; i < pool.size(); i++) free_pipes.push_back(pool[i]);
    for (size_t i = 0; i < engs.size(); i++) free_engs.push_back(engs[i]);
    int next_start = dslot, left = 0;
    for (int j = dslot; j < nchunks; j += ndrv) left++;
    int idle = 0;
    auto job_fail = [&](SolveJob &jb, int code) {
      fail(code);
      if (!jb.mine_known.load()) { jb.mine = 0; jb.mine_known.store(1, std::memory_order_release); }
      jb.state = JS_DONE;
    };
    while (left > 0) {
      bool progressed = false;
      const bool failing = err.load() != 0;
      /* 1. start chunks, in order */
      while (next_start < nchunks && jobs[(size_t)next_start].scanned.load(std::memory_order_acquire)) {
        SolveJob &jb = jobs[(size_t)next_start];
        if (failing) {                  /* after an error nothing new is queued; the chain still has to move */
          jb.mine = 0; jb.piped = true; jb.mine_known.store(1, std::memory_order_release);
          left--; next_start += ndrv; progressed = true;
          continue;
        }
        if (path >= 2 && !jb.cs.eligible) { jb.piped = true; job_fail(jb, DPC_ERR_STATE); left--; next_start += ndrv; progressed = true; continue; }
        if (!jb.cs.eligible) {
          if ((int)legacy_active.size() >= nworkers + 1) break;
          Engine *e = NULL;
          if (!free_engs.empty()) { e = free_engs.back(); free_engs.pop_back(); }
          else {
            e = new Engine();
            if (e->open(device) != DPC_OK) { delete e; jb.piped = true; job_fail(jb, DPC_ERR_CUDA); left--; next_start += ndrv; continue; }
            engs.push_back(e);
          }
          jb.eng = e; jb.piped = false;
          legacy_active.push_back(next_start);
          push(WK_LEGACY_A, next_start);
        } else {
          Pipe *pp = NULL;
          if (!free_pipes.empty()) { pp = free_pipes.back(); free_pipes.pop_back(); }
          else if ((int)pool.size() < nslots) {
            pp = new Pipe();
            if (pp->open(device) != DPC_OK) { delete pp; jb.piped = true; job_fail(jb, DPC_ERR_CUDA); left--; next_start += ndrv; continue; }
            pool.push_back(pp);
          } else break;
          jb.pp = pp; jb.piped = true;
          int r = DPC_OK;
          try { r = pipe_stage1(*pp, problems + jb.lo, jb.cnt, jb.cs, fill_gen, hp_pinned); } catch (const std::bad_alloc &) { r = DPC_ERR_NOMEM; }
          if (r < 0) { free_pipes.push_back(pp); job_fail(jb, r); left--; next_start += ndrv; continue; }
          jb.state = JS_WAIT1;
          active.push_back(next_start);
        }
        next_start += ndrv;
        progressed = true;
      }
      /* 2. advance the chunks in flight */
      for (size_t a = 0; a < active.size();) {
        SolveJob &jb = jobs[(size_t)active[a]];
        Pipe &pp = *jb.pp;
        int r = DPC_OK;
        bool moved = false;
        try {
          if (jb.state == JS_WAIT1 || jb.state == JS_WAIT2 || jb.state == JS_WAIT3) {
            const cudaError_t q = cudaEventQuery(pp.ev);
            if (q == cudaSuccess) {
              moved = true;
              if (jb.state == JS_WAIT1) {
                r = pipe_stage2(pp, jb.cnt, fill_gen);
                jb.state = JS_WAIT2;
              } else if (jb.state == JS_WAIT2) {
                const PipeCounters &pc = pp.h_pc[0];
                if (pc.err < 0) r = pc.err;
                else {
                  jb.mine = pc.pair_total;
                  /* Who expands this chunk's traceback ops into Pair records?  A worker (the compact device records
                     over PCIe, ~200 B per problem, then the host half's rebuild) or the device (16 bytes per record
                     over PCIe, ~900 B per problem).  Measured on B200 boxes (profiles/r2_route_mix.txt): the host
                     route is the faster one whenever the workers keep up, and a MIX of the two is slower than either
                     alone -- the copy engine writing 1 GB of records into host memory slows the host threads' own
                     scans and rebuilds by 2-3 x.  So: host route while the workers have at most `host_depth` chunks
                     each waiting, device route for what they cannot absorb. */
                  jb.route = 0;
                  if (pairs && jb.mine > 0) {
                    const bool workers_free = (double)rebuild_queued.load() < host_depth * (double)nworkers;
                    jb.route = forced_route >= 0 ? forced_route : (workers_free ? 1 : 0);
                    if (forced_route < 0 && env_mix >= 0) {
                      jb.route = 0;
                      if (mix_acc.fetch_add(env_mix) % 100 + env_mix >= 100) jb.route = 1;
                    }
                  }
                  if (jb.route == 0) {
                    jb.queued = pairs ? jb.mine * (int64_t)sizeof(dpc_pair_t) : 0;
                    d2h_backlog += jb.queued;
                    r = pipe_device_route_start(pp, jb.cnt, results + jb.lo, res_pinned, pairs != NULL, jb.mine);
                  } else {
                    rebuild_queued += 1;
                    host_routed += 1;
                    r = pipe_host_route_start(pp, jb.cnt, results + jb.lo, res_pinned);
                  }
                  jb.mine_known.store(1, std::memory_order_release);
                  jb.state = JS_CHAIN;
                }
              } else {
                d2h_backlog -= jb.queued; jb.queued = 0;
                push(jb.route == 0 ? WK_FINAL : WK_REBUILD, active[a]);
                jb.state = JS_WORKER;
              }
            } else if (q != cudaErrorNotReady) {
              fprintf(stderr, "libdynprog_cuda: device pipeline failed: %s\n", cudaGetErrorString(q));
              r = DPC_ERR_CUDA;
            }
          }
          if (r >= 0 && jb.state == JS_CHAIN && jb.at_known.load(std::memory_order_acquire)) {
            moved = true;
            if (err.load()) { jb.state = JS_WAIT3; CK_VOID(cudaEventRecord(pp.ev, pp.stream)); }
            else if (jb.route == 0) { r = pipe_device_route_copy(pp, pairs ? pairs + jb.at : NULL, direct, jb.mine); jb.state = JS_WAIT3; }
            else jb.state = JS_WAIT3;          /* the event of the host route was recorded with its copies */
          }
          if (jb.state == JS_WORKER && jb.worker_done.load(std::memory_order_acquire)) {
            moved = true;
            c->solve_h2d += pp.h2d_bytes; c->solve_d2h += pp.d2h_bytes; c->solve_launches += pp.nlaunch; c->solve_pipe_chunks += 1;
            jb.state = JS_DONE;
          }
        } catch (const std::bad_alloc &) { r = DPC_ERR_NOMEM; }
        if (r < 0) {
          d2h_backlog -= jb.queued; jb.queued = 0;
          cudaStreamSynchronize(pp.stream);
          job_fail(jb, r);
        }
        if (moved) progressed = true;
        if (jb.state == JS_DONE) {
          free_pipes.push_back(jb.pp);
          active[a] = active.back(); active.pop_back();
          left--;
        } else a++;
      }
      for (size_t a = 0; a < legacy_active.size();) {
        SolveJob &jb = jobs[(size_t)legacy_active[a]];
        bool finished = jb.worker_done.load(std::memory_order_acquire) != 0;
        if (!finished && err.load() && jb.mine_known.load() && jb.at_known.load()) finished = true;    /* second half never queued */
        if (finished) {
          free_engs.push_back(jb.eng);
          legacy_active[a] = legacy_active.back(); legacy_active.pop_back();
          left--; progressed = true;
        } else a++;
      }
      /* 3. the running pair offset */
      advance_chain();
      if (progressed) idle = 0;
      else if (++idle < 200) { /* spin: events complete within microseconds */ }
      else std::this_thread::yield();
    }
    drivers_left -= 1;
  };

  c->workers->run(T, [&](int w) { if (w < ndrv) driver(w); else worker(w); });
  advance_chain();
#if defined(__SSE2__) && defined(__x86_64__)
  _mm_sfence();
#endif
  if (err.load()) return err.load();
  if (pair_off) pair_off[n] = chain_total;
  if (timing)
    fprintf(stderr, "dpc_solve n=%d chunks=%d x %d devices=%d drivers=%d workers=%d slots=%d pinned(in %d res %d pairs %d): %.1f ms | %lld pipeline (%d expanded by the host) + %lld host-half chunks | worker thread-sum scan %.1f expand/hooks %.1f host-half %.1f ms\n",
            n, nchunks, chunk, ndev, ndrv, nworkers, nslots, (int)hp_pinned, (int)res_pinned, (int)direct, (now_s() - t0) * 1e3,
            (long long)c->solve_pipe_chunks.load(), host_routed.load(), (long long)c->solve_host_chunks.load(), t_scan / 1e6, t_work / 1e6, t_legacy / 1e6);
  return DPC_OK;
}

int dpc_relaunch(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  Engine &e = c->main;
  if (!e.flushed || !e.waited) return DPC_ERR_STATE;
  if (e.batch.dprobs.empty()) return 0;
  CK(cudaSetDevice(e.device));
  int rc = e.launch_all();
  if (rc != DPC_OK) return rc;
  return e.nlaunch;
}

void *dpc_stream(dpc_ctx_t *c) { return c ? (void *)c->main.stream : NULL; }

int dpc_sync(dpc_ctx_t *c) {
  if (!c) return DPC_ERR_ARG;
  Engine &e = c->main;
  CK(cudaSetDevice(e.device));
  CK(cudaStreamSynchronize(e.stream));
  if (e.flushed && !e.batch.dprobs.empty()) CK(cudaEventElapsedTime(&e.ms_total, e.ev0, e.ev1));
  return DPC_OK;
}

int dpc_measure_int_peak(int device, double *alu_gops, double *mix_gops) {
  if (!alu_gops || !mix_gops || device < 0 || device >= dpc_device_count()) return DPC_ERR_ARG;
  CK(cudaSetDevice(device));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, device));
  const int grid = p.multiProcessorCount * 8;
  int *d = NULL;
  CK(cudaMalloc(&d, (size_t)grid * 256 * sizeof(int)));
  int rc = int_peak_run<0>(d, grid, alu_gops);
  if (rc == DPC_OK) rc = int_peak_run<1>(d, grid, mix_gops);
  cudaFree(d);
  return rc;
}

int dpc_last_kernel_ms(dpc_ctx_t *c, float *ms) {
  if (!c || !ms) return DPC_ERR_ARG;
  if (!c->main.flushed) return DPC_ERR_STATE;
  *ms = c->main.ms_total;
  return DPC_OK;
}

static void add_stats(const Engine &e, dpc_stats_t *out) {
  const Batch &b = e.batch;
  out->nproblems += (int64_t)b.probs.size();
  for (size_t i = 0; i < b.dprobs.size(); i++) {
    const DevProb &p = b.dprobs[i];
    if ((p.kind == DPC_END5_GAP || p.kind == DPC_END3_GAP) && p.endalign == DPC_QUERYEND_NOGAPS) {
      out->fill_bytes += (int64_t)sizeof(DevProb) + 4 + p.L1 + (p.L2 + 3) / 4 + (int64_t)sizeof(DevRes) +
                         (p.gout != DPC_NO_GOUT ? p.L2 : 0);
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
      /* algorithmic HBM bytes: query bytes + 2-bit genome in, staged genome characters out */
      out->fill_bytes += (p.kind == DPC_CDNA_GAP) ? d.cols + (d.rows + 3) / 4 : d.rows + (d.cols + 3) / 4;
      if (p.gout != DPC_NO_GOUT) out->fill_bytes += d.cols;
    }
    out->fill_bytes += (int64_t)sizeof(DevProb) + 4 + (int64_t)sizeof(DevRes);   /* descriptor + list entry in, result out */
  }
  out->h2d_bytes += e.h2d_bytes; out->d2h_bytes += e.d2h_bytes;
  out->launches += e.nlaunch;
}

/* Stats of the batch held by the ticket-API engine, or (after dpc_solve) of the chunks of its last wave. */
int dpc_get_stats(dpc_ctx_t *c, dpc_stats_t *out) {
  if (!c || !out) return DPC_ERR_ARG;
  memset(out, 0, sizeof *out);
  if (!c->main.batch.probs.empty()) add_stats(c->main, out);
  else {
    /* after dpc_solve: cells / bytes of the chunks still held by the engines, transfer totals of the whole call */
    for (size_t d = 0; d < c->engines.size(); d++) for (size_t i = 0; i < c->engines[d].size(); i++) add_stats(*c->engines[d][i], out);
    out->nproblems = c->solve_problems;
    out->h2d_bytes = c->solve_h2d; out->d2h_bytes = c->solve_d2h; out->launches = (int32_t)c->solve_launches;
    out->pipeline_chunks = c->solve_pipe_chunks; out->host_chunks = (int32_t)c->solve_host_chunks;
  }
  return DPC_OK;
}

}  /* extern "C" */
