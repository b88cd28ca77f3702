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
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "dpc_host.h"
#include "dpc_rows.h"

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
};

/* resident blocks per SM the compiler must leave room for: the one-matrix kernels fit DPC_MIN_BLOCKS_1M x 256
 * threads (3 -> 80 registers); their narrow-band instantiation (1 or 2 diagonals per lane only) is asked for
 * DPC_MIN_BLOCKS_NARROW; the two-matrix kernels keep 2 (128 registers) */
#ifndef DPC_MIN_BLOCKS_1M
#define DPC_MIN_BLOCKS_1M 3
#endif
#ifndef DPC_MIN_BLOCKS_NARROW
#define DPC_MIN_BLOCKS_NARROW 4
#endif
/* Launch variants (one kernel instantiation each, so that each carries only its own code and address spaces):
 *   V_NARROW  arenas in shared memory, bulk region in the arena, bands of at most 64 diagonals (RowFillT<2>)
 *   V_WIDE    arenas in shared memory, bulk region in the arena, any band
 *   V_SPILL   small region in shared memory, bulk region (direction planes, nogap bands, ops) in HBM scratch
 *   V_HBM     everything in HBM scratch
 * KG: kind group (0 one-matrix solvers, 1 genome gap, 2 cDNA gap); GEN: every matrix through the memory-state fill
 * (test hook; not instantiated for V_NARROW). */
enum { V_NARROW = 0, V_WIDE = 1, V_SPILL = 2, V_HBM = 3, NVARIANT = 4 };
template <int V, int KG, bool GEN>
__global__ void __launch_bounds__(256, (KG == 0 ? (V == V_NARROW ? DPC_MIN_BLOCKS_NARROW : DPC_MIN_BLOCKS_1M) : 2)) dpc_solve_kernel(const KernelArgs a) {
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
  /* work items are claimed one ahead, so the next problem's descriptor is already on its way
     from HBM while the current problem is being solved (the list is ordered by work, not by address) */
  int nxt = 0;
  if (ln.lane == 0) nxt = (int)atomicAdd(a.counter, 1u);
  nxt = __shfl_sync(0xffffffffu, nxt, 0);
  for (;;) {
    const int i = nxt;
    if (i >= a.n) break;
    const uint32_t pi = a.list[i];
    if (ln.lane == 0) nxt = (int)atomicAdd(a.counter, 1u);
    nxt = __shfl_sync(0xffffffffu, nxt, 0);
    if (nxt < a.n) {
      const uint32_t pn = a.list[nxt];
      const char *d = (const char *)&a.probs[pn];
      if (ln.lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(d + 64 * ln.lane));
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

typedef void (*kernel_fn)(const KernelArgs);
/* [variant][kind group][generic]; the generic test hook runs the narrow class in the V_WIDE instantiation */
static kernel_fn kernel_of(int v, int kg, bool gen) {
  static const kernel_fn tab[NVARIANT][3][2] = {
    { { dpc_solve_kernel<V_NARROW, 0, false>, dpc_solve_kernel<V_WIDE, 0, true> },
      { dpc_solve_kernel<V_NARROW, 1, false>, dpc_solve_kernel<V_WIDE, 1, true> },
      { dpc_solve_kernel<V_NARROW, 2, false>, dpc_solve_kernel<V_WIDE, 2, true> } },
    { { dpc_solve_kernel<V_WIDE, 0, false>, dpc_solve_kernel<V_WIDE, 0, true> },
      { dpc_solve_kernel<V_WIDE, 1, false>, dpc_solve_kernel<V_WIDE, 1, true> },
      { dpc_solve_kernel<V_WIDE, 2, false>, dpc_solve_kernel<V_WIDE, 2, true> } },
    { { dpc_solve_kernel<V_SPILL, 0, false>, dpc_solve_kernel<V_SPILL, 0, true> },
      { dpc_solve_kernel<V_SPILL, 1, false>, dpc_solve_kernel<V_SPILL, 1, true> },
      { dpc_solve_kernel<V_SPILL, 2, false>, dpc_solve_kernel<V_SPILL, 2, true> } },
    { { dpc_solve_kernel<V_HBM, 0, false>, dpc_solve_kernel<V_HBM, 0, true> },
      { dpc_solve_kernel<V_HBM, 1, false>, dpc_solve_kernel<V_HBM, 1, true> },
      { dpc_solve_kernel<V_HBM, 2, false>, dpc_solve_kernel<V_HBM, 2, true> } } };
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
  for (int v = 0; v < V_HBM; v++)
    for (int kg = 0; kg < 3; kg++)
      for (int gen = 0; gen < 2; gen++)
        CK(cudaFuncSetAttribute((const void *)kernel_of(v, kg, gen != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
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

struct ClassLaunch {
  int variant;                     /* V_* */
  int kg;
  uint32_t arena_bytes;
  int wpb;
  size_t list_off;
  int n;
};

/* Launch classes.  Two shared-memory arena sizes: 6 KB per warp leaves most of the 228 KB to L1 (descriptors,
 * genome blocks and query bytes are read through it), 13 KB takes what is bigger (the register file allows at most
 * 2-4 blocks of 8 warps per SM, so 13 KB per warp never limits occupancy below that of the two-matrix kernels).  A
 * problem whose bulk region does not fit keeps it in HBM scratch (V_SPILL); a problem whose small region alone does
 * not fit runs entirely from HBM scratch (V_HBM).  Narrow problems (every band <= 64 diagonals) get the narrow
 * instantiation. */
#define NCLASS 6
static const int k_class_variant[NCLASS] = { V_NARROW, V_NARROW, V_WIDE, V_WIDE, V_SPILL, V_HBM };
static uint32_t k_class_bytes[NCLASS] = { 6 << 10, 13 << 10, 6 << 10, 13 << 10, 13 << 10, 0 };
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
    live = true;
    return DPC_OK;
  }
  void close() {
    if (!live) return;
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
    d_probs.release(); d_pool.release(); d_scratch.release(); d_gout.release(); d_res.release(); d_list.release();
    d_ovf.release(); d_counters.release();
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
      const int threads = L.wpb * 32;
      const size_t smem = L.variant != V_HBM ? (size_t)L.wpb * L.arena_bytes : 0;
      int per_sm = 1;
      const kernel_fn fn = kernel_of(L.variant, L.kg, fill_gen != 0);
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, threads, smem));
      if (per_sm < 1) per_sm = 1;
      int grid = (L.n + L.wpb - 1) / L.wpb;
      if (grid > d.sm_count * per_sm) grid = d.sm_count * per_sm;
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
    if (n == 0) return DPC_OK;

    fill_gen = g_force_generic;                       /* one snapshot per batch: layout and kernels must agree */
    const int with_state = fill_gen ? 1 : 2;          /* must match FILL::fillmode of the kernel */
    /* class (shared-memory arena or HBM only) and a work bucket per problem: within a class the list is ordered
       by descending work so that the long problems start first and the tail of the launch is made of short ones */
    cls.resize(n);
    size_t count[NCLASS * 3 * NBUCKET] = { 0 };
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
      const int kg = p.kind == DPC_GENOME_GAP ? 1 : p.kind == DPC_CDNA_GAP ? 2 : 0;
      cls[i] = (uint16_t)((k * 3 + kg) * NBUCKET + bucket);
      count[cls[i]]++;
    }
    list.clear();
    list.grow(n);
    size_t off[NCLASS * 3 * NBUCKET], at = 0;
    for (int k = 0; k < NCLASS * 3 * NBUCKET; k++) { off[k] = at; at += count[k]; }
    {
      size_t cur[NCLASS * 3 * NBUCKET];
      for (int k = 0; k < NCLASS * 3 * NBUCKET; k++) cur[k] = off[k];
      for (size_t i = 0; i < n; i++) list[cur[cls[i]]++] = (uint32_t)i;
    }
    for (int k = 0; k < NCLASS * 3; k++) {       /* one launch per (arena class, kind group) that has work */
      size_t cnt = 0;
      for (int q = 0; q < NBUCKET; q++) cnt += count[k * NBUCKET + q];
      if (!cnt) continue;
      ClassLaunch L;
      L.variant = k_class_variant[k / 3];
      L.kg = k % 3;
      L.arena_bytes = k_class_bytes[k / 3];
      L.wpb = 8;
      L.list_off = off[k * NBUCKET]; L.n = (int)cnt;
      launches.push_back(L);
    }
    /* ops overflow arena: problems whose worst case exceeds the inline slots (bounded) */
    if (ovf_worst > (1ull << 30)) ovf_worst = 1ull << 30;
    ovf_cap = (size_t)ovf_worst + 64;

    b.pool_align(16);
    int rc;
    if ((rc = d_probs.need(n)) || (rc = d_pool.need(b.pool.size())) || (rc = d_res.need(n)) ||
        (rc = d_list.need(n)) || (rc = d_ovf.need(ovf_cap)) || (rc = d_counters.need(NCLASS * 3 + 2)) ||
        (rc = d_scratch.need((size_t)scratch_total + 16)) || (rc = d_gout.need((size_t)b.gout_total + 64)))
      return rc;
    h_gout.clear(); h_gout.grow((size_t)b.gout_total + 64);
    b.gout_host = NULL;
    h_res.clear(); h_res.grow(n);
    h_counters.clear(); h_counters.grow(NCLASS * 3 + 2);
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

struct dpc_ctx {
  Engine main;                       /* ticket API, dpc_relaunch */
  std::vector<Engine *> subs;        /* bulk API: one engine per chunk in flight */
  Workers *workers;
  int nthreads;
  std::atomic<int64_t> solve_h2d, solve_d2h, solve_launches, solve_problems;   /* totals of the last dpc_solve */
  dpc_ctx() : workers(NULL), nthreads(1), solve_h2d(0), solve_d2h(0), solve_launches(0), solve_problems(0) {}
};

/* ---- C ABI ---------------------------------------------------------------------------------- */
extern "C" {

int dpc_init(int maxlookback, int extraquerygap, int maxpeelback, int extramaterial_end, int extramaterial_paired, int mode) {
  std::lock_guard<std::mutex> lock(g_mu);
  int rc = host_init(maxlookback, extraquerygap, maxpeelback, extramaterial_end, extramaterial_paired, mode);
  g_version++;
  const char *kb = getenv("DPC_CLASS0_KB");          /* tuning aid: size of the small shared-memory class */
  if (kb && atoi(kb) >= 1 && atoi(kb) <= 13) k_class_bytes[0] = k_class_bytes[2] = (uint32_t)atoi(kb) << 10;
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
  return c;
}

void dpc_ctx_free(dpc_ctx_t *c) {
  if (!c) return;
  delete c->workers;
  for (size_t i = 0; i < c->subs.size(); i++) delete c->subs[i];
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

/* Bulk call: the problems are cut into chunks; host threads pack a chunk, queue its copies and kernels on
 * the chunk's own stream, and finalise it when it is back, so packing, PCIe traffic, kernels and Pair
 * rebuild of different chunks overlap.  Results and pairs come out in input order. */
static double now_s() {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

int dpc_solve(dpc_ctx_t *c, const dpc_problem_t *problems, int n, dpc_result_t *results,
              dpc_pair_t *pairs, int64_t pair_cap, int64_t *pair_off) {
  if (!c || n < 0 || (n > 0 && (!problems || !results))) return DPC_ERR_ARG;
  int rc = ensure_device(c->main.device);
  if (rc != DPC_OK) return rc;
  const int device = c->main.device;
  const int T = c->nthreads;
  if (!c->workers) c->workers = new Workers(T);
  int chunk = n / (4 * T) + 1;
  if (chunk < 2048) chunk = 2048;
  if (chunk > 16384) chunk = 16384;
  { static const int forced = getenv("DPC_CHUNK") ? atoi(getenv("DPC_CHUNK")) : 0; if (forced >= 256) chunk = forced; }   /* tuning aid */
  const int nchunks = (n + chunk - 1) / chunk;
  const int nengines = std::min(nchunks, 4 * T);              /* chunks in flight */
  while ((int)c->subs.size() < nengines) {
    Engine *e = new Engine();
    if (e->open(device) != DPC_OK) { delete e; return DPC_ERR_CUDA; }
    c->subs.push_back(e);
  }
  static const bool timing = getenv("DPC_TIMING") != NULL;
  static const bool stream_pairs = getenv("DPC_NO_STREAM") == NULL;     /* non-temporal stores for the pair records */
  std::atomic<int> err(0);
  std::atomic<int64_t> t_pack(0), t_flush(0), t_wait(0), t_fin(0), t_pairs(0), t_stall(0);
  /* chunk j publishes the end offset of its pair block once every chunk before it has; a chunk's engine is
   * free again when the chunk's pairs are out */
  std::vector<std::atomic<int64_t>> chunk_end((size_t)nchunks);
  std::vector<std::atomic<int>> chunk_done((size_t)nchunks);
  for (int j = 0; j < nchunks; j++) { chunk_end[(size_t)j].store(-1); chunk_done[(size_t)j].store(0); }
  const double t0 = now_s();
  c->solve_h2d = 0; c->solve_d2h = 0; c->solve_launches = 0; c->solve_problems = n;
  std::atomic<int> next_chunk(0);
  /* first half of a chunk: pack and queue copies + kernels on the chunk's stream (returns at once) */
  auto launch = [&](int j) {
    Engine &e = *c->subs[(size_t)(j % nengines)];
    const int lo = j * chunk, cnt = std::min(chunk, n - lo);
    double a0 = timing ? now_s() : 0;
    if (j >= nengines) while (!chunk_done[(size_t)(j - nengines)].load(std::memory_order_acquire)) std::this_thread::yield();
    if (err.load()) return;
    int r = 0;
    try {
      e.reset();
      r = e.batch.add_ext(problems + lo, results + lo, cnt);
      double a1 = timing ? now_s() : 0;
      if (r >= 0) r = e.flush();
      if (timing) { t_pack += (int64_t)((a1 - a0) * 1e9); t_flush += (int64_t)((now_s() - a1) * 1e9); }
    } catch (const std::bad_alloc &) { r = DPC_ERR_NOMEM; }
    if (r < 0) { int z = 0; err.compare_exchange_strong(z, r); }
  };
  /* second half: wait for the device, finalise, take the pair offset from the previous chunk, rebuild the pairs */
  auto finish = [&](int j) {
    Engine &e = *c->subs[(size_t)(j % nengines)];
    const int lo = j * chunk, cnt = std::min(chunk, n - lo);
    int64_t mine = 0;
    double a2 = timing ? now_s() : 0, a3 = 0, a4 = 0;
    if (!err.load()) {
      int r = 0;
      try {
        r = e.wait();
        if (r >= 0) for (int i = 0; i < cnt; i++) mine += results[lo + i].npairs;
      } catch (const std::bad_alloc &) { r = DPC_ERR_NOMEM; }
      if (r < 0) { int z = 0; err.compare_exchange_strong(z, r); }
    }
    if (timing) a3 = now_s();
    int64_t at = 0;
    if (j > 0) {
      int64_t prev;
      while ((prev = chunk_end[(size_t)(j - 1)].load(std::memory_order_acquire)) < 0) std::this_thread::yield();
      at = prev;
    }
    chunk_end[(size_t)j].store(at + mine, std::memory_order_release);
    if (timing) a4 = now_s();
    if (!err.load() && (pairs || pair_off)) {
      if (pairs && at + mine > pair_cap) { int z = 0; err.compare_exchange_strong(z, DPC_ERR_NOMEM); }
      else {
        try {
          for (int i = 0; i < cnt; i++) {
            const int np = results[lo + i].npairs;
            if (pairs && i + 8 < cnt) e.batch.prefetch_genome(i + 8);
            if (pair_off) pair_off[lo + i] = at;
            if (!pairs || np == 0) { at += np; continue; }
            int k = e.pairs_into(i, pairs + at, stream_pairs);
            if (k != np) { int z = 0; err.compare_exchange_strong(z, DPC_ERR_STATE); break; }
            at += k;
          }
        } catch (const std::bad_alloc &) { int z = 0; err.compare_exchange_strong(z, DPC_ERR_NOMEM); }
      }
    }
    c->solve_h2d += e.h2d_bytes; c->solve_d2h += e.d2h_bytes; c->solve_launches += e.nlaunch;
    chunk_done[(size_t)j].store(1, std::memory_order_release);
    if (timing) {
      t_wait += (int64_t)((a3 - a2) * 1e9); t_fin += (int64_t)(e.t_finalize * 1e9);
      t_stall += (int64_t)((a4 - a3) * 1e9); t_pairs += (int64_t)((now_s() - a4) * 1e9);
    }
  };
  /* Launch tickets and finish tickets are handed out separately, both in chunk order; a thread launches one chunk
     (two before its first finish, so that every thread keeps two in flight) and then finishes the OLDEST chunk
     nobody has taken yet -- not necessarily one it launched.  Finishing in order keeps the ordered hand-off of the
     pair offsets short: chunk j-1 is always in somebody's hands before chunk j is.  (No deadlock: a launch only
     waits for the engine of chunk l - nengines, whose finish ticket is taken by then because nengines = 4T and a
     thread is never more than two launches ahead of its finishes.) */
  std::vector<std::atomic<int>> launched((size_t)nchunks);
  for (int j = 0; j < nchunks; j++) launched[(size_t)j].store(0);
  std::atomic<int> next_finish(0);
  c->workers->run(T, [&](int) {
    cudaSetDevice(device);
    auto try_launch = [&]() {
      const int l = next_chunk.fetch_add(1);
      if (l < nchunks) { launch(l); launched[(size_t)l].store(1, std::memory_order_release); }
    };
    try_launch();
    for (;;) {
      try_launch();
      const int f = next_finish.fetch_add(1);
      if (f >= nchunks) break;
      while (!launched[(size_t)f].load(std::memory_order_acquire)) std::this_thread::yield();
      finish(f);
    }
  });
#if defined(__SSE2__) && defined(__x86_64__)
  _mm_sfence();
#endif
  if (err.load()) return err.load();
  if (pair_off) pair_off[n] = nchunks ? chunk_end[(size_t)nchunks - 1].load() : 0;
#ifdef DPC_PROFILE_REBUILD
  {
    unsigned long long cg = 0, cr = 0;
    for (size_t i = 0; i < c->subs.size(); i++) { cg += c->subs[i]->scratch.cyc_gather; cr += c->subs[i]->scratch.cyc_replay; c->subs[i]->scratch.cyc_gather = c->subs[i]->scratch.cyc_replay = 0; }
    fprintf(stderr, "rebuild cycles (thread-sum, tsc): gather %.1f M, replay %.1f M\n", cg / 1e6, cr / 1e6);
  }
#endif
  if (timing)
    fprintf(stderr, "dpc_solve n=%d chunks=%d x %d threads=%d engines=%d: %.1f ms | thread-sum pack %.1f flush %.1f wait %.1f (finalize %.1f) stall %.1f pairs %.1f ms\n",
            n, nchunks, chunk, T, nengines, (now_s() - t0) * 1e3, t_pack / 1e6, t_flush / 1e6, t_wait / 1e6, t_fin / 1e6, t_stall / 1e6, t_pairs / 1e6);
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

int dpc_last_kernel_ms(dpc_ctx_t *c, float ms[3]) {
  if (!c || !ms) return DPC_ERR_ARG;
  if (!c->main.flushed) return DPC_ERR_STATE;
  ms[0] = c->main.ms_total; ms[1] = 0.0f; ms[2] = c->main.ms_total;   /* fill, bridge and traceback are one fused kernel */
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
    for (size_t i = 0; i < c->subs.size(); i++) add_stats(*c->subs[i], out);
    out->nproblems = c->solve_problems;
    out->h2d_bytes = c->solve_h2d; out->d2h_bytes = c->solve_d2h; out->launches = (int32_t)c->solve_launches;
  }
  return DPC_OK;
}

}  /* extern "C" */
