/* dpc_host.h -- host side of libdynprog_cuda: score tables, problem packing, result
 * finalisation and the rebuild of Pair records from device traceback runs.
 *
 * Pure host C++ (no CUDA calls): dynprog_cuda.cu includes it next to the kernels;
 * tests/emul includes it next to a single-lane build of dpc_core.h.  Nothing in here
 * computes a DP matrix -- that only exists as device code.
 *
 * Reference: src/dynprog.c of GMAP/GSNAP 2012-07-03 (line numbers cited per function).
 */
#ifndef DPC_HOST_H
#define DPC_HOST_H

#include <stdint.h>
#include <string.h>
#include <ctype.h>
#include <stdlib.h>
#if defined(__SSE2__) && defined(__x86_64__)
#include <immintrin.h>
#define DPC_X86_SIMD 1
#endif
#include <math.h>
#include <algorithm>
#include <functional>
#include <new>
#include <vector>
#include "../../include/dynprog_cuda.h"
#include "dpc_core.h"

namespace dpc {

enum { HIGHQ = 0, MEDQ = 1, LOWQ = 2, ENDQ = 3 };   /* Mismatchtype_T, dynprog.c:150 */

struct Globals {
  bool inited, setup_done;
  int maxlength1, maxlength2;
  int mode;
  int P[4][128][128];            /* pairdistance_array, dynprog.c:1045 */
  uint8_t CONS[128][128];        /* consistent_array, dynprog.c:1046 */
  bool acgt_plain;               /* no two different upper-case bases are "consistent" (false in the CMET modes, 1215-1219) */
  DevTables tables;
  dpc_setup_t setup;
  uint64_t genome_nbases;
};
inline Globals &G() { static Globals g; return g; }

/* permute_cases / permute_cases_oneway, dynprog.c:1053-1124 */
inline void set_pair(Globals &g, int a, int b, int score, bool oneway) {
  const int A[2] = { a, tolower(a) }, B[2] = { b, tolower(b) };
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) {
      g.CONS[A[i]][B[j]] = 1;
      for (int t = 0; t < 4; t++) g.P[t][A[i]][B[j]] = score;
      if (!oneway) {
        g.CONS[B[j]][A[i]] = 1;
        for (int t = 0; t < 4; t++) g.P[t][B[j]][A[i]] = score;
      }
    }
}

/* pairdistance_init, dynprog.c:1127-1226.  The order of the assignments matters (later ones win). */
inline void build_tables(Globals &g, int mode) {
  static const int mismatch[4] = { -3, -2, -1, -5 };                 /* dynprog.c:169-179 */
  static const char *const half[] = { "RA", "RG", "YT", "YC", "WA", "WT", "SG", "SC", "MA", "MC", "KG", "KT" };
  static const char *const ambig[] = { "HA", "HT", "HC", "BG", "BC", "BT", "VG", "VA", "VC", "DG", "DA", "DT",
                                       "NT", "NC", "NA", "NG", "XT", "XC", "XA", "XG" };
  memset(g.P, 0, sizeof g.P);
  memset(g.CONS, 0, sizeof g.CONS);
  for (int c1 = 'A'; c1 <= 'z'; c1++)
    for (int c2 = 'A'; c2 < 'z'; c2++)                               /* sic: 'z' itself is excluded, 1152 */
      for (int t = 0; t < 4; t++) g.P[t][c1][c2] = mismatch[t];
  set_pair(g, 'U', 'T', 3, false);
  for (unsigned i = 0; i < sizeof half / sizeof *half; i++) set_pair(g, half[i][0], half[i][1], 1, false);
  for (unsigned i = 0; i < sizeof ambig / sizeof *ambig; i++) set_pair(g, ambig[i][0], ambig[i][1], -1, false);
  if (mode == DPC_MODE_CMET_STRANDED || mode == DPC_MODE_CMET_NONSTRANDED) {   /* 1215-1219 */
    set_pair(g, 'T', 'C', 3, true);
    set_pair(g, 'A', 'G', 3, true);
  }
  for (int c = 'A'; c < 'Z'; c++) set_pair(g, c, c, 3, false);       /* sic: 'Z' excluded, 1221 */

  g.acgt_plain = true;
  for (const char *x = "ACGT"; *x; x++)
    for (const char *y = "ACGT"; *y; y++)
      if (*x != *y && g.CONS[(int)*x][(int)*y]) g.acgt_plain = false;

  static const char codes[6] = { 'A', 'C', 'G', 'T', 'N', '*' };
  memset(&g.tables, 0, sizeof g.tables);
  for (int t = 0; t < 4; t++)
    for (int q = 0; q < 128; q++)
      for (int c = 0; c < 6; c++) g.tables.score[t][q][c] = (int8_t)g.P[t][q][(int)codes[c]];
  for (int q = 0; q < 128; q++)
    for (int c = 0; c < 6; c++) {
      if (g.CONS[q][(int)codes[c]]) g.tables.cons[q] |= (uint8_t)(1u << c);
      if (g.CONS[(int)codes[c]][q]) g.tables.consT[q] |= (uint8_t)(1u << c);
    }
}

inline int host_init(int maxlookback, int extraquerygap, int maxpeelback, int extramaterial_end,
                     int extramaterial_paired, int mode) {
  Globals &g = G();
  /* compute_maxlengths, dynprog.c:831-852 */
  int m1 = maxlookback + maxpeelback;
  if (m1 < 500) m1 = 500;
  int m2 = m1 + extraquerygap + (extramaterial_end > extramaterial_paired ? extramaterial_end : extramaterial_paired);
  if (m2 < 2000) m2 = 2000;
  if (m1 > 1300 || m2 > 3000) return DPC_ERR_ARG;   /* nogap bands are kept in 16 bits (worst real score -7*3000 - 5*1300),
                                                        traceback ops carry 14-bit lengths, bridge keys 13-bit columns */
  g.maxlength1 = m1; g.maxlength2 = m2; g.mode = mode;
  build_tables(g, mode);
  g.inited = true;
  return DPC_OK;
}

/* ---- host view of the genome (for the rebuilt pairs' genome characters) ------------------ */
inline char host_genome_char(const uint32_t *blocks, uint32_t pos) {   /* genome.c:9325-9362 */
  return "ACGTN"[dpc_genome_code(blocks, pos)];
}
inline bool allstar(const dpc_problem_t &p) {                           /* dynprog.c:415-419 */
  uint32_t pos = p.chroffset + p.chrpos;
  return pos < p.chroffset || pos >= p.chrhigh;
}
inline char host_genomic_nt(const dpc_problem_t &p, int genomicpos) {   /* get_genomic_nt, dynprog.c:403-441 */
  static const char compl_nt[6] = { 'T', 'G', 'C', 'A', 'N', '*' };
  if (genomicpos < 0 || (uint32_t)genomicpos >= p.genomiclength || allstar(p)) return '*';
  const uint32_t *blocks = G().setup.genome_blocks;
  if (p.watsonp) return host_genome_char(blocks, p.chroffset + p.chrpos + (uint32_t)genomicpos);
  return compl_nt[dpc_genome_code(blocks, p.chroffset + p.chrpos + (p.genomiclength - 1) - (uint32_t)genomicpos)];
}

/* four bases per table entry: index = one byte of the 2-bit words (base i in bits 2i..2i+1);
 * [0] ACGT in ascending order, [1] the same four reversed, [2] complemented ascending, [3] complemented reversed */
struct NtLut {
  uint32_t t[4][256];
  NtLut() {
    static const char f[4] = { 'A', 'C', 'G', 'T' }, c[4] = { 'T', 'G', 'C', 'A' };
    for (int b = 0; b < 256; b++) {
      uint32_t a = 0, ar = 0, k = 0, kr = 0;
      for (int i = 0; i < 4; i++) {
        int code = (b >> (2 * i)) & 3;
        a |= (uint32_t)(uint8_t)f[code] << (8 * i);  ar |= (uint32_t)(uint8_t)f[code] << (8 * (3 - i));
        k |= (uint32_t)(uint8_t)c[code] << (8 * i);  kr |= (uint32_t)(uint8_t)c[code] << (8 * (3 - i));
      }
      t[0][b] = a; t[1][b] = ar; t[2][b] = k; t[3][b] = kr;
    }
  }
};
inline const NtLut &nt_lut() { static const NtLut l; return l; }

/* gathers get_genomic_nt(start +/- k) for k = 0..len-1 (dynprog.c:403-441): positions outside the segment give
 * '*'; inside, the bases come out of the (high, low, flags) blocks of genome.c:9325-9362, four at a time through a
 * byte table where a block has no N flags */
inline void gather_genome(const dpc_problem_t &p, const uint32_t *blocks, int start, int len, bool rev, char *out) {
  static const char fwd_nt[4] = { 'A', 'C', 'G', 'T' }, compl_nt[4] = { 'T', 'G', 'C', 'A' };
  if (len <= 0) return;
  const int lo = rev ? start - (len - 1) : start;          /* segment positions covered, ascending: [lo, lo+len) */
  if (allstar(p)) { memset(out, '*', (size_t)len); return; }
  const int64_t glen = p.genomiclength;
  const uint32_t base = p.chroffset + p.chrpos;
  const bool watson = p.watsonp != 0;
  const char *nt = watson ? fwd_nt : compl_nt;
  const NtLut &lut = nt_lut();
  /* bases come out in ascending genome order on Watson and descending on Crick; the output index ascends unless
     rev: the table gives the four characters of a byte in output order */
  const bool out_desc = rev;
  const bool bit_desc = !watson;
  const uint32_t *tab = lut.t[(watson ? 0 : 2) + ((out_desc != bit_desc) ? 1 : 0)];
  for (int pos = lo; pos < lo + len;) {
    const int oi = rev ? start - pos : pos - start;
    if (pos < 0 || pos >= glen) { out[oi] = '*'; pos++; continue; }
    const uint32_t g = watson ? base + (uint32_t)pos : base + (uint32_t)(glen - 1 - pos);   /* absolute genome position */
    const uint32_t *b = blocks + (uint64_t)(g >> 5) * 3;
    const uint64_t bits = ((uint64_t)b[0] << 32) | b[1];
    const uint32_t flags = b[2];
    int bit = (int)(g & 31);
    int run = watson ? 32 - bit : bit + 1;                  /* bases of this block on the way */
    if (run > lo + len - pos) run = lo + len - pos;
    if ((int64_t)pos + run > glen) run = (int)(glen - pos);
    const int od = rev ? -1 : 1, bd = watson ? 1 : -1;
    int o = oi, k = 0;
    if (flags == 0) {
      /* scalar until the bit index reaches a byte boundary in the direction of travel, then four at a time */
      while (k < run && (watson ? (bit & 3) != 0 : (bit & 3) != 3)) { out[o] = nt[(bits >> (2 * bit)) & 3u]; k++; bit += bd; o += od; }
      while (run - k >= 4) {
        const int byte = watson ? bit >> 2 : (bit - 3) >> 2;
        const uint32_t four = tab[(bits >> (8 * byte)) & 0xffu];
        memcpy(rev ? out + o - 3 : out + o, &four, 4);
        k += 4; bit += 4 * bd; o += 4 * od;
      }
    }
    for (; k < run; k++, bit += bd, o += od)
      out[o] = ((flags >> bit) & 1u) ? 'N' : nt[(bits >> (2 * bit)) & 3u];
    pos += run;
  }
}

/* ---- growable buffers whose storage the CUDA side can make page-locked ------------------------ */
struct Alloc { void *(*alloc)(size_t); void (*release)(void *); };
inline Alloc default_alloc() { Alloc a = { malloc, free }; return a; }

template <class T> struct PBuf {
  T *p; size_t n, cap; Alloc al;
  PBuf() : p(NULL), n(0), cap(0), al(default_alloc()) {}
  ~PBuf() { if (p) al.release(p); }
  void set_alloc(Alloc a) { if (p) al.release(p); p = NULL; n = cap = 0; al = a; }
  void reserve(size_t want) {
    if (want <= cap) return;
    size_t nc = cap * 2 > want ? cap * 2 : want;
    if (nc < 1024) nc = 1024;
    T *q = (T *)al.alloc(nc * sizeof(T));
    if (!q) throw std::bad_alloc();
    if (n) memcpy(q, p, n * sizeof(T));
    if (p) al.release(p);
    p = q; cap = nc;
  }
  void clear() { n = 0; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T *data() { return p; }
  const T *data() const { return p; }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
  void push_back(const T &v) { if (n == cap) reserve(n + 1); p[n++] = v; }
  T *grow(size_t k) { reserve(n + k); T *r = p + n; n += k; return r; }
 private:
  PBuf(const PBuf &); PBuf &operator=(const PBuf &);
};

/* ---- batch ---------------------------------------------------------------------------------- */
/* Dynprog_microexon_int (dynprog.c:7127-7429) on the host side of the boundary: everything but the exact-match scans */
struct MicroCand { int cL, cR, mid, textleft, query; };      /* one (cL, cR) pair with splice-site dinucleotides; query = its scan, or -1 */
struct MicroProb {
  int problem;                       /* ticket */
  std::vector<MicroCand> cands;      /* in the reference's loop order */
  std::vector<dpc_pair_t> pairs;     /* the returned list, head first (filled by finalize_micro) */
};

struct HostProb {
  uint32_t q0, q1;      /* pool offsets: first byte of the copied span(s) */
  uint32_t aux;         /* pool offset of probability / known-site arrays */
  int32_t dev;          /* index into the device arrays, -1 when resolved on the host (early returns) */
  int32_t L1, L2;       /* lengths after clipping (end gaps) */
  int32_t micro;        /* index into Batch::micro (Dynprog_microexon_int), else -1 */
  uint32_t gout;        /* offset of the device-staged genome characters in the returned byte stream, or DPC_NO_GOUT */
};

/* pair records are written through a bare cursor: every caller sizes the destination first */
struct Out {
  dpc_pair_t *p; int n;
  bool stream;
  Out() : p(NULL), n(0), stream(false) {}
  void push(int qpos, int gpos, char cdna, char comp, char genome, int idx, int gapp) {
    /* dpc_pair_t is 16 bytes: {querypos, genomepos} and {dynprogindex, cdna, comp, genome, gapp}: two 8-byte stores */
    uint64_t lo = (uint32_t)qpos | ((uint64_t)(uint32_t)gpos << 32);
    uint64_t hi = (uint32_t)idx | ((uint64_t)(uint8_t)cdna << 32) | ((uint64_t)(uint8_t)comp << 40) |
                  ((uint64_t)(uint8_t)genome << 48) | ((uint64_t)(uint8_t)gapp << 56);
    uint64_t *w = (uint64_t *)&p[n++];
#if defined(__SSE2__) && defined(__x86_64__)
    if (stream) {        /* the caller's big array: written once, not read back here -- go around the cache */
      _mm_stream_si64((long long *)w, (long long)lo);
      _mm_stream_si64((long long *)w + 1, (long long)hi);
      return;
    }
#endif
    w[0] = lo; w[1] = hi;
  }
  void push_gapholder() { push(-1, -1, ' ', ' ', ' ', 0, 1); }       /* pairpool.c:352-401 */
};

/* per-thread scratch of the rebuild */
struct Scratch {
  std::vector<char> qa, ga, qb, gb;
  std::vector<dpc_pair_t> sL, sR, out;
  unsigned long long cyc_gather, cyc_replay;      /* only maintained with -DDPC_PROFILE_REBUILD */
  Scratch() : cyc_gather(0), cyc_replay(0) {}
};

#ifdef DPC_X86_SIMD
/* Eight aligned columns at a time (the long diagonal runs are where the rebuild spends its time): characters are
 * read backwards, positions are an arithmetic sequence, records are interleaved with unpacks.  Columns whose
 * characters differ get '?' here and are fixed by the caller (it knows the case / ambiguity rules).  No column may
 * be '*' (the caller checks the device's flag).  Returns the number of columns written (a multiple of 8). */
__attribute__((target("avx2"))) inline int mrun_avx2(dpc_pair_t *dst, const char *qlast, const char *glast, int len,
                                                     int qpos, int gpos, int step, int idx, bool stream, uint32_t *diffmask,
                                                     bool acgt_plain) {
  /* qlast / glast point at the characters of the run's FIRST column; column j reads qlast[-j] */
  const __m256i rev = _mm256_setr_epi32(7, 6, 5, 4, 3, 2, 1, 0);
  const __m256i iota = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
  const __m256i vstep = _mm256_set1_epi32(step);
  __m256i vq = _mm256_sub_epi32(_mm256_set1_epi32(qpos), _mm256_mullo_epi32(iota, vstep));
  __m256i vg = _mm256_sub_epi32(_mm256_set1_epi32(gpos), _mm256_mullo_epi32(iota, vstep));
  const __m256i dec = _mm256_slli_epi32(vstep, 3);
  const __m256i vidx = _mm256_set1_epi32(idx);
  const __m256i star = _mm256_set1_epi32('*' << 8), qm = _mm256_set1_epi32('?' << 8), space = _mm256_set1_epi32(' ' << 8);
  const __m256i vplain = _mm256_set1_epi32(acgt_plain ? -1 : 0);
  const __m256i cA = _mm256_set1_epi32('A'), cC = _mm256_set1_epi32('C'), cG = _mm256_set1_epi32('G'), cT = _mm256_set1_epi32('T');
  int j = 0, w = 0;
  for (; j + 8 <= len; j += 8, w++) {
    __m256i q = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i *)(qlast - j - 7)));
    __m256i g = _mm256_cvtepu8_epi32(_mm_loadl_epi64((const __m128i *)(glast - j - 7)));
    q = _mm256_permutevar8x32_epi32(q, rev);
    g = _mm256_permutevar8x32_epi32(g, rev);
    const __m256i eq = _mm256_cmpeq_epi32(q, g);
    /* two different upper-case bases are a plain mismatch (consistent_array is false for them, dynprog.c:1150-1157);
       only columns with anything else on either side (lower case, N, IUPAC codes) go to the exact rule */
    const __m256i qb = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi32(q, cA), _mm256_cmpeq_epi32(q, cC)),
                                       _mm256_or_si256(_mm256_cmpeq_epi32(q, cG), _mm256_cmpeq_epi32(q, cT)));
    const __m256i gb = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi32(g, cA), _mm256_cmpeq_epi32(g, cC)),
                                       _mm256_or_si256(_mm256_cmpeq_epi32(g, cG), _mm256_cmpeq_epi32(g, cT)));
    const __m256i plain = _mm256_or_si256(eq, _mm256_and_si256(_mm256_and_si256(qb, gb), vplain));
    diffmask[w] = ~(uint32_t)_mm256_movemask_ps(_mm256_castsi256_ps(plain)) & 0xffu;
    const __m256i tail = _mm256_or_si256(_mm256_or_si256(q, _mm256_slli_epi32(g, 16)),
                                         _mm256_blendv_epi8(_mm256_blendv_epi8(qm, space, plain), star, eq));
    const __m256i ab_lo = _mm256_unpacklo_epi32(vq, vg), ab_hi = _mm256_unpackhi_epi32(vq, vg);
    const __m256i cd_lo = _mm256_unpacklo_epi32(vidx, tail), cd_hi = _mm256_unpackhi_epi32(vidx, tail);
    const __m256i r04 = _mm256_unpacklo_epi64(ab_lo, cd_lo), r15 = _mm256_unpackhi_epi64(ab_lo, cd_lo);
    const __m256i r26 = _mm256_unpacklo_epi64(ab_hi, cd_hi), r37 = _mm256_unpackhi_epi64(ab_hi, cd_hi);
    __m128i *o = (__m128i *)(dst + j);
    if (stream) {
      _mm_stream_si128(o + 0, _mm256_castsi256_si128(r04)); _mm_stream_si128(o + 1, _mm256_castsi256_si128(r15));
      _mm_stream_si128(o + 2, _mm256_castsi256_si128(r26)); _mm_stream_si128(o + 3, _mm256_castsi256_si128(r37));
      _mm_stream_si128(o + 4, _mm256_extracti128_si256(r04, 1)); _mm_stream_si128(o + 5, _mm256_extracti128_si256(r15, 1));
      _mm_stream_si128(o + 6, _mm256_extracti128_si256(r26, 1)); _mm_stream_si128(o + 7, _mm256_extracti128_si256(r37, 1));
    } else {
      _mm_storeu_si128(o + 0, _mm256_castsi256_si128(r04)); _mm_storeu_si128(o + 1, _mm256_castsi256_si128(r15));
      _mm_storeu_si128(o + 2, _mm256_castsi256_si128(r26)); _mm_storeu_si128(o + 3, _mm256_castsi256_si128(r37));
      _mm_storeu_si128(o + 4, _mm256_extracti128_si256(r04, 1)); _mm_storeu_si128(o + 5, _mm256_extracti128_si256(r15, 1));
      _mm_storeu_si128(o + 6, _mm256_extracti128_si256(r26, 1)); _mm_storeu_si128(o + 7, _mm256_extracti128_si256(r37, 1));
    }
    vq = _mm256_sub_epi32(vq, dec);
    vg = _mm256_sub_epi32(vg, dec);
  }
  return j;
}
inline bool have_avx2() { static const bool h = __builtin_cpu_supports("avx2"); return h; }
#endif

struct Batch {
  std::vector<HostProb> probs;
  PBuf<uint8_t> pool;               /* query bytes + aux arrays; copied verbatim to the device */
  PBuf<DevProb> dprobs;
  std::vector<uint32_t> dev2host;
  /* the caller's arrays (bulk API) or copies (ticket API) */
  const dpc_problem_t *ext; dpc_result_t *ext_res;
  std::vector<dpc_problem_t> own; std::vector<dpc_result_t> own_res;

  /* genome characters as the device staged them (one span per matrix), returned with the results: the rebuild reads
     them instead of decoding the 2-bit genome a second time at an unpredictable address per problem */
  uint32_t gout_total;
  const uint8_t *gout_host;
  /* microexon searches: their scans run on the device next to the solve kernels */
  std::vector<MicroProb> micro;
  PBuf<ScanQuery> scans;
  uint64_t hits_total;

  Batch() : ext(NULL), ext_res(NULL), gout_total(8), gout_host(NULL), hits_total(0) {}
  void clear() { probs.clear(); pool.clear(); dprobs.clear(); dev2host.clear(); own.clear(); own_res.clear(); ext = NULL; ext_res = NULL;
                 gout_total = 8; gout_host = NULL; micro.clear(); scans.clear(); hits_total = 0; }
  const dpc_problem_t &P(int i) const { return ext ? ext[i] : own[i]; }
  dpc_result_t &R(int i) { return ext_res ? ext_res[i] : own_res[i]; }

  uint32_t pool_put(const char *src, int n) {
    uint32_t at = (uint32_t)pool.size();
    memcpy(pool.grow((size_t)n), src, (size_t)n);
    return at;
  }
  void pool_align(size_t a) { while (pool.size() % a) pool.push_back(0); }

  static void result_init(dpc_result_t &r, const dpc_problem_t &p) {
    r.null_list = 1; r.dynprogindex_out = p.dynprogindex;
    r.finalscore = r.nmatches = r.nmismatches = r.nopens = r.nindels = DPC_UNSET;
    r.new_leftgenomepos = r.new_rightgenomepos = r.exonhead = r.introntype = DPC_UNSET;
    r.incompletep = DPC_UNSET; r.npairs = 0; r.reserved = 0;
    r.left_prob = r.right_prob = -1.0;
  }
  static int bump(int idx) { return idx + (idx > 0 ? 1 : -1); }      /* e.g. dynprog.c:4570 */
  static int quality(double defect_rate) {                             /* dynprog.h:27-28 */
    return defect_rate < 0.003 ? HIGHQ : defect_rate < 0.014 ? MEDQ : LOWQ;
  }
  static bool alphabet_ok(const char *s, int n) {
    unsigned char acc = 0;
    for (int i = 0; i < n; i++) acc |= (unsigned char)s[i];
    return acc < 128;
  }
  static bool segment_ok(const dpc_problem_t &p) {
    if (allstar(p)) return true;
    uint64_t end = (uint64_t)(uint32_t)(p.chroffset + p.chrpos) + p.genomiclength;
    return end <= G().genome_nbases;
  }

  /* known-site / probability positions: get_splicesite_probs 3195-3287, the arrays of 3377-3458 and 3856-3903 */
  static void site(const dpc_problem_t &p, bool left, int c, uint32_t *pos, int *which, int *sign) {
    const int lo = p.offset2, ro = p.offset2R;
    if (left) {
      if (p.watsonp) { *pos = p.chrpos + lo + c; *which = p.cdna_direction > 0 ? 0 : 3; *sign = p.cdna_direction > 0 ? +1 : -1; }
      else { *pos = p.chrpos + (p.genomiclength - 1) - lo - c + 1; *which = p.cdna_direction > 0 ? 2 : 1; *sign = p.cdna_direction > 0 ? -1 : +1; }
    } else {
      if (p.watsonp) { *pos = p.chrpos + ro - c + 1; *which = p.cdna_direction > 0 ? 1 : 2; *sign = p.cdna_direction > 0 ? +1 : -1; }
      else { *pos = p.chrpos + (p.genomiclength - 1) - ro + c; *which = p.cdna_direction > 0 ? 3 : 0; *sign = p.cdna_direction > 0 ? -1 : +1; }
    }
  }
  static bool site_known(const dpc_problem_t &p, bool left, int c) {
    const dpc_setup_t &s = G().setup;
    uint32_t pos; int which, sign;
    if (!s.splice_known) return false;
    site(p, left, c, &pos, &which, &sign);
    return s.splice_known(which, p.chrnum, pos, sign, s.user) != 0;
  }
  static double site_prob(const dpc_problem_t &p, bool left, int c, bool known) {
    const dpc_setup_t &s = G().setup;
    uint32_t pos; int which, sign;
    if (known) return 1.0;
    site(p, left, c, &pos, &which, &sign);
    return s.splice_prob(which, p.chroffset + pos, p.chroffset, s.user);
  }

  /* Dynprog_microexon_int up to its scans: p-value -> minimum microexon length (7158-7232), the mismatch-bounded
     ends (7234-7283), the (cL, cR) pairs with splice-site dinucleotides (7289-7312) and one device scan per pair */
  static void micro_introns(const dpc_problem_t &p, char *i1, char *i2, char *i3, char *i4, char *gapchar, int *type) {
    if (p.cdna_direction > 0) { *i1 = 'G'; *i2 = 'T'; *i3 = 'A'; *i4 = 'G'; *gapchar = '>'; *type = 0x20; }     /* GTAG_FWD */
    else { *i1 = 'C'; *i2 = 'T'; *i3 = 'A'; *i4 = 'C'; *gapchar = '<'; *type = 0x04; }                           /* GTAG_REV */
  }
  int add_micro(const dpc_problem_t &p, dpc_result_t &r, HostProb &h) {
    const Globals &g = G();
    const int L1 = p.length1, lo = p.offset2, ro = p.offset2R;
    char i1, i2, i3, i4, gapchar; int type;
    if (p.cdna_direction == 0 || ro - lo <= 0 || L1 <= 0 || p.seq1 == NULL) return DPC_ERR_ARG;     /* abort(), 7192, 7215 */
    if (g.setup.splice_prob == NULL) return DPC_ERR_STATE;
    if (!alphabet_ok(p.seq1, L1)) return DPC_ERR_ALPHABET;
    if (!allstar(p) && !segment_ok(p)) return DPC_ERR_ARG;
    micro_introns(p, &i1, &i2, &i3, &i4, &gapchar, &type);
    r.left_prob = r.right_prob = 0.0;
    r.introntype = type;
    const double pvalue = p.defect_rate < 0.003 ? 0.01 : p.defect_rate < 0.014 ? 0.001 : 0.0001;
    int minlen = (int)ceil(-log(1.0 - pow(1.0 - pvalue, 1.0 / (double)(ro - lo))) / log(4));
    minlen -= 8;
    if (minlen > 12) { r.introntype = 0; return 0; }                   /* MAX_MICROEXON_LENGTH, 7222-7227 */
    if (minlen < 3) minlen = 3;
    int leftbound = 0, rightbound = 0, nmm = 0, i;
    while (leftbound < L1 - 1 && nmm <= 1) { if ((char)dpc_query_uc(p.seq1[leftbound]) != host_genomic_nt(p, lo + leftbound)) nmm++; leftbound++; }
    leftbound--;
    i = L1 - 1; nmm = 0;
    while (i >= 0 && nmm <= 1) { if ((char)dpc_query_uc(p.seq1[i]) != host_genomic_nt(p, ro - rightbound)) nmm++; rightbound++; i--; }
    rightbound--;
    MicroProb mp;
    mp.problem = (int)probs.size();
    for (int cL = 1; cL <= leftbound; cL++) {
      if (!(host_genomic_nt(p, lo + cL) == i1 && host_genomic_nt(p, lo + cL + 1) == i2)) continue;
      int mincR = L1 - 12 - cL, maxcR = L1 - minlen - cL;
      if (mincR < 1) mincR = 1;
      if (maxcR > rightbound) maxcR = rightbound;
      for (int cR = mincR; cR <= maxcR; cR++) {
        if (!(host_genomic_nt(p, ro - cR - 1) == i3 && host_genomic_nt(p, ro - cR) == i4)) continue;
        MicroCand c;
        c.cL = cL; c.cR = cR; c.mid = L1 - cL - cR; c.textleft = lo + cL + DPC_MICROINTRON; c.query = -1;
        const int textlen = (ro - cR - DPC_MICROINTRON) - c.textleft;
        uint32_t pat = 0; bool okay = c.mid >= 1 && c.mid <= 16;
        for (int k = 0; k < c.mid && okay; k++) {                      /* query_okay, boyer-moore.c:268 */
          const int ch = dpc_query_uc(p.seq1[cL + k]);
          const int code = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
          if (code < 0) okay = false; else pat |= (uint32_t)code << (2 * k);
        }
        if (okay && textlen - c.mid >= 0) {
          ScanQuery q;
          memset(&q, 0, sizeof q);
          q.gbase = p.chroffset + p.chrpos; q.glen = p.genomiclength; q.textleft = c.textleft; q.npos = textlen - c.mid + 1;
          q.pat = pat; q.len = (uint8_t)c.mid; q.watson = p.watsonp ? 1 : 0;
          if (hits_total + (uint64_t)q.npos > (64ull << 20)) return DPC_ERR_NOMEM;
          q.hits_off = (uint32_t)hits_total; hits_total += (uint64_t)q.npos;
          c.query = (int)scans.size();
          scans.push_back(q);
        }
        mp.cands.push_back(c);
      }
    }
    if (mp.cands.empty()) { r.introntype = 0; return 0; }              /* no candidate pair: nothing to scan, NULL */
    h.micro = (int)micro.size();
    micro.push_back(mp);
    return 0;
  }
  /* ... and after them: flanks of every hit (7327-7330), MaxEnt probabilities through the hook (7336-7372), the best
     pair, and make_microexon_pairs_double (6941-7053) around the LAST hit looked at (sic, 7401).  hits / count: what
     the device scan returned (positions j, any order). */
  void finalize_micro(MicroProb &mp, const uint32_t *hits, const uint32_t *count) {
    const dpc_problem_t &p = P(mp.problem);
    dpc_result_t &r = R(mp.problem);
    const dpc_setup_t &su = G().setup;
    const int lo = p.offset2, ro = p.offset2R;
    char i1, i2, i3, i4, gapchar; int type;
    micro_introns(p, &i1, &i2, &i3, &i4, &gapchar, &type);
    int bestcL = -1, bestcR = -1, bestmid = 0, candidate = 0;
    double bestprob = 0.0;
    std::vector<uint32_t> js;
    for (size_t k = 0; k < mp.cands.size(); k++) {
      const MicroCand &c = mp.cands[k];
      if (c.query < 0) continue;
      const ScanQuery &q = scans[(size_t)c.query];
      js.assign(hits + q.hits_off, hits + q.hits_off + count[c.query]);
      std::sort(js.begin(), js.end(), std::greater<uint32_t>());      /* BoyerMoore_nt pushes its hits: highest position first */
      for (size_t t = 0; t < js.size(); t++) {
        candidate = c.textleft + (int)js[t];
        if (!(host_genomic_nt(p, candidate - 2) == i3 && host_genomic_nt(p, candidate - 1) == i4 &&
              host_genomic_nt(p, candidate + c.mid) == i1 && host_genomic_nt(p, candidate + c.mid + 1) == i2)) continue;
        uint32_t s2, s3; int w2, w3;
        if (p.watsonp) {
          s2 = p.chrpos + (uint32_t)(candidate - 1) + 1; s3 = p.chrpos + (uint32_t)(candidate + c.mid);
          if (p.cdna_direction > 0) { w2 = 1; w3 = 0; } else { w2 = 2; w3 = 3; }
        } else {
          s2 = p.chrpos + (p.genomiclength - 1) - (uint32_t)(candidate - 1); s3 = p.chrpos + (p.genomiclength - 1) - (uint32_t)(candidate + c.mid) + 1;
          if (p.cdna_direction > 0) { w2 = 3; w3 = 2; } else { w2 = 0; w3 = 1; }
        }
        const double prob2 = su.splice_prob(w2, p.chroffset + s2, p.chroffset, su.user);
        const double prob3 = su.splice_prob(w3, p.chroffset + s3, p.chroffset, su.user);
        if (prob2 + prob3 > bestprob) { bestcL = c.cL; bestcR = c.cR; bestmid = c.mid; r.left_prob = prob2; r.right_prob = prob3; bestprob = prob2 + prob3; }
      }
    }
    mp.pairs.clear();
    if (bestcL < 0 || bestcR < 0) { r.introntype = 0; r.npairs = 0; r.null_list = 1; return; }
    std::vector<dpc_pair_t> st;
    Out o;
    st.resize((size_t)(bestcL + bestmid + bestcR + 2));
    o.p = st.data(); o.n = 0;
    const int off1[3] = { p.offset1, p.offset1 + bestcL, p.offset1 + bestcL + bestmid };
    const int off2[3] = { lo, candidate, ro - bestcR + 1 }, len[3] = { bestcL, bestmid, bestcR };
    for (int part = 0; part < 3; part++) {
      if (part > 0) o.push(-1, -1, ' ', gapchar, ' ', 0, 1);           /* gapholder with comp = gapchar, 6979-6982 */
      for (int k = 0; k < len[part]; k++) {
        const char c1 = p.seq1[off1[part] - p.offset1 + k], c2 = host_genomic_nt(p, off2[part] + k);
        const char comp = (char)dpc_query_uc(c1) == c2 ? '*' : G().CONS[c1 & 127][c2 & 127] ? ':' : ' ';
        o.push(off1[part] + k, off2[part] + k, c1, comp, c2, p.dynprogindex, 0);
      }
    }
    mp.pairs.assign(st.rbegin(), st.rend());                          /* the list comes back as pushed: last pair first */
    r.npairs = (int32_t)mp.pairs.size(); r.null_list = 0;
    r.dynprogindex_out = bump(p.dynprogindex);
  }

  /* ticket API: copies the problem */
  int add(const dpc_problem_t &in) {
    if (ext) return DPC_ERR_STATE;
    own.push_back(in);
    own_res.push_back(dpc_result_t());
    int rc = add_impl(own.back(), own_res.back());
    if (rc < 0) { own.pop_back(); own_res.pop_back(); }
    return rc;
  }
  /* bulk API: the caller's arrays stay valid until the results are out */
  int add_ext(const dpc_problem_t *problems, dpc_result_t *results, int n) {
    if (!probs.empty()) return DPC_ERR_STATE;
    ext = problems; ext_res = results;
    probs.reserve((size_t)n); dprobs.reserve((size_t)n); dev2host.reserve((size_t)n);
    for (int i = 0; i < n; i++) {
      int rc = add_impl(problems[i], results[i]);
      if (rc < 0) return rc;
    }
    return 0;
  }

  /* Returns the ticket or a negative code.  Mirrors the argument checks and early returns of the
   * five reference entry points; everything that needs a matrix becomes a DevProb. */
  int add_impl(const dpc_problem_t &p, dpc_result_t &r) {
    Globals &g = G();
    if (!g.inited || !g.setup_done) return DPC_ERR_STATE;
    HostProb h;
    h.q0 = h.q1 = h.aux = 0; h.dev = -1; h.L1 = p.length1; h.L2 = p.length2; h.gout = DPC_NO_GOUT; h.micro = -1;
    result_init(r, p);
    dprobs.reserve(dprobs.size() + 1);
    DevProb &d = dprobs.data()[dprobs.size()];      /* built in place; committed below when it goes to the device */
    memset(&d, 0, sizeof d);
    bool todev = false;
    d.kind = (uint8_t)p.kind; d.endalign = (uint8_t)p.endalign;
    d.gbase = p.chroffset + p.chrpos; d.glen = p.genomiclength;
    d.extraband = p.extraband; d.cdna_direction = (int8_t)(p.cdna_direction > 0 ? 1 : p.cdna_direction < 0 ? -1 : 0);
    d.score_threshold = p.score_threshold;
    d.flags = (p.watsonp ? DPC_F_WATSON : 0) | (p.jump_late_p ? DPC_F_LATE : 0) | (p.widebandp ? DPC_F_WIDEBAND : 0) |
              (p.halfp ? DPC_F_HALFP : 0) | (p.finalp ? DPC_F_FINALP : 0) | (allstar(p) ? DPC_F_ALLSTAR : 0) |
              (g.setup.novelsplicingp ? DPC_F_NOVEL : 0);
    if (p.extraband < 0 || p.extraband > 4000) return DPC_ERR_ARG;

    switch (p.kind) {
    case DPC_SINGLE_GAP: {                                             /* Dynprog_single_gap, 4450-4572 */
      int L1 = p.length1, L2 = p.length2, lband, rband;
      if (L1 > g.maxlength1 || L2 > g.maxlength2) {                   /* 4509-4519 */
        r.finalscore = -10000; r.nmatches = r.nmismatches = r.nopens = r.nindels = 0;
        r.dynprogindex_out = bump(p.dynprogindex);
        break;
      }
      if (L1 <= 0 || L2 <= 0) return DPC_ERR_ARG;                     /* Matrix3_alloc aborts, 495-498 */
      dpc_bands(L1, L2, p.extraband, p.widebandp, &lband, &rband);
      if (L2 - L1 > rband || L1 - L2 > lband) {
        /* only without widebandp: the corner was never filled; the reference reads the memset 0
         * (or the forced NEG one step above the band, 1501-1506) and traces nothing back */
        if (L2 - L1 > rband + 1) return DPC_ERR_ARG;                  /* the reference writes outside its matrix here */
        r.finalscore = (L2 - L1 == rband + 1) ? DPC_NEG_INFINITY : 0;
        r.nmatches = r.nmismatches = r.nopens = r.nindels = 0;
        r.dynprogindex_out = bump(p.dynprogindex);
        break;
      }
      if (!alphabet_ok(p.seq1, L1)) return DPC_ERR_ALPHABET;
      d.type = (uint8_t)quality(p.defect_rate); d.open = -10; d.extend = -3;    /* SINGLE, 222-229 */
      d.L1 = L1; d.L2 = L2; d.off2 = p.offset2;
      d.q0 = h.q0 = pool_put(p.seq1, L1);
      todev = true;
      break;
    }
    case DPC_END5_GAP: case DPC_END3_GAP: {                            /* 5094-5284, 5556-5741 */
      const bool five = p.kind == DPC_END5_GAP;
      int L1 = p.length1, L2 = p.length2, ea = p.endalign;
      if (ea < 0 || ea > 3) return DPC_ERR_ARG;                        /* abort(), 5215 */
      if (L1 <= 0 || L2 <= 0) {                                        /* 5140-5157 */
        r.nmatches = r.nmismatches = r.nopens = r.nindels = 0; r.finalscore = 0;
        break;
      }
      if (ea != DPC_QUERYEND_NOGAPS) {
        if (L1 > g.maxlength1) L1 = g.maxlength1;
        if (L2 > g.maxlength2) L2 = g.maxlength2;
      } else {
        L1 = L2 = (L1 < L2 ? L1 : L2);                                 /* 2358-2369 */
        /* not clipped by the reference; the diagonal leaves the device as M runs of at most DPC_OP_MAXLEN columns */
        if (L1 > DPC_INLINE_OPS * DPC_OP_MAXLEN) return DPC_ERR_UNSUPPORTED;
      }
      h.L1 = L1; h.L2 = L2;
      if (!alphabet_ok(five ? p.seq1 - (L1 - 1) : p.seq1, L1)) return DPC_ERR_ALPHABET;
      d.type = ENDQ; d.open = -12; d.extend = -1;                      /* END, 240-247; ENDQ 179 */
      d.flags |= DPC_F_WIDEBAND;
      d.L1 = L1; d.L2 = L2; d.off2 = p.offset2;
      d.q0 = h.q0 = five ? pool_put(p.seq1 - (L1 - 1), L1) : pool_put(p.seq1, L1);
      todev = true;
      break;
    }
    case DPC_END5_SPLICEJUNCTION: case DPC_END3_SPLICEJUNCTION: {      /* 5411-5552, 5869-6012 */
      const bool five = p.kind == DPC_END5_SPLICEJUNCTION;
      const int L1 = p.length1, L2 = p.length2;
      if (L1 <= 0 || L1 > g.maxlength1 || L2 <= 0 || L2 > g.maxlength2) {      /* 5452-5465: no chopping here */
        r.nmatches = r.nmismatches = r.nopens = r.nindels = 0; r.finalscore = 0;
        break;
      }
      if (p.seq1R == NULL || p.length2R < 0) return DPC_ERR_ARG;
      if (!alphabet_ok(five ? p.seq1 - (L1 - 1) : p.seq1, L1)) return DPC_ERR_ALPHABET;
      d.kind = (uint8_t)(five ? DPC_END5_GAP : DPC_END3_GAP);           /* same device work as an end gap ... */
      d.endalign = DPC_QUERYEND_INDELS;                                /* ... with the end point on the last row, 5506 */
      d.type = ENDQ; d.open = -12; d.extend = -1;                      /* END, 240-247; ENDQ 179 */
      d.flags |= DPC_F_WIDEBAND | DPC_F_SEQ2;
      d.L1 = L1; d.L2 = L2; d.off2 = p.offset2;
      d.q0 = h.q0 = five ? pool_put(p.seq1 - (L1 - 1), L1) : pool_put(p.seq1, L1);
      {
        /* the junction string as genome codes in matrix order (column i = sequence2[+-i]) */
        uint8_t *codes = pool.grow((size_t)L2);
        d.q1 = h.q1 = (uint32_t)(codes - pool.data());
        for (int i = 0; i < L2; i++) {
          const int ch = (unsigned char)(five ? p.seq1R[-i] : p.seq1R[i]);
          const int code = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : ch == 'N' ? 4 : -1;
          if (code < 0) return DPC_ERR_ALPHABET;
          codes[i] = (uint8_t)code;
        }
      }
      d.flags |= DPC_F_ALLSTAR;          /* no genome access: the segment checks below do not apply */
      todev = true;
      break;
    }
    case DPC_GENOME_GAP: {                                             /* Dynprog_genome_gap, 4798-5061 */
      int L1 = p.length1, L2L = p.length2, L2R = p.length2R;
      r.nmatches = r.nmismatches = r.nopens = r.nindels = 0;
      r.left_prob = r.right_prob = 0.0;
      if (L1 <= 1) { r.finalscore = DPC_NEG_INFINITY; break; }         /* 4855-4858 */
      if (L1 > g.maxlength1 || L2L > g.maxlength2 || L2R > g.maxlength2) {     /* 4922-4954 */
        r.new_leftgenomepos = p.offset2 - 1; r.new_rightgenomepos = p.offset2R + 1; r.exonhead = p.offset1 + L1 - 1;
        r.dynprogindex_out = bump(p.dynprogindex); r.finalscore = DPC_NEG_INFINITY;
        break;
      }
      if (L2L <= 0 || L2R <= 0 || L2L < L1 - 1 || L2R < L1 - 1) return DPC_ERR_ARG;
      if ((p.finalp || p.use_probabilities_p) && g.setup.splice_prob == NULL) return DPC_ERR_STATE;
      if (!alphabet_ok(p.seq1, L1)) return DPC_ERR_ALPHABET;
      d.type = (uint8_t)quality(p.defect_rate);
      if (L1 > p.maxpeelback * 4) { d.open = -10; d.extend = -3; } else { d.open = -18; d.extend = -3; }   /* 4862-4870 */
      d.reward = (int8_t)(!p.splicingp ? 0 : (p.finalp ? 30 : 10) + 6 * d.type);     /* 277-283, 4871-4877 */
      d.L1 = L1; d.L2 = L2L; d.L2R = L2R; d.off2 = p.offset2; d.off2R = p.offset2R;
      d.gap = p.offset2R - p.offset2;
      d.q0 = h.q0 = pool_put(p.seq1, L1);
      /* novelsplicingp == false with an intron-level IIT: the bridge only looks at given introns and ignores
         use_probabilities_p (the test at 3552 comes first) */
      const bool constrained = !g.setup.novelsplicingp && g.setup.splice_known && g.setup.intron_level;
      if (constrained && g.setup.splice_intron == NULL) return DPC_ERR_STATE;
      const bool probmode = p.use_probabilities_p && !constrained;
      if (probmode || g.setup.splice_known) {
        /* aux block: [probabilities: (L2L+1) + (L2R+1) doubles] [known flags: (L2L+1) + (L2R+1) bytes]
           [given introns: uint32 count, then (cL, cR) uint16 pairs].  One spare entry per side, zero like the
           reference's CALLOC(length2+1) arrays: the bridge reads index length1-1, which may equal length2 */
        pool_align(8);
        d.aux = h.aux = (uint32_t)pool.size();
        std::vector<uint8_t> known((size_t)L2L + L2R + 2, 0);
        uint8_t *lknown = known.data(), *rknown = known.data() + L2L + 1;
        if (g.setup.splice_known) {
          d.flags |= DPC_F_KNOWN;
          for (int c = 0; c < L2L - 1; c++) lknown[c] = site_known(p, true, c);
          for (int c = 0; c < L2R - 1; c++) rknown[c] = site_known(p, false, c);
        }
        if (probmode) {
          d.flags |= DPC_F_PROBMODE;
          std::vector<double> pr((size_t)L2L + L2R + 2, 0.0);
          for (int c = 0; c < L2L - 1; c++) pr[c] = site_prob(p, true, c, lknown[c] != 0);
          for (int c = 0; c < L2R - 1; c++) pr[L2L + 1 + c] = site_prob(p, false, c, rknown[c] != 0);
          pool_put((const char *)pr.data(), (int)(pr.size() * sizeof(double)));
        }
        if (g.setup.splice_known) pool_put((const char *)known.data(), (int)known.size());
        if (constrained) {
          /* the given introns among the known (left, right) site pairs, 3596-3615: the IIT lookup happens here */
          d.flags |= DPC_F_INTRONS;
          std::vector<uint16_t> pairs;
          const dpc_setup_t &su = g.setup;
          for (int cL = 1; cL < L2L - 1; cL++) {
            if (!lknown[cL]) continue;
            for (int cR = 1; cR < L2R - 1; cR++) {
              if (!rknown[cR]) continue;
              int yes;
              if (p.watsonp) {
                const uint32_t pos1 = p.chrpos + p.offset2 + cL, pos2 = p.chrpos + p.offset2R - cR + 1;
                yes = su.splice_intron(p.chrnum, pos1, pos2 + 1U, p.cdna_direction, su.user);
              } else {
                const uint32_t pos1 = p.chrpos + (p.genomiclength - 1) - p.offset2 - cL + 1;
                const uint32_t pos2 = p.chrpos + (p.genomiclength - 1) - p.offset2R + cR;
                yes = su.splice_intron(p.chrnum, pos2, pos1 + 1U, -p.cdna_direction, su.user);
              }
              if (yes) { pairs.push_back((uint16_t)cL); pairs.push_back((uint16_t)cR); }
            }
          }
          pool_align(4);
          const uint32_t npairs = (uint32_t)(pairs.size() / 2);
          pool_put((const char *)&npairs, 4);
          if (npairs) pool_put((const char *)pairs.data(), (int)(pairs.size() * sizeof(uint16_t)));
        }
      }
      todev = true;
      break;
    }
    case DPC_CDNA_GAP: {                                               /* Dynprog_cdna_gap, 4577-4793 */
      int L1L = p.length1, L1R = p.length1R, L2 = p.length2;
      if (L2 <= 1) break;                                              /* 4605-4607: nothing is written */
      if (L2 > g.maxlength1 || L1R > g.maxlength2 || L1L > g.maxlength2) {     /* 4648-4670 */
        r.dynprogindex_out = bump(p.dynprogindex);
        break;
      }
      if (L1L <= 0 || L1R <= 0) return DPC_ERR_ARG;
      /* the two query ends must be one span: the SHORTGAP insertion (4730-4751) indexes across it */
      int span = p.offset1R - p.offset1 + 1;
      if (span < L1L || span < L1R || span > 100000 || p.seq1R != p.seq1 + (span - 1)) return DPC_ERR_ARG;
      if (!alphabet_ok(p.seq1, span)) return DPC_ERR_ALPHABET;
      d.type = (uint8_t)quality(p.defect_rate); d.open = -10; d.extend = -7;         /* CDNA, 231-238 */
      d.L1 = L1L; d.L1R = L1R; d.L2 = L2; d.off2 = p.offset2;
      d.gap = p.offset1R - p.offset1;
      d.q0 = h.q0 = pool_put(p.seq1, span);
      d.q1 = h.q1 = d.q0 + (uint32_t)(span - 1);
      todev = true;
      break;
    }
    case DPC_MICROEXON_INT: {                                          /* Dynprog_microexon_int, 7127-7429 */
      int rc = add_micro(p, r, h);
      if (rc < 0) return rc;
      break;
    }
    default:
      return DPC_ERR_ARG;
    }
    d.gout = DPC_NO_GOUT;
    if (todev && !(d.flags & DPC_F_SEQ2) && p.kind != DPC_CDNA_GAP) {
      const uint32_t need = dpc_gout_span(d.L2) + (p.kind == DPC_GENOME_GAP ? dpc_gout_span(d.L2R) : 0u) + 8u;
      if ((uint64_t)gout_total + need < 0xfff00000ull) { d.gout = h.gout = gout_total; gout_total += need; }
    }
    if (todev) {
      if (!(d.flags & DPC_F_SEQ2) && !segment_ok(p)) return DPC_ERR_ARG;
      h.dev = (int32_t)dprobs.size();
      dprobs.grow(1);
      dev2host.push_back((uint32_t)probs.size());
    }
    probs.push_back(h);
    return (int)probs.size() - 1;
  }

  /* ---- rebuild of the Pair records (dynprog.c:2372-2712 and the assembly in each entry point) */

  /* One matrix: replays the ops from (r,c).  qch / gch are in matrix order. */
  template <bool REV, bool GROWS>
  static void replay_t(Out &st, const uint16_t *ops, int nops, int r, int c, const char *qch, const char *gch,
                       int q0, int g0, int idx, bool nostar) {
    const int step = REV ? -1 : 1;
    const Globals &g = G();
    for (int i = 0; i < nops; i++) {
      const int op = ops[i] & 3, len = ops[i] >> 2;
      if (op == DPC_OP_M) {
        int qi = (GROWS ? c : r) - 1, gi = (GROWS ? r : c) - 1;
        int qpos = q0 + step * qi, gpos = g0 + step * gi;
        int j = 0;
#ifdef DPC_X86_SIMD
        if (!GROWS && nostar && len >= 8 && have_avx2()) {
          /* the bulk of the run eight columns at a time; differing columns are patched with the exact rule */
          uint32_t diff[512];
          dpc_pair_t *base = st.p + st.n;
          const bool stream = st.stream && (((uintptr_t)base) & 15) == 0;
          const int done = mrun_avx2(base, qch + qi, gch + gi, len > 4096 ? 4096 : len, qpos, gpos, step, idx, stream, diff, g.acgt_plain);
          for (int w = 0; w < done / 8; w++)
            for (uint32_t mask = diff[w]; mask; mask &= mask - 1) {
              const int jj = 8 * w + __builtin_ctz(mask);
              const char c1 = qch[qi - jj], c2 = gch[gi - jj];
              char comp = '*';
              if ((char)dpc_query_uc(c1) != c2) comp = g.CONS[c1 & 127][c2 & 127] ? ':' : ' ';
              if (stream) {        /* the record went around the cache: rewrite it whole */
                Out one; one.p = base + jj; one.n = 0; one.stream = true;
                one.push(qpos - step * jj, gpos - step * jj, c1, comp, c2, idx, 0);
              } else base[jj].comp = comp;
            }
          st.n += done;
          j = done; qi -= done; gi -= done; qpos -= step * done; gpos -= step * done;
        }
#endif
        for (; j < len; j++, qi--, gi--, qpos -= step, gpos -= step) {
          const char c1 = qch[qi], c2 = gch[gi];
          if (!GROWS && c2 == '*') continue;                          /* 2644 */
          char comp = '*';
          if (c1 != c2 && (char)dpc_query_uc(c1) != c2) {
            const bool consistent = GROWS ? g.CONS[c2 & 127][c1 & 127] : g.CONS[c1 & 127][c2 & 127];   /* 2654 vs 2752 */
            comp = consistent ? ':' : ' ';
          }
          st.push(qpos, gpos, c1, comp, c2, idx, 0);
        }
        r -= len; c -= len;
        continue;
      }
      const bool along_cols = (op == DPC_OP_QSKIP) ? GROWS : !GROWS;
      if (along_cols) c -= len; else r -= len;
      if (op == DPC_OP_GAPHOLDER) { st.push_gapholder(); continue; }  /* 2507 */
      if (op == DPC_OP_GSKIP) {                                       /* add_genomeskip dashes, 2444-2505 */
        const int lo = GROWS ? r : c, qi2 = GROWS ? c - 1 : r - 1;
        const int qpos = REV ? q0 - qi2 : q0 + qi2 + 1;
        for (int j = 0; j < len; j++) {
          const int gi2 = lo + len - 1 - j;
          st.push(qpos, g0 + step * gi2, ' ', '-', gch[gi2], idx, 0);
        }
      } else {                                                        /* add_queryskip, 2372-2413 */
        const int lo = GROWS ? c : r, gi2 = GROWS ? r - 1 : c - 1;
        const int gpos = REV ? g0 - gi2 : g0 + gi2 + 1;
        for (int j = 0; j < len; j++) {
          const int qi2 = lo + len - 1 - j;
          st.push(q0 + step * qi2, gpos, qch[qi2], '-', ' ', idx, 0);
        }
      }
    }
  }
  static void replay(Out &st, const uint16_t *ops, int nops, int r, int c, const char *qch, const char *gch,
                     int q0, int g0, bool revp, bool genome_rows, int idx, bool nostar = false) {
    if (revp) {
      if (genome_rows) replay_t<true, true>(st, ops, nops, r, c, qch, gch, q0, g0, idx, nostar);
      else replay_t<true, false>(st, ops, nops, r, c, qch, gch, q0, g0, idx, nostar);
    } else {
      if (genome_rows) replay_t<false, true>(st, ops, nops, r, c, qch, gch, q0, g0, idx, nostar);
      else replay_t<false, false>(st, ops, nops, r, c, qch, gch, q0, g0, idx, nostar);
    }
  }

  static void emit(Out &out, const dpc_pair_t *v, int n, bool reversed) {
    for (int i = 0; i < n; i++) {
      const dpc_pair_t &pr = v[reversed ? n - 1 - i : i];
      out.push(pr.querypos, pr.genomepos, pr.cdna, pr.comp, pr.genome, pr.dynprogindex, pr.gapp);
    }
  }
  static char *fit(std::vector<char> &v, int n) { if ((int)v.size() < n + 8) v.resize((size_t)n + 64); return v.data(); }
  static dpc_pair_t *fit(std::vector<dpc_pair_t> &v, int n) { if ((int)v.size() < n + 8) v.resize((size_t)n + 64); return v.data(); }

  /* The rebuild reads the genome at an unpredictable place per problem: ask for the lines ahead of time. */
  const char *staged(const HostProb &h) const {
    return (gout_host && h.gout != DPC_NO_GOUT) ? (const char *)gout_host + h.gout : NULL;
  }
  void prefetch_genome(int i) const {
    if (gout_host) return;
    const dpc_problem_t &p = P(i);
    const uint32_t *blocks = G().setup.genome_blocks;
    const uint32_t base = p.chroffset + p.chrpos;
    const int offs[2] = { p.offset2, p.kind == DPC_GENOME_GAP ? p.offset2R : p.offset2 + p.length2 - 1 };
    for (int k = 0; k < 2; k++) {
      int64_t pos = offs[k];
      if (pos < 0) pos = 0;
      if (pos >= (int64_t)p.genomiclength) pos = (int64_t)p.genomiclength - 1;
      if (pos < 0) continue;
      const uint64_t g = p.watsonp ? (uint64_t)base + (uint64_t)pos : (uint64_t)base + (uint64_t)(p.genomiclength - 1 - pos);
      if (g < G().genome_nbases) __builtin_prefetch(blocks + (g >> 5) * 3);
    }
  }

  /* Writes the pairs of problem i, in the order of the List_T the reference returns, to dst (which must
   * hold max_pairs(i) records) and returns their number. */
  int max_pairs(int i) const {
    const dpc_problem_t &p = P(i);
    const HostProb &h = probs[i];
    switch (p.kind) {
    case DPC_GENOME_GAP: return 2 * p.length1 + p.length2 + p.length2R + 8;
    case DPC_CDNA_GAP: return p.length1 + p.length1R + 2 * p.length2 + 32;
    default: return h.L1 + h.L2 + 8;
    }
  }
  int rebuild(int i, const DevRes &dr, const uint16_t *ops, dpc_pair_t *dst, Scratch &s, bool stream_dst = false) const {
    const HostProb &h = probs[i];
    return rebuild_core(P(i), (const char *)&pool[h.q0], &pool[h.q1], h.L1, h.L2, staged(h), dr, ops, dst, s, stream_dst);
  }
  /* lengths as the entry points clip them (5179-5186, 2358-2369): what the device worked on */
  static void clipped_lengths(const dpc_problem_t &p, int *L1, int *L2) {
    *L1 = p.length1; *L2 = p.length2;
    if (p.kind == DPC_END5_GAP || p.kind == DPC_END3_GAP) {
      if (p.endalign != DPC_QUERYEND_NOGAPS) {
        if (*L1 > G().maxlength1) *L1 = G().maxlength1;
        if (*L2 > G().maxlength2) *L2 = G().maxlength2;
      } else *L1 = *L2 = (*L1 < *L2 ? *L1 : *L2);
    }
  }
  /* q: the query bytes of the problem in forward order (first byte of the span the problem points at); codes2: the
     junction string as genome codes (splice-junction solvers only); L1c / L2c: clipped lengths; staged: the genome
     characters the device staged for this problem, or NULL (decode the 2-bit genome here) */
  static int rebuild_core(const dpc_problem_t &p, const char *q, const uint8_t *codes2, int L1c, int L2c, const char *staged_g,
                          const DevRes &dr, const uint16_t *ops, dpc_pair_t *dst, Scratch &s, bool stream_dst) {
    const uint32_t *blocks = G().setup.genome_blocks;
    Out out; out.p = dst; out.n = 0; out.stream = stream_dst;
    const bool nostar = !(dr.status & DPC_ST_STAR);
    switch (p.kind) {
    case DPC_SINGLE_GAP: {
      const char *ga = staged_g;
#ifdef DPC_PROFILE_REBUILD
      unsigned long long t0 = __rdtsc();
#endif
      if (!ga) { char *buf = fit(s.ga, L2c); gather_genome(p, blocks, p.offset2, L2c, false, buf); ga = buf; }
#ifdef DPC_PROFILE_REBUILD
      unsigned long long t1 = __rdtsc();
#endif
      /* List_reverse of the pushed list (4571) = push order */
      replay(out, ops, dr.nopsL, dr.bestrL, dr.bestcL, q, ga, p.offset1, p.offset2, false, false, p.dynprogindex, nostar);
#ifdef DPC_PROFILE_REBUILD
      s.cyc_gather += t1 - t0; s.cyc_replay += __rdtsc() - t1;
#endif
      break;
    }
    case DPC_END5_GAP: case DPC_END3_GAP: {
      const bool five = p.kind == DPC_END5_GAP;
      char *qa = fit(s.qa, L1c);
      const char *ga = staged_g;
      for (int k = 0; k < L1c; k++) qa[k] = five ? q[L1c - 1 - k] : q[k];
      if (!ga) { char *buf = fit(s.ga, L2c); gather_genome(p, blocks, p.offset2, L2c, five, buf); ga = buf; }
      Out sL; sL.p = fit(s.sL, L1c + L2c + 2); sL.n = 0;
      replay(sL, ops, dr.nopsL, dr.bestrL, dr.bestcL, qa, ga, p.offset1, p.offset2, five, false, p.dynprogindex, nostar);
      if ((p.endalign == DPC_QUERYEND_GAP || p.endalign == DPC_BEST_LOCAL) && dr.nmatches + 1 < dr.nmismatches) break;   /* 5259 */
      int first = 0;                                                   /* 5265-5268 */
      while (first < sL.n && sL.p[first].comp == '-') first++;
      emit(out, sL.p + first, sL.n - first, five);                     /* end5: List_reverse again 5283; end3: as is 5740 */
      break;
    }
    case DPC_END5_SPLICEJUNCTION: case DPC_END3_SPLICEJUNCTION: {
      /* traceback_local twice (2875-2969): columns above contlength carry the far offset, then the known
         gapholder, then the rest with the anchor offset.  An iteration of the reference emits one aligned column
         and the gap run that follows it, so the split can only fall before an aligned column. */
      const bool five = p.kind == DPC_END5_SPLICEJUNCTION;
      const int endc = p.length2R;
      char *qa = fit(s.qa, L1c), *ga = fit(s.ga, L2c);
      for (int k = 0; k < L1c; k++) qa[k] = five ? q[L1c - 1 - k] : q[k];
      for (int k = 0; k < L2c; k++) ga[k] = (char)dpc_code_char(codes2[k]);
      Out sL; sL.p = fit(s.sL, L1c + L2c + 4); sL.n = 0;
      std::vector<uint16_t> part;
      const int nops = dr.nopsL;
      int r = dr.bestrL, c = dr.bestcL, k = 0, used = 0;                /* used: columns of run k taken by the first call */
      while (k < nops && c > endc) {
        const int op = ops[k] & 3, len = ops[k] >> 2;
        if (op == DPC_OP_M) {
          const int take = len < c - endc ? len : c - endc;
          part.push_back((uint16_t)((take << 2) | DPC_OP_M));
          r -= take; c -= take;
          if (take < len) { used = take; break; }                      /* the rest of this run belongs to the second call */
          k++;
          if (k < nops && (ops[k] & 3) != DPC_OP_M) {                   /* the gap run after the run's last column */
            part.push_back(ops[k]);
            if ((ops[k] & 3) == DPC_OP_QSKIP) r -= ops[k] >> 2; else c -= ops[k] >> 2;
            k++;
          }
        } else {
          part.push_back(ops[k]);
          if (op == DPC_OP_QSKIP) r -= len; else c -= len;
          k++;
        }
      }
      replay(sL, part.data(), (int)part.size(), dr.bestrL, dr.bestcL, qa, ga, p.offset1, p.offset2R, five, false, p.dynprogindex, true);
      sL.push(0, five ? p.offset2 - p.offset2R : p.offset2R - p.offset2, ' ', ' ', ' ', 0, 2);   /* 5518, 5977 */
      part.clear();
      if (used > 0) { part.push_back((uint16_t)((((ops[k] >> 2) - used) << 2) | DPC_OP_M)); k++; }
      for (; k < nops; k++) part.push_back(ops[k]);
      replay(sL, part.data(), (int)part.size(), r, c, qa, ga, p.offset1, p.offset2, five, false, p.dynprogindex, true);
      int first = 0;                                                   /* 5541-5544, 6000-6003 */
      while (first < sL.n && sL.p[first].comp == '-') first++;
      emit(out, sL.p + first, sL.n - first, five);                     /* end5: List_reverse again 5551; end3: as is 6011 */
      break;
    }
    case DPC_GENOME_GAP: {
      if (!(dr.status & DPC_ST_OK)) break;
      const int L1 = p.length1, L2L = p.length2, L2R = p.length2R, revoffset1 = p.offset1 + L1 - 1;
      char *qb = fit(s.qb, L1);
      const char *ga = staged_g, *gb = ga ? ga + dpc_gout_span(L2L) : NULL;
      for (int k = 0; k < L1; k++) qb[k] = q[L1 - 1 - k];
      if (!ga) {
        char *bufa = fit(s.ga, L2L), *bufb = fit(s.gb, L2R);
        gather_genome(p, blocks, p.offset2, L2L, false, bufa);
        gather_genome(p, blocks, p.offset2R, L2R, true, bufb);
        ga = bufa; gb = bufb;
      }
      Out sR, sL; sR.p = fit(s.sR, L1 + L2R + 2); sR.n = 0; sL.p = fit(s.sL, L1 + L2L + 2); sL.n = 0;
      replay(sR, ops + dr.nopsL, dr.nopsR, dr.bestrR, dr.bestcR, qb, gb, revoffset1, p.offset2R, true, false, p.dynprogindex, nostar);
      replay(sL, ops, dr.nopsL, dr.bestrL, dr.bestcL, q, ga, p.offset1, p.offset2, false, false, p.dynprogindex, nostar);
      if (sR.n + sL.n > 0) {                                           /* List_length == 1 -> NULL, 5051 */
        emit(out, sR.p, sR.n, true);
        out.push_gapholder();
        emit(out, sL.p, sL.n, false);
      }
      break;
    }
    case DPC_CDNA_GAP: {
      if (!(dr.status & DPC_ST_OK)) break;
      const int L1L = p.length1, L1R = p.length1R, L2 = p.length2, revoffset2 = p.offset2 + L2 - 1;
      const int span = p.offset1R - p.offset1 + 1;
      char *qb = fit(s.qb, L1R), *ga = fit(s.ga, L2), *gb = fit(s.gb, L2);
      for (int k = 0; k < L1R; k++) qb[k] = q[span - 1 - k];
      gather_genome(p, blocks, p.offset2, L2, false, ga);
      gather_genome(p, blocks, revoffset2, L2, true, gb);
      Out sR, sL; sR.p = fit(s.sR, L1R + L2 + 2); sR.n = 0; sL.p = fit(s.sL, L1L + L2 + 2); sL.n = 0;
      dpc_pair_t midbuf[24];
      Out mid; mid.p = midbuf; mid.n = 0;
      replay(sR, ops + dr.nopsL, dr.nopsR, dr.bestrR, dr.bestcR, qb, gb, p.offset1R, revoffset2, true, true, p.dynprogindex);
      int queryjump = (p.offset1R - dr.bestcR) - (p.offset1 + dr.bestcL) + 1;     /* 4725-4726 */
      int genomejump = (revoffset2 - dr.bestrR) - (p.offset2 + dr.bestrL) + 1;
      if (queryjump == 9 && genomejump == 9) {                         /* INSERT_PAIRS, 4730-4751 */
        for (int k = p.offset1R - dr.bestcR; k >= p.offset1 + dr.bestcL; k--)
          mid.push(k, revoffset2 - dr.bestrR + 1, q[k - p.offset1], '~', ' ', p.dynprogindex, 0);
        for (int k = revoffset2 - dr.bestrR; k >= p.offset2 + dr.bestrL; k--)
          mid.push(p.offset1 + dr.bestcL, k, ' ', '~', ga[k - p.offset2], p.dynprogindex, 0);
      } else {
        mid.push_gapholder();
      }
      replay(sL, ops, dr.nopsL, dr.bestrL, dr.bestcL, q, ga, p.offset1, p.offset2, false, true, p.dynprogindex);
      if (sR.n + mid.n + sL.n != 1) {                                  /* 4784-4787 */
        emit(out, sR.p, sR.n, true);
        emit(out, mid.p, mid.n, false);
        emit(out, sL.p, sL.n, false);
      }
      break;
    }
    default: break;
    }
    return out.n;
  }

  /* Turns the device record of problem i into the reference's output parameters. */
  void finalize(int i, const DevRes &dr, const uint16_t *ops, Scratch &s) {
    const dpc_problem_t &p = P(i);
    dpc_result_t &r = R(i);
    const HostProb &h = probs[i];
    int npairs = -1;     /* -1: count by rebuilding */
    /* pairs pushed = aligned columns that are not '*' + dashes + gapholders */
    int pushed = dr.nmatches + dr.nmismatches;
    for (int k = 0; k < dr.nopsL + dr.nopsR; k++) {
      int op = ops[k] & 3, len = ops[k] >> 2;
      if (op == DPC_OP_GSKIP || op == DPC_OP_QSKIP) pushed += len; else if (op == DPC_OP_GAPHOLDER) pushed += 1;
    }
    const bool star = (dr.status & DPC_ST_STAR) != 0;
    switch (p.kind) {
    case DPC_SINGLE_GAP:
      r.finalscore = dr.finalscore;
      r.nmatches = dr.nmatches; r.nmismatches = dr.nmismatches; r.nopens = dr.nopens; r.nindels = dr.nindels;
      r.dynprogindex_out = bump(p.dynprogindex);
      npairs = pushed;
      break;
    case DPC_END5_GAP: case DPC_END3_GAP:
      r.finalscore = dr.finalscore;
      r.nmatches = dr.nmatches; r.nmismatches = dr.nmismatches; r.nopens = dr.nopens; r.nindels = dr.nindels;
      r.dynprogindex_out = bump(p.dynprogindex);
      if ((p.endalign == DPC_QUERYEND_GAP || p.endalign == DPC_BEST_LOCAL) && dr.nmatches + 1 < dr.nmismatches) {
        r.finalscore = 0; npairs = 0;                                  /* 5259-5262 */
      } else if (!star) npairs = pushed;                               /* a leading '-' needs a skipped '*' column before it */
      break;
    case DPC_END5_SPLICEJUNCTION: case DPC_END3_SPLICEJUNCTION:
      r.nmatches = dr.nmatches; r.nmismatches = dr.nmismatches; r.nopens = dr.nopens; r.nindels = dr.nindels;
      r.finalscore = 3 * dr.nmatches - 5 * dr.nmismatches - 12 * dr.nopens - dr.nindels;      /* 5537, 5996 */
      r.dynprogindex_out = bump(p.dynprogindex);
      break;                                                           /* npairs: counted by rebuilding */
    case DPC_GENOME_GAP: {
      r.finalscore = dr.finalscore;
      const uint32_t dflags = dprobs[h.dev].flags;
      if (dflags & DPC_F_INTRONS) r.introntype = 0;                   /* *best_introntype = NONINTRON, 3695 */
      else r.introntype = ((dr.status & DPC_ST_HAVE) && !p.use_probabilities_p) ? dr.introntype : DPC_UNSET;
      npairs = 0;
      if (dr.status & DPC_ST_OK) {
        const uint8_t *known = (G().setup.splice_known != NULL)
            ? &pool[h.aux + ((dflags & DPC_F_PROBMODE) ? 8u * (uint32_t)(p.length2 + p.length2R + 2) : 0u)] : NULL;
        if (p.finalp) {                                                /* 4104-4108 */
          r.left_prob = site_prob(p, true, dr.bestcL, known && known[dr.bestcL]);
          r.right_prob = site_prob(p, false, dr.bestcR, known && known[p.length2 + 1 + dr.bestcR]);
        }
        r.new_leftgenomepos = p.offset2 + (dr.bestcL - 1);             /* 5000-5004 */
        r.new_rightgenomepos = p.offset2R - (dr.bestcR - 1);
        r.exonhead = (p.offset1 + p.length1 - 1) - (dr.bestrR - 1);
        r.nmatches = dr.nmatches; r.nmismatches = dr.nmismatches; r.nopens = dr.nopens; r.nindels = dr.nindels;
        r.dynprogindex_out = bump(p.dynprogindex);
        npairs = pushed > 0 ? pushed + 1 : 0;
      }
      break;
    }
    case DPC_CDNA_GAP: {
      r.finalscore = dr.finalscore;
      npairs = 0;
      if (dr.status & DPC_ST_OK) {
        int revoffset2 = p.offset2 + p.length2 - 1;
        int queryjump = (p.offset1R - dr.bestcR) - (p.offset1 + dr.bestcL) + 1;
        int genomejump = (revoffset2 - dr.bestrR) - (p.offset2 + dr.bestrL) + 1;
        int mid = 1;
        if (queryjump == 9 && genomejump == 9) mid = 18; else r.incompletep = 1;
        r.dynprogindex_out = bump(p.dynprogindex);
        npairs = pushed + mid == 1 ? 0 : pushed + mid;
      }
      break;
    }
    default: break;
    }
    if (npairs < 0) npairs = rebuild(i, dr, ops, fit(s.out, max_pairs(i)), s);
    r.npairs = npairs;
    r.null_list = npairs == 0;
  }
};

inline const char *strerror_(int code) {
  switch (code) {
  case DPC_OK: return "ok";
  case DPC_ERR_CUDA: return "CUDA device or driver error (there is no CPU fallback)";
  case DPC_ERR_ARG: return "malformed problem";
  case DPC_ERR_ALPHABET: return "query byte >= 128";
  case DPC_ERR_UNSUPPORTED: return "run too long for the traceback op format (QUERYEND_NOGAPS end of more than 622 554 columns)";
  case DPC_ERR_STATE: return "library not initialised / bad ticket / missing hook";
  case DPC_ERR_NOMEM: return "out of memory or output capacity too small";
  default: return "unknown error";
  }
}

}  // namespace dpc
#endif /* DPC_HOST_H */
