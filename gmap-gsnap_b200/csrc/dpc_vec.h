/* dpc_vec.h -- "one value per lane" types for the warp-synchronous routines.
 *
 * Device build (nvcc): VI is an int, VM a bool; shuffles and ballots are the hardware's.
 * Host build (tests/emul only): VI / VM are arrays of 32 values and every operation loops over the
 * lanes, so a routine written once against this header runs as a 32-lane warp on the GPU and as a
 * lock-step simulation of that warp in the CPU test-suite.  Routines written this way must keep
 * their control flow uniform across lanes (lane-dependent choices go through vsel).
 */
#ifndef DPC_VEC_H
#define DPC_VEC_H
#include <stdint.h>

#ifdef __CUDACC__
namespace vec {
typedef int VI;
typedef bool VM;
#define DPC_V __device__ __forceinline__
DPC_V VI lane_index() { return (int)(threadIdx.x & 31); }
DPC_V VI splat(int x) { return x; }
DPC_V VI vsel(VM m, VI a, VI b) { return m ? a : b; }
DPC_V VI vmax(VI a, VI b) { return max(a, b); }
DPC_V VI vmin(VI a, VI b) { return min(a, b); }
DPC_V VI vmax3(VI a, VI b, VI c) { return __vimax3_s32(a, b, c); }
/* max(a, b) and, in the same instruction (VIMNMX with a predicate result), whether a >= b */
DPC_V VI vmax_ge(VI a, VI b, VM &ge) { return __vibmax_s32(a, b, &ge); }
DPC_V VM vand(VM a, VM b) { return a && b; }
DPC_V VM vor(VM a, VM b) { return a || b; }
DPC_V VM vnot(VM a) { return !a; }
DPC_V VM vlt_u(VI a, VI b) { return (unsigned)a < (unsigned)b; }
/* value of lane l-delta; lanes below delta get `fill` */
DPC_V VI shfl_up(VI v, int delta, VI fill) {
  int t = __shfl_up_sync(0xffffffffu, v, (unsigned)delta);
  return (int)(threadIdx.x & 31) < delta ? fill : t;
}
/* value of lane l-delta; lanes below delta keep their own (what the scan steps want) */
DPC_V VI shfl_up_keep(VI v, int delta) { return __shfl_up_sync(0xffffffffu, v, (unsigned)delta); }
/* value of lane l+1; lane 31 gets `fill` */
DPC_V VI shfl_down1(VI v, VI fill) {
  int t = __shfl_down_sync(0xffffffffu, v, 1u);
  return (threadIdx.x & 31) == 31 ? fill : t;
}
/* value of lane l+1; lane 31 keeps its own */
DPC_V VI shfl_down1_keep(VI v) { return __shfl_down_sync(0xffffffffu, v, 1u); }
/* value of lane l-1; lane 0 gets lane 31's */
DPC_V VI shfl_rot1(VI v) { return __shfl_sync(0xffffffffu, v, (int)((threadIdx.x + 31u) & 31u)); }
/* a per-lane byte pointer (base + idx), kept in a register across a loop; read at a uniform offset */
typedef const uint8_t *VP;
DPC_V VP vptr(const uint8_t *base, VI idx) { return base + idx; }
DPC_V VI load_u8p(VP p, int off) { return p[off]; }
typedef const int16_t *VPS;
DPC_V VPS vptr16(const int16_t *base, VI idx) { return base + idx; }
DPC_V VI load_i16p(VPS p, int off) { return p[off]; }
DPC_V int shfl_get(VI v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }
DPC_V uint32_t vballot(VM m) { return __ballot_sync(0xffffffffu, m); }
DPC_V VI load_u8(const uint8_t *base, VI idx) { return base[idx]; }
DPC_V VI load_i8(const int8_t *base, VI idx) { return base[idx]; }
DPC_V VI load_u32(const uint32_t *base, VI idx) { return (int)base[idx]; }
DPC_V VI load_i16(const int16_t *base, VI idx) { return base[idx]; }
DPC_V void store_i32(int32_t *base, VI idx, VI val, VM m) { if (m) base[idx] = val; }
DPC_V void store_i16(int16_t *base, VI idx, VI val, VM m) { if (m) base[idx] = (int16_t)val; }
/* lane 0 stores four consecutive words */
DPC_V void store4_lane0(uint32_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  if ((threadIdx.x & 31) == 0) *reinterpret_cast<uint4 *>(dst) = make_uint4(a, b, c, d);
}
DPC_V void store_u16_lane0(uint16_t *dst, int v) { if ((threadIdx.x & 31) == 0) *dst = (uint16_t)v; }
DPC_V void sync() { __syncwarp(); }
/* lane-wise "keep the better (score, key)" */
DPC_V void keep_better(VI &bs, VI &bk, VI s, VI k, VM cand, int late) {
  bool take = cand && (s > bs || (s == bs && (late ? k > bk : k < bk)));
  if (take) { bs = s; bk = k; }
}
DPC_V int lane0(VI v) { return __shfl_sync(0xffffffffu, v, 0); }
DPC_V int extract(VI v, int lane) { return __shfl_sync(0xffffffffu, v, lane); }
/* best (score, key) over the lanes, same rule as keep_better; every lane gets the result */
DPC_V void reduce_better(VI bs, VI bk, int late, int *s, int *k) {
  for (int o = 16; o > 0; o >>= 1) {
    int s2 = __shfl_xor_sync(0xffffffffu, bs, o), k2 = __shfl_xor_sync(0xffffffffu, bk, o);
    if (s2 > bs || (s2 == bs && (late ? k2 > bk : k2 < bk))) { bs = s2; bk = k2; }
  }
  *s = bs; *k = bk;
}
DPC_V int reduce_sum(VI v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
DPC_V int first_set(uint32_t b) { return __ffs((int)b) - 1; }
}  // namespace vec
#else
namespace vec {
#define DPC_V static inline
struct VI { int v[32]; };
struct VM { bool v[32]; };
#define DPC_VLOOP for (int l = 0; l < 32; l++)
DPC_V VI splat(int x) { VI r; DPC_VLOOP r.v[l] = x; return r; }
DPC_V VI lane_index() { VI r; DPC_VLOOP r.v[l] = l; return r; }
#define DPC_BIN(op) \
  DPC_V VI operator op(const VI &a, const VI &b) { VI r; DPC_VLOOP r.v[l] = a.v[l] op b.v[l]; return r; } \
  DPC_V VI operator op(const VI &a, int b) { VI r; DPC_VLOOP r.v[l] = a.v[l] op b; return r; } \
  DPC_V VI operator op(int a, const VI &b) { VI r; DPC_VLOOP r.v[l] = a op b.v[l]; return r; }
DPC_BIN(+) DPC_BIN(-) DPC_BIN(*) DPC_BIN(&) DPC_BIN(|) DPC_BIN(^) DPC_BIN(<<) DPC_BIN(>>)
#undef DPC_BIN
#define DPC_CMP(op) \
  DPC_V VM operator op(const VI &a, const VI &b) { VM r; DPC_VLOOP r.v[l] = a.v[l] op b.v[l]; return r; } \
  DPC_V VM operator op(const VI &a, int b) { VM r; DPC_VLOOP r.v[l] = a.v[l] op b; return r; } \
  DPC_V VM operator op(int a, const VI &b) { VM r; DPC_VLOOP r.v[l] = a op b.v[l]; return r; }
DPC_CMP(<) DPC_CMP(<=) DPC_CMP(>) DPC_CMP(>=) DPC_CMP(==) DPC_CMP(!=)
#undef DPC_CMP
DPC_V VI vsel(const VM &m, const VI &a, const VI &b) { VI r; DPC_VLOOP r.v[l] = m.v[l] ? a.v[l] : b.v[l]; return r; }
DPC_V VI vsel(const VM &m, const VI &a, int b) { return vsel(m, a, splat(b)); }
DPC_V VI vsel(const VM &m, int a, const VI &b) { return vsel(m, splat(a), b); }
DPC_V VI vsel(const VM &m, int a, int b) { return vsel(m, splat(a), splat(b)); }
DPC_V VI vmax(const VI &a, const VI &b) { VI r; DPC_VLOOP r.v[l] = a.v[l] > b.v[l] ? a.v[l] : b.v[l]; return r; }
DPC_V VI vmin(const VI &a, const VI &b) { VI r; DPC_VLOOP r.v[l] = a.v[l] < b.v[l] ? a.v[l] : b.v[l]; return r; }
DPC_V VI vmax(const VI &a, int b) { return vmax(a, splat(b)); }
DPC_V VI vmax3(const VI &a, const VI &b, const VI &c) { return vmax(vmax(a, b), c); }
DPC_V VI vmax_ge(const VI &a, const VI &b, VM &ge) { VI r; DPC_VLOOP { ge.v[l] = a.v[l] >= b.v[l]; r.v[l] = ge.v[l] ? a.v[l] : b.v[l]; } return r; }
DPC_V VI vmin(const VI &a, int b) { return vmin(a, splat(b)); }
DPC_V VM vand(const VM &a, const VM &b) { VM r; DPC_VLOOP r.v[l] = a.v[l] && b.v[l]; return r; }
DPC_V VM vor(const VM &a, const VM &b) { VM r; DPC_VLOOP r.v[l] = a.v[l] || b.v[l]; return r; }
DPC_V VM vnot(const VM &a) { VM r; DPC_VLOOP r.v[l] = !a.v[l]; return r; }
DPC_V VM vlt_u(const VI &a, const VI &b) { VM r; DPC_VLOOP r.v[l] = (unsigned)a.v[l] < (unsigned)b.v[l]; return r; }
DPC_V VM vlt_u(const VI &a, int b) { return vlt_u(a, splat(b)); }
DPC_V VI shfl_up(const VI &v, int delta, const VI &fill) { VI r; DPC_VLOOP r.v[l] = l < delta ? fill.v[l] : v.v[l - delta]; return r; }
DPC_V VI shfl_up(const VI &v, int delta, int fill) { return shfl_up(v, delta, splat(fill)); }
DPC_V VI shfl_up_keep(const VI &v, int delta) { VI r; DPC_VLOOP r.v[l] = l < delta ? v.v[l] : v.v[l - delta]; return r; }
DPC_V VI shfl_down1(const VI &v, const VI &fill) { VI r; DPC_VLOOP r.v[l] = l == 31 ? fill.v[l] : v.v[l + 1]; return r; }
DPC_V VI shfl_down1(const VI &v, int fill) { return shfl_down1(v, splat(fill)); }
DPC_V VI shfl_down1_keep(const VI &v) { VI r; DPC_VLOOP r.v[l] = l == 31 ? v.v[l] : v.v[l + 1]; return r; }
DPC_V VI shfl_rot1(const VI &v) { VI r; DPC_VLOOP r.v[l] = v.v[(l + 31) & 31]; return r; }
struct VP { const uint8_t *base; VI idx; };
DPC_V VP vptr(const uint8_t *base, const VI &idx) { VP p; p.base = base; p.idx = idx; return p; }
DPC_V VI load_u8p(const VP &p, int off) { VI r; DPC_VLOOP r.v[l] = p.base[p.idx.v[l] + off]; return r; }
struct VPS { const int16_t *base; VI idx; };
DPC_V VPS vptr16(const int16_t *base, const VI &idx) { VPS p; p.base = base; p.idx = idx; return p; }
DPC_V VI load_i16p(const VPS &p, int off) { VI r; DPC_VLOOP r.v[l] = p.base[p.idx.v[l] + off]; return r; }
DPC_V int shfl_get(const VI &v, int lane) { return v.v[lane & 31]; }
DPC_V uint32_t vballot(const VM &m) { uint32_t b = 0; DPC_VLOOP if (m.v[l]) b |= 1u << l; return b; }
DPC_V VI load_u8(const uint8_t *base, const VI &idx) { VI r; DPC_VLOOP r.v[l] = base[idx.v[l]]; return r; }
DPC_V VI load_i8(const int8_t *base, const VI &idx) { VI r; DPC_VLOOP r.v[l] = base[idx.v[l]]; return r; }
DPC_V VI load_u32(const uint32_t *base, const VI &idx) { VI r; DPC_VLOOP r.v[l] = (int)base[idx.v[l]]; return r; }
DPC_V VI load_i16(const int16_t *base, const VI &idx) { VI r; DPC_VLOOP r.v[l] = base[idx.v[l]]; return r; }
DPC_V void store_i32(int32_t *base, const VI &idx, const VI &val, const VM &m) { DPC_VLOOP if (m.v[l]) base[idx.v[l]] = val.v[l]; }
DPC_V void store_i16(int16_t *base, const VI &idx, const VI &val, const VM &m) { DPC_VLOOP if (m.v[l]) base[idx.v[l]] = (int16_t)val.v[l]; }
DPC_V void store4_lane0(uint32_t *dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d; }
DPC_V void store_u16_lane0(uint16_t *dst, int v) { *dst = (uint16_t)v; }
DPC_V void sync() {}
DPC_V void keep_better(VI &bs, VI &bk, const VI &s, const VI &k, const VM &cand, int late) {
  DPC_VLOOP {
    bool take = cand.v[l] && (s.v[l] > bs.v[l] || (s.v[l] == bs.v[l] && (late ? k.v[l] > bk.v[l] : k.v[l] < bk.v[l])));
    if (take) { bs.v[l] = s.v[l]; bk.v[l] = k.v[l]; }
  }
}
DPC_V int lane0(const VI &v) { return v.v[0]; }
DPC_V int extract(const VI &v, int lane) { return v.v[lane & 31]; }
DPC_V void reduce_better(const VI &bs, const VI &bk, int late, int *s, int *k) {
  int S = bs.v[0], K = bk.v[0];
  for (int l = 1; l < 32; l++)
    if (bs.v[l] > S || (bs.v[l] == S && (late ? bk.v[l] > K : bk.v[l] < K))) { S = bs.v[l]; K = bk.v[l]; }
  *s = S; *k = K;
}
DPC_V int reduce_sum(const VI &v) { int t = 0; DPC_VLOOP t += v.v[l]; return t; }
DPC_V int first_set(uint32_t b) { int f = 0; while (!((b >> f) & 1u)) f++; return f; }
#undef DPC_VLOOP
}  // namespace vec
#endif
#endif /* DPC_VEC_H */
