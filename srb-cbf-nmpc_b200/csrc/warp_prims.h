// warp_prims.h -- the few warp-collective primitives the solver uses.
//
// CUDA build (the product): thin wrappers over shuffles / redux / ballots of one 32-lane warp.
// CBF_HOST_EMU build (tests/emu only, never part of libcbfnmpc.so): the same names implemented on 32
// cooperative fibers so the kernel source can be single-stepped on a CPU-only machine.  The emulation is a
// debugging aid for the kernel *logic*; it is not a fallback and the C ABI never reaches it.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef CBF_HOST_EMU
// ------------------------------------------------------------------------------------------ CUDA
#define DEV __device__ __forceinline__
#define DEVNI __device__ __noinline__
#define FULL 0xffffffffu

DEV int lane_id() { return threadIdx.x & 31; }
DEV void wsync() { __syncwarp(FULL); }
DEV double shfl_xor(double v, int m) { return __shfl_xor_sync(FULL, v, m); }
DEV double shfl_idx(double v, int src) { return __shfl_sync(FULL, v, src); }
// value of lane (lane & ~3) + k : exchange inside a group of 4 lanes
DEV double shfl_g4(double v, int k) { return __shfl_sync(FULL, v, k, 4); }
DEV int shfl_idx_i(int v, int src) { return __shfl_sync(FULL, v, src); }
DEV bool wany(bool p) { return __any_sync(FULL, p); }
DEV bool wall(bool p) { return __all_sync(FULL, p); }
DEV unsigned wballot(bool p) { return __ballot_sync(FULL, p); }

DEV double wsum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
  return v;
}
// max / min of NON-NEGATIVE doubles (or +inf; NaN propagates as "largest"): two 32-bit redux on the bit pattern
DEV double wmax_nn(double v) {
  unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  unsigned mh = __reduce_max_sync(FULL, hi);
  unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
  return __hiloint2double((int)mh, (int)ml);
}
DEV double wmin_nn(double v) {
  unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  unsigned mh = __reduce_min_sync(FULL, hi);
  unsigned ml = __reduce_min_sync(FULL, hi == mh ? lo : 0xffffffffu);
  return __hiloint2double((int)mh, (int)ml);
}
DEV int wmin_i(int v) { return __reduce_min_sync(FULL, v); }
DEV unsigned wor(unsigned v) { return __reduce_or_sync(FULL, v); }
// CTA-wide barrier + OR of a predicate: keeps the warps of a CTA in lockstep (one IPM iteration per round)
DEV bool cta_or(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }
// 1/v and 1/sqrt(v) for normal, positive-range operands: hardware seed (MUFU.RCP64H / RSQ64H) + two Newton steps (~1 ulp),
// ~6 instructions instead of the ~20-instruction IEEE sequences with their slow-path calls.
DEV double rcp(double v) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  r = fma(fma(-v, r, 1.0), r, r);
  r = fma(fma(-v, r, 1.0), r, r);
  return r;
}
DEV double rsq(double v) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  double h = 0.5 * v;
  r = fma(fma(-h * r, r, 0.5), r, r);
  r = fma(fma(-h * r, r, 0.5), r, r);
  return r;
}
DEV double mul_rn(double a, double b) { return __dmul_rn(a, b); }
DEV double add_rn(double a, double b) { return __dadd_rn(a, b); }

#else
// ------------------------------------------------------------------------------------------ host emulation
#include <string.h>
#define DEV static inline
#define DEVNI static
#define FULL 0xffffffffu
namespace emu {
struct WarpCtx {
  int cur;            // fiber currently running
  double xd[32];      // exchange buffers
  uint64_t xi[32];
  void (*yield_fn)();
};
extern WarpCtx g_warp;
void barrier();  // all 32 lanes arrive before any leaves
}  // namespace emu
DEV int lane_id() { return emu::g_warp.cur; }
DEV void wsync() { emu::barrier(); }
DEV double shfl_idx(double v, int src) {
  emu::g_warp.xd[lane_id()] = v; emu::barrier();
  double r = emu::g_warp.xd[src & 31]; emu::barrier();
  return r;
}
DEV double shfl_xor(double v, int m) { return shfl_idx(v, lane_id() ^ m); }
DEV double shfl_g4(double v, int k) { return shfl_idx(v, (lane_id() & ~3) + (k & 3)); }
DEV int shfl_idx_i(int v, int src) { return (int)shfl_idx((double)v, src); }
DEV unsigned wballot(bool p) {
  emu::g_warp.xi[lane_id()] = p ? 1u : 0u; emu::barrier();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)emu::g_warp.xi[i] << i;
  emu::barrier();
  return r;
}
DEV bool wany(bool p) { return wballot(p) != 0u; }
DEV bool wall(bool p) { return wballot(p) == 0xffffffffu; }
DEV double wsum(double v) {  // same butterfly order as the device version
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
DEV double wmax_nn(double v) {
  emu::g_warp.xd[lane_id()] = v; emu::barrier();
  uint64_t best = 0;
  for (int i = 0; i < 32; i++) { uint64_t b; memcpy(&b, &emu::g_warp.xd[i], 8); if (b > best) best = b; }
  emu::barrier();
  double r; memcpy(&r, &best, 8); return r;
}
DEV double wmin_nn(double v) {
  emu::g_warp.xd[lane_id()] = v; emu::barrier();
  uint64_t best = ~0ull;
  for (int i = 0; i < 32; i++) { uint64_t b; memcpy(&b, &emu::g_warp.xd[i], 8); if (b < best) best = b; }
  emu::barrier();
  double r; memcpy(&r, &best, 8); return r;
}
DEV int wmin_i(int v) {
  emu::g_warp.xi[lane_id()] = (uint64_t)(int64_t)v; emu::barrier();
  int best = 0x7fffffff;
  for (int i = 0; i < 32; i++) { int b = (int)(int64_t)emu::g_warp.xi[i]; if (b < best) best = b; }
  emu::barrier();
  return best;
}
DEV unsigned wor(unsigned v) {
  emu::g_warp.xi[lane_id()] = v; emu::barrier();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)emu::g_warp.xi[i];
  emu::barrier();
  return r;
}
DEV bool cta_or(bool p) { return wany(p); }  // the emulated CTA is a single warp
DEV double rcp(double v) { return 1.0 / v; }
DEV double rsq(double v) { return 1.0 / sqrt(v); }
DEV double mul_rn(double a, double b) { volatile double r = a * b; return r; }
DEV double add_rn(double a, double b) { volatile double r = a + b; return r; }
#endif
