// lane_prims.h -- primitives of the lane-per-stage solver: one NMPC instance = one group of 4 adjacent lanes
// (lane j of the group owns horizon stage j), 8 instances per warp.
//
// CUDA build (the product): width-4 shuffles / ballots restricted to the group's own lane mask, so that groups of one
// warp may diverge from each other (different iteration counts, inertia corrections, line-search lengths).
// CBF_HOST_EMU build (tests/emu only, never part of libcbfnmpc.so): the same names on 4 cooperative fibers, so the
// kernel source can be checked against the oracle on a CPU-only machine.  Not a fallback; the C ABI never reaches it.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef CBF_HOST_EMU
// ------------------------------------------------------------------------------------------ CUDA
#define LDEV __device__ __forceinline__
#define LMEM __device__ __forceinline__
#define LFULL 0xffffffffu
struct Grp {
  unsigned mask;  // the 4 lanes of this group
  int j;          // stage owned by this lane (0..3)
};
// (Making the thread index opaque to stop ptxas from re-reading SR_TID wherever it needs the mask again was tried: the registers
// it then holds cost more in spills than the S2R latency -- 0.746 instead of 0.72 ms per 4096 solves.)
LDEV int l_tid() { return (int)threadIdx.x; }
LDEV Grp grp_make() { Grp g; const int lane = l_tid() & 31; g.j = lane & 3; g.mask = 0xFu << (lane & ~3); return g; }
LDEV double g_idx(const Grp& g, double v, int src) { return __shfl_sync(g.mask, v, src, 4); }
LDEV double g_up(const Grp& g, double v) { return __shfl_up_sync(g.mask, v, 1, 4); }      // from lane j-1 (lane 0: own value)
LDEV double g_down(const Grp& g, double v) { return __shfl_down_sync(g.mask, v, 1, 4); }  // from lane j+1 (lane 3: own value)
LDEV double g_xor(const Grp& g, double v, int m) { return __shfl_xor_sync(g.mask, v, m, 4); }
LDEV unsigned g_xor_u(const Grp& g, unsigned v, int m) { return __shfl_xor_sync(g.mask, v, m, 4); }
// bit k = predicate of lane k of the group.  (Flags are exchanged with votes, never as converted doubles through shuffles:
// ptxas 12.9 drops one of the F2I conversions of a shuffled (double)flag sequence -- see DESIGN.md, "toolchain notes".)
LDEV unsigned g_bits(const Grp& g, bool p) { return (__ballot_sync(g.mask, p) >> ((threadIdx.x & 31) & ~3)) & 0xFu; }
LDEV bool g_any(const Grp& g, bool p) { return (__ballot_sync(g.mask, p) & g.mask) != 0u; }
LDEV bool g_all(const Grp& g, bool p) { return (__ballot_sync(g.mask, p) & g.mask) == g.mask; }
// Loop control ("does any group of this CTA still iterate").  The warps of a CTA are kept within ONE solver iteration of each
// other so that they fetch the same instruction-cache lines, but they do not wait for each other every iteration: round r ends
// with an arrival on mbarrier r & 1 and a wait for round r - 1 only.  A warp whose own groups have all finished has nothing to
// run, so it waits for round r itself and reads the other warps' flags of that round; all warps leave in the same round.
struct Pace {
#if defined(LANE_NT) && LANE_NT > 32
  unsigned bar, flg;  // shared-space addresses: 2 mbarriers (arrival count = warps per CTA), int flag[2][warps]
  int r, me;          // round counter, thread index
#endif
};
#if defined(LANE_NT) && LANE_NT > 32
LDEV void pace_wait(unsigned bar, int parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "PACE_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra PACE_DONE;\n\t"
      "bra PACE_WAIT;\n\t"
      "PACE_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// all threads of the CTA, once, before the first pace_go
LDEV Pace pace_make(unsigned long long* bars, int* flags) {
  Pace p;
  p.bar = (unsigned)__cvta_generic_to_shared(bars); p.flg = (unsigned)__cvta_generic_to_shared(flags); p.r = 0; p.me = l_tid();
  if (p.me == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p.bar), "r"(LANE_NT / 32) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p.bar + 8), "r"(LANE_NT / 32) : "memory");
  }
  __syncthreads();
  return p;
}
LDEV bool pace_go(Pace& p, bool act) {
  constexpr int NWARP = LANE_NT / 32;
  const bool mine = __any_sync(LFULL, act);
  const int r = p.r, k = r & 1;
  if (r > 0) pace_wait(p.bar + 8 * (k ^ 1), ((r - 1) >> 1) & 1);  // everybody has finished round r - 1
  if ((p.me & 31) == 0) {
    asm volatile("st.shared.s32 [%0], %1;" ::"r"(p.flg + 4 * (k * NWARP + (p.me >> 5))), "r"(mine ? 1 : 0) : "memory");
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(p.bar + 8 * k) : "memory");
  }
  p.r = r + 1;
  if (mine) return true;
  pace_wait(p.bar + 8 * k, (r >> 1) & 1);
  int any = 0;
#pragma unroll
  for (int w = 0; w < NWARP; w++) {
    int f;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(f) : "r"(p.flg + 4 * (k * NWARP + w)) : "memory");
    any |= f;
  }
  return any != 0;
}
#else
LDEV Pace pace_make(unsigned long long*, int*) { return Pace(); }
LDEV bool pace_go(Pace&, bool act) { return __any_sync(LFULL, act); }
#endif
LDEV double l_rcp(double v) {  // 1/v: MUFU.RCP64H seed + two Newton steps (~1 ulp for normal operands)
#if defined(CBF_EXACT_DIV) || defined(CBF_EXACT_RCP)
  return 1.0 / v;
#endif
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  r = fma(fma(-v, r, 1.0), r, r);
  r = fma(fma(-v, r, 1.0), r, r);
  return r;
}
LDEV double l_rsq(double v) {  // 1/sqrt(v)
#if defined(CBF_EXACT_DIV) || defined(CBF_EXACT_RSQ)
  return 1.0 / sqrt(v);
#endif
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  const double h = 0.5 * v;
  r = fma(fma(-h * r, r, 0.5), r, r);
  r = fma(fma(-h * r, r, 0.5), r, r);
  return r;
}
LDEV double l_mul_rn(double a, double b) { return __dmul_rn(a, b); }
LDEV double l_add_rn(double a, double b) { return __dadd_rn(a, b); }
// high word of a double / double from a high word: non-negative doubles order like their high words, which makes a 1-instruction
// (integer) running maximum good to 2^-20 relative -- enough for threshold tests
LDEV int l_hiword(double v) { return __double2hiint(v); }
LDEV double l_from_hiword(int h) { return __hiloint2double(h, 0); }
#ifdef CBF_DBGBUF  // debug builds only: group 0 of warp 0 logs numbers into a device buffer (lighter than printf)
__device__ double g_dbg[8192];
__device__ int g_dbg_n;
#define LANE_LOG(v) do { if (blockIdx.x == 0 && threadIdx.x == 0 && g_dbg_n < 8192) g_dbg[g_dbg_n++] = (v); } while (0)
#else
#define LANE_LOG(v) do { } while (0)
#endif
#ifdef CBF_TRACE  // debug builds only (scripts/dev/README.md): group 0 of warp 0 narrates
#include <stdio.h>
#define LANE_DBG(...) do { if (blockIdx.x == 0 && threadIdx.x == 0) printf(__VA_ARGS__); } while (0)
#else
#define LANE_DBG(...) do { } while (0)
#endif

#else
// ------------------------------------------------------------------------------------------ host emulation
#include <stdio.h>
#include <string.h>
// everything is forced inline so that each source-level shuffle site has its own call address: the harness checks that
// the 4 lanes of a group always meet at the SAME site (what __shfl_sync with the group mask requires on the device)
#define LDEV static inline __attribute__((always_inline))
#define LMEM inline __attribute__((always_inline))
namespace lemu {
struct GroupCtx {
  int cur;         // fiber currently running (0..3)
  double xd[4];    // exchange buffers
  uint64_t xi[4];
};
extern GroupCtx g_grp;
void barrier();  // all 4 fibers arrive before any leaves
}  // namespace lemu
struct Grp { unsigned mask; int j; };
LDEV Grp grp_make() { Grp g; g.j = lemu::g_grp.cur; g.mask = 0xFu; return g; }
LDEV double g_idx(const Grp& g, double v, int src) {
  lemu::g_grp.xd[g.j] = v; lemu::barrier();
  double r = lemu::g_grp.xd[src & 3]; lemu::barrier();
  return r;
}
LDEV double g_up(const Grp& g, double v) { return g_idx(g, v, g.j > 0 ? g.j - 1 : g.j); }
LDEV double g_down(const Grp& g, double v) { return g_idx(g, v, g.j < 3 ? g.j + 1 : g.j); }
LDEV double g_xor(const Grp& g, double v, int m) { return g_idx(g, v, g.j ^ m); }
LDEV unsigned g_xor_u(const Grp& g, unsigned v, int m) {
  lemu::g_grp.xi[g.j] = v; lemu::barrier();
  unsigned r = (unsigned)lemu::g_grp.xi[(g.j ^ m) & 3]; lemu::barrier();
  return r;
}
LDEV bool g_any(const Grp& g, bool p) {
  lemu::g_grp.xi[g.j] = p ? 1u : 0u; lemu::barrier();
  bool r = (lemu::g_grp.xi[0] | lemu::g_grp.xi[1] | lemu::g_grp.xi[2] | lemu::g_grp.xi[3]) != 0; lemu::barrier();
  return r;
}
LDEV bool g_all(const Grp& g, bool p) { return !g_any(g, !p); }
LDEV unsigned g_bits(const Grp& g, bool p) {
  lemu::g_grp.xi[g.j] = p ? 1u : 0u; lemu::barrier();
  unsigned r = (unsigned)(lemu::g_grp.xi[0] | (lemu::g_grp.xi[1] << 1) | (lemu::g_grp.xi[2] << 2) | (lemu::g_grp.xi[3] << 3)); lemu::barrier();
  return r;
}
struct Pace {};
LDEV bool pace_go(Pace&, bool act) { Grp g = grp_make(); return g_any(g, act); }  // the emulated CTA is one group
#ifdef CBF_EMU_GPUMATH  // mimic the CUDA build's rounding: low-precision seed + two Newton steps (not correctly rounded)
LDEV double l_rcp(double v) { double r = (double)(float)(1.0 / v); r = fma(fma(-v, r, 1.0), r, r); r = fma(fma(-v, r, 1.0), r, r); return r; }
LDEV double l_rsq(double v) { double r = (double)(float)(1.0 / sqrt(v)); const double h = 0.5 * v; r = fma(fma(-h * r, r, 0.5), r, r); r = fma(fma(-h * r, r, 0.5), r, r); return r; }
#else
LDEV double l_rcp(double v) { return 1.0 / v; }
LDEV double l_rsq(double v) { return 1.0 / sqrt(v); }
#endif
LDEV int l_hiword(double v) { uint64_t u; memcpy(&u, &v, 8); return (int)(u >> 32); }
LDEV double l_from_hiword(int h) { uint64_t u = (uint64_t)(uint32_t)h << 32; double v; memcpy(&v, &u, 8); return v; }
LDEV double l_mul_rn(double a, double b) { volatile double r = a * b; return r; }
LDEV double l_add_rn(double a, double b) { volatile double r = a + b; return r; }
#ifdef CBF_DBGBUF
extern double g_dbg[8192]; extern int g_dbg_n;
#define LANE_LOG(v) do { if (lemu::g_grp.cur == 0 && g_dbg_n < 8192) g_dbg[g_dbg_n++] = (v); } while (0)
#else
#define LANE_LOG(v) do { } while (0)
#endif
#ifdef CBF_TRACE
#define LANE_DBG(...) do { if (lemu::g_grp.cur == 0) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define LANE_DBG(...) do { } while (0)
#endif
#endif

LDEV int l_imax(int a, int b) { return a > b ? a : b; }
// ------------------------------------------------------------------------------------------ group reductions
LDEV double g_sum(const Grp& g, double v) { v += g_xor(g, v, 1); v += g_xor(g, v, 2); return v; }
LDEV double g_max(const Grp& g, double v) { v = fmax(v, g_xor(g, v, 1)); v = fmax(v, g_xor(g, v, 2)); return v; }
LDEV double g_min(const Grp& g, double v) { v = fmin(v, g_xor(g, v, 1)); v = fmin(v, g_xor(g, v, 2)); return v; }
LDEV unsigned g_or(const Grp& g, unsigned v) { v |= g_xor_u(g, v, 1); v |= g_xor_u(g, v, 2); return v; }
