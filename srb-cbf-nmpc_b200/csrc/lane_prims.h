// lane_prims.h -- primitives of the lane-per-stage solver: one NMPC instance = one group of 4 adjacent lanes
// (lane j of the group owns horizon stage j), 8 instances per warp.
//
// CUDA build (the product): width-4 shuffles / ballots.  Their member mask is NOT the group's own nibble: `Grp::mask` is the
// WARP-UNIFORM set of lanes that execute the current region (a vote result, held in a uniform register), so that a shuffle
// compiles to the bare SHFL behind one BRA.DIV.  With per-group masks ptxas has to prove at run time that the masks of the
// active lanes partition the active set: S2R + MATCH.ANY + REDUX.OR + LOP3 in front of every cluster of shuffles -- 14.5 % of
// the stall samples of the round-1 kernel (profiles/r02a_ncu_summary.md).  The price is a discipline: groups of one warp still
// differ in iteration counts, inertia corrections, refinement solves and line-search lengths.  (1) A region whose condition is
// not warp-uniform and that contains a shuffle or a vote is bracketed by g_enter / g_leave (the mask shrinks to the lanes that
// go in), and an early return is announced by g_keep right before it; no return inside a bracketed region.  (2) A LOOP that
// contains shuffles keeps the mask it was entered with and runs while ANY group of the mask needs another trip (w_any); a group
// that needs none runs through the extra trips as a zombie -- same instructions, every write to state that outlives the loop
// guarded by the flag next to the loop.  That costs no time (the lanes of a warp share one instruction stream anyway) and keeps
// the mask out of loop-carried registers: a mask that is the phi of a loop sits in a vector register and brings MATCH.ANY back
// (measured: narrowing at the back edges left the MATCH count of the hot loops unchanged).  A fully warp-synchronous variant
// (all 32 lanes always, finished groups as zombies of whole iterations, FULL mask) was measured too: no MATCH / REDUX at all
// and 14 % less code, but 4 % slower than this one at 4096 instances and 11 % slower for a lone instance (zombie lanes still
// cost shared-memory wavefronts and local-memory sectors).
// CBF_HOST_EMU build (tests/emu only, never part of libcbfnmpc.so): the same names on 32 cooperative fibers (one warp = 8
// groups) with mask-based rendezvous -- every lane named in the mask has to arrive with the same mask, and in the sitecheck
// build at the same source-level site, before any of them leaves; anything else is reported as a deadlock -- so the kernel
// source, including the mask discipline, is checked against the oracle on a CPU-only machine.  Not a fallback; the C ABI
// never reaches it.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef CBF_HOST_EMU
// ------------------------------------------------------------------------------------------ CUDA
#define LDEV __device__ __forceinline__
#define LMEM __device__ __forceinline__
#define LFULL 0xffffffffu
struct Grp {
  unsigned mask;  // warp-uniform: the lanes that execute the current region (whole groups only)
  int j;          // stage owned by this lane (0..3)
};
LDEV int l_tid() { return (int)threadIdx.x; }
LDEV Grp grp_make() { Grp g; g.j = l_tid() & 3; g.mask = LFULL; return g; }
LDEV unsigned l_gsh() { return (unsigned)l_tid() & 28u; }  // first lane of this lane's group
LDEV double g_idx(const Grp& g, double v, int src) { return __shfl_sync(g.mask, v, src, 4); }
LDEV double g_up(const Grp& g, double v) { return __shfl_up_sync(g.mask, v, 1, 4); }      // from lane j-1 (lane 0: own value)
LDEV double g_down(const Grp& g, double v) { return __shfl_down_sync(g.mask, v, 1, 4); }  // from lane j+1 (lane 3: own value)
LDEV double g_xor(const Grp& g, double v, int m) { return __shfl_xor_sync(g.mask, v, m, 4); }
LDEV unsigned g_xor_u(const Grp& g, unsigned v, int m) { return __shfl_xor_sync(g.mask, v, m, 4); }
// bit k = predicate of lane k of the group.  (Flags are exchanged with votes, never as converted doubles through shuffles:
// ptxas 12.9 drops one of the F2I conversions of a shuffled (double)flag sequence -- see DESIGN.md, "toolchain notes".)
LDEV unsigned g_bits(const Grp& g, bool p) { return (__ballot_sync(g.mask, p) >> l_gsh()) & 0xFu; }
LDEV bool g_any(const Grp& g, bool p) { return g_bits(g, p) != 0u; }
LDEV bool g_all(const Grp& g, bool p) { return g_bits(g, !p) == 0u; }
// mask discipline (see the header comment); all of these are executed by every lane of the current mask
LDEV unsigned w_ballot_full(bool p) { return __ballot_sync(LFULL, p); }                    // all 32 lanes are present
LDEV unsigned g_enter(Grp& g, bool p) { const unsigned sv = g.mask; g.mask = __ballot_sync(sv, p); return sv; }  // before `if (p) { ... }`
LDEV void g_leave(Grp& g, unsigned sv) { g.mask = sv; }                                    // after it
LDEV void g_keep(Grp& g, bool stay) { g.mask = __ballot_sync(g.mask, stay); }              // before `if (!stay) return`
LDEV bool w_any(const Grp& g, bool p) { return __ballot_sync(g.mask, p) != 0u; }           // over all groups of the current mask
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
LDEV bool pace_go(Pace& p, bool mine) {  // mine: some group of this warp still iterates (warp-uniform)
  constexpr int NWARP = LANE_NT / 32;
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
LDEV bool pace_go(Pace&, bool mine) { return mine; }
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
#define LFULL 0xffffffffu
namespace lemu {
struct WarpCtx {
  int cur;          // fiber (lane) currently running (0..31)
  double xd[32];    // exchange buffers
  uint64_t xi[32];
};
extern WarpCtx g_warp;
void rendezvous(unsigned mask);  // every lane of `mask` arrives (with the same mask, at the same site) before any of them leaves
}  // namespace lemu
struct Grp { unsigned mask; int j; };
LDEV int l_tid() { return lemu::g_warp.cur; }
LDEV Grp grp_make() { Grp g; g.j = l_tid() & 3; g.mask = LFULL; return g; }
LDEV unsigned l_gsh() { return (unsigned)l_tid() & 28u; }
LDEV double g_idx(const Grp& g, double v, int src) {
  lemu::g_warp.xd[l_tid()] = v; lemu::rendezvous(g.mask);
  double r = lemu::g_warp.xd[l_gsh() + (src & 3)]; lemu::rendezvous(g.mask);
  return r;
}
LDEV double g_up(const Grp& g, double v) { return g_idx(g, v, g.j > 0 ? g.j - 1 : g.j); }
LDEV double g_down(const Grp& g, double v) { return g_idx(g, v, g.j < 3 ? g.j + 1 : g.j); }
LDEV double g_xor(const Grp& g, double v, int m) { return g_idx(g, v, g.j ^ m); }
LDEV unsigned g_xor_u(const Grp& g, unsigned v, int m) {
  lemu::g_warp.xi[l_tid()] = v; lemu::rendezvous(g.mask);
  unsigned r = (unsigned)lemu::g_warp.xi[l_gsh() + ((g.j ^ m) & 3)]; lemu::rendezvous(g.mask);
  return r;
}
LDEV unsigned l_ballot(unsigned mask, bool p) {
  lemu::g_warp.xi[l_tid()] = p ? 1u : 0u; lemu::rendezvous(mask);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r |= (unsigned)(lemu::g_warp.xi[l] & 1u) << l;
  lemu::rendezvous(mask);
  return r;
}
LDEV unsigned g_bits(const Grp& g, bool p) { return (l_ballot(g.mask, p) >> l_gsh()) & 0xFu; }
LDEV bool g_any(const Grp& g, bool p) { return g_bits(g, p) != 0u; }
LDEV bool g_all(const Grp& g, bool p) { return g_bits(g, !p) == 0u; }
LDEV unsigned w_ballot_full(bool p) { return l_ballot(LFULL, p); }
LDEV unsigned g_enter(Grp& g, bool p) { const unsigned sv = g.mask; g.mask = l_ballot(sv, p); return sv; }
LDEV void g_leave(Grp& g, unsigned sv) { g.mask = sv; }
LDEV void g_keep(Grp& g, bool stay) { g.mask = l_ballot(g.mask, stay); }
LDEV bool w_any(const Grp& g, bool p) { return l_ballot(g.mask, p) != 0u; }
struct Pace {};
LDEV bool pace_go(Pace&, bool mine) { return mine; }  // the emulated CTA is one warp
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
#define LANE_LOG(v) do { if (lemu::g_warp.cur == 0 && g_dbg_n < 8192) g_dbg[g_dbg_n++] = (v); } while (0)
#else
#define LANE_LOG(v) do { } while (0)
#endif
#ifdef CBF_TRACE
#define LANE_DBG(...) do { if (lemu::g_warp.cur == 0) fprintf(stderr, __VA_ARGS__); } while (0)
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
