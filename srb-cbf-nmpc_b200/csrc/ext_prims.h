// ext_prims.h -- primitives of the extended-mode solver (cbfnmpc_ext.h): one NMPC instance = one group of W adjacent lanes
// (W = number of grid points N, 4 or 8; lane j of the group owns grid point j), a team = T groups next to each other in one warp.
//
// CUDA build (the product): shuffles / votes restricted to the group's or the team's own lane mask, so that groups (and teams)
// of one warp may diverge from each other.
// CBF_HOST_EMU build (tests/emu only, never part of libcbfnmpc.so): the same names on cooperative fibers, one fiber per lane of
// a warp, with mask-based rendezvous -- so the kernel source can be checked against the oracle on a CPU-only machine.
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef CBF_HOST_EMU
#define XDEV __device__ __forceinline__
#define XMEM __device__ __forceinline__
XDEV int x_lane() { return (int)(threadIdx.x & 31); }
XDEV double x_shfl(unsigned mask, double v, int src) { return __shfl_sync(mask, v, src); }
XDEV unsigned x_shfl_u(unsigned mask, unsigned v, int src) { return __shfl_sync(mask, v, src); }
XDEV unsigned x_ballot(unsigned mask, bool p) { return __ballot_sync(mask, p) & mask; }
XDEV double x_rcp(double v) {  // MUFU.RCP64H seed + two Newton steps
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  r = fma(fma(-v, r, 1.0), r, r);
  r = fma(fma(-v, r, 1.0), r, r);
  return r;
}
XDEV double x_rsq(double v) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
  const double h = 0.5 * v;
  r = fma(fma(-h * r, r, 0.5), r, r);
  r = fma(fma(-h * r, r, 0.5), r, r);
  return r;
}
#ifdef CBF_TRACE
#include <stdio.h>
#define XDBG(...) do { if (blockIdx.x == 0 && threadIdx.x == 0) printf(__VA_ARGS__); } while (0)
#else
#define XDBG(...) do { } while (0)
#endif
#else
// ------------------------------------------------------------------------------------------ host emulation
#include <stdio.h>
#include <string.h>
#define XDEV static inline __attribute__((always_inline))
#define XMEM inline __attribute__((always_inline))
namespace xemu {
struct WarpCtx {
  int cur;          // fiber (lane) currently running
  double xd[32];    // exchange buffers
  uint64_t xi[32];
};
extern WarpCtx g_warp;
void rendezvous(unsigned mask);  // every lane of `mask` arrives (with the same mask) before any of them leaves
void yield();                    // a lane that polls (the fused exchange's sequence flags) lets the other fibers run
}  // namespace xemu
XDEV int x_lane() { return xemu::g_warp.cur; }
XDEV double x_shfl(unsigned mask, double v, int src) {
  xemu::g_warp.xd[x_lane()] = v; xemu::rendezvous(mask);
  const double r = xemu::g_warp.xd[src & 31]; xemu::rendezvous(mask);
  return r;
}
XDEV unsigned x_shfl_u(unsigned mask, unsigned v, int src) {
  xemu::g_warp.xi[x_lane()] = v; xemu::rendezvous(mask);
  const unsigned r = (unsigned)xemu::g_warp.xi[src & 31]; xemu::rendezvous(mask);
  return r;
}
XDEV unsigned x_ballot(unsigned mask, bool p) {
  xemu::g_warp.xi[x_lane()] = p ? 1u : 0u; xemu::rendezvous(mask);
  unsigned r = 0;
  for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r |= (unsigned)(xemu::g_warp.xi[l] & 1u) << l;
  xemu::rendezvous(mask);
  return r;
}
XDEV double x_rcp(double v) { return 1.0 / v; }
XDEV double x_rsq(double v) { return 1.0 / sqrt(v); }
#ifdef CBF_TRACE
#define XDBG(...) do { if (xemu::g_warp.cur == 0) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define XDBG(...) do { } while (0)
#endif
#endif

// ------------------------------------------------------------------------------------------ group / team views
template <int W>
struct XGrp {
  unsigned mask;   // the W lanes of this instance
  unsigned tmask;  // the T * W lanes of its team (== mask when T == 1)
  int j;           // grid point owned by this lane
  int base;        // first lane of the group
  int tbase;       // first lane of the team
  int a;           // index of this instance inside its team
};
template <int W>
XDEV XGrp<W> xgrp_make(int T) {
  XGrp<W> g;
  const int lane = x_lane();
  g.j = lane & (W - 1); g.base = lane & ~(W - 1);
  g.mask = ((W == 32) ? 0xffffffffu : ((1u << W) - 1u)) << g.base;
  const int tw = T * W;  // 4, 8, 16 or 32
  g.tbase = (lane / tw) * tw; g.a = (lane - g.tbase) / W;
  g.tmask = ((tw >= 32) ? 0xffffffffu : ((1u << tw) - 1u)) << g.tbase;
  return g;
}
template <int W> XDEV double g_idx(const XGrp<W>& g, double v, int k) { return x_shfl(g.mask, v, g.base + k); }
template <int W> XDEV double g_up(const XGrp<W>& g, double v) { return x_shfl(g.mask, v, g.base + (g.j > 0 ? g.j - 1 : 0)); }          // from lane j-1 (lane 0: own)
template <int W> XDEV double g_down(const XGrp<W>& g, double v) { return x_shfl(g.mask, v, g.base + (g.j < W - 1 ? g.j + 1 : W - 1)); }  // from lane j+1 (last: own)
template <int W> XDEV double g_xor(const XGrp<W>& g, double v, int m) { return x_shfl(g.mask, v, g.base + (g.j ^ m)); }
template <int W> XDEV bool g_any(const XGrp<W>& g, bool p) { return x_ballot(g.mask, p) != 0u; }
template <int W> XDEV bool g_all(const XGrp<W>& g, bool p) { return x_ballot(g.mask, p) == g.mask; }
template <int W> XDEV unsigned g_bits(const XGrp<W>& g, bool p) { return x_ballot(g.mask, p) >> g.base; }
template <int W> XDEV double g_sum(const XGrp<W>& g, double v) {
#pragma unroll
  for (int m = 1; m < W; m <<= 1) v += g_xor(g, v, m);
  return v;
}
template <int W> XDEV double g_max(const XGrp<W>& g, double v) {
#pragma unroll
  for (int m = 1; m < W; m <<= 1) v = fmax(v, g_xor(g, v, m));
  return v;
}
template <int W> XDEV double g_min(const XGrp<W>& g, double v) {
#pragma unroll
  for (int m = 1; m < W; m <<= 1) v = fmin(v, g_xor(g, v, m));
  return v;
}
template <int W> XDEV bool t_any(const XGrp<W>& g, bool p) { return x_ballot(g.tmask, p) != 0u; }
// value of lane (agent b, grid point j) of this lane's team
template <int W> XDEV double t_from(const XGrp<W>& g, double v, int b) { return x_shfl(g.tmask, v, g.tbase + b * W + g.j); }
