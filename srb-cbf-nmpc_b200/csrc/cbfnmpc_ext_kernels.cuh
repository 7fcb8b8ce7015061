// cbfnmpc_ext_kernels.cuh -- CUDA kernels of the extended mode (solver source: cbfnmpc_ext.h).  Included by cbfnmpc_api.cu.
#pragma once
#include "cbfnmpc_ext.h"

// Class lists over TEAMS (a warp runs one template instantiation): perm[0 .. n4) = teams whose members all stand on 4 feet,
// perm[n_teams .. n_teams + n2) = teams on a diagonal pair; counts = {n4, n2}.  Teams with a bad or mixed contact pattern are
// answered here (status 4 for every member).  One CTA, stable order.
__global__ void __launch_bounds__(1024, 1) cbfnmpc_ext_classify_kernel(const __grid_constant__ ExtIO io, int n_teams, int* __restrict__ perm,
                                                                       int* __restrict__ counts) {
  __shared__ int wcnt[2][32];
  __shared__ int base[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 2) base[threadIdx.x] = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_teams; t0 += 1024) {
    const int ti = t0 + threadIdx.x;
    // (fused runs in slices of teams, io.T == 1: item ti of this launch is instance t of the member-major batch)
    const int t = (io.xmode == 4 && io.slice_ns > 0) ? (ti / io.slice_ns) * io.n_teams + io.slice_t0 + ti % io.slice_ns : ti;
    const int cls = (ti < n_teams) ? ext_classify_team(io, t) : 0;
    if (ti < n_teams && cls < 0) {
      for (int m = 0; m < io.T; m++) ext_store_bad<4, 4>(io, t * io.T + m, 0);
      if (io.xmode == 4) ext_fused_publish_bad(io, t);   // its team mates (other CTAs, other GPUs) wait for its publications
    }
    const unsigned m4 = __ballot_sync(0xffffffffu, cls == 4), m2 = __ballot_sync(0xffffffffu, cls == 2);
    if (lane == 0) { wcnt[0][warp] = __popc(m4); wcnt[1][warp] = __popc(m2); }
    __syncthreads();
    if (warp < 2) {
      const int v = wcnt[warp][lane];
      int s = v;
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += u; }
      wcnt[warp][lane] = base[warp] + s - v;
      __syncwarp();
      if (lane == 31) base[warp] += s;
    }
    __syncthreads();
    const unsigned below = (1u << lane) - 1u;
    if (cls == 4) perm[wcnt[0][warp] + __popc(m4 & below)] = t;
    if (cls == 2) perm[n_teams + wcnt[1][warp] + __popc(m2 & below)] = t;
    __syncthreads();
  }
  if (threadIdx.x < 2) counts[threadIdx.x] = base[threadIdx.x];
}

// One warp per CTA; a warp holds 32 / (T W) teams of one stance class (W lanes per instance, T instances per team).
template <int W>
__global__ void __launch_bounds__(XNT, 1) cbfnmpc_ext_kernel(const __grid_constant__ XParams XP, const __grid_constant__ ExtIO io, int n_teams,
                                                             const int* __restrict__ perm, const int* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = (int)threadIdx.x;
  XPriv pr; pr.base = reinterpret_cast<double*>(smem_raw) + lane;
  const int lpt = io.T * W, tpw = XNT / lpt;   // lanes per team, teams per warp
  const int n4 = counts[0], n2 = counts[1];
  const int w4 = (n4 + tpw - 1) / tpw;
  const int warp = (int)blockIdx.x;
  const int member = (lane % lpt) / W;
  if (warp < w4) {
    const int slot = warp * tpw + lane / lpt;
    const bool valid = slot < n4;
    ext_run_group<4, W>(XP, io, valid ? perm[slot] * io.T + member : 0, valid, pr);
  } else {
    const int slot = (warp - w4) * tpw + lane / lpt;
    if ((warp - w4) * tpw >= n2) return;
    const bool valid = slot < n2;
    ext_run_group<2, W>(XP, io, valid ? perm[n_teams + slot] * io.T + member : 0, valid, pr);
  }
}

// The whole solve of agent-sharded teams in ONE launch (cbfnmpc_ext_solve_fused, io.xmode == 4): members publish into the mailboxes
// of their group's ranks (peer memory) and wait on sequence flags.  Every CTA of the grid must be resident at once -- a member
// waits for members that run in other CTAs, here and on the other GPUs -- which the host checks before it launches.
template <int W>
__global__ void __launch_bounds__(XNT, 1) cbfnmpc_ext_fused_kernel(const __grid_constant__ XParams XP, const __grid_constant__ ExtIO io, int n_teams,
                                                                   const int* __restrict__ perm, const int* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = (int)threadIdx.x;
  XPriv pr; pr.base = reinterpret_cast<double*>(smem_raw) + lane;
  const int tpw = XNT / W;   // instances per warp (io.T == 1: every instance is its own "team" of the class lists)
  const int n4 = counts[0], n2 = counts[1];
  const int w4 = (n4 + tpw - 1) / tpw;
  const int warp = (int)blockIdx.x;
  if (warp < w4) {
    const int slot = warp * tpw + lane / W;
    const bool valid = slot < n4;
    ext_run_group<4, W, true>(XP, io, valid ? perm[slot] : 0, valid, pr);
  } else {
    const int slot = (warp - w4) * tpw + lane / W;
    if ((warp - w4) * tpw >= n2) return;
    const bool valid = slot < n2;
    ext_run_group<2, W, true>(XP, io, valid ? perm[n_teams + slot] : 0, valid, pr);
  }
}

