// cbfnmpc_api.cu -- C ABI (include/cbfnmpc.h) + the CUDA kernels.  sm_100a only; no CPU path of any kind.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include "cbfnmpc_host.h"
#ifndef LANE_NT
#define LANE_NT 64  // 2 warps (16 instances, one stance class) per CTA, 2 CTAs per SM; the warps of a CTA re-synchronise every solver iteration
#endif
#include "cbfnmpc_lane.h"
#include "cbfnmpc_qpdense.h"
#include "cbfnmpc_ext_kernels.cuh"

// ---------------------------------------------------------------------------------------------------------
// The solver (cbfnmpc_lane.h): one instance = 4 adjacent lanes, 8 instances per warp, LANE_NT / 32 warps per CTA.
// A pre-pass sorts the instances by stance class (4 feet / diagonal pair) so that every warp runs one template
// instantiation; the order inside a class is the batch order (deterministic).
// ---------------------------------------------------------------------------------------------------------
// perm[0 .. n4) = instances with 4 stance feet, perm[B .. B + n2) = instances with a diagonal pair; counts = {n4, n2}.
// One CTA; bad contact patterns are answered here (status 4).
__global__ void __launch_bounds__(1024, 1) cbfnmpc_classify_kernel(const __grid_constant__ BatchIO io, int* __restrict__ perm,
                                                                   int* __restrict__ counts) {
  __shared__ int wcnt[2][32];
  __shared__ int base[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("griddepcontrol.launch_dependents;");   // a solve kernel launched as programmatic dependent may be placed now (launch_on)
  if (threadIdx.x < 2) base[threadIdx.x] = 0;
  __syncthreads();
  for (int b0 = 0; b0 < io.B; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int cls = (b < io.B) ? lane_classify(io, b) : 0;
    if (b < io.B && cls < 0) lane_store_bad(io, b);
    const unsigned m4 = __ballot_sync(0xffffffffu, cls == 4), m2 = __ballot_sync(0xffffffffu, cls == 2);
    if (lane == 0) { wcnt[0][warp] = __popc(m4); wcnt[1][warp] = __popc(m2); }
    __syncthreads();
    if (warp < 2) {  // exclusive scan of the 32 per-warp counts of class `warp`
      const int v = wcnt[warp][lane];
      int s = v;
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
      wcnt[warp][lane] = base[warp] + s - v;
      __syncwarp();
      if (lane == 31) base[warp] += s;
    }
    __syncthreads();
    const unsigned below = (1u << lane) - 1u;
    if (cls == 4) perm[wcnt[0][warp] + __popc(m4 & below)] = b;
    if (cls == 2) perm[io.B + wcnt[1][warp] + __popc(m2 & below)] = b;
    __syncthreads();
  }
  if (threadIdx.x < 2) counts[threadIdx.x] = base[threadIdx.x];
}

#ifdef CBF_CTATIME  // debug builds only (scripts/dev/cta_times.py): duration and SM of every warp of the last cbfnmpc_lane_kernel launch
__device__ long long g_ctat[3 * 2 * 8192];
#endif
__global__ void __launch_bounds__(LANE_NT, 1) cbfnmpc_lane_kernel(const __grid_constant__ KParams P, const __grid_constant__ BatchIO io,
                                                                  const int* __restrict__ perm, const int* __restrict__ counts, int sm_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = l_tid();
  Priv pr; pr.base = reinterpret_cast<double*>(smem_raw) + tid;
  constexpr int IPC = LANE_NT / 4;  // instances per CTA; a CTA runs one stance class
  __shared__ unsigned long long pace_bar[2];
  __shared__ int pace_flag[2 * ((LANE_NT + 31) / 32)];
  const int grp = tid >> 2;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // class lists of the preceding cbfnmpc_classify_kernel (no-op for an ordinary launch)
  const int n4 = counts[0], n2 = counts[1];
  const int c4 = (n4 + IPC - 1) / IPC;
  // The first 2 x (number of SMs) CTAs are resident at once and are handed out breadth-first over the SMs, so blockIdx b and
  // b + nsm share SM b: give those two the neighbouring work items 2b, 2b + 1 (same stance class except at the class boundary)
  // so that the two CTAs of an SM share instruction-cache lines.  SMs that get a single CTA take the items after the pairs.
  int cta = blockIdx.x;
#ifdef CBF_CTATIME
  long long ctat0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ctat0));
#endif
  {
    const int nsm = sm_count, total = (int)gridDim.x, first = total < 2 * nsm ? total : 2 * nsm;
    const int npair = first > nsm ? first - nsm : 0;
    if (cta < nsm) cta = (cta < npair) ? 2 * cta : 2 * npair + (cta - npair);
    else if (cta < first) cta = 2 * (cta - nsm) + 1;
  }
  if (cta < c4) {
    const int slot = cta * IPC + grp;
    const bool valid = slot < n4;
    lane_run_group<4>(P, io, valid ? perm[slot] : 0, valid, pr, pace_make(pace_bar, pace_flag));
  } else {
    const int slot = (cta - c4) * IPC + grp;
    if ((cta - c4) * IPC >= n2) return;
    const bool valid = slot < n2;
    lane_run_group<2>(P, io, valid ? perm[io.B + slot] : 0, valid, pr, pace_make(pace_bar, pace_flag));
  }
#ifdef CBF_CTATIME
  __syncwarp();
  if ((tid & 31) == 0 && blockIdx.x < 8192) {
    long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long* o = g_ctat + 3 * (2 * (size_t)blockIdx.x + (tid >> 5));
    o[0] = t1 - ctat0; o[1] = cta; o[2] = smid;
  }
#endif
}

// The same kernel for host-buffer calls: the obstacles are looked at after the QP stage, once their upload (second stream) has landed.
__global__ void __launch_bounds__(LANE_NT, 1) cbfnmpc_lane_kernel_late(const __grid_constant__ KParams P, const __grid_constant__ BatchIO io,
                                                                  const int* __restrict__ perm, const int* __restrict__ counts, int sm_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = l_tid();
  Priv pr; pr.base = reinterpret_cast<double*>(smem_raw) + tid;
  constexpr int IPC = LANE_NT / 4;  // instances per CTA; a CTA runs one stance class
  __shared__ unsigned long long pace_bar[2];
  __shared__ int pace_flag[2 * ((LANE_NT + 31) / 32)];
  const int grp = tid >> 2;
  asm volatile("griddepcontrol.wait;" ::: "memory");   // class lists of the preceding cbfnmpc_classify_kernel (no-op for an ordinary launch)
  const int n4 = counts[0], n2 = counts[1];
  const int c4 = (n4 + IPC - 1) / IPC;
  // The first 2 x (number of SMs) CTAs are resident at once and are handed out breadth-first over the SMs, so blockIdx b and
  // b + nsm share SM b: give those two the neighbouring work items 2b, 2b + 1 (same stance class except at the class boundary)
  // so that the two CTAs of an SM share instruction-cache lines.  SMs that get a single CTA take the items after the pairs.
  int cta = blockIdx.x;
  {
    const int nsm = sm_count, total = (int)gridDim.x, first = total < 2 * nsm ? total : 2 * nsm;
    const int npair = first > nsm ? first - nsm : 0;
    if (cta < nsm) cta = (cta < npair) ? 2 * cta : 2 * npair + (cta - npair);
    else if (cta < first) cta = 2 * (cta - nsm) + 1;
  }
  if (cta < c4) {
    const int slot = cta * IPC + grp;
    const bool valid = slot < n4;
    lane_run_group<4, true>(P, io, valid ? perm[slot] : 0, valid, pr, pace_make(pace_bar, pace_flag));
  } else {
    const int slot = (cta - c4) * IPC + grp;
    if ((cta - c4) * IPC >= n2) return;
    const bool valid = slot < n2;
    lane_run_group<2, true>(P, io, valid ? perm[io.B + slot] : 0, valid, pr, pace_make(pace_bar, pace_flag));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Input packing: updateState + copPlanner_eventbase + footholdsPlanner for B agents, one thread per
// agent.  Pure data movement (HBM-bound, ~0.9 KB per agent); every store is coalesced (batch index fastest).
// ---------------------------------------------------------------------------------------------------------
struct PackIO {
  cbfnmpc_raw_in r;
  double *x0, *xdes, *feet, *obs;
  int* contact;
  int B;
  double Ts;
};
__global__ void __launch_bounds__(256) cbfnmpc_pack_kernel(const __grid_constant__ PackIO io) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= io.B) return;
  const size_t B = (size_t)io.B;
  const cbfnmpc_raw_in& r = io.r;
  // measured LIP state (q[0], dq[0], q[1], dq[1])                                      src/MPC_dist.cpp:226-229
  io.x0[b] = r.q[b]; io.x0[B + b] = r.dq[b]; io.x0[2 * B + b] = r.q[B + b]; io.x0[3 * B + b] = r.dq[B + b];
  const int g = r.gait_domain[b];
  int a = r.agent_id[b], t = r.team[b];
  // a team or agent index outside the arrays is bad input: nothing is read out of bounds and the instance is answered
  // with status 4 (the contact pattern written below classifies as invalid)
  const bool bad_index = t < 0 || t >= r.n_teams || a < 0 || a >= 4;
  if (bad_index) { a = 0; t = 0; }
  // desired trajectory: columns [4 g, 4 g + 4) of rows 2a, 2a+1 of Pr_refined_ / Prd_refined_, vectorised column-major
  // as [x, xd, y, yd] x 4                                                               :713-718, :766-780
  const double* Pr = r.Pr_refined + (size_t)t * 8 * r.ref_cols;
  const double* Prd = r.Prd_refined + (size_t)t * 8 * r.ref_cols;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    int col = 4 * g + k;
    col = col < r.ref_cols ? col : r.ref_cols - 1;
    col = col < 0 ? 0 : col;
    io.xdes[(size_t)(4 * k + 0) * B + b] = Pr[2 * a + 8 * (size_t)col];
    io.xdes[(size_t)(4 * k + 1) * B + b] = Prd[2 * a + 8 * (size_t)col];
    io.xdes[(size_t)(4 * k + 2) * B + b] = Pr[2 * a + 1 + 8 * (size_t)col];
    io.xdes[(size_t)(4 * k + 3) * B + b] = Prd[2 * a + 1 + 8 * (size_t)col];
  }
  // stance feet: nominal footprint about agent_Initial_ during the first two domains, measured toes afterwards   :1204-1216
  const double fpx[4] = {0.2188, 0.2188, -0.1472, -0.1472}, fpy[4] = {-0.1320, 0.1320, -0.1320, 0.1320};
  const double ax = r.agent_initial[b], ay = r.agent_initial[B + b];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    io.feet[(size_t)i * B + b] = (g <= 1) ? fpx[i] + ax : r.toe_pos[(size_t)(3 * i) * B + b];
    io.feet[(size_t)(4 + i) * B + b] = (g <= 1) ? fpy[i] + ay : r.toe_pos[(size_t)(3 * i + 1) * B + b];
    io.contact[(size_t)i * B + b] = bad_index ? -1 : r.contact[(size_t)i * B + b];
  }
  // real obstacles of the agent's team (Pobs_real, 2 x n_obs column-major) -> x of all obstacles, then y          :371-396
  const double* Po = r.Pobs_real + (size_t)t * 2 * r.n_obs;
  const int no = (r.state_other != nullptr) ? r.n_other : 0, K = r.n_obs + no;
  for (int k = 0; k < r.n_obs; k++) {
    io.obs[(size_t)k * B + b] = Po[2 * k];
    io.obs[(size_t)(K + k) * B + b] = Po[2 * k + 1];
  }
  // the other agents, extrapolated over the horizon, as obstacle candidates (the reference's disabled rule)            :391
  for (int k = 0; k < no; k++) {
    const double* so = r.state_other + (size_t)(4 * k) * B + b;
    // separately rounded products and sum (no FMA contraction): the same bits as the host-side mirrors
    io.obs[(size_t)(r.n_obs + k) * B + b] = __dadd_rn(so[0], __dmul_rn(__dmul_rn(so[2 * B], 4.0), io.Ts));
    io.obs[(size_t)(K + r.n_obs + k) * B + b] = __dadd_rn(so[B], __dmul_rn(__dmul_rn(so[3 * B], 4.0), io.Ts));
  }
}

// state_other of team-major batches: neighbour k of agent a = k-th other member of its team
__global__ void __launch_bounds__(256) cbfnmpc_team_states_kernel(int B, int T, const double* __restrict__ q, const double* __restrict__ dq,
                                                                  double* __restrict__ so) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const size_t Bs = (size_t)B;
  const int base = (b / T) * T, me = b - base;
  for (int k = 0; k < T - 1; k++) {
    const int o = base + (k < me ? k : k + 1);
    so[(size_t)(4 * k) * Bs + b] = q[o]; so[(size_t)(4 * k + 1) * Bs + b] = q[Bs + o];
    so[(size_t)(4 * k + 2) * Bs + b] = dq[o]; so[(size_t)(4 * k + 3) * Bs + b] = dq[Bs + o];
  }
}

// ---------------------------------------------------------------------------------------------------------
// World update of the closed loop: one thread per world.
// ---------------------------------------------------------------------------------------------------------
struct RollIO {
  cbfnmpc_world_state w;
  const double *X, *abuf_next, *agent_initial, *Pobs_real;
  const int *status, *team;
  int B, n_obs, cycle, n_teams;
  double fail_radius;
};
__global__ void __launch_bounds__(256) cbfnmpc_rollout_kernel(const __grid_constant__ RollIO io) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= io.B) return;
  const size_t B = (size_t)io.B;
  if (io.status[b] != CBFNMPC_ST_CONVERGED) io.w.n_not_converged[b] += 1;
  if (io.status[b] == CBFNMPC_ST_BADINPUT) return;  // nothing was solved (X, alpha_buf_next are not written): the world stays where it is
  // updateDistance_to_fail (src/MPC_dist.cpp:21-40) at the 4 grid points of this domain
  if (io.w.success[b]) {
    const int tm = io.team[b] < 0 ? 0 : (io.team[b] >= io.n_teams ? io.n_teams - 1 : io.team[b]);
    const double* Po = io.Pobs_real + (size_t)tm * 2 * io.n_obs;
    bool failed = false;
    for (int k = 0; k < 4 && !failed; k++) {
      const double px = io.X[(size_t)(4 * k) * B + b], py = io.X[(size_t)(4 * k + 2) * B + b];
      for (int i = 0; i < io.n_obs; i++) {
        const double dx = px - Po[2 * i], dy = py - Po[2 * i + 1];
        if (sqrt(dx * dx + dy * dy) < io.fail_radius) {
          failed = true;
          io.w.success[b] = 0;
          io.w.distance_to_fail[b] = sqrt(px * px + py * py);
          break;
        }
      }
    }
  }
  // next measured state = x_4 (end of the gait domain)
  const double x = io.X[12 * B + b], xd = io.X[13 * B + b], y = io.X[14 * B + b], yd = io.X[15 * B + b];
  io.w.q[b] = x; io.w.dq[b] = xd; io.w.q[B + b] = y; io.w.dq[B + b] = yd;
  if (io.w.traj) {
    double* t = io.w.traj + (size_t)io.cycle * 4 * B;
    t[b] = x; t[B + b] = xd; t[2 * B + b] = y; t[3 * B + b] = yd;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) io.w.alpha_buf[(size_t)i * B + b] = io.abuf_next[(size_t)i * B + b];  // src/MPC_dist.cpp:798
  const int g = io.w.gait_domain[b] + 1;                                                              // :452
  io.w.gait_domain[b] = g;
  // stance feet under the body, trot contact sequence
  const double fpx[4] = {0.2188, 0.2188, -0.1472, -0.1472}, fpy[4] = {-0.1320, 0.1320, -0.1320, 0.1320};  // :1206-1209
#pragma unroll
  for (int i = 0; i < 4; i++) {
    io.w.toe_pos[(size_t)(3 * i) * B + b] = x + fpx[i];
    io.w.toe_pos[(size_t)(3 * i + 1) * B + b] = y + fpy[i];
    io.w.toe_pos[(size_t)(3 * i + 2) * B + b] = 0.0;
    const int diag = (i == 0 || i == 3) ? 1 : 0;       // FR, RL
    io.w.contact[(size_t)i * B + b] = (g <= 1) ? 1 : ((g & 1) ? diag : 1 - diag);
  }
}

// ---------------------------------------------------------------------------------------------------------
// generateReferenceTrajectory (src/MPC_dist.cpp:930-1062), SURVEY.md 8f.4: sequential in time, parallel over teams.
// One team = 4 adjacent lanes (lane k = agent k); team mates' positions travel by width-4 shuffles; all lanes of a warp run
// the same `loop` steps, so full-mask shuffles are legal.  Obstacles of the CTA's teams sit in shared memory.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) cbfnmpc_refgen_kernel(int n_teams, const double* __restrict__ Pstart, const double* __restrict__ Pobs,
                                                             int n_obs, double goal_x, double goal_y, int loop, double* __restrict__ Pr,
                                                             double* __restrict__ Prd) {
  extern __shared__ double s_obs[];  // [32 teams][2 n_obs]
  const int tl = threadIdx.x >> 2, k = threadIdx.x & 3;
  const int team = blockIdx.x * 32 + tl;
  const bool valid = team < n_teams;
  const int tc = valid ? team : n_teams - 1;   // idle lanes shadow the last team (they must execute the shuffles)
  for (int i = k; i < 2 * n_obs; i += 4) s_obs[tl * 2 * n_obs + i] = Pobs[(size_t)tc * 2 * n_obs + i];
  __syncthreads();
  const double* ob = s_obs + tl * 2 * n_obs;
  const double alpha = 150.0, eta = 400.0, dmin = 1.0, epsilon = 60.0;                          // :932-933 (sigma = 1)
  const double a13 = 0.001025061239872, a33 = 0.929527039758809;                                // Ad, :942-945
  const double b1 = 0.000000068070472960, b3 = 0.000128132654983983;                            // Bd, :957-960
  const int ramp = 1000, dec = 40, ncols = loop / dec;
  double x = Pstart[(size_t)tc * 8 + 2 * k], y = Pstart[(size_t)tc * 8 + 2 * k + 1], vx = 0.0, vy = 0.0;
  double* pr = Pr + (size_t)tc * 8 * ncols + 2 * k;
  double* prd = Prd + (size_t)tc * 8 * ncols + 2 * k;
  for (int i = 0; i <= loop; i++) {
    // output column c holds column c + 2 of the decimated history: step 40 (c + 2) for c + 2 < ncols, else step c + 2
    // (a step can be both: the two stores are independent)
    const int c1 = (i % dec == 0 && i / dec < ncols) ? i / dec - 2 : -1;
    const int c2 = (i == ncols || i == ncols + 1) ? i - 2 : -1;
    if (valid && c1 >= 0 && c1 < ncols) {
      pr[8 * (size_t)c1] = x; pr[8 * (size_t)c1 + 1] = y; prd[8 * (size_t)c1] = vx; prd[8 * (size_t)c1 + 1] = vy;
    }
    if (valid && c2 >= 0 && c2 < ncols) {
      pr[8 * (size_t)c2] = x; pr[8 * (size_t)c2 + 1] = y; prd[8 * (size_t)c2] = vx; prd[8 * (size_t)c2 + 1] = vy;
    }
    if (i == loop) break;
    const double gx = x - goal_x, gy = y - goal_y;
    const double d_goal = sqrt(gx * gx + gy * gy);
    double fx = -alpha * (gx / d_goal), fy = -alpha * (gy / d_goal);                             // F_att, :1000
    double rx = 0.0, ry = 0.0;
    for (int j = 0; j < n_obs; j++) {                                                            // F_rep, :1003-1017
      const double ox = x - ob[2 * j], oy = y - ob[2 * j + 1];
      const double d = sqrt(ox * ox + oy * oy);
      if (d < dmin) {
        const double w = eta * (1.0 / d - 1.0 / dmin) * (1.0 / (d * d));
        rx += w * (ox / d); ry += w * (oy / d);
      }
    }
    double ax = 0.0, ay = 0.0;
#pragma unroll
    for (int m = 1; m < 4; m++) {                                                                // Lennard-Jones, :1019-1029
      const double qx = __shfl_xor_sync(0xffffffffu, x, m, 4), qy = __shfl_xor_sync(0xffffffffu, y, m, 4);
      const double ox = x - qx, oy = y - qy;
      const double d2 = ox * ox + oy * oy, d = sqrt(d2), d4 = d2 * d2, d6 = d4 * d2;
      const double w = 4.0 * epsilon * (6.0 / (d6 * d) - 12.0 / (d6 * d6 * d));
      ax = ax - w * (ox / d); ay = ay - w * (oy / d);
    }
    fx += rx + ax; fy += ry + ay;
    if (d_goal < 0.001) { fx = 0.0; fy = 0.0; }                                                  // :1033-1036
    const double scale = (i < ramp) ? ((double)i / ramp) : 1.0;                                  // :1039
    const double ux = fx * scale, uy = fy * scale;
    x = x + a13 * vx + b1 * ux; y = y + a13 * vy + b1 * uy;                                      // q+ = Ae q + Be F scale, :1041
    vx = a33 * vx + b3 * ux; vy = a33 * vy + b3 * uy;
  }
}

// Dense batched QP (cbfnmpc_qpdense.h): one warp per problem, 4 problems per CTA (56.6 KB of shared memory each).
#define QD_WARPS 4
__global__ void __launch_bounds__(32 * QD_WARPS, 1) cbfnmpc_qpdense_kernel(const __grid_constant__ QDenseIO io) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  QDenseWS* ws = reinterpret_cast<QDenseWS*>(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int q = blockIdx.x * QD_WARPS + warp;
  qdense_solve_warp(ws[warp], io, q < io.B ? q : 0, q < io.B);
}

// FP64 pipe microbenchmark: 8 independent DFMA chains per thread.
__global__ void cbfnmpc_dfma_kernel(double* out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(CBFNMPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

#define CBF_MAX_CHUNKS 4
// Rows of the device output staging ([row][B] doubles).  What a caller of MPC_dist reads every cycle (X, U, slack, alpha, next
// alpha buffer, KKT error) comes first and is contiguous, so that a host block laid out the same way comes back in ONE copy per
// column range; lam and x_qp (diagnostics) follow and are only downloaded when asked for.
enum { O_X = 0, O_U = 16, O_SLACK = 24, O_ALPHA = 25, O_ABUF = 45, O_KKT = 49, O_LAM = 50, O_XQP = 66, O_ROWS = 90 };
#define CBF_REFGEN_MAX_OBS 440
struct cbfnmpc_ctx {
  int device, max_batch, max_obs;
  cbfnmpc_gains gains;
  cbfnmpc_options opts;
  KParams kp;
  cudaStream_t stream;
  cudaStream_t cstream[CBF_MAX_CHUNKS];  // host-buffer calls: chunk k of the batch travels on cstream[k] (cstream[0] == stream)
  cudaEvent_t ev0, ev1;
  cudaEvent_t ev_last;       // recorded after the last launch that used the context's class lists / staging buffers
  cudaStream_t last_stream;  // the stream it was recorded on (a call on another stream waits for it first)
  bool have_last;
  bool timed;
  int* d_obsflag; int* h_obsseq; int obs_seq;   // host-buffer calls: "the obstacles have landed" word (cbfnmpc_lane_kernel_late)
  bool timing_on;            // record the event pair around the solve kernel (cbfnmpc_set_kernel_timing; default on)
  int sm_count;
  long long launches;
  // staging for the host-pointer entry points
  double* d_in;   // (32 + 2 max_obs) * max_batch
  int* d_contact; // 4 * max_batch
  double* d_out;  // 90 * max_batch
  int* d_outi;    // 4 * max_batch
  unsigned long long* d_active;  // 2 * max_batch
  double* d_xstart;              // 41 (dense entry point)
  int* d_perm;                   // 2 * max_batch + 2: class lists + counts (grown on demand for device-pointer batches)
  int perm_cap;
  // extended mode (cbfnmpc_ext_api.cuh): class lists over teams, host staging, resumable state, published trajectories, NCCL
  int* x_perm; int x_perm_cap;
  double *x_in, *x_out; int* x_int; size_t x_in_cap, x_out_cap, x_int_cap;
  double* x_state; size_t x_state_cap;
  double* x_pub; size_t x_pub_cap;
  int *x_active, *x_active_h;
  cudaEvent_t x_ev[2], x_t0, x_t1;
  void* comm; int comm_size, comm_rank;
  // cbfnmpc_ext_solve_fused: own mailbox, the mailboxes of the communicator's ranks mapped into this process (CUDA IPC; f_peer[rank] ==
  // f_mbox), control words {error, largest team iteration count}, a few doubles for the handle exchange / the start barrier
  double* f_mbox; size_t f_mbox_cap; double* f_peer[8]; int* f_ctl; double* f_bar; bool f_disabled, f_mapped;
  // small host-buffer batches (<= small_cap instances): one pinned block up, one down (see solve_small)
  int small_cap;
  unsigned char *h_sin, *h_sout, *d_sin, *d_sout;
};

// Small host-buffer batches: inputs  = (32 + 2K) B doubles | contact 4B ints | class lists 2B ints | counts 2 ints (+ pad),
//                             outputs = 90 B doubles | active 2B u64 | status, qp_iters, nlp_iters, n_fact 4B ints.
#define CBF_SMALL_BATCH 64
static size_t small_in_bytes(int B, int K) { return sizeof(double) * (size_t)(32 + 2 * K) * B + sizeof(int) * ((size_t)6 * B + 4); }
static size_t small_out_bytes(int B) { return sizeof(double) * (size_t)92 * B + sizeof(int) * (size_t)4 * B; }

extern "C" {

int cbfnmpc_comm_destroy(cbfnmpc_ctx* c);
int cbfnmpc_version(void) { return 100; }
const char* cbfnmpc_last_error(void) { return g_err.c_str(); }
void cbfnmpc_default_gains(cbfnmpc_gains* g) { host_default_gains(g); }
void cbfnmpc_default_options(cbfnmpc_options* o) { host_default_options(o); }

int cbfnmpc_create(cbfnmpc_ctx** out, int device, int max_batch, int max_obs) {
  if (!out || max_batch < 1 || max_obs < 1) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_create: bad argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(CBFNMPC_ERR_NODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_create: device index out of range");
  CK(cudaSetDevice(device));
  cbfnmpc_ctx* c = new cbfnmpc_ctx();
  c->device = device; c->max_batch = max_batch; c->max_obs = max_obs; c->launches = 0; c->timed = false;
  { cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device)); c->sm_count = prop.multiProcessorCount; }
  c->timing_on = true;
  host_default_gains(&c->gains); host_default_options(&c->opts);
  host_fill_kparams(&c->gains, &c->opts, &c->kp);
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->cstream[0] = c->stream;
  for (int k = 1; k < CBF_MAX_CHUNKS; k++) CK(cudaStreamCreateWithFlags(&c->cstream[k], cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
  CK(cudaEventCreateWithFlags(&c->ev_last, cudaEventDisableTiming)); c->have_last = false; c->last_stream = c->stream;
  CK(cudaMalloc(&c->d_in, sizeof(double) * (size_t)(32 + 2 * max_obs) * max_batch));
  CK(cudaMalloc(&c->d_contact, sizeof(int) * 4 * (size_t)max_batch));
  CK(cudaMalloc(&c->d_out, sizeof(double) * 90 * (size_t)max_batch));
  CK(cudaMalloc(&c->d_outi, sizeof(int) * 4 * (size_t)max_batch));
  CK(cudaMalloc(&c->d_active, sizeof(unsigned long long) * 2 * (size_t)max_batch));
  CK(cudaMalloc(&c->d_xstart, sizeof(double) * 41));
  CK(cudaMalloc(&c->d_perm, sizeof(int) * (2 * (size_t)max_batch + 2 * CBF_MAX_CHUNKS)));
  c->perm_cap = max_batch;
  c->small_cap = max_batch < CBF_SMALL_BATCH ? max_batch : CBF_SMALL_BATCH;
  {
    const size_t nin = small_in_bytes(c->small_cap, max_obs), nout = small_out_bytes(c->small_cap);
    CK(cudaMalloc(&c->d_sin, nin)); CK(cudaMalloc(&c->d_sout, nout));
    CK(cudaHostAlloc(&c->h_sin, nin, cudaHostAllocDefault)); CK(cudaHostAlloc(&c->h_sout, nout, cudaHostAllocDefault));
  }
  // function attributes are per device: set them for every context, not once per process
  CK(cudaFuncSetAttribute(cbfnmpc_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LANE_NPRIV * LANE_NT * sizeof(double))));
  {
    // two CTAs per SM need 2 x (93.2 + 1) KB of shared memory: ask for exactly the carve-out step that holds them, so that the rest of
    // the 256 KB stays L1 -- the kernel's stack (spilled registers: 408 B x 128 threads per SM) then fits in it instead of going to L2
    const int need_kb = (int)((2 * (LANE_NPRIV * LANE_NT * sizeof(double) + 1024) + 1023) / 1024);
    const int pct = (100 * need_kb + 227) / 228;
    CK(cudaFuncSetAttribute(cbfnmpc_lane_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct < 100 ? pct : 100));
    CK(cudaFuncSetAttribute(cbfnmpc_lane_kernel_late, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LANE_NPRIV * LANE_NT * sizeof(double))));
    CK(cudaFuncSetAttribute(cbfnmpc_lane_kernel_late, cudaFuncAttributePreferredSharedMemoryCarveout, pct < 100 ? pct : 100));
  }
  CK(cudaMalloc(&c->d_obsflag, sizeof(int))); CK(cudaMemset(c->d_obsflag, 0, sizeof(int)));
  CK(cudaHostAlloc(&c->h_obsseq, sizeof(int), cudaHostAllocDefault)); c->obs_seq = 0;
  CK(cudaFuncSetAttribute(cbfnmpc_qpdense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(QD_WARPS * sizeof(QDenseWS))));
  CK(cudaFuncSetAttribute(cbfnmpc_refgen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * 64 * CBF_REFGEN_MAX_OBS)));
  *out = c;
  return CBFNMPC_OK;
}

void cbfnmpc_destroy(cbfnmpc_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (int k = 0; k < CBF_MAX_CHUNKS; k++) cudaStreamSynchronize(c->cstream[k]);
  if (c->have_last) cudaEventSynchronize(c->ev_last);  // the caller's stream of the last call
  cudaFree(c->d_in); cudaFree(c->d_contact); cudaFree(c->d_out); cudaFree(c->d_outi); cudaFree(c->d_active); cudaFree(c->d_xstart); cudaFree(c->d_perm);
  cudaFree(c->d_sin); cudaFree(c->d_sout); cudaFreeHost(c->h_sin); cudaFreeHost(c->h_sout);
  cbfnmpc_comm_destroy(c);
  cudaFree(c->x_perm); cudaFree(c->x_in); cudaFree(c->x_out); cudaFree(c->x_int); cudaFree(c->x_state); cudaFree(c->x_pub);
  cudaFree(c->d_obsflag); if (c->h_obsseq) cudaFreeHost(c->h_obsseq);
  for (int r = 0; r < 8; r++) if (c->f_peer[r] && c->f_peer[r] != c->f_mbox) cudaIpcCloseMemHandle(c->f_peer[r]);
  cudaFree(c->f_mbox); cudaFree(c->f_ctl); cudaFree(c->f_bar);
  if (c->x_active) { cudaFree(c->x_active); cudaFreeHost(c->x_active_h); cudaEventDestroy(c->x_ev[0]); cudaEventDestroy(c->x_ev[1]); cudaEventDestroy(c->x_t0); cudaEventDestroy(c->x_t1); }
  cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev_last);
  for (int k = 1; k < CBF_MAX_CHUNKS; k++) cudaStreamDestroy(c->cstream[k]);
  cudaStreamDestroy(c->stream);
  delete c;
}

int cbfnmpc_set_gains(cbfnmpc_ctx* c, const cbfnmpc_gains* g) {
  if (!c || !g) return fail(CBFNMPC_ERR_ARG, "null argument");
  if (!(g->h_com > 0) || !(g->Ts > 0) || !(g->Rw > 0) || !(g->slack_w > 0) || !(g->box > 0) || !(g->v_max > 0))
    return fail(CBFNMPC_ERR_ARG, "cbfnmpc_set_gains: non-positive gain");
  c->gains = *g;
  host_fill_kparams(&c->gains, &c->opts, &c->kp);
  return CBFNMPC_OK;
}

int cbfnmpc_set_options(cbfnmpc_ctx* c, const cbfnmpc_options* o) {
  if (!c || !o) return fail(CBFNMPC_ERR_ARG, "null argument");
  if (o->qp_maxit < 1 || o->max_iter < 1 || !(o->tol > 0) || !(o->push > 0) || o->filter_cap < 1 || o->max_refine < 0)
    return fail(CBFNMPC_ERR_ARG, "cbfnmpc_set_options: bad option");
  c->opts = *o;
  host_fill_kparams(&c->gains, &c->opts, &c->kp);
  return CBFNMPC_OK;
}

// One context owns one set of class lists and staging buffers, so its calls must not overlap each other.  Calls on the same
// stream are ordered by the stream; a call on a different stream than the previous one first waits for the previous call's work
// (an event, no host synchronisation).
static int ctx_enter(cbfnmpc_ctx* c, cudaStream_t st) {
  if (c->have_last && c->last_stream != st) CK(cudaStreamWaitEvent(st, c->ev_last, 0));
  return CBFNMPC_OK;
}
static int ctx_leave(cbfnmpc_ctx* c, cudaStream_t st) {
  CK(cudaEventRecord(c->ev_last, st));
  c->last_stream = st; c->have_last = true;
  return CBFNMPC_OK;
}

// B instances whose arrays have pitch ld; perm = this launch's class lists (2 B + 2 ints)
static int launch_on(cbfnmpc_ctx* c, int B, int ld, int mode, const cbfnmpc_in* in, const cbfnmpc_out* out, const double* xstart,
                     cudaStream_t st, int* perm, bool timed = true, bool classify = true, const int* obs_flag = nullptr, int obs_seq = 0) {
  BatchIO io;
  io.obs_flag = obs_flag; io.obs_seq = obs_seq;
  io.x0 = in->x0; io.xdes = in->xdes; io.feet = in->feet_xy; io.contact = in->contact; io.obs = in->obs_xy; io.n_obs = in->n_obs;
  io.abuf = in->alpha_buf; io.xstart = xstart;
  io.X = out->X; io.U = out->U; io.lam = out->lam; io.slack = out->slack; io.xqp = out->x_qp; io.alpha = out->alpha;
  io.abuf_next = out->alpha_buf_next; io.kkt = out->kkt; io.status = out->status; io.qp_iters = out->qp_iters;
  io.nlp_iters = out->nlp_iters; io.n_fact = out->n_fact; io.active = (unsigned long long*)out->active;
  io.B = B; io.ld = ld; io.mode = mode;
  int* counts = perm + 2 * (size_t)B;
  const int ipc = LANE_NT / 4;
  if (classify) {
    cbfnmpc_classify_kernel<<<1, 1024, 0, st>>>(io, perm, counts);
    CK(cudaGetLastError());
    c->launches++;
  }
  timed = timed && c->timing_on;
  if (timed) CK(cudaEventRecord(c->ev0, st));  // the event pair brackets the solve kernel alone
  {
    // Without the event in between, the solve kernel is a programmatic dependent of the classify kernel: its CTAs are placed while
    // the class lists are still being written and wait for them at griddepcontrol.wait (the launch latency leaves the step).
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)((B + ipc - 1) / ipc + 1)); cfg.blockDim = dim3(LANE_NT);
    cfg.dynamicSmemBytes = LANE_NPRIV * LANE_NT * sizeof(double); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (classify && !timed) ? 1 : 0;
    const int* cperm = perm; const int* ccounts = counts;
    if (obs_flag) CK(cudaLaunchKernelEx(&cfg, cbfnmpc_lane_kernel_late, c->kp, io, cperm, ccounts, c->sm_count));
    else CK(cudaLaunchKernelEx(&cfg, cbfnmpc_lane_kernel, c->kp, io, cperm, ccounts, c->sm_count));
  }
  CK(cudaGetLastError());
  if (timed) { CK(cudaEventRecord(c->ev1, st)); c->timed = true; }
  c->launches++;
  return CBFNMPC_OK;
}

// A batch of at most small_cap instances in host memory: the latency of such a call is the driver's per-copy cost, not
// bandwidth, so everything goes up in ONE copy from a pinned block (the class lists included: the stance classes are sorted on
// the host, which has the contact flags, instead of by cbfnmpc_classify_kernel) and comes back in ONE copy.
static int solve_small(cbfnmpc_ctx* c, int B, int mode, const cbfnmpc_in* in, const cbfnmpc_out* out, int K) {
  const size_t Bs = (size_t)B;
  double* hd = (double*)c->h_sin;
  const size_t nd = (size_t)(32 + 2 * K) * Bs;
  memcpy(hd, in->x0, sizeof(double) * 4 * Bs);
  memcpy(hd + 4 * Bs, in->xdes, sizeof(double) * 16 * Bs);
  memcpy(hd + 20 * Bs, in->feet_xy, sizeof(double) * 8 * Bs);
  if (in->alpha_buf) memcpy(hd + 28 * Bs, in->alpha_buf, sizeof(double) * 4 * Bs); else memset(hd + 28 * Bs, 0, sizeof(double) * 4 * Bs);
  if (K > 0) memcpy(hd + 32 * Bs, in->obs_xy, sizeof(double) * 2 * (size_t)K * Bs);
  int* hi = (int*)(hd + nd);
  memcpy(hi, in->contact, sizeof(int) * 4 * Bs);
  int* hperm = hi + 4 * Bs;
  int n4 = 0, n2 = 0, nbad = 0;
  int bad[CBF_SMALL_BATCH];
  for (int b = 0; b < B; b++) {  // lane_classify, stable within a class like the kernel's scan
    const int c0 = hi[b], csum = c0 + hi[Bs + b] + hi[2 * Bs + b] + hi[3 * Bs + b];
    if (csum == 4) hperm[n4++] = b;
    else if (csum == 2 && (c0 == 0 || c0 == 1)) hperm[Bs + n2++] = b;
    else bad[nbad++] = b;
  }
  hperm[2 * Bs] = n4; hperm[2 * Bs + 1] = n2;
  cudaStream_t st = c->stream;
  CK(cudaMemcpyAsync(c->d_sin, c->h_sin, small_in_bytes(B, K), cudaMemcpyHostToDevice, st));
  double* d = (double*)c->d_sin;
  int* di = (int*)(d + nd);
  cbfnmpc_in din = {d, d + 4 * Bs, d + 20 * Bs, di, d + 32 * Bs, K, d + 28 * Bs};
  double* o = (double*)c->d_sout;
  cbfnmpc_out dout;
  dout.X = o + O_X * Bs; dout.U = o + O_U * Bs; dout.lam = o + O_LAM * Bs; dout.slack = o + O_SLACK * Bs; dout.x_qp = o + O_XQP * Bs;
  dout.alpha = o + O_ALPHA * Bs; dout.alpha_buf_next = o + O_ABUF * Bs; dout.kkt = o + O_KKT * Bs;
  dout.active = (uint64_t*)(o + 90 * Bs);
  int32_t* doi = (int32_t*)(o + 92 * Bs);
  dout.status = doi; dout.qp_iters = doi + Bs; dout.nlp_iters = doi + 2 * Bs; dout.n_fact = doi + 3 * Bs;
  int rc = launch_on(c, B, B, mode, &din, &dout, nullptr, st, di + 4 * Bs, true, false);
  if (rc) return rc;
  CK(cudaMemcpyAsync(c->h_sout, c->d_sout, small_out_bytes(B), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  double* ho = (double*)c->h_sout;
  uint64_t* ha = (uint64_t*)(ho + 90 * Bs);
  int32_t* hoi = (int32_t*)(ho + 92 * Bs);
  for (int k = 0; k < nbad; k++) {  // lane_store_bad
    const int b = bad[k];
    hoi[b] = 4; hoi[Bs + b] = 0; hoi[2 * Bs + b] = 0; hoi[3 * Bs + b] = 0; ho[O_KKT * Bs + b] = 0.0; ho[O_SLACK * Bs + b] = 0.0;
    ha[b] = 0; ha[Bs + b] = 0;
  }
  struct OSeg { double* h; size_t row, rows; };
  const OSeg os[8] = {{out->X, O_X, 16}, {out->U, O_U, 8}, {out->slack, O_SLACK, 1}, {out->alpha, O_ALPHA, 20}, {out->alpha_buf_next, O_ABUF, 4},
                      {out->kkt, O_KKT, 1}, {out->lam, O_LAM, 16}, {out->x_qp, O_XQP, 24}};
  for (int i = 0; i < 8; i++) if (os[i].h) memcpy(os[i].h, ho + os[i].row * Bs, sizeof(double) * os[i].rows * Bs);
  int32_t* const oi[4] = {out->status, out->qp_iters, out->nlp_iters, out->n_fact};
  for (int i = 0; i < 4; i++) if (oi[i]) memcpy(oi[i], hoi + (size_t)i * Bs, sizeof(int32_t) * Bs);
  if (out->active) memcpy(out->active, ha, sizeof(uint64_t) * 2 * Bs);
  return CBFNMPC_OK;
}

static int launch(cbfnmpc_ctx* c, int B, int mode, const cbfnmpc_in* in, const cbfnmpc_out* out, const double* xstart,
                  cudaStream_t st) {
  if (B > c->perm_cap) {  // device-pointer batches larger than max_batch: grow the class lists
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(c->d_perm));
    CK(cudaMalloc(&c->d_perm, sizeof(int) * (2 * (size_t)B + 2 * CBF_MAX_CHUNKS)));
    c->perm_cap = B;
  }
  int rc = ctx_enter(c, st);
  if (rc) return rc;
  rc = launch_on(c, B, B, mode, in, out, xstart, st, c->d_perm);
  if (rc) return rc;
  return ctx_leave(c, st);
}

int cbfnmpc_solve_batch(cbfnmpc_ctx* c, int B, int mode, const cbfnmpc_in* in, const cbfnmpc_out* out, void* cuda_stream) {
  if (!c || !in || !out || B < 1) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_solve_batch: bad argument");
  if (mode != CBFNMPC_MODE_QP && mode != CBFNMPC_MODE_QP_NLP) return fail(CBFNMPC_ERR_ARG, "bad mode");
  if (!in->x0 || !in->xdes || !in->feet_xy || !in->contact) return fail(CBFNMPC_ERR_ARG, "missing input array");
  if (mode == CBFNMPC_MODE_QP_NLP && (!in->obs_xy || in->n_obs < 1)) return fail(CBFNMPC_ERR_ARG, "NLP mode needs obstacles");
  CK(cudaSetDevice(c->device));
  return launch(c, B, mode, in, out, nullptr, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

int cbfnmpc_solve_batch_host(cbfnmpc_ctx* c, int B, int mode, const cbfnmpc_in* in, const cbfnmpc_out* out) {
  if (!c || !in || !out || B < 1) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_solve_batch_host: bad argument");
  if (B > c->max_batch) return fail(CBFNMPC_ERR_ARG, "batch larger than max_batch");
  if (mode != CBFNMPC_MODE_QP && mode != CBFNMPC_MODE_QP_NLP) return fail(CBFNMPC_ERR_ARG, "bad mode");
  if (!in->x0 || !in->xdes || !in->feet_xy || !in->contact) return fail(CBFNMPC_ERR_ARG, "missing input array");
  const int K = (mode == CBFNMPC_MODE_QP_NLP) ? in->n_obs : 0;
  if (mode == CBFNMPC_MODE_QP_NLP && (!in->obs_xy || K < 1 || K > c->max_obs)) return fail(CBFNMPC_ERR_ARG, "NLP mode needs 1..max_obs obstacles");
  CK(cudaSetDevice(c->device));
  if (c->have_last) { CK(cudaEventSynchronize(c->ev_last)); c->have_last = false; }  // an earlier asynchronous call on a caller's stream
  if (B <= c->small_cap) return solve_small(c, B, mode, in, out, K);
  const size_t Bs = (size_t)B;
  // Column ranges, each on its own stream, let the upload of one range overlap the solves of the earlier ones and the downloads of
  // the earlier ones overlap the solves of the later ones; every range costs a classify + solve launch pair of its own, and side by
  // side those pairs take longer than one (four quarter batches of 1024 on four streams: 0.573 ms against 0.529 ms for one launch of
  // 4096, device-resident).  With the round-1 kernel (0.75 ms per 4096) four ranges were the optimum (1 / 2 / 4 / 8 ranges: 0.932 /
  // 0.890 / 0.880 / 0.903 ms); with the round-2 kernel the copies of a 4096 batch (0.06 + 0.045 ms) are worth less than the extra
  // launches -- measured host call (scripts/dev/e2e_parts.py), 1 / 2 / 4 ranges: 2048: 0.543 / 0.554 / 0.580; 4096: 0.647 / 0.654 /
  // 0.666; 8192: 0.995 / 1.062 / 1.022; 16384: 1.899 / 1.863 / 2.048; 32768: 3.487 / 3.317 / 3.425 ms -- so one range below 16384
  // instances and two from there on.  All uploads and launches are enqueued before any download.
  // Device staging (pitch B): x0 4 | xdes 16 | feet 8 | alpha_buf 4 | obs 2K rows of doubles; outputs in the order of
  // the O_* rows (X 16 | U 8 | slack 1 | alpha 20 | abuf_next 4 | kkt 1 | lam 16 | x_qp 24).
  int nch = B >= 16384 ? 2 : 1;
  if (const char* e = getenv("CBFNMPC_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= CBF_MAX_CHUNKS) nch = v; }   // tuning experiments only
  size_t cut[CBF_MAX_CHUNKS + 1];
  for (int k = 0; k <= nch; k++) cut[k] = Bs * k / nch;
  double* d = c->d_in;
  double* o = c->d_out;
  struct Seg { const double* h; size_t row, rows; };
  const Seg segs[5] = {{in->x0, 0, 4}, {in->xdes, 4, 16}, {in->feet_xy, 20, 8}, {in->alpha_buf, 28, 4}, {in->obs_xy, 32, 2 * (size_t)K}};
  struct OSeg { double* h; size_t row, rows; };
  const OSeg os[8] = {{out->X, O_X, 16}, {out->U, O_U, 8}, {out->slack, O_SLACK, 1}, {out->alpha, O_ALPHA, 20}, {out->alpha_buf_next, O_ABUF, 4},
                      {out->kkt, O_KKT, 1}, {out->lam, O_LAM, 16}, {out->x_qp, O_XQP, 24}};
  int32_t* const oi[4] = {out->status, out->qp_iters, out->nlp_iters, out->n_fact};
  const size_t pitch = sizeof(double) * Bs;
  // One range: the obstacles (more than half of the upload) travel on a second stream while the classify pre-pass and the QP stage
  // run; a word copied behind them is what cbfnmpc_lane_kernel_late waits for before it looks at them.
  const bool late = nch == 1 && K > 0 && getenv("CBFNMPC_NO_LATE_OBS") == nullptr;
  const int nseg = late ? 4 : 5;
  for (int k = 0; k < nch; k++) {
    const size_t b0 = cut[k], n = cut[k + 1] - b0;
    cudaStream_t st = c->cstream[k];
    // adjacent host arrays (one pinned block, as bench.py allocates) are merged into a single 2-D copy
    for (int i = 0; i < nseg;) {
      if (!segs[i].h || segs[i].rows == 0) { i++; continue; }
      int j = i; size_t rows = segs[i].rows;
      while (j + 1 < nseg && segs[j + 1].h == segs[j].h + segs[j].rows * Bs && segs[j + 1].rows > 0 && segs[j + 1].row == segs[j].row + segs[j].rows) { j++; rows += segs[j].rows; }
      CK(cudaMemcpy2DAsync(d + segs[i].row * Bs + b0, pitch, segs[i].h + b0, pitch, sizeof(double) * n, rows, cudaMemcpyHostToDevice, st));
      i = j + 1;
    }
    if (!in->alpha_buf) CK(cudaMemset2DAsync(d + 28 * Bs + b0, pitch, 0, sizeof(double) * n, 4, st));
    CK(cudaMemcpy2DAsync(c->d_contact + b0, sizeof(int) * Bs, in->contact + b0, sizeof(int) * Bs, sizeof(int) * n, 4, cudaMemcpyHostToDevice, st));
    cbfnmpc_in din = {d + b0, d + 4 * Bs + b0, d + 20 * Bs + b0, c->d_contact + b0, d + 32 * Bs + b0, K, d + 28 * Bs + b0};
    cbfnmpc_out dout;
    dout.X = o + O_X * Bs + b0; dout.U = o + O_U * Bs + b0; dout.lam = o + O_LAM * Bs + b0; dout.slack = o + O_SLACK * Bs + b0;
    dout.x_qp = o + O_XQP * Bs + b0; dout.alpha = o + O_ALPHA * Bs + b0; dout.alpha_buf_next = o + O_ABUF * Bs + b0; dout.kkt = o + O_KKT * Bs + b0;
    dout.status = c->d_outi + b0; dout.qp_iters = c->d_outi + Bs + b0; dout.nlp_iters = c->d_outi + 2 * Bs + b0;
    dout.n_fact = c->d_outi + 3 * Bs + b0;
    dout.active = (uint64_t*)c->d_active + b0;
    if (late) {
      *c->h_obsseq = ++c->obs_seq;
      CK(cudaMemcpyAsync(d + 32 * Bs, in->obs_xy, sizeof(double) * 2 * (size_t)K * Bs, cudaMemcpyHostToDevice, c->cstream[1]));
      CK(cudaMemcpyAsync(c->d_obsflag, c->h_obsseq, sizeof(int), cudaMemcpyHostToDevice, c->cstream[1]));
    }
    int rc = launch_on(c, (int)n, B, mode, &din, &dout, nullptr, st, c->d_perm + 2 * b0 + 2 * k, k == 0, true, late ? c->d_obsflag : nullptr, c->obs_seq);
    if (rc) return rc;
  }
  for (int k = 0; k < nch; k++) {
    const size_t b0 = cut[k], n = cut[k + 1] - b0;
    cudaStream_t st = c->cstream[k];
    for (int i = 0; i < 8;) {
      if (!os[i].h) { i++; continue; }
      int j = i; size_t rows = os[i].rows;
      while (j + 1 < 8 && os[j + 1].h == os[j].h + os[j].rows * Bs && os[j + 1].row == os[j].row + os[j].rows) { j++; rows += os[j].rows; }
      CK(cudaMemcpy2DAsync(os[i].h + b0, pitch, o + os[i].row * Bs + b0, pitch, sizeof(double) * n, rows, cudaMemcpyDeviceToHost, st));
      i = j + 1;
    }
    for (int i = 0; i < 4;) {
      if (!oi[i]) { i++; continue; }
      int j = i; size_t rows = 1;
      while (j + 1 < 4 && oi[j + 1] == oi[j] + Bs) { j++; rows++; }
      CK(cudaMemcpy2DAsync(oi[i] + b0, sizeof(int) * Bs, c->d_outi + (size_t)i * Bs + b0, sizeof(int) * Bs, sizeof(int) * n, rows, cudaMemcpyDeviceToHost, st));
      i = j + 1;
    }
    if (out->active)
      CK(cudaMemcpy2DAsync(out->active + b0, sizeof(uint64_t) * Bs, c->d_active + b0, sizeof(uint64_t) * Bs, sizeof(uint64_t) * n, 2, cudaMemcpyDeviceToHost, st));
  }
  for (int k = 0; k < nch; k++) CK(cudaStreamSynchronize(c->cstream[k]));
  if (late) CK(cudaStreamSynchronize(c->cstream[1]));
  return CBFNMPC_OK;
}

#ifdef CBF_CTATIME
int cbfnmpc_debug_ctatime(long long* out, int n) {  // debug builds only
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(out, g_ctat, sizeof(long long) * (size_t)n);
}
#endif
#ifdef CBF_DBGBUF
int cbfnmpc_debug_read(double* out, int reset) {  // debug builds only
  int n = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&n, g_dbg_n, sizeof(int));
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(double) * 8192);
  if (reset) { int z = 0; cudaMemcpyToSymbol(g_dbg_n, &z, sizeof(int)); }
  return n;
}
#endif

static int check_raw(const cbfnmpc_raw_in* r) {
  if (!r->q || !r->dq || !r->contact || !r->toe_pos || !r->gait_domain || !r->agent_id || !r->team || !r->agent_initial ||
      !r->Pr_refined || !r->Prd_refined || !r->Pobs_real)
    return fail(CBFNMPC_ERR_ARG, "cbfnmpc_raw_in: missing array");
  if (r->nq < 2 || r->ref_cols < 1 || r->n_teams < 1 || r->n_obs < 1) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_raw_in: bad size");
  if (r->n_other < 0 || (r->n_other > 0 && !r->state_other)) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_raw_in: n_other without state_other");
  return CBFNMPC_OK;
}

int cbfnmpc_pack_inputs(cbfnmpc_ctx* c, int B, const cbfnmpc_raw_in* raw, const cbfnmpc_in* packed, void* cuda_stream) {
  if (!c || !raw || !packed || B < 1) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_pack_inputs: bad argument");
  if (!packed->x0 || !packed->xdes || !packed->feet_xy || !packed->contact || !packed->obs_xy) return fail(CBFNMPC_ERR_ARG, "missing output array");
  int rc = check_raw(raw);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  PackIO io;
  io.r = *raw; io.B = B; io.Ts = c->gains.Ts;
  io.x0 = (double*)packed->x0; io.xdes = (double*)packed->xdes; io.feet = (double*)packed->feet_xy; io.obs = (double*)packed->obs_xy;
  io.contact = (int*)packed->contact;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  rc = ctx_enter(c, st);
  if (rc) return rc;
  cbfnmpc_pack_kernel<<<(B + 255) / 256, 256, 0, st>>>(io);
  CK(cudaGetLastError());
  c->launches++;
  return ctx_leave(c, st);
}

int cbfnmpc_team_states(cbfnmpc_ctx* c, int B, int team_size, const double* q, const double* dq, double* state_other, void* cuda_stream) {
  if (!c || !q || !dq || !state_other || B < 1 || team_size < 2 || B % team_size != 0)
    return fail(CBFNMPC_ERR_ARG, "cbfnmpc_team_states: bad argument (B must be a multiple of team_size >= 2)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  cbfnmpc_team_states_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, team_size, q, dq, state_other);
  CK(cudaGetLastError());
  c->launches++;
  return CBFNMPC_OK;
}

int cbfnmpc_run_cycle(cbfnmpc_ctx* c, int B, int mode, const cbfnmpc_raw_in* raw, const double* alpha_buf, const cbfnmpc_out* out,
                      void* cuda_stream) {
  if (!c || !raw || !out || B < 1) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_run_cycle: bad argument");
  if (B > c->max_batch) return fail(CBFNMPC_ERR_ARG, "batch larger than max_batch");
  const int K = raw->n_obs + (raw->state_other ? raw->n_other : 0);
  if (K > c->max_obs) return fail(CBFNMPC_ERR_ARG, "more obstacle columns (n_obs + n_other) than max_obs");
  if (mode != CBFNMPC_MODE_QP && mode != CBFNMPC_MODE_QP_NLP) return fail(CBFNMPC_ERR_ARG, "bad mode");
  const size_t Bs = (size_t)B;
  double* d = c->d_in;  // staging layout of cbfnmpc_solve_batch_host: x0 4 | xdes 16 | feet 8 | alpha_buf 4 | obs 2K
  cbfnmpc_in din = {d, d + 4 * Bs, d + 20 * Bs, c->d_contact, d + 32 * Bs, K, alpha_buf};
  int rc = cbfnmpc_pack_inputs(c, B, raw, &din, cuda_stream);
  if (rc) return rc;
  return launch(c, B, mode, &din, out, nullptr, cuda_stream ? (cudaStream_t)cuda_stream : c->stream);
}

int cbfnmpc_qp_dense_batch(cbfnmpc_ctx* c, int B, int n, int p, int m, const double* P, const double* cv, const double* A, const double* b,
                           const double* G, const double* h, double* x, double* y, double* z, int32_t* status, int32_t* iters, void* cuda_stream) {
  if (!c || B < 1 || !P || !cv || !G || !h || !x || !status || !iters) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_qp_dense_batch: bad argument");
  if (n < 1 || n > QD_NMAX || p < 0 || p > QD_PMAX || m < 1 || m > QD_MMAX) return fail(CBFNMPC_ERR_ARG, "dimensions outside n <= 32, p <= 24, 1 <= m <= 64");
  if (p > 0 && (!A || !b)) return fail(CBFNMPC_ERR_ARG, "equality data missing");
  CK(cudaSetDevice(c->device));
  QDenseIO io;
  io.B = B; io.n = n; io.p = p; io.m = m; io.P = P; io.c = cv; io.A = A; io.b = b; io.G = G; io.h = h;
  io.reltol = c->opts.qp_reltol; io.abstol = c->opts.qp_abstol; io.maxit = c->opts.qp_maxit;
  io.x = x; io.y = y; io.z = z; io.status = status; io.iters = iters;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  cbfnmpc_qpdense_kernel<<<(B + QD_WARPS - 1) / QD_WARPS, 32 * QD_WARPS, QD_WARPS * sizeof(QDenseWS), st>>>(io);
  CK(cudaGetLastError());
  c->launches++;
  return CBFNMPC_OK;
}

int cbfnmpc_generate_reference(cbfnmpc_ctx* c, int n_teams, const double* Pstart, const double* Pobs, int n_obs, double goal_x, double goal_y,
                               int loop, double* Pr, double* Prd, void* cuda_stream) {
  // the obstacles of a CTA's 32 teams sit in shared memory: 512 B per obstacle, 227 KB per CTA at most
  if (!c || !Pstart || !Pobs || !Pr || !Prd || n_teams < 1 || n_obs < 1 || n_obs > CBF_REFGEN_MAX_OBS || loop < 120)
    return fail(CBFNMPC_ERR_ARG, "cbfnmpc_generate_reference: bad argument (1 <= n_obs <= 440, loop >= 120)");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  cbfnmpc_refgen_kernel<<<(n_teams + 31) / 32, 128, sizeof(double) * 64 * (size_t)n_obs, st>>>(n_teams, Pstart, Pobs, n_obs, goal_x, goal_y, loop, Pr, Prd);
  CK(cudaGetLastError());
  c->launches++;
  return CBFNMPC_OK;
}

int cbfnmpc_closed_loop(cbfnmpc_ctx* c, int B, int n_cycles, int mode, const cbfnmpc_raw_in* raw, const cbfnmpc_world_state* w,
                        double fail_radius, void* cuda_stream) {
  if (!c || !raw || !w || B < 1 || n_cycles < 1) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_closed_loop: bad argument");
  if (!w->q || !w->dq || !w->contact || !w->toe_pos || !w->gait_domain || !w->alpha_buf || !w->success || !w->distance_to_fail ||
      !w->n_not_converged)
    return fail(CBFNMPC_ERR_ARG, "cbfnmpc_world_state: missing array");
  if (w->q != raw->q || w->dq != raw->dq || w->contact != raw->contact || w->toe_pos != raw->toe_pos || w->gait_domain != raw->gait_domain)
    return fail(CBFNMPC_ERR_ARG, "cbfnmpc_world_state must alias the arrays of cbfnmpc_raw_in");
  if (B > c->max_batch) return fail(CBFNMPC_ERR_ARG, "batch larger than max_batch");
  if (!(fail_radius > 0.0)) return fail(CBFNMPC_ERR_ARG, "fail_radius must be positive");
  const size_t Bs = (size_t)B;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  double* o = c->d_out;  // staging layout of cbfnmpc_solve_batch_host
  cbfnmpc_out dout;
  memset(&dout, 0, sizeof(dout));
  dout.X = o + O_X * Bs; dout.U = o + O_U * Bs; dout.alpha_buf_next = o + O_ABUF * Bs; dout.kkt = o + O_KKT * Bs; dout.status = c->d_outi;
  for (int cyc = 0; cyc < n_cycles; cyc++) {
    int rc = cbfnmpc_run_cycle(c, B, mode, raw, w->alpha_buf, &dout, st);
    if (rc) return rc;
    RollIO io;
    io.w = *w; io.X = dout.X; io.abuf_next = dout.alpha_buf_next; io.agent_initial = raw->agent_initial; io.Pobs_real = raw->Pobs_real;
    io.status = dout.status; io.team = raw->team; io.B = B; io.n_obs = raw->n_obs; io.cycle = cyc; io.fail_radius = fail_radius; io.n_teams = raw->n_teams;
    cbfnmpc_rollout_kernel<<<(B + 255) / 256, 256, 0, st>>>(io);
    CK(cudaGetLastError());
    c->launches++;
  }
  return ctx_leave(c, st);
}

long long cbfnmpc_launch_count(const cbfnmpc_ctx* c) { return c ? c->launches : 0; }

int cbfnmpc_set_kernel_timing(cbfnmpc_ctx* c, int on) {
  if (!c) return fail(CBFNMPC_ERR_ARG, "null argument");
  c->timing_on = on != 0;
  return CBFNMPC_OK;
}

int cbfnmpc_last_kernel_ms(cbfnmpc_ctx* c, float* ms) {
  if (!c || !ms) return fail(CBFNMPC_ERR_ARG, "null argument");
  if (!c->timed) return fail(CBFNMPC_ERR_ARG, "no kernel launched yet");
  CK(cudaSetDevice(c->device));
  CK(cudaEventSynchronize(c->ev1));
  CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return CBFNMPC_OK;
}

int cbfnmpc_measure_fp64_peak(cbfnmpc_ctx* c, double* tflops) {
  if (!c || !tflops) return fail(CBFNMPC_ERR_ARG, "null argument");
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  double* d = nullptr;
  CK(cudaMalloc(&d, sizeof(double) * (size_t)blocks * threads));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaEventRecord(a, c->stream));
    cbfnmpc_dfma_kernel<<<blocks, threads, 0, c->stream>>>(d, iters);
    CK(cudaEventRecord(b, c->stream));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (rep > 0 && ms < best) best = ms;
  }
  c->launches += 4;
  CK(cudaEventDestroy(a)); CK(cudaEventDestroy(b)); CK(cudaFree(d));
  *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) * 1e-12;
  return CBFNMPC_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Dense nlpTuple entry point (include/pack_nlp_data.h:11-29): recover the compact form, check the structure by
// re-packing exactly as src/MPC_dist.cpp:164-321, then solve one instance.
// ---------------------------------------------------------------------------------------------------------
static void host_pack_dense(const cbfnmpc_gains* g, const double* x0, const double* xdes, const double* F /*2x4, stride 4*/, int c,
                            std::vector<double>& Q, std::vector<double>& f, std::vector<double>& A, std::vector<double>& b,
                            std::vector<double>& G, std::vector<double>& h) {
  const int n = 4, m = 2, N = 4, nv = 25 + 4 * c, p = 28, mi = 60 + 8 * c;
  double Ad[16], Bd[8];
  host_lip(g, Ad, Bd);
  Q.assign((size_t)nv * nv, 0.0); f.assign(nv, 0.0); A.assign((size_t)p * nv, 0.0); b.assign(p, 0.0); G.assign((size_t)mi * nv, 0.0); h.assign(mi, 1.0);
  for (int i = 0; i < 16; i++) { double w = i >= 12 ? g->Pw : g->Qw; Q[i + (size_t)i * nv] = w; f[i] = -xdes[i] * w; }
  for (int i = 0; i < 8; i++) Q[(16 + i) + (size_t)(16 + i) * nv] = g->Rw;
  Q[(nv - 1) + (size_t)(nv - 1) * nv] = g->slack_w;
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < n; i++) {
      A[(4 * k + i) + (size_t)(4 * k + i) * p] = -1.0;
      if (k >= 1) for (int j = 0; j < n; j++) A[(4 * k + i) + (size_t)(4 * (k - 1) + j) * p] += Ad[4 * i + j];
      for (int j = 0; j < m; j++) A[(4 * k + i) + (size_t)(16 + 2 * k + j) * p] = Bd[2 * i + j];
    }
    for (int a = 0; a < 2; a++) {
      A[(16 + 2 * k + a) + (size_t)(16 + 2 * k + a) * p] = 1.0;
      for (int j = 0; j < c; j++) A[(16 + 2 * k + a) + (size_t)(24 + c * k + j) * p] = -F[4 * a + j];
    }
    for (int j = 0; j < c; j++) A[(24 + k) + (size_t)(24 + c * k + j) * p] = 1.0;
    b[24 + k] = 1.0;
  }
  for (int i = 0; i < 4; i++) { double s = 0; for (int j = 0; j < 4; j++) s += Ad[4 * i + j] * x0[j]; b[i] = -s; }
  const int r = 6;
  for (int i = 0; i < 3; i++)
    for (int a = 0; a < 2; a++) {
      G[(2 * i + a) + (size_t)(4 * i + 2 * a) * mi] = 1.0; G[(2 * i + a) + (size_t)(16 + 2 * (i + 1) + a) * mi] = -1.0;
      G[(r + 2 * i + a) + (size_t)(4 * i + 2 * a) * mi] = -1.0; G[(r + 2 * i + a) + (size_t)(16 + 2 * (i + 1) + a) * mi] = 1.0;
    }
  int o = 12;
  for (int i = 0; i < 16; i++) { G[(o + i) + (size_t)i * mi] = 1.0; G[(o + 16 + i) + (size_t)i * mi] = -1.0; }
  o += 32;
  for (int i = 0; i < 8; i++) { G[(o + i) + (size_t)(16 + i) * mi] = 1.0; G[(o + 8 + i) + (size_t)(16 + i) * mi] = -1.0; }
  o += 16;
  for (int i = 0; i < 4 * c; i++) { G[(o + i) + (size_t)(24 + i) * mi] = -1.0; G[(o + 4 * c + i) + (size_t)(24 + i) * mi] = 1.0; }
  const double fb = g->mu_fric * g->h_com / sqrt(2.0);
  for (int i = 0; i < 12; i++) h[i] = fb;
  for (int i = 12; i < 60; i++) h[i] = g->box;
  for (int i = 0; i < 4 * c; i++) h[60 + i] = 0.0;
}

static bool close_vec(const double* a, const std::vector<double>& b, double tol) {
  for (size_t i = 0; i < b.size(); i++) if (!(fabs(a[i] - b[i]) <= tol * (1.0 + fabs(b[i])))) return false;
  return true;
}

int cbfnmpc_pack_dense(const cbfnmpc_ctx* c, const double x0[4], const double xdes[16], const double* F, int cc, double* Q, double* f,
                       double* Aeq, double* beq, double* G, double* h) {
  if (!c || !x0 || !xdes || !F || !Q || !f || !Aeq || !beq || !G || !h || (cc != 2 && cc != 4)) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_pack_dense: bad argument");
  std::vector<double> Q2, f2, A2, b2, G2, h2;
  host_pack_dense(&c->gains, x0, xdes, F, cc, Q2, f2, A2, b2, G2, h2);
  memcpy(Q, Q2.data(), sizeof(double) * Q2.size()); memcpy(f, f2.data(), sizeof(double) * f2.size());
  memcpy(Aeq, A2.data(), sizeof(double) * A2.size()); memcpy(beq, b2.data(), sizeof(double) * b2.size());
  memcpy(G, G2.data(), sizeof(double) * G2.size()); memcpy(h, h2.data(), sizeof(double) * h2.size());
  return CBFNMPC_OK;
}

int cbfnmpc_solve_dense_nlptuple(cbfnmpc_ctx* c, int nv, int p, int mi, const double* Q, const double* f, const double* Aeq,
                                 const double* beq, const double* G, const double* h, const double* x0_start, const double obs[2],
                                 int mode_nlp, double* x_out, int* status, double* kkt) {
  if (!c || !Q || !f || !Aeq || !beq || !G || !h || !x_out) return fail(CBFNMPC_ERR_ARG, "null argument");
  int cc = (nv - 25) / 4;
  if ((cc != 2 && cc != 4) || nv != 25 + 4 * cc || p != 28 || mi != 60 + 8 * cc)
    return fail(CBFNMPC_ERR_STRUCTURE, "dimensions are not those of MPC_dist (nv = 25+4c, p = 28, mi = 60+8c, c in {2,4})");
  if (mode_nlp && (!x0_start || !obs)) return fail(CBFNMPC_ERR_ARG, "NLP mode needs x0 and the obstacle");
  const cbfnmpc_gains* g = &c->gains;
  // compact form: F from the u = F lambda rows, x0 from beq = -Ad x0, xdes from f = -Pbig xdes
  double F[8] = {0}, x0[4], xdes[16], Ad[16], Bd[8];
  for (int a = 0; a < 2; a++) for (int j = 0; j < cc; j++) F[4 * a + j] = -Aeq[(16 + a) + (size_t)(24 + j) * p];
  host_lip(g, Ad, Bd);
  {  // per-axis 2x2 inverse
    double det = Ad[0] * Ad[5] - Ad[1] * Ad[4];
    for (int bx = 0; bx < 2; bx++) {
      double r0 = -beq[2 * bx], r1 = -beq[2 * bx + 1];
      x0[2 * bx] = (Ad[5] * r0 - Ad[1] * r1) / det;
      x0[2 * bx + 1] = (-Ad[4] * r0 + Ad[0] * r1) / det;
    }
  }
  for (int i = 0; i < 16; i++) xdes[i] = -f[i] / (i >= 12 ? g->Pw : g->Qw);
  std::vector<double> Q2, f2, A2, b2, G2, h2;
  host_pack_dense(g, x0, xdes, F, cc, Q2, f2, A2, b2, G2, h2);
  if (!close_vec(Q, Q2, 1e-12) || !close_vec(f, f2, 1e-12) || !close_vec(Aeq, A2, 1e-12) || !close_vec(beq, b2, 1e-9) ||
      !close_vec(G, G2, 1e-12) || !close_vec(h, h2, 1e-12))
    return fail(CBFNMPC_ERR_STRUCTURE, "matrices do not have the MPC_dist structure for the context's gains");
  // one-instance SoA inputs (B = 1): contact pattern that reproduces F's column order
  CK(cudaSetDevice(c->device));
  double hin[34];
  double feet[8] = {0};
  int contact[4];
  if (cc == 4) { for (int i = 0; i < 8; i++) feet[i] = F[i]; contact[0] = contact[1] = contact[2] = contact[3] = 1; }
  else { feet[1] = F[0]; feet[2] = F[1]; feet[5] = F[4]; feet[6] = F[5]; contact[0] = 0; contact[1] = 1; contact[2] = 1; contact[3] = 0; }
  memcpy(hin, x0, 32); memcpy(hin + 4, xdes, 128); memcpy(hin + 20, feet, 64);
  for (int i = 0; i < 4; i++) hin[28 + i] = 0.0;
  hin[32] = obs ? obs[0] : 0.0; hin[33] = obs ? obs[1] : 0.0;
  cudaStream_t st = c->stream;
  if (c->have_last) { CK(cudaEventSynchronize(c->ev_last)); c->have_last = false; }
  CK(cudaMemcpyAsync(c->d_in, hin, sizeof(hin), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->d_contact, contact, sizeof(contact), cudaMemcpyHostToDevice, st));
  double xs[41] = {0};
  if (mode_nlp) { memcpy(xs, x0_start, sizeof(double) * nv); CK(cudaMemcpyAsync(c->d_xstart, xs, sizeof(xs), cudaMemcpyHostToDevice, st)); }
  double* d = c->d_in;
  cbfnmpc_in din = {d, d + 4, d + 20, c->d_contact, d + 32, 1, d + 28};
  double* o = c->d_out;
  cbfnmpc_out dout;
  memset(&dout, 0, sizeof(dout));
  dout.X = o + O_X; dout.U = o + O_U; dout.lam = o + O_LAM; dout.slack = o + O_SLACK; dout.kkt = o + O_KKT; dout.status = c->d_outi;
  int rc = launch(c, 1, mode_nlp ? CBFNMPC_MODE_QP_NLP : CBFNMPC_MODE_QP, &din, &dout, mode_nlp ? c->d_xstart : nullptr, st);
  if (rc) return rc;
  double hall[O_ROWS]; double ho[41]; int hs = 0; double hk = 0;
  CK(cudaMemcpyAsync(hall, o, sizeof(double) * O_ROWS, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&hs, c->d_outi, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  memcpy(ho, hall + O_X, sizeof(double) * 24); memcpy(ho + 24, hall + O_LAM, sizeof(double) * 16); ho[40] = hall[O_SLACK]; hk = hall[O_KKT];
  memcpy(x_out, ho, sizeof(double) * 24);
  for (int i = 0; i < 4 * cc; i++) x_out[24 + i] = ho[24 + i];
  x_out[nv - 1] = ho[40];
  if (status) *status = hs;
  if (kkt) *kkt = hk;
  return CBFNMPC_OK;
}

#include "cbfnmpc_ext_api.cuh"

}  // extern "C"
