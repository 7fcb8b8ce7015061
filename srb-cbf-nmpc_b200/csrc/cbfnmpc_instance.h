// cbfnmpc_instance.h -- one warp = one run_NMPC(): load (updateState + planners' outputs), closest obstacle,
// QP stage, NLP stage, Bezier epilogue, store.   Shared by the CUDA kernel and the tests/emu harness.
#pragma once
#include "cbfnmpc_solver.h"


template <int C>
DEV void active_bits(const LaneSlots& ls, int q, int slot, unsigned w[4]) {
  typedef Dm<C> d;
  const SlotMeta& m = ls.m[q];
  int rlo = -1, rhi = -1;  // reference row numbers (DESIGN.md row map)
  if (slot < 16) { rhi = 12 + slot; rlo = 28 + slot; }
  else if (slot < 24) { rhi = 44 + (slot - 16); rlo = 52 + (slot - 16); }
  else if (slot < d::S_VEL) { rlo = 60 + (slot - 24); rhi = 60 + 4 * C + (slot - 24); }
  else if (slot < d::S_FRIC) { rhi = d::MQP + 4 + (slot - d::S_VEL); rlo = d::MQP + 12 + (slot - d::S_VEL); }
  else if (slot < d::S_DIST) { rhi = slot - d::S_FRIC; rlo = 6 + (slot - d::S_FRIC); }
  else if (slot < d::NSLOT) { rlo = d::MQP + (slot - d::S_DIST); }
  if (m.vlo && rlo >= 0 && ls.zlo[q] > ls.slo[q]) w[rlo >> 5] |= 1u << (rlo & 31);
  if (m.vhi && rhi >= 0 && ls.zhi[q] > ls.shi[q]) w[rhi >> 5] |= 1u << (rhi & 31);
}

template <int C>
DEV void run_instance(WSB& ws, const KParams& P, const BatchIO& io, int b) {
  typedef Dm<C> d;
  constexpr int NW = d::NW;
  const int lane = lane_id();
  const int B = io.B;
  // D = F E, Bt = Bd D   (E = [I; -1'])
  if (lane < 2 * NW) { int k = lane / NW, a = lane % NW; ws.D[k * NW + a] = ws.F[4 * k + a] - ws.F[4 * k + C - 1]; }
  wsync();
  if (lane < 4 * NW) {
    int m = lane / NW, a = lane % NW;
    ws.Bt[m * NW + a] = ((m & 1) ? P.bv : P.bp) * ws.D[(m >> 1) * NW + a];
  }
  LaneSlots ls;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    ls.m[q] = slot_meta<C>(lane + 32 * q, P);
    ls.zlo[q] = ls.zhi[q] = 0.0; ls.slo[q] = ls.shi[q] = 1.0;
  }
  wsync();
  SolveInfo info; info.status = 1; info.qp_iters = 0; info.qp_flag = 3; info.nlp_iters = 0; info.n_fact = 0; info.kkt = 0.0;
  // Lockstep state machine: every warp of the CTA executes one iteration of its current stage per round and then meets
  // the others at a CTA barrier, so that the warps of an SM share the instruction stream (DESIGN.md "lockstep").
  enum { PH_QP = 0, PH_NLP = 1, PH_DONE = 2 };
  int phase = PH_QP;
  QpState qs; NlpState ns;
  bool qp_done;
  if (io.xstart == nullptr) {
    qp_done = qp_init<C>(ws, P, ls, info, qs);
  } else {
#pragma unroll
    for (int q = 0; q < 2; q++) { int v = lane + 32 * q; if (v < d::NV) ws.x[v] = io.xstart[(size_t)v * B + b]; }
    info.qp_flag = 0;
    qp_done = true;
    wsync();
  }
  for (;;) {
    if (phase == PH_QP) {
      if (!qp_done) qp_done = qp_iter<C>(ws, P, ls, info, qs);
      if (qp_done) {
        if (io.xqp && lane < 24) io.xqp[(size_t)lane * B + b] = ws.x[lane];
        if (io.mode == 1) { nlp_init<C>(ws, P, ls, info, ns); phase = PH_NLP; }
        else {
          // QP-only answer: unscaled KKT error of the QP
          double* sb = ws.u.a.sb;
          double ec = 0.0, ep = 0.0;
#pragma unroll
          for (int q = 0; q < 2; q++) {
            sb[lane + 32 * q] = ls.zlo[q] - ls.zhi[q];
            if (ls.m[q].qp) {
              double e = slot_e(ls.m[q], ws.x, 0.0, 0.0, 0.0);
              ec = fmax(ec, fmax(fabs(ls.slo[q] * ls.zlo[q]), fabs(ls.shi[q] * ls.zhi[q])));
              ep = fmax(ep, fmax(fmax(ls.m[q].lo - e, e - ls.m[q].hi), 0.0));
            }
          }
          wsync();
          double rd = fabs(grad_plus_Aty<C>(ws, P, ws.x, ws.y, lane, true) - gatherJT<C>(sb, ws.gv, lane));
          if (lane + 32 < d::NV) rd = fmax(rd, fabs(grad_plus_Aty<C>(ws, P, ws.x, ws.y, lane + 32, true) - gatherJT<C>(sb, ws.gv, lane + 32)));
          double re = (lane < 28) ? fabs(eq_residual<C>(ws, P, ws.x, lane, false)) : 0.0;
          info.kkt = wmax_nn(fmax(fmax(rd, re), fmax(ec, ep)));
          info.status = (info.qp_flag == 0) ? 0 : 1;
          wsync();
          phase = PH_DONE;
        }
      }
    } else if (phase == PH_NLP) {
      if (nlp_iter<C>(ws, P, ls, info, ns)) phase = PH_DONE;
    }
    if (!cta_or(phase != PH_DONE)) break;
  }
  // ---- store
  if (lane < 16) { if (io.X) io.X[(size_t)lane * B + b] = ws.x[lane]; }
  else if (lane < 24) { if (io.U) io.U[(size_t)(lane - 16) * B + b] = ws.x[lane]; }
  if (io.lam && lane < 16) io.lam[(size_t)lane * B + b] = (lane < 4 * C) ? ws.x[24 + lane] : 0.0;
  // Bezier fit  alpha = Y Minv'   (MPC_dist.cpp:784-855, effective semantics)
  if (lane < 20) {
    int r = lane / 5, j = lane % 5;
    double s = P.minv[5 * j] * ws.abuf[r];
#pragma unroll
    for (int i = 1; i < 5; i++) s += P.minv[5 * j + i] * ws.x[4 * (i - 1) + r];
    if (io.alpha) io.alpha[(size_t)lane * B + b] = s;
  } else if (lane < 24) {
    if (io.abuf_next) io.abuf_next[(size_t)(lane - 20) * B + b] = ws.x[12 + (lane - 20)];
  }
  unsigned w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int q = 0; q < 2; q++) active_bits<C>(ls, q, lane + 32 * q, w);
#pragma unroll
  for (int k = 0; k < 4; k++) w[k] = wor(w[k]);
  if (lane == 0) {
    if (io.slack) io.slack[b] = ws.x[d::XD];
    if (io.kkt) io.kkt[b] = info.kkt;
    if (io.status) io.status[b] = info.status;
    if (io.qp_iters) io.qp_iters[b] = info.qp_iters;
    if (io.nlp_iters) io.nlp_iters[b] = info.nlp_iters;
    if (io.n_fact) io.n_fact[b] = info.n_fact;
    if (io.active) {
      io.active[b] = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
      io.active[(size_t)B + b] = (unsigned long long)w[2] | ((unsigned long long)w[3] << 32);
    }
  }
}

// One warp: instance b.  Implements updateState's copies and the argmin of MPC_dist.cpp:371-396.
DEV void solve_warp(WSB& ws, const KParams& P, const BatchIO& io, int b) {
  const int lane = lane_id();
  const int B = io.B;
  if (b >= B) {  // padding warp of the last CTA: only keeps the CTA barriers balanced
    for (;;) if (!cta_or(false)) break;
    return;
  }
  // ---- load (SoA, one instance)
  if (lane < 4) { ws.x0[lane] = io.x0[(size_t)lane * B + b]; ws.abuf[lane] = io.abuf ? io.abuf[(size_t)lane * B + b] : 0.0; }
  if (lane >= 16) ws.xdes[lane - 16] = io.xdes[(size_t)(lane - 16) * B + b];
  double foot = (lane >= 4 && lane < 12) ? io.feet[(size_t)(lane - 4) * B + b] : 0.0;
  int c0 = io.contact[b], c1 = io.contact[(size_t)B + b], c2 = io.contact[2 * (size_t)B + b], c3 = io.contact[3 * (size_t)B + b];
  int csum = c0 + c1 + c2 + c3;
  // stance matrix (MPC_dist.cpp:1238-1252)
  int C = -1;
  if (csum == 4) C = 4;
  else if (csum == 2 && (c0 == 0 || c0 == 1)) C = 2;
  if (lane >= 4 && lane < 12) {
    int a = (lane - 4) >> 2, i = (lane - 4) & 3;  // feet[a][i]
    if (C == 4) ws.F[4 * a + i] = foot;
    else if (C == 2) {
      if (c0 == 0) { if (i == 1) ws.F[4 * a] = foot; if (i == 2) ws.F[4 * a + 1] = foot; }
      else { if (i == 0) ws.F[4 * a] = foot; if (i == 3) ws.F[4 * a + 1] = foot; }
    }
  }
  // closest obstacle to the current position: first strict minimum of the Euclidean distance
  if (io.mode == 1) {
    double px = io.x0[b], py = io.x0[2 * (size_t)B + b];
    double best = 1000.0; int bi = 0x7fffffff;
    for (int k = lane; k < io.n_obs; k += 32) {
      double dx = px - io.obs[(size_t)k * B + b], dy = py - io.obs[(size_t)(io.n_obs + k) * B + b];
      double dd = sqrt(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)));
      if (dd < best) { best = dd; bi = k; }
    }
    double gbest = wmin_nn(best);
    int cand = (best == gbest && bi != 0x7fffffff) ? bi : 0x7fffffff;
    int gi = wmin_i(cand);
    if (gi == 0x7fffffff) gi = 0;
    if (lane == 0) { ws.obs[0] = io.obs[(size_t)gi * B + b]; ws.obs[1] = io.obs[(size_t)(io.n_obs + gi) * B + b]; }
  } else if (lane == 0) { ws.obs[0] = 0.0; ws.obs[1] = 0.0; }
  wsync();
  if (C == 4) run_instance<4>(ws, P, io, b);
  else if (C == 2) run_instance<2>(ws, P, io, b);
  else {
    if (lane == 0) {
      if (io.status) io.status[b] = 4;
      if (io.kkt) io.kkt[b] = 0.0;
      if (io.qp_iters) io.qp_iters[b] = 0;
      if (io.nlp_iters) io.nlp_iters[b] = 0;
      if (io.n_fact) io.n_fact[b] = 0;
      if (io.slack) io.slack[b] = 0.0;
      if (io.active) { io.active[b] = 0ull; io.active[(size_t)B + b] = 0ull; }
    }
    for (;;) if (!cta_or(false)) break;
  }
}
