// cbfnmpc_solver.h -- one CBF-NMPC instance per warp, FP64, everything resident in shared memory / registers.
//
// Replaces, for a whole batch at once, what MPC_dist::run_NMPC() does per agent per gait domain
// (/root/reference/src/MPC_dist.cpp:81-454):
//   stage 1 (QP)  : iSWIFT's Mehrotra predictor-corrector (optimization/iSWIFT/src/Prime.c:127-230) on the LIP
//                   tracking QP of MPC_dist.cpp:164-321;
//   stage 2 (NLP) : the same QP plus the 4 squared-distance rows and 8 velocity boxes of
//                   include/dec_vars_constr_cost.h:245-324, solved by the primal-dual filter interior-point method
//                   "IPM-F" of DESIGN.md (feasible w.r.t. inequalities, exact Hessian, inertia correction);
//   epilogue      : Bezier fit of MPC_dist.cpp:784-855.
// The Newton/KKT systems of both stages are solved by one structured factorisation: per-stage elimination of
// (u_j, lambda_j) through u = F lambda, 1'lambda = 1, a Riccati/Cholesky recursion over the 4-step horizon, and a
// 5x5 Schur complement for the global slack variable and the (capped) distance rows.
//
// Lane roles: 32 lanes <-> up to 64 "slots" (two-sided scalar constraints, 2 per lane); <-> 41 variables (2 per
// lane); <-> 28 equality rows; Riccati matrix products use up to 28 lanes; right-hand sides run 4 lanes per RHS.
#pragma once
#include "warp_prims.h"
#include "cbfnmpc_types.h"


#if defined(CBF_HOST_EMU) && defined(CBF_TRACE)
#include <stdio.h>
#define CBF_DBG(...) do { if (lane_id() == 0) fprintf(stderr, __VA_ARGS__); } while (0)
#else
#define CBF_DBG(...) do { } while (0)
#endif

template <int C>
struct Dm {
  static constexpr int NW = C - 1, NV = 25 + 4 * C, XD = 24 + 4 * C;
  static constexpr int S_VEL = 24 + 4 * C, S_FRIC = S_VEL + 8, S_DIST = S_FRIC + 6, NSLOT = S_DIST + 4;
  static constexpr int NR = NW * (NW + 1) / 2, NS = 4 * NW, MQP = 60 + 8 * C;
};

// Per-warp shared-memory workspace (one instance), ~6.8 KB.  The layout does not depend on the number of stance
// feet; arrays are sized for 4.
struct WSB {
  double x[42], y[28], dx[42], dy[28], r1[42], r2[28];
  double x0[4], xdes[16], F[8], D[6], Bt[12], obs[2], abuf[4];
  double hx[16], hxy[4], hu[8], hxu[8], hl[16];
  double gv[8], bd[8], sc[4], srinv[4], zdist[4], hd;
  double P[4][16];
  double St[4][12], K[4][12], Lc[4][6];
  double rw[12], cv[16], wsol[12];
  double pos[6][8];
  double T[10], cb[4], Tic[4], tb[4], den, ddl;
  double filt[16];
  int capped[4];
  double pvb[5][16], kvb[5][12];  // backward-pass vectors of the 5 border columns (live until the next factorisation)
  union {
    struct { double sa[64], sb[64], PA[16], PB[12], Rt[6], Qt[10]; } a;  // assembly + factor temporaries
    struct { double pv0[16], kv0[12]; } s;                               // backward-pass vectors of the main RHS
  } u;
};
template <int C> using WS = WSB;
DEV double* pv_of(WSB& ws, int r) { return r == 0 ? ws.u.s.pv0 : ws.pvb[r - 1]; }
DEV double* kv_of(WSB& ws, int r) { return r == 0 ? ws.u.s.kv0 : ws.kvb[r - 1]; }

// ---------------------------------------------------------------------------------------------------------
// slots: two-sided scalar constraints  lo <= e(x) <= hi   (reference rows: DESIGN.md "row map")
// ---------------------------------------------------------------------------------------------------------
struct SlotMeta { int kind, i1, i2; double lo, hi; bool vlo, vhi, qp; };

template <int C>
DEV SlotMeta slot_meta(int s, const KParams& P) {
  typedef Dm<C> d;
  SlotMeta m; m.kind = 3; m.i1 = 0; m.i2 = 0; m.lo = -1.0; m.hi = 1.0; m.vlo = m.vhi = m.qp = false;
  if (s < 24) { m.kind = 0; m.i1 = s; m.lo = -P.box; m.hi = P.box; m.vlo = m.vhi = m.qp = true; }          // X,U boxes  MD:294-297,317-318
  else if (s < d::S_VEL) { m.kind = 0; m.i1 = s; m.lo = 0.0; m.hi = 1.0; m.vlo = m.vhi = m.qp = true; }       // 0<=lambda<=1 MD:299-300,320-321
  else if (s < d::S_FRIC) { int r = s - d::S_VEL; m.kind = 0; m.i1 = (r < 4) ? 4 * r + 1 : 4 * (r - 4) + 3;   // velocity   DV:270-279,306-317
    m.lo = -P.vmax; m.hi = P.vmax; m.vlo = m.vhi = true; }
  else if (s < d::S_DIST) { int q = s - d::S_FRIC, i = q >> 1, a = q & 1; m.kind = 1; m.i1 = 4 * i + 2 * a;    // |p-u|<=fb  MD:278-292,315
    m.i2 = 16 + 2 * (i + 1) + a; m.lo = -P.fb; m.hi = P.fb; m.vlo = m.vhi = m.qp = true; }
  else if (s < d::NSLOT) { int k = s - d::S_DIST; m.kind = 2; m.i1 = 4 * k; m.i2 = 4 * k + 2; m.lo = P.eps;    // distance   DV:262-265,299-302
    m.hi = 0.0; m.vlo = true; }
  return m;
}

DEV double slot_e(const SlotMeta& m, const double* x, double ox, double oy, double xdl) {
  if (m.kind == 0) return x[m.i1];
  if (m.kind == 1) return x[m.i1] - x[m.i2];
  if (m.kind == 2) { double a = x[m.i1] - ox, b = x[m.i2] - oy; return a * a + b * b + xdl; }
  return 0.0;
}
// directional derivative of e along dx (gx, gy = 2 (p - o) for distance slots)
DEV double slot_de(const SlotMeta& m, const double* dx, double gx, double gy, double ddl) {
  if (m.kind == 0) return dx[m.i1];
  if (m.kind == 1) return dx[m.i1] - dx[m.i2];
  if (m.kind == 2) return gx * dx[m.i1] + gy * dx[m.i2] + ddl;
  return 0.0;
}

// (J' w)_v for per-slot weights w (weight of +grad e), v = variable index
template <int C>
DEV double gatherJT(const double* w, const double* gv, int v) {
  typedef Dm<C> d;
  if (v < 16) {
    int j = v >> 2, i = v & 3;
    double t = w[v];
    if (i & 1) t += w[d::S_VEL + ((i == 1) ? j : 4 + j)];
    else {
      int a = i >> 1;
      if (j <= 2) t += w[d::S_FRIC + 2 * j + a];
      t += gv[2 * j + a] * w[d::S_DIST + j];
    }
    return t;
  }
  if (v < 24) { int q = v - 16, j = q >> 1, a = q & 1; double t = w[v]; if (j >= 1) t -= w[d::S_FRIC + 2 * (j - 1) + a]; return t; }
  if (v < d::XD) return w[v];
  return (w[d::S_DIST] + w[d::S_DIST + 1]) + (w[d::S_DIST + 2] + w[d::S_DIST + 3]);
}

// grad f (x) + A' y  for variable v          (cost: DV:423-438, MD:190-210; equality rows: MD:232-261)
template <int C>
DEV double grad_plus_Aty(const WS<C>& ws, const KParams& P, const double* xx, const double* yy, int v, bool with_grad) {
  typedef Dm<C> d;
  if (v < 16) {
    int j = v >> 2, i = v & 3, b = i >> 1, r = i & 1;
    double t = -yy[v];
    if (j < 3) t += (r ? P.a12 : P.a11) * yy[4 * (j + 1) + 2 * b] + (r ? P.a22 : P.a21) * yy[4 * (j + 1) + 2 * b + 1];
    if (with_grad) t += ((j == 3) ? P.Pw : P.Qw) * (xx[v] - ws.xdes[v]);
    return t;
  }
  if (v < 24) {
    int q = v - 16, j = q >> 1, a = q & 1;
    double t = P.bp * yy[4 * j + 2 * a] + P.bv * yy[4 * j + 2 * a + 1] + yy[16 + q];
    if (with_grad) t += P.Rw * xx[v];
    return t;
  }
  if (v < d::XD) {
    int q = v - 24, j = q / C, i = q % C;
    return -(ws.F[i] * yy[16 + 2 * j] + ws.F[4 + i] * yy[17 + 2 * j]) + yy[24 + j];
  }
  return with_grad ? P.slw * xx[v] : 0.0;
}

// (Aeq xx - beq)_t, or (Aeq xx)_t when homogeneous
template <int C>
DEV double eq_residual(const WS<C>& ws, const KParams& P, const double* xx, int t, bool homogeneous) {
  if (t < 16) {
    int j = t >> 2, i = t & 3, b = i >> 1, r = i & 1;
    double xa, xb;
    if (j == 0) { xa = homogeneous ? 0.0 : ws.x0[2 * b]; xb = homogeneous ? 0.0 : ws.x0[2 * b + 1]; }
    else { xa = xx[4 * (j - 1) + 2 * b]; xb = xx[4 * (j - 1) + 2 * b + 1]; }
    return (r ? P.a21 : P.a11) * xa + (r ? P.a22 : P.a12) * xb - xx[t] + (r ? P.bv : P.bp) * xx[16 + 2 * j + b];
  }
  if (t < 24) {
    int q = t - 16, j = q >> 1, a = q & 1;
    double s = xx[16 + q];
#pragma unroll
    for (int i = 0; i < C; i++) s -= ws.F[4 * a + i] * xx[24 + C * j + i];
    return s;
  }
  int j = t - 24;
  double s = homogeneous ? 0.0 : -1.0;
#pragma unroll
  for (int i = 0; i < C; i++) s += xx[24 + C * j + i];
  return s;
}

// ---------------------------------------------------------------------------------------------------------
// stage Hessians from the slot weights in ws.u.a.sa (dw is added later, inside the factorisation)
// ---------------------------------------------------------------------------------------------------------
template <int C>
DEVNI void assemble_hessian(WS<C>& ws, const KParams& P, bool nl, double rho) {
  typedef Dm<C> d;
  const int lane = lane_id();
  const double* sa = ws.u.a.sa;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    int v = lane + 32 * q;
    if (v < 16) {
      int j = v >> 2, i = v & 3;
      double h = ((j == 3) ? P.Pw : P.Qw) + sa[v];
      if (i & 1) h += sa[d::S_VEL + ((i == 1) ? j : 4 + j)];
      else {
        int a = i >> 1;
        if (j <= 2) h += sa[d::S_FRIC + 2 * j + a];
        if (nl) { double g = ws.gv[2 * j + a]; h += ws.sc[j] * g * g - 2.0 * ws.zdist[j]; }
      }
      ws.hx[v] = h;
    } else if (v < 24) {
      int qq = v - 16, j = qq >> 1, a = qq & 1;
      double sf = (j >= 1) ? sa[d::S_FRIC + 2 * (j - 1) + a] : 0.0;
      ws.hu[qq] = P.Rw + sa[v] + sf;
      ws.hxu[qq] = -sf;
    } else if (v < d::XD) {
      ws.hl[v - 24] = rho + sa[v];
    } else if (v < d::XD + 4) {  // four spare lanes: distance-row border data
      int j = v - d::XD;
      double s = nl ? ws.sc[j] : 0.0;
      ws.hxy[j] = s * ws.gv[2 * j] * ws.gv[2 * j + 1];
      ws.bd[2 * j] = s * ws.gv[2 * j];
      ws.bd[2 * j + 1] = s * ws.gv[2 * j + 1];
    } else if (v == d::XD + 4) {
      ws.hd = P.slw + (nl ? ((ws.sc[0] + ws.sc[1]) + (ws.sc[2] + ws.sc[3])) : 0.0);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Riccati / Cholesky factorisation over the horizon.  Returns false when a pivot block is not positive definite.
// P_j (j = 1..4) -> ws.P[j-1];  S~_j, K_j, chol(R~_j) -> ws.St[j], ws.K[j], ws.Lc[j].
// ---------------------------------------------------------------------------------------------------------
template <int NW>
DEV bool chol_small(const double* R, double* L) {  // R packed lower (00,10,11,20,21,22); L = (1/l00, l10, 1/l11, l20, l21, 1/l22)
  bool ok = true;
  double d0 = R[0];
  ok = ok && (d0 > 0.0) && (d0 < 1e300);
  double i0 = rsq(d0);
  L[0] = i0;
  if (NW == 3) {
    double l10 = R[1] * i0, l20 = R[3] * i0;
    double d1 = R[2] - l10 * l10;
    ok = ok && (d1 > 0.0) && (d1 < 1e300);
    double i1 = rsq(d1);
    double l21 = (R[4] - l20 * l10) * i1;
    double d2 = R[5] - l20 * l20 - l21 * l21;
    ok = ok && (d2 > 0.0) && (d2 < 1e300);
    double i2 = rsq(d2);
    L[1] = l10; L[2] = i1; L[3] = l20; L[4] = l21; L[5] = i2;
  }
  return ok;
}
template <int NW>
DEV void chol_solve(const double* L, double* b) {  // b <- (L L')^-1 b
  if (NW == 1) { b[0] = b[0] * L[0] * L[0]; return; }
  double y0 = b[0] * L[0];
  double y1 = (b[1] - L[1] * y0) * L[2];
  double y2 = (b[2] - L[3] * y0 - L[4] * y1) * L[5];
  double k2 = y2 * L[5];
  double k1 = (y1 - L[4] * k2) * L[2];
  double k0 = (y0 - L[1] * k1 - L[3] * k2) * L[0];
  b[0] = k0; b[1] = k1; b[2] = k2;
}

template <int C>
DEVNI bool riccati_factor(WS<C>& ws, const KParams& P, double dw) {
  typedef Dm<C> d;
  constexpr int NW = d::NW;
  const int lane = lane_id();
  double* PA = ws.u.a.PA; double* PB = ws.u.a.PB; double* Rt = ws.u.a.Rt; double* Qt = ws.u.a.Qt;
  if (lane < 16) {
    int i = lane >> 2, k = lane & 3;
    double v = 0.0;
    if (i == k) v = ws.hx[12 + i] + dw;
    else if ((i == 0 && k == 2) || (i == 2 && k == 0)) v = ws.hxy[3];
    ws.P[3][lane] = v;
  }
  wsync();
  for (int j = 3; j >= 0; --j) {
    const double* Pn = ws.P[j];
    // round 1: PA = P+ A (16), PB = P+ Bt (4 NW)
    if (lane < 16) {
      int i = lane >> 2, k = lane & 3, b = k >> 1, r = k & 1;
      PA[lane] = Pn[4 * i + 2 * b] * (r ? P.a12 : P.a11) + Pn[4 * i + 2 * b + 1] * (r ? P.a22 : P.a21);
    } else if (lane < 16 + 4 * NW) {
      int t = lane - 16, i = t / NW, a = t % NW;
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < 4; m++) s += Pn[4 * i + m] * ws.Bt[m * NW + a];
      PB[t] = s;
    }
    wsync();
    // round 2: R~ (NR), S~ (4 NW), Q~ (10)
    if (lane < d::NR) {
      int a = 0, b = 0;
      if (NW == 3) { a = (lane >= 3) ? 2 : ((lane >= 1) ? 1 : 0); b = lane - ((a * (a + 1)) >> 1); }
      double s = ws.D[a] * (ws.hu[2 * j] + dw) * ws.D[b] + ws.D[NW + a] * (ws.hu[2 * j + 1] + dw) * ws.D[NW + b];
      s += ws.hl[C * j + C - 1] + dw;
      if (a == b) s += ws.hl[C * j + a] + dw;
#pragma unroll
      for (int m = 0; m < 4; m++) s += ws.Bt[m * NW + a] * PB[m * NW + b];
      Rt[lane] = s;
    } else if (lane < d::NR + d::NS) {
      if (j >= 1) {
        int t = lane - d::NR, a = t >> 2, k = t & 3;
        double s = (k & 1) ? 0.0 : ws.D[(k >> 1) * NW + a] * ws.hxu[2 * j + (k >> 1)];
#pragma unroll
        for (int m = 0; m < 4; m++) s += ws.Bt[m * NW + a] * PA[4 * m + k];
        ws.St[j][a * 4 + k] = s;
      }
    } else if (lane < d::NR + d::NS + 10) {
      if (j >= 1) {
        int t = lane - d::NR - d::NS, i = (t >= 6) ? 3 : ((t >= 3) ? 2 : ((t >= 1) ? 1 : 0)), k = t - ((i * (i + 1)) >> 1), bi = i >> 1, ri = i & 1;
        double s = (ri ? P.a12 : P.a11) * PA[4 * (2 * bi) + k] + (ri ? P.a22 : P.a21) * PA[4 * (2 * bi + 1) + k];
        if (i == k) s += ws.hx[4 * (j - 1) + i] + dw;
        else if (i == 2 && k == 0) s += ws.hxy[j - 1];
        Qt[t] = s;
      }
    }
    wsync();
    // round 3: Cholesky of R~ (every lane, same data -> same result)
    double L[6];
    bool ok = chol_small<NW>(Rt, L);
    if (!ok) return false;
    if (lane < d::NR) ws.Lc[j][lane] = L[lane];
    if (j == 0) break;
    // round 4: K = R~^-1 S~ (one column per lane)
    if (lane < 4) {
      double col[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int a = 0; a < NW; a++) col[a] = ws.St[j][a * 4 + lane];
      chol_solve<NW>(L, col);
#pragma unroll
      for (int a = 0; a < NW; a++) ws.K[j][a * 4 + lane] = col[a];
    }
    wsync();
    // round 5: P_j = Q~ - S~' K
    if (lane < 10) {
      int i = (lane >= 6) ? 3 : ((lane >= 3) ? 2 : ((lane >= 1) ? 1 : 0)), k = lane - ((i * (i + 1)) >> 1);
      double s = Qt[lane];
#pragma unroll
      for (int a = 0; a < NW; a++) s -= ws.St[j][a * 4 + i] * ws.K[j][a * 4 + k];
      ws.P[j - 1][4 * i + k] = s;
      ws.P[j - 1][4 * k + i] = s;
    }
    wsync();
  }
  wsync();
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// Backward / forward substitution, 4 lanes per right-hand side.  RHS r: 0 main, 1 slack-variable column,
// 2..5 distance-row columns.  `rmask` selects which RHS run.
// ---------------------------------------------------------------------------------------------------------
template <int C>
DEV double rhs_x(const WS<C>& ws, int r, int j, int i) {  // x-part of RHS r at state stage j (0..3 <-> x_1..x_4)
  if (r == 0) return ws.r1[4 * j + i];
  if (i & 1) return 0.0;
  if (r == 1) return ws.bd[2 * j + (i >> 1)];
  return (r - 2 == j) ? ws.gv[2 * j + (i >> 1)] : 0.0;
}

template <int C>
DEVNI void kkt_backward(WS<C>& ws, const KParams& P, unsigned rmask) {
  constexpr int NW = Dm<C>::NW;
  const int lane = lane_id(), r = lane >> 2, i = lane & 3;
  const bool on = (r < 6) && ((rmask >> r) & 1u);
  const int rr = on ? r : 0;
  double pn = on ? rhs_x<C>(ws, r, 3, i) : 0.0;  // p_4
  double* pvr = pv_of(ws, rr); double* kvr = kv_of(ws, rr);
  if (on) pvr[12 + i] = pn;
  for (int j = 3; j >= 0; --j) {
    double qi = -pn;
    if (rr == 0 && on) {
#pragma unroll
      for (int m = 0; m < 4; m++) qi += ws.P[j][4 * i + m] * ws.cv[4 * j + m];
    }
    double q[4];
#pragma unroll
    for (int m = 0; m < 4; m++) q[m] = shfl_g4(qi, m);
    double k[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < NW; a++) {
      double s = (rr == 0) ? ws.rw[NW * j + a] : 0.0;
#pragma unroll
      for (int m = 0; m < 4; m++) s -= ws.Bt[m * NW + a] * q[m];
      k[a] = s;
    }
    chol_solve<NW>(ws.Lc[j], k);
    if (on && i < NW) kvr[NW * j + i] = (i == 0) ? k[0] : ((i == 1) ? k[1] : k[2]);
    if (j >= 1) {
      int b = i >> 1, ri = i & 1;
      double s = rhs_x<C>(ws, rr, j - 1, i) - ((ri ? P.a12 : P.a11) * q[2 * b] + (ri ? P.a22 : P.a21) * q[2 * b + 1]);
#pragma unroll
      for (int a = 0; a < NW; a++) s -= ws.St[j][a * 4 + i] * k[a];
      pn = on ? s : 0.0;
      if (on) pvr[4 * (j - 1) + i] = pn;
    }
  }
  wsync();
}

// mode 0: store only positions of x_1..x_4 into ws.pos[r]; mode 1 (r = 0 only): store dx, w, dpi
template <int C>
DEVNI void kkt_forward(WS<C>& ws, const KParams& P, unsigned rmask, int mode) {
  constexpr int NW = Dm<C>::NW;
  const int lane = lane_id(), r = lane >> 2, i = lane & 3;
  const bool on = (r < 6) && ((rmask >> r) & 1u);
  const int rr = on ? r : 0;
  const double* pvr = pv_of(ws, rr); const double* kvr = kv_of(ws, rr);
  double xp = 0.0;
  for (int j = 0; j < 4; j++) {
    double xv[4];
#pragma unroll
    for (int m = 0; m < 4; m++) xv[m] = shfl_g4(xp, m);
    double w[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < NW; a++) {
      double s = kvr[NW * j + a];
      if (j > 0) {
#pragma unroll
        for (int m = 0; m < 4; m++) s -= ws.K[j][a * 4 + m] * xv[m];
      }
      w[a] = s;
    }
    int b = i >> 1, ri = i & 1;
    double xn = (ri ? P.a21 : P.a11) * xv[2 * b] + (ri ? P.a22 : P.a12) * xv[2 * b + 1];
#pragma unroll
    for (int a = 0; a < NW; a++) xn += ws.Bt[i * NW + a] * w[a];
    if (rr == 0) xn += ws.cv[4 * j + i];
    if (!on) xn = 0.0;
    double xnv[4];
#pragma unroll
    for (int m = 0; m < 4; m++) xnv[m] = shfl_g4(xn, m);
    if (on) {
      if (mode == 0) {
        if (!(i & 1)) ws.pos[rr][2 * j + (i >> 1)] = xn;
      } else if (rr == 0) {
        double s = -pvr[4 * j + i];
#pragma unroll
        for (int m = 0; m < 4; m++) s += ws.P[j][4 * i + m] * xnv[m];
        ws.dx[4 * j + i] = xn;
        ws.dy[4 * j + i] = s;
        if (i < NW) ws.wsol[NW * j + i] = (i == 0) ? w[0] : ((i == 1) ? w[1] : w[2]);
      }
    }
    xp = xn;
  }
  wsync();
}

// Border part of the factorisation (NLP stage): columns of the slack variable and of the capped distance rows.
template <int C>
DEVNI bool border_factor(WS<C>& ws, const KParams& P, double dw) {
  const int lane = lane_id();
  const bool anycap = (ws.capped[0] | ws.capped[1] | ws.capped[2] | ws.capped[3]) != 0;
  const unsigned rmask = anycap ? 0x3eu : 0x02u;
  kkt_backward<C>(ws, P, rmask);
  kkt_forward<C>(ws, P, rmask, 0);
  // T (4x4, capped rows only), cb, a  -- computed redundantly by every lane
  double T[10], cb[4];
  {
    const int ti[10] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3}, tk[10] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3};
#pragma unroll
    for (int t = 0; t < 10; t++) {
      int k = ti[t], l = tk[t];
      double v;
      if (ws.capped[k] && ws.capped[l]) {
        v = ws.gv[2 * k] * ws.pos[2 + l][2 * k] + ws.gv[2 * k + 1] * ws.pos[2 + l][2 * k + 1];
        if (k == l) v += ws.srinv[k];
      } else v = (k == l) ? 1.0 : 0.0;
      T[t] = v;
    }
  }
  double a = ws.hd + dw;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    cb[k] = ws.capped[k] ? 1.0 - (ws.gv[2 * k] * ws.pos[1][2 * k] + ws.gv[2 * k + 1] * ws.pos[1][2 * k + 1]) : 0.0;
    a -= ws.bd[2 * k] * ws.pos[1][2 * k] + ws.bd[2 * k + 1] * ws.pos[1][2 * k + 1];
  }
  // Cholesky 4x4 (packed lower), inverse diagonals stored
  bool ok = true;
  double L[10];
  {
    double d0 = T[0]; ok = ok && d0 > 0.0 && d0 < 1e300; double i0 = rsq(d0);
    double l10 = T[1] * i0, l20 = T[3] * i0, l30 = T[6] * i0;
    double d1 = T[2] - l10 * l10; ok = ok && d1 > 0.0; double i1 = rsq(d1);
    double l21 = (T[4] - l20 * l10) * i1, l31 = (T[7] - l30 * l10) * i1;
    double d2 = T[5] - l20 * l20 - l21 * l21; ok = ok && d2 > 0.0; double i2 = rsq(d2);
    double l32 = (T[8] - l30 * l20 - l31 * l21) * i2;
    double d3 = T[9] - l30 * l30 - l31 * l31 - l32 * l32; ok = ok && d3 > 0.0; double i3 = rsq(d3);
    L[0] = i0; L[1] = l10; L[2] = i1; L[3] = l20; L[4] = l21; L[5] = i2; L[6] = l30; L[7] = l31; L[8] = l32; L[9] = i3;
  }
  if (!ok) return false;
  // Tic = T^-1 cb
  double y0 = cb[0] * L[0], y1 = (cb[1] - L[1] * y0) * L[2], y2 = (cb[2] - L[3] * y0 - L[4] * y1) * L[5],
         y3 = (cb[3] - L[6] * y0 - L[7] * y1 - L[8] * y2) * L[9];
  double k3 = y3 * L[9], k2 = (y2 - L[8] * k3) * L[5], k1 = (y1 - L[4] * k2 - L[7] * k3) * L[2],
         k0 = (y0 - L[1] * k1 - L[3] * k2 - L[6] * k3) * L[0];
  double den = a + cb[0] * k0 + cb[1] * k1 + cb[2] * k2 + cb[3] * k3;
  if (!(den > 0.0)) return false;
  wsync();
  if (lane == 0) {
#pragma unroll
    for (int t = 0; t < 10; t++) ws.T[t] = L[t];
    ws.Tic[0] = k0; ws.Tic[1] = k1; ws.Tic[2] = k2; ws.Tic[3] = k3; ws.den = den;
#pragma unroll
    for (int k = 0; k < 4; k++) ws.cb[k] = cb[k];
  }
  wsync();
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// Solve  [H A'; A 0][dx; dy] = [r1; r2]  (r1 = ws.r1 per variable, r2 = ws.r2 per equality row) with the current
// factorisation; dx, dy -> ws.dx, ws.dy; border multipliers t -> ws.tb.  ws.r1's X entries are modified.
// ---------------------------------------------------------------------------------------------------------
template <int C>
DEVNI void kkt_solve(WS<C>& ws, const KParams& P, double dw, bool border) {
  typedef Dm<C> d;
  constexpr int NW = d::NW;
  const int lane = lane_id();
  // stage elimination of (u_j, lambda_j): one lane per stage
  if (lane < 4) {
    const int j = lane;
    const double e0 = ws.r2[24 + j];
    double d0[2];
#pragma unroll
    for (int a = 0; a < 2; a++) d0[a] = ws.F[4 * a + C - 1] * e0 + ws.r2[16 + 2 * j + a];
    const double last = ws.r1[24 + C * j + C - 1] - (ws.hl[C * j + C - 1] + dw) * e0;
#pragma unroll
    for (int a = 0; a < NW; a++) {
      double s = ws.r1[24 + C * j + a] - last;
#pragma unroll
      for (int k = 0; k < 2; k++) s += ws.D[k * NW + a] * (ws.r1[16 + 2 * j + k] - (ws.hu[2 * j + k] + dw) * d0[k]);
      ws.rw[NW * j + a] = s;
    }
#pragma unroll
    for (int b = 0; b < 2; b++) {
      ws.cv[4 * j + 2 * b] = P.bp * d0[b] - ws.r2[4 * j + 2 * b];
      ws.cv[4 * j + 2 * b + 1] = P.bv * d0[b] - ws.r2[4 * j + 2 * b + 1];
      if (j >= 1) ws.r1[4 * (j - 1) + 2 * b] -= ws.hxu[2 * j + b] * d0[b];
    }
  }
  wsync();
  kkt_backward<C>(ws, P, 0x1u);
  double ddl;
  if (border) {
    kkt_forward<C>(ws, P, 0x1u, 0);
    double rt[4], rd = ws.r1[d::XD];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double pk0 = ws.pos[0][2 * k], pk1 = ws.pos[0][2 * k + 1];
      rt[k] = ws.capped[k] ? ws.gv[2 * k] * pk0 + ws.gv[2 * k + 1] * pk1 : 0.0;
      rd -= ws.bd[2 * k] * pk0 + ws.bd[2 * k + 1] * pk1;
    }
    const double* L = ws.T;
    double y0 = rt[0] * L[0], y1 = (rt[1] - L[1] * y0) * L[2], y2 = (rt[2] - L[3] * y0 - L[4] * y1) * L[5],
           y3 = (rt[3] - L[6] * y0 - L[7] * y1 - L[8] * y2) * L[9];
    double k3 = y3 * L[9], k2 = (y2 - L[8] * k3) * L[5], k1 = (y1 - L[4] * k2 - L[7] * k3) * L[2],
           k0 = (y0 - L[1] * k1 - L[3] * k2 - L[6] * k3) * L[0];
    ddl = (rd - (ws.cb[0] * k0 + ws.cb[1] * k1 + ws.cb[2] * k2 + ws.cb[3] * k3)) / ws.den;
    double t[4] = {ws.Tic[0] * ddl + k0, ws.Tic[1] * ddl + k1, ws.Tic[2] * ddl + k2, ws.Tic[3] * ddl + k3};
    wsync();
    // combine the backward-pass vectors: p <- p0 - ddl p_delta - sum t_k p_k ; same for k
    if (lane < 16) {
      double s = ws.u.s.pv0[lane] - ddl * ws.pvb[0][lane];
#pragma unroll
      for (int k = 0; k < 4; k++) if (ws.capped[k]) s -= t[k] * ws.pvb[1 + k][lane];
      ws.u.s.pv0[lane] = s;
    } else if (lane < 16 + 4 * NW) {
      int q = lane - 16;
      double s = ws.u.s.kv0[q] - ddl * ws.kvb[0][q];
#pragma unroll
      for (int k = 0; k < 4; k++) if (ws.capped[k]) s -= t[k] * ws.kvb[1 + k][q];
      ws.u.s.kv0[q] = s;
    } else if (lane == 31) {
      ws.tb[0] = t[0]; ws.tb[1] = t[1]; ws.tb[2] = t[2]; ws.tb[3] = t[3];
    }
    wsync();
  } else {
    ddl = ws.r1[d::XD] / (ws.hd + dw);
  }
  kkt_forward<C>(ws, P, 0x1u, 1);
  // recover du, dlambda, dnu, deta : one lane per stage
  if (lane < 4) {
    const int j = lane;
    const double e0 = ws.r2[24 + j];
    double wsum_ = 0.0, du[2] = {ws.F[C - 1] * e0 + ws.r2[16 + 2 * j], ws.F[4 + C - 1] * e0 + ws.r2[17 + 2 * j]};
#pragma unroll
    for (int a = 0; a < NW; a++) {
      double w = ws.wsol[NW * j + a];
      ws.dx[24 + C * j + a] = w;
      wsum_ += w;
      du[0] += ws.D[a] * w; du[1] += ws.D[NW + a] * w;
    }
    const double dll = e0 - wsum_;
    ws.dx[24 + C * j + C - 1] = dll;
    double dnu[2];
#pragma unroll
    for (int a = 0; a < 2; a++) {
      ws.dx[16 + 2 * j + a] = du[a];
      double s = ws.r1[16 + 2 * j + a] - (ws.hu[2 * j + a] + dw) * du[a] - (P.bp * ws.dy[4 * j + 2 * a] + P.bv * ws.dy[4 * j + 2 * a + 1]);
      if (j >= 1) s -= ws.hxu[2 * j + a] * ws.dx[4 * (j - 1) + 2 * a];
      dnu[a] = s;
      ws.dy[16 + 2 * j + a] = s;
    }
    ws.dy[24 + j] = ws.r1[24 + C * j + C - 1] - (ws.hl[C * j + C - 1] + dw) * dll + ws.F[C - 1] * dnu[0] + ws.F[4 + C - 1] * dnu[1];
  }
  if (lane == 4) { ws.dx[d::XD] = ddl; ws.ddl = ddl; }
  wsync();
}

DEV double wmax_any(double v) {  // signed max via butterfly
  for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor(v, m));
  return v;
}
DEV double wmin_any(double v) {
  for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor(v, m));
  return v;
}

// ---------------------------------------------------------------------------------------------------------
// Per-lane slot state
// ---------------------------------------------------------------------------------------------------------
struct LaneSlots {
  SlotMeta m[2];
  double zlo[2], zhi[2], slo[2], shi[2];
};

struct SolveInfo { int status, qp_iters, qp_flag, nlp_iters, n_fact; double kkt; };
struct QpState { double sigma; int it; };
struct NlpState { double mu, tau, dw_last, theta_max, theta_min, E0; int it, nfilt, status; bool first_ic; };

// =========================================================================================================
// Stage 1: QP, iSWIFT restated (Prime.c:127-230; kkt_initialize Auxilary.c:680-755)
// =========================================================================================================
// qp_init: initial point.  Returns true when the stage is already finished (factorisation failure).
template <int C>
DEV bool qp_init(WS<C>& ws, const KParams& P, LaneSlots& ls, SolveInfo& info, QpState& st) {
  typedef Dm<C> d;
  const int lane = lane_id();
  const int v0 = lane, v1 = lane + 32;
  const bool hv1 = v1 < d::NV;
  const int mrows = d::MQP;
  double* sa = ws.u.a.sa; double* sb = ws.u.a.sb;
  // ---- initial point: W = I, rhs [-c; b; h]
#pragma unroll
  for (int q = 0; q < 2; q++) {
    int s = lane + 32 * q;
    bool act = ls.m[q].qp;
    sa[s] = act ? 2.0 : 0.0;
    sb[s] = act ? (ls.m[q].hi + ls.m[q].lo) : 0.0;
  }
  if (lane < 8) ws.gv[lane] = 0.0;
  if (lane < 4) { ws.sc[lane] = 0.0; ws.capped[lane] = 0; ws.zdist[lane] = 0.0; ws.srinv[lane] = 0.0; }
  if (lane < 28) ws.y[lane] = 0.0;
  ws.x[v0] = 0.0; if (hv1) ws.x[v1] = 0.0;
  wsync();
  assemble_hessian<C>(ws, P, false, 0.0);
  {
    double c0 = (v0 < 16) ? -((v0 >= 12) ? P.Pw : P.Qw) * ws.xdes[v0] : 0.0;
    ws.r1[v0] = -c0 + gatherJT<C>(sb, ws.gv, v0);
    if (hv1) ws.r1[v1] = gatherJT<C>(sb, ws.gv, v1);
    if (lane < 28) ws.r2[lane] = -eq_residual<C>(ws, P, ws.x, lane, false);
  }
  wsync();
  info.qp_flag = 3; info.qp_iters = 0; st.it = 0; st.sigma = 100.0;  // GlobalOptions.h:26
  if (!riccati_factor<C>(ws, P, 0.0)) { info.qp_flag = 1; return true; }
  kkt_solve<C>(ws, P, 0.0, false);
  ws.x[v0] = ws.dx[v0]; if (hv1) ws.x[v1] = ws.dx[v1];
  if (lane < 28) ws.y[lane] = ws.dy[lane];
  wsync();
  {
    double rmin = 1e300, rmax = -1e300, rl[2], rh[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      double e = slot_e(ls.m[q], ws.x, 0.0, 0.0, 0.0);
      rl[q] = e - ls.m[q].lo; rh[q] = ls.m[q].hi - e;
      if (ls.m[q].qp) { rmin = fmin(rmin, fmin(rl[q], rh[q])); rmax = fmax(rmax, fmax(rl[q], rh[q])); }
    }
    rmin = wmin_any(rmin); rmax = wmax_any(rmax);
    double ap = -rmin, ad = rmax;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      bool act = ls.m[q].qp;
      ls.slo[q] = act ? ((ap < 0) ? rl[q] : rl[q] + (1.0 + ap)) : 1.0;
      ls.shi[q] = act ? ((ap < 0) ? rh[q] : rh[q] + (1.0 + ap)) : 1.0;
      ls.zlo[q] = act ? ((ad < 0) ? -rl[q] : -rl[q] + (1.0 + ad)) : 0.0;
      ls.zhi[q] = act ? ((ad < 0) ? -rh[q] : -rh[q] + (1.0 + ad)) : 0.0;
    }
  }
  info.qp_flag = 2;
  return false;
}

// qp_iter: one Mehrotra predictor-corrector iteration (Prime.c:138-218).  Returns true when the stage is finished.
template <int C>
DEV bool qp_iter(WS<C>& ws, const KParams& P, LaneSlots& ls, SolveInfo& info, QpState& st) {
  typedef Dm<C> d;
  const int lane = lane_id();
  const int v0 = lane, v1 = lane + 32;
  const bool hv1 = v1 < d::NV;
  const int mrows = d::MQP;
  double* sa = ws.u.a.sa; double* sb = ws.u.a.sb;
  double sigma = st.sigma;
  if (st.it >= P.qp_maxit) return true;
  {
    // residuals (computeresiduals, Auxilary.c:524-553)
    double rzl[2], rzh[2], acc_rz = 0.0, acc_sz = 0.0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      double e = slot_e(ls.m[q], ws.x, 0.0, 0.0, 0.0);
      bool act = ls.m[q].qp;
      rzh[q] = act ? (ls.m[q].hi - ls.shi[q] - e) : 0.0;
      rzl[q] = act ? (-ls.m[q].lo - ls.slo[q] + e) : 0.0;
      acc_rz += rzh[q] * rzh[q] + rzl[q] * rzl[q];
      acc_sz += ls.slo[q] * ls.zlo[q] + ls.shi[q] * ls.zhi[q];
      sb[lane + 32 * q] = ls.zlo[q] - ls.zhi[q];
    }
    wsync();
    double rx0 = -(grad_plus_Aty<C>(ws, P, ws.x, ws.y, v0, true) - gatherJT<C>(sb, ws.gv, v0));
    double rx1 = hv1 ? -(grad_plus_Aty<C>(ws, P, ws.x, ws.y, v1, true) - gatherJT<C>(sb, ws.gv, v1)) : 0.0;
    double ry = (lane < 28) ? -eq_residual<C>(ws, P, ws.x, lane, false) : 0.0;
    double nrx = wsum(rx0 * rx0 + rx1 * rx1), nry = wsum(ry * ry), nrz = wsum(acc_rz), sz = wsum(acc_sz);
    const double lim = P.qp_reltol / sqrt(3.0);
    if (sqrt(nrx) < lim && sqrt(nrz) < lim && sqrt(nry) < lim && sz / mrows < P.qp_abstol) { info.qp_flag = 0; return true; }
    const double mu = sz / mrows;
    double dsl[2], dsh[2], dzl[2], dzh[2];   // affine (then final) slack / multiplier steps
    double cds_l[2], cds_h[2];               // "ds" of form_ds (Auxilary.c:250-267)
    double isl[2], ish[2], izl[2], izh[2], wl[2], wh[2];  // 1/s, 1/z, W^-1 = z/s  (one reciprocal each per iteration)
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const bool act = ls.m[q].qp;
      isl[q] = rcp(ls.slo[q]); ish[q] = rcp(ls.shi[q]);
      izl[q] = act ? rcp(ls.zlo[q]) : 0.0; izh[q] = act ? rcp(ls.zhi[q]) : 0.0;
      wl[q] = act ? ls.zlo[q] * isl[q] : 0.0; wh[q] = act ? ls.zhi[q] * ish[q] : 0.0;
    }
    wsync();
    if (sigma > 0.0) {
      // predictor: W = s/z, ds = -s z, rhs3 = rz - ds/z = rz + s
#pragma unroll
      for (int q = 0; q < 2; q++) {
        sa[lane + 32 * q] = wl[q] + wh[q];
        sb[lane + 32 * q] = wh[q] * (rzh[q] + ls.shi[q]) - wl[q] * (rzl[q] + ls.slo[q]);
      }
      wsync();
      assemble_hessian<C>(ws, P, false, 0.0);
      ws.r1[v0] = rx0 + gatherJT<C>(sb, ws.gv, v0);
      if (hv1) ws.r1[v1] = rx1 + gatherJT<C>(sb, ws.gv, v1);
      if (lane < 28) ws.r2[lane] = ry;
      wsync();
      info.n_fact++;
      if (!riccati_factor<C>(ws, P, 0.0)) { info.qp_flag = 1; return true; }
      kkt_solve<C>(ws, P, 0.0, false);
      double rp = 0.0, rdm = 0.0;   // max(-ds/s), max(-dz/z): alpha^max = 1/that (findsteplength, Auxilary.c:271-294)
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const bool act = ls.m[q].qp;
        double de = slot_de(ls.m[q], ws.dx, 0.0, 0.0, 0.0);
        dzh[q] = wh[q] * (de - (rzh[q] + ls.shi[q]));
        dzl[q] = wl[q] * (-de - (rzl[q] + ls.slo[q]));
        dsh[q] = act ? (-ls.shi[q] * ls.zhi[q] - ls.shi[q] * dzh[q]) * izh[q] : 0.0;
        dsl[q] = act ? (-ls.slo[q] * ls.zlo[q] - ls.slo[q] * dzl[q]) * izl[q] : 0.0;
        rp = fmax(rp, fmax(-dsh[q] * ish[q], -dsl[q] * isl[q]));
        rdm = fmax(rdm, fmax(-dzh[q] * izh[q], -dzl[q] * izl[q]));
      }
      rp = wmax_nn(rp); rdm = wmax_nn(rdm);
      const double ap = (rp > 0.0) ? 1.0 / rp : 1.0, ad = (rdm > 0.0) ? 1.0 / rdm : 1.0;
      double num = 0.0;
#pragma unroll
      for (int q = 0; q < 2; q++)
        if (ls.m[q].qp)
          num += (ls.slo[q] + ap * dsl[q]) * (ls.zlo[q] + ad * dzl[q]) + (ls.shi[q] + ap * dsh[q]) * (ls.zhi[q] + ad * dzh[q]);
      num = wsum(num);
      double rho = num / sz, mr = rho < 1.0 ? rho : 1.0;
      sigma = mr * mr * mr;
      if (sigma < 0.0) sigma = 0.0;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        cds_l[q] = -ls.slo[q] * ls.zlo[q] - dsl[q] * dzl[q] + sigma * mu;
        cds_h[q] = -ls.shi[q] * ls.zhi[q] - dsh[q] * dzh[q] + sigma * mu;
      }
    } else {
      sigma = 0.0;
#pragma unroll
      for (int q = 0; q < 2; q++) { cds_l[q] = -ls.slo[q] * ls.zlo[q]; cds_h[q] = -ls.shi[q] * ls.zhi[q]; }
    }
    // corrector solve with the same factor
    double r3l[2], r3h[2];
    wsync();
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const bool act = ls.m[q].qp;
      r3l[q] = act ? rzl[q] - cds_l[q] * izl[q] : 0.0;
      r3h[q] = act ? rzh[q] - cds_h[q] * izh[q] : 0.0;
      sb[lane + 32 * q] = wh[q] * r3h[q] - wl[q] * r3l[q];
    }
    wsync();
    ws.r1[v0] = rx0 + gatherJT<C>(sb, ws.gv, v0);
    if (hv1) ws.r1[v1] = rx1 + gatherJT<C>(sb, ws.gv, v1);
    if (lane < 28) ws.r2[lane] = ry;
    wsync();
    kkt_solve<C>(ws, P, 0.0, false);
    double rp = 0.0, rdm = 0.0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const bool act = ls.m[q].qp;
      double de = slot_de(ls.m[q], ws.dx, 0.0, 0.0, 0.0);
      dzh[q] = wh[q] * (de - r3h[q]);
      dzl[q] = wl[q] * (-de - r3l[q]);
      dsh[q] = act ? (cds_h[q] - ls.shi[q] * dzh[q]) * izh[q] : 0.0;
      dsl[q] = act ? (cds_l[q] - ls.slo[q] * dzl[q]) * izl[q] : 0.0;
      rp = fmax(rp, fmax(-dsh[q] * ish[q], -dsl[q] * isl[q]));
      rdm = fmax(rdm, fmax(-dzh[q] * izh[q], -dzl[q] * izl[q]));
    }
    rp = wmax_nn(rp); rdm = wmax_nn(rdm);
    double ap = (rp > 0.0) ? 1.0 / rp : 1.0, ad = (rdm > 0.0) ? 1.0 / rdm : 1.0;
    ap = fmin(0.99 * ap, 1.0); ad = fmin(0.99 * ad, 1.0);
    ws.x[v0] += ap * ws.dx[v0]; if (hv1) ws.x[v1] += ap * ws.dx[v1];
    if (lane < 28) ws.y[lane] += ad * ws.dy[lane];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      ls.slo[q] += ap * dsl[q]; ls.shi[q] += ap * dsh[q];
      ls.zlo[q] += ad * dzl[q]; ls.zhi[q] += ad * dzh[q];
    }
    wsync();
  }
  st.sigma = sigma;
  st.it++;
  info.qp_iters = st.it;
  return false;
}

// =========================================================================================================
// Stage 2: NLP, IPM-F
// =========================================================================================================
template <int C>
DEV void nlp_init(WS<C>& ws, const KParams& P, LaneSlots& ls, SolveInfo& info, NlpState& st) {
  typedef Dm<C> d;
  const int lane = lane_id();
  const int v0 = lane, v1 = lane + 32;
  const bool hv1 = v1 < d::NV;
  double* sa = ws.u.a.sa; double* sb = ws.u.a.sb;
  const double ox = ws.obs[0], oy = ws.obs[1];
  // ---- step 0: strictly interior start (X, then U against friction, then lambda, then the slack variable)
#pragma unroll
  for (int q = 0; q < 2; q++) {
    int v = lane + 32 * q;
    if (v < 16) {
      double lim = (v & 1) ? fmin(P.box, P.vmax) : P.box;
      ws.x[v] = fmin(fmax(ws.x[v], -lim + P.push), lim - P.push);
    } else if (v >= 24 && v < d::XD) ws.x[v] = fmin(fmax(ws.x[v], P.push), 1.0 - P.push);
  }
  wsync();
  if (lane >= 16 && lane < 24) {
    int q = lane - 16, k = q >> 1, a = q & 1;
    double lo = -P.box + P.push, hi = P.box - P.push;
    if (k >= 1) { double pk = ws.x[4 * (k - 1) + 2 * a]; lo = fmax(lo, pk - P.fb + P.push); hi = fmin(hi, pk + P.fb - P.push); }
    ws.x[lane] = fmin(fmax(ws.x[lane], lo), hi);
  }
  {
    double cmin = 1e300;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double a = ws.x[4 * k] - ox, b = ws.x[4 * k + 2] - oy;
      cmin = fmin(cmin, a * a + b * b + ws.x[d::XD] - P.eps);
    }
    wsync();
    if (lane == 0 && cmin < P.push) ws.x[d::XD] += P.push - cmin;
  }
  if (lane < 28) ws.y[lane] = 0.0;
  if (lane < 16) ws.filt[lane] = 0.0;
  wsync();
  const double mu = P.mu_init;
  st.mu = mu; st.tau = fmax(P.tau_min, 1.0 - mu); st.dw_last = 0.0; st.first_ic = true; st.nfilt = 0; st.it = 0; st.status = 1; st.E0 = 1e300;
#pragma unroll
  for (int q = 0; q < 2; q++) {
    double e = slot_e(ls.m[q], ws.x, ox, oy, ws.x[d::XD]);
    ls.slo[q] = ls.m[q].vlo ? e - ls.m[q].lo : 1.0;
    ls.shi[q] = ls.m[q].vhi ? ls.m[q].hi - e : 1.0;
    ls.zlo[q] = ls.m[q].vlo ? mu * rcp(ls.slo[q]) : 0.0;
    ls.zhi[q] = ls.m[q].vhi ? mu * rcp(ls.shi[q]) : 0.0;
  }
  {
    double r = (lane < 28) ? fabs(eq_residual<C>(ws, P, ws.x, lane, false)) : 0.0;
    double th0 = wsum(r);
    st.theta_max = 1e4 * fmax(1.0, th0); st.theta_min = 1e-4 * fmax(1.0, th0);
  }
  info.status = 1; info.nlp_iters = 0;
}

// nlp_iter: one IPM-F iteration.  Returns true when the stage is finished (info.status set).
template <int C>
DEV bool nlp_iter(WS<C>& ws, const KParams& P, LaneSlots& ls, SolveInfo& info, NlpState& st) {
  typedef Dm<C> d;
  const int lane = lane_id();
  const int v0 = lane, v1 = lane + 32;
  const bool hv1 = v1 < d::NV;
  double* sa = ws.u.a.sa; double* sb = ws.u.a.sb;
  const double ox = ws.obs[0], oy = ws.obs[1];
  double mu = st.mu, tau = st.tau, dw_last = st.dw_last, E0;
  const double theta_max = st.theta_max, theta_min = st.theta_min;
  bool first_ic = st.first_ic;
  int nfilt = st.nfilt;
  const int it = st.it;
  int status = 1;
  {
    // ---- slots at x: s = c(x) (feasible IPM), multiplier safeguard, gradients of the distance rows
    double gx[2] = {0, 0}, gy[2] = {0, 0}, e_c = 0.0, zmax = 0.0, smin = 1e300;
    const double xdl = ws.x[d::XD];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const SlotMeta& m = ls.m[q];
      double e = slot_e(m, ws.x, ox, oy, xdl);
      ls.slo[q] = m.vlo ? e - m.lo : 1.0;
      ls.shi[q] = m.vhi ? m.hi - e : 1.0;
      smin = fmin(smin, fmin(ls.slo[q], ls.shi[q]));
      if (it > 0) {  // z reset (Ipopt eq. 16) with the mu of the step just taken
        const double mr_lo = mu * rcp(ls.slo[q]), mr_hi = mu * rcp(ls.shi[q]), iks = 1.0 / P.kappa_sigma;
        if (m.vlo) ls.zlo[q] = fmin(fmax(ls.zlo[q], mr_lo * iks), P.kappa_sigma * mr_lo);
        if (m.vhi) ls.zhi[q] = fmin(fmax(ls.zhi[q], mr_hi * iks), P.kappa_sigma * mr_hi);
      }
      if (m.kind == 2) {
        gx[q] = 2.0 * (ws.x[m.i1] - ox); gy[q] = 2.0 * (ws.x[m.i2] - oy);
        int k = (lane + 32 * q) - d::S_DIST;
        ws.gv[2 * k] = gx[q]; ws.gv[2 * k + 1] = gy[q];
      }
      e_c = fmax(e_c, fmax(fabs(ls.slo[q] * ls.zlo[q]), fabs(ls.shi[q] * ls.zhi[q])));
      zmax = fmax(zmax, fmax(ls.zlo[q], ls.zhi[q]));
      sb[lane + 32 * q] = ls.zlo[q] - ls.zhi[q];
    }
    wsync();
    // ---- KKT error
    const double gfy0 = grad_plus_Aty<C>(ws, P, ws.x, ws.y, v0, true);
    const double gfy1 = hv1 ? grad_plus_Aty<C>(ws, P, ws.x, ws.y, v1, true) : 0.0;
    const double rd0 = gfy0 - gatherJT<C>(sb, ws.gv, v0);
    const double rd1 = hv1 ? gfy1 - gatherJT<C>(sb, ws.gv, v1) : 0.0;
    const double re = (lane < 28) ? eq_residual<C>(ws, P, ws.x, lane, false) : 0.0;
    const double e_d = wmax_nn(fmax(fabs(rd0), fabs(rd1)));
    const double e_p = wmax_nn(fabs(re));
    e_c = wmax_nn(e_c);
    zmax = wmax_nn(zmax);
    const double th = wsum(fabs(re));
    E0 = fmax(e_d, fmax(e_p, e_c));
    info.kkt = E0; info.nlp_iters = it;
    CBF_DBG("it %d mu %.2e E0 %.3e ed %.2e ep %.2e ec %.2e zmax %.2e th %.2e\n", it, mu, E0, e_d, e_p, e_c, zmax, th);
    if (E0 <= P.tol) { info.status = 0; return true; }
    if (it == P.max_iter) { info.status = 1; return true; }
    if (!(zmax < P.z_div) || !(E0 < 1e300) || wany(!(smin > 0.0))) { info.status = 2; return true; }
    // ---- barrier update (Ipopt eq. 7)
    for (;;) {
      double emu = 0.0;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        if (ls.m[q].vlo) emu = fmax(emu, fabs(ls.slo[q] * ls.zlo[q] - mu));
        if (ls.m[q].vhi) emu = fmax(emu, fabs(ls.shi[q] * ls.zhi[q] - mu));
      }
      emu = fmax(wmax_nn(emu), fmax(e_d, e_p));
      if (emu <= P.kappa_eps * mu && mu > P.tol / 10.0) {
        mu = fmax(P.tol / 10.0, fmin(P.kappa_mu * mu, mu * sqrt(mu)));
        tau = fmax(P.tau_min, 1.0 - mu);
        nfilt = 0;
      } else break;
    }
    // ---- slot weights: Sigma -> sa, mu/s_lo - mu/s_hi -> sb ; distance rows: capped split
    double rlo[2], rhi[2];
    wsync();
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const SlotMeta& m = ls.m[q];
      rlo[q] = rcp(ls.slo[q]); rhi[q] = rcp(ls.shi[q]);
      double sg = ls.zlo[q] * rlo[q] + ls.zhi[q] * rhi[q];
      sa[lane + 32 * q] = sg;
      sb[lane + 32 * q] = (m.vlo ? mu * rlo[q] : 0.0) - (m.vhi ? mu * rhi[q] : 0.0);
      if (m.kind == 2) {
        int k = (lane + 32 * q) - d::S_DIST;
        double scap = fmin(sg, P.sigma_cap);
        bool cap = sg > P.sigma_cap;
        ws.sc[k] = scap; ws.capped[k] = cap ? 1 : 0; ws.srinv[k] = cap ? rcp(sg - scap) : 0.0; ws.zdist[k] = ls.zlo[q];
      }
    }
    wsync();
    assemble_hessian<C>(ws, P, true, P.rho_lambda);
    ws.r1[v0] = -gfy0 + gatherJT<C>(sb, ws.gv, v0);
    if (hv1) ws.r1[v1] = -gfy1 + gatherJT<C>(sb, ws.gv, v1);
    if (lane < 28) ws.r2[lane] = -re;
    wsync();
    // ---- factorisation with inertia correction (Ipopt algorithm IC)
    double dw = 0.0;
    bool ok = false;
    for (;;) {
      info.n_fact++;
      ok = riccati_factor<C>(ws, P, dw);
      if (ok) ok = border_factor<C>(ws, P, dw);
      if (ok) break;
      if (dw == 0.0) dw = (dw_last == 0.0) ? P.dw_first : fmax(P.dw_min, P.kw_minus * dw_last);
      else dw *= (first_ic && dw_last == 0.0) ? P.kw_plus_first : P.kw_plus;
      if (dw > P.dw_max) break;
      wsync();
    }
    if (!ok) { info.status = 3; return true; }
    if (dw > 0.0) { dw_last = dw; first_ic = false; }
    kkt_solve<C>(ws, P, dw, true);
    // ---- slot steps
    double de[2], dzl[2], dzh[2], quad[2];
    const double ddl = ws.dx[d::XD];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const SlotMeta& m = ls.m[q];
      de[q] = slot_de(m, ws.dx, gx[q], gy[q], ddl);
      quad[q] = (m.kind == 2) ? ws.dx[m.i1] * ws.dx[m.i1] + ws.dx[m.i2] * ws.dx[m.i2] : 0.0;
      dzl[q] = m.vlo ? (mu - ls.slo[q] * ls.zlo[q]) * rlo[q] - ls.zlo[q] * rlo[q] * de[q] : 0.0;
      dzh[q] = m.vhi ? (mu - ls.shi[q] * ls.zhi[q]) * rhi[q] + ls.zhi[q] * rhi[q] * de[q] : 0.0;
      if (m.kind == 2) {
        int k = (lane + 32 * q) - d::S_DIST;
        if (ws.capped[k]) dzl[q] = (mu - ls.slo[q] * ls.zlo[q]) * rlo[q] - ws.sc[k] * de[q] - ws.tb[k];
      }
    }
    // ---- iterative refinement on the unreduced Newton system
    double dxa0 = ws.dx[v0], dxa1 = hv1 ? ws.dx[v1] : 0.0, dya = (lane < 28) ? ws.dy[lane] : 0.0;
    for (int rf = 0; rf < P.max_refine; rf++) {
      double rcl[2], rch[2];
      wsync();
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const SlotMeta& m = ls.m[q];
        rcl[q] = m.vlo ? (mu - ls.slo[q] * ls.zlo[q]) - (ls.slo[q] * dzl[q] + ls.zlo[q] * de[q]) : 0.0;
        rch[q] = m.vhi ? (mu - ls.shi[q] * ls.zhi[q]) - (ls.shi[q] * dzh[q] - ls.zhi[q] * de[q]) : 0.0;
        sa[lane + 32 * q] = (ls.zlo[q] + dzl[q]) - (ls.zhi[q] + dzh[q]);
        sb[lane + 32 * q] = rcl[q] * rlo[q] - rch[q] * rhi[q];
      }
      wsync();
      // rho1 = -(grad f + A'(y+dy) + Htilde dx - J'(z+dz)),  rho2 = -(A (x+dx) - b)
      double ht0, ht1 = 0.0;
      {
        int v = v0;
        if (v < 16) { ht0 = ((v >= 12) ? P.Pw : P.Qw) + dw - ((v & 1) ? 0.0 : 2.0 * ws.zdist[v >> 2]); }
        else if (v < 24) ht0 = P.Rw + dw;
        else ht0 = dw + P.rho_lambda;
        if (hv1) ht1 = (v1 < d::XD) ? dw + P.rho_lambda : P.slw + dw;
      }
      double rho0 = -(gfy0 + grad_plus_Aty<C>(ws, P, ws.dx, ws.dy, v0, false) + ht0 * dxa0 - gatherJT<C>(sa, ws.gv, v0));
      double rho1 = hv1 ? -(gfy1 + grad_plus_Aty<C>(ws, P, ws.dx, ws.dy, v1, false) + ht1 * dxa1 - gatherJT<C>(sa, ws.gv, v1)) : 0.0;
      double rho2 = (lane < 28) ? -(re + eq_residual<C>(ws, P, ws.dx, lane, true)) : 0.0;
      double nres = wmax_nn(fmax(fmax(fabs(rho0), fabs(rho1)), fabs(rho2)));
      double nrhs = fmax(e_d, e_p);
      CBF_DBG("   refine %d nres %.2e nrhs %.2e\n", rf, nres, nrhs);
      if (!(nres > fmax(P.refine_rel * nrhs, P.refine_abs))) break;
      wsync();
      ws.r1[v0] = rho0 + gatherJT<C>(sb, ws.gv, v0);
      if (hv1) ws.r1[v1] = rho1 + gatherJT<C>(sb, ws.gv, v1);
      if (lane < 28) ws.r2[lane] = rho2;
      wsync();
      kkt_solve<C>(ws, P, dw, true);
      const double dd = ws.dx[d::XD];
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const SlotMeta& m = ls.m[q];
        double dde = slot_de(m, ws.dx, gx[q], gy[q], dd);
        double cl = m.vlo ? rcl[q] * rlo[q] - ls.zlo[q] * rlo[q] * dde : 0.0;
        double ch = m.vhi ? rch[q] * rhi[q] + ls.zhi[q] * rhi[q] * dde : 0.0;
        if (m.kind == 2) {
          int k = (lane + 32 * q) - d::S_DIST;
          if (ws.capped[k]) cl = rcl[q] * rlo[q] - ws.sc[k] * dde - ws.tb[k];
        }
        dzl[q] += cl; dzh[q] += ch; de[q] += dde;
      }
      dxa0 += ws.dx[v0]; if (hv1) dxa1 += ws.dx[v1];
      if (lane < 28) dya += ws.dy[lane];
      wsync();
      ws.dx[v0] = dxa0; if (hv1) ws.dx[v1] = dxa1;
      if (lane < 28) ws.dy[lane] = dya;
      wsync();
#pragma unroll
      for (int q = 0; q < 2; q++)
        quad[q] = (ls.m[q].kind == 2) ? ws.dx[ls.m[q].i1] * ws.dx[ls.m[q].i1] + ws.dx[ls.m[q].i2] * ws.dx[ls.m[q].i2] : 0.0;
    }
    // ---- fraction to the boundary
    // alpha_max = min(1, tau / max_i(-ds_i / s_i)): one reciprocal per lane instead of one division per row
    double rp = 0.0, rdm = 0.0;
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const SlotMeta& m = ls.m[q];
      if (m.vlo) rp = fmax(rp, -de[q] * rlo[q]);
      if (m.vhi) rp = fmax(rp, de[q] * rhi[q]);
      double zq = fmax(m.vlo ? -dzl[q] * rcp(ls.zlo[q]) : 0.0, m.vhi ? -dzh[q] * rcp(ls.zhi[q]) : 0.0);
      rdm = fmax(rdm, zq);
    }
    rp = wmax_nn(rp); rdm = wmax_nn(rdm);
    double a_pmax = (rp > tau) ? tau / rp : 1.0, a_dmax = (rdm > tau) ? tau / rdm : 1.0;
    CBF_DBG("   dw %.1e apmax %.3e admax %.3e ddl %.3e\n", dw, a_pmax, a_dmax, ws.dx[d::XD]);
    // ---- filter line search on phi = f - mu sum log s  (theta = ||A x - b||_1 shrinks by (1 - alpha))
    double ph, dphi, dQd, gdx;
    {
      // f(x) = sum_j x_j (1/2 Q_jj x_j + f_j)   (DV:423-428)
      double f0 = 0.0, gd = 0.0, qd = 0.0;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        int v = lane + 32 * q;
        if (v < d::NV) {
          double xv = ws.x[v], dv = ws.dx[v], qv, fv = 0.0;
          if (v < 16) { qv = (v >= 12) ? P.Pw : P.Qw; fv = -qv * ws.xdes[v]; }
          else if (v < 24) qv = P.Rw;
          else if (v < d::XD) qv = 0.0;
          else qv = P.slw;
          f0 += xv * (0.5 * qv * xv + fv);
          gd += (qv * xv + fv) * dv;
          qd += qv * dv * dv;
        }
      }
      double prod = 1.0, dsum = 0.0;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        prod *= ls.slo[q] * ls.shi[q];
        dsum += (ls.m[q].vlo ? de[q] * rlo[q] : 0.0) - (ls.m[q].vhi ? de[q] * rhi[q] : 0.0);
      }
      ph = wsum(f0 - mu * log(prod));
      gdx = wsum(gd);
      dphi = gdx - mu * wsum(dsum);
      dQd = wsum(qd);
    }
    double alpha = a_pmax;
    bool accepted = false, ftype = false;
    while (alpha >= P.alpha_min) {
      // phi(x + alpha dx) - phi(x) = alpha g'dx + 1/2 alpha^2 dx'Q dx - mu sum log(s_t / s)
      double prod = 1.0;
      bool pos = true;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const SlotMeta& m = ls.m[q];
        double et = alpha * de[q] + alpha * alpha * quad[q];
        double sl = m.vlo ? ls.slo[q] + et : 1.0, sh = m.vhi ? ls.shi[q] - et : 1.0;
        pos = pos && (sl > 0.0) && (sh > 0.0);
        prod *= (sl * rlo[q]) * (sh * rhi[q]);
      }
      bool allpos = wall(pos);
      double dlog = wsum(log(allpos ? prod : 1.0));
      if (allpos) {
        double tht = (1.0 - alpha) * th;
        double dph = alpha * gdx + 0.5 * alpha * alpha * dQd - mu * dlog;
        double pht = ph + dph;
        bool okf = tht <= theta_max;
        for (int f = 0; f < nfilt && okf; f++)
          if (tht >= ws.filt[2 * f] && pht >= ws.filt[2 * f + 1]) okf = false;
        if (okf) {
          bool sw = dphi < 0.0 && alpha * (dphi * dphi) > th;  // switching condition, s_phi = 2, s_theta = 1, delta = 1
          if (th <= theta_min && sw) {
            if (dph - 10.0 * CBF_EPSM * fabs(ph) <= P.eta_phi * alpha * dphi) { accepted = true; ftype = true; }
          } else if (tht <= (1.0 - P.gamma_theta) * th || dph - 10.0 * CBF_EPSM * fabs(ph) <= -P.gamma_phi * th) {
            accepted = true; ftype = false;
          }
        }
      }
      if (accepted) break;
      alpha *= 0.5;
    }
    CBF_DBG("   alpha %.3e acc %d ftype %d ph %.10e dphi %.3e\n", alpha, (int)accepted, (int)ftype, ph, dphi);
    if (!accepted) { info.status = 2; return true; }
    wsync();
    if (!ftype) {
      if (nfilt >= P.filter_cap) {
        double t = (lane < 2 * (P.filter_cap - 1)) ? ws.filt[lane + 2] : 0.0;
        wsync();
        if (lane < 2 * (P.filter_cap - 1)) ws.filt[lane] = t;
        nfilt--;
        wsync();
      }
      if (lane == 0) { ws.filt[2 * nfilt] = (1.0 - P.gamma_theta) * th; ws.filt[2 * nfilt + 1] = ph - P.gamma_phi * th; }
      nfilt++;
    }
    // ---- update
    ws.x[v0] += alpha * ws.dx[v0]; if (hv1) ws.x[v1] += alpha * ws.dx[v1];
    if (lane < 28) ws.y[lane] += alpha * ws.dy[lane];
#pragma unroll
    for (int q = 0; q < 2; q++) { ls.zlo[q] += a_dmax * dzl[q]; ls.zhi[q] += a_dmax * dzh[q]; }
    wsync();
  }
  (void)status;
  st.mu = mu; st.tau = tau; st.dw_last = dw_last; st.first_ic = first_ic; st.nfilt = nfilt; st.it = it + 1;
  return false;
}
