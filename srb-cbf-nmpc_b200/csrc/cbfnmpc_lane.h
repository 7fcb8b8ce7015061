// cbfnmpc_lane.h -- lane-per-stage CBF-NMPC solver: one instance = 4 adjacent lanes (lane j owns horizon stage j:
// x_{j+1}, u_j, lambda_j, their multipliers and their constraint rows), 8 instances per warp, FP64.
//
// Replaces, for a whole batch at once, what MPC_dist::run_NMPC() does per agent per gait domain
// (/root/reference/src/MPC_dist.cpp:81-454):
//   stage 1 (QP)  : iSWIFT's Mehrotra predictor-corrector (optimization/iSWIFT/src/Prime.c:127-230) on the LIP
//                   tracking QP of MPC_dist.cpp:164-321;
//   stage 2 (NLP) : the same QP plus the 4 squared-distance rows and 8 velocity boxes of
//                   include/dec_vars_constr_cost.h:245-324, solved by the primal-dual filter interior-point method
//                   "IPM-F" of DESIGN.md (feasible w.r.t. inequalities, exact Hessian, inertia correction);
//   epilogue      : Bezier fit of MPC_dist.cpp:784-855.
// KKT systems: per-stage elimination of (u_j, lambda_j) through u = F lambda, 1'lambda = 1; Riccati/Cholesky
// recursion over the horizon as a chain lane 3 -> 2 -> 1 -> 0 (each lane keeps its stage's factors in registers) on the
// 5-state (x, delta): the slack variable of the distance rows rides along as a fifth state, a distance row whose Sigma
// exceeds the cap keeps its remainder inside its stage in quasi-definite form.
//
// The iterate, the Hessian blocks and the Riccati factors live in registers (compile-time indices only); per-row state
// (slacks, multipliers, reciprocals, steps) and a few vectors that would otherwise be spilled live in thread-private
// columns of shared memory (element f of thread t at priv[f * LANE_NT + t]: conflict-free).  Formulas that are the same for
// every row run as rolled loops over those columns: the kernel is sensitive to its instruction footprint.
#pragma once
#include "cbfnmpc_types.h"
#include "lane_prims.h"

#ifndef LANE_NT
#define LANE_NT 64  // threads per CTA = stride of the thread-private shared-memory columns (set by cbfnmpc_api.cu)
#endif
#define LUNR _Pragma("unroll")

// ---------------------------------------------------------------------------------------------------------------
// KKT vector restricted to one lane: primal part (xs = x_{j+1}, us = u_j, lm = lambda_j, dl = slack variable,
// replicated in the 4 lanes) and dual part (pi: dynamics rows of stage j, nu: u = F lambda rows, eta: 1'lambda = 1).
// ---------------------------------------------------------------------------------------------------------------
struct KV { double xs[4], us[2], lm[4], dl, pi[4], nu[2], eta; };

struct LaneConst {       // per-instance constants of this lane
  // (the desired x_{j+1}, MPC_dist.cpp:702-782, and the measured state x_0, :226-229, live in the private columns F_XDES / F_XIN:
  //  they are read two or three times per iteration and would otherwise be spilled)
  double Fl[2];          // last stance foot (column C-1 of the stance matrix F)  (MPC_dist.cpp:1238-1252); F[a][i] = D[a][i] + Fl[a]
  double D[6];           // D = F E, D[c * NW + a] = F[c][a] - F[c][C-1]
  double ox, oy;         // closest obstacle                                   (MPC_dist.cpp:371-396)
  double qw;             // state weight of this stage (Pw at the last stage)
};

struct LaneHess {        // stage blocks of H = W + dw I + J' Sigma J  (dw is added inside the factorisation)
  double hx[4], hxy;     // x_{j+1}: diagonal, (p_x, p_y) coupling
  double hu[2], hxu[2];  // u_j diagonal; coupling x_j position <-> u_j (friction rows)
  double hl[4];          // lambda_j diagonal
  double bd[2], hd;      // coupling p_{j+1} <-> this stage's copy of the slack variable; its diagonal (the row's Sigma; + P.slw at stage 3)
  double gv[2];          // gradient of the distance row wrt p_{j+1}: 2 (p - o)
  double sc, srinv;      // capped Sigma of the distance row, 1 / (Sigma - sc) when capped
  int capped;            // the remainder of the row is kept in quasi-definite form (l_factor)
};

struct LaneFact {        // Riccati factors of stage j on the 5-state xi = (x, delta)
  double Pn[15];         // cost-to-go at xi_{j+1}, packed lower (i >= k: i (i + 1) / 2 + k)
  double K[15];          // K_j = R~^-1 S~_j: [a * 5 + m]  (S~_j itself is not kept: S~'R~^-1 v = K'v in the substitutions)
  double Lc[6];          // chol(R~_j): (1/l00, l10, 1/l11, l20, l21, 1/l22)
  double Zm[3], Nt[5], Ym[3], Tc;  // quasi-definite row: R~^-1 M', N~ = N - M R~^-1 S~, L^-1 M', 1 / sqrt(T)
  double p0dd;           // cost-to-go curvature of delta_0 (lane 0's value, replicated)
  int nfix;              // R~_j was factorised in the basis T = [e0, e1, w_n] (l_null_dir); K, Zm, Ym and the solves follow it
};

LDEV int sy(int i, int k) { return i >= k ? (i * (i + 1)) / 2 + k : (k * (k + 1)) / 2 + i; }
// Four stance feet: the support weights have ONE direction that moves nothing else -- F n = 0, 1'n = 0 -- and in the control
// basis of the stage elimination (lambda = lambda_p + E w, D = F E) it is the null vector of D: w_n = (n0, n1, 1) / scale = the
// cross product of D's two rows.  Its curvature is only what the bounds of the support weights give (Sigma_lambda ~ mu / s^2 ~
// 1e-5 late in a solve) while the rest of R~ = D' N2 D + ... carries the stiffest state bound of the stage (Sigma ~ 1e13 for an
// active velocity bound): formed entry by entry, the rounding of the large part (1e13 x 2e-16) buries the small one, the Newton
// step along n is noise and the support-weight rows of the dual residual hover at ~mu / s (measured: one instance of the
// 8 x 4096 bench set needed 21 instead of 12 iterations and set the time of its whole launch).  The factorisation therefore
// works in the basis T = [e0, e1, w_n]: row and column 2 of T'R~T have no D term at all (D w_n = 0) and are written down from the
// lambda block alone; right-hand sides get their third component the same way, and w = T w' at the end.
template <int C>
LDEV bool l_null_dir(const LaneConst& c, double& n0, double& n1) {
  if (C != 4) { n0 = 0.0; n1 = 0.0; return false; }
  const double c0 = c.D[1] * c.D[5] - c.D[2] * c.D[4], c1 = c.D[2] * c.D[3] - c.D[0] * c.D[5], c2 = c.D[0] * c.D[4] - c.D[1] * c.D[3];
  const bool ok = fabs(c2) > 1e-3 * fmax(fabs(c0), fabs(c1));   // feet 0, 1 and the pivot foot not (nearly) in line
  const double r = ok ? l_rcp(c2) : 0.0;
  n0 = c0 * r; n1 = c1 * r;
  return ok;
}
LDEV double l_sel4(const double* v, int j) { return j == 0 ? v[0] : (j == 1 ? v[1] : (j == 2 ? v[2] : v[3])); }

// thread-private shared-memory column
struct Priv {
  double* base;
  LMEM double& operator()(int f) const { return base[f * LANE_NT]; }
};

template <int C>
struct LaneDims {
  static constexpr int NW = C - 1, NR = NW * (NW + 1) / 2, MQP = 60 + 8 * C;
  // two-sided slots of one lane
  // slot order: X 4 | U 2 | lambda C | then QP stage: friction 2;  NLP stage: velocity 2 | distance 1 | friction 2.
  // The friction slots come last so that lane 0 (which has none: x_0 is fixed) simply runs shorter slot loops.
  static constexpr int SX = 0, SU = 4, SL = 6, SFQ = 6 + C, NSQP = 8 + C, SV = 6 + C, SD = 8 + C, SFN = 9 + C, NS = 11 + C;
  // thread-private shared-memory fields.  QP stage: 12 per slot, stride NSQP; NLP stage: 11 per slot, stride NS (the two
  // stages never overlap); then the filter and a few parked vectors.
  static constexpr int Q_SLO = 0, Q_SHI = NSQP, Q_ZLO = 2 * NSQP, Q_ZHI = 3 * NSQP, Q_A = 4 * NSQP, Q_B = 5 * NSQP, Q_C = 6 * NSQP,
                       Q_D = 7 * NSQP, Q_E = 8 * NSQP, Q_G = 9 * NSQP, Q_EV = 10 * NSQP, Q_DE = 11 * NSQP, Q_END = 12 * NSQP;
  static constexpr int N_SLO = 0, N_SHI = NS, N_ZLO = 2 * NS, N_ZHI = 3 * NS, N_A = 4 * NS, N_B = 5 * NS, N_C = 6 * NS, N_D = 7 * NS,
                       N_E = 8 * NS, N_G = 9 * NS, N_EV = 10 * NS, N_END = 11 * NS;
  static constexpr int F_FILT = Q_END > N_END ? Q_END : N_END;  // this lane's share of the filter: entries p = 4 q + j (q = 0, 1), (theta, phi) each
  static constexpr int F_XDES = F_FILT + 4;    // desired state of this lane's stage (4)
  static constexpr int F_XIN = F_XDES + 4;     // measured state x_0 (4; used by lane 0)
  static constexpr int F_PERS = F_XIN + 4;     // values that live across iterations and are read once or twice per iteration:
                                               // tau, dw_last, theta_max, theta_min, ad_prev (out of the register file)
  static constexpr int NPRIV = F_PERS + 5;
};
constexpr int LANE_NPRIV = LaneDims<4>::NPRIV;

// ---------------------------------------------------------------------------------------------------------------
// small dense pieces
// ---------------------------------------------------------------------------------------------------------------
// Pivot rule of the 3 x 3 Cholesky.  When every lambda of a stage is strictly inside (0, 1), the reduced Hessian
// D'N D + Sigma_lambda has one direction (F n = 0, 1'n = 0) whose curvature is only Sigma_lambda ~ mu / s^2, below the
// rounding error of D'N D, so the last pivot comes out as +-1e-13 |R|.  That direction moves neither X, U nor the
// objective.  A pivot inside the noise window (> -CHOL_NOISE * diagonal) is therefore replaced by CHOL_FLOOR * diagonal
// in the QP stage (FLOOR; this only damps the motion of lambda inside its optimal face).  The NLP stage reports every
// non-positive pivot instead and lets the inertia correction add dw I, which is what keeps its line search well scaled.  iSWIFT's LDL' / the oracle's Bunch-Kaufman work on the
// unreduced KKT matrix and do not meet this cancellation (Auxilary.c:332-420).
#define LANE_QD_CHECK 1e11   // largest Sigma of a group's distance rows above which the Newton residual is checked (iterative refinement)
#define CHOL_FLOOR 1e-12
#define CHOL_NOISE 1e-9
template <bool FLOOR>
LDEV double l_pivot(double d, double diag, bool& ok) {
  ok = ok && (d > (FLOOR ? -CHOL_NOISE * diag : 0.0)) && (d < 1e300);
  return FLOOR ? fmax(d, CHOL_FLOOR * diag) : d;
}
template <int NW, bool FLOOR>
LDEV bool l_chol(const double* R, double* L) {  // R packed lower (00,10,11,20,21,22)
  bool ok = true;
  const double d0 = R[0];
  ok = ok && (d0 > 0.0) && (d0 < 1e300);
  const double i0 = l_rsq(d0);
  L[0] = i0;
  if (NW == 3) {
    const double l10 = R[1] * i0, l20 = R[3] * i0;
    const double d1 = l_pivot<FLOOR>(R[2] - l10 * l10, R[2], ok);
    const double i1 = l_rsq(d1);
    const double l21 = (R[4] - l20 * l10) * i1;
    const double d2 = l_pivot<FLOOR>(R[5] - l20 * l20 - l21 * l21, R[5], ok);
    const double i2 = l_rsq(d2);
    L[1] = l10; L[2] = i1; L[3] = l20; L[4] = l21; L[5] = i2;
  }
  return ok;
}
template <int NW>
LDEV void l_chol_fwd(const double* L, double* b) {  // b <- L^-1 b
  if (NW == 1) { b[0] = b[0] * L[0]; return; }
  const double y0 = b[0] * L[0];
  const double y1 = (b[1] - L[1] * y0) * L[2];
  const double y2 = (b[2] - L[3] * y0 - L[4] * y1) * L[5];
  b[0] = y0; b[1] = y1; b[2] = y2;
}
template <int NW>
LDEV void l_chol_bwd(const double* L, double* b) {  // b <- L^-T b
  if (NW == 1) { b[0] = b[0] * L[0]; return; }
  const double k2 = b[2] * L[5];
  const double k1 = (b[1] - L[4] * k2) * L[2];
  const double k0 = (b[0] - L[1] * k1 - L[3] * k2) * L[0];
  b[0] = k0; b[1] = k1; b[2] = k2;
}
template <int NW>
LDEV void l_chol_solve(const double* L, double* b) { l_chol_fwd<NW>(L, b); l_chol_bwd<NW>(L, b); }

// LIP maps (per axis b: position index 2b, velocity index 2b+1; MPC_dist.cpp:117-127)
LDEV void l_Ax(const KParams& P, const double* x, double* y) {        // y = A x
  LUNR for (int b = 0; b < 2; b++) {
    y[2 * b] = P.a11 * x[2 * b] + P.a12 * x[2 * b + 1];
    y[2 * b + 1] = P.a21 * x[2 * b] + P.a22 * x[2 * b + 1];
  }
}
LDEV void l_ATx(const KParams& P, const double* x, double* y) {       // y = A' x
  LUNR for (int b = 0; b < 2; b++) {
    y[2 * b] = P.a11 * x[2 * b] + P.a21 * x[2 * b + 1];
    y[2 * b + 1] = P.a12 * x[2 * b] + P.a22 * x[2 * b + 1];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Residuals that do not involve the inequality rows
// ---------------------------------------------------------------------------------------------------------------
// (Aeq x - beq) of this lane's rows; xp = x_j (previous lane's xs; lane 0: measured state, or 0 when homogeneous)
template <int C>
LDEV void l_eq_residual(const KParams& P, const LaneConst& c, const KV& v, const double* xp, bool homogeneous, double* rpi, double* rnu, double& reta) {
  double ax[4];
  l_Ax(P, xp, ax);
  LUNR for (int b = 0; b < 2; b++) {
    rpi[2 * b] = ax[2 * b] - v.xs[2 * b] + P.bp * v.us[b];
    rpi[2 * b + 1] = ax[2 * b + 1] - v.xs[2 * b + 1] + P.bv * v.us[b];
  }
  LUNR for (int a = 0; a < 2; a++) {
    double s = v.us[a];
    // F lambda = D lambda_{0..C-2} + Fl 1'lambda
    double sl = 0.0;
    LUNR for (int i = 0; i < C; i++) sl += v.lm[i];
    LUNR for (int i = 0; i < C - 1; i++) s -= c.D[a * (C - 1) + i] * v.lm[i];
    s -= c.Fl[a] * sl;
    rnu[a] = s;
  }
  double s = homogeneous ? 0.0 : -1.0;
  LUNR for (int i = 0; i < C; i++) s += v.lm[i];
  reta = s;
}

// grad f(x) + A' y restricted to this lane's variables; pin = pi of lane j+1 (0 at lane 3)
template <int C>
LDEV void l_grad_Aty(const KParams& P, const LaneConst& c, const Priv& pr, const KV& x, const KV& y, const double* pin, bool with_grad, int j, KV& g) {
  double at[4];
  l_ATx(P, pin, at);
  LUNR for (int i = 0; i < 4; i++) {
    double t = at[i] - y.pi[i];
    if (with_grad) t += c.qw * (x.xs[i] - pr(LaneDims<C>::F_XDES + i));
    g.xs[i] = t;
  }
  LUNR for (int a = 0; a < 2; a++) {
    double t = P.bp * y.pi[2 * a] + P.bv * y.pi[2 * a + 1] + y.nu[a];
    if (with_grad) t += P.Rw * x.us[a];
    g.us[a] = t;
  }
  {
    const double gl = -(c.Fl[0] * y.nu[0] + c.Fl[1] * y.nu[1]) + y.eta;   // foot C-1
    LUNR for (int i = 0; i < C - 1; i++) g.lm[i] = gl - (c.D[i] * y.nu[0] + c.D[(C - 1) + i] * y.nu[1]);
    g.lm[C - 1] = gl;
  }
  g.dl = (with_grad && j == 3) ? P.slw * x.dl : 0.0;   // this lane's share: the slack variable's cost sits at the last stage's copy
}

// shuffle helpers: 4-vector from the previous / next lane (zero at the chain ends unless `first` supplies lane 0's value)
LDEV void l_from_prev(const Grp& g, const double* v, double* out) {
  LUNR for (int i = 0; i < 4; i++) { const double t = g_up(g, v[i]); out[i] = (g.j > 0) ? t : 0.0; }
}
LDEV void l_from_next(const Grp& g, const double* v, double* out) {
  LUNR for (int i = 0; i < 4; i++) { const double t = g_down(g, v[i]); out[i] = (g.j < 3) ? t : 0.0; }
}

// ---------------------------------------------------------------------------------------------------------------
// Riccati / Cholesky factorisation as a chain over the 4 lanes, on the 5-state xi = (x, delta): the slack variable rides along
// as a fifth state (delta_{j+1} = delta_j, delta_0 free; its cost P.slw and the inertia shift dw sit at the last stage's copy),
// so a distance row is an ordinary row of its own stage and there is no border.  (Round 1 eliminated the slack variable and the
// capped part of the 4 distance rows through a 5 x 5 Schur complement whose 4 border columns went through a backward / forward
// chain of their own at every factorisation: 24 chain steps per iteration, 11 % of the stall samples, 28 private columns.  This
// form needs 12 steps; it is the recursion of the extended-mode kernel, cbfnmpc_ext.h, where it reproduces the oracle's iterates
// to 1e-14 without refinement on the cfg-3 set.)
// Every lane evaluates its stage at every step; lane j holds valid data from step 3 - j on, so after 4 steps all stages are
// factorised (no divergent code, no extra issue slots: the idle lanes of a step would execute the same instructions masked off).
// A distance row whose Sigma exceeds P.sigma_cap is condensed only up to the cap; the remainder stays inside its stage in
// quasi-definite form and is eliminated AFTER the stage's controls:  T = 1 / (Sigma - cap) + M R~^-1 M',  N~ = N - M R~^-1 S~,
// P_j += N~' T^-1 N~  (all numbers stay O(curvature) while the stage's controls have authority over the row; when they do not,
// the stiffness moves to the previous stage and the recursion loses digits there: iterative refinement in nlp_iter).
// Returns false when a pivot block of this group is not positive definite.
// ---------------------------------------------------------------------------------------------------------------
template <int C, bool FLOOR>
LDEV bool l_factor(const Grp& g, const KParams& P, const LaneConst& c, const LaneHess& H, double dw, LaneFact& f) {
  constexpr int NW = LaneDims<C>::NW;
  // own blocks of the cost on xi_{j+1}; the cost-to-go of lanes < 3 adds what lane j+1 hands down
  double Pown[15];
  LUNR for (int t = 0; t < 15; t++) Pown[t] = 0.0;
  LUNR for (int i = 0; i < 4; i++) Pown[sy(i, i)] = H.hx[i] + dw;
  Pown[sy(2, 0)] = H.hxy; Pown[sy(4, 0)] = H.bd[0]; Pown[sy(4, 2)] = H.bd[1];
  Pown[sy(4, 4)] = H.hd + ((g.j == 3) ? dw : 0.0);
  double Pout[15];
  LUNR for (int t = 0; t < 15; t++) { Pout[t] = 0.0; f.Pn[t] = Pown[t]; }
  bool ok = true;
  double n0, n1;
  const bool nullok = !FLOOR && l_null_dir<C>(c, n0, n1);   // NLP stage only (the QP stage has no state bound beyond the +-1e3 boxes)
  const double nsum = n0 + n1 + 1.0;
  f.nfix = 0;
#pragma unroll 1
  for (int step = 0; step < 4; step++) {
    if (step > 0) {
      LUNR for (int t = 0; t < 15; t++) { const double v = g_down(g, Pout[t]); if (g.j < 3) f.Pn[t] = Pown[t] + v; }
    }
    const double* Pn = f.Pn;
    double PA[25];   // Pn A^ (A^ = diag(A2, A2, 1))
    LUNR for (int i = 0; i < 5; i++) {
      LUNR for (int b = 0; b < 2; b++) {
        const double p0 = Pn[sy(i, 2 * b)], p1 = Pn[sy(i, 2 * b + 1)];
        PA[5 * i + 2 * b] = p0 * P.a11 + p1 * P.a21;
        PA[5 * i + 2 * b + 1] = p0 * P.a12 + p1 * P.a22;
      }
      PA[5 * i + 4] = Pn[sy(i, 4)];
    }
    double GA[10];   // B'Pn A^ (2 x 5)
    LUNR for (int cc = 0; cc < 2; cc++) {
      LUNR for (int m = 0; m < 5; m++) GA[5 * cc + m] = P.bp * PA[5 * (2 * cc) + m] + P.bv * PA[5 * (2 * cc + 1) + m];
    }
    // N2 = B'Pn B + diag(hu + dw)
    double N2[3];
    {
      const double g00 = P.bp * Pn[sy(0, 0)] + P.bv * Pn[sy(1, 0)], g01 = P.bp * Pn[sy(0, 1)] + P.bv * Pn[sy(1, 1)];
      const double g22 = P.bp * Pn[sy(2, 2)] + P.bv * Pn[sy(3, 2)], g23 = P.bp * Pn[sy(2, 3)] + P.bv * Pn[sy(3, 3)];
      const double g10 = P.bp * Pn[sy(2, 0)] + P.bv * Pn[sy(3, 0)], g11 = P.bp * Pn[sy(2, 1)] + P.bv * Pn[sy(3, 1)];
      N2[0] = P.bp * g00 + P.bv * g01 + (H.hu[0] + dw);
      N2[1] = P.bp * g10 + P.bv * g11;
      N2[2] = P.bp * g22 + P.bv * g23 + (H.hu[1] + dw);
    }
    // R~ = D' N2 D + (hl_last + dw) 11' + diag(hl + dw)
    double Rt[6];
    {
      double T1[6];
      LUNR for (int b = 0; b < NW; b++) { T1[b] = N2[0] * c.D[b] + N2[1] * c.D[NW + b]; T1[NW + b] = N2[1] * c.D[b] + N2[2] * c.D[NW + b]; }
      const double hlast = H.hl[C - 1] + dw;
      double dmax = 0.0;   // largest diagonal entry of the D part
      LUNR for (int a = 0; a < NW; a++) LUNR for (int b = 0; b <= a; b++) {
        const double dd = c.D[a] * T1[b] + c.D[NW + a] * T1[NW + b];
        double s = dd + hlast;
        if (a == b) { s += H.hl[a] + dw; dmax = fmax(dmax, dd); }
        Rt[(a * (a + 1)) / 2 + b] = s;
      }
      // basis T = [e0, e1, w_n] (l_null_dir) when the D part is what would bury the null direction's curvature -- three orders above
      // everything the lambda block puts into row 2 of T'R~T: that row and column then hold the lambda block alone.  Otherwise
      // (stiff support-weight bounds: they sit on the diagonal of the plain form, where the factorisation takes them first, and
      // would be spread over row 2 by T) the plain basis is the accurate one.
      if (nullok) {
        const double h0 = H.hl[0] + dw, h1 = H.hl[1] + dw, h2 = H.hl[NW - 1] + dw, hs = hlast * nsum;
        const double t0 = h0 * n0 * n0, t1 = h1 * n1 * n1, t2 = hs * nsum;
        f.nfix = (1e3 * fmax(fmax(t0, t1), fmax(t2, h2)) < dmax) ? 1 : 0;
        if (f.nfix) { Rt[3] = hs + h0 * n0; Rt[4] = hs + h1 * n1; Rt[5] = t2 + t0 + t1 + h2; }
      }
    }
    const bool nfix = f.nfix != 0;
    const bool okc = l_chol<NW, FLOOR>(Rt, f.Lc);
    // S~ = D' (GA + hxu on the position columns);  Y = L^-1 S~,  K = L^-T Y, column by column
    GA[0] += H.hxu[0]; GA[5 + 2] += H.hxu[1];
    double Y[15];
    LUNR for (int m = 0; m < 5; m++) {
      double col[3] = {0.0, 0.0, 0.0};
      LUNR for (int a = 0; a < NW; a++) col[a] = c.D[a] * GA[m] + c.D[NW + a] * GA[5 + m];
      if (nfix) col[NW - 1] = 0.0;   // w_n' D'(.) = 0
      l_chol_fwd<NW>(f.Lc, col);
      LUNR for (int a = 0; a < NW; a++) Y[5 * a + m] = col[a];
      l_chol_bwd<NW>(f.Lc, col);
      LUNR for (int a = 0; a < NW; a++) f.K[5 * a + m] = col[a];
    }
    // P_j (without lane j-1's own blocks) = A^'PA - Y'Y
    LUNR for (int i = 0; i < 5; i++) LUNR for (int k = 0; k <= i; k++) {
      double s;
      if (i < 4) { const int bi = i >> 1, ri = i & 1; s = (ri ? P.a12 : P.a11) * PA[5 * (2 * bi) + k] + (ri ? P.a22 : P.a21) * PA[5 * (2 * bi + 1) + k]; }
      else s = PA[5 * 4 + k];
      LUNR for (int a = 0; a < NW; a++) s -= Y[5 * a + i] * Y[5 * a + k];
      Pout[sy(i, k)] = s;
    }
    // the row kept in quasi-definite form (the distance row of x_{j+1} acts on xi_j through A^ and B)
    if (H.capped) {
      double col[3] = {0.0, 0.0, 0.0};
      LUNR for (int a = 0; a < NW; a++) col[a] = P.bp * (H.gv[0] * c.D[a] + H.gv[1] * c.D[NW + a]);
      if (nfix) col[NW - 1] = 0.0;
      l_chol_fwd<NW>(f.Lc, col);
      LUNR for (int a = 0; a < 3; a++) f.Ym[a] = (a < NW) ? col[a] : 0.0;
      l_chol_bwd<NW>(f.Lc, col);
      LUNR for (int a = 0; a < 3; a++) f.Zm[a] = (a < NW) ? col[a] : 0.0;
      const double Nq[5] = {H.gv[0] * P.a11, H.gv[0] * P.a12, H.gv[1] * P.a11, H.gv[1] * P.a12, 1.0};
      LUNR for (int m = 0; m < 5; m++) {
        double s = Nq[m];
        LUNR for (int a = 0; a < NW; a++) s -= f.Ym[a] * Y[5 * a + m];
        f.Nt[m] = s;
      }
      const double t00 = H.srinv + f.Ym[0] * f.Ym[0] + f.Ym[1] * f.Ym[1] + f.Ym[2] * f.Ym[2];
      f.Tc = l_rsq(t00);
      double V0[5];
      LUNR for (int m = 0; m < 5; m++) V0[m] = f.Nt[m] * f.Tc;
      LUNR for (int i = 0; i < 5; i++) LUNR for (int k = 0; k <= i; k++) Pout[sy(i, k)] += V0[i] * V0[k];
    }
    ok = okc;
  }
  // x_0 is fixed (the chains feed lane 0 with zeros) but delta_0 is free: its curvature is lane 0's Pout(4,4), which must be positive
  const double p0 = g_idx(g, Pout[sy(4, 4)], 0);
  f.p0dd = p0;
  return g_all(g, ok) && (p0 > 0.0) && (p0 < 1e300);
}

// ---------------------------------------------------------------------------------------------------------------
// Solve  [H A'; A 0][dx; dy] = [r1; r2]  with the current factorisation (r.xs,us,lm = r1 of this lane's variables, r.dl = this
// lane's SHARE of the slack variable's row; r.pi,nu,eta = r2).  One backward chain (lane j receives from lane j+1, adds its own
// right-hand side, keeps pv and kv, passes on), the slack variable from lane 0's cost-to-go, one forward chain.
// rq: right-hand side of this lane's quasi-definite row (g'dx - t / (Sigma - cap) = rq; zero in the first solve, the residual
// in refinement); tq: its multiplier t.
// ---------------------------------------------------------------------------------------------------------------
struct LaneMain { double rx[5], rw[3], cv[4], pv[5], kv[3], tt; };   // (rw, kv and the factors K, Zm live in the basis T of l_null_dir)

template <int C>
LDEV void l_stage_eliminate(const Grp& g, const KParams& P, const LaneConst& c, const LaneHess& H, double dw, const KV& r, LaneMain& m, double* d0,
                            bool nfix, double n0, double n1) {
  constexpr int NW = LaneDims<C>::NW;
  const double e0 = r.eta;
  LUNR for (int a = 0; a < 2; a++) d0[a] = c.Fl[a] * e0 + r.nu[a];
  const double last = r.lm[C - 1] - (H.hl[C - 1] + dw) * e0;
  const double ru0 = r.us[0] - (H.hu[0] + dw) * d0[0], ru1 = r.us[1] - (H.hu[1] + dw) * d0[1];
  LUNR for (int a = 0; a < 3; a++) m.rw[a] = 0.0;
  LUNR for (int a = 0; a < NW; a++) m.rw[a] = r.lm[a] - last + c.D[a] * ru0 + c.D[NW + a] * ru1;
  if (nfix) m.rw[NW - 1] = n0 * (r.lm[0] - last) + n1 * (r.lm[1] - last) + (r.lm[NW - 1] - last);   // w_n'(.): no D term
  LUNR for (int b = 0; b < 2; b++) {
    m.cv[2 * b] = P.bp * d0[b] - r.pi[2 * b];
    m.cv[2 * b + 1] = P.bv * d0[b] - r.pi[2 * b + 1];
  }
  // the friction coupling hxu_j d0_j acts on the position rows of x_j, i.e. of the previous lane
  LUNR for (int b = 0; b < 2; b++) {
    const double t = g_down(g, H.hxu[b] * d0[b]);
    m.rx[2 * b] = r.xs[2 * b] - ((g.j < 3) ? t : 0.0);
    m.rx[2 * b + 1] = r.xs[2 * b + 1];
  }
  m.rx[4] = r.dl;
}

template <int C>
LDEV void l_recover(const Grp& g, const KParams& P, const LaneConst& c, const LaneHess& H, double dw, const KV& r, const double* d0,
                    const double* w, KV& d) {
  constexpr int NW = LaneDims<C>::NW;
  const double e0 = r.eta;
  double wsum_ = 0.0, du0 = d0[0], du1 = d0[1];
  LUNR for (int a = 0; a < NW; a++) { d.lm[a] = w[a]; wsum_ += w[a]; du0 += c.D[a] * w[a]; du1 += c.D[NW + a] * w[a]; }
  const double dll = e0 - wsum_;
  d.lm[C - 1] = dll;
  LUNR for (int i = C; i < 4; i++) d.lm[i] = 0.0;
  d.us[0] = du0; d.us[1] = du1;
  // dx of the previous lane's positions (x_j), zero at lane 0
  const double xp0 = g_up(g, d.xs[0]), xp2 = g_up(g, d.xs[2]);
  const double du[2] = {du0, du1};
  const double xpp[2] = {(g.j > 0) ? xp0 : 0.0, (g.j > 0) ? xp2 : 0.0};
  LUNR for (int a = 0; a < 2; a++)
    d.nu[a] = r.us[a] - (H.hu[a] + dw) * du[a] - (P.bp * d.pi[2 * a] + P.bv * d.pi[2 * a + 1]) - H.hxu[a] * xpp[a];
  d.eta = r.lm[C - 1] - (H.hl[C - 1] + dw) * dll + c.Fl[0] * d.nu[0] + c.Fl[1] * d.nu[1];
}

template <int C, bool NL>
LDEV void l_kkt_solve(const Grp& g, const KParams& P, const LaneConst& c, const LaneHess& H, const LaneFact& f, double dw, const KV& r,
                      double rq, KV& d, double& tq) {
  constexpr int NW = LaneDims<C>::NW;
  LaneMain m; double d0[2];
  double n0, n1;
  const bool nfix = NL && l_null_dir<C>(c, n0, n1) && f.nfix != 0;
  l_stage_eliminate<C>(g, P, c, H, dw, r, m, d0, nfix, n0, n1);
  // ---- backward chain
  double out[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int step = 0; step < 4; step++) {
    double inc[5];
    LUNR for (int i = 0; i < 5; i++) { const double t = g_down(g, out[i]); inc[i] = (step > 0 && g.j < 3) ? t : 0.0; }
    double pn[5], q[5];
    LUNR for (int i = 0; i < 5; i++) pn[i] = m.rx[i] + inc[i];
    LUNR for (int i = 0; i < 5; i++) {
      double s = -pn[i];
      LUNR for (int k = 0; k < 4; k++) s += f.Pn[sy(i, k)] * m.cv[k];
      q[i] = s;
    }
    const double g0 = P.bp * q[0] + P.bv * q[1], g1 = P.bp * q[2] + P.bv * q[3];
    double k0[3] = {0.0, 0.0, 0.0}, kr[3] = {0.0, 0.0, 0.0}, ykr[3] = {0.0, 0.0, 0.0};
    LUNR for (int a = 0; a < NW; a++) { kr[a] = m.rw[a] - (c.D[a] * g0 + c.D[NW + a] * g1); k0[a] = kr[a]; }
    if (nfix) { kr[NW - 1] = m.rw[NW - 1]; k0[NW - 1] = kr[NW - 1]; }
    l_chol_fwd<NW>(f.Lc, k0);
    LUNR for (int a = 0; a < NW; a++) ykr[a] = k0[a];
    l_chol_bwd<NW>(f.Lc, k0);
    double tt = 0.0;
    if (H.capped) {
      double s = H.gv[0] * m.cv[0] + H.gv[1] * m.cv[2] - rq;
      LUNR for (int a = 0; a < NW; a++) s += f.Ym[a] * ykr[a];
      tt = s * f.Tc * f.Tc;
    }
    double aq[4];
    l_ATx(P, q, aq);
    LUNR for (int i = 0; i < 5; i++) {
      double s = (i < 4) ? -aq[i] : -q[4];
      LUNR for (int a = 0; a < NW; a++) s -= f.K[5 * a + i] * kr[a];
      if (H.capped) s -= f.Nt[i] * tt;
      out[i] = s;
    }
    LUNR for (int i = 0; i < 5; i++) m.pv[i] = pn[i];
    LUNR for (int a = 0; a < 3; a++) m.kv[a] = k0[a];
    m.tt = tt;
  }
  // ---- delta_0 (x_0 is fixed): P_0(4,4) ddelta = p_0(4)
  const double ddl = g_idx(g, out[4], 0) * l_rcp(f.p0dd);
  // ---- forward chain
  double xn[4] = {0.0, 0.0, 0.0, 0.0}, w[3] = {0.0, 0.0, 0.0};
  tq = 0.0;
#pragma unroll 1
  for (int step = 0; step < 4; step++) {
    double xp[5];
    LUNR for (int i = 0; i < 4; i++) { const double t = g_up(g, xn[i]); xp[i] = (step > 0 && g.j > 0) ? t : 0.0; }
    xp[4] = ddl;
    double tl = 0.0;
    if (H.capped) {
      double s = 0.0;
      LUNR for (int i = 0; i < 5; i++) s += f.Nt[i] * xp[i];
      tl = s * f.Tc * f.Tc + m.tt;
    }
    LUNR for (int a = 0; a < 3; a++) w[a] = 0.0;
    LUNR for (int a = 0; a < NW; a++) {
      double s = m.kv[a];
      LUNR for (int k = 0; k < 5; k++) s -= f.K[5 * a + k] * xp[k];
      if (H.capped) s -= f.Zm[a] * tl;
      w[a] = s;
    }
    if (nfix) { w[0] += n0 * w[NW - 1]; w[1] += n1 * w[NW - 1]; }   // w = T w'
    double e0 = 0.0, e1 = 0.0;
    LUNR for (int a = 0; a < NW; a++) { e0 += c.D[a] * w[a]; e1 += c.D[NW + a] * w[a]; }
    double ax[4];
    l_Ax(P, xp, ax);
    ax[0] += P.bp * e0; ax[1] += P.bv * e0; ax[2] += P.bp * e1; ax[3] += P.bv * e1;
    LUNR for (int i = 0; i < 4; i++) { ax[i] += m.cv[i]; xn[i] = ax[i]; }
    LUNR for (int i = 0; i < 4; i++) {
      double s = -m.pv[i] + f.Pn[sy(i, 4)] * ddl;
      LUNR for (int k = 0; k < 4; k++) s += f.Pn[sy(i, k)] * ax[k];
      d.xs[i] = ax[i]; d.pi[i] = s;
    }
    if (H.capped) { d.pi[0] += H.gv[0] * tl; d.pi[2] += H.gv[1] * tl; }
    tq = tl;
  }
  d.dl = ddl;
  l_recover<C>(g, P, c, H, dw, r, d0, w, d);
}

// ===============================================================================================================
// Slots: two-sided scalar constraints lo <= e(x) <= hi owned by this lane (reference rows: DESIGN.md "row map")
//   SX i : x_{j+1}[i]            +-box   MD:294-297,317-318        SU a : u_j[a]  +-box
//   SL i : lambda_j[i] in [0,1]  MD:299-300,320-321                SF a : p_j[a] - u_j[a], +-mu h / sqrt 2 (j >= 1)  MD:278-292,315
//   SV a : v_{j+1}[a]  +-v_max   DV:270-279,306-317 (NLP only)     SD   : |p_{j+1} - o|^2 + delta >= eps   DV:262-265,299-302 (NLP only)
// Per-slot numbers live in the thread-private columns.  Everything that is the same formula for every slot runs as a
// ROLLED loop over the slot index (small code: the kernel is instruction-fetch bound otherwise);
// only the maps between slots and variables (l_scatter: e = J v, l_gather: J'w, l_assemble) are unrolled.
// ===============================================================================================================
#ifndef LSLOT_UNROLL
#define LSLOT_UNROLL 2
#endif
#define LPRAGMA_(x) _Pragma(#x)
#define LPRAGMA(x) LPRAGMA_(x)
#define LSLOTS LPRAGMA(unroll LSLOT_UNROLL)
template <int C, bool NL> LDEV int l_nslots(const Grp& g) { typedef LaneDims<C> d; return (NL ? d::NS : d::NSQP) - (g.j == 0 ? 2 : 0); }
template <int C, bool NL> LDEV double l_slot_lo(const KParams& P, int s) { return P.slot_lo[NL ? 1 : 0][C == 4 ? 0 : 1][s]; }
template <int C, bool NL> LDEV double l_slot_hi(const KParams& P, int s) { return P.slot_hi[NL ? 1 : 0][C == 4 ? 0 : 1][s]; }
// e = J v of the linear slots -> field fld (stride-1 slot index); v = lane vector, vp = previous lane's xs (x_j).
// The distance slot (NLP) is written by the caller.
template <int C, bool NL>
LDEV void l_scatter(const Grp& g, const Priv& pr, int fld, const KV& v, const double* vp) {
  typedef LaneDims<C> d;
  LUNR for (int i = 0; i < 4; i++) pr(fld + d::SX + i) = v.xs[i];
  LUNR for (int a = 0; a < 2; a++) pr(fld + d::SU + a) = v.us[a];
  LUNR for (int i = 0; i < C; i++) pr(fld + d::SL + i) = v.lm[i];
  constexpr int SF = NL ? d::SFN : d::SFQ;
  LUNR for (int a = 0; a < 2; a++) pr(fld + SF + a) = vp[2 * a] - v.us[a];
  if (NL) { LUNR for (int a = 0; a < 2; a++) pr(fld + d::SV + a) = v.xs[2 * a + 1]; }
}
// acc.(xs,us,lm) = J'w for the per-slot weights w in field fld (weight of +grad e; 0 in the invalid slots);
// wd = weight of this lane's distance row = its share of the slack variable's entry
template <int C, bool NL>
LDEV void l_gather(const Grp& g, const Priv& pr, int fld, const double* gv, KV& acc, double& wd) {
  typedef LaneDims<C> d;
  constexpr int SF = NL ? d::SFN : d::SFQ;
  const double wf0 = (g.j > 0) ? pr(fld + SF) : 0.0, wf1 = (g.j > 0) ? pr(fld + SF + 1) : 0.0;
  LUNR for (int i = 0; i < 4; i++) acc.xs[i] = pr(fld + d::SX + i);
  acc.us[0] = pr(fld + d::SU) - wf0; acc.us[1] = pr(fld + d::SU + 1) - wf1;
  LUNR for (int i = 0; i < 4; i++) acc.lm[i] = (i < C) ? pr(fld + d::SL + i) : 0.0;
  { const double t = g_down(g, wf0); acc.xs[0] += (g.j < 3) ? t : 0.0; }
  { const double t = g_down(g, wf1); acc.xs[2] += (g.j < 3) ? t : 0.0; }
  wd = 0.0;
  if (NL) {
    const double wdj = pr(fld + d::SD);
    LUNR for (int a = 0; a < 2; a++) { acc.xs[2 * a + 1] += pr(fld + d::SV + a); acc.xs[2 * a] += gv[a] * wdj; }
    wd = wdj;
  }
}
// stage Hessian blocks from the slot weights Sigma in field fld; NL: distance-row data already in H (gv, sc, zd)
template <int C, bool NL>
LDEV void l_assemble(const Grp& g, const KParams& P, const LaneConst& c, const Priv& pr, int fld, double rho, double zd, LaneHess& H) {
  typedef LaneDims<C> d;
  LUNR for (int i = 0; i < 4; i++) H.hx[i] = c.qw + pr(fld + d::SX + i);
  LUNR for (int a = 0; a < 2; a++) {
    const double sf = (g.j > 0) ? pr(fld + (NL ? d::SFN : d::SFQ) + a) : 0.0;
    H.hu[a] = P.Rw + pr(fld + d::SU + a) + sf;
    H.hxu[a] = -sf;
    const double t = g_down(g, sf);
    H.hx[2 * a] += (g.j < 3) ? t : 0.0;
  }
  LUNR for (int i = 0; i < 4; i++) H.hl[i] = (i < C) ? rho + pr(fld + d::SL + i) : 0.0;
  if (NL) {
    LUNR for (int a = 0; a < 2; a++) { H.hx[2 * a + 1] += pr(fld + d::SV + a); H.hx[2 * a] += H.sc * H.gv[a] * H.gv[a] - 2.0 * zd; H.bd[a] = H.sc * H.gv[a]; }
    H.hxy = H.sc * H.gv[0] * H.gv[1];
    H.hd = H.sc + ((g.j == 3) ? P.slw : 0.0);
  } else {
    H.hxy = 0.0; H.bd[0] = 0.0; H.bd[1] = 0.0; H.hd = (g.j == 3) ? P.slw : 0.0;
  }
}

struct LaneInfo { int status, qp_iters, qp_flag, nlp_iters, n_fact; double kkt; };

// everything a lane keeps in registers across iterations
template <int C>
struct LaneSolver {
  typedef LaneDims<C> d;
  Grp g;
  Priv pr;
  LaneConst c;
  KV x;            // iterate: xs, us, lm, dl | pi, nu, eta
  LaneHess H;
  LaneFact f;
  LaneInfo info;
  int nsq, nsn;    // slot-loop lengths of this lane in the QP / NLP stage (lane 0 has no friction slots)

  LMEM void prev_state(const KV& v, bool homogeneous, double* xp) const {  // x_j: previous lane's xs; lane 0: measured state
    LUNR for (int i = 0; i < 4; i++) { const double t = g_up(g, v.xs[i]); xp[i] = (g.j > 0) ? t : (homogeneous ? 0.0 : pr(d::F_XIN + i)); }
  }
  LMEM void next_pi(const KV& v, double* pin) const { l_from_next(g, v.pi, pin); }

  // =============================================================================================================
  // Stage 1: QP, iSWIFT restated (Prime.c:127-230; kkt_initialize Auxilary.c:680-755)
  // Slot fields (stride NSQP): SLO SHI ZLO ZHI | A = 1/s_lo  B = 1/s_hi  C = 1/z_lo  D = 1/z_hi (then the step) |
  //                            E G = ds dz of the affine step (before that: Sigma, rhs weights) | EV = e(x) | DE = e(dx), weights
  // =============================================================================================================
  double sigma; int qp_it;

  // ---------------------------------------------------------------------------------------------------------------
  // Pivot foot.  The stage elimination removes ONE support weight through 1'lambda = 1 -- the one in the last place -- so the
  // reduced Hessian carries h_last 11' (h = the lambda block of H).  A last weight that sits on its bound (h = z / s ~ 1e10 ..
  // 1e16 late in a stage) buries the O(1) curvature of the other directions under that term: the Newton step lost its lambda
  // rows, the dual residual stalled (1.6e-4 against iSWIFT's 5.8e-7 limit) and the QP stage ran into its iteration limit where
  // iSWIFT converges -- the round-1 "known limit", one wide-fuzz instance in 1200.  The feet are therefore RELABELLED when the last
  // one becomes much stiffer than the softest (one always is soft: the weights sum to 1): the two change places in everything
  // this lane indexes by foot -- lambda, its slot columns, D and Fl -- and `perm` (position -> foot, 2 bits each) remembers the
  // order for the outputs.  Lane-local, no shuffle, at most a few times per solve (wide fuzz: 18 % of the 4-foot lanes, once); huge entries then sit on the
  // diagonal of the reduced Hessian only, where the Cholesky factorisation takes them first.
  // ---------------------------------------------------------------------------------------------------------------
  int perm;
  LMEM void repivot(bool nlp) {
    constexpr int NW = d::NW;
    int p = C - 1; double hm = H.hl[C - 1];
    LUNR for (int a = 0; a < NW; a++) { if (H.hl[a] < hm) { hm = H.hl[a]; p = a; } }
    if (p == C - 1 || !(H.hl[C - 1] > 1e3 * hm)) return;
    double dp0 = 0.0, dp1 = 0.0;
    LUNR for (int a = 0; a < NW; a++) {
      if (a == p) {
        dp0 = c.D[a]; dp1 = c.D[NW + a];
        const double t = x.lm[a]; x.lm[a] = x.lm[C - 1]; x.lm[C - 1] = t;
        const double h = H.hl[a]; H.hl[a] = H.hl[C - 1]; H.hl[C - 1] = h;
      }
    }
    LUNR for (int a = 0; a < NW; a++) { c.D[a] = ((a == p) ? 0.0 : c.D[a]) - dp0; c.D[NW + a] = ((a == p) ? 0.0 : c.D[NW + a]) - dp1; }
    c.Fl[0] += dp0; c.Fl[1] += dp1;
    // slot columns that carry state from one iteration to the next: slacks and multipliers (QP); multipliers and their pending steps (NLP)
    const int f0 = nlp ? d::N_ZLO : d::Q_SLO, f1 = nlp ? d::N_ZHI : d::Q_SHI, f2 = nlp ? d::N_D : d::Q_ZLO, f3 = nlp ? d::N_E : d::Q_ZHI;
    const int sa = d::SL + p, sb = d::SL + C - 1;
    { const double t = pr(f0 + sa); pr(f0 + sa) = pr(f0 + sb); pr(f0 + sb) = t; }
    { const double t = pr(f1 + sa); pr(f1 + sa) = pr(f1 + sb); pr(f1 + sb) = t; }
    { const double t = pr(f2 + sa); pr(f2 + sa) = pr(f2 + sb); pr(f2 + sb) = t; }
    { const double t = pr(f3 + sa); pr(f3 + sa) = pr(f3 + sb); pr(f3 + sb) = t; }
    const int fa = (perm >> (2 * p)) & 3, fb = (perm >> (2 * (C - 1))) & 3;
    perm = (perm & ~((3 << (2 * p)) | (3 << (2 * (C - 1))))) | (fb << (2 * p)) | (fa << (2 * (C - 1)));
  }

  // residual parts shared by qp_iter and qp_final_kkt: -(Px + c + A'y + G'z) and b - Ax; z-weights must be in Q_DE
  LMEM void qp_dual_eq(const KParams& P, const double* xp, const double* pin, KV& rx, double& acc_rx, double& acc_ry) {
    KV gfy, jz; double wd;
    l_grad_Aty<C>(P, c, pr, x, x, pin, true, g.j, gfy);
    l_gather<C, false>(g, pr, d::Q_DE, H.gv, jz, wd);
    acc_rx = 0.0; acc_ry = 0.0;
    LUNR for (int i = 0; i < 4; i++) { rx.xs[i] = -(gfy.xs[i] - jz.xs[i]); acc_rx += rx.xs[i] * rx.xs[i]; }
    LUNR for (int a = 0; a < 2; a++) { rx.us[a] = -(gfy.us[a] - jz.us[a]); acc_rx += rx.us[a] * rx.us[a]; }
    LUNR for (int i = 0; i < 4; i++) { rx.lm[i] = (i < C) ? -(gfy.lm[i] - jz.lm[i]) : 0.0; acc_rx += rx.lm[i] * rx.lm[i]; }
    rx.dl = -gfy.dl;
    acc_rx += rx.dl * rx.dl;
    double rpi[4], rnu[2], reta;
    l_eq_residual<C>(P, c, x, xp, false, rpi, rnu, reta);
    LUNR for (int i = 0; i < 4; i++) { rx.pi[i] = -rpi[i]; acc_ry += rpi[i] * rpi[i]; }
    LUNR for (int a = 0; a < 2; a++) { rx.nu[a] = -rnu[a]; acc_ry += rnu[a] * rnu[a]; }
    rx.eta = -reta; acc_ry += reta * reta;
  }

  // initial point.  Returns true when the stage is already finished (factorisation failure).
  LMEM bool qp_init(const KParams& P) {
    // W = I, rhs [-c; b; h]
    LSLOTS for (int s = 0; s < nsq; s++) {
      const bool act = true;
      pr(d::Q_E + s) = act ? 2.0 : 0.0;
      pr(d::Q_G + s) = act ? l_slot_hi<C, false>(P, s) + l_slot_lo<C, false>(P, s) : 0.0;
    }
    LUNR for (int i = 0; i < 4; i++) { x.xs[i] = 0.0; x.lm[i] = 0.0; x.pi[i] = 0.0; }
    x.us[0] = x.us[1] = 0.0; x.dl = 0.0; x.nu[0] = x.nu[1] = 0.0; x.eta = 0.0;
    H.gv[0] = H.gv[1] = 0.0; H.sc = 0.0; H.srinv = 0.0; H.capped = 0;
    l_assemble<C, false>(g, P, c, pr, d::Q_E, 0.0, 0.0, H);
    KV r, dd; double wd;
    l_gather<C, false>(g, pr, d::Q_G, H.gv, r, wd);
    LUNR for (int i = 0; i < 4; i++) r.xs[i] += c.qw * pr(d::F_XDES + i);
    r.dl = 0.0;
    {
      double xp[4]; prev_state(x, false, xp);
      double rpi[4], rnu[2], reta;
      l_eq_residual<C>(P, c, x, xp, false, rpi, rnu, reta);
      LUNR for (int i = 0; i < 4; i++) r.pi[i] = -rpi[i];
      r.nu[0] = -rnu[0]; r.nu[1] = -rnu[1]; r.eta = -reta;
    }
    info.qp_flag = 3; info.qp_iters = 0; qp_it = 0; sigma = 100.0;  // GlobalOptions.h:26
    {
      const bool okr = l_factor<C, true>(g, P, c, H, 0.0, f);
      g_keep(g, okr);
      if (!okr) { info.qp_flag = 1; return true; }
    }
    { double tq; l_kkt_solve<C, false>(g, P, c, H, f, 0.0, r, 0.0, dd, tq); }
    x = dd;
    LUNR for (int i = C; i < 4; i++) x.lm[i] = 0.0;
    // slacks and multipliers shifted positive (Auxilary.c:724-746)
    double xp[4]; prev_state(x, false, xp);
    l_scatter<C, false>(g, pr, d::Q_EV, x, xp);
    double rmin = 1e300, rmax = -1e300;
    LSLOTS for (int s = 0; s < nsq; s++) {
      {
        const double e = pr(d::Q_EV + s);
        const double rl = e - l_slot_lo<C, false>(P, s), rh = l_slot_hi<C, false>(P, s) - e;
        rmin = fmin(rmin, fmin(rl, rh)); rmax = fmax(rmax, fmax(rl, rh));
      }
    }
    rmin = g_min(g, rmin); rmax = g_max(g, rmax);
    const double ap = -rmin, ad = rmax;
    LSLOTS for (int s = 0; s < nsq; s++) {
      const bool act = true;
      const double e = pr(d::Q_EV + s);
      const double rl = e - l_slot_lo<C, false>(P, s), rh = l_slot_hi<C, false>(P, s) - e;
      pr(d::Q_SLO + s) = act ? ((ap < 0) ? rl : rl + (1.0 + ap)) : 1.0;
      pr(d::Q_SHI + s) = act ? ((ap < 0) ? rh : rh + (1.0 + ap)) : 1.0;
      pr(d::Q_ZLO + s) = act ? ((ad < 0) ? -rl : -rl + (1.0 + ad)) : 0.0;
      pr(d::Q_ZHI + s) = act ? ((ad < 0) ? -rh : -rh + (1.0 + ad)) : 0.0;
    }
    info.qp_flag = 2;
    return false;
  }

  // one Mehrotra predictor-corrector iteration (Prime.c:138-218).  Returns true when the stage is finished.
  // Slot passes: A (residuals, reciprocals, predictor weights), B (affine step statistics), C (corrector right-hand side),
  // D (final step and step lengths), E (update).  One instance of the KKT solve serves predictor and corrector.
  LMEM bool qp_iter(const KParams& P) {   // (the iteration limit is tested at the end: the caller never enters with qp_it >= qp_maxit)
    if (sigma > 0.0) repivot(false);   // (a corrector-only step keeps the previous factor, and with it the order of the feet)
    double xp[4], pin[4];
    prev_state(x, false, xp); next_pi(x, pin);
    l_scatter<C, false>(g, pr, d::Q_EV, x, xp);
    // ---- pass A: residuals (computeresiduals, Auxilary.c:524-553); 1/s -> A, B ; 1/z -> C, D ; predictor: Sigma -> E, rhs -> G
    KV rx;  // rx.(xs,us,lm,dl) = -(Px + G'z + A'y + c), rx.(pi,nu,eta) = b - Ax
    double sz, mu;
    const bool predictor = sigma > 0.0;
    {
      double acc_rz = 0.0, acc_sz = 0.0;
      LSLOTS for (int s = 0; s < nsq; s++) {
        const double lo = l_slot_lo<C, false>(P, s), hi = l_slot_hi<C, false>(P, s), e = pr(d::Q_EV + s);
        const double slo = pr(d::Q_SLO + s), shi = pr(d::Q_SHI + s), zlo = pr(d::Q_ZLO + s), zhi = pr(d::Q_ZHI + s);
        const double rzh = hi - shi - e, rzl = -lo - slo + e;
        acc_rz += rzh * rzh + rzl * rzl;
        acc_sz += slo * zlo + shi * zhi;
        pr(d::Q_DE + s) = zlo - zhi;
        const double isl = l_rcp(slo), ish = l_rcp(shi);
        pr(d::Q_A + s) = isl; pr(d::Q_B + s) = ish;
        pr(d::Q_C + s) = l_rcp(zlo); pr(d::Q_D + s) = l_rcp(zhi);
        // predictor: W = s/z, ds = -s z, rhs3 = rz - ds/z = rz + s
        const double wl = zlo * isl, wh = zhi * ish;
        pr(d::Q_E + s) = predictor ? wl + wh : 0.0;
        pr(d::Q_G + s) = predictor ? wh * (rzh + shi) - wl * (rzl + slo) : 0.0;
      }
      double acc_rx, acc_ry;
      qp_dual_eq(P, xp, pin, rx, acc_rx, acc_ry);
      const double nrx = g_sum(g, acc_rx), nry = g_sum(g, acc_ry), nrz = g_sum(g, acc_rz);
      sz = g_sum(g, acc_sz);
      const double lim = P.qp_reltol / sqrt(3.0);
      LANE_DBG("qp it %d rx %.3e ry %.3e rz %.3e mu %.3e sigma %.3e\n", qp_it, sqrt(nrx), sqrt(nry), sqrt(nrz), sz / d::MQP, sigma);
      const bool conv = sqrt(nrx) < lim && sqrt(nrz) < lim && sqrt(nry) < lim && sz / d::MQP < P.qp_abstol;
      g_keep(g, !conv);
      if (conv) { info.qp_flag = 0; return true; }
      mu = sz / d::MQP;
    }
    {
      bool okr = true;
      const unsigned sv = g_enter(g, predictor);   // (sigma == 0 -- a pure corrector step with the previous factor, Prime.c:176-180 -- is rare)
      if (predictor) {
        l_assemble<C, false>(g, P, c, pr, d::Q_E, 0.0, 0.0, H);
        info.n_fact++;
        okr = l_factor<C, true>(g, P, c, H, 0.0, f);
      } else {
        sigma = 0.0;  // E, G hold zeros
      }
      g_leave(g, sv);
      g_keep(g, okr);
      if (!okr) { info.qp_flag = 1; return true; }
    }
    double ap = 1.0, ad = 1.0;
    KV dd;
    const bool any_pred = w_any(g, predictor);
#pragma unroll 1
    for (int ph = any_pred ? 0 : 1; ph < 2; ph++) {   // (a group without a predictor step runs ph 0 as a zombie)
      const bool mine = ph == 1 || predictor;
      const double sm = sigma * mu;
      if (ph == 1) {
        // ---- pass C: corrector right-hand side, ds = -s z - ds_aff dz_aff + sigma mu (form_ds, Auxilary.c:250-267) -> DE
        LSLOTS for (int s = 0; s < nsq; s++) {
          const double lo = l_slot_lo<C, false>(P, s), hi = l_slot_hi<C, false>(P, s), e = pr(d::Q_EV + s);
          const double slo = pr(d::Q_SLO + s), shi = pr(d::Q_SHI + s), zlo = pr(d::Q_ZLO + s), zhi = pr(d::Q_ZHI + s);
          const double wl = zlo * pr(d::Q_A + s), wh = zhi * pr(d::Q_B + s);
          const double rzh = hi - shi - e, rzl = -lo - slo + e;
          const double cdl = -slo * zlo - pr(d::Q_E + s) + sm, cdh = -shi * zhi - pr(d::Q_G + s) + sm;
          const double r3l = rzl - cdl * pr(d::Q_C + s), r3h = rzh - cdh * pr(d::Q_D + s);
          pr(d::Q_DE + s) = wh * r3h - wl * r3l;
        }
      }
      KV r; double wd;
      l_gather<C, false>(g, pr, ph == 0 ? d::Q_G : d::Q_DE, H.gv, r, wd);
      LUNR for (int i = 0; i < 4; i++) { r.xs[i] += rx.xs[i]; r.lm[i] += rx.lm[i]; r.pi[i] = rx.pi[i]; }
      LUNR for (int a = 0; a < 2; a++) { r.us[a] += rx.us[a]; r.nu[a] = rx.nu[a]; }
      r.dl = rx.dl; r.eta = rx.eta;
      { double tq; l_kkt_solve<C, false>(g, P, c, H, f, 0.0, r, 0.0, dd, tq); }
      double dxp[4]; prev_state(dd, true, dxp);
      l_scatter<C, false>(g, pr, d::Q_DE, dd, dxp);
      // ---- pass B (affine: max step lengths, findsteplength Auxilary.c:271-294, and rho, Auxilary.c:600-609)
      //      or pass D (final step: ds -> A, B ; dz -> C, D)
      double rp = 0.0, rdm = 0.0, s_dsz = 0.0, s_sdz = 0.0, s_dsdz = 0.0;
      const bool aff = ph == 0;
      LSLOTS for (int s = 0; s < nsq; s++) {
        const double lo = l_slot_lo<C, false>(P, s), hi = l_slot_hi<C, false>(P, s), e = pr(d::Q_EV + s), de = pr(d::Q_DE + s);
        const double slo = pr(d::Q_SLO + s), shi = pr(d::Q_SHI + s), zlo = pr(d::Q_ZLO + s), zhi = pr(d::Q_ZHI + s);
        const double isl = pr(d::Q_A + s), ish = pr(d::Q_B + s), izl = pr(d::Q_C + s), izh = pr(d::Q_D + s);
        const double wl = zlo * isl, wh = zhi * ish;
        const double rzh = hi - shi - e, rzl = -lo - slo + e;
        double dzh, dzl, dsh, dsl;
        if (aff) {
          dzh = wh * (de - (rzh + shi)); dzl = wl * (-de - (rzl + slo));
          dsh = (-shi * zhi - shi * dzh) * izh; dsl = (-slo * zlo - slo * dzl) * izl;
        } else {
          const double cdl = -slo * zlo - pr(d::Q_E + s) + sm, cdh = -shi * zhi - pr(d::Q_G + s) + sm;
          const double r3l = rzl - cdl * izl, r3h = rzh - cdh * izh;
          dzh = wh * (de - r3h); dzl = wl * (-de - r3l);
          dsh = (cdh - shi * dzh) * izh; dsl = (cdl - slo * dzl) * izl;
        }
        rp = fmax(rp, fmax(-dsh * ish, -dsl * isl));
        rdm = fmax(rdm, fmax(-dzh * izh, -dzl * izl));
        if (aff) {
          s_dsz += dsl * zlo + dsh * zhi; s_sdz += slo * dzl + shi * dzh; s_dsdz += dsl * dzl + dsh * dzh;
          if (mine) { pr(d::Q_E + s) = dsl * dzl; pr(d::Q_G + s) = dsh * dzh; }
        } else {
          pr(d::Q_A + s) = dsl; pr(d::Q_B + s) = dsh; pr(d::Q_C + s) = dzl; pr(d::Q_D + s) = dzh;
        }
      }
      rp = g_max(g, rp); rdm = g_max(g, rdm);
      ap = (rp > 0.0) ? 1.0 / rp : 1.0; ad = (rdm > 0.0) ? 1.0 / rdm : 1.0;
      if (aff) {
        // sum (s + ap ds)(z + ad dz) = s'z + ap ds'z + ad s'dz + ap ad ds'dz
        const double num = sz + g_sum(g, ap * s_dsz + ad * s_sdz + (ap * ad) * s_dsdz);
        const double rho = num / sz, mr = rho < 1.0 ? rho : 1.0;
        if (mine) {
          sigma = mr * mr * mr;
          if (sigma < 0.0) sigma = 0.0;
        }
      }
    }
    ap = fmin(0.99 * ap, 1.0); ad = fmin(0.99 * ad, 1.0);
    LUNR for (int i = 0; i < 4; i++) { x.xs[i] += ap * dd.xs[i]; x.pi[i] += ad * dd.pi[i]; }
    LUNR for (int i = 0; i < C; i++) x.lm[i] += ap * dd.lm[i];
    LUNR for (int a = 0; a < 2; a++) { x.us[a] += ap * dd.us[a]; x.nu[a] += ad * dd.nu[a]; }
    x.dl += ap * dd.dl; x.eta += ad * dd.eta;
    // ---- pass E
    LSLOTS for (int s = 0; s < nsq; s++) {
      pr(d::Q_SLO + s) += ap * pr(d::Q_A + s); pr(d::Q_SHI + s) += ap * pr(d::Q_B + s);
      pr(d::Q_ZLO + s) += ad * pr(d::Q_C + s); pr(d::Q_ZHI + s) += ad * pr(d::Q_D + s);
    }
    qp_it++;
    info.qp_iters = qp_it;
    return qp_it >= P.qp_maxit;
  }

  // QP-only answer (use_snopt == false): unscaled KKT error of the QP
  LMEM void qp_final_kkt(const KParams& P) {
    double xp[4], pin[4];
    prev_state(x, false, xp); next_pi(x, pin);
    l_scatter<C, false>(g, pr, d::Q_EV, x, xp);
    double ec = 0.0, ep = 0.0;
    LSLOTS for (int s = 0; s < nsq; s++) {
      double wz = 0.0;
      {
        const double lo = l_slot_lo<C, false>(P, s), hi = l_slot_hi<C, false>(P, s), e = pr(d::Q_EV + s);
        const double slo = pr(d::Q_SLO + s), shi = pr(d::Q_SHI + s), zlo = pr(d::Q_ZLO + s), zhi = pr(d::Q_ZHI + s);
        wz = zlo - zhi;
        ec = fmax(ec, fmax(fabs(slo * zlo), fabs(shi * zhi)));
        ep = fmax(ep, fmax(fmax(lo - e, e - hi), 0.0));
      }
      pr(d::Q_DE + s) = wz;
    }
    KV rx; double acc_rx, acc_ry;
    qp_dual_eq(P, xp, pin, rx, acc_rx, acc_ry);
    double rd = fabs(rx.dl), re = fabs(rx.eta);
    LUNR for (int i = 0; i < 4; i++) { rd = fmax(rd, fabs(rx.xs[i])); re = fmax(re, fabs(rx.pi[i])); }
    LUNR for (int a = 0; a < 2; a++) { rd = fmax(rd, fabs(rx.us[a])); re = fmax(re, fabs(rx.nu[a])); }
    LUNR for (int i = 0; i < C; i++) rd = fmax(rd, fabs(rx.lm[i]));
    info.kkt = g_max(g, fmax(fmax(rd, re), fmax(ec, ep)));
    info.status = (info.qp_flag == 0) ? 0 : 1;
  }

  // =============================================================================================================
  // Stage 2: NLP, IPM-F
  // Slot fields (stride NS): SLO SHI ZLO ZHI | A = 1/s_lo  B = 1/s_hi | C = e(dx)  D = dz_lo  E = dz_hi |
  //                          G, EV = scratch weights (e(x) at the top of the iteration; refinement)
  // The distance slot SD is one-sided: its "hi" half is inert (s_hi = 1, z_hi = 0, never updated).
  // =============================================================================================================
  double mu; int nl_it, nfilt; bool first_ic;
  // tau, dw_last, theta_max, theta_min, ad_prev: private columns F_PERS + 0 .. 4
#define PERS_TAU pr(d::F_PERS + 0)
#define PERS_DWLAST pr(d::F_PERS + 1)
#define PERS_THMAX pr(d::F_PERS + 2)
#define PERS_THMIN pr(d::F_PERS + 3)
#define PERS_ADPREV pr(d::F_PERS + 4)

  // e(x) of every slot -> field fld
  LMEM void nlp_scatter_e(int fld, const double* xp) {
    l_scatter<C, true>(g, pr, fld, x, xp);
    const double a = x.xs[0] - c.ox, b = x.xs[2] - c.oy;
    pr(fld + d::SD) = a * a + b * b + x.dl;
  }
  // e(dx) (directional derivative) of every slot -> field fld
  LMEM void nlp_scatter_de(int fld, const KV& dv, const double* dvp) {
    l_scatter<C, true>(g, pr, fld, dv, dvp);
    pr(fld + d::SD) = H.gv[0] * dv.xs[0] + H.gv[1] * dv.xs[2] + dv.dl;
  }

  // The filter is a ring of P.filter_cap (theta, phi) pairs shared by the group: pair p lives in lane p & 3, slot p >> 2.  Empty
  // slots hold theta = +inf, which never blocks a trial point, so the acceptance test needs no count.
  LMEM void filter_reset() {
    LUNR for (int q = 0; q < 2; q++) { pr(d::F_FILT + 2 * q) = CBF_INF; pr(d::F_FILT + 2 * q + 1) = CBF_INF; }
    nfilt = 0;
  }
  LMEM void filter_add(const KParams& P, double th_f, double ph_f) {  // overwrites the oldest pair once the ring is full
    const int p = nfilt % P.filter_cap;
    if ((p & 3) == g.j) {
      if (p >> 2) { pr(d::F_FILT + 2) = th_f; pr(d::F_FILT + 3) = ph_f; } else { pr(d::F_FILT) = th_f; pr(d::F_FILT + 1) = ph_f; }
    }
    nfilt++;
  }
  LMEM bool filter_ok(double tht, double pht) const {
    bool okl = true;
    LUNR for (int q = 0; q < 2; q++) okl = okl && !(tht >= pr(d::F_FILT + 2 * q) && pht >= pr(d::F_FILT + 2 * q + 1));
    return g_all(g, okl);
  }

  // Returns true when the stage is finished already: data that leave some row without an interior start (non-finite inputs)
  // are answered with status 4, like the oracle does.
  LMEM bool nlp_init(const KParams& P) {
    double chk = 0.0 * x.dl + 0.0 * x.us[0] + 0.0 * x.us[1];  // a QP stage that broke down leaves NaN behind: no start point
    LUNR for (int i = 0; i < 4; i++) chk += 0.0 * x.xs[i];
    LUNR for (int i = 0; i < C; i++) chk += 0.0 * x.lm[i];
    // ---- step 0: strictly interior start (X, then U against friction, then lambda, then the slack variable)
    LUNR for (int i = 0; i < 4; i++) {
      const double lim = (i & 1) ? fmin(P.box, P.vmax) : P.box;
      x.xs[i] = fmin(fmax(x.xs[i], -lim + P.push), lim - P.push);
    }
    LUNR for (int i = 0; i < C; i++) x.lm[i] = fmin(fmax(x.lm[i], P.push), 1.0 - P.push);
    double xp[4]; prev_state(x, false, xp);
    LUNR for (int a = 0; a < 2; a++) {
      double lo = -P.box + P.push, hi = P.box - P.push;
      if (g.j >= 1) { lo = fmax(lo, xp[2 * a] - P.fb + P.push); hi = fmin(hi, xp[2 * a] + P.fb - P.push); }
      x.us[a] = fmin(fmax(x.us[a], lo), hi);
    }
    {
      const double a = x.xs[0] - c.ox, b = x.xs[2] - c.oy;
      const double cmin = g_min(g, a * a + b * b + x.dl - P.eps);
      if (cmin < P.push) x.dl += P.push - cmin;
    }
    LUNR for (int i = 0; i < 4; i++) x.pi[i] = 0.0;
    x.nu[0] = x.nu[1] = 0.0; x.eta = 0.0;
    filter_reset();
    mu = P.mu_init; PERS_TAU = fmax(P.tau_min, 1.0 - mu); PERS_DWLAST = 0.0; first_ic = true; nfilt = 0; nl_it = 0; PERS_ADPREV = 0.0;
    nlp_scatter_e(d::N_EV, xp);
    bool interior = chk == 0.0;
    LSLOTS for (int s = 0; s < nsn; s++) {
      const bool v = true, vhi = s != d::SD;
      const double e = pr(d::N_EV + s);
      const double slo = v ? e - l_slot_lo<C, true>(P, s) : 1.0, shi = vhi ? l_slot_hi<C, true>(P, s) - e : 1.0;
      interior = interior && (slo > 0.0) && (shi > 0.0);
      pr(d::N_SLO + s) = slo; pr(d::N_SHI + s) = shi;
      pr(d::N_ZLO + s) = v ? mu * l_rcp(slo) : 0.0; pr(d::N_ZHI + s) = vhi ? mu * l_rcp(shi) : 0.0;
    }
    {
      double rpi[4], rnu[2], reta;
      l_eq_residual<C>(P, c, x, xp, false, rpi, rnu, reta);
      double r = fabs(reta) + fabs(rnu[0]) + fabs(rnu[1]);
      LUNR for (int i = 0; i < 4; i++) r += fabs(rpi[i]);
      const double th0 = g_sum(g, r);
      PERS_THMAX = 1e4 * fmax(1.0, th0); PERS_THMIN = 1e-4 * fmax(1.0, th0);
    }
    info.status = 1; info.nlp_iters = 0;
    if (!g_all(g, interior)) { info.status = 4; info.kkt = CBF_INF; return true; }
    return false;
  }

  // one IPM-F iteration.  Returns true when the stage is finished (info.status set).
  // Slot passes per iteration: A (slacks, pending multiplier step + reset, reciprocals, weights), B (multiplier steps and
  // the residual weights of the unreduced Newton system; B' after a refinement solve), C (step bounds, barrier terms),
  // D (one per line-search trial).  The bordered KKT solve is instantiated once and shared by the main solve and the
  // refinement solves (code footprint: DESIGN.md section 5).
  LMEM bool nlp_iter(const KParams& P) {
    const int it = nl_it;
    if (it > 0) repivot(true);
    double xp[4], pin[4];
    prev_state(x, false, xp); next_pi(x, pin);
    // ---- pass A
    KV gfy; double re_pi[4], re_nu[2], re_eta, e_d, e_p, th, zd, szmin, szmax;
    {
      int zhw = 0;          // running maximum of the multipliers' high words (z >= 0; a NaN shows up as a huge value)
      bool spos = true;     // every slack positive
      szmin = 1e300; szmax = 0.0;
      const double iks = 1.0 / P.kappa_sigma, adp = PERS_ADPREV;
      nlp_scatter_e(d::N_EV, xp);
      LSLOTS for (int s = 0; s < nsn; s++) {
        const bool vhi = s != d::SD;
        const double e = pr(d::N_EV + s);
        const double slo = e - l_slot_lo<C, true>(P, s), shi = vhi ? l_slot_hi<C, true>(P, s) - e : 1.0;
        const double rlo = l_rcp(slo), rhi = l_rcp(shi);
        double zlo = pr(d::N_ZLO + s), zhi = pr(d::N_ZHI + s);
        spos = spos && (slo > 0.0) && (shi > 0.0);
        if (it > 0) {  // multiplier step of the previous iteration, then the z reset (Ipopt eq. 16) with that iteration's mu
          zlo += adp * pr(d::N_D + s);
          if (vhi) zhi += adp * pr(d::N_E + s);
          // z in [mu / (kappa s), kappa mu / s] <=> s z in [mu / kappa, kappa mu]: with kappa = 1e10 the clamp is almost never active
          const double pl = slo * zlo, ph_ = shi * zhi;
          if (!(pl >= mu * iks && pl <= P.kappa_sigma * mu)) { const double mr_lo = mu * rlo; zlo = fmin(fmax(zlo, mr_lo * iks), P.kappa_sigma * mr_lo); }
          if (vhi && !(ph_ >= mu * iks && ph_ <= P.kappa_sigma * mu)) { const double mr_hi = mu * rhi; zhi = fmin(fmax(zhi, mr_hi * iks), P.kappa_sigma * mr_hi); }
          pr(d::N_ZLO + s) = zlo; pr(d::N_ZHI + s) = zhi;
        }
        pr(d::N_SLO + s) = slo; pr(d::N_SHI + s) = shi; pr(d::N_A + s) = rlo; pr(d::N_B + s) = rhi;
        const double szl = slo * zlo, szh = shi * zhi;
        const double szh_ = vhi ? szh : szl;   // the distance slot has no upper half
        szmin = fmin(szmin, fmin(szl, szh_)); szmax = fmax(szmax, fmax(szl, szh_));
        zhw = l_imax(zhw, l_imax(l_hiword(zlo), l_hiword(zhi)));
        pr(d::N_G + s) = zlo - zhi;                    // J'z weights
        pr(d::N_C + s) = zlo * rlo + zhi * rhi;        // Sigma
        pr(d::N_D + s) = rlo - (vhi ? rhi : 0.0);      // barrier gradient weights / mu
      }
      H.gv[0] = 2.0 * (x.xs[0] - c.ox); H.gv[1] = 2.0 * (x.xs[2] - c.oy);
      zd = pr(d::N_ZLO + d::SD);
      // ---- KKT error
      KV jz; double wd;
      l_grad_Aty<C>(P, c, pr, x, x, pin, true, g.j, gfy);
      l_gather<C, true>(g, pr, d::N_G, H.gv, jz, wd);
      double rdm = fabs(P.slw * x.dl - g_sum(g, wd));   // the slack variable's row: the sum of the lanes' shares
      LUNR for (int i = 0; i < 4; i++) rdm = fmax(rdm, fabs(gfy.xs[i] - jz.xs[i]));
      LUNR for (int a = 0; a < 2; a++) rdm = fmax(rdm, fabs(gfy.us[a] - jz.us[a]));
      LUNR for (int i = 0; i < C; i++) rdm = fmax(rdm, fabs(gfy.lm[i] - jz.lm[i]));
      l_eq_residual<C>(P, c, x, xp, false, re_pi, re_nu, re_eta);
      double rem = fabs(re_eta), res = fabs(re_eta);
      LUNR for (int i = 0; i < 4; i++) { rem = fmax(rem, fabs(re_pi[i])); res += fabs(re_pi[i]); }
      LUNR for (int a = 0; a < 2; a++) { rem = fmax(rem, fabs(re_nu[a])); res += fabs(re_nu[a]); }
      e_d = g_max(g, rdm); e_p = g_max(g, rem);
      szmin = g_min(g, szmin); szmax = g_max(g, szmax);
      const double e_c = szmax;                                  // complementarity error max_i s_i z_i (s, z > 0)
      const double zmax = g_max(g, l_from_hiword(zhw));          // to 2^-20 relative: only compared with z_div
      th = g_sum(g, res);
      const bool nan_seen = g_any(g, !(rdm == rdm) || !(rem == rem) || !(e_c == e_c) || !(zmax == zmax));
      const double E0 = nan_seen ? 1e300 : fmax(e_d, fmax(e_p, e_c));
      info.kkt = E0; info.nlp_iters = it;
      LANE_LOG(1000.0 + it); LANE_LOG(mu); LANE_LOG(E0); LANE_LOG(e_d); LANE_LOG(e_p); LANE_LOG(e_c); LANE_LOG(th);
      LANE_DBG("it %d mu %.2e E0 %.3e ed %.2e ep %.2e ec %.2e zmax %.2e th %.2e\n", it, mu, E0, e_d, e_p, e_c, zmax, th);
#ifdef CBF_DBGBUF
      {   // dual residual by variable class: X, U, lambda, slack variable
        double tx = 0, tu = 0, tl = 0;
        LUNR for (int i = 0; i < 4; i++) tx = fmax(tx, fabs(gfy.xs[i] - jz.xs[i]));
        LUNR for (int a = 0; a < 2; a++) tu = fmax(tu, fabs(gfy.us[a] - jz.us[a]));
        LUNR for (int i = 0; i < C; i++) tl = fmax(tl, fabs(gfy.lm[i] - jz.lm[i]));
        const double tdl = fabs(P.slw * x.dl - g_sum(g, wd));
        for (int k = 0; k < 4; k++) {
          const double v0 = g_idx(g, tx, k), v1 = g_idx(g, tu, k), v2 = g_idx(g, tl, k);
          LANE_LOG(-3000.0 - k); LANE_LOG(v0); LANE_LOG(v1); LANE_LOG(v2); LANE_LOG(tdl);
        }
      }
#endif
#ifdef CBF_TRACE
      {
        double tx = 0, tu = 0, tl = 0, gx = 0, jx = 0, py = 0, ny = 0;
        LUNR for (int i = 0; i < 4; i++) { tx = fmax(tx, fabs(gfy.xs[i] - jz.xs[i])); gx = fmax(gx, fabs(gfy.xs[i])); jx = fmax(jx, fabs(jz.xs[i])); py = fmax(py, fabs(x.pi[i])); }
        LUNR for (int a = 0; a < 2; a++) { tu = fmax(tu, fabs(gfy.us[a] - jz.us[a])); ny = fmax(ny, fabs(x.nu[a])); }
        LUNR for (int i = 0; i < C; i++) tl = fmax(tl, fabs(gfy.lm[i] - jz.lm[i]));
        tx = g_max(g, tx); tu = g_max(g, tu); tl = g_max(g, tl); gx = g_max(g, gx); jx = g_max(g, jx); py = g_max(g, py); ny = g_max(g, ny);
        const double ey = g_max(g, fabs(x.eta)), xm = g_max(g, fmax(fabs(x.xs[0]), fabs(x.xs[2])));
        LANE_DBG("   ed parts: x %.2e u %.2e lam %.2e dl %.2e | |gfy.x| %.2e |jz.x| %.2e |pi| %.2e |nu| %.2e |eta| %.2e |pos| %.2e szmin %.2e szmax %.2e\n",
                 tx, tu, tl, fabs(P.slw * x.dl - g_sum(g, wd)), gx, jx, py, ny, ey, xm, szmin, szmax);
      }
#endif
      // converged: KKT error <= tol, or <= acc_tol once the barrier parameter sits at its floor (there the dual residual is at
      // its FP64 noise level, s = mu / z ~ 1e-13 for active rows, and may hover just above tol for many iterations)
      const bool anyneg = g_any(g, !spos);
      int ex = -1;
      if (E0 <= P.tol || (mu <= P.tol / 10.0 && E0 <= P.acc_tol)) ex = 0;
      else if (it == P.max_iter) ex = 1;
      else if (!(zmax < P.z_div) || !(E0 < 1e300) || anyneg) ex = 2;
      g_keep(g, ex < 0);
      if (ex >= 0) { info.status = ex; return true; }
    }
    // ---- barrier update (Ipopt eq. 7): max_i |s_i z_i - mu| from the extreme products
    for (;;) {
      const double emu = fmax(fmax(szmax - mu, mu - szmin), fmax(e_d, e_p));
      if (emu <= P.kappa_eps * mu && mu > P.tol / 10.0) {
        mu = fmax(P.tol / 10.0, fmin(P.kappa_mu * mu, mu * sqrt(mu)));
        PERS_TAU = fmax(P.tau_min, 1.0 - mu);
        filter_reset();
      } else break;
    }
    // ---- Hessian blocks and right-hand side (distance row: capped split)
    KV r;
    bool qd_mine, any_qd;   // some row of this group / of some group of the warp is stiff enough for the Newton residual to be worth
                            // a check (see the refinement loop; the groups of a warp go through the check together)
    {
      const double sg = pr(d::N_C + d::SD);
      qd_mine = g_any(g, sg > LANE_QD_CHECK);
      any_qd = w_any(g, qd_mine);
      const bool cap = sg > P.sigma_cap;
      H.sc = fmin(sg, P.sigma_cap); H.capped = cap ? 1 : 0; H.srinv = cap ? l_rcp(sg - H.sc) : 0.0;
      l_assemble<C, true>(g, P, c, pr, d::N_C, P.rho_lambda, zd, H);
      double wd;
      l_gather<C, true>(g, pr, d::N_D, H.gv, r, wd);
      LUNR for (int i = 0; i < 4; i++) { r.xs[i] = mu * r.xs[i] - gfy.xs[i]; r.lm[i] = (i < C) ? mu * r.lm[i] - gfy.lm[i] : 0.0; r.pi[i] = -re_pi[i]; }
      LUNR for (int a = 0; a < 2; a++) { r.us[a] = mu * r.us[a] - gfy.us[a]; r.nu[a] = -re_nu[a]; }
      r.dl = mu * wd - gfy.dl; r.eta = -re_eta;
    }
    // ---- factorisation with inertia correction (Ipopt algorithm IC)
    double dw = 0.0;
    {
      const double dw_last = PERS_DWLAST;
      bool ok = false;
      bool need = true;  // this group wants (another) factorisation; one that does not reproduces its factors (same H, same dw)
      for (;;) {
        if (need) info.n_fact++;
        const bool okf = l_factor<C, false>(g, P, c, H, dw, f);
        if (need) {
          ok = okf;
          need = !ok;
          if (need) {
            if (dw == 0.0) dw = (dw_last == 0.0) ? P.dw_first : fmax(P.dw_min, P.kw_minus * dw_last);
            else dw *= (first_ic && dw_last == 0.0) ? P.kw_plus_first : P.kw_plus;
            if (dw > P.dw_max) need = false;
          }
        }
        if (!w_any(g, need)) break;
      }
      g_keep(g, ok);
      if (!ok) { info.status = 3; return true; }
      if (dw > 0.0) { PERS_DWLAST = dw; first_ic = false; }
    }
    // ---- solve, then iterative refinement on the UNREDUCED Newton system (one instance of the solve for both).  A row whose
    //      Sigma is ~1e15 and over which the stage's own controls have little authority (a diagonal-pair stance has one control
    //      direction) hands its stiffness on to the previous stage, where the recursion then loses digits: measured residual 5e-9
    //      of the right-hand side at Sigma ~ 1e11, 1e-6 .. 1 at 1e13 .. 1e15; below that the recursion is accurate to round-off,
    //      so the residual is evaluated only when some row of the group is above LANE_QD_CHECK.
    KV dd; double tq = 0.0;
    {
      double rq = 0.0;
      bool refining = qd_mine;  // this group checks its residual and takes the correction of the next solve
      for (int rf = 0;; rf++) {
        KV d2; double t2;
        l_kkt_solve<C, true>(g, P, c, H, f, dw, r, rq, d2, t2);
        if (rf == 0) { dd = d2; tq = t2; }
        else if (refining) {
          LUNR for (int i = 0; i < 4; i++) { dd.xs[i] += d2.xs[i]; dd.lm[i] += d2.lm[i]; dd.pi[i] += d2.pi[i]; }
          LUNR for (int a = 0; a < 2; a++) { dd.us[a] += d2.us[a]; dd.nu[a] += d2.nu[a]; }
          dd.dl += d2.dl; dd.eta += d2.eta; tq += t2;
        }
        double ddp[4]; prev_state(dd, true, ddp);
        nlp_scatter_de(d::N_C, dd, ddp);
        // pass B: multiplier steps dz_lo -> D, dz_hi -> E from the slot steps de (C); J'(z + dz) weights -> G for the residual
        LSLOTS for (int s = 0; s < nsn; s++) {
          const bool vhi = s != d::SD;
          const double slo = pr(d::N_SLO + s), shi = pr(d::N_SHI + s), zlo = pr(d::N_ZLO + s), zhi = pr(d::N_ZHI + s);
          const double rlo = pr(d::N_A + s), rhi = pr(d::N_B + s), de = pr(d::N_C + s);
          double dzl = (mu - slo * zlo) * rlo - zlo * rlo * de;
          const double dzh = vhi ? (mu - shi * zhi) * rhi + zhi * rhi * de : 0.0;
          if (s == d::SD && H.capped) dzl = (mu - slo * zlo) * rlo - H.sc * de - tq;
          pr(d::N_D + s) = dzl; pr(d::N_E + s) = dzh;
          if (any_qd) pr(d::N_G + s) = (zlo + dzl) - (zhi + dzh);
        }
        if (rf >= P.max_refine || !any_qd) break;
        // rho1 = -(grad f + A'(y+dy) + Htilde dx - J'(z+dz)),  rho2 = -(A (x+dx) - b),  rho_q = t / (Sigma - cap) - g'dx
        // (grad f + A'y and the equality residual of the iterate are recomputed here, rarely, instead of living across the solve)
        const double rqn = H.capped ? H.srinv * tq - pr(d::N_C + d::SD) : 0.0;
        KV gd, jz2, gf; double wd2, dpin[4], rp_[4], rn_[2], re_;
        {
          double xp0[4], pin0[4];
          prev_state(x, false, xp0); next_pi(x, pin0);
          l_grad_Aty<C>(P, c, pr, x, x, pin0, true, g.j, gf);
          l_eq_residual<C>(P, c, x, xp0, false, rp_, rn_, re_);
        }
        next_pi(dd, dpin);
        l_grad_Aty<C>(P, c, pr, dd, dd, dpin, false, g.j, gd);
        l_gather<C, true>(g, pr, d::N_G, H.gv, jz2, wd2);
        double nres = fabs(rqn);
        LUNR for (int i = 0; i < 4; i++) {
          const double ht = c.qw + dw - ((i & 1) ? 0.0 : 2.0 * zd);
          r.xs[i] = -(gf.xs[i] + gd.xs[i] + ht * dd.xs[i] - jz2.xs[i]);
          nres = fmax(nres, fabs(r.xs[i]));
        }
        LUNR for (int a = 0; a < 2; a++) { r.us[a] = -(gf.us[a] + gd.us[a] + (P.Rw + dw) * dd.us[a] - jz2.us[a]); nres = fmax(nres, fabs(r.us[a])); }
        LUNR for (int i = 0; i < 4; i++) {
          r.lm[i] = (i < C) ? -(gf.lm[i] + gd.lm[i] + (dw + P.rho_lambda) * dd.lm[i] - jz2.lm[i]) : 0.0;
          nres = fmax(nres, fabs(r.lm[i]));
        }
        r.dl = -(gf.dl + ((g.j == 3) ? (P.slw + dw) * dd.dl : 0.0) - wd2);   // this lane's share of the slack variable's row
        nres = fmax(nres, fabs(g_sum(g, r.dl)));
        {
          double hp[4], hn[2], he;
          l_eq_residual<C>(P, c, dd, ddp, true, hp, hn, he);
          LUNR for (int i = 0; i < 4; i++) { r.pi[i] = -(rp_[i] + hp[i]); nres = fmax(nres, fabs(r.pi[i])); }
          LUNR for (int a = 0; a < 2; a++) { r.nu[a] = -(rn_[a] + hn[a]); nres = fmax(nres, fabs(r.nu[a])); }
          r.eta = -(re_ + he); nres = fmax(nres, fabs(r.eta));
        }
        nres = g_max(g, nres);
        const double nrhs = fmax(e_d, e_p);
        LANE_DBG("   refine %d nres %.2e nrhs %.2e\n", rf, nres, nrhs);
        refining = refining && nres > fmax(P.refine_rel * nrhs, P.refine_abs);
        if (!w_any(g, refining)) break;
        rq = rqn;
      }
    }
    const double quad = dd.xs[0] * dd.xs[0] + dd.xs[2] * dd.xs[2];
    // ---- pass C: fraction to the boundary (alpha_max = min(1, tau / max_i(-ds_i / s_i))) and the line-search constants
    //      phi = f - mu sum log s, its directional derivative, sum of slack products (for log)
    double a_pmax, a_dmax, ph, dphi, dQd, gdx;
    {
      double rp = 0.0, rdm = 0.0, prod = 1.0, prod2 = 1.0, dsum = 0.0;
#ifdef CBF_DBGBUF
      double dbg_zq = -1e300, dbg_de = 0, dbg_slo = 0, dbg_shi = 0, dbg_zlo = 0, dbg_zhi = 0; int dbg_slot = -1;
#endif
      LSLOTS for (int s = 0; s < nsn; s++) {
        const bool vhi = s != d::SD;
        const double de = pr(d::N_C + s), rlo = pr(d::N_A + s), rhi = pr(d::N_B + s);
        rp = fmax(rp, -de * rlo);
        if (vhi) rp = fmax(rp, de * rhi);
        double zq = -pr(d::N_D + s) * l_rcp(pr(d::N_ZLO + s));
        if (vhi) zq = fmax(zq, -pr(d::N_E + s) * l_rcp(pr(d::N_ZHI + s)));
        rdm = fmax(rdm, zq);
#ifdef CBF_DBGBUF
        if (zq > dbg_zq) { dbg_zq = zq; dbg_slot = s; dbg_de = de; dbg_slo = pr(d::N_SLO + s); dbg_shi = pr(d::N_SHI + s); dbg_zlo = pr(d::N_ZLO + s); dbg_zhi = pr(d::N_ZHI + s); }
#endif
        const double ss = pr(d::N_SLO + s) * pr(d::N_SHI + s);
        if (s & 1) prod *= ss; else prod2 *= ss;
        dsum += de * rlo - (vhi ? de * rhi : 0.0);
      }
      rp = g_max(g, rp); rdm = g_max(g, rdm);
      const double tau = PERS_TAU;
      a_pmax = (rp > tau) ? tau / rp : 1.0; a_dmax = (rdm > tau) ? tau / rdm : 1.0;
      LANE_LOG(dw); LANE_LOG(a_pmax); LANE_LOG(a_dmax);
#ifdef CBF_DBGBUF
      for (int k = 0; k < 4; k++) {   // per lane: the slot that limits the dual step
        const double v0 = g_idx(g, dbg_zq, k), v1 = g_idx(g, (double)dbg_slot, k), v2 = g_idx(g, dbg_de, k), v3 = g_idx(g, dbg_slo, k), v4 = g_idx(g, dbg_shi, k), v5 = g_idx(g, dbg_zlo, k), v6 = g_idx(g, dbg_zhi, k);
        LANE_LOG(-2000.0 - k); LANE_LOG(v0); LANE_LOG(v1); LANE_LOG(v2); LANE_LOG(v3); LANE_LOG(v4); LANE_LOG(v5); LANE_LOG(v6);
      }
#endif
      LANE_DBG("   dw %.1e apmax %.3e admax %.3e ddl %.3e\n", dw, a_pmax, a_dmax, dd.dl);
      // f(x) = sum_j x_j (1/2 Q_jj x_j + f_j)   (DV:423-428)
      double f0 = 0.0, gd = 0.0, qd = 0.0;
      LUNR for (int i = 0; i < 4; i++) {
        const double xv = x.xs[i], dv = dd.xs[i], fv = -c.qw * pr(d::F_XDES + i);
        f0 += xv * (0.5 * c.qw * xv + fv); gd += (c.qw * xv + fv) * dv; qd += c.qw * dv * dv;
      }
      LUNR for (int a = 0; a < 2; a++) { const double xv = x.us[a], dv = dd.us[a]; f0 += xv * (0.5 * P.Rw * xv); gd += P.Rw * xv * dv; qd += P.Rw * dv * dv; }
      if (g.j == 0) { const double xv = x.dl, dv = dd.dl; f0 += xv * (0.5 * P.slw * xv); gd += P.slw * xv * dv; qd += P.slw * dv * dv; }
      ph = g_sum(g, f0 - mu * (log(prod) + log(prod2)));
      gdx = g_sum(g, gd);
      dphi = gdx - mu * g_sum(g, dsum);
      dQd = g_sum(g, qd);
    }
    // ---- filter line search (theta = ||A x - b||_1 shrinks by (1 - alpha)); pass D per trial
    double alpha = a_pmax;
    bool accepted = false, ftype = false;
    bool going = alpha >= P.alpha_min;   // this group still looks for its step length
    while (w_any(g, going)) {
      // phi(x + alpha dx) - phi(x) = alpha g'dx + 1/2 alpha^2 dx'Q dx - mu sum log(s_t / s)
      double prod = 1.0, prod2 = 1.0;
      bool pos = true;
      double smin = 1.0;
      const double aq = alpha * alpha * quad;
      LSLOTS for (int s = 0; s < nsn; s++) {
        const bool vhi = s != d::SD;
        const double et = alpha * pr(d::N_C + s) + ((s == d::SD) ? aq : 0.0);
        const double sl = pr(d::N_SLO + s) + et, sh = vhi ? pr(d::N_SHI + s) - et : 1.0;
        pos = pos && (sl > 0.0) && (sh > 0.0);
        smin = fmin(smin, fmin(sl, sh));
        const double t = (sl * pr(d::N_A + s)) * (vhi ? sh * pr(d::N_B + s) : 1.0);
        if (s & 1) prod *= t; else prod2 *= t;
      }
      const bool allpos = g_all(g, pos);
      const double dlog = g_sum(g, allpos ? log(prod) + log(prod2) : 0.0);
      const bool anysmall = g_any(g, smin < 1e-12);
      {
        const double tht = (1.0 - alpha) * th;
        const double dph = alpha * gdx + 0.5 * alpha * alpha * dQd - mu * dlog;
        const double pht = ph + dph;
        const bool fok = filter_ok(tht, pht);   // (a vote: evaluated by every group, used when the trial point is interior)
        if (going && allpos && tht <= PERS_THMAX && fok) {
          const bool sw = dphi < 0.0 && alpha * (dphi * dphi) > th;  // switching condition, s_phi = 2, s_theta = 1, delta = 1
          if (th <= PERS_THMIN && sw) {
            if (dph - 10.0 * CBF_EPSM * fabs(ph) <= P.eta_phi * alpha * dphi) { accepted = true; ftype = true; }
          } else if (tht <= (1.0 - P.gamma_theta) * th || dph - 10.0 * CBF_EPSM * fabs(ph) <= -P.gamma_phi * th) {
            accepted = true; ftype = false;
          }
        }
      }
      const bool recheck = going && accepted && anysmall;
      if (w_any(g, recheck)) {
        // The test above works on the linearised slacks s + alpha ds.  Late in the run tau = 1 - mu lets an active slack shrink
        // by a factor mu in one step, to below the rounding (~1e-15) of the O(1) positions it is computed from.  When a trial
        // slack is that small, the slacks that pass A will compute at x + alpha dx -- by the same expressions, from the same
        // bits -- must be positive too (the oracle's line search evaluates c(x + alpha dx) itself); otherwise the step is halved.
        KV xt;
        LUNR for (int i = 0; i < 4; i++) { xt.xs[i] = x.xs[i] + alpha * dd.xs[i]; xt.lm[i] = (i < C) ? x.lm[i] + alpha * dd.lm[i] : 0.0; }
        LUNR for (int a = 0; a < 2; a++) xt.us[a] = x.us[a] + alpha * dd.us[a];
        xt.dl = x.dl + alpha * dd.dl;
        double xp2[4]; prev_state(xt, false, xp2);
        l_scatter<C, true>(g, pr, d::N_EV, xt, xp2);
        { const double a_ = xt.xs[0] - c.ox, b_ = xt.xs[2] - c.oy; pr(d::N_EV + d::SD) = a_ * a_ + b_ * b_ + xt.dl; }
        bool okp = true;
        LSLOTS for (int s = 0; s < nsn; s++) {
          const double e = pr(d::N_EV + s);
          okp = okp && (e - l_slot_lo<C, true>(P, s) > 0.0) && ((s != d::SD) ? (l_slot_hi<C, true>(P, s) - e > 0.0) : true);
        }
        const bool okall = g_all(g, okp);
        if (recheck) accepted = okall;
      }
      if (going && !accepted) alpha *= 0.5;
      going = going && !accepted && alpha >= P.alpha_min;
    }
    LANE_LOG(alpha); LANE_LOG(ph); LANE_LOG(dphi);
    LANE_DBG("   alpha %.3e acc %d ftype %d ph %.10e dphi %.3e\n", alpha, (int)accepted, (int)ftype, ph, dphi);
    g_keep(g, accepted);
    if (!accepted) { info.status = 2; return true; }
    if (!ftype) filter_add(P, (1.0 - P.gamma_theta) * th, ph - P.gamma_phi * th);
    // ---- update (the multiplier step z += a_dmax dz is applied by pass A of the next iteration)
    LUNR for (int i = 0; i < 4; i++) { x.xs[i] += alpha * dd.xs[i]; x.pi[i] += alpha * dd.pi[i]; }
    LUNR for (int i = 0; i < C; i++) x.lm[i] += alpha * dd.lm[i];
    LUNR for (int a = 0; a < 2; a++) { x.us[a] += alpha * dd.us[a]; x.nu[a] += alpha * dd.nu[a]; }
    x.dl += alpha * dd.dl; x.eta += alpha * dd.eta;
    PERS_ADPREV = a_dmax;
    nl_it = it + 1;
    return false;
  }
};

// ===============================================================================================================
// One group = one run_NMPC(): load (updateState + planners' outputs), closest obstacle, QP stage, NLP stage,
// Bezier epilogue, store.  Shared by the CUDA kernel and the tests/emu harness.
// ===============================================================================================================
// LATE: the obstacles are looked at AFTER the QP stage (which does not use them) and only once io.obs_flag shows io.obs_seq -- the
// host-buffer call uploads them on a second stream while the QP stage runs.  Same arithmetic, same answers.
template <int C, bool LATE>
LDEV bool lane_closest_obstacle(const BatchIO& io, const Grp& g, const Priv& pr, int b, double& ox, double& oy) {
  typedef LaneDims<C> d;
  const size_t B = (size_t)io.ld;
  const int j = g.j;
  if (LATE && io.obs_flag) {
#ifndef CBF_HOST_EMU
    long long t0 = clock64();
    for (;;) {
      int v; asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(io.obs_flag) : "memory");
      if (v == io.obs_seq || clock64() - t0 > 20000000000ll) break;   // (10 s: the copy was never issued; the answers will say so)
      __nanosleep(100);
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
  }
  // closest obstacle to the current position: first strict minimum of the Euclidean distance (MPC_dist.cpp:371-396)
  const double px = pr(d::F_XIN + 0), py = pr(d::F_XIN + 2);
  double best = 1000.0; int bi = 0x7fffffff;
  for (int k = j; k < io.n_obs; k += 4) {
    const double dx = px - io.obs[(size_t)k * B + b], dy = py - io.obs[(size_t)(io.n_obs + k) * B + b];
    const double dd = sqrt(l_add_rn(l_mul_rn(dx, dx), l_mul_rn(dy, dy)));
    if (dd < best) { best = dd; bi = k; }
  }
  LUNR for (int m = 1; m < 4; m <<= 1) {
    const double ob = g_xor(g, best, m); const int oi = (int)g_xor_u(g, (unsigned)bi, m);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (bi == 0x7fffffff) bi = 0;
  ox = io.obs[(size_t)bi * B + b]; oy = io.obs[(size_t)(io.n_obs + bi) * B + b];
  if (!LATE) return true;   // (the caller's own finiteness check covers them)
  return g_all(g, 0.0 * ox + 0.0 * oy == 0.0);
}

template <int C, bool LATE = false>
LDEV void lane_run_group(const KParams& P, const BatchIO& io, int b, bool valid, Priv pr, Pace pace = Pace()) {
  typedef LaneDims<C> d;
  constexpr int NW = d::NW;
  LaneSolver<C> S;
  S.g = grp_make(); S.pr = pr;
  S.nsq = l_nslots<C, false>(S.g); S.nsn = l_nslots<C, true>(S.g);
  Grp& g = S.g;   // g.mask follows the mask discipline of lane_prims.h: set here at every point where all 32 lanes are present
  const int j = g.j;
  const size_t B = (size_t)io.ld;  // array pitch
  S.info.status = 1; S.info.qp_iters = 0; S.info.qp_flag = 3; S.info.nlp_iters = 0; S.info.n_fact = 0; S.info.kkt = 0.0;
  S.perm = 0xE4;   // position i holds foot i
  double abuf_j = 0.0;
  bool qp_done = true, badin = false;
  g.mask = w_ballot_full(valid);
  if (valid) {
    // ---- load (SoA, one instance; lane j takes stage j)
    LUNR for (int i = 0; i < 4; i++) { pr(d::F_XDES + i) = io.xdes[(size_t)(4 * j + i) * B + b]; pr(d::F_XIN + i) = io.x0[(size_t)i * B + b]; }
    abuf_j = io.abuf ? io.abuf[(size_t)j * B + b] : 0.0;
    double feet[8];
    LUNR for (int t = 0; t < 8; t++) feet[t] = io.feet[(size_t)t * B + b];
    const int c0 = io.contact[b];
    // stance matrix (MPC_dist.cpp:1238-1252)
    double F[8];
    LUNR for (int a = 0; a < 2; a++) {
      if (C == 4) { LUNR for (int i = 0; i < 4; i++) F[4 * a + i] = feet[4 * a + i]; }
      else {
        F[4 * a] = (c0 == 0) ? feet[4 * a + 1] : feet[4 * a]; F[4 * a + 1] = (c0 == 0) ? feet[4 * a + 2] : feet[4 * a + 3];
        F[4 * a + 2] = 0.0; F[4 * a + 3] = 0.0;
      }
      S.c.Fl[a] = F[4 * a + C - 1];
    }
    LUNR for (int t = 0; t < 6; t++) S.c.D[t] = 0.0;
    LUNR for (int k = 0; k < 2; k++) LUNR for (int a = 0; a < NW; a++) S.c.D[k * NW + a] = F[4 * k + a] - F[4 * k + C - 1];
    S.c.qw = (j == 3) ? P.Pw : P.Qw;
    S.c.ox = 0.0; S.c.oy = 0.0;
    if (!LATE && io.mode == 1) (void)lane_closest_obstacle<C, false>(io, g, pr, b, S.c.ox, S.c.oy);
    // non-finite data (0 * v is 0 only for finite v) are answered with status 4 before anything is solved
    double chk = 0.0 * S.c.ox + 0.0 * S.c.oy;
    LUNR for (int i = 0; i < 4; i++) chk += 0.0 * pr(d::F_XDES + i) + 0.0 * pr(d::F_XIN + i);
    LUNR for (int t = 0; t < 8; t++) chk += 0.0 * F[t];
    badin = !g_all(g, chk == 0.0);
    const unsigned sv = g_enter(g, !badin);
    if (badin) {
      S.info.status = 4;
    } else if (io.xstart == nullptr) {
      qp_done = S.qp_init(P) || P.qp_maxit <= 0;
    } else {
      LUNR for (int i = 0; i < 4; i++) { S.x.xs[i] = io.xstart[(size_t)(4 * j + i) * B + b]; S.x.lm[i] = (i < C) ? io.xstart[(size_t)(24 + C * j + i) * B + b] : 0.0; S.x.pi[i] = 0.0; }
      LUNR for (int a = 0; a < 2; a++) { S.x.us[a] = io.xstart[(size_t)(16 + 2 * j + a) * B + b]; S.x.nu[a] = 0.0; }
      S.x.dl = io.xstart[(size_t)(24 + 4 * C) * B + b]; S.x.eta = 0.0;
      S.info.qp_flag = 0;
    }
    g_leave(g, sv);
  }
  // ---- QP stage: the groups of a CTA iterate together (pace_go) until the last one has finished
  for (;;) {
    const bool act = valid && !qp_done;
    g.mask = w_ballot_full(act);
    if (!pace_go(pace, g.mask != 0u)) break;
    if (act) qp_done = S.qp_iter(P);
  }
  bool nl_done = true;
  g.mask = w_ballot_full(valid && !badin);
  if (LATE && io.mode == 1) {   // the obstacles now (they were on their way during the QP stage); non-finite ones are bad input
    bool late_bad = false;
    if (valid && !badin) late_bad = !lane_closest_obstacle<C, true>(io, g, pr, b, S.c.ox, S.c.oy);
    if (late_bad) { badin = true; S.info.status = 4; }
    g.mask = w_ballot_full(valid && !badin);
  }
  if (valid && !badin) {
    if (io.xqp) {
      LUNR for (int i = 0; i < 4; i++) io.xqp[(size_t)(4 * j + i) * B + b] = S.x.xs[i];
      LUNR for (int a = 0; a < 2; a++) io.xqp[(size_t)(16 + 2 * j + a) * B + b] = S.x.us[a];
    }
    if (io.mode == 1) nl_done = S.nlp_init(P);
    else S.qp_final_kkt(P);
  }
  // ---- NLP stage
  for (;;) {
    const bool act = valid && !nl_done;
    g.mask = w_ballot_full(act);
    if (!pace_go(pace, g.mask != 0u)) break;
    if (act) nl_done = S.nlp_iter(P);
  }
  g.mask = w_ballot_full(valid && !badin);
  if (!valid) return;
  if (badin) {  // non-finite inputs: nothing was solved; only the status fields are written (include/cbfnmpc.h), like lane_store_bad
    if (j == 0) {
      if (io.status) io.status[b] = 4;
      if (io.kkt) io.kkt[b] = 0.0;
      if (io.qp_iters) io.qp_iters[b] = 0;
      if (io.nlp_iters) io.nlp_iters[b] = 0;
      if (io.n_fact) io.n_fact[b] = 0;
      if (io.slack) io.slack[b] = 0.0;
      if (io.active) { io.active[b] = 0ull; io.active[B + b] = 0ull; }
    }
    return;
  }
  // ---- store
  if (io.X) { LUNR for (int i = 0; i < 4; i++) io.X[(size_t)(4 * j + i) * B + b] = S.x.xs[i]; }
  if (io.U) { LUNR for (int a = 0; a < 2; a++) io.U[(size_t)(2 * j + a) * B + b] = S.x.us[a]; }
  if (io.lam) {
    LUNR for (int i = 0; i < C; i++) io.lam[(size_t)(C * j + ((S.perm >> (2 * i)) & 3)) * B + b] = S.x.lm[i];
    LUNR for (int i = 0; i < 4 - C; i++) io.lam[(size_t)(4 * C + (4 - C) * j + i) * B + b] = 0.0;
  }
  // Bezier fit  alpha = Y Minv'   (MPC_dist.cpp:784-855, effective semantics): lane r fits state row r
  {
    double Y[5];
    Y[0] = abuf_j;
    LUNR for (int st = 0; st < 4; st++) {
      double v = 0.0;
      LUNR for (int t = 0; t < 4; t++) { const double w = g_idx(g, S.x.xs[t], st); if (t == j) v = w; }
      Y[1 + st] = v;
    }
    LUNR for (int jj = 0; jj < 5; jj++) {
      double s = 0.0;
      LUNR for (int i = 0; i < 5; i++) s += P.minv[5 * jj + i] * Y[i];
      if (io.alpha) io.alpha[(size_t)(5 * j + jj) * B + b] = s;
    }
    if (io.abuf_next) io.abuf_next[(size_t)j * B + b] = Y[4];
  }
  // active-set mask in the reference's row numbering (DESIGN.md row map): bit set <=> z > s
  unsigned w[4] = {0u, 0u, 0u, 0u};
  {
    const bool nlp = io.mode == 1;
    const int st = nlp ? d::NS : d::NSQP;  // field stride, and slot order, of the stage that ran last
    const int SF = nlp ? d::SFN : d::SFQ;
    LUNR for (int s = 0; s < d::NS; s++) {
      if (s < st && !(j == 0 && s >= SF && s < SF + 2)) {
        int rlo = -1, rhi = -1;
        if (s < d::SU) { rhi = 12 + 4 * j + s; rlo = 28 + 4 * j + s; }
        else if (s < d::SL) { rhi = 44 + 2 * j + (s - d::SU); rlo = 52 + 2 * j + (s - d::SU); }
        else if (s < d::SL + C) { const int ft = (S.perm >> (2 * (s - d::SL))) & 3; rlo = 60 + C * j + ft; rhi = 60 + 4 * C + C * j + ft; }
        else if (s >= SF) { rhi = 2 * (j - 1) + (s - SF); rlo = 6 + rhi; }
        else if (s < d::SD) { const int q = (s - d::SV) ? 4 + j : j; rhi = d::MQP + 4 + q; rlo = d::MQP + 12 + q; }
        else { rlo = d::MQP + j; }
        if (rlo >= 0 && pr(2 * st + s) > pr(s)) w[rlo >> 5] |= 1u << (rlo & 31);
        if (rhi >= 0 && !(nlp && s == d::SD) && pr(3 * st + s) > pr(st + s)) w[rhi >> 5] |= 1u << (rhi & 31);
      }
    }
    LUNR for (int k = 0; k < 4; k++) w[k] = g_or(g, w[k]);
  }
  if (j == 0) {
    if (io.slack) io.slack[b] = S.x.dl;
    if (io.kkt) io.kkt[b] = S.info.kkt;
    if (io.status) io.status[b] = S.info.status;
    if (io.qp_iters) io.qp_iters[b] = S.info.qp_iters;
    if (io.nlp_iters) io.nlp_iters[b] = S.info.nlp_iters;
    if (io.n_fact) io.n_fact[b] = S.info.n_fact;
    if (io.active) {
      io.active[b] = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
      io.active[B + b] = (unsigned long long)w[2] | ((unsigned long long)w[3] << 32);
    }
  }
}

// contact pattern -> number of stance feet: 4, 2 (a diagonal pair) or -1 (MPC_dist.cpp:1238-1252)
LDEV int lane_classify(const BatchIO& io, int b) {
  const size_t B = (size_t)io.ld;  // array pitch
  const int c0 = io.contact[b], c1 = io.contact[B + b], c2 = io.contact[2 * B + b], c3 = io.contact[3 * B + b];
  const int csum = c0 + c1 + c2 + c3;
  if (csum == 4) return 4;
  if (csum == 2 && (c0 == 0 || c0 == 1)) return 2;
  return -1;
}
LDEV void lane_store_bad(const BatchIO& io, int b) {
  const size_t B = (size_t)io.ld;  // array pitch
  if (io.status) io.status[b] = 4;
  if (io.kkt) io.kkt[b] = 0.0;
  if (io.qp_iters) io.qp_iters[b] = 0;
  if (io.nlp_iters) io.nlp_iters[b] = 0;
  if (io.n_fact) io.n_fact[b] = 0;
  if (io.slack) io.slack[b] = 0.0;
  if (io.active) { io.active[b] = 0ull; io.active[B + b] = 0ull; }
}
