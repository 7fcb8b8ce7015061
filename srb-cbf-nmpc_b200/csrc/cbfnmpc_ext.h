// cbfnmpc_ext.h -- EXTENDED-MODE CBF-NMPC solver (BASELINE configs 4 and 5).
//
// The reference's NLP (src/MPC_dist.cpp:81-454, include/dec_vars_constr_cost.h:245-438) generalised along the reference's own
// hints -- NO REFERENCE IMPLEMENTATION of this mode exists; its checker is oracle/cbfnmpc_oracle_ext.c:
//   * N = 4 or 8 grid points (N / 4 gait domains, one stance matrix per domain),
//   * K_obs obstacle rows per grid point  |p_k - o_i|^2 + delta >= eps_obs     (the reference's row DV:262-265, per obstacle),
//   * n_nbr neighbour rows per grid point |p_k - p^j_k|^2 + delta >= eps_nbr   (DV:402 "epsilon = 2.2 // robot to robot"),
//   * team mode (IPM-F-J): the T members of a team iterate together and read each other's published trajectories at the top of
//     every interior-point iteration (under-relaxed Jacobi) -- through warp shuffles when the team sits in one warp, through
//     global arrays (filled by an NCCL all-gather, cbfnmpc_api.cu) when its members live on different GPUs.
// With N = 4, K_obs = 1, n_nbr = 0 it solves the reference-faithful problem of cbfnmpc_lane.h.
//
// One instance = W = N adjacent lanes (lane j owns grid point j: x_{j+1}, u_j, lambda_j, their multipliers and rows).
// Linear algebra (differs from cbfnmpc_lane.h): the slack variable delta rides along as a FIFTH STATE (delta_{j+1} = delta_j,
// delta_0 free), so every distance row is an ordinary row of its own grid point and K rows per grid point cost no border.  Rows
// whose barrier weight Sigma = z / s exceeds sigma_cap (active rows late in the run: Sigma ~ 1e16) are condensed only up to the
// cap; the remainder stays in the stage as a quasi-definite 2 x 2 (at most XKC rows per grid point) that is eliminated AFTER the
// stage's controls: T = 1 / (Sigma - cap) + M R~^-1 M' -- all numbers stay O(curvature), nothing of size 1e16 enters the
// Riccati recursion.
#pragma once
#include "cbfnmpc_types.h"
#include "ext_prims.h"

#define XNT 32   // threads per CTA = one warp = stride of the thread-private shared-memory columns
#ifndef XKC
#define XKC 1    // rows per grid point kept in quasi-definite form when their Sigma exceeds the cap (1 or 2).  A second stiff row of the same
                 // grid point is condensed with its full Sigma; the accuracy that costs is restored by the iterative refinement.
#endif
#define XUNR _Pragma("unroll")
#ifndef XSLOT_UNROLL
#define XSLOT_UNROLL 2
#endif
#define XPRAGMA_(x) _Pragma(#x)
#define XPRAGMA(x) XPRAGMA_(x)
#define XSLOTS XPRAGMA(unroll XSLOT_UNROLL)

struct XParams {
  KParams k;
  double eps_obs, eps_nbr, relax, pub_tol;
  int mu_stall;      // team mode: iterations without a decrease after which the barrier parameter is lowered anyway
  int max_team_it;   // iterations of a team's loop (a team that has not settled by then keeps the statuses its members have)
};

struct ExtIO {
  // inputs, SoA [field][ld]
  const double *x0, *xdes, *feet; const int* contact; const double* obs; const double* abuf;
  const double* nbr[3];  // xmode 2: published trajectory of neighbour i, [2N + 2][ld] (rows 2k, 2k+1 = x, y at grid point k; then flags)
  double* pub;           // xmode 2: this instance's published trajectory [2N + 2][ld]
  // outputs
  double *X, *U, *lam, *slack, *xqp, *alpha, *abuf_next, *kkt;
  int *status, *qp_iters, *nlp_iters, *team_iters;
  // resumable runs (xmode 2 with one launch per iteration): state[XSTATE][ld * W]
  double* state;
  int* n_active;         // incremented by every team that still iterates when the launch ends
  int B, ld, N, K_obs, n_nbr, T;
  int xmode;             // 0: no neighbours; 1: team in one warp (shuffles); 2: fixed neighbour trajectories in io.nbr;
                         // 3: members of a team in different launches / on different GPUs: published trajectories of ALL members
                         //    in io.gathered, [member][2N + 2][n_teams] (filled by an all-gather between the launches); this launch
                         //    holds members member0 .. of every team, member-major: instance b = (member - member0) * n_teams + team
  const double* gathered;
  int n_teams, member0;
  // xmode 4 (cbfnmpc_ext_solve_fused): as xmode 3, but ONE launch for the whole solve -- every member writes its published block
  // straight into the mailboxes of all ranks of its team's group (peer memory over NVLink: mbox[r] is rank r's mailbox mapped into
  // this process; own rank included) and waits on per-(member, team) sequence flags in its OWN mailbox: no launch and no collective
  // between the iterations.  Mailbox: double [2 parities][T members][2N + 2 rows][n_teams], then int flags [2][T][n_teams].
  double* mbox[8];
  int n_grp;             // ranks of the group (<= 8)
  int slice_t0, slice_ns;   // this launch runs teams [slice_t0, slice_t0 + slice_ns) only (0, 0: all): a launch must be resident at once,
                            // so a batch beyond that is solved in slices; item i of the class lists is instance (i / ns) n_teams + t0 + i % ns
  int* fused_err;        // set when a wait ran into the watchdog (a peer that never published)
  int* fused_its;        // atomicMax of the team iteration counts
  int phase;             // 0: whole solve in one launch.  Resumable (xmode 2, an all-gather between launches): 1: QP stage, publish, save;
                         // 2: NLP start, publish, save; 3: one iteration, publish, save; 4: store the outputs
  int n_run;
};

// thread-private shared-memory column
struct XPriv {
  double* base;
  XMEM double& operator()(int f) const { return base[f * XNT]; }
};

template <int C>
struct XDims {
  static constexpr int NW = C - 1;
  // two-sided slots of one lane.  QP stage: X 4 | U 2 | lambda C | friction 2;  NLP stage: X 4 | U 2 | lambda C | velocity 2 | friction 2.
  // The friction slots come last: lane 0 has none (x_0 is fixed) and simply runs shorter loops.
  static constexpr int SX = 0, SU = 4, SL = 6, SFQ = 6 + C, NSQ = 8 + C, SV = 6 + C, SFN = 8 + C, NS = 10 + C;
  static constexpr int Q_SLO = 0, Q_SHI = NSQ, Q_ZLO = 2 * NSQ, Q_ZHI = 3 * NSQ, Q_A = 4 * NSQ, Q_B = 5 * NSQ, Q_C = 6 * NSQ, Q_D = 7 * NSQ,
                       Q_E = 8 * NSQ, Q_G = 9 * NSQ, Q_EV = 10 * NSQ, Q_DE = 11 * NSQ, Q_END = 12 * NSQ;
  static constexpr int N_SLO = 0, N_SHI = NS, N_ZLO = 2 * NS, N_ZHI = 3 * NS, N_A = 4 * NS, N_B = 5 * NS, N_C = 6 * NS, N_D = 7 * NS,
                       N_E = 8 * NS, N_G = 9 * NS, N_END = 10 * NS;
  static constexpr int F_XDES = Q_END > N_END ? Q_END : N_END;  // desired state of this lane's grid point (4)
  static constexpr int F_XIN = F_XDES + 4;      // measured state x_0 (4)
  static constexpr int F_FILT = F_XIN + 4;      // this lane's share of the filter (2 pairs)
  static constexpr int F_PERS = F_FILT + 4;     // tau, dw_last, theta_max, theta_min, ad_prev, pub_x, pub_y, moved
  static constexpr int F_ROWS = F_PERS + 8;     // K distance rows x XROW fields
  static constexpr int NFIX = F_ROWS;
};
// per distance row: centre (2), slack, 1/slack, multiplier, e(dx), dz, slack of the previous iteration
#define XJ_SFLOOR 1e-12
#define XQD_CHECK 1e11   // largest Sigma of a lane's rows above which the Newton residual is checked (iterative refinement)
#define XROW 8
#define XR_CX 0
#define XR_CY 1
#define XR_S 2
#define XR_RS 3
#define XR_Z 4
#define XR_DE 5
#define XR_DZ 6
#define XR_SP 7
#define XPERS_TAU pr(d::F_PERS + 0)
#define XPERS_DWLAST pr(d::F_PERS + 1)
#define XPERS_THMAX pr(d::F_PERS + 2)
#define XPERS_THMIN pr(d::F_PERS + 3)
#define XPERS_ADPREV pr(d::F_PERS + 4)
#define XPERS_PUBX pr(d::F_PERS + 5)
#define XPERS_PUBY pr(d::F_PERS + 6)
#define XPERS_MOVED pr(d::F_PERS + 7)
#ifdef CBF_HOST_EMU
#define XHD
#else
#define XHD __host__ __device__
#endif
XHD constexpr int ext_npriv(int K) { return XDims<4>::NFIX + XROW * K; }
// registers of a lane that survive between launches of a resumable run (after the private columns)
#define XSTATE_REGS 44
XHD constexpr int ext_nstate(int K) { return ext_npriv(K) + XSTATE_REGS; }

struct XKV { double xs[4], us[2], lm[4], dl, pi[4], nu[2], eta; };
struct XConst { double Fl[2], D[6], qw; };
struct XHess {
  double hx[4], hxy, bd[2], hd;  // xi_{j+1} = (x_{j+1}, delta): x diagonal, (p_x, p_y), (p, delta), (delta, delta)
  double hu[2], hxu[2], hl[4];
};
struct XCap { double gx, gy, srinv; int row; };   // a row kept in quasi-definite form
struct XFact {
  double Pn[15];          // cost-to-go at xi_{j+1}, 5 x 5 packed lower
  double K[15];           // R~^-1 S~ : [a * 5 + m]
  double Lc[6];           // chol(R~): (1/l00, l10, 1/l11, l20, l21, 1/l22)
  double Zm[XKC][3];      // R~^-1 M_q'
  double Nt[XKC][5];      // N~_q = N_q - M_q R~^-1 S~
  double Ym[XKC][3];      // L^-1 M_q'
  double Tc[3];           // chol(T): (1/l00, l10, 1/l11)
  double p0dd;            // lane 0: cost-to-go curvature of delta_0 (all lanes hold lane 0's value after the factorisation)
  int kc;
};
struct XMain { double rx[5], rw[3], cv[4], pv[5], kv[3], tt[XKC]; };

XDEV int xsy(int i, int k) { return i >= k ? (i * (i + 1)) / 2 + k : (k * (k + 1)) / 2 + i; }
#define XCHOL_FLOOR 1e-12
#define XCHOL_NOISE 1e-9
template <bool FLOOR>
XDEV double x_pivot(double d, double diag, bool& ok) {
  ok = ok && (d > (FLOOR ? -XCHOL_NOISE * diag : 0.0)) && (d < 1e300);
  return FLOOR ? fmax(d, XCHOL_FLOOR * diag) : d;
}
template <int NW, bool FLOOR>
XDEV bool x_chol(const double* R, double* L) {  // R packed lower (00,10,11,20,21,22); pivot rule: cbfnmpc_lane.h l_chol
  bool ok = true;
  const double d0 = R[0];
  ok = ok && (d0 > 0.0) && (d0 < 1e300);
  const double i0 = x_rsq(d0);
  L[0] = i0; L[1] = 0.0; L[2] = 1.0; L[3] = 0.0; L[4] = 0.0; L[5] = 1.0;
  if (NW == 3) {
    const double l10 = R[1] * i0, l20 = R[3] * i0;
    const double d1 = x_pivot<FLOOR>(R[2] - l10 * l10, R[2], ok);
    const double i1 = x_rsq(d1);
    const double l21 = (R[4] - l20 * l10) * i1;
    const double d2 = x_pivot<FLOOR>(R[5] - l20 * l20 - l21 * l21, R[5], ok);
    const double i2 = x_rsq(d2);
    L[1] = l10; L[2] = i1; L[3] = l20; L[4] = l21; L[5] = i2;
  }
  return ok;
}
template <int NW> XDEV void x_fwd(const double* L, double* b) {
  if (NW == 1) { b[0] = b[0] * L[0]; return; }
  const double y0 = b[0] * L[0], y1 = (b[1] - L[1] * y0) * L[2], y2 = (b[2] - L[3] * y0 - L[4] * y1) * L[5];
  b[0] = y0; b[1] = y1; b[2] = y2;
}
template <int NW> XDEV void x_bwd(const double* L, double* b) {
  if (NW == 1) { b[0] = b[0] * L[0]; return; }
  const double k2 = b[2] * L[5], k1 = (b[1] - L[4] * k2) * L[2], k0 = (b[0] - L[1] * k1 - L[3] * k2) * L[0];
  b[0] = k0; b[1] = k1; b[2] = k2;
}
// T = [t00 t10; t10 t11] (kc = 2) or [t00]: Cholesky with inverse diagonals, and solve
XDEV void x_chol2(int kc, double t00, double t10, double t11, double* Tc) {
  const double i0 = x_rsq(t00);
  Tc[0] = i0; Tc[1] = 0.0; Tc[2] = 1.0;
  if (XKC == 2 && kc == 2) { const double l10 = t10 * i0; Tc[1] = l10; Tc[2] = x_rsq(t11 - l10 * l10); }
}
XDEV void x_chol2_solve(int kc, const double* Tc, double* b) {
  const double y0 = b[0] * Tc[0];
  if (XKC == 2 && kc == 2) { const double y1 = (b[1] - Tc[1] * y0) * Tc[2]; b[1] = y1 * Tc[2]; b[0] = (y0 - Tc[1] * b[1]) * Tc[0]; }
  else b[0] = y0 * Tc[0];
}
XDEV void x_Ax(const KParams& P, const double* x, double* y) {
  XUNR for (int b = 0; b < 2; b++) { y[2 * b] = P.a11 * x[2 * b] + P.a12 * x[2 * b + 1]; y[2 * b + 1] = P.a21 * x[2 * b] + P.a22 * x[2 * b + 1]; }
}
XDEV void x_ATx(const KParams& P, const double* x, double* y) {
  XUNR for (int b = 0; b < 2; b++) { y[2 * b] = P.a11 * x[2 * b] + P.a21 * x[2 * b + 1]; y[2 * b + 1] = P.a12 * x[2 * b] + P.a22 * x[2 * b + 1]; }
}

// =================================================================================================================
template <int C, int W>
struct XSolver {
  typedef XDims<C> d;
  static constexpr int NW = d::NW;
  XGrp<W> g;
  XPriv pr;
  XConst c;
  XKV x;
  XHess H;
  XFact f;
  XCap cap[XKC];
  int K, K_obs;            // distance rows per grid point (obstacles first, then neighbours)
  int nsq, nsn;
  int status, qp_iters, qp_flag, nlp_iters, n_fact;
  double kkt, mu, sigma;
  int qp_it, nl_it, nfilt, mu_age;
  bool first_ic, hold;

  XMEM double& row(int r, int fld) const { return pr(d::F_ROWS + XROW * r + fld); }
  XMEM double slot_lo(const KParams& P, bool nl, int s) const {
    return s < d::SL ? -P.box : (s < d::SL + C ? 0.0 : ((nl && s < d::SFN) ? -P.vmax : -P.fb));
  }
  XMEM double slot_hi(const KParams& P, bool nl, int s) const {
    return s < d::SL ? P.box : (s < d::SL + C ? 1.0 : ((nl && s < d::SFN) ? P.vmax : P.fb));
  }
  XMEM void prev_state(const XKV& v, bool homogeneous, double* xp) const {
    XUNR for (int i = 0; i < 4; i++) { const double t = g_up(g, v.xs[i]); xp[i] = (g.j > 0) ? t : (homogeneous ? 0.0 : pr(d::F_XIN + i)); }
  }
  XMEM void next_pi(const XKV& v, double* pin) const {
    XUNR for (int i = 0; i < 4; i++) { const double t = g_down(g, v.pi[i]); pin[i] = (g.j < W - 1) ? t : 0.0; }
  }

  // ---- residual pieces (rows of Aeq, gradient of the quadratic cost; MPC_dist.cpp:164-261)
  XMEM void eq_residual(const KParams& P, const XKV& v, const double* xp, bool homogeneous, double* rpi, double* rnu, double& reta) const {
    double ax[4];
    x_Ax(P, xp, ax);
    XUNR for (int b = 0; b < 2; b++) {
      rpi[2 * b] = ax[2 * b] - v.xs[2 * b] + P.bp * v.us[b];
      rpi[2 * b + 1] = ax[2 * b + 1] - v.xs[2 * b + 1] + P.bv * v.us[b];
    }
    double sl = 0.0;
    XUNR for (int i = 0; i < C; i++) sl += v.lm[i];
    XUNR for (int a = 0; a < 2; a++) {
      double s = v.us[a];
      XUNR for (int i = 0; i < C - 1; i++) s -= c.D[a * (C - 1) + i] * v.lm[i];
      rnu[a] = s - c.Fl[a] * sl;
    }
    reta = sl - (homogeneous ? 0.0 : 1.0);
  }
  XMEM void grad_Aty(const KParams& P, const XKV& xv, const XKV& y, const double* pin, bool with_grad, XKV& o) const {
    double at[4];
    x_ATx(P, pin, at);
    XUNR for (int i = 0; i < 4; i++) {
      double t = at[i] - y.pi[i];
      if (with_grad) t += c.qw * (xv.xs[i] - pr(d::F_XDES + i));
      o.xs[i] = t;
    }
    XUNR for (int a = 0; a < 2; a++) {
      double t = P.bp * y.pi[2 * a] + P.bv * y.pi[2 * a + 1] + y.nu[a];
      if (with_grad) t += P.Rw * xv.us[a];
      o.us[a] = t;
    }
    const double gl = -(c.Fl[0] * y.nu[0] + c.Fl[1] * y.nu[1]) + y.eta;
    XUNR for (int i = 0; i < 4; i++) o.lm[i] = 0.0;
    XUNR for (int i = 0; i < C - 1; i++) o.lm[i] = gl - (c.D[i] * y.nu[0] + c.D[(C - 1) + i] * y.nu[1]);
    o.lm[C - 1] = gl;
    o.dl = (with_grad && g.j == W - 1) ? P.slw * xv.dl : 0.0;   // the slack variable's cost sits at the last grid point's copy
  }

  // ---- slots <-> variables
  template <bool NL>
  XMEM void scatter(int fld, const XKV& v, const double* vp) const {
    XUNR for (int i = 0; i < 4; i++) pr(fld + d::SX + i) = v.xs[i];
    XUNR for (int a = 0; a < 2; a++) pr(fld + d::SU + a) = v.us[a];
    XUNR for (int i = 0; i < C; i++) pr(fld + d::SL + i) = v.lm[i];
    constexpr int SF = NL ? d::SFN : d::SFQ;
    XUNR for (int a = 0; a < 2; a++) pr(fld + SF + a) = vp[2 * a] - v.us[a];
    if (NL) { XUNR for (int a = 0; a < 2; a++) pr(fld + d::SV + a) = v.xs[2 * a + 1]; }
  }
  template <bool NL>
  XMEM void gather(int fld, XKV& acc) const {   // acc.(xs,us,lm) = J'w of the slot weights in field fld; acc.dl = 0
    constexpr int SF = NL ? d::SFN : d::SFQ;
    const double wf0 = (g.j > 0) ? pr(fld + SF) : 0.0, wf1 = (g.j > 0) ? pr(fld + SF + 1) : 0.0;
    XUNR for (int i = 0; i < 4; i++) acc.xs[i] = pr(fld + d::SX + i);
    acc.us[0] = pr(fld + d::SU) - wf0; acc.us[1] = pr(fld + d::SU + 1) - wf1;
    XUNR for (int i = 0; i < 4; i++) acc.lm[i] = (i < C) ? pr(fld + d::SL + i) : 0.0;
    { const double t = g_down(g, wf0); acc.xs[0] += (g.j < W - 1) ? t : 0.0; }
    { const double t = g_down(g, wf1); acc.xs[2] += (g.j < W - 1) ? t : 0.0; }
    if (NL) { XUNR for (int a = 0; a < 2; a++) acc.xs[2 * a + 1] += pr(fld + d::SV + a); }
    acc.dl = 0.0;
  }
  template <bool NL>
  XMEM void assemble(const KParams& P, int fld, double rho) {   // slot part of the stage Hessian (rows are added by the caller)
    XUNR for (int i = 0; i < 4; i++) H.hx[i] = c.qw + pr(fld + d::SX + i);
    XUNR for (int a = 0; a < 2; a++) {
      const double sf = (g.j > 0) ? pr(fld + (NL ? d::SFN : d::SFQ) + a) : 0.0;
      H.hu[a] = P.Rw + pr(fld + d::SU + a) + sf;
      H.hxu[a] = -sf;
      const double t = g_down(g, sf);
      H.hx[2 * a] += (g.j < W - 1) ? t : 0.0;
    }
    XUNR for (int i = 0; i < 4; i++) H.hl[i] = (i < C) ? rho + pr(fld + d::SL + i) : 0.0;
    if (NL) { XUNR for (int a = 0; a < 2; a++) H.hx[2 * a + 1] += pr(fld + d::SV + a); }
    H.hxy = 0.0; H.bd[0] = 0.0; H.bd[1] = 0.0; H.hd = (g.j == W - 1) ? P.slw : 0.0;
  }

  // ===============================================================================================================
  // Factorisation: Riccati recursion over the W lanes on the 5-state (x, delta).  Returns false when a pivot is not positive.
  // ===============================================================================================================
  template <bool FLOOR>
  XMEM bool factor(const KParams& P, double dw, int kc_in) {
    // cost on xi_j = the previous lane's blocks
    double hxp[4], hxyp, bdp[2], hdp;
    XUNR for (int i = 0; i < 4; i++) hxp[i] = g_up(g, H.hx[i]);
    hxyp = g_up(g, H.hxy); bdp[0] = g_up(g, H.bd[0]); bdp[1] = g_up(g, H.bd[1]); hdp = g_up(g, H.hd + ((g.j == W - 1) ? dw : 0.0));
    if (g.j == 0) { XUNR for (int i = 0; i < 4; i++) hxp[i] = 0.0; hxyp = 0.0; bdp[0] = bdp[1] = 0.0; hdp = 0.0; }
    double Pown[15];
    XUNR for (int t = 0; t < 15; t++) Pown[t] = 0.0;
    XUNR for (int i = 0; i < 4; i++) Pown[xsy(i, i)] = H.hx[i] + dw;
    Pown[xsy(2, 0)] = H.hxy; Pown[xsy(4, 0)] = H.bd[0]; Pown[xsy(4, 2)] = H.bd[1];
    Pown[xsy(4, 4)] = H.hd + ((g.j == W - 1) ? dw : 0.0);
    double Pout[15];
    XUNR for (int t = 0; t < 15; t++) { Pout[t] = 0.0; f.Pn[t] = Pown[t]; }
    f.kc = kc_in;
    bool ok = true;
#pragma unroll 1
    for (int step = 0; step < W; step++) {
      if (step > 0) {   // P_{j+1} = own blocks + what lane j+1 produced in the previous step
        XUNR for (int t = 0; t < 15; t++) { const double v = g_down(g, Pout[t]); if (g.j < W - 1) f.Pn[t] = Pown[t] + v; }
      }
      const double* Pn = f.Pn;
      double PA[25];   // Pn A^ (A^ = diag(A2, A2, 1))
      XUNR for (int i = 0; i < 5; i++) {
        XUNR for (int b = 0; b < 2; b++) {
          const double p0 = Pn[xsy(i, 2 * b)], p1 = Pn[xsy(i, 2 * b + 1)];
          PA[5 * i + 2 * b] = p0 * P.a11 + p1 * P.a21;
          PA[5 * i + 2 * b + 1] = p0 * P.a12 + p1 * P.a22;
        }
        PA[5 * i + 4] = Pn[xsy(i, 4)];
      }
      double G2[4], GA[10];   // B'Pn restricted to the x columns (2 x 2 block products), B'Pn A^ (2 x 5)
      XUNR for (int cc = 0; cc < 2; cc++) {
        XUNR for (int m = 0; m < 5; m++) GA[5 * cc + m] = P.bp * PA[5 * (2 * cc) + m] + P.bv * PA[5 * (2 * cc + 1) + m];
      }
      // N2 = B'Pn B + diag(hu + dw)
      G2[0] = P.bp * Pn[xsy(0, 0)] + P.bv * Pn[xsy(1, 0)]; G2[1] = P.bp * Pn[xsy(0, 1)] + P.bv * Pn[xsy(1, 1)];
      G2[2] = P.bp * Pn[xsy(2, 2)] + P.bv * Pn[xsy(3, 2)]; G2[3] = P.bp * Pn[xsy(2, 3)] + P.bv * Pn[xsy(3, 3)];
      const double g10 = P.bp * Pn[xsy(2, 0)] + P.bv * Pn[xsy(3, 0)], g11 = P.bp * Pn[xsy(2, 1)] + P.bv * Pn[xsy(3, 1)];
      double N2[3];
      N2[0] = P.bp * G2[0] + P.bv * G2[1] + (H.hu[0] + dw);
      N2[1] = P.bp * g10 + P.bv * g11;
      N2[2] = P.bp * G2[2] + P.bv * G2[3] + (H.hu[1] + dw);
      double Rt[6] = {0.0, 0.0, 1.0, 0.0, 0.0, 1.0};
      {
        double T1[6];
        XUNR for (int b = 0; b < NW; b++) { T1[b] = N2[0] * c.D[b] + N2[1] * c.D[NW + b]; T1[NW + b] = N2[1] * c.D[b] + N2[2] * c.D[NW + b]; }
        const double hlast = H.hl[C - 1] + dw;
        XUNR for (int a = 0; a < NW; a++) XUNR for (int b = 0; b <= a; b++) {
          double s = c.D[a] * T1[b] + c.D[NW + a] * T1[NW + b] + hlast;
          if (a == b) s += H.hl[a] + dw;
          Rt[(a * (a + 1)) / 2 + b] = s;
        }
      }
      const bool okc = x_chol<NW, FLOOR>(Rt, f.Lc);
      GA[0] += H.hxu[0]; GA[5 + 2] += H.hxu[1];
      double Y[15];
      XUNR for (int m = 0; m < 5; m++) {
        double col[3] = {0.0, 0.0, 0.0};
        XUNR for (int a = 0; a < NW; a++) col[a] = c.D[a] * GA[m] + c.D[NW + a] * GA[5 + m];
        x_fwd<NW>(f.Lc, col);
        XUNR for (int a = 0; a < 3; a++) Y[5 * a + m] = (a < NW) ? col[a] : 0.0;
        x_bwd<NW>(f.Lc, col);
        XUNR for (int a = 0; a < 3; a++) f.K[5 * a + m] = (a < NW) ? col[a] : 0.0;
      }
      // P_j = A^'PA + (cost on xi_j) - Y'Y
      XUNR for (int i = 0; i < 5; i++) XUNR for (int k = 0; k <= i; k++) {
        double s;
        if (i < 4) { const int bi = i >> 1, ri = i & 1; s = (ri ? P.a12 : P.a11) * PA[5 * (2 * bi) + k] + (ri ? P.a22 : P.a21) * PA[5 * (2 * bi + 1) + k]; }
        else s = PA[5 * 4 + k];
        XUNR for (int a = 0; a < NW; a++) s -= Y[5 * a + i] * Y[5 * a + k];
        Pout[xsy(i, k)] = s;
      }
      // rows kept in quasi-definite form: T = diag(1 / (Sigma - cap)) + M R~^-1 M', N~ = N - M R~^-1 S~, P_j += N~' T^-1 N~
      if (f.kc > 0) {
        double t00 = 0.0, t10 = 0.0, t11 = 1.0;
        XUNR for (int q = 0; q < XKC; q++) {
          if (q < f.kc) {
            double col[3] = {0.0, 0.0, 0.0};
            XUNR for (int a = 0; a < NW; a++) col[a] = P.bp * (cap[q].gx * c.D[a] + cap[q].gy * c.D[NW + a]);
            x_fwd<NW>(f.Lc, col);
            XUNR for (int a = 0; a < 3; a++) f.Ym[q][a] = (a < NW) ? col[a] : 0.0;
            x_bwd<NW>(f.Lc, col);
            XUNR for (int a = 0; a < 3; a++) f.Zm[q][a] = (a < NW) ? col[a] : 0.0;
            const double Nq[5] = {cap[q].gx * P.a11, cap[q].gx * P.a12, cap[q].gy * P.a11, cap[q].gy * P.a12, 1.0};
            XUNR for (int m = 0; m < 5; m++) {
              double s = Nq[m];
              XUNR for (int a = 0; a < NW; a++) s -= f.Ym[q][a] * Y[5 * a + m];
              f.Nt[q][m] = s;
            }
          }
        }
        t00 = cap[0].srinv + f.Ym[0][0] * f.Ym[0][0] + f.Ym[0][1] * f.Ym[0][1] + f.Ym[0][2] * f.Ym[0][2];
        if (XKC == 2 && f.kc == 2) {
          t10 = f.Ym[XKC - 1][0] * f.Ym[0][0] + f.Ym[XKC - 1][1] * f.Ym[0][1] + f.Ym[XKC - 1][2] * f.Ym[0][2];
          t11 = cap[XKC - 1].srinv + f.Ym[XKC - 1][0] * f.Ym[XKC - 1][0] + f.Ym[XKC - 1][1] * f.Ym[XKC - 1][1] + f.Ym[XKC - 1][2] * f.Ym[XKC - 1][2];
        }
        x_chol2(f.kc, t00, t10, t11, f.Tc);
        // V = Lt^-1 N~ ; P_j += V'V
        double V0[5], V1[5];
        XUNR for (int m = 0; m < 5; m++) {
          V0[m] = f.Nt[0][m] * f.Tc[0];
          V1[m] = (XKC == 2 && f.kc == 2) ? (f.Nt[XKC - 1][m] - f.Tc[1] * V0[m]) * f.Tc[2] : 0.0;
        }
        XUNR for (int i = 0; i < 5; i++) XUNR for (int k = 0; k <= i; k++) Pout[xsy(i, k)] += V0[i] * V0[k] + V1[i] * V1[k];
      }
      // what is handed to lane j-1 excludes that lane's own blocks (it adds them itself); lane 0 keeps its P_0 for delta_0
      ok = okc;
    }
    (void)hxp; (void)hxyp; (void)bdp; (void)hdp;
    // curvature of delta_0: lane 0's Pout(4,4) (x_0 is fixed); must be positive
    const double p0 = g_idx(g, Pout[xsy(4, 4)], 0);
    f.p0dd = p0;
    return g_all(g, ok) && (p0 > 0.0) && (p0 < 1e300);
  }

  // right-hand side after the elimination of (u_j, lambda_j) through u = F lambda, 1'lambda = 1 (cbfnmpc_lane.h l_stage_eliminate)
  XMEM void stage_eliminate(const KParams& P, double dw, const XKV& r, XMain& m, double* d0) const {
    const double e0 = r.eta;
    XUNR for (int a = 0; a < 2; a++) d0[a] = c.Fl[a] * e0 + r.nu[a];
    const double last = r.lm[C - 1] - (H.hl[C - 1] + dw) * e0;
    const double ru0 = r.us[0] - (H.hu[0] + dw) * d0[0], ru1 = r.us[1] - (H.hu[1] + dw) * d0[1];
    XUNR for (int a = 0; a < 3; a++) m.rw[a] = 0.0;
    XUNR for (int a = 0; a < NW; a++) m.rw[a] = r.lm[a] - last + c.D[a] * ru0 + c.D[NW + a] * ru1;
    XUNR for (int b = 0; b < 2; b++) { m.cv[2 * b] = P.bp * d0[b] - r.pi[2 * b]; m.cv[2 * b + 1] = P.bv * d0[b] - r.pi[2 * b + 1]; }
    XUNR for (int b = 0; b < 2; b++) {
      const double t = g_down(g, H.hxu[b] * d0[b]);
      m.rx[2 * b] = r.xs[2 * b] - ((g.j < W - 1) ? t : 0.0);
      m.rx[2 * b + 1] = r.xs[2 * b + 1];
    }
    m.rx[4] = r.dl;   // this lane's share of the slack variable's right-hand side
  }
  XMEM void recover(const KParams& P, double dw, const XKV& r, const double* d0, const double* w, XKV& dd) const {
    const double e0 = r.eta;
    double wsum = 0.0, du0 = d0[0], du1 = d0[1];
    XUNR for (int i = 0; i < 4; i++) dd.lm[i] = 0.0;
    XUNR for (int a = 0; a < NW; a++) { dd.lm[a] = w[a]; wsum += w[a]; du0 += c.D[a] * w[a]; du1 += c.D[NW + a] * w[a]; }
    const double dll = e0 - wsum;
    dd.lm[C - 1] = dll;
    dd.us[0] = du0; dd.us[1] = du1;
    const double xp0 = g_up(g, dd.xs[0]), xp2 = g_up(g, dd.xs[2]);
    const double du[2] = {du0, du1};
    const double xpp[2] = {(g.j > 0) ? xp0 : 0.0, (g.j > 0) ? xp2 : 0.0};
    XUNR for (int a = 0; a < 2; a++)
      dd.nu[a] = r.us[a] - (H.hu[a] + dw) * du[a] - (P.bp * dd.pi[2 * a] + P.bv * dd.pi[2 * a + 1]) - H.hxu[a] * xpp[a];
    dd.eta = r.lm[C - 1] - (H.hl[C - 1] + dw) * dll + c.Fl[0] * dd.nu[0] + c.Fl[1] * dd.nu[1];
  }

  // Solve [H A'; A 0][dx; dy] = r with the current factorisation.  tq = multipliers t of this lane's quasi-definite rows.
  // rq = right-hand side of the quasi-definite rows (g'dx - t / (Sigma - cap) = rq; zero in the first solve, the residual in refinement)
  XMEM void kkt_solve(const KParams& P, double dw, const XKV& r, const double* rq, XKV& dd, double* tq) const {
    XMain m; double d0[2];
    stage_eliminate(P, dw, r, m, d0);
    // ---- backward chain
    double out[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
    for (int step = 0; step < W; step++) {
      double inc[5];
      XUNR for (int i = 0; i < 5; i++) { const double t = g_down(g, out[i]); inc[i] = (step > 0 && g.j < W - 1) ? t : 0.0; }
      double pn[5], q[5];
      XUNR for (int i = 0; i < 5; i++) pn[i] = m.rx[i] + inc[i];
      XUNR for (int i = 0; i < 5; i++) {
        double s = -pn[i];
        XUNR for (int k = 0; k < 4; k++) s += f.Pn[xsy(i, k)] * m.cv[k];
        q[i] = s;
      }
      const double g0 = P.bp * q[0] + P.bv * q[1], g1 = P.bp * q[2] + P.bv * q[3];
      double k0[3] = {0.0, 0.0, 0.0}, kr[3] = {0.0, 0.0, 0.0}, ykr[3];
      XUNR for (int a = 0; a < NW; a++) { kr[a] = m.rw[a] - (c.D[a] * g0 + c.D[NW + a] * g1); k0[a] = kr[a]; }
      x_fwd<NW>(f.Lc, k0);
      XUNR for (int a = 0; a < 3; a++) ykr[a] = k0[a];
      x_bwd<NW>(f.Lc, k0);
      double tt[2] = {0.0, 0.0};
      if (f.kc > 0) {
        XUNR for (int qq = 0; qq < XKC; qq++)
          if (qq < f.kc) {
            double s = cap[qq].gx * m.cv[0] + cap[qq].gy * m.cv[2] - rq[qq];
            XUNR for (int a = 0; a < NW; a++) s += f.Ym[qq][a] * ykr[a];
            tt[qq] = s;
          }
        x_chol2_solve(f.kc, f.Tc, tt);
      }
      double aq[4];
      x_ATx(P, q, aq);
      XUNR for (int i = 0; i < 5; i++) {
        double s = (i < 4) ? -aq[i] : -q[4];
        XUNR for (int a = 0; a < NW; a++) s -= f.K[5 * a + i] * kr[a];
        if (f.kc > 0) { XUNR for (int qq = 0; qq < XKC; qq++) if (qq < f.kc) s -= f.Nt[qq][i] * tt[qq]; }
        out[i] = s;
      }
      XUNR for (int i = 0; i < 5; i++) m.pv[i] = pn[i];
      XUNR for (int a = 0; a < 3; a++) m.kv[a] = k0[a];
      XUNR for (int qq = 0; qq < XKC; qq++) m.tt[qq] = tt[qq];
    }
    // ---- delta_0 (x_0 is fixed): P_0(4,4) ddelta = p_0(4)
    const double ddl = g_idx(g, out[4], 0) * x_rcp(f.p0dd);
    // ---- forward chain
    double xn[5] = {0.0, 0.0, 0.0, 0.0, ddl}, w[3] = {0.0, 0.0, 0.0};
    tq[0] = 0.0; tq[1] = 0.0;
#pragma unroll 1
    for (int step = 0; step < W; step++) {
      double xp[5];
      XUNR for (int i = 0; i < 4; i++) { const double t = g_up(g, xn[i]); xp[i] = (step > 0 && g.j > 0) ? t : 0.0; }
      xp[4] = ddl;
      double tl[2] = {0.0, 0.0};
      if (f.kc > 0) {
        XUNR for (int qq = 0; qq < XKC; qq++)
          if (qq < f.kc) { double s = 0.0; XUNR for (int i = 0; i < 5; i++) s += f.Nt[qq][i] * xp[i]; tl[qq] = s; }
        x_chol2_solve(f.kc, f.Tc, tl);
        XUNR for (int qq = 0; qq < XKC; qq++) tl[qq] += m.tt[qq];
      }
      XUNR for (int a = 0; a < 3; a++) w[a] = 0.0;
      XUNR for (int a = 0; a < NW; a++) {
        double s = m.kv[a];
        XUNR for (int k = 0; k < 5; k++) s -= f.K[5 * a + k] * xp[k];
        if (f.kc > 0) { XUNR for (int qq = 0; qq < XKC; qq++) if (qq < f.kc) s -= f.Zm[qq][a] * tl[qq]; }
        w[a] = s;
      }
      double e0 = 0.0, e1 = 0.0;
      XUNR for (int a = 0; a < NW; a++) { e0 += c.D[a] * w[a]; e1 += c.D[NW + a] * w[a]; }
      double ax[4];
      x_Ax(P, xp, ax);
      ax[0] += P.bp * e0; ax[1] += P.bv * e0; ax[2] += P.bp * e1; ax[3] += P.bv * e1;
      XUNR for (int i = 0; i < 4; i++) { ax[i] += m.cv[i]; xn[i] = ax[i]; }
      XUNR for (int i = 0; i < 4; i++) {
        double s = -m.pv[i] + f.Pn[xsy(i, 4)] * ddl;
        XUNR for (int k = 0; k < 4; k++) s += f.Pn[xsy(i, k)] * ax[k];
        dd.xs[i] = ax[i]; dd.pi[i] = s;
      }
      if (f.kc > 0) {
        XUNR for (int qq = 0; qq < XKC; qq++) if (qq < f.kc) { dd.pi[0] += cap[qq].gx * tl[qq]; dd.pi[2] += cap[qq].gy * tl[qq]; }
      }
      tq[0] = tl[0]; tq[1] = tl[1];
    }
    dd.dl = ddl;
    recover(P, dw, r, d0, w, dd);
  }

  // ===============================================================================================================
  // Stage 1: QP, iSWIFT restated (Prime.c:127-230; kkt_initialize Auxilary.c:680-755) -- cbfnmpc_lane.h qp_init / qp_iter on W lanes
  // ===============================================================================================================
  XMEM void qp_dual_eq(const KParams& P, const double* xp, const double* pin, XKV& rx, double& acc_rx, double& acc_ry) {
    XKV gfy, jz;
    grad_Aty(P, x, x, pin, true, gfy);
    gather<false>(d::Q_DE, jz);
    acc_rx = 0.0; acc_ry = 0.0;
    XUNR for (int i = 0; i < 4; i++) { rx.xs[i] = -(gfy.xs[i] - jz.xs[i]); acc_rx += rx.xs[i] * rx.xs[i]; }
    XUNR for (int a = 0; a < 2; a++) { rx.us[a] = -(gfy.us[a] - jz.us[a]); acc_rx += rx.us[a] * rx.us[a]; }
    XUNR for (int i = 0; i < 4; i++) { rx.lm[i] = (i < C) ? -(gfy.lm[i] - jz.lm[i]) : 0.0; acc_rx += rx.lm[i] * rx.lm[i]; }
    rx.dl = -gfy.dl;
    acc_rx += rx.dl * rx.dl;
    double rpi[4], rnu[2], reta;
    eq_residual(P, x, xp, false, rpi, rnu, reta);
    XUNR for (int i = 0; i < 4; i++) { rx.pi[i] = -rpi[i]; acc_ry += rpi[i] * rpi[i]; }
    XUNR for (int a = 0; a < 2; a++) { rx.nu[a] = -rnu[a]; acc_ry += rnu[a] * rnu[a]; }
    rx.eta = -reta; acc_ry += reta * reta;
  }
  XMEM bool qp_init(const KParams& P) {
    XSLOTS for (int s = 0; s < nsq; s++) { pr(d::Q_E + s) = 2.0; pr(d::Q_G + s) = slot_hi(P, false, s) + slot_lo(P, false, s); }
    XUNR for (int i = 0; i < 4; i++) { x.xs[i] = 0.0; x.lm[i] = 0.0; x.pi[i] = 0.0; }
    x.us[0] = x.us[1] = 0.0; x.dl = 0.0; x.nu[0] = x.nu[1] = 0.0; x.eta = 0.0;
    assemble<false>(P, d::Q_E, 0.0);
    XKV r, dd;
    gather<false>(d::Q_G, r);
    XUNR for (int i = 0; i < 4; i++) r.xs[i] += c.qw * pr(d::F_XDES + i);
    r.dl = 0.0;
    {
      double xp[4]; prev_state(x, false, xp);
      double rpi[4], rnu[2], reta;
      eq_residual(P, x, xp, false, rpi, rnu, reta);
      XUNR for (int i = 0; i < 4; i++) r.pi[i] = -rpi[i];
      r.nu[0] = -rnu[0]; r.nu[1] = -rnu[1]; r.eta = -reta;
    }
    qp_flag = 3; qp_iters = 0; qp_it = 0; sigma = 100.0;
    if (!factor<true>(P, 0.0, 0)) { qp_flag = 1; return true; }
    double tq[2]; const double rq0[2] = {0.0, 0.0};
    kkt_solve(P, 0.0, r, rq0, dd, tq);
    x = dd;
    XUNR for (int i = C; i < 4; i++) x.lm[i] = 0.0;
    double xp[4]; prev_state(x, false, xp);
    scatter<false>(d::Q_EV, x, xp);
    double rmin = 1e300, rmax = -1e300;
    XSLOTS for (int s = 0; s < nsq; s++) {
      const double e = pr(d::Q_EV + s);
      const double rl = e - slot_lo(P, false, s), rh = slot_hi(P, false, s) - e;
      rmin = fmin(rmin, fmin(rl, rh)); rmax = fmax(rmax, fmax(rl, rh));
    }
    rmin = g_min(g, rmin); rmax = g_max(g, rmax);
    const double ap = -rmin, ad = rmax;
    XSLOTS for (int s = 0; s < nsq; s++) {
      const double e = pr(d::Q_EV + s);
      const double rl = e - slot_lo(P, false, s), rh = slot_hi(P, false, s) - e;
      pr(d::Q_SLO + s) = (ap < 0) ? rl : rl + (1.0 + ap);
      pr(d::Q_SHI + s) = (ap < 0) ? rh : rh + (1.0 + ap);
      pr(d::Q_ZLO + s) = (ad < 0) ? -rl : -rl + (1.0 + ad);
      pr(d::Q_ZHI + s) = (ad < 0) ? -rh : -rh + (1.0 + ad);
    }
    qp_flag = 2;
    return false;
  }
  XMEM bool qp_iter(const KParams& P) {
    if (qp_it >= P.qp_maxit) return true;
    const int MQP = 16 * W - 4 + 2 * C * W;   // rows of G (MPC_dist.cpp:264-321 for W grid points)
    double xp[4], pin[4];
    prev_state(x, false, xp); next_pi(x, pin);
    scatter<false>(d::Q_EV, x, xp);
    XKV rx;
    double sz, mu_;
    const bool predictor = sigma > 0.0;
    {
      double acc_rz = 0.0, acc_sz = 0.0;
      XSLOTS for (int s = 0; s < nsq; s++) {
        const double lo = slot_lo(P, false, s), hi = slot_hi(P, false, s), e = pr(d::Q_EV + s);
        const double slo = pr(d::Q_SLO + s), shi = pr(d::Q_SHI + s), zlo = pr(d::Q_ZLO + s), zhi = pr(d::Q_ZHI + s);
        const double rzh = hi - shi - e, rzl = -lo - slo + e;
        acc_rz += rzh * rzh + rzl * rzl;
        acc_sz += slo * zlo + shi * zhi;
        pr(d::Q_DE + s) = zlo - zhi;
        const double isl = x_rcp(slo), ish = x_rcp(shi);
        pr(d::Q_A + s) = isl; pr(d::Q_B + s) = ish;
        pr(d::Q_C + s) = x_rcp(zlo); pr(d::Q_D + s) = x_rcp(zhi);
        const double wl = zlo * isl, wh = zhi * ish;
        pr(d::Q_E + s) = predictor ? wl + wh : 0.0;
        pr(d::Q_G + s) = predictor ? wh * (rzh + shi) - wl * (rzl + slo) : 0.0;
      }
      double acc_rx, acc_ry;
      qp_dual_eq(P, xp, pin, rx, acc_rx, acc_ry);
      const double nrx = g_sum(g, acc_rx), nry = g_sum(g, acc_ry), nrz = g_sum(g, acc_rz);
      sz = g_sum(g, acc_sz);
      const double lim = P.qp_reltol / sqrt(3.0);
      if (sqrt(nrx) < lim && sqrt(nrz) < lim && sqrt(nry) < lim && sz / MQP < P.qp_abstol) { qp_flag = 0; return true; }
      mu_ = sz / MQP;
    }
    if (predictor) {
      assemble<false>(P, d::Q_E, 0.0);
      n_fact++;
      if (!factor<true>(P, 0.0, 0)) { qp_flag = 1; return true; }
    } else {
      sigma = 0.0;
    }
    double ap = 1.0, ad = 1.0;
    XKV dd;
    for (int ph = predictor ? 0 : 1; ph < 2; ph++) {
      const double sm = sigma * mu_;
      if (ph == 1) {
        XSLOTS for (int s = 0; s < nsq; s++) {
          const double lo = slot_lo(P, false, s), hi = slot_hi(P, false, s), e = pr(d::Q_EV + s);
          const double slo = pr(d::Q_SLO + s), shi = pr(d::Q_SHI + s), zlo = pr(d::Q_ZLO + s), zhi = pr(d::Q_ZHI + s);
          const double wl = zlo * pr(d::Q_A + s), wh = zhi * pr(d::Q_B + s);
          const double rzh = hi - shi - e, rzl = -lo - slo + e;
          const double cdl = -slo * zlo - pr(d::Q_E + s) + sm, cdh = -shi * zhi - pr(d::Q_G + s) + sm;
          const double r3l = rzl - cdl * pr(d::Q_C + s), r3h = rzh - cdh * pr(d::Q_D + s);
          pr(d::Q_DE + s) = wh * r3h - wl * r3l;
        }
      }
      XKV r;
      gather<false>(ph == 0 ? d::Q_G : d::Q_DE, r);
      XUNR for (int i = 0; i < 4; i++) { r.xs[i] += rx.xs[i]; r.lm[i] += rx.lm[i]; r.pi[i] = rx.pi[i]; }
      XUNR for (int a = 0; a < 2; a++) { r.us[a] += rx.us[a]; r.nu[a] = rx.nu[a]; }
      r.dl = rx.dl; r.eta = rx.eta;
      double tq[2]; const double rq0[2] = {0.0, 0.0};
      kkt_solve(P, 0.0, r, rq0, dd, tq);
      double dxp[4]; prev_state(dd, true, dxp);
      scatter<false>(d::Q_DE, dd, dxp);
      double rp = 0.0, rdm = 0.0, s_dsz = 0.0, s_sdz = 0.0, s_dsdz = 0.0;
      const bool aff = ph == 0;
      XSLOTS for (int s = 0; s < nsq; s++) {
        const double lo = slot_lo(P, false, s), hi = slot_hi(P, false, s), e = pr(d::Q_EV + s), de = pr(d::Q_DE + s);
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
          pr(d::Q_E + s) = dsl * dzl; pr(d::Q_G + s) = dsh * dzh;
        } else {
          pr(d::Q_A + s) = dsl; pr(d::Q_B + s) = dsh; pr(d::Q_C + s) = dzl; pr(d::Q_D + s) = dzh;
        }
      }
      rp = g_max(g, rp); rdm = g_max(g, rdm);
      ap = (rp > 0.0) ? 1.0 / rp : 1.0; ad = (rdm > 0.0) ? 1.0 / rdm : 1.0;
      if (aff) {
        const double num = sz + g_sum(g, ap * s_dsz + ad * s_sdz + (ap * ad) * s_dsdz);
        const double rho = num / sz, mr = rho < 1.0 ? rho : 1.0;
        sigma = mr * mr * mr;
        if (sigma < 0.0) sigma = 0.0;
      }
    }
    ap = fmin(0.99 * ap, 1.0); ad = fmin(0.99 * ad, 1.0);
    XUNR for (int i = 0; i < 4; i++) { x.xs[i] += ap * dd.xs[i]; x.pi[i] += ad * dd.pi[i]; }
    XUNR for (int i = 0; i < C; i++) x.lm[i] += ap * dd.lm[i];
    XUNR for (int a = 0; a < 2; a++) { x.us[a] += ap * dd.us[a]; x.nu[a] += ad * dd.nu[a]; }
    x.dl += ap * dd.dl; x.eta += ad * dd.eta;
    XSLOTS for (int s = 0; s < nsq; s++) {
      pr(d::Q_SLO + s) += ap * pr(d::Q_A + s); pr(d::Q_SHI + s) += ap * pr(d::Q_B + s);
      pr(d::Q_ZLO + s) += ad * pr(d::Q_C + s); pr(d::Q_ZHI + s) += ad * pr(d::Q_D + s);
    }
    qp_it++;
    qp_iters = qp_it;
    return false;
  }

  // ===============================================================================================================
  // Stage 2: NLP, IPM-F with K distance rows per grid point; team mode = IPM-F-J
  // ===============================================================================================================
  XMEM void filter_reset() {
    XUNR for (int q = 0; q < 2; q++) { pr(d::F_FILT + 2 * q) = CBF_INF; pr(d::F_FILT + 2 * q + 1) = CBF_INF; }
    nfilt = 0;
  }
  XMEM void filter_add(const KParams& P, double th_f, double ph_f) {   // ring of filter_cap pairs shared by the group: pair p in lane p % W... (W >= 4)
    const int p = nfilt % P.filter_cap;
    if ((p & 3) == g.j) {
      if (p >> 2) { pr(d::F_FILT + 2) = th_f; pr(d::F_FILT + 3) = ph_f; } else { pr(d::F_FILT) = th_f; pr(d::F_FILT + 1) = ph_f; }
    }
    nfilt++;
  }
  XMEM bool filter_ok(double tht, double pht) const {
    bool okl = true;
    if (g.j < 4) { XUNR for (int q = 0; q < 2; q++) okl = okl && !(tht >= pr(d::F_FILT + 2 * q) && pht >= pr(d::F_FILT + 2 * q + 1)); }
    return g_all(g, okl);
  }
  XMEM void slots_e(int fld, const double* xp) const { scatter<true>(fld, x, xp); }

  // centres of the neighbour rows <- nbx[i], nby[i] (neighbour i's published position at this lane's grid point)
  XMEM void set_nbr(const double* nbx, const double* nby) {
    for (int i = 0; i < K - K_obs; i++) { row(K_obs + i, XR_CX) = nbx[i]; row(K_obs + i, XR_CY) = nby[i]; }
  }
  XMEM double row_eps(const XParams& XP, int r) const { return r < K_obs ? XP.eps_obs : XP.eps_nbr; }

  // Returns true when the stage is finished already (no interior start: status 4)
  XMEM bool nlp_init(const XParams& XP) {
    const KParams& P = XP.k;
    double chk = 0.0 * x.dl + 0.0 * x.us[0] + 0.0 * x.us[1];
    XUNR for (int i = 0; i < 4; i++) chk += 0.0 * x.xs[i];
    XUNR for (int i = 0; i < C; i++) chk += 0.0 * x.lm[i];
    XUNR for (int i = 0; i < 4; i++) {
      const double lim = (i & 1) ? fmin(P.box, P.vmax) : P.box;
      x.xs[i] = fmin(fmax(x.xs[i], -lim + P.push), lim - P.push);
    }
    XUNR for (int i = 0; i < C; i++) x.lm[i] = fmin(fmax(x.lm[i], P.push), 1.0 - P.push);
    double xp[4]; prev_state(x, false, xp);
    XUNR for (int a = 0; a < 2; a++) {
      double lo = -P.box + P.push, hi = P.box - P.push;
      if (g.j >= 1) { lo = fmax(lo, xp[2 * a] - P.fb + P.push); hi = fmin(hi, xp[2 * a] + P.fb - P.push); }
      x.us[a] = fmin(fmax(x.us[a], lo), hi);
    }
    {
      double cmin = 1e300;
      for (int r = 0; r < K; r++) {
        const double a = x.xs[0] - row(r, XR_CX), b = x.xs[2] - row(r, XR_CY);
        cmin = fmin(cmin, a * a + b * b + x.dl - row_eps(XP, r));
      }
      cmin = g_min(g, cmin);
      if (cmin < P.push) x.dl += P.push - cmin;
    }
    XUNR for (int i = 0; i < 4; i++) x.pi[i] = 0.0;
    x.nu[0] = x.nu[1] = 0.0; x.eta = 0.0;
    filter_reset();
    mu = P.mu_init; XPERS_TAU = fmax(P.tau_min, 1.0 - mu); XPERS_DWLAST = 0.0; first_ic = true; nl_it = 0; XPERS_ADPREV = 0.0; hold = false; mu_age = 0;
    slots_e(d::N_G, xp);
    bool interior = chk == 0.0;
    XSLOTS for (int s = 0; s < nsn; s++) {
      const double e = pr(d::N_G + s);
      const double slo = e - slot_lo(P, true, s), shi = slot_hi(P, true, s) - e;
      interior = interior && (slo > 0.0) && (shi > 0.0);
      pr(d::N_SLO + s) = slo; pr(d::N_SHI + s) = shi;
      pr(d::N_ZLO + s) = mu * x_rcp(slo); pr(d::N_ZHI + s) = mu * x_rcp(shi);
    }
    for (int r = 0; r < K; r++) {
      const double a = x.xs[0] - row(r, XR_CX), b = x.xs[2] - row(r, XR_CY);
      const double s = a * a + b * b + x.dl - row_eps(XP, r);
      interior = interior && (s > 0.0);
      row(r, XR_S) = s; row(r, XR_SP) = s; row(r, XR_Z) = mu * x_rcp(s); row(r, XR_DZ) = 0.0; row(r, XR_DE) = 0.0;
    }
    {
      double rpi[4], rnu[2], reta;
      eq_residual(P, x, xp, false, rpi, rnu, reta);
      double rr = fabs(reta) + fabs(rnu[0]) + fabs(rnu[1]);
      XUNR for (int i = 0; i < 4; i++) rr += fabs(rpi[i]);
      const double th0 = g_sum(g, rr);
      XPERS_THMAX = 1e4 * fmax(1.0, th0); XPERS_THMIN = 1e-4 * fmax(1.0, th0);
    }
    status = 1; nlp_iters = 0;
    if (!g_all(g, interior)) { status = 4; kkt = CBF_INF; return true; }
    return false;
  }

  // One IPM-F(-J) iteration.  Returns 0: stepped; 1: converged for the current neighbour positions (no step taken; team mode only:
  // a later call may step again); 2: finished for good (status set).
  XMEM int nlp_iter(const XParams& XP, bool jacobi) {
    const KParams& P = XP.k;
    const int it = nl_it;
    if (jacobi) {
      // the neighbours have moved: keep s = c(x) > 0 by raising the shared slack variable when a neighbour row would lose more than
      // 90 % of its slack, or -- for rows that are (nearly) active, slack <= XJ_SFLOOR -- any of it (an active slack sits at
      // mu / z ~ 1e-13, three digits above the rounding of the O(1) distances it is computed from); the filter was measured
      // against other neighbour positions
      double bump = 0.0;
      for (int r = K_obs; r < K; r++) {
        const double a = x.xs[0] - row(r, XR_CX), b = x.xs[2] - row(r, XR_CY);
        const double sp = row(r, XR_SP);
        bump = fmax(bump, fmax(0.1 * sp, fmin(sp, XJ_SFLOOR)) - (a * a + b * b + x.dl - XP.eps_nbr));
      }
      bump = g_max(g, bump);
      if (bump > 0.0) x.dl += bump;
      filter_reset();
    }
    double xp[4], pin[4];
    prev_state(x, false, xp); next_pi(x, pin);
    // ---- pass A: slacks, pending multiplier step + reset, reciprocals, weights; KKT error
    XKV gfy; double re_pi[4], re_nu[2], re_eta, e_d, e_p, th, szmin = 1e300, szmax = 0.0, zmaxl = 0.0;
    bool spos = true;
    const double iks = 1.0 / P.kappa_sigma, adp = XPERS_ADPREV;
    slots_e(d::N_G, xp);
    XSLOTS for (int s = 0; s < nsn; s++) {
      const double e = pr(d::N_G + s);
      const double slo = e - slot_lo(P, true, s), shi = slot_hi(P, true, s) - e;
      const double rlo = x_rcp(slo), rhi = x_rcp(shi);
      double zlo = pr(d::N_ZLO + s), zhi = pr(d::N_ZHI + s);
      spos = spos && (slo > 0.0) && (shi > 0.0);
      if (it > 0) {
        zlo += adp * pr(d::N_D + s); zhi += adp * pr(d::N_E + s);
        const double pl = slo * zlo, ph_ = shi * zhi;
        if (!(pl >= mu * iks && pl <= P.kappa_sigma * mu)) { const double mr = mu * rlo; zlo = fmin(fmax(zlo, mr * iks), P.kappa_sigma * mr); }
        if (!(ph_ >= mu * iks && ph_ <= P.kappa_sigma * mu)) { const double mr = mu * rhi; zhi = fmin(fmax(zhi, mr * iks), P.kappa_sigma * mr); }
        pr(d::N_ZLO + s) = zlo; pr(d::N_ZHI + s) = zhi;
      }
      pr(d::N_SLO + s) = slo; pr(d::N_SHI + s) = shi; pr(d::N_A + s) = rlo; pr(d::N_B + s) = rhi;
      const double szl = slo * zlo, szh = shi * zhi;
      szmin = fmin(szmin, fmin(szl, szh)); szmax = fmax(szmax, fmax(szl, szh));
      zmaxl = fmax(zmaxl, fmax(zlo, zhi));
      pr(d::N_G + s) = zlo - zhi;                // J'z weights
      pr(d::N_C + s) = zlo * rlo + zhi * rhi;    // Sigma
    }
    // distance rows: J'z, and the running sums for the right-hand side / Hessian are taken in the assembly sweep below
    double jzx = 0.0, jzy = 0.0, jzd = 0.0;
    for (int r = 0; r < K; r++) {
      const double a = x.xs[0] - row(r, XR_CX), b = x.xs[2] - row(r, XR_CY);
      const double s = a * a + b * b + x.dl - row_eps(XP, r);
      const double rs = x_rcp(s);
      double z = row(r, XR_Z);
      spos = spos && (s > 0.0);
      if (it > 0) {
        z += adp * row(r, XR_DZ);
        const double pl = s * z;
        if (!(pl >= mu * iks && pl <= P.kappa_sigma * mu)) { const double mr = mu * rs; z = fmin(fmax(z, mr * iks), P.kappa_sigma * mr); }
        row(r, XR_Z) = z;
      }
      row(r, XR_S) = s; row(r, XR_RS) = rs; row(r, XR_SP) = s;
      const double sz = s * z;
      szmin = fmin(szmin, sz); szmax = fmax(szmax, sz); zmaxl = fmax(zmaxl, z);
#if defined(CBF_TRACE) && defined(CBF_HOST_EMU)
      if (!(z < 1e300) && x_lane() < 4) fprintf(stderr, "      lane %d row %d s %.3e z %.3e rs %.3e dz %.3e adp %.3e zold %.3e mu %.3e\n", x_lane(), r, s, z, rs, row(r, XR_DZ), adp, row(r, XR_Z), mu);
#endif
      jzx += 2.0 * a * z; jzy += 2.0 * b * z; jzd += z;
    }
    XPERS_ADPREV = 0.0;   // the pending multiplier step is applied once (an iteration that ends without a step must not repeat it)
    {
      XKV jz;
      grad_Aty(P, x, x, pin, true, gfy);
      gather<true>(d::N_G, jz);
      jz.xs[0] += jzx; jz.xs[2] += jzy;
      const double jzd_all = g_sum(g, jzd), gdl_all = g_sum(g, gfy.dl);
      double rdm = fabs(gdl_all - jzd_all);
      XUNR for (int i = 0; i < 4; i++) rdm = fmax(rdm, fabs(gfy.xs[i] - jz.xs[i]));
      XUNR for (int a = 0; a < 2; a++) rdm = fmax(rdm, fabs(gfy.us[a] - jz.us[a]));
      XUNR for (int i = 0; i < C; i++) rdm = fmax(rdm, fabs(gfy.lm[i] - jz.lm[i]));
      eq_residual(P, x, xp, false, re_pi, re_nu, re_eta);
      double rem = fabs(re_eta), res = fabs(re_eta);
      XUNR for (int i = 0; i < 4; i++) { rem = fmax(rem, fabs(re_pi[i])); res += fabs(re_pi[i]); }
      XUNR for (int a = 0; a < 2; a++) { rem = fmax(rem, fabs(re_nu[a])); res += fabs(re_nu[a]); }
      e_d = g_max(g, rdm); e_p = g_max(g, rem);
      szmin = g_min(g, szmin); szmax = g_max(g, szmax);
      const double zmax = g_max(g, zmaxl);
      th = g_sum(g, res);
      const bool nan_seen = g_any(g, !(rdm == rdm) || !(rem == rem) || !(szmax == szmax) || !(zmax == zmax));
      const double E0 = nan_seen ? 1e300 : fmax(e_d, fmax(e_p, szmax));
      kkt = E0; nlp_iters = it;
      XDBG("it %d mu %.2e E0 %.3e ed %.2e ep %.2e ec %.2e zmax %.2e th %.2e dl %.6f\n", it, mu, E0, e_d, e_p, szmax, zmax, th, x.dl);
      if (E0 <= P.tol || (mu <= P.tol / 10.0 && E0 <= P.acc_tol)) { status = 0; hold = true; return jacobi ? 1 : 2; }
      status = 1; hold = false;
      if (it == P.max_iter) return 2;
      if (!(zmax < P.z_div) || !(E0 < 1e300) || g_any(g, !spos)) { status = 2; return 2; }
    }
    {
      bool decreased = false;
      for (;;) {   // barrier update (Ipopt eq. 7)
        const double emu = fmax(fmax(szmax - mu, mu - szmin), fmax(e_d, e_p));
        if (emu <= P.kappa_eps * mu && mu > P.tol / 10.0) {
          mu = fmax(P.tol / 10.0, fmin(P.kappa_mu * mu, mu * sqrt(mu)));
          XPERS_TAU = fmax(P.tau_min, 1.0 - mu);
          filter_reset();
          decreased = true;
        } else break;
      }
      // IPM-F-J only: neighbours that keep moving can hold the subproblem's error just above kappa_eps mu for good (each move costs the
      // agent a little slack, the repair perturbs its dual residual); after mu_stall iterations without a decrease mu is lowered anyway
      if (decreased) mu_age = 0;
      else if (jacobi && ++mu_age >= XP.mu_stall && mu > P.tol / 10.0) {
        mu = fmax(P.tol / 10.0, fmin(P.kappa_mu * mu, mu * sqrt(mu)));
        XPERS_TAU = fmax(P.tau_min, 1.0 - mu);
        mu_age = 0;
      }
    }
    // ---- Hessian blocks and right-hand side
    XKV r;
    int kc = 0;
    double sig_max = 0.0;
    {
      assemble<true>(P, d::N_C, P.rho_lambda);
      // barrier-gradient weights of the slots (1/s_lo - 1/s_hi) -> N_D (the multiplier steps are written there afterwards)
      XSLOTS for (int s = 0; s < nsn; s++) pr(d::N_D + s) = pr(d::N_A + s) - pr(d::N_B + s);
      gather<true>(d::N_D, r);
      double rdl = 0.0, rx0 = 0.0, rx2 = 0.0;
      XUNR for (int q = 0; q < XKC; q++) { cap[q].srinv = 1.0; cap[q].gx = 0.0; cap[q].gy = 0.0; cap[q].row = -1; }
      double s0 = 0.0, s1 = 0.0;   // the (two) largest Sigma above the cap
      for (int rr = 0; rr < K; rr++) {
        const double sg = row(rr, XR_Z) * row(rr, XR_RS);
        if (sg > P.sigma_cap) {
          if (sg > s0) { if (XKC == 2) { s1 = s0; cap[XKC - 1] = cap[0]; } s0 = sg; cap[0].row = rr; }
          else if (XKC == 2 && sg > s1) { s1 = sg; cap[XKC - 1].row = rr; }
        }
      }
      kc = (cap[0].row >= 0 ? 1 : 0) + ((XKC == 2 && cap[XKC - 1].row >= 0) ? 1 : 0);
      sig_max = s0;
      for (int rr = 0; rr < K; rr++) {
        const double gx = 2.0 * (x.xs[0] - row(rr, XR_CX)), gy = 2.0 * (x.xs[2] - row(rr, XR_CY));
        const double z = row(rr, XR_Z), rs = row(rr, XR_RS), sg = z * rs;
        const bool qd = (rr == cap[0].row) || (XKC == 2 && rr == cap[XKC - 1].row);
        const double sc = qd ? P.sigma_cap : sg;
        H.hx[0] += sc * gx * gx - 2.0 * z; H.hx[2] += sc * gy * gy - 2.0 * z; H.hxy += sc * gx * gy;
        H.bd[0] += sc * gx; H.bd[1] += sc * gy; H.hd += sc;
        rx0 += mu * rs * gx; rx2 += mu * rs * gy; rdl += mu * rs;
        if (rr == cap[0].row) { cap[0].gx = gx; cap[0].gy = gy; cap[0].srinv = x_rcp(sg - sc); }
        if (XKC == 2 && rr == cap[XKC - 1].row) { cap[XKC - 1].gx = gx; cap[XKC - 1].gy = gy; cap[XKC - 1].srinv = x_rcp(sg - sc); }
      }
      XUNR for (int i = 0; i < 4; i++) { r.xs[i] = mu * r.xs[i] - gfy.xs[i]; r.lm[i] = (i < C) ? mu * r.lm[i] - gfy.lm[i] : 0.0; r.pi[i] = -re_pi[i]; }
      r.xs[0] += rx0; r.xs[2] += rx2;
      XUNR for (int a = 0; a < 2; a++) { r.us[a] = mu * r.us[a] - gfy.us[a]; r.nu[a] = -re_nu[a]; }
      r.dl = rdl - gfy.dl; r.eta = -re_eta;
    }
    // ---- factorisation with inertia correction (Ipopt algorithm IC)
    double dw = 0.0;
    {
      const double dw_last = XPERS_DWLAST;
      bool ok = false;
      for (;;) {
        n_fact++;
        ok = factor<false>(P, dw, kc);
        if (ok) break;
        if (dw == 0.0) dw = (dw_last == 0.0) ? P.dw_first : fmax(P.dw_min, P.kw_minus * dw_last);
        else dw *= (first_ic && dw_last == 0.0) ? P.kw_plus_first : P.kw_plus;
        if (dw > P.dw_max) break;
      }
      if (!ok) { status = 3; return 2; }
      if (dw > 0.0) { XPERS_DWLAST = dw; first_ic = false; }
    }
    // ---- solve, then iterative refinement on the UNREDUCED Newton system (a row whose Sigma is ~1e15 and over which the stage's own
    //      controls have little authority -- a 2-foot stance, say -- hands its stiffness on to the previous stage, where the recursion
    //      then loses digits; one or two refinement solves with the same factors restore them)
    XKV dd; double tq[2] = {0.0, 0.0};
    double rp = 0.0, rdm = 0.0, prod = 1.0, prod2 = 1.0, dsum = 0.0;
    {
      XKV& rcur = r; double rq[2] = {0.0, 0.0};   // the right-hand side is consumed by the first solve; refinement reuses its storage
      // checked only when some row is stiff enough to matter (measured: residual 5e-9 of the right-hand side at Sigma ~ 1e11,
      // 1e-6 .. 1 at 1e13 .. 1e15; below that the recursion is accurate to round-off)
      const bool any_qd = g_any(g, sig_max > XQD_CHECK);
      for (int rf = 0;; rf++) {
        XKV d2; double t2[2];
        kkt_solve(P, dw, rcur, rq, d2, t2);
        if (rf == 0) { dd = d2; tq[0] = t2[0]; tq[1] = t2[1]; }
        else {
          XUNR for (int i = 0; i < 4; i++) { dd.xs[i] += d2.xs[i]; dd.lm[i] += d2.lm[i]; dd.pi[i] += d2.pi[i]; }
          XUNR for (int a = 0; a < 2; a++) { dd.us[a] += d2.us[a]; dd.nu[a] += d2.nu[a]; }
          dd.dl += d2.dl; dd.eta += d2.eta; tq[0] += t2[0]; tq[1] += t2[1];
        }
        double ddp[4]; prev_state(dd, true, ddp);
        scatter<true>(d::N_C, dd, ddp);     // e(dx) of the slots -> N_C
        // pass B / C: multiplier steps (slots: N_D, N_E; rows: XR_DZ), the weights z + dz for the residual (slots: N_G), step bounds
        // and line-search constants (recomputed when a refinement solve follows)
        rp = 0.0; rdm = 0.0; prod = 1.0; prod2 = 1.0; dsum = 0.0;
        XSLOTS for (int s = 0; s < nsn; s++) {
          const double slo = pr(d::N_SLO + s), shi = pr(d::N_SHI + s), zlo = pr(d::N_ZLO + s), zhi = pr(d::N_ZHI + s);
          const double rlo = pr(d::N_A + s), rhi = pr(d::N_B + s), de = pr(d::N_C + s);
          const double dzl = (mu - slo * zlo) * rlo - zlo * rlo * de;
          const double dzh = (mu - shi * zhi) * rhi + zhi * rhi * de;
          pr(d::N_D + s) = dzl; pr(d::N_E + s) = dzh;
          if (any_qd) pr(d::N_G + s) = (zlo + dzl) - (zhi + dzh);
          rp = fmax(rp, fmax(-de * rlo, de * rhi));
          rdm = fmax(rdm, fmax(-dzl * x_rcp(zlo), -dzh * x_rcp(zhi)));
          const double ss = slo * shi;
          if (s & 1) prod *= ss; else prod2 *= ss;
          dsum += de * rlo - de * rhi;
        }
        double jx0 = 0.0, jx2 = 0.0, jdl = 0.0, zsum = 0.0, rqn[2] = {0.0, 0.0}, prodr = 1.0;
        for (int rr = 0; rr < K; rr++) {
          const double gx = 2.0 * (x.xs[0] - row(rr, XR_CX)), gy = 2.0 * (x.xs[2] - row(rr, XR_CY));
          const double s = row(rr, XR_S), rs = row(rr, XR_RS), z = row(rr, XR_Z);
          const double de = gx * dd.xs[0] + gy * dd.xs[2] + dd.dl;
          double dz = (mu - s * z) * rs - z * rs * de;
          if (rr == cap[0].row) { dz = (mu - s * z) * rs - P.sigma_cap * de - tq[0]; rqn[0] = cap[0].srinv * tq[0] - de; }
          if (XKC == 2 && rr == cap[XKC - 1].row) { dz = (mu - s * z) * rs - P.sigma_cap * de - tq[1]; rqn[1] = cap[XKC - 1].srinv * tq[1] - de; }
          row(rr, XR_DE) = de; row(rr, XR_DZ) = dz;
          const double zn = z + dz;
          jx0 += gx * zn; jx2 += gy * zn; jdl += zn; zsum += z;
          rp = fmax(rp, -de * rs);
          rdm = fmax(rdm, -dz * x_rcp(z));
          prodr *= s;
          dsum += de * rs;
          if ((rr & 3) == 3) { prod *= prodr; prodr = 1.0; }
        }
        prod *= prodr;
        // without a quasi-definite row the recursion is accurate to round-off (measured: residual 1e-12 of the right-hand side)
        if (rf >= P.max_refine || !any_qd) break;
        // rho1 = -(grad f + A'(y + dy) + W~ dx - J'(z + dz)),  rho2 = -(A (x + dx) - b),  rho_q = t / (Sigma - cap) - g'dx
        // (grad f + A'y and the equality residual of the iterate are recomputed here, rarely, instead of living across the solve)
        XKV jz2, gd, gf0; double dpin[4], xp0[4], pin0[4], rp0[4], rn0[2], re0;
        prev_state(x, false, xp0); next_pi(x, pin0);
        grad_Aty(P, x, x, pin0, true, gf0);
        eq_residual(P, x, xp0, false, rp0, rn0, re0);
        gather<true>(d::N_G, jz2);
        jz2.xs[0] += jx0; jz2.xs[2] += jx2;
        next_pi(dd, dpin);
        grad_Aty(P, dd, dd, dpin, false, gd);
        double nres = fmax(fabs(rqn[0]), fabs(rqn[1]));
        XUNR for (int i = 0; i < 4; i++) {
          const double ht = c.qw + dw - ((i & 1) ? 0.0 : 2.0 * zsum);
          rcur.xs[i] = -(gf0.xs[i] + gd.xs[i] + ht * dd.xs[i] - jz2.xs[i]);
          nres = fmax(nres, fabs(rcur.xs[i]));
        }
        XUNR for (int a = 0; a < 2; a++) { rcur.us[a] = -(gf0.us[a] + gd.us[a] + (P.Rw + dw) * dd.us[a] - jz2.us[a]); nres = fmax(nres, fabs(rcur.us[a])); }
        XUNR for (int i = 0; i < 4; i++) {
          rcur.lm[i] = (i < C) ? -(gf0.lm[i] + gd.lm[i] + (dw + P.rho_lambda) * dd.lm[i] - jz2.lm[i]) : 0.0;
          nres = fmax(nres, fabs(rcur.lm[i]));
        }
        rcur.dl = -(gf0.dl + ((g.j == W - 1) ? (P.slw + dw) * dd.dl : 0.0) - jdl);   // this lane's share of the slack variable's row
        const double rdl_all = g_sum(g, rcur.dl);
        nres = fmax(nres, fabs(rdl_all));
        {
          double hp[4], hn[2], he;
          eq_residual(P, dd, ddp, true, hp, hn, he);
          XUNR for (int i = 0; i < 4; i++) { rcur.pi[i] = -(rp0[i] + hp[i]); nres = fmax(nres, fabs(rcur.pi[i])); }
          XUNR for (int a = 0; a < 2; a++) { rcur.nu[a] = -(rn0[a] + hn[a]); nres = fmax(nres, fabs(rcur.nu[a])); }
          rcur.eta = -(re0 + he); nres = fmax(nres, fabs(rcur.eta));
        }
        nres = g_max(g, nres);
        XDBG("   refine %d: newton residual %.3e (rhs %.3e)\n", rf, nres, fmax(e_d, e_p));
        if (!(nres > fmax(P.refine_rel * fmax(e_d, e_p), P.refine_abs))) break;
        rq[0] = rqn[0]; rq[1] = rqn[1];
      }
    }
    rp = g_max(g, rp); rdm = g_max(g, rdm);
    const double tau = XPERS_TAU;
    const double a_pmax = (rp > tau) ? tau / rp : 1.0, a_dmax = (rdm > tau) ? tau / rdm : 1.0;
    const double quad = dd.xs[0] * dd.xs[0] + dd.xs[2] * dd.xs[2];
    double ph, dphi, dQd, gdx;
    {
      double f0 = 0.0, gd = 0.0, qd = 0.0;
      XUNR for (int i = 0; i < 4; i++) {
        const double xv = x.xs[i], dv = dd.xs[i], fv = -c.qw * pr(d::F_XDES + i);
        f0 += xv * (0.5 * c.qw * xv + fv); gd += (c.qw * xv + fv) * dv; qd += c.qw * dv * dv;
      }
      XUNR for (int a = 0; a < 2; a++) { const double xv = x.us[a], dv = dd.us[a]; f0 += xv * (0.5 * P.Rw * xv); gd += P.Rw * xv * dv; qd += P.Rw * dv * dv; }
      if (g.j == 0) { const double xv = x.dl, dv = dd.dl; f0 += xv * (0.5 * P.slw * xv); gd += P.slw * xv * dv; qd += P.slw * dv * dv; }
      ph = g_sum(g, f0 - mu * (log(prod) + log(prod2)));
      gdx = g_sum(g, gd);
      dphi = gdx - mu * g_sum(g, dsum);
      dQd = g_sum(g, qd);
    }
    XDBG("   dw %.1e kc %d apmax %.3e admax %.3e ddl %.3e dphi %.3e\n", dw, kc, a_pmax, a_dmax, dd.dl, dphi);
    // ---- filter line search
    double alpha = a_pmax;
    bool accepted = false, ftype = false;
    while (alpha >= P.alpha_min) {
      double pr1 = 1.0, pr2 = 1.0;
      bool pos = true;
      double smin = 1.0;
      const double aq = alpha * alpha * quad;
      XSLOTS for (int s = 0; s < nsn; s++) {
        const double et = alpha * pr(d::N_C + s);
        const double sl = pr(d::N_SLO + s) + et, sh = pr(d::N_SHI + s) - et;
        pos = pos && (sl > 0.0) && (sh > 0.0);
        smin = fmin(smin, fmin(sl, sh));
        const double t = (sl * pr(d::N_A + s)) * (sh * pr(d::N_B + s));
        if (s & 1) pr1 *= t; else pr2 *= t;
      }
      for (int rr = 0; rr < K; rr++) {
        const double st = row(rr, XR_S) + alpha * row(rr, XR_DE) + aq;
        pos = pos && (st > 0.0);
        smin = fmin(smin, st);
        pr1 *= st * row(rr, XR_RS);
      }
      const bool allpos = g_all(g, pos);
      const double dlog = g_sum(g, allpos ? log(pr1) + log(pr2) : 0.0);
      if (allpos) {
        const double tht = (1.0 - alpha) * th;
        const double dph = alpha * gdx + 0.5 * alpha * alpha * dQd - mu * dlog;
        const double pht = ph + dph;
        if (tht <= XPERS_THMAX && filter_ok(tht, pht)) {
          const bool sw = dphi < 0.0 && alpha * (dphi * dphi) > th;
          if (th <= XPERS_THMIN && sw) {
            if (dph - 10.0 * CBF_EPSM * fabs(ph) <= P.eta_phi * alpha * dphi) { accepted = true; ftype = true; }
          } else if (tht <= (1.0 - P.gamma_theta) * th || dph - 10.0 * CBF_EPSM * fabs(ph) <= -P.gamma_phi * th) {
            accepted = true; ftype = false;
          }
        }
      }
      if (accepted && g_any(g, smin < 1e-12)) {
        // trial slacks below the rounding of the O(1) numbers they come from: re-evaluate them exactly as the next pass A will
        XKV xt;
        XUNR for (int i = 0; i < 4; i++) { xt.xs[i] = x.xs[i] + alpha * dd.xs[i]; xt.lm[i] = (i < C) ? x.lm[i] + alpha * dd.lm[i] : 0.0; }
        XUNR for (int a = 0; a < 2; a++) xt.us[a] = x.us[a] + alpha * dd.us[a];
        xt.dl = x.dl + alpha * dd.dl;
        double xp2[4]; prev_state(xt, false, xp2);
        scatter<true>(d::N_G, xt, xp2);
        bool okp = true;
        XSLOTS for (int s = 0; s < nsn; s++) {
          const double e = pr(d::N_G + s);
          okp = okp && (e - slot_lo(P, true, s) > 0.0) && (slot_hi(P, true, s) - e > 0.0);
        }
        for (int rr = 0; rr < K; rr++) {
          const double a = xt.xs[0] - row(rr, XR_CX), b = xt.xs[2] - row(rr, XR_CY);
          okp = okp && (a * a + b * b + xt.dl - row_eps(XP, rr) > 0.0);
        }
        accepted = g_all(g, okp);
      }
      if (accepted) break;
      alpha *= 0.5;
    }
    XDBG("   alpha %.3e acc %d ftype %d ph %.10e\n", alpha, (int)accepted, (int)ftype, ph);
    if (!accepted) { status = 2; return 2; }
    if (!ftype) filter_add(P, (1.0 - P.gamma_theta) * th, ph - P.gamma_phi * th);
    XUNR for (int i = 0; i < 4; i++) { x.xs[i] += alpha * dd.xs[i]; x.pi[i] += alpha * dd.pi[i]; }
    XUNR for (int i = 0; i < C; i++) x.lm[i] += alpha * dd.lm[i];
    XUNR for (int a = 0; a < 2; a++) { x.us[a] += alpha * dd.us[a]; x.nu[a] += alpha * dd.nu[a]; }
    x.dl += alpha * dd.dl; x.eta += alpha * dd.eta;
    XPERS_ADPREV = a_dmax;
    nl_it = it + 1;
    // slacks at the new point (the next iteration's bump rule measures the neighbours' moves against them)
    for (int rr = 0; rr < K; rr++) {
      const double a = x.xs[0] - row(rr, XR_CX), b = x.xs[2] - row(rr, XR_CY);
      row(rr, XR_SP) = a * a + b * b + x.dl - row_eps(XP, rr);
    }
    return 0;
  }
};

// ===================================================================================================================
// One group = one extended-mode solve; the groups of a team run the NLP stage in lock step (one exchange per iteration).
// `b` = instance index (team-major: b = team * T + member), valid per TEAM.  Shared by the CUDA kernel and tests/emu.
// ===================================================================================================================
#ifndef CBF_HOST_EMU
XDEV void atomicAdd_int(int* p, int v) { atomicAdd(p, v); }
#else
XDEV void atomicAdd_int(int* p, int v) { *p += v; }
#endif
template <int W> XDEV double t_max(const XGrp<W>& g, double v, int T) {   // maximum over the team (v already uniform in each group)
  for (int m = W; m < T * W; m <<= 1) v = fmax(v, x_shfl(g.tmask, v, g.tbase + ((x_lane() - g.tbase) ^ m)));
  return v;
}

template <int C, int W>
XDEV void ext_store_bad(const ExtIO& io, int b, int j) {
  if (j != 0) return;
  if (io.status) io.status[b] = 4;
  if (io.kkt) io.kkt[b] = 0.0;
  if (io.qp_iters) io.qp_iters[b] = 0;
  if (io.nlp_iters) io.nlp_iters[b] = 0;
  if (io.slack) io.slack[b] = 0.0;
  if (io.team_iters) io.team_iters[b] = 0;
}

// Registers of a lane that survive between the launches of a resumable run (agent-sharded teams: one launch per iteration, an
// all-gather of the published trajectories in between); stored behind the private columns in io.state.
template <int C, int W>
struct XCarry {
  double abuf_j, pubx, puby, moved, moved_prev;
  int badin, fin, team_it, stepped_prev, team_done;
};
template <int C, int W>
XDEV void ext_save(const ExtIO& io, int b, int npriv, const XSolver<C, W>& S, const XCarry<C, W>& cy) {
  double* st = io.state + (size_t)b * W + S.g.j;
  const size_t sd = (size_t)io.ld * W;
  for (int f = 0; f < npriv; f++) st[(size_t)f * sd] = S.pr(f);
  double* r = st + (size_t)npriv * sd;
  int q = 0;
  XUNR for (int i = 0; i < 4; i++) { r[(q++) * sd] = S.x.xs[i]; r[(q++) * sd] = S.x.lm[i]; r[(q++) * sd] = S.x.pi[i]; }
  XUNR for (int a = 0; a < 2; a++) { r[(q++) * sd] = S.x.us[a]; r[(q++) * sd] = S.x.nu[a]; r[(q++) * sd] = S.c.Fl[a]; }
  XUNR for (int t = 0; t < 6; t++) r[(q++) * sd] = S.c.D[t];
  r[(q++) * sd] = S.x.dl; r[(q++) * sd] = S.x.eta; r[(q++) * sd] = S.c.qw; r[(q++) * sd] = S.mu; r[(q++) * sd] = S.kkt;
  r[(q++) * sd] = cy.abuf_j; r[(q++) * sd] = cy.pubx; r[(q++) * sd] = cy.puby; r[(q++) * sd] = cy.moved; r[(q++) * sd] = cy.moved_prev;
  // flags and counters, packed into two doubles' worth of integers each
  r[(q++) * sd] = (double)(S.status | (S.qp_flag << 4) | ((int)S.first_ic << 8) | ((int)S.hold << 9) | (cy.badin << 10) | (cy.fin << 11) |
                           (cy.stepped_prev << 12) | (cy.team_done << 13) | (S.nfilt << 16));
  r[(q++) * sd] = (double)S.qp_iters; r[(q++) * sd] = (double)S.nlp_iters; r[(q++) * sd] = (double)S.nl_it; r[(q++) * sd] = (double)S.n_fact;
  r[(q++) * sd] = (double)cy.team_it; r[(q++) * sd] = (double)S.mu_age;
}
template <int C, int W>
XDEV void ext_restore(const ExtIO& io, int b, int npriv, XSolver<C, W>& S, XCarry<C, W>& cy) {
  const double* st = io.state + (size_t)b * W + S.g.j;
  const size_t sd = (size_t)io.ld * W;
  for (int f = 0; f < npriv; f++) S.pr(f) = st[(size_t)f * sd];
  const double* r = st + (size_t)npriv * sd;
  int q = 0;
  XUNR for (int i = 0; i < 4; i++) { S.x.xs[i] = r[(q++) * sd]; S.x.lm[i] = r[(q++) * sd]; S.x.pi[i] = r[(q++) * sd]; }
  XUNR for (int a = 0; a < 2; a++) { S.x.us[a] = r[(q++) * sd]; S.x.nu[a] = r[(q++) * sd]; S.c.Fl[a] = r[(q++) * sd]; }
  XUNR for (int t = 0; t < 6; t++) S.c.D[t] = r[(q++) * sd];
  S.x.dl = r[(q++) * sd]; S.x.eta = r[(q++) * sd]; S.c.qw = r[(q++) * sd]; S.mu = r[(q++) * sd]; S.kkt = r[(q++) * sd];
  cy.abuf_j = r[(q++) * sd]; cy.pubx = r[(q++) * sd]; cy.puby = r[(q++) * sd]; cy.moved = r[(q++) * sd]; cy.moved_prev = r[(q++) * sd];
  const int fl = (int)r[(q++) * sd];
  S.status = fl & 15; S.qp_flag = (fl >> 4) & 15; S.first_ic = (fl >> 8) & 1; S.hold = (fl >> 9) & 1; cy.badin = (fl >> 10) & 1; cy.fin = (fl >> 11) & 1;
  cy.stepped_prev = (fl >> 12) & 1; cy.team_done = (fl >> 13) & 1; S.nfilt = fl >> 16;
  S.qp_iters = (int)r[(q++) * sd]; S.nlp_iters = (int)r[(q++) * sd]; S.nl_it = (int)r[(q++) * sd]; S.n_fact = (int)r[(q++) * sd];
  cy.team_it = (int)r[(q++) * sd]; S.mu_age = (int)r[(q++) * sd];
}

#ifndef CBF_HOST_EMU
#define XFUSED_WATCHDOG 40000000000ll   // cycles (~20 s) a member waits for a neighbour's flag before it gives up
// (flags are polled and raised with relaxed system-scope accesses; the ordering against the data comes from ONE fence.acq_rel.sys on
//  either side -- a release / acquire per access costs a system-scope membar each: 18 % of the stall samples of the first version)
XDEV int x_ld_flag(const int* p) { int v; asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
XDEV void x_st_flag(int* p, int v) { asm volatile("st.relaxed.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
XDEV double x_ld_peer(const double* p) { double v; asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
XDEV void x_st_peer(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
XDEV void x_fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
XDEV long long x_clock() { return clock64(); }
XDEV void x_sleep() { __nanosleep(200); }
XDEV void x_atomic_max(int* p, int v) { atomicMax(p, v); }
XDEV int x_atomic_exch(int* p, int v) { return atomicExch(p, v); }
#else
#define XFUSED_WATCHDOG 4000000ll   /* polls */
XDEV int x_ld_flag(const int* p) { return *(volatile const int*)p; }
XDEV void x_st_flag(int* p, int v) { *(volatile int*)p = v; }
XDEV double x_ld_peer(const double* p) { return *(volatile const double*)p; }
XDEV void x_st_peer(double* p, double v) { *(volatile double*)p = v; }
XDEV void x_fence_sys() { }
XDEV long long x_clock() { static long long polls = 0; return polls++; }   // (emulation: the watchdog counts polls)
XDEV void x_sleep() { xemu::yield(); }
XDEV void x_atomic_max(int* p, int v) { if (v > *p) *p = v; }
XDEV int x_atomic_exch(int* p, int v) { const int o = *p; *p = v; return o; }
#endif

template <int C, int W, bool FUSED = false>
XDEV void ext_run_group(const XParams& XP, const ExtIO& io, int b, bool valid, XPriv pr) {
  typedef XDims<C> d;
  constexpr int NW = d::NW;
  const KParams& P = XP.k;
  XSolver<C, W> S;
  S.g = xgrp_make<W>(io.T); S.pr = pr;
  const XGrp<W>& g = S.g;
  const int j = g.j;
  S.nsq = d::NSQ - (j == 0 ? 2 : 0); S.nsn = d::NS - (j == 0 ? 2 : 0);
  S.K_obs = io.K_obs; S.K = io.K_obs + io.n_nbr;
  S.status = 1; S.qp_iters = 0; S.qp_flag = 3; S.nlp_iters = 0; S.n_fact = 0; S.kkt = 0.0; S.hold = false; S.nl_it = 0; S.mu = 0.0; S.first_ic = true; S.nfilt = 0;
  S.sigma = 0.0; S.qp_it = 0; S.mu_age = 0;
  const size_t ld = (size_t)io.ld;
  const bool jacobi = io.n_nbr > 0;
  const int npriv = ext_npriv(S.K);
  XCarry<C, W> cy;
  cy.abuf_j = 0.0; cy.pubx = 0.0; cy.puby = 0.0; cy.moved = 0.0; cy.moved_prev = 0.0; cy.badin = 0; cy.fin = 0; cy.team_it = 0; cy.stepped_prev = 1; cy.team_done = 0;
  if (io.phase <= 1) {
    bool qp_done = true, badin = false;
    if (valid) {
      XUNR for (int i = 0; i < 4; i++) { pr(d::F_XDES + i) = io.xdes[(size_t)(4 * j + i) * ld + b]; pr(d::F_XIN + i) = io.x0[(size_t)i * ld + b]; }
      cy.abuf_j = (io.abuf && j < 4) ? io.abuf[(size_t)j * ld + b] : 0.0;
      const int dm = j >> 2;   // gait domain of this grid point
      double feet[8];
      XUNR for (int t = 0; t < 8; t++) feet[t] = io.feet[(size_t)(8 * dm + t) * ld + b];
      const int c0 = io.contact[b];
      double F[8];
      XUNR for (int a = 0; a < 2; a++) {   // stance matrix (MPC_dist.cpp:1238-1252)
        if (C == 4) { XUNR for (int i = 0; i < 4; i++) F[4 * a + i] = feet[4 * a + i]; }
        else {
          F[4 * a] = (c0 == 0) ? feet[4 * a + 1] : feet[4 * a]; F[4 * a + 1] = (c0 == 0) ? feet[4 * a + 2] : feet[4 * a + 3];
          F[4 * a + 2] = 0.0; F[4 * a + 3] = 0.0;
        }
        S.c.Fl[a] = F[4 * a + C - 1];
      }
      XUNR for (int t = 0; t < 6; t++) S.c.D[t] = 0.0;
      XUNR for (int k = 0; k < 2; k++) XUNR for (int a = 0; a < NW; a++) S.c.D[k * NW + a] = F[4 * k + a] - F[4 * k + C - 1];
      S.c.qw = (j == W - 1) ? P.Pw : P.Qw;
      double chk = 0.0;
      for (int r = 0; r < io.K_obs; r++) {
        const double ox = io.obs[(size_t)r * ld + b], oy = io.obs[(size_t)(io.K_obs + r) * ld + b];
        S.row(r, XR_CX) = ox; S.row(r, XR_CY) = oy;
        chk += 0.0 * ox + 0.0 * oy;
      }
      XUNR for (int i = 0; i < 4; i++) chk += 0.0 * pr(d::F_XDES + i) + 0.0 * pr(d::F_XIN + i);
      XUNR for (int t = 0; t < 8; t++) chk += 0.0 * F[t];
      badin = !g_all(g, chk == 0.0);
      if (badin) S.status = 4;
      else qp_done = S.qp_init(P);
    }
    // ---- QP stage (independent per instance)
    for (;;) {
      const bool act = valid && !badin && !qp_done;
      if (!g_any(g, act)) break;
      qp_done = S.qp_iter(P);
    }
    cy.badin = badin ? 1 : 0;
    cy.fin = (!valid || badin) ? 1 : 0;
    if (valid && !badin && io.xqp) {
      XUNR for (int i = 0; i < 4; i++) io.xqp[(size_t)(4 * j + i) * ld + b] = S.x.xs[i];
      XUNR for (int a = 0; a < 2; a++) io.xqp[(size_t)(4 * W + 2 * j + a) * ld + b] = S.x.us[a];
    }
    // published trajectory: the QP positions first (a member with bad input stays where it was measured)
    if (valid) { cy.pubx = badin ? pr(d::F_XIN + 0) : S.x.xs[0]; cy.puby = badin ? pr(d::F_XIN + 2) : S.x.xs[2]; }
  } else if (valid) {
    ext_restore<C, W>(io, b, npriv, S, cy);
  }
  const bool resumable = io.phase != 0;
  const int tix = (io.xmode >= 3) ? b % io.n_teams : b;                  // team index / column inside a member block
  const int mloc = (io.xmode >= 3) ? b / io.n_teams : 0;                // member block of this launch
  const size_t mstride = (size_t)(2 * W + 2) * (size_t)io.n_teams;      // one member block of the published trajectories
  auto write_pub = [&](int stepped, double mv) {   // send buffer: positions, then (stepped, moved) of the iteration just done
    if (!valid || !io.pub) return;
    double* pb = io.pub + (size_t)mloc * mstride;
    const size_t pl = (size_t)io.n_teams;
    pb[(size_t)(2 * j) * pl + tix] = cy.pubx; pb[(size_t)(2 * j + 1) * pl + tix] = cy.puby;
    if (j == 0) { pb[(size_t)(2 * W) * pl + tix] = (double)stepped; pb[(size_t)(2 * W + 1) * pl + tix] = mv; }
  };
  auto nbr_block = [&](int i) {   // published trajectory of this instance's i-th neighbour (xmode 3)
    const int a = io.member0 + mloc, bb = i < a ? i : i + 1;
    return io.gathered + (size_t)bb * mstride;
  };
  if (io.phase == 1) { write_pub(1, 0.0); if (valid) ext_save<C, W>(io, b, npriv, S, cy); return; }
  double nbx[3] = {0.0, 0.0, 0.0}, nby[3] = {0.0, 0.0, 0.0};
  // ---- xmode 4 (FUSED): published blocks go straight into the mailboxes of the group's ranks; sequence flags instead of a collective
  int fseq = 0; bool fst = false; double fmv = 0.0;   // publications so far; neighbours' "stepped" / "distance still to go" of the last exchange
  const int t_all = io.n_nbr + 1;
  const size_t fl_off = (size_t)2 * t_all * mstride;   // flags follow the two parities of member blocks
  auto fused_pub = [&](int stepped, double mv) {
    fseq++;
    const int par = fseq & 1, a = io.member0 + mloc;
    if (valid) {
      const size_t off = (size_t)(par * t_all + a) * mstride, pl = (size_t)io.n_teams;
      for (int r = 0; r < io.n_grp; r++) {
        double* pb = io.mbox[r] + off;
        x_st_peer(pb + (size_t)(2 * j) * pl + tix, cy.pubx); x_st_peer(pb + (size_t)(2 * j + 1) * pl + tix, cy.puby);
        if (j == 0) { x_st_peer(pb + (size_t)(2 * W) * pl + tix, (double)stepped); x_st_peer(pb + (size_t)(2 * W + 1) * pl + tix, mv); }
      }
      x_fence_sys();
    }
    (void)g_all(g, true);   // every lane of the instance has written and fenced its rows before lane 0 raises the flags
    if (valid && j == 0) {   // (the other lanes' fences have completed before they arrived at the vote: their rows are visible)
      for (int r = 0; r < io.n_grp; r++) {
        int* fl = reinterpret_cast<int*>(io.mbox[r] + fl_off);
        x_st_flag(fl + (size_t)(par * t_all + a) * io.n_teams + tix, fseq);
      }
    }
  };
  auto fused_exchange = [&]() {   // wait for publication number fseq of every neighbour, then read it from the own mailbox
    bool timed_out = false;
    fst = false; fmv = 0.0;
    if (valid) {
      const int par = fseq & 1, a = io.member0 + mloc;
      const double* loc = io.mbox[io.member0 / (t_all / io.n_grp)];   // this rank's own mailbox (ranks hold t_all / n_grp members each)
      const int* fl = reinterpret_cast<const int*>(loc + fl_off);
      const size_t pl = (size_t)io.n_teams;
      for (int i = 0; i < io.n_nbr; i++) {
        const int bb = i < a ? i : i + 1;
        const int* f = fl + (size_t)(par * t_all + bb) * io.n_teams + tix;
        const long long t0 = x_clock();
        while (x_ld_flag(f) < fseq) {   // (a member with a bad contact pattern is published "for good" by the classify kernel: INT_MAX)
          x_sleep();
          if (x_clock() - t0 > XFUSED_WATCHDOG || x_ld_flag(io.fused_err) != 0) { timed_out = true; break; }
        }
      }
      x_fence_sys();   // all flags were seen: the rows behind them are visible to the loads below (one fence per exchange)
      for (int i = 0; i < io.n_nbr; i++) {
        const int bb = i < a ? i : i + 1;
        const double* nb = loc + (size_t)(par * t_all + bb) * mstride;
        nbx[i] = x_ld_peer(nb + (size_t)(2 * j) * pl + tix); nby[i] = x_ld_peer(nb + (size_t)(2 * j + 1) * pl + tix);
        fst = fst || x_ld_peer(nb + (size_t)(2 * W) * pl + tix) != 0.0;
        fmv = fmax(fmv, x_ld_peer(nb + (size_t)(2 * W + 1) * pl + tix));
      }
    }
    if (g_any(g, timed_out)) {   // a neighbour never published: give up (the host reports the error), leave the loops
      if (valid && j == 0) x_atomic_exch(io.fused_err, 1);
      cy.team_done = 1; cy.fin = 1; fst = false; fmv = 0.0;
    }
  };
  if (FUSED) fused_pub(1, 0.0);   // the QP positions (phase 1 of the resumable path)
  auto exchange = [&]() {
    if (FUSED) { fused_exchange(); return; }
    if (io.xmode == 1) {
      for (int i = 0; i < io.n_nbr; i++) { const int bb = i < g.a ? i : i + 1; nbx[i] = t_from(g, cy.pubx, bb); nby[i] = t_from(g, cy.puby, bb); }
    } else if (io.xmode == 2 && valid) {
      for (int i = 0; i < io.n_nbr; i++) { nbx[i] = io.nbr[i][(size_t)(2 * j) * ld + b]; nby[i] = io.nbr[i][(size_t)(2 * j + 1) * ld + b]; }
    } else if (io.xmode == 3 && valid) {
      for (int i = 0; i < io.n_nbr; i++) {
        const double* nb = nbr_block(i);
        nbx[i] = nb[(size_t)(2 * j) * io.n_teams + tix]; nby[i] = nb[(size_t)(2 * j + 1) * io.n_teams + tix];
      }
    }
  };
  auto publish = [&]() {   // pos <- pos + omega (iterate - pos); moved = distance still to go
    if (valid && !cy.badin) {
      cy.pubx += XP.relax * (S.x.xs[0] - cy.pubx); cy.puby += XP.relax * (S.x.xs[2] - cy.puby);
      cy.moved = g_max(g, fmax(fabs(S.x.xs[0] - cy.pubx), fabs(S.x.xs[2] - cy.puby)));
    }
  };
  if (io.phase == 0 || io.phase == 2) {
    exchange();
    if (!cy.fin) { S.set_nbr(nbx, nby); cy.fin = S.nlp_init(XP) ? 1 : 0; }
    cy.moved_prev = 0.0;
    publish();
    if (io.phase == 2) {   // the first resumed launch counts its (vacuous) "previous iteration stepped" check: start one below
      cy.team_it = -1; cy.stepped_prev = 1;
      write_pub(1, 0.0); if (valid) ext_save<C, W>(io, b, npriv, S, cy); return;
    }
    if (FUSED) { cy.team_it = -1; cy.stepped_prev = 1; fused_pub(1, 0.0); }   // as phase 2: same counts as the resumable path
  }
  // ---- iterations: [exchange; team finished (flags of the previous iteration)?; iterate; publish]
  bool first = io.phase == 0 && !FUSED;
  if (io.phase != 4) for (;;) {
    exchange();
    if (!first && !cy.team_done) {
      bool more;
      if (FUSED) {
        more = (cy.stepped_prev != 0 || fst) || fmax(cy.moved_prev, fmv) > XP.pub_tol;
      } else if (io.xmode == 1) {
        more = t_any(g, cy.stepped_prev != 0) || t_max(g, cy.moved_prev, io.T) > XP.pub_tol;
      } else if (io.xmode == 3) {
        bool st = cy.stepped_prev != 0; double mv = cy.moved_prev;
        if (valid) for (int i = 0; i < io.n_nbr; i++) {
          const double* nb = nbr_block(i);
          st = st || nb[(size_t)(2 * W) * io.n_teams + tix] != 0.0; mv = fmax(mv, nb[(size_t)(2 * W + 1) * io.n_teams + tix]);
        }
        more = st || mv > XP.pub_tol;
      } else {
        more = cy.stepped_prev != 0;
      }
      if (more) { cy.team_it++; if (cy.team_it >= XP.max_team_it) more = false; }
      if (!more) cy.team_done = 1;
    }
    first = false;
    if (cy.team_done && !resumable) break;
    int stepped = 0;
    if (!cy.team_done && !cy.fin) {
      S.set_nbr(nbx, nby);
      const int rc = S.nlp_iter(XP, jacobi);
      stepped = rc == 0;
      if (rc == 2) cy.fin = 1;
    }
    cy.stepped_prev = stepped; cy.moved_prev = cy.moved;
    if (!cy.team_done) publish();
    if (FUSED) fused_pub(stepped, cy.moved_prev);
    if (resumable) {
      write_pub(stepped, cy.moved_prev);
      if (valid && !cy.team_done && j == 0 && io.n_active) atomicAdd_int(io.n_active, 1);
      break;
    }
  }
  if (resumable && !(io.phase == 4)) { if (valid) ext_save<C, W>(io, b, npriv, S, cy); return; }
  if (!valid) return;
  if (cy.badin) { ext_store_bad<C, W>(io, b, j); return; }
  const double abuf_j = cy.abuf_j, pubx = cy.pubx, puby = cy.puby;
  const int team_it = cy.team_it;
  if (FUSED && j == 0 && io.fused_its) x_atomic_max(io.fused_its, team_it);
  // ---- store
  if (io.X) { XUNR for (int i = 0; i < 4; i++) io.X[(size_t)(4 * j + i) * ld + b] = S.x.xs[i]; }
  if (io.U) { XUNR for (int a = 0; a < 2; a++) io.U[(size_t)(2 * j + a) * ld + b] = S.x.us[a]; }
  if (io.lam) {
    XUNR for (int i = 0; i < C; i++) io.lam[(size_t)(C * j + i) * ld + b] = S.x.lm[i];
    XUNR for (int i = 0; i < 4 - C; i++) io.lam[(size_t)(W * C + (4 - C) * j + i) * ld + b] = 0.0;
  }
  if (io.pub && !resumable) { io.pub[(size_t)(2 * j) * ld + b] = pubx; io.pub[(size_t)(2 * j + 1) * ld + b] = puby; }
  (void)pubx; (void)puby;
  {  // Bezier fit of the first gait domain (MPC_dist.cpp:784-855): lane r < 4 fits state row r through x_1 .. x_4
    double Y[5];
    Y[0] = abuf_j;
    XUNR for (int st = 0; st < 4; st++) {
      double v = 0.0;
      XUNR for (int t = 0; t < 4; t++) { const double w = g_idx(g, S.x.xs[t], st); if (t == j) v = w; }
      Y[1 + st] = v;
    }
    if (j < 4) {
      XUNR for (int jj = 0; jj < 5; jj++) {
        double s = 0.0;
        XUNR for (int i = 0; i < 5; i++) s += P.minv[5 * jj + i] * Y[i];
        if (io.alpha) io.alpha[(size_t)(5 * j + jj) * ld + b] = s;
      }
      if (io.abuf_next) io.abuf_next[(size_t)j * ld + b] = Y[4];
    }
  }
  if (j == 0) {
    if (io.slack) io.slack[b] = S.x.dl;
    if (io.kkt) io.kkt[b] = S.kkt;
    if (io.status) io.status[b] = S.status;
    if (io.qp_iters) io.qp_iters[b] = S.qp_iters;
    if (io.nlp_iters) io.nlp_iters[b] = S.nlp_iters;
    if (io.team_iters) io.team_iters[b] = team_it;
  }
}

// xmode 4: an instance that no lane will run (bad contact pattern, status 4) still has team mates waiting for its publications: the
// classify kernel publishes it once and for all -- its measured position at every grid point, "did not step", nothing left to go --
// into both parities of every mailbox of the group, with flags that satisfy every wait.
XDEV void ext_fused_publish_bad(const ExtIO& io, int b) {
  const int W = io.N, t_all = io.n_nbr + 1;
  const int tix = b % io.n_teams, a = io.member0 + b / io.n_teams;
  const size_t mstride = (size_t)(2 * W + 2) * (size_t)io.n_teams, pl = (size_t)io.n_teams, ld = (size_t)io.ld;
  const double px = io.x0[b], py = io.x0[2 * ld + b];
  for (int r = 0; r < io.n_grp; r++)
    for (int par = 0; par < 2; par++) {
      double* pb = io.mbox[r] + (size_t)(par * t_all + a) * mstride;
      for (int k = 0; k < W; k++) { x_st_peer(pb + (size_t)(2 * k) * pl + tix, px); x_st_peer(pb + (size_t)(2 * k + 1) * pl + tix, py); }
      x_st_peer(pb + (size_t)(2 * W) * pl + tix, 0.0); x_st_peer(pb + (size_t)(2 * W + 1) * pl + tix, 0.0);
    }
  x_fence_sys();
  for (int r = 0; r < io.n_grp; r++)
    for (int par = 0; par < 2; par++) {
      int* fl = reinterpret_cast<int*>(io.mbox[r] + (size_t)2 * t_all * mstride);
      x_st_flag(fl + (size_t)(par * t_all + a) * io.n_teams + tix, 0x7fffffff);
    }
}

// contact pattern -> number of stance feet (4, 2 or -1) of a TEAM: every member must have the same pattern class
XDEV int ext_classify_team(const ExtIO& io, int t) {
  const size_t ld = (size_t)io.ld;
  int cls = 0;
  for (int a = 0; a < io.T; a++) {
    const int b = t * io.T + a;
    const int c0 = io.contact[b], csum = c0 + io.contact[ld + b] + io.contact[2 * ld + b] + io.contact[3 * ld + b];
    const int ca = (csum == 4) ? 4 : ((csum == 2 && (c0 == 0 || c0 == 1)) ? 2 : -1);
    if (a == 0) cls = ca;
    else if (ca != cls) cls = -1;
  }
  return cls;
}
