// cbfnmpc_host.h -- host-side helpers shared by the C-ABI implementation (cbfnmpc_api.cu) and the tests/emu
// harness: defaults, and the constant model / Bernstein data handed to the kernel.
#pragma once
#include <math.h>
#include <string.h>
#include "../../include/cbfnmpc.h"
#include "cbfnmpc_types.h"

static inline void host_default_gains(cbfnmpc_gains* g) {
  g->Qw = 3e2; g->Rw = 1e-1; g->Pw = 2e3; g->slack_w = 0.3e4;            // src/MPC_dist.cpp:172-175
  g->mu_fric = 0.7; g->h_com = 0.29; g->g_acc = 9.81; g->Ts = 43 * 0.001; // :90, :99-100, :104 (TSOPTTICK 43 ms)
  g->eps_dist = (double)1.9f;                                            // dec_vars_constr_cost.h:401 (float literal)
  g->v_max = (double)0.35f;                                              // :306-307 (float literal)
  g->box = 1e3;                                                          // src/MPC_dist.cpp:317-318
}

static inline void host_default_options(cbfnmpc_options* o) {
  o->qp_reltol = 1e-6; o->qp_abstol = 1e-6; o->qp_maxit = 25;            // iSWIFT GlobalOptions.h:23-26
  o->tol = 3e-9; o->mu_init = 0.1; o->kappa_eps = 10.0; o->kappa_mu = 0.2; o->tau_min = 0.99; o->push = 1e-2;  // tol: see DESIGN.md 4
  o->gamma_theta = 1e-5; o->gamma_phi = 1e-5; o->eta_phi = 1e-4; o->kappa_sigma = 1e10; o->alpha_min = 1e-10;
  o->dw_first = 1e-4; o->dw_min = 1e-20; o->kw_minus = 1.0 / 3.0; o->kw_plus = 8.0; o->kw_plus_first = 100.0;
  o->dw_max = 1e40; o->rho_lambda = 1e-6; o->z_div = 1e13;
  o->sigma_cap = 1e8; o->refine_rel = 1e-4; o->refine_abs = 1e-11;
  o->max_iter = 200; o->filter_cap = 8; o->max_refine = 2;
  o->acc_tol = 9e-9;  // BASELINE bar: KKT residual below 1e-8; see DESIGN.md 4
}

// LIP discretisation exactly as src/MPC_dist.cpp:117-127 (3rd-order Taylor of exp(A Ts); Bd = A^-1 (Ad - I) B)
static inline void host_lip(const cbfnmpc_gains* g, double Ad[16], double Bd[8]) {
  const double w = g->g_acc / g->h_com, Ts = g->Ts;
  double A[16] = {0}, A2[16], A3[16];
  A[1] = 1; A[4] = w; A[11] = 1; A[14] = w;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += A[4 * i + k] * A[4 * k + j]; A2[4 * i + j] = s; }
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += A2[4 * i + k] * A[4 * k + j]; A3[4 * i + j] = s; }
  for (int i = 0; i < 16; i++) Ad[i] = ((i % 5 == 0) ? 1.0 : 0.0) + A[i] * Ts + 0.5 * A2[i] * Ts * Ts + A3[i] * Ts * Ts * Ts / 6.0;
  double Ainv[16] = {0}, M[16], T[16];
  Ainv[1] = 1.0 / w; Ainv[4] = 1; Ainv[11] = 1.0 / w; Ainv[14] = 1;
  for (int i = 0; i < 16; i++) M[i] = Ad[i] - ((i % 5 == 0) ? 1.0 : 0.0);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double s = 0; for (int k = 0; k < 4; k++) s += Ainv[4 * i + k] * M[4 * k + j]; T[4 * i + j] = s; }
  for (int i = 0; i < 4; i++) { Bd[2 * i] = T[4 * i + 1] * (-w); Bd[2 * i + 1] = T[4 * i + 3] * (-w); }
}

// inverse of the 5x5 Bernstein collocation matrix M_ij = C(4,j) t_i^j (1-t_i)^(4-j), t_i = i/4 (math_define.h:409-414)
static inline void host_bernstein_inv(double Minv[25]) {
  static const double binom[5] = {1, 4, 6, 4, 1};
  double Aug[5][10];
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 10; j++)
      Aug[i][j] = j < 5 ? binom[j] * pow(i * 0.25, j) * pow(1 - i * 0.25, 4 - j) : (j - 5 == i ? 1.0 : 0.0);
  for (int k = 0; k < 5; k++) {
    int piv = k;
    for (int i = k + 1; i < 5; i++) if (fabs(Aug[i][k]) > fabs(Aug[piv][k])) piv = i;
    for (int j = 0; j < 10; j++) { double t = Aug[k][j]; Aug[k][j] = Aug[piv][j]; Aug[piv][j] = t; }
    double d = Aug[k][k];
    for (int j = 0; j < 10; j++) Aug[k][j] /= d;
    for (int i = 0; i < 5; i++) if (i != k) { double f = Aug[i][k]; for (int j = 0; j < 10; j++) Aug[i][j] -= f * Aug[k][j]; }
  }
  for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) Minv[5 * i + j] = Aug[i][5 + j];
}

static inline void host_fill_kparams(const cbfnmpc_gains* g, const cbfnmpc_options* o, KParams* P) {
  double Ad[16], Bd[8];
  host_lip(g, Ad, Bd);
  P->a11 = Ad[0]; P->a12 = Ad[1]; P->a21 = Ad[4]; P->a22 = Ad[5]; P->bp = Bd[0]; P->bv = Bd[2];
  P->Qw = g->Qw; P->Pw = g->Pw; P->Rw = g->Rw; P->slw = g->slack_w;
  P->fb = g->mu_fric * g->h_com / sqrt(2.0);  // src/MPC_dist.cpp:315
  P->box = g->box; P->vmax = g->v_max; P->eps = g->eps_dist;
  host_bernstein_inv(P->minv);
  P->qp_reltol = o->qp_reltol; P->qp_abstol = o->qp_abstol; P->qp_maxit = o->qp_maxit;
  P->tol = o->tol; P->mu_init = o->mu_init; P->kappa_eps = o->kappa_eps; P->kappa_mu = o->kappa_mu; P->tau_min = o->tau_min;
  P->push = o->push; P->gamma_theta = o->gamma_theta; P->gamma_phi = o->gamma_phi; P->eta_phi = o->eta_phi;
  P->kappa_sigma = o->kappa_sigma; P->alpha_min = o->alpha_min; P->dw_first = o->dw_first; P->dw_min = o->dw_min;
  P->kw_minus = o->kw_minus; P->kw_plus = o->kw_plus; P->kw_plus_first = o->kw_plus_first; P->dw_max = o->dw_max;
  P->rho_lambda = o->rho_lambda; P->z_div = o->z_div; P->sigma_cap = o->sigma_cap; P->refine_rel = o->refine_rel;
  P->refine_abs = o->refine_abs; P->max_iter = o->max_iter; P->filter_cap = o->filter_cap > 8 ? 8 : o->filter_cap;
  P->max_refine = o->max_refine; P->acc_tol = o->acc_tol > o->tol ? o->acc_tol : o->tol;
  for (int k = 0; k < 2; k++) {  // slots: X 4 | U 2 | lambda C | QP: friction 2 ; NLP: velocity 2 | distance 1 | friction 2
    const int C = k == 0 ? 4 : 2;
    for (int s = 0; s < 16; s++) {
      double lo = 0.0, hi = 0.0, lon = 0.0, hin = 0.0;
      if (s < 6) { lo = lon = -P->box; hi = hin = P->box; }
      else if (s < 6 + C) { lo = lon = 0.0; hi = hin = 1.0; }
      if (s >= 6 + C && s < 8 + C) { lo = -P->fb; hi = P->fb; lon = -P->vmax; hin = P->vmax; }
      if (s == 8 + C) { lon = P->eps; hin = 0.0; }
      if (s >= 9 + C && s < 11 + C) { lon = -P->fb; hin = P->fb; }
      P->slot_lo[0][k][s] = lo; P->slot_hi[0][k][s] = hi;
      P->slot_lo[1][k][s] = lon; P->slot_hi[1][k][s] = hin;
    }
  }
}
