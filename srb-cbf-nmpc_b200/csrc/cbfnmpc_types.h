// cbfnmpc_types.h -- plain data shared by the kernels and the host side: solver constants and the batch I/O descriptor.
#pragma once

struct KParams {
  double a11, a12, a21, a22, bp, bv;  // per-axis LIP: A2 = [[a11,a12],[a21,a22]], B2 = [bp; bv]  (MPC_dist.cpp:117-127)
  double Qw, Pw, Rw, slw, fb, box, vmax, eps;
  double minv[25];                    // inverse Bernstein matrix (row-major)                     (MPC_dist.cpp:803-807)
  double qp_reltol, qp_abstol;
  double tol, mu_init, kappa_eps, kappa_mu, tau_min, push, gamma_theta, gamma_phi, eta_phi, kappa_sigma, alpha_min;
  double dw_first, dw_min, kw_minus, kw_plus, kw_plus_first, dw_max, rho_lambda, z_div, sigma_cap, refine_rel, refine_abs;
  int qp_maxit, max_iter, filter_cap, max_refine;
  double acc_tol;
  // slot bounds by [stage: 0 QP, 1 NLP][class: 0 = 4 stance feet, 1 = 2][slot] (slot order of cbfnmpc_lane.h, LaneDims)
  double slot_lo[2][2][16], slot_hi[2][2][16];
};

#define CBF_EPSM 2.220446049250313e-16
#define CBF_INF (__builtin_huge_val())

struct BatchIO {
  // inputs, SoA [field][B]
  const double *x0, *xdes, *feet; const int* contact; const double* obs; int n_obs; const double* abuf;
  const double* xstart;  // optional [41][B]: skip the QP stage and start the NLP here (dense nlpTuple entry point)
  // outputs, SoA
  double *X, *U, *lam, *slack, *xqp, *alpha, *abuf_next, *kkt;
  int *status, *qp_iters, *nlp_iters, *n_fact;
  unsigned long long* active;
  int B, mode;
  int ld;  // leading dimension of every array (>= B; > B when the batch is a column range of a larger SoA block)
  // host-buffer calls (cbfnmpc_lane_kernel_late): the obstacles travel on a second stream WHILE the QP stage runs; a word copied
  // behind them in that stream is what the kernel waits for before it looks at them (obs_flag == nullptr: they are there)
  const int* obs_flag; int obs_seq;
};
