/* TEST INFRASTRUCTURE -- see cbfnmpc_oracle.h for scope and pinning status ("parity unpinned" for the NLP stage).
 * Plain C99, no dependencies.  Dense, structure-agnostic on purpose: it must not share the structured
 * (Riccati) algebra of the CUDA path it checks. */
#include "cbfnmpc_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define EPSM 2.220446049250313e-16
#define NGRID 4
static const int VEL_IDX[8] = {1, 5, 9, 13, 3, 7, 11, 15}; /* include/dec_vars_constr_cost.h:270-279 */

void orc_default_options(orc_options *o) {
  o->qp_reltol = 1e-6; o->qp_abstol = 1e-6; o->qp_maxit = 25;
  o->tol = 3e-9; o->mu_init = 0.1; o->kappa_eps = 10.0; o->kappa_mu = 0.2; o->theta_mu = 1.5; o->tau_min = 0.99;
  o->push = 1e-2; o->gamma_theta = 1e-5; o->gamma_phi = 1e-5; o->delta = 1.0; o->s_theta = 1.1; o->s_phi = 2.3;
  o->eta_phi = 1e-4; o->kappa_sigma = 1e10; o->alpha_min = 1e-10;
  o->dw_first = 1e-4; o->dw_min = 1e-20; o->kw_minus = 1.0 / 3.0; o->kw_plus = 8.0; o->kw_plus_first = 100.0;
  o->dw_max = 1e40; o->rho_lambda = 1e-6; o->z_div = 1e13; o->acc_tol = 9e-9;
  o->max_iter = 200; o->filter_cap = 8;
}

void orc_default_gains(orc_gains *g) {
  g->Qw = 3e2; g->Rw = 1e-1; g->Pw = 2e3; g->slack_w = 0.3e4;             /* src/MPC_dist.cpp:172-175 */
  g->mu_fric = 0.7; g->h_com = 0.29; g->g_acc = 9.81; g->Ts = 43 * 0.001; /* :90, :99-100, :104 */
  g->eps_dist = (double)1.9f;                                             /* dec_vars_constr_cost.h:401 */
  g->v_max = (double)0.35f;                                               /* :306-307 */
  g->box = 1e3;                                                           /* src/MPC_dist.cpp:317-318 */
}

/* ------------------------------------------------------------------------------------------------ */
/* small dense helpers (column-major)                                                               */
/* ------------------------------------------------------------------------------------------------ */
static void mat4_mul(const double *A, const double *B, double *C) { /* row-major 4x4 */
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += A[4 * i + k] * B[4 * k + j];
      C[4 * i + j] = s;
    }
}

void orc_lip_discrete(const orc_gains *g, double *Ad, double *Bd) {
  /* Ad = I + A Ts + 1/2 A^2 Ts^2 + 1/6 A^3 Ts^3 ; Bd = A^-1 (Ad - I) B   (src/MPC_dist.cpp:117-127) */
  double w = g->g_acc / g->h_com, Ts = g->Ts;
  double A[16] = {0}, A2[16], A3[16];
  A[0 * 4 + 1] = 1; A[1 * 4 + 0] = w; A[2 * 4 + 3] = 1; A[3 * 4 + 2] = w;
  mat4_mul(A, A, A2); mat4_mul(A2, A, A3);
  for (int i = 0; i < 16; i++)
    Ad[i] = ((i % 5 == 0) ? 1.0 : 0.0) + A[i] * Ts + 0.5 * A2[i] * Ts * Ts + A3[i] * Ts * Ts * Ts / 6.0;
  /* A^-1 = [[0,1/w,0,0],[1,0,0,0],[0,0,0,1/w],[0,0,1,0]] ; B = [[0,0],[-w,0],[0,0],[0,-w]] */
  double Ainv[16] = {0}, M[16], T[16];
  Ainv[0 * 4 + 1] = 1.0 / w; Ainv[1 * 4 + 0] = 1; Ainv[2 * 4 + 3] = 1.0 / w; Ainv[3 * 4 + 2] = 1;
  for (int i = 0; i < 16; i++) M[i] = Ad[i] - ((i % 5 == 0) ? 1.0 : 0.0);
  mat4_mul(Ainv, M, T);
  for (int i = 0; i < 4; i++) {
    Bd[2 * i + 0] = T[4 * i + 1] * (-w);
    Bd[2 * i + 1] = T[4 * i + 3] * (-w);
  }
}

int orc_stance_matrix(const double *feet, const int *contact, double *F) {
  int c = contact[0] + contact[1] + contact[2] + contact[3];
  if (c == 4) {
    for (int i = 0; i < 4; i++) { F[i] = feet[i]; F[4 + i] = feet[4 + i]; }
    return 4;
  }
  if (c == 2 && contact[0] == 0) { /* 0 1 1 0  -> FL, RR (src/MPC_dist.cpp:1243-1247) */
    F[0] = feet[1]; F[1] = feet[2]; F[2] = feet[4 + 1]; F[3] = feet[4 + 2];
    return 2;
  }
  if (c == 2 && contact[0] == 1) { /* 1 0 0 1 -> FR, RL (:1248-1252) */
    F[0] = feet[0]; F[1] = feet[3]; F[2] = feet[4 + 0]; F[3] = feet[4 + 3];
    return 2;
  }
  return -1;
}

void orc_pack_dense(const orc_gains *g, const double *x0, const double *xdes, const double *F, int c, double *Q,
                    double *f, double *Aeq, double *beq, double *G, double *h) {
  const int n = 4, m = 2, N = NGRID;
  const int nv = n * N + m * N + N * c + 1, p = n * N + m * N + N, mi = 2 * m * (N - 1) + 2 * n * N + 2 * m * N + 2 * N * c;
  double Ad[16], Bd[8];
  orc_lip_discrete(g, Ad, Bd);
  memset(Q, 0, sizeof(double) * nv * nv); memset(f, 0, sizeof(double) * nv);
  memset(Aeq, 0, sizeof(double) * p * nv); memset(beq, 0, sizeof(double) * p);
  memset(G, 0, sizeof(double) * mi * nv);
#define QQ(i, j) Q[(i) + (j) * nv]
#define AE(i, j) Aeq[(i) + (j) * p]
#define GG(i, j) G[(i) + (j) * mi]
  for (int i = 0; i < n * N; i++) {
    double w = (i >= n * (N - 1)) ? g->Pw : g->Qw; /* :193-197 */
    QQ(i, i) = w;
    f[i] = -xdes[i] * w; /* :209 */
  }
  for (int i = 0; i < m * N; i++) QQ(n * N + i, n * N + i) = g->Rw;
  QQ(nv - 1, nv - 1) = g->slack_w; /* :205 */
  /* equalities :232-261 */
  for (int k = 0; k < N; k++) {
    for (int i = 0; i < n; i++) {
      AE(n * k + i, n * k + i) = -1.0;
      if (k >= 1)
        for (int j = 0; j < n; j++) AE(n * k + i, n * (k - 1) + j) += Ad[4 * i + j];
      for (int j = 0; j < m; j++) AE(n * k + i, n * N + m * k + j) = Bd[2 * i + j];
    }
    for (int a = 0; a < m; a++) {
      AE(n * N + m * k + a, n * N + m * k + a) = 1.0;
      for (int j = 0; j < c; j++) AE(n * N + m * k + a, n * N + m * N + c * k + j) = -F[a * c + j];
    }
    for (int j = 0; j < c; j++) AE(n * N + m * N + k, n * N + m * N + c * k + j) = 1.0;
    beq[n * N + m * N + k] = 1.0;
  }
  for (int i = 0; i < n; i++) {
    double s = 0;
    for (int j = 0; j < n; j++) s += Ad[4 * i + j] * x0[j];
    beq[i] = -s;
  }
  /* inequalities :278-321 */
  const int r = m * (N - 1);
  for (int i = 0; i < N - 1; i++)
    for (int a = 0; a < 2; a++) {
      GG(i * m + a, i * n + 2 * a) = 1.0;                   /* Gsubx: xportion picks (p_x, p_y) of state block i */
      GG(i * m + a, n * N + (i + 1) * m + a) = -1.0;        /* -Gsubu: control block i+1 */
      GG(r + i * m + a, i * n + 2 * a) = -1.0;
      GG(r + i * m + a, n * N + (i + 1) * m + a) = 1.0;
    }
  int o = 2 * r;
  for (int i = 0; i < n * N; i++) { GG(o + i, i) = 1.0; GG(o + n * N + i, i) = -1.0; }
  o += 2 * n * N;
  for (int i = 0; i < m * N; i++) { GG(o + i, n * N + i) = 1.0; GG(o + m * N + i, n * N + i) = -1.0; }
  o += 2 * m * N;
  for (int i = 0; i < N * c; i++) { GG(o + i, n * N + m * N + i) = -1.0; GG(o + N * c + i, n * N + m * N + i) = 1.0; }
  const double fb = g->mu_fric * g->h_com / sqrt(2.0);
  for (int i = 0; i < mi; i++) h[i] = 1.0;
  for (int i = 0; i < 2 * r; i++) h[i] = fb;
  for (int i = 2 * r; i < 2 * r + 2 * n * N + 2 * m * N; i++) h[i] = g->box;
  for (int i = 0; i < N * c; i++) h[2 * r + 2 * n * N + 2 * m * N + i] = 0.0;
#undef QQ
#undef AE
#undef GG
}

/* generateReferenceTrajectory (src/MPC_dist.cpp:930-1062): potential-field + Lennard-Jones recurrence for the 4 agents of one
 * team, `loop` (100000) steps of the hard-coded discrete double integrator, force ramp over the first 1000 steps, decimation
 * by 40.  Pstart: 8 (x,y of the 4 agents); Pobs: 2 x n_obs column-major (the PLANNING obstacles, setPobs); outputs Pr, Prd:
 * 8 x ncols column-major with ncols = loop / 40.  Column c of the outputs is column c+2 of the decimated state history
 * (:1044-1047); the decimation overwrites only the first ncols columns of the history (:1034-1037), so the last two output
 * columns are the UNdecimated steps ncols and ncols+1 -- kept, since that is what the reference produces. */
void orc_generate_reference(const double *Pstart, const double *Pobs, int n_obs, double goal_x, double goal_y, int loop,
                            double *Pr, double *Prd) {
  const double alpha = 150.0, eta = 400.0, dmin = 1.0, epsilon = 60, sigma = 1.0;               /* :932-933 */
  const double a13 = 0.001025061239872, a33 = 0.929527039758809;                                /* Ad, :942-945 */
  const double b1 = 0.000000068070472960, b3 = 0.000128132654983983;                            /* Bd, :957-960 */
  const int ramp = 1000, dec = 40, ncols = loop / dec;
  double q[16];
  double *hist = (double *)malloc(sizeof(double) * 16 * (size_t)(ncols + 2));  /* decimated history, columns 0 .. ncols+1 */
  for (int k = 0; k < 4; k++) { q[4 * k] = Pstart[2 * k]; q[4 * k + 1] = Pstart[2 * k + 1]; q[4 * k + 2] = 0; q[4 * k + 3] = 0; }
  for (int i = 0; i <= loop; i++) {
    /* record: column c of the decimated history is step 40 c for c < ncols, step c for c = ncols, ncols + 1 */
    if (i % dec == 0 && i / dec < ncols) memcpy(hist + 16 * (size_t)(i / dec), q, sizeof(q));
    if (i == ncols || i == ncols + 1) memcpy(hist + 16 * (size_t)i, q, sizeof(q));
    if (i == loop) break;
    double F[8];
    for (int k = 0; k < 4; k++) {
      const double px = q[4 * k], py = q[4 * k + 1];
      const double gx = px - goal_x, gy = py - goal_y;
      const double d_goal = sqrt(gx * gx + gy * gy);
      double fx = -alpha * (gx / d_goal), fy = -alpha * (gy / d_goal);                           /* F_att, :1000 */
      double rx = 0, ry = 0;
      for (int j = 0; j < n_obs; j++) {                                                          /* F_rep, :1003-1017 */
        const double ox = px - Pobs[2 * j], oy = py - Pobs[2 * j + 1];
        const double d = sqrt(ox * ox + oy * oy);
        if (d < dmin) {
          const double w = eta * (1 / d - 1 / dmin) * (1 / pow(d, 2));
          rx += w * (ox / d); ry += w * (oy / d);
        }
      }
      double ax = 0, ay = 0;
      for (int j = 0; j < 4; j++) {                                                              /* Lennard-Jones, :1019-1029 */
        if (j == k) continue;
        const double ox = px - q[4 * j], oy = py - q[4 * j + 1];
        const double d = sqrt(ox * ox + oy * oy);
        const double w = 4 * epsilon * ((6 * pow(sigma, 6)) / pow(d, 7) - (12 * pow(sigma, 12)) / pow(d, 13));
        ax = ax - w * (ox / d); ay = ay - w * (oy / d);
      }
      F[2 * k] = fx + rx + ax; F[2 * k + 1] = fy + ry + ay;
      if (d_goal < 0.001) { F[2 * k] = 0; F[2 * k + 1] = 0; }                                    /* :1033-1036 */
    }
    const double scale = (i < ramp) ? ((double)i / ramp) : 1.0;                                  /* :1039 */
    for (int k = 0; k < 4; k++) {                                                                /* q+ = Ae q + Be F scale, :1041 */
      const double x = q[4 * k], y = q[4 * k + 1], vx = q[4 * k + 2], vy = q[4 * k + 3];
      const double ux = F[2 * k] * scale, uy = F[2 * k + 1] * scale;
      q[4 * k] = x + a13 * vx + b1 * ux; q[4 * k + 1] = y + a13 * vy + b1 * uy;
      q[4 * k + 2] = a33 * vx + b3 * ux; q[4 * k + 3] = a33 * vy + b3 * uy;
    }
  }
  for (int c = 0; c < ncols; c++)
    for (int k = 0; k < 4; k++) {
      const double *h = hist + 16 * (size_t)(c + 2) + 4 * k;
      Pr[2 * k + 8 * (size_t)c] = h[0]; Pr[2 * k + 1 + 8 * (size_t)c] = h[1];
      Prd[2 * k + 8 * (size_t)c] = h[2]; Prd[2 * k + 1 + 8 * (size_t)c] = h[3];
    }
  free(hist);
}

/* copPlanner_eventbase (src/MPC_dist.cpp:702-782, effective semantics: use_bezier is false): columns [N g, N g + N) of rows
 * 2a, 2a+1 of Pr_refined_ / Prd_refined_ (8 x ref_cols, column-major) -> com_desired_Traj_vec_eventbased_ ([x,xd,y,yd] x N).
 * Columns past the end are clamped (the reference has no bound check). */
void orc_cop_planner(const double *Pr, const double *Prd, int ref_cols, int agent_id, int gait_domain, double *xdes16) {
  for (int k = 0; k < 4; k++) {
    int col = 4 * gait_domain + k;
    if (col > ref_cols - 1) col = ref_cols - 1;
    if (col < 0) col = 0;
    xdes16[4 * k + 0] = Pr[2 * agent_id + 8 * (size_t)col];      /* :715 */
    xdes16[4 * k + 1] = Prd[2 * agent_id + 8 * (size_t)col];     /* :716 */
    xdes16[4 * k + 2] = Pr[2 * agent_id + 1 + 8 * (size_t)col];  /* :717 */
    xdes16[4 * k + 3] = Prd[2 * agent_id + 1 + 8 * (size_t)col]; /* :718 */
  }
}

/* footholdsPlanner, first part (src/MPC_dist.cpp:1204-1216): toePos (3x4 column-major) -> feet_xy (x of FR,FL,RR,RL then y);
 * during the first two gait domains the nominal footprint about agent_Initial_ replaces the measured toes. */
void orc_footholds(const double *toe_pos12, int gait_domain, const double *agent_initial, double *feet_xy8) {
  static const double fp[4][2] = {{0.2188, -0.1320}, {0.2188, 0.1320}, {-0.1472, -0.1320}, {-0.1472, 0.1320}}; /* :1206-1209 */
  for (int i = 0; i < 4; i++) {
    feet_xy8[i] = (gait_domain <= 1) ? fp[i][0] + agent_initial[0] : toe_pos12[3 * i];
    feet_xy8[4 + i] = (gait_domain <= 1) ? fp[i][1] + agent_initial[1] : toe_pos12[3 * i + 1];
  }
}

int orc_closest_obstacle(const double *pos, const double *obs_xy, int K, double *out) {
  double min_dist = 1000; /* src/MPC_dist.cpp:371 */
  int min_i = 0;
  for (int i = 0; i < K; i++) {
    double d = sqrt(pow(pos[0] - obs_xy[i], 2) + pow(pos[1] - obs_xy[K + i], 2));
    if (d < min_dist) { min_dist = d; min_i = i; }
  }
  out[0] = obs_xy[min_i]; out[1] = obs_xy[K + min_i];
  return min_i;
}

/* ------------------------------------------------------------------------------------------------ */
/* Bunch-Kaufman LDL' (lower, unblocked) : factor, solve, inertia                                   */
/* ------------------------------------------------------------------------------------------------ */
static int bk_factor(int n, double *A, int lda, int *ipiv) {
#define AA(i, j) A[(i) + (size_t)(j) * lda]
  const double alpha = (1.0 + sqrt(17.0)) / 8.0;
  int info = 0, k = 0;
  while (k < n) {
    int kstep = 1, kp = k, imax = k;
    double absakk = fabs(AA(k, k)), colmax = 0.0;
    for (int i = k + 1; i < n; i++)
      if (fabs(AA(i, k)) > colmax) { colmax = fabs(AA(i, k)); imax = i; }
    if (fmax(absakk, colmax) == 0.0) {
      if (!info) info = k + 1;
      kp = k;
    } else {
      if (absakk >= alpha * colmax) kp = k;
      else {
        double rowmax = 0.0;
        for (int j = k; j < imax; j++) rowmax = fmax(rowmax, fabs(AA(imax, j)));
        for (int i = imax + 1; i < n; i++) rowmax = fmax(rowmax, fabs(AA(i, imax)));
        if (absakk >= alpha * colmax * (colmax / rowmax)) kp = k;
        else if (fabs(AA(imax, imax)) >= alpha * rowmax) kp = imax;
        else { kp = imax; kstep = 2; }
      }
      int kk = k + kstep - 1;
      if (kp != kk) {
        for (int i = kp + 1; i < n; i++) { double t = AA(i, kk); AA(i, kk) = AA(i, kp); AA(i, kp) = t; }
        for (int j = kk + 1; j < kp; j++) { double t = AA(j, kk); AA(j, kk) = AA(kp, j); AA(kp, j) = t; }
        { double t = AA(kk, kk); AA(kk, kk) = AA(kp, kp); AA(kp, kp) = t; }
        if (kstep == 2) { double t = AA(k + 1, k); AA(k + 1, k) = AA(kp, k); AA(kp, k) = t; }
      }
      if (kstep == 1) {
        if (k < n - 1) {
          double r1 = 1.0 / AA(k, k);
          for (int j = k + 1; j < n; j++) {
            double t = r1 * AA(j, k);
            if (t != 0.0)
              for (int i = j; i < n; i++) AA(i, j) -= AA(i, k) * t;
          }
          for (int i = k + 1; i < n; i++) AA(i, k) *= r1;
        }
      } else if (k < n - 2) {
        double d21 = AA(k + 1, k), d11 = AA(k + 1, k + 1) / d21, d22 = AA(k, k) / d21;
        double t = 1.0 / (d11 * d22 - 1.0);
        d21 = t / d21;
        for (int j = k + 2; j < n; j++) {
          double wk = d21 * (d11 * AA(j, k) - AA(j, k + 1));
          double wkp1 = d21 * (d22 * AA(j, k + 1) - AA(j, k));
          for (int i = j; i < n; i++) AA(i, j) -= AA(i, k) * wk + AA(i, k + 1) * wkp1;
          AA(j, k) = wk; AA(j, k + 1) = wkp1;
        }
      }
    }
    if (kstep == 1) ipiv[k] = kp;
    else { ipiv[k] = -(kp + 1); ipiv[k + 1] = -(kp + 1); }
    k += kstep;
  }
  return info;
#undef AA
}

static void bk_inertia(int n, const double *A, int lda, const int *ipiv, int *npos, int *nneg, int *nzero) {
  *npos = *nneg = *nzero = 0;
  int k = 0;
  while (k < n) {
    if (ipiv[k] >= 0) {
      double d = A[k + (size_t)k * lda];
      if (d > 0) (*npos)++; else if (d < 0) (*nneg)++; else (*nzero)++;
      k++;
    } else {
      double a = A[k + (size_t)k * lda], b = A[k + 1 + (size_t)k * lda], d = A[k + 1 + (size_t)(k + 1) * lda];
      double det = a * d - b * b, tr = a + d;
      if (det < 0) { (*npos)++; (*nneg)++; }
      else if (det > 0) { if (tr > 0) *npos += 2; else *nneg += 2; }
      else { (*nzero)++; if (tr > 0) (*npos)++; else if (tr < 0) (*nneg)++; else (*nzero)++; }
      k += 2;
    }
  }
}

static void bk_solve(int n, const double *A, int lda, const int *ipiv, double *b) {
#define AA(i, j) A[(i) + (size_t)(j) * lda]
  int k = 0;
  while (k < n) {
    if (ipiv[k] >= 0) {
      int kp = ipiv[k];
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      for (int i = k + 1; i < n; i++) b[i] -= AA(i, k) * b[k];
      b[k] /= AA(k, k);
      k++;
    } else {
      int kp = -ipiv[k] - 1;
      if (kp != k + 1) { double t = b[k + 1]; b[k + 1] = b[kp]; b[kp] = t; }
      for (int i = k + 2; i < n; i++) b[i] -= AA(i, k) * b[k] + AA(i, k + 1) * b[k + 1];
      double akm1k = AA(k + 1, k), akm1 = AA(k, k) / akm1k, ak = AA(k + 1, k + 1) / akm1k, denom = akm1 * ak - 1.0;
      double bkm1 = b[k] / akm1k, bk = b[k + 1] / akm1k;
      b[k] = (ak * bkm1 - bk) / denom;
      b[k + 1] = (akm1 * bk - bkm1) / denom;
      k += 2;
    }
  }
  k = n - 1;
  while (k >= 0) {
    if (ipiv[k] >= 0) {
      double s = 0;
      for (int i = k + 1; i < n; i++) s += AA(i, k) * b[i];
      b[k] -= s;
      int kp = ipiv[k];
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k--;
    } else {
      double s0 = 0, s1 = 0;
      for (int i = k + 1; i < n; i++) { s0 += AA(i, k) * b[i]; s1 += AA(i, k - 1) * b[i]; }
      b[k] -= s0; b[k - 1] -= s1;
      int kp = -ipiv[k] - 1;
      if (kp != k) { double t = b[k]; b[k] = b[kp]; b[kp] = t; }
      k -= 2;
    }
  }
#undef AA
}

/* ------------------------------------------------------------------------------------------------ */
/* QP stage: iSWIFT restated (Prime.c:127-230, Auxilary.c)                                          */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  int n, m, p, N;
  const double *P, *A, *G;
  double *K;   /* (n+p)^2 */
  int *ipiv;
  double *w;   /* m : s/z of the current factorisation */
} qp_kkt;

/* factor [P + G' W^-1 G , A' ; A , 0] */
static int qp_factor(qp_kkt *k, const double *w) {
  int n = k->n, m = k->m, p = k->p, N = n + p;
  memset(k->K, 0, sizeof(double) * N * N);
  for (int j = 0; j < n; j++)
    for (int i = j; i < n; i++) k->K[i + (size_t)j * N] = k->P[i + (size_t)j * n];
  for (int r = 0; r < m; r++) {
    double wi = 1.0 / w[r];
    for (int j = 0; j < n; j++) {
      double gj = k->G[r + (size_t)j * m];
      if (gj == 0.0) continue;
      for (int i = j; i < n; i++) {
        double gi = k->G[r + (size_t)i * m];
        if (gi != 0.0) k->K[i + (size_t)j * N] += gi * wi * gj;
      }
    }
  }
  for (int j = 0; j < n; j++)
    for (int i = 0; i < p; i++) k->K[(n + i) + (size_t)j * N] = k->A[i + (size_t)j * p];
  memcpy(k->w, w, sizeof(double) * m);
  return bk_factor(N, k->K, N, k->ipiv);
}

/* solve [P A' G'; A 0 0; G 0 -W][dx;dy;dz] = [r1;r2;r3] with the stored factor */
static void qp_solve(const qp_kkt *k, const double *r1, const double *r2, const double *r3, double *dx, double *dy,
                     double *dz, double *work) {
  int n = k->n, m = k->m, p = k->p, N = n + p;
  for (int j = 0; j < n; j++) {
    double s = r1[j];
    for (int r = 0; r < m; r++) {
      double gj = k->G[r + (size_t)j * m];
      if (gj != 0.0) s += gj * r3[r] / k->w[r];
    }
    work[j] = s;
  }
  for (int i = 0; i < p; i++) work[n + i] = r2[i];
  bk_solve(N, k->K, N, k->ipiv, work);
  memcpy(dx, work, sizeof(double) * n);
  memcpy(dy, work + n, sizeof(double) * p);
  for (int r = 0; r < m; r++) {
    double s = -r3[r];
    for (int j = 0; j < n; j++) {
      double gj = k->G[r + (size_t)j * m];
      if (gj != 0.0) s += gj * dx[j];
    }
    dz[r] = s / k->w[r];
  }
}

static double vnorm2(const double *v, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += v[i] * v[i];
  return sqrt(s);
}

static void steplength(const double *s, const double *ds, const double *z, const double *dz, int m, double *ap,
                       double *ad) { /* Auxilary.c:271-294 */
  double a = 1e10, d = 1e10;
  int f1 = 0, f2 = 0;
  for (int i = 0; i < m; i++) {
    if (ds[i] < 0 && (-s[i] / ds[i]) < a) { a = -(s[i] / ds[i]); f1 = 1; }
    if (dz[i] < 0 && (-z[i] / dz[i]) < d) { d = -(z[i] / dz[i]); f2 = 1; }
  }
  *ap = f1 ? a : 1.0;
  *ad = f2 ? d : 1.0;
}

int orc_qp_mehrotra(int n, int m, int p, const double *P, const double *c, const double *A, const double *b,
                    const double *G, const double *h, double reltol, double abstol, int maxit, double sigma_d,
                    double *x, double *y, double *z, double *s, int *iters) {
  int N = n + p, flag = 3;
  qp_kkt k = {n, m, p, N, P, A, G, NULL, NULL, NULL};
  k.K = (double *)malloc(sizeof(double) * N * N);
  k.ipiv = (int *)malloc(sizeof(int) * N);
  k.w = (double *)malloc(sizeof(double) * m);
  double *buf = (double *)calloc((size_t)(6 * n + 4 * p + 12 * m + N), sizeof(double));
  double *rx = buf, *dx = rx + n, *r1 = dx + n, *ry = r1 + n, *dy = ry + p, *rz = dy + p, *dz = rz + m, *ds = dz + m,
         *dsl = ds + m, *r3 = dsl + m, *w = r3 + m, *lam2 = w + m, *work = lam2 + m;
  /* kkt_initialize (Auxilary.c:680-755): W = I, rhs [-c; b; h] */
  for (int i = 0; i < m; i++) w[i] = 1.0;
  for (int j = 0; j < n; j++) r1[j] = -c[j];
  *iters = 0;
  if (qp_factor(&k, w)) { flag = 1; goto done; }
  qp_solve(&k, r1, b, h, x, y, dz, work);
  {
    double rmin = 1e300, rmax = -1e300;
    for (int r = 0; r < m; r++) {
      double t = h[r];
      for (int j = 0; j < n; j++) {
        double gj = G[r + (size_t)j * m];
        if (gj != 0.0) t -= gj * x[j];
      }
      rz[r] = t;
      if (t < rmin) rmin = t;
      if (t > rmax) rmax = t;
    }
    double ap = -rmin, ad = rmax;
    for (int r = 0; r < m; r++) {
      s[r] = (ap < 0) ? rz[r] : rz[r] + (1 + ap);
      z[r] = (ad < 0) ? -rz[r] : -rz[r] + (1 + ad);
    }
  }
  double sigma = 100.0; /* GlobalOptions.h:26 */
  flag = 2;
  int it;
  for (it = 0; it < maxit; it++) {
    /* computeresiduals (Auxilary.c:524-553) */
    for (int j = 0; j < n; j++) {
      double t = c[j];
      for (int i = 0; i < n; i++) t += P[j + (size_t)i * n] * x[i];
      for (int r = 0; r < m; r++) {
        double gj = G[r + (size_t)j * m];
        if (gj != 0.0) t += gj * z[r];
      }
      for (int i = 0; i < p; i++) t += A[i + (size_t)j * p] * y[i];
      rx[j] = -t;
    }
    for (int i = 0; i < p; i++) {
      double t = b[i];
      for (int j = 0; j < n; j++) t -= A[i + (size_t)j * p] * x[j];
      ry[i] = t;
    }
    double sz = 0;
    for (int r = 0; r < m; r++) {
      double t = h[r] - s[r];
      for (int j = 0; j < n; j++) {
        double gj = G[r + (size_t)j * m];
        if (gj != 0.0) t -= gj * x[j];
      }
      rz[r] = t;
      sz += s[r] * z[r];
    }
    double lim = reltol / sqrt(3.0);
    if (vnorm2(rx, n) < lim && vnorm2(rz, m) < lim && (p == 0 || vnorm2(ry, p) < lim) && sz / m < abstol) {
      flag = 0;
      break;
    }
    double mu = 0;
    for (int r = 0; r < m; r++) { lam2[r] = sqrt(s[r] * z[r]); lam2[r] *= lam2[r]; mu += lam2[r]; }
    mu /= m;
    if (sigma > sigma_d) {
      for (int r = 0; r < m; r++) { w[r] = s[r] / z[r]; ds[r] = -lam2[r]; r3[r] = rz[r] - ds[r] / z[r]; }
      if (qp_factor(&k, w)) { flag = 1; break; }
      qp_solve(&k, rx, ry, r3, dx, dy, dz, work);
      for (int r = 0; r < m; r++) dsl[r] = (ds[r] - s[r] * dz[r]) / z[r];
      double ap, ad;
      steplength(s, dsl, z, dz, m, &ap, &ad);
      double num = 0;
      for (int r = 0; r < m; r++) num += (s[r] + ap * dsl[r]) * (z[r] + ad * dz[r]);
      double rho = num / sz;
      double mr = rho < 1.0 ? rho : 1.0;
      sigma = mr * mr * mr;
      if (sigma < sigma_d) sigma = sigma_d;
      for (int r = 0; r < m; r++) ds[r] = -lam2[r] - dsl[r] * dz[r] + sigma * mu;
    } else {
      sigma = sigma_d;
      for (int r = 0; r < m; r++) ds[r] = -lam2[r] + sigma * mu;
    }
    for (int r = 0; r < m; r++) r3[r] = rz[r] - ds[r] / z[r];
    qp_solve(&k, rx, ry, r3, dx, dy, dz, work);
    for (int r = 0; r < m; r++) dsl[r] = (ds[r] - s[r] * dz[r]) / z[r];
    double ap, ad;
    steplength(s, dsl, z, dz, m, &ap, &ad);
    ap = fmin(0.99 * ap, 1.0); ad = fmin(0.99 * ad, 1.0);
    for (int j = 0; j < n; j++) x[j] += ap * dx[j];
    for (int i = 0; i < p; i++) y[i] += ad * dy[i];
    for (int r = 0; r < m; r++) { s[r] += ap * dsl[r]; z[r] += ad * dz[r]; }
  }
  *iters = it;
done:
  free(buf); free(k.K); free(k.ipiv); free(k.w);
  return flag;
}

/* ------------------------------------------------------------------------------------------------ */
/* NLP stage: IPM-F                                                                                  */
/* ------------------------------------------------------------------------------------------------ */
typedef struct {
  int nv, mi, m, p, c, with_nl;   /* with_nl: the 8 + 8 velocity rows are present */
  const double *Q, *f, *A, *b, *G, *h;
  double v_max;
  int nd;                          /* distance rows: |p_{dstage+1} - dctr|^2 + delta >= deps */
  int dstage[ORC_ND_MAX];
  double dctr[ORC_ND_MAX][2], deps[ORC_ND_MAX];
} nlp_prob;

static void nlp_cons(const nlp_prob *P, const double *x, double *cv) { /* c(x) >= 0, row order [h-Gx | d-eps | vmax-v | vmax+v] */
  for (int r = 0; r < P->mi; r++) {
    double t = P->h[r];
    for (int j = 0; j < P->nv; j++) {
      double gj = P->G[r + (size_t)j * P->mi];
      if (gj != 0.0) t -= gj * x[j];
    }
    cv[r] = t;
  }
  for (int r = 0; r < P->nd; r++) { /* dec_vars_constr_cost.h:262-265 */
    const int k = P->dstage[r];
    cv[P->mi + r] = pow(x[4 * k] - P->dctr[r][0], 2) + pow(x[4 * k + 2] - P->dctr[r][1], 2) * 1 + x[P->nv - 1] - P->deps[r];
  }
  if (P->with_nl) {
    for (int r = 0; r < 8; r++) {
      cv[P->mi + P->nd + r] = P->v_max - x[VEL_IDX[r]];
      cv[P->mi + P->nd + 8 + r] = P->v_max + x[VEL_IDX[r]];
    }
  }
}

/* J (m x nv, column-major, ld = m) */
static void nlp_jac(const nlp_prob *P, const double *x, double *J) {
  int m = P->m, nv = P->nv;
  memset(J, 0, sizeof(double) * m * nv);
  for (int j = 0; j < nv; j++)
    for (int r = 0; r < P->mi; r++) J[r + (size_t)j * m] = -P->G[r + (size_t)j * P->mi];
  for (int r = 0; r < P->nd; r++) { /* dec_vars_constr_cost.h:363-377 */
    const int k = P->dstage[r];
    J[(P->mi + r) + (size_t)(4 * k) * m] = 2 * (x[4 * k] - P->dctr[r][0]);
    J[(P->mi + r) + (size_t)(4 * k + 2) * m] = 2 * (x[4 * k + 2] - P->dctr[r][1]);
    J[(P->mi + r) + (size_t)(nv - 1) * m] = 1.0;
  }
  if (P->with_nl) {
    for (int r = 0; r < 8; r++) {
      J[(P->mi + P->nd + r) + (size_t)VEL_IDX[r] * m] = -1.0;
      J[(P->mi + P->nd + 8 + r) + (size_t)VEL_IDX[r] * m] = 1.0;
    }
  }
}

static double nlp_obj(const nlp_prob *P, const double *x) { /* dec_vars_constr_cost.h:423-428 */
  double v = 0;
  for (int j = 0; j < P->nv; j++) {
    double t = 0;
    for (int i = 0; i < P->nv; i++) t += P->Q[i + (size_t)j * P->nv] * x[i];
    v += x[j] * (0.5 * t + P->f[j]);
  }
  return v;
}

static void nlp_grad(const nlp_prob *P, const double *x, double *g) { /* :430-438 */
  for (int j = 0; j < P->nv; j++) {
    double t = P->f[j];
    for (int i = 0; i < P->nv; i++) t += P->Q[j + (size_t)i * P->nv] * x[i];
    g[j] = t;
  }
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* IPM-F step 0: strictly interior start w.r.t. every inequality row */
static void interior_start(const nlp_prob *P, const orc_gains *g, double push, double *x) {
  const double fb = g->mu_fric * g->h_com / sqrt(2.0);
  int c = P->c;
  for (int j = 0; j < 16; j++) {
    double lim = g->box;
    if (P->with_nl && (j & 1)) lim = fmin(lim, P->v_max);
    x[j] = clampd(x[j], -lim + push, lim - push);
  }
  for (int k = 0; k < 4; k++)
    for (int a = 0; a < 2; a++) {
      double lo = -g->box + push, hi = g->box - push;
      if (k >= 1) {
        double pk = x[4 * (k - 1) + 2 * a];
        lo = fmax(lo, pk - fb + push); hi = fmin(hi, pk + fb - push);
      }
      x[16 + 2 * k + a] = clampd(x[16 + 2 * k + a], lo, hi);
    }
  for (int j = 24; j < 24 + 4 * c; j++) x[j] = clampd(x[j], push, 1 - push);
  if (P->nd > 0) {
    double cmin = 1e300;
    for (int r = 0; r < P->nd; r++) {
      const int k = P->dstage[r];
      double d = pow(x[4 * k] - P->dctr[r][0], 2) + pow(x[4 * k + 2] - P->dctr[r][1], 2) + x[P->nv - 1] - P->deps[r];
      if (d < cmin) cmin = d;
    }
    if (cmin < push) x[P->nv - 1] += push - cmin;
  }
}

int orc_nlp_ipm(int nv, int mi, const double *Q, const double *f, const double *Aeq, const double *beq,
                const double *G, const double *h, const double *obs, const double *x_start, int with_nl,
                const orc_gains *g, const orc_options *o, orc_result *res) {
  int dstage[4] = {0, 1, 2, 3};
  double dctr[8], deps[4];
  for (int k = 0; k < 4; k++) { dctr[2 * k] = obs ? obs[0] : 0; dctr[2 * k + 1] = obs ? obs[1] : 0; deps[k] = g->eps_dist; }
  return orc_nlp_ipm_rows(nv, mi, Q, f, Aeq, beq, G, h, with_nl ? 4 : 0, dstage, dctr, deps, x_start, with_nl, g, o, res);
}

int orc_nlp_ipm_rows(int nv, int mi, const double *Q, const double *f, const double *Aeq, const double *beq,
                     const double *G, const double *h, int nd, const int *dstage, const double *dctr, const double *deps,
                     const double *x_start, int with_nl, const orc_gains *g, const orc_options *o, orc_result *res) {
  nlp_prob P;
  if (nd < 0 || nd > ORC_ND_MAX) { memset(res, 0, sizeof(*res)); res->status = 4; return 4; }
  P.nv = nv; P.mi = mi; P.with_nl = with_nl; P.nd = nd; P.m = mi + nd + (with_nl ? 16 : 0); P.p = ORC_NEQ; P.c = (nv - 25) / 4;
  P.Q = Q; P.f = f; P.A = Aeq; P.b = beq; P.G = G; P.h = h;
  for (int r = 0; r < nd; r++) { P.dstage[r] = dstage[r]; P.dctr[r][0] = dctr[2 * r]; P.dctr[r][1] = dctr[2 * r + 1]; P.deps[r] = deps[r]; }
  P.v_max = g->v_max;
  const int n = nv, m = P.m, p = P.p;
  /* general (multi-variable) rows stay in the augmented system; single-variable bound rows are condensed */
  int gen[ORC_MROWS_MAX], ng = 0, isgen[ORC_MROWS_MAX];
  for (int r = 0; r < m; r++) {
    int nnz = 0;
    if (r < mi) {
      for (int j = 0; j < n; j++) nnz += (G[r + (size_t)j * mi] != 0.0);
    } else nnz = (r < mi + nd) ? 3 : 1;
    isgen[r] = nnz > 1;
    if (isgen[r]) gen[ng++] = r;
  }
  const int NK = n + p + ng;
  double *buf = (double *)calloc((size_t)NK * NK + (size_t)m * n + 8 * (size_t)NK + 12 * (size_t)m + 8 * (size_t)n, sizeof(double));
  double *K = buf, *J = K + (size_t)NK * NK, *rhs = J + (size_t)m * n, *sol = rhs + NK;
  double *x = sol + NK, *xt = x + n, *gr = xt + n, *dx = gr + n, *rd = dx + n;
  double *s = rd + n, *z = s + m, *st = z + m, *ds = st + m, *dz = ds + m, *sig = dz + m;
  double *y = sig + m, *dy = y + p, *re = dy + p;
  int *ipiv = (int *)malloc(sizeof(int) * NK);
  double filt_t[16], filt_p[16];
  int nfilt = 0;
  memcpy(x, x_start, sizeof(double) * n);
  interior_start(&P, g, o->push, x);
  nlp_cons(&P, x, s);
  double mu = o->mu_init, tau = fmax(o->tau_min, 1 - mu), dw_last = 0.0;
  int first_ic = 1, status = 1, it = 0, n_fact = 0, n_ls = 0;
  for (int r = 0; r < m; r++) {
    if (!(s[r] > 0)) status = 4;
    z[r] = mu / s[r];
  }
  double theta0 = 0;
  for (int i = 0; i < p; i++) {
    double t = -beq[i];
    for (int j = 0; j < n; j++) t += Aeq[i + (size_t)j * p] * x[j];
    theta0 += fabs(t);
  }
  const double theta_max = 1e4 * fmax(1.0, theta0), theta_min = 1e-4 * fmax(1.0, theta0);
  double E0 = 1e300;
  if (status == 4) goto finish;
  for (it = 0; it <= o->max_iter; it++) {
    nlp_cons(&P, x, s);
    nlp_jac(&P, x, J);
    nlp_grad(&P, x, gr);
    double e_d = 0, e_p = 0, e_c = 0, th = 0;
    for (int j = 0; j < n; j++) {
      double t = gr[j];
      for (int i = 0; i < p; i++) t += Aeq[i + (size_t)j * p] * y[i];
      for (int r = 0; r < m; r++) {
        double jr = J[r + (size_t)j * m];
        if (jr != 0.0) t -= jr * z[r];
      }
      rd[j] = t;
      e_d = fmax(e_d, fabs(t));
    }
    for (int i = 0; i < p; i++) {
      double t = -beq[i];
      for (int j = 0; j < n; j++) t += Aeq[i + (size_t)j * p] * x[j];
      re[i] = t;
      e_p = fmax(e_p, fabs(t));
      th += fabs(t);
    }
    for (int r = 0; r < m; r++) e_c = fmax(e_c, fabs(s[r] * z[r]));
    E0 = fmax(e_d, fmax(e_p, e_c));
    if (E0 <= o->tol || (mu <= o->tol / 10 && E0 <= o->acc_tol)) { status = 0; break; }
    if (it == o->max_iter) break;
    double zmax = 0;
    for (int r = 0; r < m; r++) zmax = fmax(zmax, z[r]);
    if (!(zmax < o->z_div) || !isfinite(E0)) { status = 2; break; } /* diverging multipliers: infeasible */
    /* barrier update (Ipopt eq. 7) */
    for (;;) {
      double emu = 0;
      for (int r = 0; r < m; r++) emu = fmax(emu, fabs(s[r] * z[r] - mu));
      emu = fmax(emu, fmax(e_d, e_p));
      if (emu <= o->kappa_eps * mu && mu > o->tol / 10) {
        mu = fmax(o->tol / 10, fmin(o->kappa_mu * mu, pow(mu, o->theta_mu)));
        tau = fmax(o->tau_min, 1 - mu);
        nfilt = 0;
      } else break;
    }
    /* Newton system  [W + Jb' Sig Jb + dw I + rho_lambda I_lam , A' , Jg' ; A 0 0 ; Jg 0 -S/Z] */
    for (int r = 0; r < m; r++) sig[r] = z[r] / s[r];
    double dw = 0.0;
    int ok = 0;
    for (;;) {
      memset(K, 0, sizeof(double) * NK * NK);
      for (int j = 0; j < n; j++) {
        for (int i = j; i < n; i++) K[i + (size_t)j * NK] = Q[i + (size_t)j * n];
        K[j + (size_t)j * NK] += dw;
      }
      for (int j = 24; j < n - 1; j++) K[j + (size_t)j * NK] += o->rho_lambda;
      for (int r = 0; r < nd; r++) { /* Hessian of the Lagrangian: -2 z_r on the two position diagonals of the row's stage */
        const int k = P.dstage[r];
        K[(4 * k) + (size_t)(4 * k) * NK] -= 2 * z[mi + r];
        K[(4 * k + 2) + (size_t)(4 * k + 2) * NK] -= 2 * z[mi + r];
      }
      for (int r = 0; r < m; r++) {
        if (isgen[r]) continue;
        for (int j = 0; j < n; j++) {
          double jj = J[r + (size_t)j * m];
          if (jj != 0.0) K[j + (size_t)j * NK] += sig[r] * jj * jj; /* single-variable rows */
        }
      }
      for (int j = 0; j < n; j++) {
        for (int i = 0; i < p; i++) K[(n + i) + (size_t)j * NK] = Aeq[i + (size_t)j * p];
        for (int q = 0; q < ng; q++) K[(n + p + q) + (size_t)j * NK] = J[gen[q] + (size_t)j * m];
      }
      for (int q = 0; q < ng; q++) K[(n + p + q) + (size_t)(n + p + q) * NK] = -s[gen[q]] / z[gen[q]];
      n_fact++;
      bk_factor(NK, K, NK, ipiv);
      int np_, nn_, nz_;
      bk_inertia(NK, K, NK, ipiv, &np_, &nn_, &nz_);
      if (np_ == n && nn_ == p + ng && nz_ == 0) { ok = 1; break; }
      if (dw == 0.0) dw = (dw_last == 0.0) ? o->dw_first : fmax(o->dw_min, o->kw_minus * dw_last);
      else dw *= (first_ic && dw_last == 0.0) ? o->kw_plus_first : o->kw_plus;
      if (dw > o->dw_max) break;
    }
    if (!ok) { status = 3; break; }
    if (dw > 0) { dw_last = dw; first_ic = 0; }
    for (int j = 0; j < n; j++) {
      double t = -rd[j];
      for (int r = 0; r < m; r++) {
        if (isgen[r]) continue;
        double jr = J[r + (size_t)j * m];
        if (jr != 0.0) t -= jr * (z[r] - mu / s[r]);
      }
      rhs[j] = t;
    }
    for (int i = 0; i < p; i++) rhs[n + i] = -re[i];
    for (int q = 0; q < ng; q++) rhs[n + p + q] = mu / z[gen[q]] - s[gen[q]];
    memcpy(sol, rhs, sizeof(double) * NK);
    bk_solve(NK, K, NK, ipiv, sol);
    memcpy(dx, sol, sizeof(double) * n);
    memcpy(dy, sol + n, sizeof(double) * p);
    for (int r = 0; r < m; r++) {
      double t = 0;
      for (int j = 0; j < n; j++) {
        double jr = J[r + (size_t)j * m];
        if (jr != 0.0) t += jr * dx[j];
      }
      ds[r] = t;
      dz[r] = mu / s[r] - z[r] - sig[r] * t;
    }
    for (int q = 0; q < ng; q++) dz[gen[q]] = -sol[n + p + q]; /* third block solves for -dz */
    double a_pmax = 1.0, a_dmax = 1.0;
    for (int r = 0; r < m; r++) {
      if (ds[r] < 0) a_pmax = fmin(a_pmax, -tau * s[r] / ds[r]);
      if (dz[r] < 0) a_dmax = fmin(a_dmax, -tau * z[r] / dz[r]);
    }
    double ph = nlp_obj(&P, x), dphi = 0;
    for (int r = 0; r < m; r++) { ph -= mu * log(s[r]); dphi -= mu * ds[r] / s[r]; }
    for (int j = 0; j < n; j++) dphi += gr[j] * dx[j];
    double alpha = a_pmax;
    int accepted = 0, ftype = 0;
    while (alpha >= o->alpha_min) {
      n_ls++;
      for (int j = 0; j < n; j++) xt[j] = x[j] + alpha * dx[j];
      nlp_cons(&P, xt, st);
      double smin = 1e300;
      for (int r = 0; r < m; r++) smin = fmin(smin, st[r]);
      if (smin > 0) {
        double tht = (1 - alpha) * th, pht = nlp_obj(&P, xt);
        for (int r = 0; r < m; r++) pht -= mu * log(st[r]);
        int okf = tht <= theta_max;
        for (int q = 0; q < nfilt && okf; q++)
          if (tht >= filt_t[q] && pht >= filt_p[q]) okf = 0;
        if (okf) {
          int sw = dphi < 0 && alpha * pow(-dphi, o->s_phi) > o->delta * pow(th, o->s_theta);
          if (th <= theta_min && sw) {
            if (pht - ph - 10 * EPSM * fabs(ph) <= o->eta_phi * alpha * dphi) { accepted = 1; ftype = 1; }
          } else if (tht <= (1 - o->gamma_theta) * th || pht - ph - 10 * EPSM * fabs(ph) <= -o->gamma_phi * th) {
            accepted = 1; ftype = 0;
          }
        }
      }
      if (accepted) break;
      alpha *= 0.5;
    }
    if (!accepted) { status = 2; break; }
    if (!ftype) {
      if (nfilt >= o->filter_cap) {
        memmove(filt_t, filt_t + 1, sizeof(double) * (nfilt - 1));
        memmove(filt_p, filt_p + 1, sizeof(double) * (nfilt - 1));
        nfilt--;
      }
      filt_t[nfilt] = (1 - o->gamma_theta) * th;
      filt_p[nfilt] = ph - o->gamma_phi * th;
      nfilt++;
    }
    for (int j = 0; j < n; j++) x[j] += alpha * dx[j];
    for (int i = 0; i < p; i++) y[i] += alpha * dy[i];
    nlp_cons(&P, x, s);
    for (int r = 0; r < m; r++) {
      double zz = z[r] + a_dmax * dz[r];
      double lo = mu / (o->kappa_sigma * s[r]), hi = o->kappa_sigma * mu / s[r];
      z[r] = fmin(fmax(zz, lo), hi);
    }
  }
finish:
  memset(res->x, 0, sizeof(res->x));
  memcpy(res->x, x, sizeof(double) * n);
  memcpy(res->y, y, sizeof(double) * p);
  memset(res->z, 0, sizeof(res->z)); memset(res->s, 0, sizeof(res->s));
  memcpy(res->z, z, sizeof(double) * m); memcpy(res->s, s, sizeof(double) * m);
  res->kkt = E0; res->mu = mu; res->status = status; res->nlp_iters = it; res->n_fact = n_fact; res->n_ls = n_ls;
  res->nv = n; res->c = P.c; res->m = m;
  free(buf); free(ipiv);
  return status;
}

/* ------------------------------------------------------------------------------------------------ */
static void bernstein_inv(double *Minv) { /* inverse of M_ij = C(4,j) t_i^j (1-t_i)^(4-j), t_i = i/4 (math_define.h:409-414) */
  static const double binom[5] = {1, 4, 6, 4, 1};
  double M[25], Aug[5][10];
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) M[5 * i + j] = binom[j] * pow(i * 0.25, j) * pow(1 - i * 0.25, 4 - j);
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 10; j++) Aug[i][j] = j < 5 ? M[5 * i + j] : (j - 5 == i ? 1.0 : 0.0);
  for (int k = 0; k < 5; k++) {
    int piv = k;
    for (int i = k + 1; i < 5; i++)
      if (fabs(Aug[i][k]) > fabs(Aug[piv][k])) piv = i;
    for (int j = 0; j < 10; j++) { double t = Aug[k][j]; Aug[k][j] = Aug[piv][j]; Aug[piv][j] = t; }
    double d = Aug[k][k];
    for (int j = 0; j < 10; j++) Aug[k][j] /= d;
    for (int i = 0; i < 5; i++)
      if (i != k) {
        double fct = Aug[i][k];
        for (int j = 0; j < 10; j++) Aug[i][j] -= fct * Aug[k][j];
      }
  }
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) Minv[5 * i + j] = Aug[i][5 + j];
}

void orc_bezier_fit(const double *alpha_buf, const double *X16, double *alpha20, double *alpha_buf_next) {
  double Minv[25];
  bernstein_inv(Minv);
  for (int r = 0; r < 4; r++) {
    double Y[5] = {alpha_buf[r], X16[r], X16[4 + r], X16[8 + r], X16[12 + r]};
    for (int j = 0; j < 5; j++) {
      double t = 0;
      for (int i = 0; i < 5; i++) t += Minv[5 * j + i] * Y[i];
      alpha20[5 * r + j] = t;
    }
    alpha_buf_next[r] = X16[12 + r]; /* src/MPC_dist.cpp:798 */
  }
}

/* QP stage of orc_solve_instance: the restatement above, or (bench.py's CPU baseline) the reference's own iSWIFT core from
 * oracle/_ref, handed in as a function pointer with the signature of ref_driver.c's iswift_ref_solve_dense */
static orc_qp_backend g_qp_backend = 0;
void orc_set_qp_backend(orc_qp_backend f) { g_qp_backend = f; }

int orc_solve_instance(const orc_gains *g, const orc_options *o, const double *x0, const double *xdes,
                       const double *feet_xy, const int *contact, const double *obs_xy, int K,
                       const double *alpha_buf, int mode, orc_result *res) {
  double F[8];
  memset(res, 0, sizeof(*res));
  int c = orc_stance_matrix(feet_xy, contact, F);
  if (c < 0) { res->status = 4; return 4; }
  {
    /* non-finite data are bad input too (the reference would hand NaN to its solvers) */
    double obs0[2] = {0, 0}, chk = 0.0;
    if (mode == 1) { double pos[2] = {x0[0], x0[2]}; orc_closest_obstacle(pos, obs_xy, K, obs0); }
    for (int i = 0; i < 4; i++) chk += 0.0 * x0[i];
    for (int i = 0; i < 16; i++) chk += 0.0 * xdes[i];
    for (int i = 0; i < 2 * c; i++) chk += 0.0 * F[i];
    chk += 0.0 * obs0[0] + 0.0 * obs0[1];
    if (!(chk == 0.0)) { res->status = 4; res->c = c; return 4; }
  }
  const int nv = 25 + 4 * c, p = ORC_NEQ, mi = 60 + 8 * c;
  double *buf = (double *)malloc(sizeof(double) * ((size_t)nv * nv + nv + (size_t)p * nv + p + (size_t)mi * nv + mi + 2 * nv + p + 2 * mi));
  double *Q = buf, *f = Q + (size_t)nv * nv, *A = f + nv, *b = A + (size_t)p * nv, *G = b + p, *h = G + (size_t)mi * nv;
  double *xq = h + mi, *yq = xq + nv, *zq = yq + p, *sq = zq + mi;
  orc_pack_dense(g, x0, xdes, F, c, Q, f, A, b, G, h);
  int qit = 0;
  int qflag;
  if (g_qp_backend) { double t1 = 0, t2 = 0; qflag = g_qp_backend(nv, mi, p, Q, f, A, b, G, h, xq, yq, zq, sq, &qit, &t1, &t2); }
  else qflag = orc_qp_mehrotra(nv, mi, p, Q, f, A, b, G, h, o->qp_reltol, o->qp_abstol, o->qp_maxit, 0.0, xq, yq, zq, sq, &qit);
  double obs[2] = {0, 0};
  if (mode == 1) {
    double pos[2] = {x0[0], x0[2]};
    orc_closest_obstacle(pos, obs_xy, K, obs);
    orc_nlp_ipm(nv, mi, Q, f, A, b, G, h, obs, xq, 1, g, o, res);
  } else {
    memcpy(res->x, xq, sizeof(double) * nv);
    memcpy(res->y, yq, sizeof(double) * p);
    memcpy(res->z, zq, sizeof(double) * mi);
    memcpy(res->s, sq, sizeof(double) * mi);
    res->status = qflag == 0 ? 0 : 1;
    res->nv = nv; res->c = c; res->m = mi;
  }
  memcpy(res->x_qp, xq, sizeof(double) * nv);
  res->qp_flag = qflag; res->qp_iters = qit; res->obs[0] = obs[0]; res->obs[1] = obs[1];
  orc_bezier_fit(alpha_buf, res->x, res->alpha, res->alpha_buf_next);
  free(buf);
  return res->status;
}

typedef struct {
  const orc_gains *g; const orc_options *o; int B, K, mode, tid, nthreads;
  const double *x0, *xdes, *feet, *obs, *abuf; const int *contact; orc_result *res;
} batch_arg;

static void *batch_worker(void *vp) {
  batch_arg *a = (batch_arg *)vp;
  for (int i = a->tid; i < a->B; i += a->nthreads)
    orc_solve_instance(a->g, a->o, a->x0 + 4 * (size_t)i, a->xdes + 16 * (size_t)i, a->feet + 8 * (size_t)i,
                       a->contact + 4 * (size_t)i, a->obs + 2 * (size_t)a->K * i, a->K, a->abuf + 4 * (size_t)i, a->mode,
                       a->res + i);
  return NULL;
}

int orc_solve_batch(const orc_gains *g, const orc_options *o, int B, const double *x0, const double *xdes,
                    const double *feet_xy, const int *contact, const double *obs_xy, int K,
                    const double *alpha_buf, int mode, int nthreads, orc_result *res) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  batch_arg args[256];
  for (int t = 0; t < nthreads; t++) {
    batch_arg a = {g, o, B, K, mode, t, nthreads, x0, xdes, feet_xy, obs_xy, alpha_buf, contact, res};
    args[t] = a;
    if (nthreads > 1) pthread_create(&th[t], NULL, batch_worker, &args[t]);
  }
  if (nthreads == 1) batch_worker(&args[0]);
  else
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return 0;
}

/* dense LDL' of this file for cbfnmpc_oracle_ext.c (the extended-mode checker) */
int orc_bk_factor(int n, double *A, int lda, int *ipiv) { return bk_factor(n, A, lda, ipiv); }
void orc_bk_inertia(int n, const double *A, int lda, const int *ipiv, int *npos, int *nneg, int *nzero) { bk_inertia(n, A, lda, ipiv, npos, nneg, nzero); }
void orc_bk_solve(int n, const double *A, int lda, const int *ipiv, double *b) { bk_solve(n, A, lda, ipiv, b); }
