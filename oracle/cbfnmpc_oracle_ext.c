/* TEST INFRASTRUCTURE -- extended-mode checker; see cbfnmpc_oracle_ext.h ("no reference implementation exists").
 * Plain C99, dense, structure-agnostic: the same interior-point method IPM-F as orc_nlp_ipm_rows (cbfnmpc_oracle.c), written as
 * begin / iterate so that the members of a team can exchange their iterates between iterations. */
#include "cbfnmpc_oracle_ext.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#define EPSM 2.220446049250313e-16
#define ORC_EXT_SFLOOR 1e-12

struct orc_ext {
  orc_gains g; orc_options o;
  int N, ND, c, nv, p, mi, K_obs, n_nbr, K, nd, m, ng, NK;
  double eps_obs, eps_nbr;
  double *Q, *f, *A, *b, *G, *h;           /* dense data of the linear part */
  double *obs;                             /* 2 x K_obs row-major */
  double *dctr, *deps; int *dstage;        /* nd distance rows: row r = i * N + k (row type i, grid point k) */
  int *isgen, *gen;
  double *K_, *J, *rhs, *sol, *x, *xt, *gr, *dx, *rd, *s, *z, *st, *ds, *dz, *sig, *y, *dy, *re, *sprev;
  int *ipiv;
  double filt_t[16], filt_p[16]; int nfilt;
  double mu, tau, dw_last, theta_max, theta_min, E0;
  int first_ic, status, it, n_fact, n_ls, n_bump, started, finished, mu_age, mu_stall;
  double x_qp[ORC_EXT_NV_MAX]; int qp_flag, qp_iters;
};

void orc_ext_pack_dense(const orc_gains *g, int N, const double *x0, const double *xdes, const double *F, int c, double *Q, double *f,
                        double *Aeq, double *beq, double *G, double *h) {
  const int n = 4, m = 2;
  const int nv = n * N + m * N + N * c + 1, p = n * N + m * N + N, mi = 2 * m * (N - 1) + 2 * n * N + 2 * m * N + 2 * N * c;
  double Ad[16], Bd[8];
  orc_lip_discrete(g, Ad, Bd);
  memset(Q, 0, sizeof(double) * nv * nv); memset(f, 0, sizeof(double) * nv);
  memset(Aeq, 0, sizeof(double) * p * nv); memset(beq, 0, sizeof(double) * p);
  memset(G, 0, sizeof(double) * mi * nv);
#define QQ(i, j) Q[(i) + (size_t)(j) * nv]
#define AE(i, j) Aeq[(i) + (size_t)(j) * p]
#define GG(i, j) G[(i) + (size_t)(j) * mi]
  for (int i = 0; i < n * N; i++) {
    double w = (i >= n * (N - 1)) ? g->Pw : g->Qw; /* src/MPC_dist.cpp:193-197: terminal weight on the last grid point */
    QQ(i, i) = w;
    f[i] = -xdes[i] * w;
  }
  for (int i = 0; i < m * N; i++) QQ(n * N + i, n * N + i) = g->Rw;
  QQ(nv - 1, nv - 1) = g->slack_w;
  for (int k = 0; k < N; k++) { /* :232-261; the stance matrix of grid point k is that of its gait domain k / 4 */
    const double *Fk = F + (size_t)(k / 4) * 2 * c;
    for (int i = 0; i < n; i++) {
      AE(n * k + i, n * k + i) = -1.0;
      if (k >= 1)
        for (int j = 0; j < n; j++) AE(n * k + i, n * (k - 1) + j) += Ad[4 * i + j];
      for (int j = 0; j < m; j++) AE(n * k + i, n * N + m * k + j) = Bd[2 * i + j];
    }
    for (int a = 0; a < m; a++) {
      AE(n * N + m * k + a, n * N + m * k + a) = 1.0;
      for (int j = 0; j < c; j++) AE(n * N + m * k + a, n * N + m * N + c * k + j) = -Fk[a * c + j];
    }
    for (int j = 0; j < c; j++) AE(n * N + m * N + k, n * N + m * N + c * k + j) = 1.0;
    beq[n * N + m * N + k] = 1.0;
  }
  for (int i = 0; i < n; i++) {
    double s = 0;
    for (int j = 0; j < n; j++) s += Ad[4 * i + j] * x0[j];
    beq[i] = -s;
  }
  const int r = m * (N - 1); /* :278-321 */
  for (int i = 0; i < N - 1; i++)
    for (int a = 0; a < 2; a++) {
      GG(i * m + a, i * n + 2 * a) = 1.0;
      GG(i * m + a, n * N + (i + 1) * m + a) = -1.0;
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

/* velocity variable of one-sided row q (0 .. 2N-1): v_x of grid points 0..N-1, then v_y (DV:270-279 order) */
/* ORC_EXT_TRACE=1 in the environment: one line per iteration and agent on stderr (debugging aid) */
static int ext_trace(void) { static int t = -1; if (t < 0) t = getenv("ORC_EXT_TRACE") ? 1 : 0; return t; }

static int vel_idx(int N, int q) { return q < N ? 4 * q + 1 : 4 * (q - N) + 3; }

static void ext_cons(const orc_ext *e, const double *x, double *cv) { /* c(x) >= 0: [h - Gx | d - eps | vmax - v | vmax + v] */
  const int nv = e->nv, mi = e->mi, N = e->N;
  for (int r = 0; r < mi; r++) {
    double t = e->h[r];
    for (int j = 0; j < nv; j++) {
      double gj = e->G[r + (size_t)j * mi];
      if (gj != 0.0) t -= gj * x[j];
    }
    cv[r] = t;
  }
  for (int r = 0; r < e->nd; r++) {
    const int k = e->dstage[r];
    cv[mi + r] = pow(x[4 * k] - e->dctr[2 * r], 2) + pow(x[4 * k + 2] - e->dctr[2 * r + 1], 2) * 1 + x[nv - 1] - e->deps[r];
  }
  for (int q = 0; q < 2 * N; q++) {
    cv[mi + e->nd + q] = e->g.v_max - x[vel_idx(N, q)];
    cv[mi + e->nd + 2 * N + q] = e->g.v_max + x[vel_idx(N, q)];
  }
}

static void ext_jac(const orc_ext *e, const double *x, double *J) {
  const int m = e->m, nv = e->nv, mi = e->mi, N = e->N;
  memset(J, 0, sizeof(double) * (size_t)m * nv);
  for (int j = 0; j < nv; j++)
    for (int r = 0; r < mi; r++) J[r + (size_t)j * m] = -e->G[r + (size_t)j * mi];
  for (int r = 0; r < e->nd; r++) {
    const int k = e->dstage[r];
    J[(mi + r) + (size_t)(4 * k) * m] = 2 * (x[4 * k] - e->dctr[2 * r]);
    J[(mi + r) + (size_t)(4 * k + 2) * m] = 2 * (x[4 * k + 2] - e->dctr[2 * r + 1]);
    J[(mi + r) + (size_t)(nv - 1) * m] = 1.0;
  }
  for (int q = 0; q < 2 * N; q++) {
    J[(mi + e->nd + q) + (size_t)vel_idx(N, q) * m] = -1.0;
    J[(mi + e->nd + 2 * N + q) + (size_t)vel_idx(N, q) * m] = 1.0;
  }
}

static double ext_obj(const orc_ext *e, const double *x) {
  double v = 0;
  for (int j = 0; j < e->nv; j++) {
    double t = 0;
    for (int i = 0; i < e->nv; i++) t += e->Q[i + (size_t)j * e->nv] * x[i];
    v += x[j] * (0.5 * t + e->f[j]);
  }
  return v;
}

static void ext_grad(const orc_ext *e, const double *x, double *g) {
  for (int j = 0; j < e->nv; j++) {
    double t = e->f[j];
    for (int i = 0; i < e->nv; i++) t += e->Q[j + (size_t)i * e->nv] * x[i];
    g[j] = t;
  }
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void ext_set_nbr(orc_ext *e, const double *nbr) { /* centres of the neighbour rows: row type K_obs + j, grid point k */
  for (int j = 0; j < e->n_nbr; j++)
    for (int k = 0; k < e->N; k++) {
      const int r = (e->K_obs + j) * e->N + k;
      e->dctr[2 * r] = nbr[(size_t)(j * e->N + k) * 2]; e->dctr[2 * r + 1] = nbr[(size_t)(j * e->N + k) * 2 + 1];
    }
}

static void ext_interior_start(orc_ext *e, double *x) { /* IPM-F step 0, as interior_start of cbfnmpc_oracle.c for N grid points */
  const orc_gains *g = &e->g;
  const double push = e->o.push, fb = g->mu_fric * g->h_com / sqrt(2.0);
  const int N = e->N, c = e->c;
  for (int j = 0; j < 4 * N; j++) {
    double lim = g->box;
    if (j & 1) lim = fmin(lim, g->v_max);
    x[j] = clampd(x[j], -lim + push, lim - push);
  }
  for (int k = 0; k < N; k++)
    for (int a = 0; a < 2; a++) {
      double lo = -g->box + push, hi = g->box - push;
      if (k >= 1) {
        double pk = x[4 * (k - 1) + 2 * a];
        lo = fmax(lo, pk - fb + push); hi = fmin(hi, pk + fb - push);
      }
      x[4 * N + 2 * k + a] = clampd(x[4 * N + 2 * k + a], lo, hi);
    }
  for (int j = 6 * N; j < 6 * N + N * c; j++) x[j] = clampd(x[j], push, 1 - push);
  if (e->nd > 0) {
    double cmin = 1e300;
    for (int r = 0; r < e->nd; r++) {
      const int k = e->dstage[r];
      double d = pow(x[4 * k] - e->dctr[2 * r], 2) + pow(x[4 * k + 2] - e->dctr[2 * r + 1], 2) + x[e->nv - 1] - e->deps[r];
      if (d < cmin) cmin = d;
    }
    if (cmin < push) x[e->nv - 1] += push - cmin;
  }
}

orc_ext *orc_ext_create(const orc_gains *g, const orc_options *o, int N, const double *x0, const double *xdes, const double *feet_xy,
                        const int *contact, int K_obs, const double *obs_xy, double eps_obs, int n_nbr, double eps_nbr) {
  if ((N != 4 && N != 8) || K_obs < 0 || n_nbr < 0 || K_obs + n_nbr > ORC_EXT_KMAX || K_obs + n_nbr < 1) return NULL;
  const int ND = N / 4;
  double F[2 * 8], Ftmp[8];
  const int c = orc_stance_matrix(feet_xy, contact, Ftmp);
  if (c < 0) return NULL;
  for (int d = 0; d < ND; d++) orc_stance_matrix(feet_xy + 8 * d, contact, F + (size_t)d * 2 * c); /* 2 x c each, one after the other */
  double chk = 0.0;
  for (int i = 0; i < 4; i++) chk += 0.0 * x0[i];
  for (int i = 0; i < 4 * N; i++) chk += 0.0 * xdes[i];
  for (int i = 0; i < 2 * c * ND; i++) chk += 0.0 * F[i];
  for (int i = 0; i < 2 * K_obs; i++) chk += 0.0 * obs_xy[i];
  if (!(chk == 0.0)) return NULL;
  orc_ext *e = (orc_ext *)calloc(1, sizeof(orc_ext));
  e->g = *g; e->o = *o; e->N = N; e->ND = ND; e->c = c; e->K_obs = K_obs; e->n_nbr = n_nbr; e->K = K_obs + n_nbr;
  e->eps_obs = eps_obs; e->eps_nbr = eps_nbr;
  e->nv = N * (6 + c) + 1; e->p = 7 * N; e->mi = 16 * N - 4 + 2 * c * N; e->nd = e->K * N; e->m = e->mi + e->nd + 4 * N;
  const int nv = e->nv, p = e->p, mi = e->mi, m = e->m;
  e->Q = (double *)calloc((size_t)nv * nv + nv + (size_t)p * nv + p + (size_t)mi * nv + mi, sizeof(double));
  e->f = e->Q + (size_t)nv * nv; e->A = e->f + nv; e->b = e->A + (size_t)p * nv; e->G = e->b + p; e->h = e->G + (size_t)mi * nv;
  orc_ext_pack_dense(g, N, x0, xdes, F, c, e->Q, e->f, e->A, e->b, e->G, e->h);
  e->obs = (double *)calloc(2 * (size_t)(K_obs > 0 ? K_obs : 1), sizeof(double));
  memcpy(e->obs, obs_xy, sizeof(double) * 2 * K_obs);
  e->dctr = (double *)calloc(2 * (size_t)e->nd, sizeof(double)); e->deps = (double *)calloc(e->nd, sizeof(double));
  e->dstage = (int *)calloc(e->nd, sizeof(int));
  for (int i = 0; i < e->K; i++)
    for (int k = 0; k < N; k++) {
      const int r = i * N + k;
      e->dstage[r] = k; e->deps[r] = i < K_obs ? eps_obs : eps_nbr;
      if (i < K_obs) { e->dctr[2 * r] = obs_xy[i]; e->dctr[2 * r + 1] = obs_xy[K_obs + i]; }
    }
  e->isgen = (int *)calloc(m, sizeof(int)); e->gen = (int *)calloc(m, sizeof(int));
  e->ng = 0;
  for (int r = 0; r < m; r++) { /* multi-variable rows stay in the augmented system (as orc_nlp_ipm_rows) */
    int nnz = 0;
    if (r < mi) { for (int j = 0; j < nv; j++) nnz += (e->G[r + (size_t)j * mi] != 0.0); }
    else nnz = (r < mi + e->nd) ? 3 : 1;
    e->isgen[r] = nnz > 1;
    if (e->isgen[r]) e->gen[e->ng++] = r;
  }
  e->NK = nv + p + e->ng;
  const int NK = e->NK;
  e->K_ = (double *)calloc((size_t)NK * NK + (size_t)m * nv + 8 * (size_t)NK + 13 * (size_t)m + 8 * (size_t)nv + 4 * (size_t)p, sizeof(double));
  e->J = e->K_ + (size_t)NK * NK; e->rhs = e->J + (size_t)m * nv; e->sol = e->rhs + NK;
  e->x = e->sol + NK; e->xt = e->x + nv; e->gr = e->xt + nv; e->dx = e->gr + nv; e->rd = e->dx + nv;
  e->s = e->rd + nv; e->z = e->s + m; e->st = e->z + m; e->ds = e->st + m; e->dz = e->ds + m; e->sig = e->dz + m; e->sprev = e->sig + m;
  e->y = e->sprev + m; e->dy = e->y + p; e->re = e->dy + p;
  e->ipiv = (int *)malloc(sizeof(int) * NK);
  e->status = 1; e->mu_stall = 6;
  return e;
}

void orc_ext_free(orc_ext *e) {
  if (!e) return;
  free(e->Q); free(e->obs); free(e->dctr); free(e->deps); free(e->dstage); free(e->isgen); free(e->gen); free(e->K_); free(e->ipiv);
  free(e);
}

int orc_ext_qp(orc_ext *e) {
  const int nv = e->nv, p = e->p, mi = e->mi;
  double *yq = (double *)calloc((size_t)p + 2 * (size_t)mi, sizeof(double)), *zq = yq + p, *sq = zq + mi;
  int it = 0;
  e->qp_flag = orc_qp_mehrotra(nv, mi, p, e->Q, e->f, e->A, e->b, e->G, e->h, e->o.qp_reltol, e->o.qp_abstol, e->o.qp_maxit, 0.0, e->x_qp, yq,
                               zq, sq, &it);
  e->qp_iters = it;
  memcpy(e->x, e->x_qp, sizeof(double) * nv);
  free(yq);
  return e->qp_flag;
}

void orc_ext_set_start(orc_ext *e, const double *x) { memcpy(e->x, x, sizeof(double) * e->nv); memcpy(e->x_qp, x, sizeof(double) * e->nv); }

void orc_ext_begin(orc_ext *e, const double *nbr) {
  const int nv = e->nv, p = e->p, m = e->m;
  const orc_options *o = &e->o;
  if (e->n_nbr > 0) ext_set_nbr(e, nbr);
  ext_interior_start(e, e->x);
  ext_cons(e, e->x, e->s);
  e->mu = o->mu_init; e->tau = fmax(o->tau_min, 1 - e->mu); e->dw_last = 0.0;
  e->mu_age = 0; e->first_ic = 1; e->status = 1; e->it = 0; e->n_fact = 0; e->n_ls = 0; e->n_bump = 0; e->nfilt = 0; e->finished = 0; e->started = 1;
  for (int r = 0; r < m; r++) {
    if (!(e->s[r] > 0)) e->status = 4;
    e->z[r] = e->mu / e->s[r];
    e->sprev[r] = e->s[r];
  }
  memset(e->y, 0, sizeof(double) * p);
  double theta0 = 0;
  for (int i = 0; i < p; i++) {
    double t = -e->b[i];
    for (int j = 0; j < nv; j++) t += e->A[i + (size_t)j * p] * e->x[j];
    theta0 += fabs(t);
  }
  e->theta_max = 1e4 * fmax(1.0, theta0); e->theta_min = 1e-4 * fmax(1.0, theta0);
  e->E0 = 1e300;
  if (e->status == 4) e->finished = 1;
}

int orc_ext_iter(orc_ext *e, const double *nbr) {
  if (e->finished) return 2;
  const orc_options *o = &e->o;
  const int n = e->nv, p = e->p, m = e->m, mi = e->mi, nd = e->nd, ng = e->ng, NK = e->NK;
  double *x = e->x, *s = e->s, *z = e->z, *J = e->J, *gr = e->gr, *rd = e->rd, *re = e->re, *y = e->y, *K = e->K_;
  const double *Aeq = e->A, *beq = e->b, *Q = e->Q;
  if (e->n_nbr > 0) {
    /* IPM-F-J: the neighbours have moved since the last iteration.  The method keeps s = c(x) > 0, so when a neighbour row would
     * lose more than 90 % of its slack -- or, for (nearly) active rows with slack <= 1e-12, any of it -- the shared slack variable
     * is raised by the missing amount (an active slack sits at mu / z ~ 1e-13, close to the rounding of the O(1) distances); the filter
     * is reset every iteration because its entries were measured against other neighbour positions. */
    ext_set_nbr(e, nbr);
    double bump = 0.0;
    for (int r = e->K_obs * e->N; r < nd; r++) {
      const int k = e->dstage[r];
      const double cr = pow(x[4 * k] - e->dctr[2 * r], 2) + pow(x[4 * k + 2] - e->dctr[2 * r + 1], 2) + x[n - 1] - e->deps[r];
      const double sp = e->sprev[mi + r];
      bump = fmax(bump, fmax(0.1 * sp, fmin(sp, ORC_EXT_SFLOOR)) - cr);
    }
    if (bump > 0.0) { x[n - 1] += bump; e->n_bump++; }
    e->nfilt = 0;
  }
  const int it = e->it;
  ext_cons(e, x, s);
  ext_jac(e, x, J);
  ext_grad(e, x, gr);
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
  double E0 = fmax(e_d, fmax(e_p, e_c));
  e->E0 = E0;
  if (ext_trace()) fprintf(stderr, "[%p] it %d mu %.2e E0 %.3e ed %.2e ep %.2e ec %.2e dl %.6f\n", (void *)e, it, e->mu, E0, e_d, e_p, e_c, x[n - 1]);
  for (int r = 0; r < m; r++) e->sprev[r] = s[r];
  {
    int spos = 1;
    for (int r = 0; r < m; r++) if (!(s[r] > 0)) spos = 0;
    if (E0 <= o->tol || (e->mu <= o->tol / 10 && E0 <= o->acc_tol)) { e->status = 0; if (e->n_nbr == 0) e->finished = 1; return e->n_nbr == 0 ? 2 : 1; }
    e->status = 1;
    if (it == o->max_iter) { e->finished = 1; return 2; }
    double zmax = 0;
    for (int r = 0; r < m; r++) zmax = fmax(zmax, z[r]);
    if (!(zmax < o->z_div) || !isfinite(E0) || !spos) { e->status = 2; e->finished = 1; return 2; }
  }
  {
    int decreased = 0;
    for (;;) { /* barrier update (Ipopt eq. 7) */
      double emu = 0;
      for (int r = 0; r < m; r++) emu = fmax(emu, fabs(s[r] * z[r] - e->mu));
      emu = fmax(emu, fmax(e_d, e_p));
      if (emu <= o->kappa_eps * e->mu && e->mu > o->tol / 10) {
        e->mu = fmax(o->tol / 10, fmin(o->kappa_mu * e->mu, pow(e->mu, o->theta_mu)));
        e->tau = fmax(o->tau_min, 1 - e->mu);
        e->nfilt = 0;
        decreased = 1;
      } else break;
    }
    /* IPM-F-J only: neighbours that keep moving can hold the barrier subproblem's error just above kappa_eps mu for good (each
     * move costs the agent a little slack, the repair perturbs its dual residual: 0.2 % of the config-4 teams); after mu_stall
     * iterations without a decrease the barrier parameter is lowered anyway */
    if (decreased) e->mu_age = 0;
    else if (e->n_nbr > 0 && ++e->mu_age >= e->mu_stall && e->mu > o->tol / 10) {
      e->mu = fmax(o->tol / 10, fmin(o->kappa_mu * e->mu, pow(e->mu, o->theta_mu)));
      e->tau = fmax(o->tau_min, 1 - e->mu);
      e->mu_age = 0;
    }
  }
  const double mu = e->mu, tau = e->tau;
  double *sig = e->sig, *rhs = e->rhs, *sol = e->sol, *dx = e->dx, *dy = e->dy, *ds = e->ds, *dz = e->dz, *xt = e->xt, *st = e->st;
  const int *isgen = e->isgen, *gen = e->gen;
  for (int r = 0; r < m; r++) sig[r] = z[r] / s[r];
  double dw = 0.0;
  int ok = 0;
  const int lam0 = 6 * e->N;
  for (;;) {
    memset(K, 0, sizeof(double) * (size_t)NK * NK);
    for (int j = 0; j < n; j++) {
      for (int i = j; i < n; i++) K[i + (size_t)j * NK] = Q[i + (size_t)j * n];
      K[j + (size_t)j * NK] += dw;
    }
    for (int j = lam0; j < n - 1; j++) K[j + (size_t)j * NK] += o->rho_lambda;
    for (int r = 0; r < nd; r++) {
      const int k = e->dstage[r];
      K[(4 * k) + (size_t)(4 * k) * NK] -= 2 * z[mi + r];
      K[(4 * k + 2) + (size_t)(4 * k + 2) * NK] -= 2 * z[mi + r];
    }
    for (int r = 0; r < m; r++) {
      if (isgen[r]) continue;
      for (int j = 0; j < n; j++) {
        double jj = J[r + (size_t)j * m];
        if (jj != 0.0) K[j + (size_t)j * NK] += sig[r] * jj * jj;
      }
    }
    for (int j = 0; j < n; j++) {
      for (int i = 0; i < p; i++) K[(n + i) + (size_t)j * NK] = Aeq[i + (size_t)j * p];
      for (int q = 0; q < ng; q++) K[(n + p + q) + (size_t)j * NK] = J[gen[q] + (size_t)j * m];
    }
    for (int q = 0; q < ng; q++) K[(n + p + q) + (size_t)(n + p + q) * NK] = -s[gen[q]] / z[gen[q]];
    e->n_fact++;
    orc_bk_factor(NK, K, NK, e->ipiv);
    int np_, nn_, nz_;
    orc_bk_inertia(NK, K, NK, e->ipiv, &np_, &nn_, &nz_);
    if (np_ == n && nn_ == p + ng && nz_ == 0) { ok = 1; break; }
    if (dw == 0.0) dw = (e->dw_last == 0.0) ? o->dw_first : fmax(o->dw_min, o->kw_minus * e->dw_last);
    else dw *= (e->first_ic && e->dw_last == 0.0) ? o->kw_plus_first : o->kw_plus;
    if (dw > o->dw_max) break;
  }
  if (!ok) { e->status = 3; e->finished = 1; return 2; }
  if (dw > 0) { e->dw_last = dw; e->first_ic = 0; }
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
  orc_bk_solve(NK, K, NK, e->ipiv, sol);
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
  for (int q = 0; q < ng; q++) dz[gen[q]] = -sol[n + p + q];
  double a_pmax = 1.0, a_dmax = 1.0;
  for (int r = 0; r < m; r++) {
    if (ds[r] < 0) a_pmax = fmin(a_pmax, -tau * s[r] / ds[r]);
    if (dz[r] < 0) a_dmax = fmin(a_dmax, -tau * z[r] / dz[r]);
  }
  double ph = ext_obj(e, x), dphi = 0;
  for (int r = 0; r < m; r++) { ph -= mu * log(s[r]); dphi -= mu * ds[r] / s[r]; }
  for (int j = 0; j < n; j++) dphi += gr[j] * dx[j];
  double alpha = a_pmax;
  int accepted = 0, ftype = 0;
  while (alpha >= o->alpha_min) {
    e->n_ls++;
    for (int j = 0; j < n; j++) xt[j] = x[j] + alpha * dx[j];
    ext_cons(e, xt, st);
    double smin = 1e300;
    for (int r = 0; r < m; r++) smin = fmin(smin, st[r]);
    if (smin > 0) {
      double tht = (1 - alpha) * th, pht = ext_obj(e, xt);
      for (int r = 0; r < m; r++) pht -= mu * log(st[r]);
      int okf = tht <= e->theta_max;
      for (int q = 0; q < e->nfilt && okf; q++)
        if (tht >= e->filt_t[q] && pht >= e->filt_p[q]) okf = 0;
      if (okf) {
        int sw = dphi < 0 && alpha * pow(-dphi, o->s_phi) > o->delta * pow(th, o->s_theta);
        if (th <= e->theta_min && sw) {
          if (pht - ph - 10 * EPSM * fabs(ph) <= o->eta_phi * alpha * dphi) { accepted = 1; ftype = 1; }
        } else if (tht <= (1 - o->gamma_theta) * th || pht - ph - 10 * EPSM * fabs(ph) <= -o->gamma_phi * th) {
          accepted = 1; ftype = 0;
        }
      }
    }
    if (accepted) break;
    alpha *= 0.5;
  }
  if (ext_trace()) fprintf(stderr, "[%p]    dw %.1e apmax %.3e admax %.3e ddl %.3e dphi %.3e alpha %.3e acc %d ftype %d ph %.10e nfilt %d th %.2e\n", (void *)e, dw, a_pmax, a_dmax, dx[n - 1], dphi, alpha, accepted, ftype, ph, e->nfilt, th);
  if (!accepted) { e->status = 2; e->finished = 1; return 2; }
  if (!ftype) {
    if (e->nfilt >= o->filter_cap) {
      memmove(e->filt_t, e->filt_t + 1, sizeof(double) * (e->nfilt - 1));
      memmove(e->filt_p, e->filt_p + 1, sizeof(double) * (e->nfilt - 1));
      e->nfilt--;
    }
    e->filt_t[e->nfilt] = (1 - o->gamma_theta) * th;
    e->filt_p[e->nfilt] = ph - o->gamma_phi * th;
    e->nfilt++;
  }
  for (int j = 0; j < n; j++) x[j] += alpha * dx[j];
  for (int i = 0; i < p; i++) y[i] += alpha * dy[i];
  ext_cons(e, x, s);
  for (int r = 0; r < m; r++) {
    double zz = z[r] + a_dmax * dz[r];
    double lo = mu / (o->kappa_sigma * s[r]), hi = o->kappa_sigma * mu / s[r];
    z[r] = fmin(fmax(zz, lo), hi);
    e->sprev[r] = s[r];
  }
  e->it = it + 1;
  return 0;
}

void orc_ext_positions(const orc_ext *e, double *pos) {
  for (int k = 0; k < e->N; k++) { pos[2 * k] = e->x[4 * k]; pos[2 * k + 1] = e->x[4 * k + 2]; }
}

void orc_ext_result_get(const orc_ext *e, orc_ext_result *r) {
  memset(r, 0, sizeof(*r));
  memcpy(r->x, e->x, sizeof(double) * e->nv); memcpy(r->x_qp, e->x_qp, sizeof(double) * e->nv);
  r->kkt = e->E0; r->mu = e->mu; r->N = e->N; r->c = e->c; r->nv = e->nv; r->n_rows = e->K; r->status = e->status;
  r->qp_flag = e->qp_flag; r->qp_iters = e->qp_iters; r->nlp_iters = e->it; r->n_fact = e->n_fact; r->n_ls = e->n_ls; r->n_bump = e->n_bump;
  for (int q = 0; q < e->nd; q++) { r->zd[q] = e->z[e->mi + q]; r->sd[q] = e->s[e->mi + q]; }
}

static void bad_result(orc_ext_result *r, int N) { memset(r, 0, sizeof(*r)); r->status = 4; r->N = N; }

int orc_ext_solve(const orc_gains *g, const orc_options *o, int N, const double *x0, const double *xdes, const double *feet_xy,
                  const int *contact, int K_obs, const double *obs_xy, double eps_obs, int n_nbr, double eps_nbr, const double *nbr,
                  orc_ext_result *res) {
  orc_ext *e = orc_ext_create(g, o, N, x0, xdes, feet_xy, contact, K_obs, obs_xy, eps_obs, n_nbr, eps_nbr);
  if (!e) { bad_result(res, N); return 4; }
  orc_ext_qp(e);
  orc_ext_begin(e, nbr);
  for (;;) {
    int r = orc_ext_iter(e, nbr);
    if (r != 0) break; /* fixed neighbour positions: "converged for these positions" is final */
  }
  orc_ext_result_get(e, res);
  orc_ext_free(e);
  return res->status;
}

int orc_ext_solve_team(const orc_gains *g, const orc_options *o, int T, int N, const double *x0, const double *xdes, const double *feet_xy,
                       const int *contact, int K_obs, const double *obs_xy, double eps_obs, double eps_nbr, double relax, double pub_tol,
                       int max_team_iters, int mu_stall, orc_ext_result *res, int *team_iters) {
  const int ND = N / 4, nn = T - 1;
  orc_ext *ag[8];
  double pos[8][2 * ORC_EXT_NMAX], nbr[7 * 2 * ORC_EXT_NMAX];
  if (T < 1 || T > 8) return -1;
  for (int a = 0; a < T; a++) {
    ag[a] = orc_ext_create(g, o, N, x0 + 4 * a, xdes + 4 * N * a, feet_xy + 8 * ND * a, contact + 4 * a, K_obs, obs_xy + 2 * (size_t)K_obs * a,
                           eps_obs, nn, eps_nbr);
    if (ag[a]) { ag[a]->mu_stall = mu_stall; orc_ext_qp(ag[a]); }
  }
  for (int a = 0; a < T; a++) { /* a member with bad input stays where it was measured */
    if (ag[a]) orc_ext_positions(ag[a], pos[a]);
    else for (int k = 0; k < N; k++) { pos[a][2 * k] = x0[4 * a]; pos[a][2 * k + 1] = x0[4 * a + 2]; }
  }
#define GATHER(a)                                                                                   \
  for (int j = 0, b = 0; b < T; b++) { if (b == (a)) continue; memcpy(nbr + (size_t)j * 2 * N, pos[b], sizeof(double) * 2 * N); j++; }
  for (int a = 0; a < T; a++) { if (!ag[a]) continue; GATHER(a); orc_ext_begin(ag[a], nn > 0 ? nbr : NULL); }
  int iters = 0;
  for (;;) {
    /* published trajectories: pos <- pos + omega (iterate - pos) (under-relaxed Jacobi; omega = 1: the iterates themselves) */
    double moved = 0.0;
    for (int a = 0; a < T; a++) if (ag[a]) {
      double cur[2 * ORC_EXT_NMAX];
      orc_ext_positions(ag[a], cur);
      for (int q = 0; q < 2 * N; q++) { const double d = relax * (cur[q] - pos[a][q]); pos[a][q] += d; moved = fmax(moved, fabs(cur[q] - pos[a][q])); }
    }
    int stepped = 0;
    for (int a = 0; a < T; a++) {
      if (!ag[a]) continue;
      GATHER(a);
      if (orc_ext_iter(ag[a], nn > 0 ? nbr : NULL) == 0) stepped = 1;
    }
    if (!stepped && moved <= pub_tol) break;
    iters++;
    if (iters >= max_team_iters) break;
  }
#undef GATHER
  int all = 1;
  for (int a = 0; a < T; a++) {
    if (ag[a]) { orc_ext_result_get(ag[a], res + a); orc_ext_free(ag[a]); } else bad_result(res + a, N);
    if (res[a].status != 0) all = 0;
  }
  if (team_iters) *team_iters = iters;
  return all ? 0 : 1;
}

typedef struct {
  const orc_gains *g; const orc_options *o; int nteams, T, N, K_obs, tid, nthreads; double eps_obs, eps_nbr, relax, pub_tol; int max_team_iters, mu_stall;
  const double *x0, *xdes, *feet, *obs; const int *contact; orc_ext_result *res; int *team_iters;
} teams_arg;

static void *teams_worker(void *vp) {
  teams_arg *a = (teams_arg *)vp;
  const int T = a->T, N = a->N, ND = N / 4;
  for (int t = a->tid; t < a->nteams; t += a->nthreads)
    orc_ext_solve_team(a->g, a->o, T, N, a->x0 + 4 * (size_t)T * t, a->xdes + 4 * (size_t)N * T * t, a->feet + 8 * (size_t)ND * T * t,
                       a->contact + 4 * (size_t)T * t, a->K_obs, a->obs + 2 * (size_t)a->K_obs * T * t, a->eps_obs, a->eps_nbr, a->relax, a->pub_tol, a->max_team_iters, a->mu_stall,
                       a->res + (size_t)T * t, a->team_iters ? a->team_iters + t : NULL);
  return NULL;
}

int orc_ext_solve_teams(const orc_gains *g, const orc_options *o, int nteams, int T, int N, const double *x0, const double *xdes,
                        const double *feet_xy, const int *contact, int K_obs, const double *obs_xy, double eps_obs, double eps_nbr,
                        double relax, double pub_tol, int max_team_iters, int mu_stall, int nthreads, orc_ext_result *res, int *team_iters) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  teams_arg args[256];
  for (int t = 0; t < nthreads; t++) {
    teams_arg a = {g, o, nteams, T, N, K_obs, t, nthreads, eps_obs, eps_nbr, relax, pub_tol, max_team_iters, mu_stall, x0, xdes, feet_xy, obs_xy, contact, res, team_iters};
    args[t] = a;
    if (nthreads > 1) pthread_create(&th[t], NULL, teams_worker, &args[t]);
  }
  if (nthreads == 1) teams_worker(&args[0]);
  else
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  return 0;
}
