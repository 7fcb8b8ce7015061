/* TEST INFRASTRUCTURE -- thin driver around the reference's own iSWIFT C core, compiled in place from
 * /root/reference/optimization/iSWIFT/{src,ldl/src} by oracle/Makefile into oracle/_ref/libiswift_ref.so.
 * No reference source is copied: this file only restates the Eigen wrapper's job
 * (optimization/iSWIFT/cpp_wrapper/iswift_qp.cpp:78-210: dense -> CCS, a symmetric ordering, QP_SETUP /
 * QP_SOLVE / QP_CLEANUP).  Eigen's AMDOrdering is not available here; the ordering used instead eliminates
 * the slack block first, then the variables, then the equality rows (no zero pivots) -- the solution does not
 * depend on the ordering beyond round-off. */
#include "Prime.h"
#include <string.h>

static void dense_to_ccs(const double *M, int rows, int cols, idxint *jc, idxint *ir, realqp *pr) {
  int nnz = 0;
  for (int j = 0; j < cols; j++) {
    jc[j] = nnz;
    for (int i = 0; i < rows; i++) {
      double v = M[i + (size_t)j * rows];
      if (v != 0.0) { ir[nnz] = i; pr[nnz] = v; nnz++; } /* Eigen sparseView(): exact zeros pruned */
    }
  }
  jc[cols] = nnz;
}

static int count_nnz(const double *M, size_t len) {
  int c = 0;
  for (size_t i = 0; i < len; i++) c += (M[i] != 0.0);
  return c;
}

int iswift_ref_solve_ccs(int n, int m, int p, idxint *Pjc, idxint *Pir, realqp *Ppr, idxint *Ajc, idxint *Air,
                         realqp *Apr, idxint *Gjc, idxint *Gir, realqp *Gpr, realqp *c, realqp *h, realqp *b,
                         realqp sigma_d, idxint *permut, double *x, double *y, double *z, double *s, int *iters,
                         double *t_setup, double *t_solve) {
  QP *qp = QP_SETUP(n, m, p, Pjc, Pir, Ppr, Ajc, Air, Apr, Gjc, Gir, Gpr, c, h, b, sigma_d, permut);
  idxint flag = QP_SOLVE(qp);
  memcpy(x, qp->x, sizeof(double) * n);
  if (y && p) memcpy(y, qp->y, sizeof(double) * p);
  if (z) memcpy(z, qp->z, sizeof(double) * m);
  if (s) memcpy(s, qp->s, sizeof(double) * m);
  if (iters) *iters = (int)qp->stats->IterationCount;
  if (t_setup) *t_setup = qp->stats->tsetup;
  if (t_solve) *t_solve = qp->stats->tsolve;
  QP_CLEANUP(qp);
  return (int)flag;
}

/* dense column-major inputs, as the Eigen matrices MPC_dist passes to iswiftQp_e (src/MPC_dist.cpp:348) */
int iswift_ref_solve_dense(int n, int m, int p, const double *P, const double *c, const double *A, const double *b,
                           const double *G, const double *h, double *x, double *y, double *z, double *s, int *iters,
                           double *t_setup, double *t_solve) {
  int nP = count_nnz(P, (size_t)n * n), nA = count_nnz(A, (size_t)p * n), nG = count_nnz(G, (size_t)m * n);
  idxint *Pjc = malloc(sizeof(idxint) * (n + 1)), *Pir = malloc(sizeof(idxint) * (nP + 1));
  idxint *Ajc = malloc(sizeof(idxint) * (n + 1)), *Air = malloc(sizeof(idxint) * (nA + 1));
  idxint *Gjc = malloc(sizeof(idxint) * (n + 1)), *Gir = malloc(sizeof(idxint) * (nG + 1));
  realqp *Ppr = malloc(sizeof(realqp) * (nP + 1)), *Apr = malloc(sizeof(realqp) * (nA + 1)), *Gpr = malloc(sizeof(realqp) * (nG + 1));
  realqp *cc = malloc(sizeof(realqp) * n), *hh = malloc(sizeof(realqp) * m), *bb = malloc(sizeof(realqp) * (p + 1));
  idxint *perm = malloc(sizeof(idxint) * (n + m + p));
  dense_to_ccs(P, n, n, Pjc, Pir, Ppr);
  dense_to_ccs(A, p, n, Ajc, Air, Apr);
  dense_to_ccs(G, m, n, Gjc, Gir, Gpr);
  memcpy(cc, c, sizeof(realqp) * n); memcpy(hh, h, sizeof(realqp) * m); memcpy(bb, b, sizeof(realqp) * p);
  int k = 0;
  for (int i = 0; i < m; i++) perm[k++] = n + p + i;
  for (int i = 0; i < n; i++) perm[k++] = i;
  for (int i = 0; i < p; i++) perm[k++] = n + i;
  int flag = iswift_ref_solve_ccs(n, m, p, Pjc, Pir, Ppr, Ajc, Air, Apr, Gjc, Gir, Gpr, cc, hh, bb, 0.0, perm, x, y, z, s,
                                  iters, t_setup, t_solve);
  free(Pjc); free(Pir); free(Ppr); free(Ajc); free(Air); free(Apr); free(Gjc); free(Gir); free(Gpr);
  free(cc); free(hh); free(bb); free(perm);
  return flag;
}
