// cbfnmpc_qpdense.h -- batched dense convex QP, iSWIFT's algorithm (optimization/iSWIFT/src/Prime.c:127-230, Auxilary.c), one
// warp per problem.  Second client of the reference's `iswiftQp_e` (optimization/iSWIFT/src/iswift_qp.cpp:44-210): the 1 kHz
// low-level controller QP of src/LowLevelCtrl.cpp:18-52 (31 variables, 18 equalities, 45 inequalities at 4 stance feet)
// has no stage structure, so it is solved dense:
//     min 1/2 x'Px + c'x   s.t.  A x = b,  G x <= h            n <= 32, p <= 24, m <= 64
//   * initial point from the W = I system and the min/max shifts (kkt_initialize, Auxilary.c:680-755);
//   * Mehrotra predictor-corrector, sigma = min(1, rho)^3, step 0.99 alpha_max (Prime.c:138-218);
//   * stop when ||rx||, ||ry||, ||rz|| < reltol / sqrt(3) and s'z / m < abstol (Prime.c:148-160), at most maxit iterations.
// Linear algebra: K = P + G' diag(z/s) G + rho A'A (n x n, Cholesky, lane i owns row i), Schur complement A K^-1 A' (p x p,
// Cholesky).  The rho A'A term (with rho A'r2 added to the right-hand side: the solution is unchanged) makes K positive
// definite whenever the KKT matrix is non-singular, e.g. for the reference's canned QP whose P is only semidefinite on
// variables that no inequality touches.  iSWIFT factors the same quasi-definite system by a sparse LDL' -- same solution,
// different elimination order.
// Everything of one problem lives in that warp's shared memory (52 KB at the size limits).
#pragma once
#include <cuda_runtime.h>

#define QD_NMAX 32
#define QD_PMAX 24
#define QD_MMAX 64
#define QD_LDN 33   // leading dimensions of the row-major matrices: odd, so that "lane i reads row i" is conflict-free (and
#define QD_LDP 25   // "lane j reads column j" and broadcast reads are conflict-free anyway)
#define QD_LDG 33   // G (m x n) and A (p x n); columns n .. 31 are zero-filled so that the register loops can run to 32

struct QDenseIO {
  int B, n, p, m;
  const double *P, *c, *A, *b, *G, *h;   // per problem, contiguous: P n*n, A p*n, G m*n column-major (Eigen .data()); c n, b p, h m
  double reltol, abstol;
  int maxit;
  double* x;      // [B][n]
  double* y;      // optional [B][p]
  double* z;      // optional [B][m]
  int* status;    // [B] 0 converged, 1 factorisation failed, 2 iteration limit
  int* iters;     // [B]
};

struct QDenseWS {                       // one warp
  double P[QD_NMAX * QD_LDN], K[QD_NMAX * QD_LDN];         // row-major n x n (K: lower triangle = Cholesky factor)
  double G[QD_MMAX * QD_LDG];                              // row-major m x n
  double A[QD_PMAX * QD_LDG];                              // row-major p x n
  double V[QD_NMAX * QD_PMAX];                             // L^-1 A'  (n x p, row-major)
  double S[QD_PMAX * QD_LDP];                              // chol(V'V)
  double c[QD_NMAX], b[QD_PMAX], h[QD_MMAX];
  double x[QD_NMAX], y[QD_PMAX], z[QD_MMAX], s[QD_MMAX];
  double rx[QD_NMAX], ry[QD_PMAX], rz[QD_MMAX];
  double dx[QD_NMAX], dy[QD_PMAX], dz[QD_MMAX], ds[QD_MMAX], dsl[QD_MMAX], w[QD_MMAX], iw[QD_MMAX];
  double idK[QD_NMAX], idS[QD_PMAX];                       // 1 / diagonal of the two Cholesky factors
  double rho;
};

#define QDW 0xffffffffu
__device__ __forceinline__ double qd_wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(QDW, v, o);
  return v;
}
__device__ __forceinline__ double qd_wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(QDW, v, o));
  return v;
}
__device__ __forceinline__ double qd_wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(QDW, v, o));
  return v;
}

// Cholesky of a symmetric matrix whose row i sits in lane i's registers (row[0 .. NM), entries beyond the dimension n and the
// strict upper triangle are ignored).  Left-looking: column k needs row k of the factor, which lane k publishes entry by entry
// in shared memory (M, leading dimension ld) as it is produced, so that everybody reads it by broadcast.  On return M holds
// the factor (lower triangle), idg the reciprocals of its diagonal, row[] the lane's row of the factor.
template <int NM>
__device__ __forceinline__ bool qd_chol_reg(double (&row)[NM], double* M, double* idg, int n, int ld, int lane) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < NM; k++) {
    if (k < n) {
      double v = row[k];
#pragma unroll
      for (int j = 0; j < k; j++) v -= row[j] * M[k * ld + j];
      const double dk = __shfl_sync(QDW, v, k);
      if (!(dk > 0.0) || !(dk < 1e300)) ok = false;
      const double ik = rsqrt(dk);
      const double l = (lane == k) ? dk * ik : v * ik;
      row[k] = l;
      if (lane >= k && lane < n) M[lane * ld + k] = l;
      if (lane == k) idg[k] = ik;
      __syncwarp();
    }
  }
  return ok;
}
// v <- L^-1 v / v <- L^-T v for one vector, entry i in lane i's register (shuffles carry the pivots; no barriers)
__device__ __forceinline__ double qd_fwd(const double* L, const double* idg, int n, int ld, double v, int lane) {
  for (int k = 0; k < n; k++) {
    const double vk = __shfl_sync(QDW, v, k) * idg[k];
    if (lane == k) v = vk;
    else if (lane > k && lane < n) v -= L[lane * ld + k] * vk;
  }
  return v;
}
__device__ __forceinline__ double qd_bwd(const double* L, const double* idg, int n, int ld, double v, int lane) {
  for (int k = n - 1; k >= 0; k--) {
    const double vk = __shfl_sync(QDW, v, k) * idg[k];
    if (lane == k) v = vk;
    else if (lane < k) v -= L[k * ld + lane] * vk;
  }
  return v;
}

// factor [P + G'W^-1 G, A'; A, 0] for the weights w (s/z).  Lane i accumulates row i of K = P + G' W^-1 G + rho A'A in
// registers as a sum of outer products (the rows of G and A are read by broadcast), factorises it there (qd_chol_reg), then
// lane q substitutes column q of V = L^-1 A' in registers, accumulates row q of V'V the same way and factorises that.
__device__ __noinline__ bool qd_factor(QDenseWS& ws, int n, int p, int m, int lane) {  // one copy: 70 KB of unrolled code
  for (int r = lane; r < m; r += 32) ws.iw[r] = 1.0 / ws.w[r];
  __syncwarp();
  double kr[QD_NMAX];
#pragma unroll
  for (int j = 0; j < QD_NMAX; j++) kr[j] = ws.P[lane * QD_LDN + j];   // rows / columns n .. 31 of P are zero-filled at load
  double kd = ws.P[lane * QD_LDN + lane], ad = 0.0;                      // diagonal entries of P + G'W^-1 G and of A'A
#pragma unroll 1
  for (int r = 0; r < m; r++) {
    const double gl = ws.G[r * QD_LDG + lane], gi = gl * ws.iw[r];
    kd += gi * gl;
#pragma unroll
    for (int j = 0; j < QD_NMAX; j++) kr[j] += gi * ws.G[r * QD_LDG + j];
  }
  for (int q = 0; q < p; q++) { const double al = ws.A[q * QD_LDG + lane]; ad += al * al; }
  const double skd = qd_wsum(kd), sad = qd_wsum(ad);
  const double rho = (p > 0 && sad > 0.0) ? skd / sad : 0.0;
  if (lane == 0) ws.rho = rho;
#pragma unroll 1
  for (int q = 0; q < p; q++) {
    const double ai = rho * ws.A[q * QD_LDG + lane];
#pragma unroll
    for (int j = 0; j < QD_NMAX; j++) kr[j] += ai * ws.A[q * QD_LDG + j];
  }
  __syncwarp();
  bool ok = qd_chol_reg<QD_NMAX>(kr, ws.K, ws.idK, n, QD_LDN, lane);
  if (p > 0) {
    // V = L^-1 A': lane q substitutes column q (= row q of A) in registers; rows of L by broadcast
    double vq[QD_NMAX];
#pragma unroll
    for (int k = 0; k < QD_NMAX; k++) {
      vq[k] = 0.0;
      if (k < n) {
        double v = (lane < p) ? ws.A[lane * QD_LDG + k] : 0.0;
#pragma unroll
        for (int j = 0; j < k; j++) v -= ws.K[k * QD_LDN + j] * vq[j];
        vq[k] = v * ws.idK[k];
        if (lane < p) ws.V[k * QD_PMAX + lane] = vq[k];
      }
    }
    __syncwarp();
    double sr[QD_PMAX];
#pragma unroll
    for (int j = 0; j < QD_PMAX; j++) sr[j] = 0.0;
#pragma unroll
    for (int i = 0; i < QD_NMAX; i++) {
      if (i < n) {
#pragma unroll
        for (int j = 0; j < QD_PMAX; j++) sr[j] += vq[i] * ws.V[i * QD_PMAX + j];   // columns p .. 23 of V are zero-filled at load
      }
    }
    ok = qd_chol_reg<QD_PMAX>(sr, ws.S, ws.idS, p, QD_LDP, lane) && ok;
  }
  return ok;
}

// solve [P A' G'; A 0 0; G 0 -W][dx; dy; dz] = [r1; r2; r3] with the stored factor.  r3 may be the array dz (entry r is read
// before it is written, by the same lane); dx and dy double as scratch for L^-1 r and for the multiplier step.
__device__ __noinline__ void qd_solve(QDenseWS& ws, int n, int p, int m, const double* r1, const double* r2, const double* r3, double* dx, double* dy,
                         double* dz, int lane) {
  double t = 0.0;
  if (lane < n) {
    t = r1[lane];
    _Pragma("unroll 4")
    for (int r = 0; r < m; r++) t += ws.G[r * QD_LDG + lane] * (r3[r] * ws.iw[r]);
    for (int q = 0; q < p; q++) t += ws.rho * ws.A[q * QD_LDG + lane] * r2[q];
  }
  t = qd_fwd(ws.K, ws.idK, n, QD_LDN, t, lane);                  // t = L^-1 r
  if (p > 0) {
    if (lane < n) dx[lane] = t;
    __syncwarp();
    double u = 0.0;
    if (lane < p) {
      u = -r2[lane];
      _Pragma("unroll 4")
      for (int i = 0; i < n; i++) u += ws.V[i * QD_PMAX + lane] * dx[i];   // V't - r2
    }
    u = qd_fwd(ws.S, ws.idS, p, QD_LDP, u, lane);
    u = qd_bwd(ws.S, ws.idS, p, QD_LDP, u, lane);                // dy
    if (lane < p) dy[lane] = u;
    __syncwarp();
    if (lane < n) for (int q = 0; q < p; q++) t -= ws.V[lane * QD_PMAX + q] * dy[q];
  }
  t = qd_bwd(ws.K, ws.idK, n, QD_LDN, t, lane);                  // dx
  if (lane < n) dx[lane] = t;
  __syncwarp();
  for (int r = lane; r < m; r += 32) {
    double v = -r3[r];
    _Pragma("unroll 4")
    for (int j = 0; j < n; j++) v += ws.G[r * QD_LDG + j] * dx[j];
    dz[r] = v * ws.iw[r];
  }
  __syncwarp();
}

// max step lengths (findsteplength, Auxilary.c:271-294)
__device__ void qd_steplength(const QDenseWS& ws, const double* dsv, const double* dzv, int m, int lane, double& ap, double& ad) {
  double rp = 0.0, rd = 0.0;
  for (int r = lane; r < m; r += 32) { rp = fmax(rp, -dsv[r] / ws.s[r]); rd = fmax(rd, -dzv[r] / ws.z[r]); }
  rp = qd_wmax(rp); rd = qd_wmax(rd);
  ap = (rp > 0.0) ? 1.0 / rp : 1.0; ad = (rd > 0.0) ? 1.0 / rd : 1.0;
}

// valid == false: a warp without a problem (tail of the batch); it only takes part in the CTA-wide votes
__device__ void qdense_solve_warp(QDenseWS& ws, const QDenseIO& io, int q, bool valid) {
  const int lane = threadIdx.x & 31, n = io.n, p = io.p, m = io.m;
  int flag = 3, it = 0;
  bool run = false;
  if (valid) {
    // ---- load (column-major global -> row-major shared, zero-filled up to the register-loop bounds)
    const double* Pg = io.P + (size_t)q * n * n; const double* Ag = io.A + (size_t)q * p * n; const double* Gg = io.G + (size_t)q * m * n;
    for (int t = lane; t < QD_NMAX * QD_LDN; t += 32) ws.P[t] = 0.0;
    for (int t = lane; t < QD_MMAX * QD_LDG; t += 32) ws.G[t] = 0.0;
    for (int t = lane; t < QD_PMAX * QD_LDG; t += 32) ws.A[t] = 0.0;
    for (int t = lane; t < QD_NMAX * QD_PMAX; t += 32) ws.V[t] = 0.0;
    __syncwarp();
    for (int t = lane; t < n * n; t += 32) { const int i = t % n, j = t / n; ws.P[i * QD_LDN + j] = Pg[t]; }
    for (int t = lane; t < p * n; t += 32) { const int i = t % p, j = t / p; ws.A[i * QD_LDG + j] = Ag[t]; }
    for (int t = lane; t < m * n; t += 32) { const int i = t % m, j = t / m; ws.G[i * QD_LDG + j] = Gg[t]; }
    for (int t = lane; t < n; t += 32) ws.c[t] = io.c[(size_t)q * n + t];
    for (int t = lane; t < p; t += 32) ws.b[t] = io.b[(size_t)q * p + t];
    for (int t = lane; t < m; t += 32) ws.h[t] = io.h[(size_t)q * m + t];
    __syncwarp();
    // ---- kkt_initialize: W = I, rhs [-c; b; h]
    for (int r = lane; r < m; r += 32) ws.w[r] = 1.0;
    if (lane < n) ws.rx[lane] = -ws.c[lane];
    __syncwarp();
    if (!qd_factor(ws, n, p, m, lane)) flag = 1;
    else {
      qd_solve(ws, n, p, m, ws.rx, ws.b, ws.h, ws.x, ws.y, ws.dz, lane);
      double rmin = 1e300, rmax = -1e300;
      for (int r = lane; r < m; r += 32) {
        double t = ws.h[r];
        for (int j = 0; j < n; j++) t -= ws.G[r * QD_LDG + j] * ws.x[j];
        ws.rz[r] = t; rmin = fmin(rmin, t); rmax = fmax(rmax, t);
      }
      rmin = qd_wmin(rmin); rmax = qd_wmax(rmax);
      const double ap0 = -rmin, ad0 = rmax;
      for (int r = lane; r < m; r += 32) {
        ws.s[r] = (ap0 < 0) ? ws.rz[r] : ws.rz[r] + (1 + ap0);
        ws.z[r] = (ad0 < 0) ? -ws.rz[r] : -ws.rz[r] + (1 + ad0);
      }
      __syncwarp();
      flag = 2; run = true;
    }
  }
  double sigma = 100.0;  // GlobalOptions.h:26
  // The warps of a CTA start every iteration together, so that they fetch the same instruction-cache lines (the unrolled
  // register loops of qd_factor are 0.4 MB of code).
  for (;;) {
    const bool act = run && it < io.maxit;
    if (!__syncthreads_or(act ? 1 : 0)) break;
    if (!act) continue;
    // ---- residuals (computeresiduals, Auxilary.c:524-553)
    double nrx = 0.0, nry = 0.0, nrz = 0.0, sz = 0.0;
    if (lane < n) {
      double t = ws.c[lane];
      _Pragma("unroll 4")
      for (int i = 0; i < n; i++) t += ws.P[lane * QD_LDN + i] * ws.x[i];
      _Pragma("unroll 4")
      for (int r = 0; r < m; r++) t += ws.G[r * QD_LDG + lane] * ws.z[r];
      for (int i = 0; i < p; i++) t += ws.A[i * QD_LDG + lane] * ws.y[i];
      ws.rx[lane] = -t; nrx = t * t;
    }
    if (lane < p) {
      double t = ws.b[lane];
      for (int j = 0; j < n; j++) t -= ws.A[lane * QD_LDG + j] * ws.x[j];
      ws.ry[lane] = t; nry = t * t;
    }
    for (int r = lane; r < m; r += 32) {
      double t = ws.h[r] - ws.s[r];
      _Pragma("unroll 4")
      for (int j = 0; j < n; j++) t -= ws.G[r * QD_LDG + j] * ws.x[j];
      ws.rz[r] = t; nrz += t * t; sz += ws.s[r] * ws.z[r];
    }
    nrx = sqrt(qd_wsum(nrx)); nry = sqrt(qd_wsum(nry)); nrz = sqrt(qd_wsum(nrz)); sz = qd_wsum(sz);
    const double lim = io.reltol / sqrt(3.0);
    if (nrx < lim && nrz < lim && (p == 0 || nry < lim) && sz / m < io.abstol) { flag = 0; run = false; continue; }
    // lambda^2 = (sqrt(s z))^2 as iSWIFT forms it; entry r belongs to lane r % 32 throughout
    double mu = 0.0;
    for (int r = lane; r < m; r += 32) { const double l = sqrt(ws.s[r] * ws.z[r]); mu += l * l; }
    mu = qd_wsum(mu) / m;
    if (sigma > 0.0) {
      for (int r = lane; r < m; r += 32) {
        const double l = sqrt(ws.s[r] * ws.z[r]);
        ws.w[r] = ws.s[r] / ws.z[r]; ws.ds[r] = -(l * l); ws.dz[r] = ws.rz[r] - ws.ds[r] / ws.z[r];
      }
      __syncwarp();
      if (!qd_factor(ws, n, p, m, lane)) { flag = 1; run = false; continue; }
      qd_solve(ws, n, p, m, ws.rx, ws.ry, ws.dz, ws.dx, ws.dy, ws.dz, lane);
      for (int r = lane; r < m; r += 32) ws.dsl[r] = (ws.ds[r] - ws.s[r] * ws.dz[r]) / ws.z[r];
      __syncwarp();
      double ap, ad;
      qd_steplength(ws, ws.dsl, ws.dz, m, lane, ap, ad);
      double num = 0.0;
      for (int r = lane; r < m; r += 32) num += (ws.s[r] + ap * ws.dsl[r]) * (ws.z[r] + ad * ws.dz[r]);
      num = qd_wsum(num);
      const double rho = num / sz, mr = rho < 1.0 ? rho : 1.0;
      sigma = mr * mr * mr;
      if (sigma < 0.0) sigma = 0.0;
      for (int r = lane; r < m; r += 32) { const double l = sqrt(ws.s[r] * ws.z[r]); ws.ds[r] = -(l * l) - ws.dsl[r] * ws.dz[r] + sigma * mu; }
    } else {
      sigma = 0.0;
      for (int r = lane; r < m; r += 32) { const double l = sqrt(ws.s[r] * ws.z[r]); ws.ds[r] = -(l * l) + sigma * mu; }
    }
    __syncwarp();
    for (int r = lane; r < m; r += 32) ws.dz[r] = ws.rz[r] - ws.ds[r] / ws.z[r];
    __syncwarp();
    qd_solve(ws, n, p, m, ws.rx, ws.ry, ws.dz, ws.dx, ws.dy, ws.dz, lane);
    for (int r = lane; r < m; r += 32) ws.dsl[r] = (ws.ds[r] - ws.s[r] * ws.dz[r]) / ws.z[r];
    __syncwarp();
    double ap, ad;
    qd_steplength(ws, ws.dsl, ws.dz, m, lane, ap, ad);
    ap = fmin(0.99 * ap, 1.0); ad = fmin(0.99 * ad, 1.0);
    if (lane < n) ws.x[lane] += ap * ws.dx[lane];
    if (lane < p) ws.y[lane] += ad * ws.dy[lane];
    for (int r = lane; r < m; r += 32) { ws.s[r] += ap * ws.dsl[r]; ws.z[r] += ad * ws.dz[r]; }
    __syncwarp();
    it++;
  }
  if (!valid) return;
  // ---- store
  for (int t = lane; t < n; t += 32) io.x[(size_t)q * n + t] = ws.x[t];
  if (io.y) for (int t = lane; t < p; t += 32) io.y[(size_t)q * p + t] = ws.y[t];
  if (io.z) for (int t = lane; t < m; t += 32) io.z[(size_t)q * m + t] = ws.z[t];
  if (lane == 0) { io.status[q] = (flag == 0) ? 0 : ((flag == 1) ? 1 : 2); io.iters[q] = it; }
}
