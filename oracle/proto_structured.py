"""TEST INFRASTRUCTURE -- numpy model of the *structured* Newton-step algebra the CUDA kernel uses.

Stage-wise Riccati recursion over the horizon with
  * (u_j, lambda_j) eliminated per stage through u = F lambda, 1'lambda = 1 (residuals carried, so the
    step equals the Newton step of the full KKT system of SURVEY.md Appendix A),
  * all single-variable bound rows and the friction rows condensed into the stage Hessians,
  * the 4 distance rows split  Sigma = min(Sigma, cap) (condensed) + remainder (kept as a 4-row border
    together with the global slack variable), solved by a 5x5 Schur complement.
It is checked against the dense augmented-system direction of oracle/proto_numpy.py in
tests/test_structured_model.py.  Nothing here is used by the product path.
"""
import numpy as np
from proto_numpy import (lip_discrete, FRIC, BOX, VMAX, EPS_DIST, QW, RW, PW, SLW)

SIGMA_CAP = 1e8


class Stagewise:
    """Holds per-instance constants: F (2 x c), x0, xdes, obstacle."""

    def __init__(self, x0, xdes, F, obs):
        self.x0, self.xdes, self.F, self.obs = x0, xdes, F, obs
        self.c = F.shape[1]
        Ad, Bd = lip_discrete()
        self.A, self.B = Ad, Bd
        c = self.c
        self.E = np.vstack([np.eye(c - 1), -np.ones((1, c - 1))])
        self.D = F @ self.E
        self.Bt = Bd @ self.D
        self.qdiag = np.array([[QW] * 4] * 3 + [[PW] * 4])


def chol_lower(M):
    """Plain Cholesky; None unless every pivot is finite and > 0 (the inertia test of DESIGN.md)."""
    n = M.shape[0]
    L = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            v = M[i, j] - L[i, :j] @ L[j, :j]
            if i == j:
                if not (v > 0.0) or not np.isfinite(v):
                    return None
                L[i, i] = np.sqrt(v)
            else:
                L[i, j] = v / L[j, j]
    return L


def riccati_factor(st, hx, hxy, hu, hxu, hl):
    """hx[4][4] diag of x_1..x_4 Hessians, hxy[4] (px,py) off-diagonal, hu[4][2], hxu[4][2] (stage 0 unused),
    hl[4][c].  Returns factor dict or None when some pivot block is not positive definite."""
    A, Bt, D, E = st.A, st.Bt, st.D, st.E
    nw = st.c - 1
    P = [None] * 5; K = [None] * 4; L = [None] * 4; St = [None] * 4
    Hx4 = np.diag(hx[3]); Hx4[0, 2] = Hx4[2, 0] = hxy[3]
    P[4] = Hx4
    for j in (3, 2, 1, 0):
        Rw = D.T @ np.diag(hu[j]) @ D + E.T @ np.diag(hl[j]) @ E
        Rt = Rw + Bt.T @ P[j + 1] @ Bt
        L[j] = chol_lower(Rt)
        if L[j] is None:
            return None
        if j == 0:
            break
        S = np.zeros((nw, 4))
        S[:, 0] = D[0] * hxu[j][0]
        S[:, 2] = D[1] * hxu[j][1]
        St[j] = S + Bt.T @ P[j + 1] @ A
        K[j] = np.linalg.solve(L[j].T, np.linalg.solve(L[j], St[j]))
        Hx = np.diag(hx[j - 1]); Hx[0, 2] = Hx[2, 0] = hxy[j - 1]
        Pj = Hx + A.T @ P[j + 1] @ A - St[j].T @ K[j]
        P[j] = 0.5 * (Pj + Pj.T)
    return dict(P=P, K=K, L=L, St=St)


def riccati_solve(st, fac, rx, rw, cvec):
    """rx[4][4] (stages 1..4), rw[4][nw], cvec[4][4] affine dynamics terms (dx_{j+1} = A dx_j + Bt w_j + cvec[j]).
    Returns dx[4][4] (x_1..x_4), w[4][nw], dpi[4][4] (costates of x_1..x_4)."""
    A, Bt = st.A, st.Bt
    P, K, L, St = fac['P'], fac['K'], fac['L'], fac['St']
    p = [None] * 5; k = [None] * 4
    p[4] = rx[3].copy()
    for j in (3, 2, 1, 0):
        q = P[j + 1] @ cvec[j] - p[j + 1]
        rt = rw[j] - Bt.T @ q
        k[j] = np.linalg.solve(L[j].T, np.linalg.solve(L[j], rt))
        if j == 0:
            break
        p[j] = rx[j - 1] - A.T @ q - St[j].T @ k[j]
    dx = np.zeros((4, 4)); w = np.zeros((4, st.c - 1)); dpi = np.zeros((4, 4))
    xprev = np.zeros(4)
    for j in range(4):
        w[j] = k[j] - (K[j] @ xprev if j > 0 else 0.0)
        xn = A @ xprev + Bt @ w[j] + cvec[j]
        dx[j] = xn
        dpi[j] = P[j + 1] @ xn - p[j + 1]
        xprev = xn
    return dx, w, dpi


class StructuredKKT:
    """factor(): assemble + Riccati/border factorisation at (x, z, s, dw); solve(r1, r2, rc): one pass for
          [W+dw+rho  A'  -J'] [dx]   [r1]
          [A         0    0 ] [dy] = [r2]        ds = J dx
          [Z J       0    S ] [dz]   [rc]
    apply(dx, dy, dz): left-hand side of that system (for iterative refinement)."""

    def __init__(self, st, xvec, z, s, dw, rho_lambda, with_nl=True, sigma_cap=SIGMA_CAP):
        self.st = st; self.with_nl = with_nl
        c = st.c; self.c = c; self.nv = 25 + 4 * c; self.mG = 60 + 8 * c
        self.z, self.s, self.dw, self.rho = z, s, dw, rho_lambda
        X = xvec[:16].reshape(4, 4)
        mG = self.mG
        Sig = z / s
        self.Sig = Sig
        hx = st.qdiag.copy() + dw; hxy = np.zeros(4)
        hu = np.full((4, 2), RW + dw); hxu = np.zeros((4, 2)); hl = np.full((4, c), dw + rho_lambda)
        for j in range(16):
            hx[j // 4, j % 4] += Sig[12 + j] + Sig[28 + j]
        for j in range(8):
            hu[j // 2, j % 2] += Sig[44 + j] + Sig[52 + j]
        for j in range(4 * c):
            hl[j // c, j % c] += Sig[60 + j] + Sig[60 + 4 * c + j]
        for i in range(3):
            for a in range(2):
                sf = Sig[i * 2 + a] + Sig[6 + i * 2 + a]
                hx[i, 2 * a] += sf; hu[i + 1, a] += sf; hxu[i + 1, a] = -sf
        hd = SLW + dw; bd = np.zeros((4, 2)); gvec = np.zeros((4, 2)); sr_inv = np.zeros(4); capped = np.zeros(4, bool)
        sc_ = np.zeros(4)
        if with_nl:
            for r, j in enumerate(VEL):
                hx[j // 4, j % 4] += Sig[mG + 4 + r] + Sig[mG + 12 + r]
            for k in range(4):
                g = np.array([2 * (X[k, 0] - st.obs[0]), 2 * (X[k, 2] - st.obs[1])])
                gvec[k] = g
                zk = z[mG + k]; sg = Sig[mG + k]
                sc = min(sg, sigma_cap); sc_[k] = sc
                capped[k] = sg > sigma_cap
                sr_inv[k] = 1.0 / (sg - sc) if capped[k] else 0.0
                hx[k, 0] += sc * g[0] * g[0] - 2 * zk
                hx[k, 2] += sc * g[1] * g[1] - 2 * zk
                hxy[k] += sc * g[0] * g[1]
                hd += sc
                bd[k] = sc * g
        self.hx, self.hxy, self.hu, self.hxu, self.hl, self.hd, self.bd = hx, hxy, hu, hxu, hl, hd, bd
        self.gvec, self.sr_inv, self.capped, self.sc = gvec, sr_inv, capped, sc_
        self.ok = False
        self.fac = riccati_factor(st, hx, hxy, hu, hxu, hl)
        if self.fac is None:
            return
        if with_nl:
            nw = c - 1
            zero_w = np.zeros((4, nw)); zero_c = np.zeros((4, 4))
            rb = np.zeros((4, 4)); rb[:, 0] = bd[:, 0]; rb[:, 2] = bd[:, 1]
            self.sd = riccati_solve(st, self.fac, rb, zero_w, zero_c)
            self.sk = []
            for k in range(4):
                rk = np.zeros((4, 4)); rk[k, 0] = gvec[k, 0]; rk[k, 2] = gvec[k, 1]
                self.sk.append(riccati_solve(st, self.fac, rk, zero_w, zero_c))
            pos = lambda dxs, j: np.array([dxs[j, 0], dxs[j, 2]])
            T = np.zeros((4, 4)); cv = np.zeros(4)
            for k in range(4):
                for l in range(4):
                    T[k, l] = gvec[k] @ pos(self.sk[l][0], k)
                if capped[k]:
                    T[k, k] += sr_inv[k]
                cv[k] = 1.0 - gvec[k] @ pos(self.sd[0], k)
            self.a = hd - sum(bd[j] @ pos(self.sd[0], j) for j in range(4))
            self.T, self.cv = T, cv
            act = capped
            self.den = self.a
            if act.any():
                Ta = T[np.ix_(act, act)]
                Ta = 0.5 * (Ta + Ta.T)
                if chol_lower(Ta) is None:
                    return
                self.Ta = Ta
                self.Tic = np.linalg.solve(Ta, cv[act])
                self.den = self.a + cv[act] @ self.Tic
            if not (self.den > 0):
                return
        self.ok = True

    # J' v for a row vector v (reference row order), returned per variable block
    def _jt(self, X, v):
        c, mG = self.c, self.mG
        gx = np.zeros((4, 4)); gu = np.zeros((4, 2)); gl = np.zeros((4, c)); gd = 0.0
        for j in range(16):
            gx[j // 4, j % 4] += -v[12 + j] + v[28 + j]
        for j in range(8):
            gu[j // 2, j % 2] += -v[44 + j] + v[52 + j]
        for j in range(4 * c):
            gl[j // c, j % c] += v[60 + j] - v[60 + 4 * c + j]
        for i in range(3):
            for a in range(2):
                gf = -v[i * 2 + a] + v[6 + i * 2 + a]
                gx[i, 2 * a] += gf; gu[i + 1, a] -= gf
        if self.with_nl:
            for r, j in enumerate(VEL):
                gx[j // 4, j % 4] += -v[mG + 4 + r] + v[mG + 12 + r]
            for k in range(4):
                gx[k, 0] += v[mG + k] * self.gvec[k, 0]; gx[k, 2] += v[mG + k] * self.gvec[k, 1]
                gd += v[mG + k]
        return gx, gu, gl, gd

    def _j(self, dxv):
        """J dx (row order)."""
        c, mG = self.c, self.mG
        DX = dxv[:16].reshape(4, 4); DU = dxv[16:24].reshape(4, 2); DL = dxv[24:24 + 4 * c]
        m = mG + (20 if self.with_nl else 0)
        out = np.zeros(m)
        for i in range(3):
            for a in range(2):
                e = DX[i, 2 * a] - DU[i + 1, a]
                out[i * 2 + a] = -e; out[6 + i * 2 + a] = e
        out[12:28] = -dxv[:16]; out[28:44] = dxv[:16]
        out[44:52] = -dxv[16:24]; out[52:60] = dxv[16:24]
        out[60:60 + 4 * c] = DL; out[60 + 4 * c:60 + 8 * c] = -DL
        if self.with_nl:
            for k in range(4):
                out[mG + k] = self.gvec[k, 0] * DX[k, 0] + self.gvec[k, 1] * DX[k, 2] + dxv[-1]
            out[mG + 4:mG + 12] = -dxv[VEL]; out[mG + 12:mG + 20] = dxv[VEL]
        return out

    def apply(self, dxv, dyv, dzv):
        st = self.st; c = self.c; A, B, F = st.A, st.B, st.F
        DX = dxv[:16].reshape(4, 4); DU = dxv[16:24].reshape(4, 2); DL = dxv[24:24 + 4 * c].reshape(4, c)
        PI = dyv[:16].reshape(4, 4); NU = dyv[16:24].reshape(4, 2); ETA = dyv[24:28]
        gx, gu, gl, gd = self._jt(None, dzv)
        o1 = np.zeros(self.nv); O1x = o1[:16].reshape(4, 4); O1u = o1[16:24].reshape(4, 2); O1l = o1[24:24 + 4 * c].reshape(4, c)
        for j in range(4):
            w = (st.qdiag[j] + self.dw) * DX[j]
            if self.with_nl:
                zk = self.z[self.mG + j]
                w[0] -= 2 * zk * DX[j, 0]; w[2] -= 2 * zk * DX[j, 2]
            O1x[j] = w - PI[j] + (A.T @ PI[j + 1] if j < 3 else 0.0) - gx[j]
            O1u[j] = (RW + self.dw) * DU[j] + B.T @ PI[j] + NU[j] - gu[j]
            O1l[j] = (self.dw + self.rho) * DL[j] - F.T @ NU[j] + ETA[j] - gl[j]
        o1[-1] = (SLW + self.dw) * dxv[-1] - gd
        o2 = np.zeros(28)
        xprev = np.zeros(4)
        for j in range(4):
            o2[4 * j:4 * j + 4] = A @ xprev - DX[j] + B @ DU[j]
            o2[16 + 2 * j:18 + 2 * j] = DU[j] - F @ DL[j]
            o2[24 + j] = DL[j].sum()
            xprev = DX[j]
        oc = self.s * dzv + self.z * self._j(dxv)
        return o1, o2, oc

    def solve(self, r1, r2, rc):
        st = self.st; c = self.c; nw = c - 1
        A, B, F, D, E = st.A, st.B, st.F, st.D, st.E
        hx, hu, hxu, hl = self.hx, self.hu, self.hxu, self.hl
        gx, gu, gl, gd = self._jt(None, rc / self.s)
        rx = r1[:16].reshape(4, 4) + gx; ru = r1[16:24].reshape(4, 2) + gu; rl = r1[24:24 + 4 * c].reshape(4, c) + gl
        rdl = r1[-1] + gd
        re = -r2[:16].reshape(4, 4); rnu = -r2[16:24].reshape(4, 2); reta = -r2[24:28]
        e0 = np.zeros((4, c)); d0 = np.zeros((4, 2)); rw = np.zeros((4, nw)); cvec = np.zeros((4, 4)); rxp = rx.copy()
        for j in range(4):
            e0[j, c - 1] = -reta[j]
            d0[j] = F @ e0[j] - rnu[j]
            rw[j] = D.T @ (ru[j] - hu[j] * d0[j]) + E.T @ (rl[j] - hl[j] * e0[j])
            cvec[j] = B @ d0[j] + re[j]
            if j >= 1:
                rxp[j - 1, 0] -= hxu[j, 0] * d0[j, 0]
                rxp[j - 1, 2] -= hxu[j, 1] * d0[j, 1]
        dx0, w0, dpi0 = riccati_solve(st, self.fac, rxp, rw, cvec)
        t = np.zeros(4)
        if self.with_nl:
            pos = lambda dxs, j: np.array([dxs[j, 0], dxs[j, 2]])
            rt = np.array([self.gvec[k] @ pos(dx0, k) for k in range(4)])
            rd = rdl - sum(self.bd[j] @ pos(dx0, j) for j in range(4))
            act = self.capped
            num = rd
            if act.any():
                Tir = np.linalg.solve(self.Ta, rt[act])
                num = rd - self.cv[act] @ Tir
            ddl = num / self.den
            if act.any():
                t[act] = self.Tic * ddl + Tir
            dxs = dx0 - self.sd[0] * ddl - sum(self.sk[k][0] * t[k] for k in range(4))
            ws = w0 - self.sd[1] * ddl - sum(self.sk[k][1] * t[k] for k in range(4))
            dpis = dpi0 - self.sd[2] * ddl - sum(self.sk[k][2] * t[k] for k in range(4))
        else:
            ddl = rdl / self.hd
            dxs, ws, dpis = dx0, w0, dpi0
        dlam = np.zeros((4, c)); du = np.zeros((4, 2)); dnu = np.zeros((4, 2)); deta = np.zeros(4)
        for j in range(4):
            dlam[j] = E @ ws[j] + e0[j]
            du[j] = D @ ws[j] + d0[j]
            coup = np.zeros(2)
            if j >= 1:
                coup = hxu[j] * np.array([dxs[j - 1, 0], dxs[j - 1, 2]])
            dnu[j] = ru[j] - hu[j] * du[j] - coup - B.T @ dpis[j]
            q = hl[j] * dlam[j] - F.T @ dnu[j] - rl[j]
            deta[j] = -q[c - 1]
        dxv = np.concatenate([dxs.ravel(), du.ravel(), dlam.ravel(), [ddl]])
        dyv = np.concatenate([dpis.ravel(), dnu.ravel(), deta])
        ds = self._j(dxv)
        dzv = (rc - self.z * ds) / self.s
        if self.with_nl:
            for k in range(4):
                if self.capped[k]:
                    dzv[self.mG + k] = rc[self.mG + k] / self.s[self.mG + k] - self.sc[k] * ds[self.mG + k] - t[k]
        return dxv, dyv, dzv


VEL = [1, 5, 9, 13, 3, 7, 11, 15]


def structured_direction(st, xvec, y, z, s, mu, dw, rho_lambda, with_nl=True, sigma_cap=SIGMA_CAP, refine=2, stats=None):
    """Newton direction of the NLP-stage barrier problem at (xvec, y, z), s = c(x) > 0, with `refine`
    passes of iterative refinement on the unreduced system.  None when the inertia is wrong."""
    kk = StructuredKKT(st, xvec, z, s, dw, rho_lambda, with_nl, sigma_cap)
    if not kk.ok:
        return None
    c = st.c
    X = xvec[:16].reshape(4, 4); U = xvec[16:24].reshape(4, 2); Lm = xvec[24:24 + 4 * c].reshape(4, c); dl = xvec[-1]
    A, B, F = st.A, st.B, st.F
    PI = y[:16].reshape(4, 4); NU = y[16:24].reshape(4, 2); ETA = y[24:28]
    gx, gu, gl, gd = kk._jt(None, z)
    r1 = np.zeros(kk.nv); R1x = r1[:16].reshape(4, 4); R1u = r1[16:24].reshape(4, 2); R1l = r1[24:24 + 4 * c].reshape(4, c)
    for j in range(4):
        gradf = st.qdiag[j] * (X[j] - st.xdes[4 * j:4 * j + 4])
        R1x[j] = -(gradf - PI[j] + (A.T @ PI[j + 1] if j < 3 else 0.0) - gx[j])
        R1u[j] = -(RW * U[j] + B.T @ PI[j] + NU[j] - gu[j])
        R1l[j] = -(-F.T @ NU[j] + ETA[j] - gl[j])
    r1[-1] = -(SLW * dl - gd)
    r2 = np.zeros(28)
    xprev = st.x0
    for j in range(4):
        r2[4 * j:4 * j + 4] = -(A @ xprev - X[j] + B @ U[j])
        r2[16 + 2 * j:18 + 2 * j] = -(U[j] - F @ Lm[j])
        r2[24 + j] = -(Lm[j].sum() - 1.0)
        xprev = X[j]
    rc = mu - s * z
    dxv, dyv, dzv = kk.solve(r1, r2, rc)
    for it in range(refine):
        o1, o2, oc = kk.apply(dxv, dyv, dzv)
        e1, e2, ec = r1 - o1, r2 - o2, rc - oc
        if stats is not None:
            stats.append((it, np.abs(e1).max(), np.abs(e2).max(), np.abs(ec).max(), np.abs(r1).max()))
        ddx, ddy, ddz = kk.solve(e1, e2, ec)
        dxv += ddx; dyv += ddy; dzv += ddz
    return dxv, dyv, dzv, kk.capped
