"""TEST INFRASTRUCTURE (oracle side) -- numpy restatement of the reference hot path.

Independent, dense, structure-agnostic numpy version of what MPC_dist::run_NMPC() poses and solves.
Used to (a) prototype / pin the algorithm that oracle/cbfnmpc_oracle.c and the CUDA path implement,
(b) cross-check the C oracle on small cases, (c) generate golden vectors together with scipy.

Reference lines restated (paths relative to /root/reference):
  lip_discrete        src/MPC_dist.cpp:99-127
  pack_dense          src/MPC_dist.cpp:132-321 (+ footholdsPlanner 1204-1266 for the F block)
  mehrotra_qp         optimization/iSWIFT/src/Prime.c:127-230, src/Auxilary.c (kkt_initialize 680-755,
                      form_ds 250-267, findsteplength 271-294, formrho 600-609)
  nlp rows            include/dec_vars_constr_cost.h:245-324 (values/bounds), 326-391 (Jacobian), 423-438 (cost)
  closest_obstacle    src/MPC_dist.cpp:371-396
  bezier_fit          src/MPC_dist.cpp:784-855, include/math_define.h:409-414

parity unpinned for the NLP stage: the reference solves it with ifopt+SNOPT, which is not in
/root/reference and not installable here; `nlp_ipm` is this repo's own primal-dual filter IPM
(the algorithm DESIGN.md specifies) on the restated NLP.
"""
import numpy as np

G_ACC = 9.81
H_COM = 0.29
TS = 0.043
MU = 0.7
EPS_DIST = float(np.float32(1.9))      # DV:401  `const float epsilon = 1.9`
VMAX = float(np.float32(0.35))         # DV:306-307 `float x_vel_sat = 0.35`
FRIC = MU * H_COM / np.sqrt(2.0)       # MD:315
EPSM = 2.220446049250313e-16
BOX = 1e3                              # MD:317-318
QW, RW, PW, SLW = 3e2, 1e-1, 2e3, 0.3e4  # MD:172-175
NGRID = 4


def lip_discrete(g=G_ACC, h=H_COM, Ts=TS):
    A = np.zeros((4, 4)); B = np.zeros((4, 2))
    A[0, 1] = 1; A[1, 0] = g / h; A[2, 3] = 1; A[3, 2] = g / h
    B[1, 0] = -g / h; B[3, 1] = -g / h
    Ad = np.eye(4) + A * Ts + 0.5 * A @ A * Ts * Ts + A @ A @ A * Ts * Ts * Ts / 6
    Bd = np.linalg.inv(A) @ (Ad - np.eye(4)) @ B
    return Ad, Bd


def stance_matrix(feet_xy, contact):
    """feet_xy: (2,4) toe xy (FR,FL,RR,RL); contact: 4 ints. MD:1230-1252."""
    c = int(sum(contact))
    if c == 4:
        return feet_xy.copy()
    if c == 2 and contact[0] == 0:
        return feet_xy[:, [1, 2]].copy()
    if c == 2 and contact[0] == 1:
        return feet_xy[:, [0, 3]].copy()
    raise ValueError("unsupported contact pattern")


def pack_dense(x0, xdes, F):
    """Dense Q_qp,f_qp,Aeq,beq,Gineq,hineq exactly as MD:164-321. F is 2 x c."""
    n, m, N = 4, 2, NGRID
    c = F.shape[1]
    nv = n * N + m * N + N * c + 1
    Ad, Bd = lip_discrete()
    Pbig = np.zeros((n * N, n * N))
    for i in range(N):
        Pbig[n * i:n * i + n, n * i:n * i + n] = QW * np.eye(n)
    Pbig[n * (N - 1):, n * (N - 1):] = PW * np.eye(n)
    Q = np.zeros((nv, nv))
    Q[:n * N, :n * N] = Pbig
    Q[n * N:n * N + m * N, n * N:n * N + m * N] = RW * np.eye(m * N)
    Q[-1, -1] = SLW
    f = np.zeros(nv)
    f[:n * N] = -(xdes @ Pbig)
    # equalities
    Abig = np.zeros((n * N, n * N)); Bbig = np.zeros((n * N, m * N))
    for i in range(N - 1):
        Abig[(i + 1) * n:(i + 2) * n, i * n:(i + 1) * n] = Ad
    for i in range(N):
        Bbig[i * n:(i + 1) * n, i * m:(i + 1) * m] = Bd
    p_eq = n * N + m * N + N
    Aeq = np.zeros((p_eq, nv)); beq = np.zeros(p_eq)
    Aeq[:n * N, :n * N] = Abig - np.eye(n * N)
    Aeq[:n * N, n * N:n * N + m * N] = Bbig
    Aeq[n * N:n * N + m * N, n * N:n * N + m * N] = np.eye(m * N)
    for k in range(N):
        Aeq[n * N + 2 * k:n * N + 2 * k + 2, n * N + m * N + c * k:n * N + m * N + c * (k + 1)] = -F
        Aeq[n * N + m * N + k, n * N + m * N + c * k:n * N + m * N + c * (k + 1)] = 1.0
    beq[:n] = -(Ad @ x0)
    beq[n * N + m * N:] = 1.0
    # inequalities
    mi = 2 * m * (N - 1) + 2 * n * N + 2 * m * N + 2 * N * c
    G = np.zeros((mi, nv)); h = np.ones(mi)
    Gsubx = np.zeros((m * (N - 1), n * N)); Gsubu = np.zeros((m * (N - 1), m * N))
    xportion = np.array([[1.0, 0, 0, 0], [0, 0, 1, 0]])
    for i in range(N - 1):
        Gsubx[i * m:(i + 1) * m, i * n:(i + 1) * n] = xportion
        Gsubu[i * m:(i + 1) * m, (i + 1) * m:(i + 2) * m] = np.eye(2)
    r = m * (N - 1)
    G[0:r, :n * N] = Gsubx; G[0:r, n * N:n * N + m * N] = -Gsubu
    G[r:2 * r, :n * N] = -Gsubx; G[r:2 * r, n * N:n * N + m * N] = Gsubu
    o = 2 * r
    G[o:o + n * N, :n * N] = np.eye(n * N); o += n * N
    G[o:o + n * N, :n * N] = -np.eye(n * N); o += n * N
    G[o:o + m * N, n * N:n * N + m * N] = np.eye(m * N); o += m * N
    G[o:o + m * N, n * N:n * N + m * N] = -np.eye(m * N); o += m * N
    G[o:o + N * c, n * N + m * N:n * N + m * N + N * c] = -np.eye(N * c); o += N * c
    G[o:o + N * c, n * N + m * N:n * N + m * N + N * c] = np.eye(N * c)
    h[:2 * r] = FRIC
    h[2 * r:2 * r + 2 * n * N + 2 * m * N] = BOX
    h[2 * r + 2 * n * N + 2 * m * N:2 * r + 2 * n * N + 2 * m * N + N * c] = 0.0
    return Q, f, Aeq, beq, G, h


def closest_obstacle(pos, obs_xy):
    """MD:371-396: first strict minimum of the Euclidean distance; obs_xy is (2, K)."""
    min_dist, min_i = 1000.0, 0
    for i in range(obs_xy.shape[1]):
        d = np.sqrt((pos[0] - obs_xy[0, i]) ** 2 + (pos[1] - obs_xy[1, i]) ** 2)
        if d < min_dist:
            min_dist, min_i = d, i
    return obs_xy[:, min_i].copy()


# ----------------------------------------------------------------------------------------------
# QP stage: iSWIFT's Mehrotra predictor-corrector, dense restatement (Prime.c:127-230)
# ----------------------------------------------------------------------------------------------
def mehrotra_qp(P, c, A, b, G, h, reltol=1e-6, abstol=1e-6, maxit=25, sigma_d=0.0, trace=None):
    n, m, p = P.shape[0], G.shape[0], A.shape[0]

    def kkt_solve(w, r1, r2, r3):
        # [P A' G'; A 0 0; G 0 -W] [dx;dy;dz] = [r1;r2;r3]
        H = P + G.T @ (G / w[:, None])
        K = np.block([[H, A.T], [A, np.zeros((p, p))]])
        rhs = np.concatenate([r1 + G.T @ (r3 / w), r2])
        sol = np.linalg.solve(K, rhs)
        dx, dy = sol[:n], sol[n:]
        dz = (G @ dx - r3) / w
        return dx, dy, dz

    x, y, _ = kkt_solve(np.ones(m), -c, b, h)
    r = h - G @ x
    ap = -r.min()
    s = r.copy() if ap < 0 else r + (1 + ap)
    ad = r.max()
    z = -r if ad < 0 else -r + (1 + ad)
    sigma = 100.0
    flag = 2
    it = 0
    w_fact = None
    for it in range(maxit):
        rx = -(P @ x + G.T @ z + A.T @ y + c)
        ry = b - A @ x
        rz = h - s - G @ x
        if trace is not None:
            trace.append((np.linalg.norm(rx), np.linalg.norm(ry), np.linalg.norm(rz), s @ z / m))
        if (np.linalg.norm(rx) < reltol / np.sqrt(3) and np.linalg.norm(rz) < reltol / np.sqrt(3)
                and np.linalg.norm(ry) < reltol / np.sqrt(3) and s @ z / m < abstol):
            flag = 0
            break
        lam2 = s * z
        mu = lam2.sum() / m
        if sigma > sigma_d:
            w_fact = s / z
            ds = -lam2
            dx, dy, dz = kkt_solve(w_fact, rx, ry, rz - ds / z)
            dsl = (ds - s * dz) / z
            a_p = min([(-s[i] / dsl[i]) for i in range(m) if dsl[i] < 0], default=1.0)
            a_d = min([(-z[i] / dz[i]) for i in range(m) if dz[i] < 0], default=1.0)
            rho = ((s + a_p * dsl) @ (z + a_d * dz)) / (s @ z)
            sigma = max(sigma_d, min(1.0, rho) ** 3)
            ds = -lam2 - dsl * dz + sigma * mu
        else:
            sigma = sigma_d
            ds = -lam2 + sigma * mu
        dx, dy, dz = kkt_solve(w_fact, rx, ry, rz - ds / z)
        dsl = (ds - s * dz) / z
        a_p = min([(-s[i] / dsl[i]) for i in range(m) if dsl[i] < 0], default=1.0)
        a_d = min([(-z[i] / dz[i]) for i in range(m) if dz[i] < 0], default=1.0)
        a_p = min(0.99 * a_p, 1.0); a_d = min(0.99 * a_d, 1.0)
        x = x + a_p * dx; s = s + a_p * dsl
        y = y + a_d * dy; z = z + a_d * dz
    else:
        it = maxit
    return x, y, z, s, it, flag


# ----------------------------------------------------------------------------------------------
# NLP stage
# ----------------------------------------------------------------------------------------------
VEL_IDX = [1, 5, 9, 13, 3, 7, 11, 15]  # DV:270-279


class NlpProblem:
    """min 1/2 x'Qx+f'x  s.t. Aeq x=beq, c(x)>=0 with c = [h-Gx ; d_k-eps ; vmax - v ; vmax + v]."""

    def __init__(self, Q, f, Aeq, beq, G, h, obs, with_nl=True):
        self.Q, self.f, self.A, self.b, self.G, self.h = Q, f, Aeq, beq, G, h
        self.obs = np.asarray(obs, float)
        self.n = Q.shape[0]
        self.mG = G.shape[0]
        self.with_nl = with_nl
        self.m = self.mG + (4 + 16 if with_nl else 0)
        Jv = np.zeros((8, self.n))
        for r, j in enumerate(VEL_IDX):
            Jv[r, j] = 1.0
        self.Jv = Jv

    def cons(self, x):
        parts = [self.h - self.G @ x]
        if self.with_nl:
            d = np.array([(x[4 * k] - self.obs[0]) ** 2 + (x[4 * k + 2] - self.obs[1]) ** 2 + x[-1] for k in range(4)])
            v = x[VEL_IDX]
            parts += [d - EPS_DIST, VMAX - v, VMAX + v]
        return np.concatenate(parts)

    def jac(self, x):
        parts = [-self.G]
        if self.with_nl:
            Jd = np.zeros((4, self.n))
            for k in range(4):
                Jd[k, 4 * k] = 2 * (x[4 * k] - self.obs[0])
                Jd[k, 4 * k + 2] = 2 * (x[4 * k + 2] - self.obs[1])
                Jd[k, -1] = 1.0
            parts += [Jd, -self.Jv, self.Jv]
        return np.vstack(parts)

    def hess_lag(self, z):
        W = self.Q.copy()
        if self.with_nl:
            for k in range(4):
                zk = z[self.mG + k]
                W[4 * k, 4 * k] -= 2 * zk
                W[4 * k + 2, 4 * k + 2] -= 2 * zk
        return W

    def obj(self, x):
        return 0.5 * x @ self.Q @ x + self.f @ x

    def general_rows(self):
        r = list(range(12))                       # friction rows  MD:289-292
        if self.with_nl:
            r += list(range(self.mG, self.mG + 4))  # distance rows  DV:262-265
        return np.array(r)

    def grad(self, x):
        return self.Q @ x + self.f


DEFAULT_OPTS = dict(tol=3e-9, mu_init=0.1, kappa_eps=10.0, kappa_mu=0.2, theta_mu=1.5, tau_min=0.99,
                    s_push=1e-2, max_iter=200, gamma_theta=1e-5, gamma_phi=1e-5, delta=1.0, s_theta=1.1,
                    s_phi=2.3, eta_phi=1e-4, kappa_sigma=1e10, alpha_min=1e-10, filter_cap=8,
                    dw_first=1e-4, dw_min=1e-20, kw_minus=1.0 / 3, kw_plus=8.0, kw_plus_first=100.0,
                    dw_max=1e40, rho_lambda=1e-6)


def _ldl_inertia(K):
    from scipy.linalg import ldl
    _, d, _ = ldl(K)
    npos = nneg = nzero = 0
    i = 0
    nK = K.shape[0]
    while i < nK:
        if i + 1 < nK and d[i + 1, i] != 0.0:
            ev = np.linalg.eigvalsh(d[i:i + 2, i:i + 2])
            for e in ev:
                if e > 0: npos += 1
                elif e < 0: nneg += 1
                else: nzero += 1
            i += 2
        else:
            e = d[i, i]
            if e > 0: npos += 1
            elif e < 0: nneg += 1
            else: nzero += 1
            i += 1
    return (npos, nneg, nzero)


def _inertia_ok(H, A):
    # reduced Hessian on null(A) positive definite <=> KKT inertia (n, p, 0)
    u, sv, vt = np.linalg.svd(A)
    Z = vt[len(sv):].T
    ev = np.linalg.eigvalsh(Z.T @ H @ Z)
    return ev.min() > 0


def interior_start(prob, x_start, push):
    """Strictly interior start w.r.t. every inequality row (IPM-F step 0).
    Order: box/velocity on X, then U against friction+box, then Lambda, then the slack variable."""
    x = x_start.copy()
    n = prob.n
    c = (n - 25) // 4
    for j in range(16):
        lim = BOX
        if prob.with_nl and j in VEL_IDX:
            lim = min(lim, VMAX)
        x[j] = min(max(x[j], -lim + push), lim - push)
    for k in range(4):
        for a in range(2):
            lo, hi = -BOX + push, BOX - push
            if k >= 1:  # friction rows couple (p_k, u_k), k = 1..3   MD:280-283
                pk = x[4 * (k - 1) + 2 * a]
                lo = max(lo, pk - FRIC + push); hi = min(hi, pk + FRIC - push)
            j = 16 + 2 * k + a
            x[j] = min(max(x[j], lo), hi)
    for j in range(24, 24 + 4 * c):
        x[j] = min(max(x[j], push), 1 - push)
    if prob.with_nl:
        cd = prob.cons(x)[prob.mG:prob.mG + 4]
        if cd.min() < push:
            x[-1] += push - cd.min()
    return x


def nlp_ipm(prob, x_start, opts=None, verbose=False, direction_fn=None, compare=None):
    o = dict(DEFAULT_OPTS)
    if opts:
        o.update(opts)
    n, m, A, b = prob.n, prob.m, prob.A, prob.b
    p = A.shape[0]
    x = interior_start(prob, x_start, o['s_push'])
    s = prob.cons(x)
    assert s.min() > 0
    mu = o['mu_init']
    z = mu / s
    y = np.zeros(p)
    tau = max(o['tau_min'], 1 - mu)
    filt = []
    dw_last = 0.0
    first_ic = True
    info = dict(status=1, iters=0, n_fact=0, n_ls=0)

    def phi_of(x_, s_, mu_):
        return prob.obj(x_) - mu_ * np.log(s_).sum()

    theta0 = np.abs(A @ x - b).sum()
    theta_max = 1e4 * max(1.0, theta0)
    theta_min = 1e-4 * max(1.0, theta0)
    for it in range(o['max_iter'] + 1):
        s = prob.cons(x)
        J = prob.jac(x)
        g = prob.grad(x)
        r_d = g + A.T @ y - J.T @ z
        r_e = A @ x - b
        e_d = np.abs(r_d).max(); e_p = np.abs(r_e).max()
        E0 = max(e_d, e_p, np.abs(s * z).max())
        info.update(iters=it, kkt=E0, mu=mu)
        if verbose:
            print(f"it {it:3d} mu {mu:.2e} E0 {E0:.3e} ed {e_d:.2e} ep {e_p:.2e} obj {prob.obj(x):.8f} dw {dw_last:.1e} slack {x[-1]:.5f}")
        if E0 <= o['tol'] or (mu <= o['tol'] / 10 and E0 <= o.get('acc_tol', 9e-9)):
            info['status'] = 0
            break
        if it == o['max_iter']:
            break
        while True:
            Emu = max(e_d, e_p, np.abs(s * z - mu).max())
            if Emu <= o['kappa_eps'] * mu and mu > o['tol'] / 10:
                mu = max(o['tol'] / 10, min(o['kappa_mu'] * mu, mu ** o['theta_mu']))
                tau = max(o['tau_min'], 1 - mu)
                filt = []
            else:
                break
        Sig = z / s
        W = prob.hess_lag(z)
        # rows that are bounds on a single variable are condensed (axis aligned, no cancellation);
        # friction and distance rows stay in the augmented system  [H A' Jg'; A 0 0; Jg 0 -S/Z]
        gen = prob.general_rows()
        bnd = np.setdiff1d(np.arange(m), gen)
        Jb, Jg = J[bnd], J[gen]
        Hc = W + Jb.T @ (Sig[bnd, None] * Jb)
        Hc[24:n - 1, 24:n - 1] += o['rho_lambda'] * np.eye(n - 25)
        ng = len(gen)
        rhs = np.concatenate([-(g + A.T @ y - J.T @ z) - Jb.T @ (z[bnd] - mu / s[bnd]), -r_e, mu / z[gen] - s[gen]])
        dw = 0.0
        while True:
            if direction_fn is not None:
                info['n_fact'] += 1
                out = direction_fn(x, y, z, s, mu, dw)
                if out is not None:
                    break
            else:
              H = Hc + dw * np.eye(n)
              K = np.block([[H, A.T, Jg.T], [A, np.zeros((p, p + ng))], [Jg, np.zeros((ng, p)), -np.diag(s[gen] / z[gen])]])
              info['n_fact'] += 1
              if _ldl_inertia(K) == (n, p + ng, 0):
                break
            if dw == 0.0:
                dw = o['dw_first'] if dw_last == 0.0 else max(o['dw_min'], o['kw_minus'] * dw_last)
            else:
                dw *= o['kw_plus_first'] if first_ic and dw_last == 0.0 else o['kw_plus']
            if dw > o['dw_max']:
                info['status'] = 3
                return x, y, z, s, info
        if dw > 0:
            dw_last = dw
            first_ic = False
        if direction_fn is not None:
            dx, dy = out[0], out[1]
            if compare is not None:
                H = Hc + dw * np.eye(n)
                K = np.block([[H, A.T, Jg.T], [A, np.zeros((p, p + ng))], [Jg, np.zeros((ng, p)), -np.diag(s[gen] / z[gen])]])
                sol = np.linalg.solve(K, rhs)
                compare.append((np.abs(sol[:n] - dx).max() / max(np.abs(sol[:n]).max(), 1e-300), np.abs(sol[n:n+p] - dy).max() / max(np.abs(sol[n:n+p]).max(), 1e-300), mu))
        else:
            sol = np.linalg.solve(K, rhs)
            dx, dy = sol[:n], sol[n:n + p]
            out = (dx, dy, None)
            dz_gen = -sol[n + p:]
        ds = J @ dx
        dz = mu / s - z - Sig * ds
        if direction_fn is not None and len(out) > 2 and out[2] is not None and out[2].shape == dz.shape:
            dz = out[2]
        elif direction_fn is None:
            dz[gen] = dz_gen
        neg = ds < 0
        a_pmax = min(1.0, (-tau * s[neg] / ds[neg]).min()) if neg.any() else 1.0
        neg = dz < 0
        a_dmax = min(1.0, (-tau * z[neg] / dz[neg]).min()) if neg.any() else 1.0
        th = np.abs(r_e).sum()
        ph = phi_of(x, s, mu)
        dphi = g @ dx - mu * (ds / s).sum()
        alpha = a_pmax
        accepted = False
        ftype = False
        while alpha >= o['alpha_min']:
            info['n_ls'] += 1
            xt = x + alpha * dx
            st = prob.cons(xt)
            if st.min() > 0:
                tht = (1 - alpha) * th
                pht = phi_of(xt, st, mu)
                ok_filter = tht <= theta_max and all(not (tht >= ft and pht >= fp) for ft, fp in filt)
                if ok_filter:
                    switching = dphi < 0 and alpha * (-dphi) ** o['s_phi'] > o['delta'] * th ** o['s_theta']
                    if th <= theta_min and switching:
                        if pht - ph - 10 * EPSM * abs(ph) <= o['eta_phi'] * alpha * dphi:
                            accepted = True; ftype = True
                    else:
                        if tht <= (1 - o['gamma_theta']) * th or pht - ph - 10 * EPSM * abs(ph) <= -o['gamma_phi'] * th:
                            accepted = True; ftype = False
            if accepted:
                break
            alpha *= 0.5
        if not accepted:
            info['ls_fail'] = info.get('ls_fail', 0) + 1
            info['status'] = 2
            break
        if not ftype:
            if len(filt) >= o['filter_cap']:
                filt.pop(0)
            filt.append(((1 - o['gamma_theta']) * th, ph - o['gamma_phi'] * th))
        x = x + alpha * dx; y = y + alpha * dy
        z = z + a_dmax * dz
        s = prob.cons(x)
        lo = mu / (o['kappa_sigma'] * s); hi = o['kappa_sigma'] * mu / s
        z = np.minimum(np.maximum(z, lo), hi)
    return x, y, z, s, info


# ----------------------------------------------------------------------------------------------
def bernstein4():
    from math import comb
    M = np.zeros((5, 5))
    for i in range(5):
        t = i * 0.25
        for j in range(5):
            M[i, j] = comb(4, j) * t ** j * (1 - t) ** (4 - j)
    return M


def bezier_fit(alpha_buf, X16):
    """MD:784-855 effective semantics: alpha(4x5) = Y @ inv(M)' with Y=[buf,x1..x4]."""
    Y = np.column_stack([alpha_buf, X16.reshape(4, 4).T])  # 4 x 5 (rows = state comps)
    Minv = np.linalg.inv(bernstein4())
    return Y @ Minv.T, X16.reshape(4, 4)[3].copy()


NOMINAL_FEET = np.array([[0.2188, 0.2188, -0.1472, -0.1472], [-0.1320, 0.1320, -0.1320, 0.1320]])  # MD:1206-1209


def cfg1_instance(c=4):
    x0 = np.array([0.0, 0.05, 0.0, 0.0])
    xdes = np.zeros(16)
    for k in range(4):
        xdes[4 * k:4 * k + 4] = [0.1 * (k + 1) * TS, 0.1, 0.0, 0.0]
    contact = [1, 1, 1, 1] if c == 4 else [0, 1, 1, 0]
    return x0, xdes, NOMINAL_FEET.copy(), contact, np.array([[1.0], [0.2]])


def cfg3_instances(B, seed=20251018 + 3000):
    """SURVEY 8(d) cfg 3 generator (seeded PCG64, one spawned stream per instance)."""
    ss = np.random.SeedSequence(seed)
    out = []
    goal = np.array([10.0, 0.0])
    for child in ss.spawn(B):
        rng = np.random.Generator(np.random.PCG64(child))
        p0 = np.array([rng.uniform(0, 9), rng.uniform(-2, 2)])
        v0 = rng.uniform(-0.3, 0.3, 2)
        x0 = np.array([p0[0], v0[0], p0[1], v0[1]])
        dirg = (goal - p0) / np.linalg.norm(goal - p0)
        vref = 0.2 * dirg
        xdes = np.zeros(16)
        for k in range(4):
            pk = p0 + vref * (k + 1) * TS
            xdes[4 * k:4 * k + 4] = [pk[0], vref[0], pk[1], vref[1]]
        u = rng.uniform()
        contact = [1, 1, 1, 1] if u < 0.5 else ([0, 1, 1, 0] if u < 0.75 else [1, 0, 0, 1])
        nominal = np.array([[0.183, 0.183, -0.183, -0.183], [-0.132, 0.132, -0.132, 0.132]])
        feet = p0[:, None] + nominal + rng.uniform(-0.03, 0.03, (2, 4))
        ox = rng.uniform(0, 9, 20) + rng.uniform(-0.6, 0.6, 20)
        oy = rng.uniform(-2, 2, 20) + rng.uniform(-0.6, 0.6, 20)
        obs = np.vstack([np.clip(ox, 1, 9), np.clip(oy, -3, 3)])
        out.append((x0, xdes, feet, contact, obs))
    return out


def solve_instance(x0, xdes, feet, contact, obs, use_nlp=True, opts=None, verbose=False, qp_tol=1e-6):
    F = stance_matrix(feet, contact)
    Q, f, Aeq, beq, G, h = pack_dense(x0, xdes, F)
    xq, yq, zq, sq, itq, flagq = mehrotra_qp(Q, f, Aeq, beq, G, h, reltol=qp_tol, abstol=qp_tol)
    res = dict(x_qp=xq, qp_iters=itq, qp_flag=flagq)
    if use_nlp:
        o = closest_obstacle(np.array([x0[0], x0[2]]), obs)
        prob = NlpProblem(Q, f, Aeq, beq, G, h, o)
        x, y, z, s, info = nlp_ipm(prob, xq, opts, verbose)
        res.update(x=x, y=y, z=z, s=s, info=info, obs=o, prob=prob)
    return res


if __name__ == "__main__":
    import sys, time
    r = solve_instance(*cfg1_instance(4), verbose=True)
    print("qp iters", r['qp_iters'], r['qp_flag'], "x1", r['x_qp'][:4])
    print(r['info'], "obj", r['prob'].obj(r['x']), "slack", r['x'][-1])
    print("dist rows", r['prob'].cons(r['x'])[r['prob'].mG:r['prob'].mG + 4] + EPS_DIST)
    r = solve_instance(*cfg1_instance(2), verbose=False)
    print("c=2:", r['qp_iters'], r['info'], r['prob'].obj(r['x']))
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    t = time.time()
    its, fails, nf, nls = [], 0, [], []
    for k, inst in enumerate(cfg3_instances(B)):
        r = solve_instance(*inst)
        its.append(r['info']['iters']); nf.append(r['info']['n_fact']); nls.append(r['info']['n_ls'])
        if r['info']['status'] != 0 or r['info'].get('ls_fail'):
            fails += 1
            print("FAIL", k, r['info'])
    print("B", B, "iters mean/max", np.mean(its), np.max(its), "fact", np.mean(nf), "ls", np.mean(nls), "fails", fails, "time", time.time() - t)
