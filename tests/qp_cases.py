"""Seeded dense QPs with the dimensions and the structure of the reference's low-level controller QP
(src/LowLevelCtrl.cpp:18-52, 156-163: 31 variables = 12 contact forces + 12 joint torques + 6 output accelerations + 1 CLF
slack; 18 equalities; 45 inequalities = 4 x 5 friction-pyramid rows + 2 x 12 torque bounds + 1 CLF row).  The data are
synthetic (the whole-body model that fills them in the reference is outside this repo's scope); feasibility is by construction."""
import numpy as np


def lowlevel_like_qps(count, seed=20251018 + 7000, n=31, p=18, m=45):
    rng = np.random.default_rng(seed)
    out = []
    mu_f = 0.6
    for _ in range(count):
        M = rng.normal(size=(n, n)) * 0.3
        P = M.T @ M + np.diag(rng.uniform(0.05, 2.0, n))               # strictly convex cost
        c = rng.normal(size=n) * 5.0
        x_feas = np.concatenate([np.tile([0.0, 0.0, 30.0], 4) + rng.normal(size=12), rng.uniform(-8, 8, 12), rng.normal(size=6), [0.5]])
        A = rng.normal(size=(p, n)); b = A @ x_feas
        G = np.zeros((m, n)); h = np.zeros(m)
        r = 0
        for leg in range(4):                                            # friction pyramid + unilateral force (LowLevelCtrl.cpp:156-163)
            fx, fy, fz = 3 * leg, 3 * leg + 1, 3 * leg + 2
            for row in ([1, 0, -mu_f], [-1, 0, -mu_f], [0, 1, -mu_f], [0, -1, -mu_f], [0, 0, -1]):
                G[r, fx], G[r, fy], G[r, fz] = row; h[r] = 0.0 if row[2] != -1 else -5.0; r += 1
        for j in range(12):                                             # torque saturation +-33.5
            G[r, 12 + j] = 1.0; h[r] = 33.5; r += 1
            G[r, 12 + j] = -1.0; h[r] = 33.5; r += 1
        G[r, 24:30] = rng.normal(size=6); G[r, 30] = -1.0; h[r] = G[r] @ x_feas + 1.0   # CLF row with slack
        assert (G @ x_feas < h).all()
        out.append((P, c, A, b, G, h))
    return out
