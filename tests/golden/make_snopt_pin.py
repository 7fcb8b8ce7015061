"""Mints tests/golden/snopt_print_file.npz from the ONE NLP output the reference tree holds: /root/reference/print_file.out,
a SNOPT 7.7.7 print file of one nv = 41 instance of MPC_dist::run_NMPC (src/MPC_dist.cpp:402-427).  Run HERE (the build
container, where /root/reference exists):

    python tests/golden/make_snopt_pin.py

What the print file holds and what is reconstructed from it (nothing else about that run was recorded):
  * derivative-check block (print_file.out:25-49): the START point SNOPT was given = iSWIFT's QP optimum (MPC_dist.cpp:356-361, 403)
    -- x(1..24) and x(41) -- and the objective gradient there, g = Q x + f, both to 9 digits
        -> x_des_j = x_j - g_j / w_j  (w = 300 for j <= 12, 2000 for 13..16; MPC_dist.cpp:164-210)          [~1e-9]
        -> A_d x_0 = x_1 - B_d u_0    (the QP optimum satisfies the dynamics rows to iSWIFT's 1e-6)          [~1e-7]
           cross-check: right-hand sides of rows 2-5 (print_file.out:95-98, 5 digits) = -A_d x_0
  * stance feet: the nominal footprint about (0, 0) (MPC_dist.cpp:1206-1216; first two gait domains); cross-check: u = F lambda
    with the printed final u and lambda (5 digits)
  * final point (print_file.out:231-271, 5 digits), objective 803.70032441, slack x41 = 0.76247, distance rows
    1.90000 / 1.89969 / 1.90105 / 1.90000 with duals 702.09383 / 0 / 0 / 1583.39966 (print_file.out:215-218)
        -> the obstacle = least-squares fit of (o_x, o_y) to the four printed distance-row values                  [~1e-3]
           (SNOPT's own exit: "requested accuracy could not be achieved", max primal infeasibility 3.1e-4 on distance row 2,
            max dual infeasibility 1.6e3 -- so stationarity cannot be used to sharpen the obstacle, and SNOPT's point is a
            loosely converged one)
The fixture is a WEAK pin (5-digit data, inputs reconstructed), and the tests that use it say so.
"""
import os
import re
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
SRC = "/root/reference/print_file.out"


def parse(path):
    lines = open(path).read().splitlines()
    start_x, start_g = {}, {}
    rows, cols = {}, {}
    obj = None
    for ln in lines:
        m = re.match(r"^\s+(\d+)\s+(-?\d\.\d+E[+-]\d+)\s+\d\.\d+E[+-]\d+\s+(-?\d\.\d+E[+-]\d+)\s+(-?\d\.\d+E[+-]\d+)\s+(ok|bad\?)", ln)
        if m:
            start_x[int(m.group(1))] = float(m.group(2)); start_g[int(m.group(1))] = float(m.group(3))
        m = re.match(r"^\s+\d+\s+r\s+(\d+)\s+(?:[A-Z]\s+)?([A-Z]+)\s+(.*)$", ln)
        if m:
            f = m.group(3).split()
            num = lambda t: 0.0 if t == "." else (None if t == "None" else float(t))
            rows[int(m.group(1))] = dict(state=m.group(2), value=num(f[0]), lower=num(f[2]), upper=num(f[3]), dual=num(f[4]))
        m = re.match(r"^\s+\d+\s+x\s+(\d+)\s+(?:[A-Z]\s+)?([A-Z]+)\s+(.*)$", ln)
        if m:
            f = m.group(3).split()
            num = lambda t: 0.0 if t == "." else (None if t == "None" else float(t))
            cols[int(m.group(1))] = dict(value=num(f[0]), grad=num(f[1]))
        m = re.search(r"Objective Value\s+(-?\d\.\d+E[+-]\d+)", ln)
        if m:
            obj = float(m.group(1))
    return start_x, start_g, rows, cols, obj


def main():
    import oracle as O
    sx, sg, rows, cols, obj = parse(SRC)
    assert len(rows) == 133 and len(cols) == 41 and sorted(sx) == list(range(1, 25)) + [41]
    g = O.default_gains()
    Ad, Bd = O.lip_discrete(g)
    x_start = np.array([sx[j] for j in range(1, 25)]); g_start = np.array([sg[j] for j in range(1, 25)])
    w = np.array([g.Qw] * 12 + [g.Pw] * 4)
    xdes = x_start[:16] - g_start[:16] / w
    assert np.abs(g_start[16:24] - g.Rw * x_start[16:24]).max() < 1e-9          # R = 0.1 on the inputs
    Adx0 = x_start[:4] - Bd @ x_start[16:18]
    rhs = np.array([rows[r]["lower"] for r in (2, 3, 4, 5)])
    assert np.abs(-Adx0 - rhs).max() < 6e-6, (Adx0, rhs)                          # 5-digit cross-check
    x0 = np.linalg.solve(Ad, Adx0)
    F = np.array([[0.2188, 0.2188, -0.1472, -0.1472], [-0.1320, 0.1320, -0.1320, 0.1320]])   # MPC_dist.cpp:1206-1209 about (0, 0)
    xf = np.array([cols[j]["value"] for j in range(1, 42)]); gf = np.array([cols[j]["grad"] for j in range(1, 42)])
    for k in range(4):
        assert np.abs(F @ xf[24 + 4 * k:28 + 4 * k] - xf[16 + 2 * k:18 + 2 * k]).max() < 2e-5, k   # u = F lambda at the printed solution
    assert np.abs((xf[:16] - gf[:16] / w) - xdes).max() < 2e-5                    # the same x_des from the final point's gradient
    # rows: 1 objective | 2..29 equalities | 30..121 linear inequalities | 122..125 distance | 126..133 velocity
    dist = np.array([rows[r]["value"] for r in (122, 123, 124, 125)]); dist_dual = np.array([rows[r]["dual"] for r in (122, 123, 124, 125)])
    assert [rows[r]["lower"] for r in (122, 123, 124, 125)] == [1.9] * 4 and rows[126]["upper"] == 0.35 and abs(rows[30]["upper"] - 0.14354) < 1e-9
    delta = xf[40]
    P = np.stack([xf[0:16:4], xf[2:16:4]], 1)                                     # the 4 planned positions
    # least-squares obstacle: |p_k - o|^2 = dist_k - delta (Gauss-Newton from both intersection candidates of circles 1 and 4)
    r2 = dist - delta
    best = None
    for sign in (1.0, -1.0):
        mid = 0.5 * (P[0] + P[3]); t = P[3] - P[0]; n = np.array([-t[1], t[0]]) / np.linalg.norm(t)
        o = mid + sign * np.sqrt(max(r2[[0, 3]].mean() - 0.25 * t @ t, 0.0)) * n
        for _ in range(50):
            res = ((P - o) ** 2).sum(1) - r2
            J = -2 * (P - o)
            o = o - np.linalg.lstsq(J, res, rcond=None)[0]
        res = ((P - o) ** 2).sum(1) - r2
        if best is None or np.abs(res).max() < best[1]:
            best = (o.copy(), np.abs(res).max())
    obs, fit = best
    # The fit lands on an obstacle of the reference's own 4-agent scenario: (0.85, -0.65), third column of the obstacle matrix of
    # src/A1_Sim_4_agents.cpp:872-873.  The fixture keeps the fit and uses the scenario's exact value.
    obs_exact = np.array([0.85, -0.65])
    assert np.abs(obs - obs_exact).max() < 5e-4, obs
    print("x0", x0, "\nobstacle", obs, "fit residual", fit, "\ndelta", delta, "objective", obj)
    out = os.path.join(HERE, "snopt_print_file.npz")
    np.savez(out, x0=x0, xdes=xdes, feet=F, contact=np.array([1, 1, 1, 1], np.int32), obs=obs_exact, obs_fit=obs, obs_fit_residual=fit,
             x_start=x_start, g_start=g_start, x_final=xf, grad_final=gf, objective=obj, dist_rows=dist, dist_duals=dist_dual,
             eq_duals=np.array([rows[r]["dual"] for r in range(2, 30)]), vel_rows=np.array([rows[r]["value"] for r in range(126, 134)]),
             fric_rows=np.array([rows[r]["value"] for r in range(30, 42)]), source="/root/reference/print_file.out (SNOPT 7.7.7, Feb 2021)")
    print("wrote", out)


if __name__ == "__main__":
    main()
