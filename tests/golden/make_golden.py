"""Generates the committed golden fixtures.  Run HERE (the build container, where /root/reference exists):

    python tests/golden/make_golden.py

 * iswift_small.npz   -- the reference's canned QP (optimization/iSWIFT/include/Matrices_small.h, parsed in place) solved by
                         the reference's own C core (oracle/_ref/libiswift_ref.so, compiled from /root/reference): inputs, the
                         reference's permutation, solution, iteration count.
 * mpc_qp_golden.npz  -- restated MPC QPs (src/MPC_dist.cpp:164-321) solved by the reference's own iSWIFT core.
 * nlp_golden.npz     -- NLP-stage solutions: the C oracle and, independently, scipy SLSQP from the same QP start.
The GPU box has no /root/reference; tests there read only these files.
"""
import ctypes as C
import os
import re
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
import proto_numpy as P  # noqa: E402

REF_H = "/root/reference/optimization/iSWIFT/include/Matrices_small.h"


def parse_header(path):
    txt = open(path).read()
    out = {}
    for m in re.finditer(r"(realqp|idxint)\s+(\w+)\s*\[\d+\]\s*=\s*\{([^}]*)\}", txt):
        typ, name, body = m.groups()
        vals = [v for v in re.split(r"[,\s]+", body.strip()) if v]
        out[name] = np.array(vals, dtype=np.float64 if typ == "realqp" else np.int32)
    for m in re.finditer(r"(realqp|idxint)\s+(\w+)\s*=\s*([-\d.eE]+)\s*;", txt):
        out[m.group(2)] = float(m.group(3)) if m.group(1) == "realqp" else int(m.group(3))
    return out


def ccs_to_dense(jc, ir, pr, rows, cols):
    M = np.zeros((rows, cols))
    for j in range(cols):
        for k in range(jc[j], jc[j + 1]):
            M[ir[k], j] += pr[k]
    return M


def main():
    ref = O.ref_lib()
    assert ref is not None, "oracle/_ref/libiswift_ref.so missing: run make -C oracle ref"
    # ---- 1. canned QP through the reference core, with the reference's own permutation
    d = parse_header(REF_H)
    n, m, p = d["n"], d["m"], d["p"]
    x = np.zeros(n); y = np.zeros(p); z = np.zeros(m); s = np.zeros(m); it = C.c_int(0); ts = C.c_double(0); tv = C.c_double(0)
    cp = lambda a: np.ascontiguousarray(a.copy())
    flag = ref.iswift_ref_solve_ccs(n, m, p, cp(d["Pjc"]), cp(d["Pir"]), cp(d["Ppr"]), cp(d["Ajc"]), cp(d["Air"]), cp(d["Apr"]),
                                    cp(d["Gjc"]), cp(d["Gir"]), cp(d["Gpr"]), cp(d["c"]), cp(d["h"]), cp(d["b"]), float(d["sigma_d"]),
                                    cp(d["P"]), x, y, z, s, C.byref(it), C.byref(ts), C.byref(tv))
    Pm = ccs_to_dense(d["Pjc"], d["Pir"], d["Ppr"], n, n); Am = ccs_to_dense(d["Ajc"], d["Air"], d["Apr"], p, n)
    Gm = ccs_to_dense(d["Gjc"], d["Gir"], d["Gpr"], m, n)
    print("Matrices_small: flag", flag, "iterations", it.value, "x[:4]", x[:4])
    np.savez(os.path.join(HERE, "iswift_small.npz"), P=Pm, A=Am, G=Gm, c=d["c"], h=d["h"], b=d["b"], perm=d["P"], x=x, y=y, z=z, s=s,
             iters=it.value, flag=flag)
    # ---- 2./3. MPC instances
    from importlib import import_module
    S = import_module("srb_cbf_nmpc_b200.scenarios")
    insts = [S.cfg1_instance(4), S.cfg1_instance(2)] + S.cfg3_instances(8)
    recs = dict(x0=[], xdes=[], feet=[], contact=[], obs=[], xqp_ref=[], qp_iters_ref=[], x_oracle=[], status=[], kkt=[], x_scipy=[],
                obj_oracle=[], obj_scipy=[], nv=[])
    from scipy.optimize import minimize
    for (x0, xdes, feet, contact, obs) in insts:
        F = P.stance_matrix(feet, contact)
        Q, f, A, b, G, h = O.pack_dense(x0, xdes, F)
        xr, yr, zr, sr, itr, fl, _, _ = O.iswift_ref_dense(Q, f, A, b, G, h)
        obs20 = np.hstack([obs, np.repeat(obs[:, -1:], 20 - obs.shape[1], 1)]) if obs.shape[1] < 20 else obs
        r = O.solve_instance(x0, xdes, feet, contact, obs20)
        o = r["obs"]
        prob = P.NlpProblem(Q, f, A, b, G, h, o)
        xs = None; fs = np.nan
        if r["status"] == 0:
            cons = [dict(type="eq", fun=lambda v: A @ v - b, jac=lambda v: A), dict(type="ineq", fun=prob.cons, jac=prob.jac)]
            res = minimize(prob.obj, r["x_qp"], jac=prob.grad, constraints=cons, method="SLSQP", options=dict(ftol=1e-15, maxiter=500))
            xs, fs = res.x, res.fun
        nv = Q.shape[0]
        pad = lambda v: np.concatenate([v, np.full(41 - len(v), np.nan)]) if v is not None else np.full(41, np.nan)
        recs["x0"].append(x0); recs["xdes"].append(xdes); recs["feet"].append(feet); recs["contact"].append(contact); recs["obs"].append(obs20)
        recs["xqp_ref"].append(pad(xr)); recs["qp_iters_ref"].append(itr); recs["x_oracle"].append(pad(r["x"])); recs["status"].append(r["status"])
        recs["kkt"].append(r["kkt"]); recs["x_scipy"].append(pad(xs)); recs["obj_oracle"].append(prob.obj(r["x"])); recs["obj_scipy"].append(fs); recs["nv"].append(nv)
        if xs is not None:
            print("nv", nv, "status", r["status"], "obj oracle %.8f scipy %.8f  |dXU| %.2e |dslack| %.2e" % (prob.obj(r["x"]), fs, np.abs(xs[:24] - r["x"][:24]).max(), abs(xs[-1] - r["x"][-1])))
        else:
            print("nv", nv, "status", r["status"], "(no scipy run)")
    np.savez(os.path.join(HERE, "mpc_golden.npz"), **{k: np.array(v) for k, v in recs.items()})
    # ---- 4. dense QPs of the low-level controller's size (src/LowLevelCtrl.cpp:18-52: 31 variables, 18 equalities, 45 inequalities
    #         at 4 stance feet) through the reference's own iSWIFT core
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from qp_cases import lowlevel_like_qps
    qps = lowlevel_like_qps(12, seed=20251018 + 7000)
    sol = dict(x=[], iters=[], flag=[])
    for (Pq, cq, Aq, bq, Gq, hq) in qps:
        xr, yr, zr, sr, itr, fl, _, _ = O.iswift_ref_dense(Pq, cq, Aq, bq, Gq, hq)
        sol["x"].append(xr); sol["iters"].append(itr); sol["flag"].append(fl)
    print("low-level-size QPs through the reference core: flags", sol["flag"], "iterations", sol["iters"])
    np.savez(os.path.join(HERE, "lowlevel_qp_golden.npz"), x=np.array(sol["x"]), iters=np.array(sol["iters"]), flag=np.array(sol["flag"]))


if __name__ == "__main__":
    main()
