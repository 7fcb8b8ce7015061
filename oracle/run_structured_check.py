"""Dev check: run the IPM with the structured (Riccati+border) direction and compare with the dense one."""
import sys, numpy as np
sys.path.insert(0, __file__.rsplit('/', 1)[0])
import proto_numpy as P
import proto_structured as S

def solve_structured(inst, compare=None, verbose=False, opts=None):
    x0, xdes, feet, contact, obs = inst
    F = P.stance_matrix(feet, contact)
    Q, f, Aeq, beq, G, h = P.pack_dense(x0, xdes, F)
    xq, *_ = P.mehrotra_qp(Q, f, Aeq, beq, G, h)
    o = P.closest_obstacle(np.array([x0[0], x0[2]]), obs)
    prob = P.NlpProblem(Q, f, Aeq, beq, G, h, o)
    st = S.Stagewise(x0, xdes, F, o)
    oo = dict(P.DEFAULT_OPTS); oo.update(opts or {})
    def dirfn(x, y, z, s, mu, dw):
        return S.structured_direction(st, x, y, z, s, mu, dw, oo['rho_lambda'])
    return P.nlp_ipm(prob, xq, opts, verbose, direction_fn=dirfn, compare=compare), prob

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    insts = P.cfg3_instances(B)
    worst = 0; fails = 0; its = []; nf = []
    for k, inst in enumerate(insts):
        cmp_ = []
        (x, y, z, s, info), prob = solve_structured(inst, cmp_)
        rd = P.solve_instance(*inst)
        its.append(info['iters']); nf.append(info['n_fact'])
        if cmp_:
            w = max(c[0] for c in cmp_); worst = max(worst, w)
        if info['status'] != rd['info']['status']:
            print("STATUS MISMATCH", k, info, rd['info'])
        if info['status'] == 0:
            dxs = np.abs(x[:24] - rd['x'][:24]).max(); dsl = abs(x[-1] - rd['x'][-1])
            if dxs > 1e-7 or dsl > 1e-7:
                print("SOLUTION MISMATCH", k, dxs, dsl, info['iters'], rd['info']['iters'])
        else:
            fails += 1
    print("B", B, "iters", np.mean(its), max(its), "facts", np.mean(nf), "fails", fails, "worst rel dir err", worst)
