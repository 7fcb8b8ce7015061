"""The structured Newton-step algebra the CUDA kernel uses (stage-wise Riccati recursion, support weights reduced by u = F lambda and
1'lambda = 1, bound rows condensed, capped distance rows + slack variable as a 5x5 border), as a numpy model
(oracle/proto_structured.py), against the dense augmented-system direction of oracle/proto_numpy.py along whole IPM-F runs."""
import numpy as np
import proto_numpy as P
import proto_structured as PS


def _run(inst, cap, refine):
    x0, xdes, feet, contact, obs = inst
    F = P.stance_matrix(feet, contact)
    Q, f, A, b, G, h = P.pack_dense(x0, xdes, F)
    xq = P.mehrotra_qp(Q, f, A, b, G, h)[0]
    o = P.closest_obstacle(np.array([x0[0], x0[2]]), obs)
    prob = P.NlpProblem(Q, f, A, b, G, h, o)
    st = PS.Stagewise(x0, xdes, F, o)
    cmp_ = []

    def dfn(x, y, z, s, mu, dw):
        out = PS.structured_direction(st, x, y, z, s, mu, dw, 1e-6, True, cap, refine)
        return None if out is None else out[:3]
    r = P.nlp_ipm(prob, xq, direction_fn=dfn, compare=cmp_)
    info = r[-1] if isinstance(r, tuple) else r
    return info, cmp_


def test_structured_direction_equals_dense_direction():
    """Both stance classes: the primal direction of every iteration agrees with the dense solve to 1e-5 relative, the equality
    multipliers' to 1e-3 (they are the last to settle: 4e-5 at the barrier floor, where the step itself is ~1e-9), with two
    refinement passes and the Sigma cap 1e8 -- the kernel's settings --, and the runs converge."""
    insts = P.cfg3_instances(16)
    seen = set()
    for inst in insts:
        info, cmp_ = _run(inst, PS.SIGMA_CAP, 2)
        if info["status"] != 0:
            continue
        seen.add(int(np.sum(inst[3])))
        assert cmp_ and max(c[0] for c in cmp_) < 1e-5 and max(c[1] for c in cmp_) < 1e-3
    assert seen == {2, 4}


def test_sigma_cap_is_what_keeps_the_recursion_accurate():
    """Without the cap (the distance rows condensed into the stage Hessians whatever their Sigma) the late directions lose most
    of their digits -- the reason for the capped split and the border."""
    worst_capped = worst_uncapped = 0.0
    for inst in P.cfg3_instances(10):
        info, c1 = _run(inst, PS.SIGMA_CAP, 2)
        info2, c2 = _run(inst, np.inf, 2)
        if info["status"] == 0 and c1: worst_capped = max(worst_capped, max(c[0] for c in c1))
        if c2: worst_uncapped = max(worst_uncapped, max(c[0] for c in c2))
    assert worst_capped < 1e-5 and worst_uncapped > 1e-3
