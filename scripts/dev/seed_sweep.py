"""Dev tool: tails over several seeds (status histogram, max iterations, kernel time) at 32768 instances per launch."""
import sys, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = 32768
s = m.BatchSolver(max_batch=B)
for seed in (1, 2, 3, 4, 5, 6):
    d = capi.soa_from_instances(S.cfg3_instances(B, seed=seed))
    for rep in range(2): out = s.solve_host(d, mode=1)
    st = out["status"][0]; it = out["nlp_iters"][0]; qp = out["qp_iters"][0]
    outq = s.solve_host(d, mode=0)
    print("seed %d kernel %.3f ms status %s nlp max %d top3 %s qp max %d | QP-only status %s" % (seed, s.last_kernel_ms(), np.bincount(st, minlength=5), it.max(), np.sort(it)[-3:], qp.max(), np.bincount(outq["status"][0], minlength=5)))
