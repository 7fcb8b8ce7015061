"""Dev tool: effect of the NLP tolerance (and with it the final barrier parameter tol/10) on iteration tails and kernel time."""
import sys, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
for B in (4096, 32768):
    d = capi.soa_from_instances(S.cfg3_instances(B))
    for tol in (1e-9, 3e-9, 9e-9):
        s = m.BatchSolver(max_batch=B, options=dict(tol=tol))
        for rep in range(3): out = s.solve_host(d, mode=1)
        st = out["status"][0]; it = out["nlp_iters"][0]
        print("B %5d tol %.0e kernel %.3f ms status %s nlp its mean %.2f max %d top5 %s kkt max %.2e nfact-max %d" % (B, tol, s.last_kernel_ms(), np.bincount(st, minlength=5), it[st == 0].mean(), it.max(), np.sort(it)[-5:], out["kkt"][0][st == 0].max(), out["n_fact"][0].max()))
        s.close()
