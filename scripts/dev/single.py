"""Dev tool: one cfg-1 instance, a few launches (for ncu: the critical path of a single group)."""
import sys, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
x0, xdes, feet, contact, obs = S.cfg1_instance(4)
obs = np.hstack([obs, np.repeat(obs[:, -1:], 20 - obs.shape[1], 1)])
d = capi.soa_from_instances([(x0, xdes, feet, contact, obs)])
s = m.BatchSolver(max_batch=8)
for i in range(8):
    out = s.solve_host(d, mode=1)
print("status", out["status"][0], "its", out["qp_iters"][0], out["nlp_iters"][0], "kernel ms", s.last_kernel_ms())
