"""Dev tool: solve selected cfg-3 instances on the GPU, together and one by one."""
import sys, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
idx = [int(a) for a in sys.argv[1:]]
allinst = S.cfg3_instances(4096)
sub = [allinst[i] for i in idx]
s = m.BatchSolver(max_batch=4096)
out = s.solve_host(capi.soa_from_instances(sub), mode=1)
print("together  status", out["status"][0], "nlp", out["nlp_iters"][0])
st = []; it = []
for i in idx:
    o = s.solve_host(capi.soa_from_instances([allinst[i]]), mode=1); st.append(int(o["status"][0, 0])); it.append(int(o["nlp_iters"][0, 0]))
print("one by one status", st, "nlp", it)
# neighbours in the original batch order (same warp in the full run): 8 consecutive of the same class
