"""Dev tool: GPU answers against the oracle over several seeds of the cfg-3 generator (status agreement, worst deviation of X, U,
slack on the instances both converge on, iteration tails)."""
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
import helpers as H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = m.BatchSolver(max_batch=B)
for seed in (11, 12, 13, 14, 15, 16):
    insts = S.cfg3_instances(B, seed=seed)
    out = s.solve_host(capi.soa_from_instances(insts), mode=1)
    ref = H.oracle_batch(insts, mode=1, nthreads=16)
    c = H.compare(out, ref)
    st = out["status"][0]
    print("seed %d: status equal %.5f (%d differ), worst rel. deviation %.2e, status hist %s, nlp max gpu %d oracle %d, iteration counts equal %.4f"
          % (seed, c["status_equal"].mean(), int((~c["status_equal"]).sum()), c["worst"], np.bincount(st, minlength=5), out["nlp_iters"][0].max(),
             ref["nlp_iters"].max(), (out["nlp_iters"][0] == ref["nlp_iters"]).mean()))
