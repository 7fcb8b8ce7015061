"""Dev tool: run the cfg-3 batch through the C ABI on the GPU and dump per-instance status / iteration counts."""
import sys, numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as m
from srb_cbf_nmpc_b200 import capi, scenarios as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
tag = sys.argv[2] if len(sys.argv) > 2 else "lane"
d = capi.soa_from_instances(S.cfg3_instances(B))
s = m.BatchSolver(max_batch=B)
res = {}
for mode in (0, 1):
    out = s.solve_host(d, mode=mode)
    for k in ("status", "qp_iters", "nlp_iters", "n_fact", "kkt", "X", "U", "slack", "x_qp"):
        res["m%d_%s" % (mode, k)] = out[k]
np.savez_compressed("gpurun_out/status_%s_%d.npz" % (tag, B), **res)
print(tag, "mode1 status", np.bincount(res["m1_status"][0], minlength=5), "max nlp it", res["m1_nlp_iters"].max())
