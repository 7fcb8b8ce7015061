"""Dev tool: per-team statuses / iteration counts of a config-4 (or config-5) batch on the GPU -> gpurun_out/ext_dump_<N>.npz"""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as pkg
from srb_cbf_nmpc_b200 import capi, scenarios as S
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
K = 1 if N == 4 else 16
insts = S.cfg4_teams(nt, N=N, K_obs=K, seed=20251018 + (4000 if N == 4 else 5000))
s = pkg.BatchSolver(device=0, max_batch=len(insts), max_obs=20)
out = s.ext_solve_host(capi.ext_soa_from_instances(insts, N), N=N, team_size=4, eps_nbr=S.EPS_NBR)
st = out["status"][0].reshape(-1, 4); ti = out["team_iters"][0][::4]
bad = np.nonzero((st == 1).any(1))[0]
print("teams with a member at the iteration limit:", bad, "their team_iters", ti[bad])
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/ext_dump_%d.npz" % N, status=out["status"][0], team_iters=out["team_iters"][0], nlp_iters=out["nlp_iters"][0], kkt=out["kkt"][0], X=out["X"], slack=out["slack"][0])
