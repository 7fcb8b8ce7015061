"""Dev tool: agent-sharded config-4 teams on ONE GPU (the four members of a team in different CTAs): per-iteration path against the
fused path (mailbox + sequence flags in the own memory).  usage: ext_fused_quick.py [n_teams]   (the command the ncu capture wraps)"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import srb_cbf_nmpc_b200 as pkg
from srb_cbf_nmpc_b200 import capi, scenarios as S
n_teams = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = 4; dev = torch.device("cuda", 0)
insts = S.cfg4_teams(n_teams, N=N, K_obs=1)
B = 4 * n_teams
s = pkg.BatchSolver(device=0, max_batch=B, max_obs=20)
mm = [insts[4 * t + m] for m in range(4) for t in range(n_teams)]
d = capi.ext_soa_from_instances(mm, N)
tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
for name, fn in (("per-iteration", s.ext_solve_sharded), ("fused", s.ext_solve_fused)):
    tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(N).items() if k != "pub"}
    tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
    for rep in range(3):
        st = fn(n_teams, {k: v.data_ptr() for k, v in tin.items()}, {k: v.data_ptr() for k, v in tout.items()}, d["n_obs"], N=N, team_size=4, eps_nbr=S.EPS_NBR)
        torch.cuda.synchronize()
    conv = int((tout["status"][0] == 0).sum().item())
    print("%-14s %d teams: %.3f ms, %d launches, %d iterations, fused %d, %.3f M converged solves/s" % (name, n_teams, st["total_ms"], st["launches"], st["iterations"], st["fused"], conv / st["total_ms"] / 1e3))
