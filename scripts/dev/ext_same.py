import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
import srb_cbf_nmpc_b200 as pkg
from srb_cbf_nmpc_b200 import capi, scenarios as S
dev = torch.device("cuda", 0)
def run(name, insts, N, T, reps=4, **kw):
    B = len(insts)
    d = capi.ext_soa_from_instances(insts, N)
    s = pkg.BatchSolver(device=0, max_batch=max(B, 64), max_obs=20)
    tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
    tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(N).items()}
    tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
    ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
    ms = []
    for _ in range(reps):
        s.ext_solve_device(B, ip, op, d["n_obs"], N=N, team_size=T, **kw); torch.cuda.synchronize(); ms.append(s.last_kernel_ms())
    ti = tout["team_iters"][0].cpu().numpy()[::T]; qi = tout["qp_iters"][0].cpu().numpy()
    print("%-28s kernel ms %.3f | team it mean %.1f max %d | qp it max %d | ms per (max team it + max qp it) %.4f" % (name, min(ms), ti.mean(), ti.max(), qi.max(), min(ms) / (ti.max() + qi.max())))
    s.close()
teams = S.cfg4_teams(2048)
run("cfg4 2048 different teams", teams, 4, 4, eps_nbr=S.EPS_NBR)
for t in (0, 7, 100):
    same = teams[4 * t:4 * t + 4] * 2048
    run("cfg4 2048 x team %d" % t, same, 4, 4, eps_nbr=S.EPS_NBR)
run("cfg4 1 team", teams[:4], 4, 4, eps_nbr=S.EPS_NBR)
run("cfg4 64 teams", teams[:256], 4, 4, eps_nbr=S.EPS_NBR)
