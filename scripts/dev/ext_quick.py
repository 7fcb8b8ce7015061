"""Dev tool: extended-mode kernel timings and iteration statistics on one GPU.
    python scripts/dev/ext_quick.py [what ...]      what in: faithful cfg4 cfg5 cfg5single"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import srb_cbf_nmpc_b200 as pkg
from srb_cbf_nmpc_b200 import capi, scenarios as S

dev = torch.device("cuda", 0)
what = sys.argv[1:] or ["faithful", "cfg4", "cfg5"]


def run(name, insts, N, T, reps=5, **kw):
    B = len(insts)
    d = capi.ext_soa_from_instances(insts, N)
    s = pkg.BatchSolver(device=0, max_batch=max(B, 64), max_obs=20)
    tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
    tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(N).items()}
    tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
    ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
    ms = []
    for _ in range(reps):
        s.ext_solve_device(B, ip, op, d["n_obs"], N=N, team_size=T, **kw)
        torch.cuda.synchronize()
        ms.append(s.last_kernel_ms())
    st = tout["status"][0].cpu().numpy(); ti = tout["team_iters"][0].cpu().numpy()[::T]; ni = tout["nlp_iters"][0].cpu().numpy(); qi = tout["qp_iters"][0].cpu().numpy()
    conv = (st == 0).sum()
    print("%-12s B %6d N %d T %d  kernel ms min %.3f med %.3f  converged %.4f  -> %.3f M converged solves/s | status %s | qp it mean %.1f max %d | nlp it mean %.1f max %d | team it mean %.1f p90 %d p99 %d max %d"
          % (name, B, N, T, min(ms), sorted(ms)[len(ms) // 2], conv / B, conv / (min(ms) * 1e-3) / 1e6, np.bincount(st, minlength=5), qi.mean(), qi.max(),
             ni[st == 0].mean(), ni.max(), ti.mean(), np.percentile(ti, 90), np.percentile(ti, 99), ti.max()))
    s.close()
    return tout


if "faithful" in what:
    insts = S.cfg3_instances(4096)
    f = pkg.BatchSolver(device=0, max_batch=4096, max_obs=20)
    fo = f.solve_host(capi.soa_from_instances(insts), mode=capi.MODE_QP_NLP)
    fo = f.solve_host(capi.soa_from_instances(insts), mode=capi.MODE_QP_NLP)
    print("faithful kernel (cbfnmpc_lane_kernel) 4096: %.3f ms" % f.last_kernel_ms())
    # closest obstacle per instance (first strict minimum)
    ext = []
    for (x0, xdes, feet, contact, obs) in insts:
        dd = np.hypot(x0[0] - obs[0], x0[2] - obs[1]); k = int(np.argmin(dd))
        ext.append((x0, xdes, np.asarray(feet)[None], contact, obs[:, k:k + 1].copy()))
    o = run("ext faithful", ext, 4, 1)
    ok = (fo["status"][0] == 0) & (o["status"][0].cpu().numpy() == 0)
    print("   |X ext - X faithful| max", np.abs(o["X"].cpu().numpy() - fo["X"])[:, ok].max(), "status equal", (fo["status"][0] == o["status"][0].cpu().numpy()).mean())
    for B in (8192, 32768):
        insts = S.cfg3_instances(B, seed=7)
        ext = []
        for (x0, xdes, feet, contact, obs) in insts:
            dd = np.hypot(x0[0] - obs[0], x0[2] - obs[1]); k = int(np.argmin(dd))
            ext.append((x0, xdes, np.asarray(feet)[None], contact, obs[:, k:k + 1].copy()))
        run("ext faithful", ext, 4, 1, reps=3)
if "cfg4" in what:
    run("cfg4 teams", S.cfg4_teams(2048), 4, 4, eps_nbr=S.EPS_NBR)
if "cfg5" in what:
    run("cfg5 teams", S.cfg5_instances(8192), 8, 4, reps=3, eps_nbr=S.EPS_NBR)
if "cfg5single" in what:
    run("cfg5 single", S.cfg5_instances(8192, team=False), 8, 1, reps=3)
