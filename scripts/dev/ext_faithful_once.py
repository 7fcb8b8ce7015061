"""Dev tool: the extended-mode kernel on the 4096 cfg-3 instances (faithful problem: closest obstacle as the one row); the command
the ncu capture of cbfnmpc_ext_kernel wraps."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import torch
import srb_cbf_nmpc_b200 as pkg
from srb_cbf_nmpc_b200 import capi, scenarios as S
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
insts = S.cfg3_instances(B)
ext = []
for (x0, xdes, feet, contact, obs) in insts:
    dd = np.hypot(x0[0] - obs[0], x0[2] - obs[1]); k = int(np.argmin(dd))
    ext.append((x0, xdes, np.asarray(feet)[None], contact, obs[:, k:k + 1].copy()))
d = capi.ext_soa_from_instances(ext, 4)
s = pkg.BatchSolver(device=0, max_batch=B, max_obs=20)
tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(4).items()}
tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
for _ in range(4):
    s.ext_solve_device(B, ip, op, d["n_obs"], N=4, team_size=1); torch.cuda.synchronize()
    print("ext kernel ms", s.last_kernel_ms())
