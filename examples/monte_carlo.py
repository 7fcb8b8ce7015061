#!/usr/bin/env python
"""Closed-loop Monte-Carlo on one B200: the reference's "QP-only vs QP->NLP" success-rate study (src/A1_Sim.cpp:983-1266,
A1_Sim_Series.cpp), on the LIP template, thousands of worlds at once and entirely on the device (cbfnmpc_closed_loop).

    python examples/monte_carlo.py [worlds] [cycles]
"""
import os
import sys
import time
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srb_cbf_nmpc_b200 as pkg  # noqa: E402
from srb_cbf_nmpc_b200 import capi, scenarios as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dev = torch.device("cuda", 0)
for name, mode in (("QP only (use_snopt = false)", capi.MODE_QP), ("QP -> NLP (use_snopt = true)", capi.MODE_QP_NLP)):
    w = S.mc_worlds(B, n_cycles)
    keys = ("q", "dq", "contact", "toe_pos", "gait_domain", "agent_id", "team", "agent_initial", "Pr_refined", "Prd_refined", "Pobs_real",
            "alpha_buf", "success", "distance_to_fail", "n_not_converged")
    t = {k: torch.from_numpy(np.ascontiguousarray(w[k])).to(dev) for k in keys}
    raw = {k: t[k].data_ptr() for k in keys[:11]}
    raw.update(nq=w["nq"], ref_cols=w["ref_cols"], n_teams=w["n_teams"], n_obs=w["n_obs"])
    world = {k: t[k].data_ptr() for k in ("q", "dq", "contact", "toe_pos", "gait_domain", "alpha_buf", "success", "distance_to_fail", "n_not_converged")}
    s = pkg.BatchSolver(device=0, max_batch=B, max_obs=w["n_obs"])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.closed_loop(B, n_cycles, raw, world, mode=mode)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    succ = t["success"].cpu().numpy(); nnc = t["n_not_converged"].cpu().numpy(); dtf = t["distance_to_fail"].cpu().numpy()
    print("%-30s %d worlds x %d cycles in %.1f ms (%.2f M solves/s): success %.1f %%, mean distance to fail of the failed worlds %.2f m, "
          "cycles not converged %.2f %%" % (name, B, n_cycles, dt * 1e3, B * n_cycles / dt / 1e6, 100 * succ.mean(),
                                            dtf[succ == 0].mean() if (succ == 0).any() else float("nan"), 100 * nnc.sum() / (B * n_cycles)))
    s.close()
