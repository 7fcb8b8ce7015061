#!/usr/bin/env python
"""BASELINE config 4 as worded: four-robot teams whose members live on DIFFERENT GPUs, the neighbours' predicted trajectories
all-gathered every interior-point iteration (extended mode, SURVEY.md A.6 / 8e collective C1; no reference implementation exists).
The collective is the library's own: cbfnmpc_comm_init creates an NCCL communicator inside libcbfnmpc.so and
cbfnmpc_ext_solve_sharded calls ncclAllGather between its per-iteration launches; torch.distributed is used here only to hand the
128-byte NCCL id to the other ranks and to collect the answers for the check.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/agent_sharded_ext.py [teams per group] [N] [K_obs]

World 2: each rank holds 2 members of every team; world 4: one member each; world 8: two groups of 4 ranks, each with its own teams.
Rank 0 checks the gathered answers against the same teams solved team-major on its own GPU (members of a team in one warp,
exchange by shuffles) and prints one JSON line.  The same teams are then solved by cbfnmpc_ext_solve_fused -- the exchange inside
ONE launch, members storing their published trajectories into each other's GPU memory (NVLink peer access) and waiting on sequence
flags -- and compared bit for bit ("fused" block of the line)."""
import json
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srb_cbf_nmpc_b200 as pkg  # noqa: E402
from srb_cbf_nmpc_b200 import capi, sharding as SH, scenarios as S  # noqa: E402

n_teams = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K_obs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
T = 4
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
grp, grank, G, mpr, n_groups = SH.sharded_group(rank, world, T)
# the library's communicator over the ranks of this group: the group's first rank makes the id, torch hands it over
uid = torch.zeros(128, dtype=torch.uint8, device=dev)
if grank == 0:
    uid = torch.frombuffer(bytearray(pkg.BatchSolver.comm_unique_id()), dtype=torch.uint8).to(dev)
subgroups = [dist.new_group(list(range(g * G, (g + 1) * G))) for g in range(n_groups)]
dist.broadcast(uid, src=grp * G, group=subgroups[grp])
solver = pkg.BatchSolver(device=local, max_batch=mpr * n_teams, max_obs=max(K_obs, 1))
solver.comm_init(G, grank, bytes(uid.cpu().numpy().tobytes()))
# every group has its own teams (seeded by the group index); a rank takes its member blocks of them
insts = S.cfg4_teams(n_teams, seed=20251018 + 4000 + grp, N=N, K_obs=K_obs)
members = list(range(grank * mpr, (grank + 1) * mpr))
idx = SH.member_major(n_teams, T, members)
mine = [insts[i] for i in idx]
B = len(mine)
d = capi.ext_soa_from_instances(mine, N)
tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(N).items() if k != "pub"}
tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
ip = {k: v.data_ptr() for k, v in tin.items()}; op = {k: v.data_ptr() for k, v in tout.items()}
stats = None
for rep in range(3):   # the last repetition is the one reported
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    stats = solver.ext_solve_sharded(n_teams, ip, op, d["n_obs"], N=N, team_size=T, eps_nbr=S.EPS_NBR)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
# the same teams with the exchange inside ONE launch (cbfnmpc_ext_solve_fused: peer-memory mailboxes instead of a collective per
# iteration); answers must be bit-identical to the per-iteration path's
tout2 = {k: torch.zeros_like(v) for k, v in tout.items()}
op2 = {k: v.data_ptr() for k, v in tout2.items()}
fstats = None
for rep in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    fstats = solver.ext_solve_fused(n_teams, ip, op2, d["n_obs"], N=N, team_size=T, eps_nbr=S.EPS_NBR)
    torch.cuda.synchronize()
    fwall = time.perf_counter() - t0
same = all(torch.equal(tout[k], tout2[k]) for k in tout)
ft = torch.tensor([fstats["total_ms"], fwall * 1e3, 0.0 if same else 1.0, float(fstats["fused"])], dtype=torch.float64, device=dev)
fmax = ft.clone(); dist.all_reduce(fmax, op=dist.ReduceOp.MAX)
fmin = ft.clone(); dist.all_reduce(fmin, op=dist.ReduceOp.MIN)
# collect (X, slack, status, team_iters) of every rank on rank 0
pack = torch.cat([tout["X"], tout["slack"], tout["status"].double(), tout["team_iters"].double()]).contiguous()
allp = [torch.empty_like(pack) for _ in range(world)]
dist.all_gather(allp, pack)
tmax = torch.tensor([stats["total_ms"], wall * 1e3], dtype=torch.float64, device=dev)
dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
if rank == 0:
    res = dict(world=world, ranks_per_team=G, groups=n_groups, teams_per_group=n_teams, N=N, K_obs=K_obs, instances=int(T * n_teams * n_groups),
               iterations=stats["iterations"], launches=stats["launches"], exchanges=stats["exchanges"],
               allgather_bytes_per_rank=stats["bytes_per_exchange"], device_ms=float(tmax[0]), wall_ms=float(tmax[1]),
               ms_per_iteration=float(tmax[0]) / max(stats["launches"], 1))
    # check group 0 against the team-major solve on this GPU
    s0 = pkg.BatchSolver(device=local, max_batch=T * n_teams, max_obs=max(K_obs, 1))
    ref_insts = S.cfg4_teams(n_teams, seed=20251018 + 4000, N=N, K_obs=K_obs)
    t1 = time.perf_counter()
    ref = s0.ext_solve_host(capi.ext_soa_from_instances(ref_insts, N), N=N, team_size=T, eps_nbr=S.EPS_NBR)
    res["team_major_host_call_ms"] = 1e3 * (time.perf_counter() - t1)
    X = np.zeros((4 * N, T * n_teams)); sl = np.zeros(T * n_teams); st = np.zeros(T * n_teams, int); ti = np.zeros(T * n_teams, int)
    for r in range(G):
        ii = SH.member_major(n_teams, T, list(range(r * mpr, (r + 1) * mpr)))
        a = allp[r].cpu().numpy()
        X[:, ii] = a[:4 * N]; sl[ii] = a[4 * N]; st[ii] = a[4 * N + 1].astype(int); ti[ii] = a[4 * N + 2].astype(int)
    ok = np.repeat(((st == 0) & (ref["status"][0] == 0)).reshape(-1, T).all(1), T)
    res.update(status_equal=float((st == ref["status"][0]).mean()), converged_fraction=float((st == 0).mean()),
               max_abs_diff_X=float(np.abs(X - ref["X"])[:, ok].max()), max_abs_diff_slack=float(np.abs(sl - ref["slack"][0])[ok].max()),
               team_iters_equal=float((ti == ref["team_iters"][0])[ok].mean()), teams_compared=int(ok.sum() // T))
    conv = int((st == 0).sum()) * n_groups
    res["converged_solves_per_s"] = conv / (res["device_ms"] * 1e-3)
    res["fused"] = dict(in_one_launch=bool(fmin[3] == 1.0), device_ms=float(fmax[0]), wall_ms=float(fmax[1]), launches=fstats["launches"],
                        iterations=fstats["iterations"], peer_bytes_per_rank_and_iteration=fstats["bytes_per_exchange"],
                        ms_per_iteration=float(fmax[0]) / max(fstats["iterations"], 1), converged_solves_per_s=conv / (float(fmax[0]) * 1e-3),
                        bit_identical_to_the_per_iteration_path=bool(fmax[2] == 0.0))
    print(json.dumps(res))
    assert res["status_equal"] == 1.0 and res["max_abs_diff_X"] < 1e-9, res
    assert res["fused"]["bit_identical_to_the_per_iteration_path"], res["fused"]
solver.comm_destroy()
dist.barrier(); dist.destroy_process_group()
