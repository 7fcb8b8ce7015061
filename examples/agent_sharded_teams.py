#!/usr/bin/env python
"""Teams whose members live on different GPUs (rank r holds member r of every team; world size = team size): the one exchange
of the path -- every agent receives the other members' (x, y, vx, vy) once per control cycle, the `state_vec` of
MPC_dist::updateState (src/MPC_dist.cpp:1195-1202) -- is one NCCL all-gather per cycle (sharding.exchange_state_other); the
solves themselves stay independent.  Here the neighbours, extrapolated over the horizon, compete with the real obstacles for the
distance rows (the reference's disabled rule, src/MPC_dist.cpp:391, `state_other` / `n_other` of cbfnmpc_raw_in); the answers
are checked against the same teams stored team by team on rank 0 (cbfnmpc_team_states, no communication), bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/agent_sharded_teams.py [teams] [cycles]
"""
import json
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srb_cbf_nmpc_b200 as pkg  # noqa: E402
from srb_cbf_nmpc_b200 import capi, sharding as SH, mpc_dist as MD  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
A, K, nq, ref_cols, Ts = world, 8, 18, 4 * n_cycles + 8, 43 * 0.001


def team_world(seed=7):
    """T teams of A agents walking side by side at 0.2 m/s through K obstacles; every rank builds the same world."""
    rng = np.random.default_rng(seed)
    p0 = np.stack([rng.uniform(0, 6, T), rng.uniform(-1.5, 1.5, T)], 1)
    off = np.array([[0.0, 0.0], [0.0, 0.6], [-0.7, 0.0], [-0.7, 0.6]])[:A]
    Pobs = np.stack([rng.uniform(0, 9, (T, K)), rng.uniform(-2, 2, (T, K))], 1)
    Pr = np.zeros((T, 8, ref_cols)); Prd = np.zeros((T, 8, ref_cols))
    for a in range(A):
        Pr[:, 2 * a] = (p0[:, 0] + off[a, 0])[:, None] + 0.2 * Ts * np.arange(ref_cols)[None]
        Pr[:, 2 * a + 1] = (p0[:, 1] + off[a, 1])[:, None]; Prd[:, 2 * a] = 0.2
    return p0, off, Pobs, Pr, Prd


def run(members, teams_major):
    """members: list of member indices this process steps (one for the sharded run, all for the team-major check)."""
    p0, off, Pobs, Pr, Prd = team_world()
    nA = len(members); B = T * nA
    # batch order: team-major over the members held here
    team = np.repeat(np.arange(T), nA).astype(np.int32); agent = np.tile(np.array(members), T).astype(np.int32)
    q = np.zeros((nq, B)); dq = np.zeros((nq, B))
    q[0] = p0[team, 0] + off[agent, 0]; q[1] = p0[team, 1] + off[agent, 1]
    cm = lambda x: np.ascontiguousarray(np.transpose(x, (0, 2, 1)))
    toe = np.zeros((12, B))
    tens = dict(q=q, dq=dq, contact=np.ones((4, B), np.int32), toe_pos=toe, gait_domain=np.zeros(B, np.int32), agent_id=agent, team=team,
                agent_initial=np.stack([q[0], q[1]]), Pr_refined=cm(Pr), Prd_refined=cm(Prd), Pobs_real=cm(Pobs))
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in tens.items()}
    abuf = torch.stack([t["q"][0], 0 * t["q"][0], t["q"][1], 0 * t["q"][0]]).contiguous()
    s = pkg.BatchSolver(device=local, max_batch=B, max_obs=K + A - 1)
    out = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.OUT_F64.items()}
    out.update({k: torch.zeros((n, B), dtype=torch.int32, device=dev) for k, n in capi.OUT_I32.items()})
    out["active"] = torch.zeros((2, B), dtype=torch.int64, device=dev)
    so = torch.zeros((4 * (A - 1), B), dtype=torch.float64, device=dev)
    raw = {k: v.data_ptr() for k, v in t.items()}
    raw.update(nq=nq, ref_cols=ref_cols, n_teams=T, n_obs=K, state_other=so.data_ptr(), n_other=A - 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_x = t_s = 0.0
    traj = []
    st = torch.cuda.Stream(device=dev)   # one explicit stream for torch, NCCL (which orders itself against it) and the library
    torch.cuda.synchronize()
    torch.cuda.set_stream(st)
    for cyc in range(n_cycles):
        ev[0].record()
        if teams_major:
            s.team_states(B, A, t["q"].data_ptr(), t["dq"].data_ptr(), so.data_ptr(), stream=st.cuda_stream)
        else:
            so.copy_(SH.exchange_state_other(dist, t["q"], t["dq"]))
        ev[1].record()
        s.run_cycle(B, raw, abuf.data_ptr(), {k: v.data_ptr() for k, v in out.items()}, mode=capi.MODE_QP_NLP, stream=st.cuda_stream)
        # perfect tracking of the template: the state at the end of the domain is x_4 of the solution (cf. cbfnmpc_closed_loop)
        X = out["X"]
        t["q"][0].copy_(X[12]); t["dq"][0].copy_(X[13]); t["q"][1].copy_(X[14]); t["dq"][1].copy_(X[15])
        abuf.copy_(out["alpha_buf_next"]); t["gait_domain"] += 1
        for i, (fx, fy) in enumerate(MD.INIT_FOOTPRINT):
            t["toe_pos"][3 * i].copy_(t["q"][0] + fx); t["toe_pos"][3 * i + 1].copy_(t["q"][1] + fy)
        ev[2].record(); torch.cuda.synchronize()
        if cyc > 0: t_x += ev[0].elapsed_time(ev[1]); t_s += ev[1].elapsed_time(ev[2])
        traj.append(torch.cat([X, out["status"].double()]).clone())
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    s.close()
    return torch.stack(traj), t_x / max(n_cycles - 1, 1), t_s / max(n_cycles - 1, 1)


mine, ms_x, ms_s = run([rank], teams_major=False)                         # [cycles][17][T]
parts = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(parts, mine)
tt = torch.tensor([ms_x, ms_s], dtype=torch.float64, device=dev)
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    ref, _, _ = run(list(range(A)), teams_major=True)                     # [cycles][17][T*A], team-major
    same = all(torch.equal(parts[a], ref[:, :, a::A]) for a in range(A))
    conv = float((ref[:, 16] == 0).double().mean())
    print(json.dumps({"example": "agent_sharded_teams", "n_gpus": world, "teams": T, "cycles": n_cycles, "agents": T * A,
                      "exchange": "all_gather of 4 x %d doubles per rank and cycle (NCCL)" % T, "exchange_ms": float(tt[0]),
                      "solve_ms": float(tt[1]), "identical_to_team_major_single_gpu": bool(same), "converged_fraction": conv}))
    assert same
dist.barrier(); dist.destroy_process_group()
