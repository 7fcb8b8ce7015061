#!/usr/bin/env python
"""Stress of cbfnmpc_ext_solve_fused on 2 or 4 GPUs (one process per GPU): batch sizes that grow and shrink (the mailboxes are
re-allocated and re-mapped), N = 4 and N = 8, a member with a bad contact pattern, repeated calls, the communicator destroyed and
made again; after every call the answers must equal cbfnmpc_ext_solve_sharded's bit for bit on every rank.  With CBFNMPC_NO_IPC=1 in
the environment no rank can map a peer: every rank must then take the per-iteration path (and still agree).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/fused_exchange_stress.py

Status at the end of round 2: passes on 2 GPUs (also with CBFNMPC_NO_IPC=1).  The first attempt on 4 GPUs ran into its time limit; a
second one with per-call progress output (STRESS_VERBOSE=1) showed where: in the PER-ITERATION path of the case with the one-footed
member -- the rank that holds that member (and, with one member per rank, no team mate of it) counted the team out, left its launch
loop and with it the all-gathers its peers were still calling.  cbfnmpc_ext_solve_sharded now gathers and sums that counter over the
ranks; the fix is verified on 2 GPUs (where the old code could not fail either), the 4-GPU re-run could not be made any more (the GPU
budget of the round was spent).  The fused path itself, and the plain 4- and 8-GPU runs of both paths (examples/agent_sharded_ext.py,
bench.py at N = 4 and 8), were green throughout."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import srb_cbf_nmpc_b200 as pkg  # noqa: E402
from srb_cbf_nmpc_b200 import capi, sharding as SH, scenarios as S  # noqa: E402

T = 4
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
grp, grank, G, mpr, n_groups = SH.sharded_group(rank, world, T)
solver = pkg.BatchSolver(device=local, max_batch=mpr * 4096, max_obs=20)


def make_comm():
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if grank == 0:
        uid = torch.frombuffer(bytearray(pkg.BatchSolver.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(uid, src=grp * G)
    solver.comm_init(G, grank, bytes(uid.cpu().numpy().tobytes()))


make_comm()
log = []
cases = [(4, 1, 64, False), (4, 1, 1024, False), (8, 16, 96, False), (4, 1, 32, True), (4, 1, 4096, False), (4, 1, 64, False)]
for rnd in range(2):
    for N, K, n_teams, bad in cases:
        insts = S.cfg4_teams(n_teams, seed=99 + n_teams + N, N=N, K_obs=K)
        if bad:
            a, b_, c_, d_, e = insts[4 * 3 + 1]; insts[4 * 3 + 1] = (a, b_, c_, [1, 0, 0, 0], e)
        idx = SH.member_major(n_teams, T, list(range(grank * mpr, (grank + 1) * mpr)))
        mine = [insts[i] for i in idx]; B = len(mine)
        d = capi.ext_soa_from_instances(mine, N)
        tin = {k: torch.from_numpy(d[k]).to(dev) for k in ("x0", "xdes", "feet_xy", "contact", "obs_xy", "alpha_buf")}
        outs = []
        for fi, fn in enumerate((solver.ext_solve_sharded, solver.ext_solve_fused, solver.ext_solve_fused)):
            if os.environ.get("STRESS_VERBOSE"): print("rank", rank, "round", rnd, "case", N, n_teams, bad, "call", fi, flush=True)
            tout = {k: torch.zeros((n, B), dtype=torch.float64, device=dev) for k, n in capi.ext_out_rows(N).items() if k != "pub"}
            tout.update({k: torch.zeros((1, B), dtype=torch.int32, device=dev) for k in capi.EXT_OUT_I32})
            dist.barrier()
            st = fn(n_teams, {k: v.data_ptr() for k, v in tin.items()}, {k: v.data_ptr() for k, v in tout.items()}, d["n_obs"], N=N, team_size=T, eps_nbr=S.EPS_NBR)
            torch.cuda.synchronize()
            outs.append((st, tout))
        keep = torch.ones(B, dtype=torch.bool, device=dev)
        if bad:   # the per-iteration path reads an unwritten block for the one-footed member: its team is not compared
            for m in range(mpr): keep[m * n_teams + 3] = False
        same = all(torch.equal(outs[0][1][k][..., keep], outs[1][1][k][..., keep]) and torch.equal(outs[1][1][k], outs[2][1][k]) for k in outs[0][1])
        flag = torch.tensor([0.0 if same else 1.0, -float(outs[1][0]["fused"])], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        log.append(dict(round=rnd, N=N, n_teams=n_teams, bad=bad, fused=bool(flag[1] == -1.0), identical=bool(flag[0] == 0.0),
                        sharded_ms=outs[0][0]["total_ms"], fused_ms=outs[2][0]["total_ms"], launches=outs[2][0]["launches"]))
        if rank == 0: print("case", json.dumps(log[-1]), flush=True)
        assert flag[0] == 0.0, log[-1]
        assert bool(flag[1] == -1.0) == (os.environ.get("CBFNMPC_NO_IPC") is None or G == 1), log[-1]   # (CBFNMPC_NO_IPC=1: every rank falls back)
    torch.cuda.synchronize(); dist.barrier()
    solver.comm_destroy(); dist.barrier(); make_comm()   # the communicator goes and comes back: the mailboxes are mapped again
if rank == 0:
    print(json.dumps(dict(world=world, ranks_per_team=G, cases=log, all_identical=all(c["identical"] for c in log), all_fused=all(c["fused"] for c in log))))
solver.comm_destroy()
dist.barrier(); dist.destroy_process_group()
