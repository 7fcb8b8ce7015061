"""Multi-GPU plumbing of the batch path: instances are independent, so ranks own contiguous ranges of the
global instance set and exchange nothing in the data path; torch.distributed is used only for the barrier and for the
max-over-ranks / sum-over-ranks reduction of the timing and work counters.

The one exchange the reference has is `state_other`: every agent receives the other agent's (x, y, vx, vy) once per control
cycle (MPC_dist::updateState, src/MPC_dist.cpp:1200) -- unused by the shipped run_NMPC, consumed by its disabled inter-agent rule
(:391).  For batches stored team by team it is a local gather (cbfnmpc_team_states); when the members of a team live on
different GPUs (rank r holds member r of every team) it is one all-gather per cycle: `exchange_state_other`."""
import torch


def shard_range(n_total, rank, world):
    """Contiguous range [lo, hi) of rank `rank`; remainders go to the first ranks.  Team-major: with `team` > 1 use shard_teams."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_teams(n_teams, team, rank, world):
    """Same, in whole teams of `team` agents (a team's agents stay on one GPU)."""
    lo, hi = shard_range(n_teams, rank, world)
    return lo * team, hi * team


def reduce_over_ranks(dist, times, work, device):
    """MAX over ranks of `times` (list of floats), SUM over ranks of `work` (list of floats)."""
    t = torch.tensor(times, dtype=torch.float64, device=device)
    w = torch.tensor(work, dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return [float(x) for x in t], [float(x) for x in w]


def gather_results(dist, local, device):
    """Optional final gather of per-rank result arrays (dict of [n][B_local] numpy arrays) onto every rank, rank-major."""
    import numpy as np
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    out = {}
    for k, v in local.items():
        t = torch.from_numpy(np.ascontiguousarray(v.T)).to(device)          # [B_local][n]: batch-major so that ranks concatenate
        sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(dist.get_world_size())]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64, device=device))
        bmax = int(max(int(s) for s in sizes))
        pad = torch.zeros((bmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=device); pad[:t.shape[0]] = t
        parts = [torch.zeros_like(pad) for _ in sizes]
        dist.all_gather(parts, pad)
        out[k] = torch.cat([p[:int(s)] for p, s in zip(parts, sizes)]).cpu().numpy().T
    return out


def exchange_state_other(dist, q, dq, device=None):
    """Agent-sharded layout: rank r holds member r of every team, team t at local index t on every rank; world size = team size.
    q, dq: [nq][B_local] tensors (rows 0, 1 are read).  Returns state_other [4*(world-1)][B_local]: rows 4k..4k+3 = (x, y, vx, vy)
    of the k-th OTHER member of the team (ranks in increasing order, own rank skipped) -- the layout cbfnmpc_raw_in.state_other
    takes.  One all-gather of 4 x B_local doubles per rank and cycle (NCCL on GPUs, gloo in the CPU tests)."""
    mine = torch.stack([q[0], q[1], dq[0], dq[1]]).contiguous()
    if device is not None:
        mine = mine.to(device)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return mine.new_zeros((0, mine.shape[1]))
    world, rank = dist.get_world_size(), dist.get_rank()
    allb = torch.empty((world,) + tuple(mine.shape), dtype=mine.dtype, device=mine.device)
    try:
        dist.all_gather_into_tensor(allb, mine)           # one collective into one buffer
    except (RuntimeError, NotImplementedError, AttributeError):
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        allb = torch.stack(parts)
    others = torch.tensor([r for r in range(world) if r != rank], dtype=torch.long, device=mine.device)
    return allb.index_select(0, others).reshape(4 * (world - 1), mine.shape[1])


# ---------------------------------------------------------------------------------------------------------------- extended mode
def sharded_group(rank, world, team_size):
    """Agent-sharded teams over `world` ranks (cbfnmpc_ext_solve_sharded): ranks are cut into groups of G = min(world, team_size)
    ranks that share teams; inside a group rank g holds members [g m, (g + 1) m), m = team_size / G, of every team of the group.
    Returns (group index, rank inside the group, G, members per rank, number of groups)."""
    G = min(world, team_size)
    if team_size % G != 0 or world % G != 0:
        raise ValueError("world size %d and team size %d do not tile" % (world, team_size))
    return rank // G, rank % G, G, team_size // G, world // G


def member_major(n_teams, team_size, members):
    """Indices into a team-major batch (instance = team * team_size + member) of the member-major local batch that holds `members`
    of every team (local instance = local member * n_teams + team)."""
    import numpy as np
    return np.array([t * team_size + m for m in members for t in range(n_teams)], dtype=np.int64)
