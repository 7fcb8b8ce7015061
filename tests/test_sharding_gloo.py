"""N > 1 host logic on CPU: world_size-2 gloo processes (contiguous shards, max/sum reduction, result gather)."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from srb_cbf_nmpc_b200 import sharding as SH, scenarios as S, capi
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 37
    lo, hi = SH.shard_range(n_total, rank, world)
    insts = S.cfg3_instances(n_total)[lo:hi]
    d = capi.soa_from_instances(insts)
    # stand-in for the per-rank solve: a deterministic function of the instance (no GPU in this test)
    local = {"X": d["xdes"] * 2.0, "status": (d["contact"].sum(0, keepdims=True) == 4).astype(np.int32)}
    times, work = SH.reduce_over_ranks(dist, [10.0 + rank, 1.0 - 0.1 * rank], [hi - lo, float(local["status"].sum())], torch.device("cpu"))
    allr = SH.gather_results(dist, local, torch.device("cpu"))
    dist.barrier()
    q.put((rank, lo, hi, times, work, allr["X"].shape, float(allr["X"].sum()), int(allr["status"].sum())))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    import sys
    sys.path.insert(0, ROOT)
    from srb_cbf_nmpc_b200 import sharding as SH, scenarios as S, capi
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps: p.join(timeout=60)
    assert [p.exitcode for p in ps] == [0, 0]
    (r0, lo0, hi0, t0, w0, shp0, sx0, ss0), (r1, lo1, hi1, t1, w1, shp1, sx1, ss1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 19, 19, 37)                 # contiguous, remainder to the first rank
    assert t0 == t1 == [11.0, 1.0]                                 # MAX over ranks
    d = capi.soa_from_instances(S.cfg3_instances(37))
    assert w0 == w1 == [37.0, float((d["contact"].sum(0) == 4).sum())]   # SUM over ranks
    assert shp0 == shp1 == (16, 37)
    assert abs(sx0 - 2.0 * d["xdes"].sum()) < 1e-9 and sx0 == sx1 and ss0 == ss1 == int(w0[1])


def test_shard_ranges_cover_exactly():
    import sys
    sys.path.insert(0, ROOT)
    from srb_cbf_nmpc_b200 import sharding as SH
    for n in (0, 1, 7, 4096, 65536):
        for world in (1, 2, 4, 8):
            rs = [SH.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    assert SH.shard_teams(16384, 4, 3, 8) == (3 * 2048 * 4, 4 * 2048 * 4)


def _worker_exchange(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    from srb_cbf_nmpc_b200 import sharding as SH
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    qq = torch.randn((18, 5), generator=g, dtype=torch.float64); dq = torch.randn((18, 5), generator=g, dtype=torch.float64)
    so = SH.exchange_state_other(dist, qq, dq)
    q.put((rank, qq[:2].numpy(), dq[:2].numpy(), so.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_state_other_exchange():
    """Agent-sharded teams (rank r holds member r of every team): one all-gather per cycle hands every agent the other member's
    (x, y, vx, vy) -- the state_vec of MPC_dist::updateState (src/MPC_dist.cpp:1195-1202)."""
    world, port = 2, 31500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_exchange, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted((q.get(timeout=120) for _ in ps), key=lambda t: t[0])
    for p in ps: p.join(timeout=60)
    assert [p.exitcode for p in ps] == [0, 0]
    (_, q0, dq0, so0), (_, q1, dq1, so1) = res
    assert so0.shape == so1.shape == (4, 5)
    assert np.array_equal(so0, np.vstack([q1, dq1])) and np.array_equal(so1, np.vstack([q0, dq0]))


def test_agent_sharded_layout_of_the_extended_mode():
    """Host-side logic of cbfnmpc_ext_solve_sharded's layout: groups of ranks that share teams, member blocks per rank, and the
    member-major <-> team-major index maps (what examples/agent_sharded_ext.py and bench.py build their local batches with)."""
    import numpy as np
    from srb_cbf_nmpc_b200 import sharding as SH
    assert SH.sharded_group(0, 1, 4) == (0, 0, 1, 4, 1)
    assert SH.sharded_group(1, 2, 4) == (0, 1, 2, 2, 1)
    assert SH.sharded_group(3, 4, 4) == (0, 3, 4, 1, 1)
    assert SH.sharded_group(6, 8, 4) == (1, 2, 4, 1, 2)
    n_teams, T = 5, 4
    seen = np.zeros(n_teams * T, int)
    for world in (1, 2, 4):
        seen[:] = 0
        for rank in range(world):
            grp, grank, G, mpr, ng = SH.sharded_group(rank, world, T)
            idx = SH.member_major(n_teams, T, list(range(grank * mpr, (grank + 1) * mpr)))
            assert len(idx) == mpr * n_teams
            # local instance = local member * n_teams + team; the neighbour blocks of the gathered buffer are indexed by GLOBAL member
            for li, gi in enumerate(idx):
                assert gi // T == li % n_teams and gi % T == grank * mpr + li // n_teams
            seen[idx] += 1
        assert (seen == 1).all()
    import pytest
    with pytest.raises(ValueError):
        SH.sharded_group(0, 3, 4)
