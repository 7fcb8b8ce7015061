"""Seeded synthetic scenario generators of SURVEY.md 8(d) (PCG64, seed = 20251018 + 1000 * config_id, one spawned
stream per instance).  Instances are tuples (x0[4], xdes[16], feet_xy[2,4], contact[4], obs_xy[2,K])."""
import numpy as np

TS = 0.043
NOMINAL_FEET = np.array([[0.2188, 0.2188, -0.1472, -0.1472], [-0.1320, 0.1320, -0.1320, 0.1320]])  # MPC_dist.cpp:1206-1209


def cfg1_instance(c=4):
    """Single A1 solve: x0 = (0, 0.05, 0, 0), straight-line reference at 0.1 m/s, nominal stance, obstacle (1.0, 0.2)."""
    x0 = np.array([0.0, 0.05, 0.0, 0.0])
    xdes = np.zeros(16)
    for k in range(4):
        xdes[4 * k:4 * k + 4] = [0.1 * (k + 1) * TS, 0.1, 0.0, 0.0]
    contact = [1, 1, 1, 1] if c == 4 else [0, 1, 1, 0]
    return x0, xdes, NOMINAL_FEET.copy(), contact, np.array([[1.0], [0.2]])


def cfg3_instances(B, seed=20251018 + 3000):
    """B independent single-robot instances with randomised initial states and obstacle layouts."""
    ss = np.random.SeedSequence(seed)
    out = []
    goal = np.array([10.0, 0.0])
    nominal = np.array([[0.183, 0.183, -0.183, -0.183], [-0.132, 0.132, -0.132, 0.132]])
    for child in ss.spawn(B):
        rng = np.random.Generator(np.random.PCG64(child))
        p0 = np.array([rng.uniform(0, 9), rng.uniform(-2, 2)])
        v0 = rng.uniform(-0.3, 0.3, 2)
        x0 = np.array([p0[0], v0[0], p0[1], v0[1]])
        dirg = (goal - p0) / np.linalg.norm(goal - p0)
        vref = 0.2 * dirg
        xdes = np.zeros(16)
        for k in range(4):
            pk = p0 + vref * (k + 1) * TS
            xdes[4 * k:4 * k + 4] = [pk[0], vref[0], pk[1], vref[1]]
        u = rng.uniform()
        contact = [1, 1, 1, 1] if u < 0.5 else ([0, 1, 1, 0] if u < 0.75 else [1, 0, 0, 1])
        feet = p0[:, None] + nominal + rng.uniform(-0.03, 0.03, (2, 4))
        ox = rng.uniform(0, 9, 20) + rng.uniform(-0.6, 0.6, 20)
        oy = rng.uniform(-2, 2, 20) + rng.uniform(-0.6, 0.6, 20)
        obs = np.vstack([np.clip(ox, 1, 9), np.clip(oy, -3, 3)])
        out.append((x0, xdes, feet, contact, obs))
    return out


def cfg2_team():
    """4-agent team of SIM4:862,872-873 in reference-faithful form (neighbour state carried but unused): each agent at
    its start, v = 0, nominal stance, straight-line reference toward the goal (10, 0) at 0.2 m/s."""
    pstart = np.array([[0.0, 0.0], [0.0, -0.9], [-1.0, 0.0], [-1.0, -0.9]])
    obs = np.array([[2.2, 1, 0.85, 1, 1, 3, 3.0, 3, -100], [0.9, 1.2, -0.65, 2, -1.75, 0.2, -0.70, -2.2, -0.5 + 100]])  # Pobs_real, SIM4:872-873
    goal = np.array([10.0, 0.0])
    out = []
    for p0 in pstart:
        x0 = np.array([p0[0], 0.0, p0[1], 0.0])
        vref = 0.2 * (goal - p0) / np.linalg.norm(goal - p0)
        xdes = np.zeros(16)
        for k in range(4):
            pk = p0 + vref * (k + 1) * TS
            xdes[4 * k:4 * k + 4] = [pk[0], vref[0], pk[1], vref[1]]
        feet = p0[:, None] + NOMINAL_FEET
        out.append((x0, xdes, feet, [1, 1, 1, 1], obs.copy()))
    return out


def mc_worlds(B, n_cycles, seed=20251018 + 6000, n_obs=20, v_ref=0.2):
    """Closed-loop Monte-Carlo worlds (the reference's series of randomised runs, src/A1_Sim.cpp:944-1266, on the
    LIP template): one agent per world, start near the origin at rest, goal (10, 0), a straight-line reference at v_ref and
    n_obs obstacles scattered along the way (x ~ U(0.8, 4), y ~ U(-1, 1) about the line).  Returns host arrays in the layout of
    cbfnmpc_raw_in / cbfnmpc_world_state (SoA, per-world matrices column-major)."""
    rng = np.random.default_rng(seed)
    ref_cols = 4 * (n_cycles + 2)
    p0 = np.stack([rng.uniform(-0.2, 0.2, B), rng.uniform(-0.5, 0.5, B)])             # 2 x B
    goal = np.array([[10.0], [0.0]])
    dirv = (goal - p0) / np.linalg.norm(goal - p0, axis=0)
    v = v_ref * dirv                                                                  # 2 x B
    t = (np.arange(ref_cols) + 1) * TS
    Pr = np.zeros((B, ref_cols, 8)); Prd = np.zeros((B, ref_cols, 8))                 # [team][col][row] = column-major 8 x ref_cols
    Pr[:, :, 0] = p0[0][:, None] + v[0][:, None] * t[None, :]; Pr[:, :, 1] = p0[1][:, None] + v[1][:, None] * t[None, :]
    Prd[:, :, 0] = v[0][:, None]; Prd[:, :, 1] = v[1][:, None]
    obs = np.stack([rng.uniform(0.8, 4.0, (B, n_obs)), rng.uniform(-1.0, 1.0, (B, n_obs))], axis=2)  # [team][k][xy] = column-major 2 x K
    q = np.zeros((18, B)); dq = np.zeros((18, B))
    q[0] = p0[0]; q[1] = p0[1]
    toe = np.zeros((12, B))
    for i in range(4):
        toe[3 * i] = p0[0] + NOMINAL_FEET[0, i]; toe[3 * i + 1] = p0[1] + NOMINAL_FEET[1, i]
    return dict(q=q, dq=dq, nq=18, contact=np.ones((4, B), np.int32), toe_pos=toe, gait_domain=np.zeros(B, np.int32),
                agent_id=np.zeros(B, np.int32), team=np.arange(B, dtype=np.int32), agent_initial=p0.copy(),
                Pr_refined=Pr, Prd_refined=Prd, ref_cols=ref_cols, n_teams=B, Pobs_real=obs, n_obs=n_obs,
                alpha_buf=np.stack([p0[0], 0 * p0[0], p0[1], 0 * p0[0]]),
                success=np.ones(B, np.int32), distance_to_fail=np.full(B, 10.0), n_not_converged=np.zeros(B, np.int32))


# ---------------------------------------------------------------------------------------------------------------------
# Extended mode (BASELINE configs 4 and 5; no reference implementation exists).  Instances are tuples
# (x0[4], xdes[4N], feet_xy[ND,2,4], contact[4], obs_xy[2,K_obs]).
# ---------------------------------------------------------------------------------------------------------------------
TEAM_OFFSETS = np.array([[0.0, 0.0], [0.0, -0.9], [-1.0, 0.0], [-1.0, -0.9]])   # Pstart of SIM:1013 / SIM4:862 relative to the leader
EPS_NBR = float(np.float32(2.2))   # dec_vars_constr_cost.h:402 "epsilon = 2.2 // for robot to robot" (a float literal there)


def _ext_agent(rng, p0, v0, goal, N, K_obs, contact, obs_all):
    nominal = np.array([[0.183, 0.183, -0.183, -0.183], [-0.132, 0.132, -0.132, 0.132]])
    x0 = np.array([p0[0], v0[0], p0[1], v0[1]])
    vref = 0.2 * (goal - p0) / np.linalg.norm(goal - p0)
    xdes = np.zeros(4 * N)
    for k in range(N):
        pk = p0 + vref * (k + 1) * TS
        xdes[4 * k:4 * k + 4] = [pk[0], vref[0], pk[1], vref[1]]
    feet = np.zeros((N // 4, 2, 4))
    feet[0] = p0[:, None] + nominal + rng.uniform(-0.03, 0.03, (2, 4))
    for d in range(1, N // 4):                       # later domains: nominal stance about the reference position at the domain's start
        feet[d] = (p0 + vref * 4 * d * TS)[:, None] + nominal
    d2 = ((obs_all - p0[:, None]) ** 2).sum(0)
    near = np.argsort(d2, kind="stable")[:K_obs]     # the K_obs nearest obstacles, nearest first
    return x0, xdes, feet, contact, obs_all[:, near].copy()


def cfg4_teams(n_teams, seed=20251018 + 4000, N=4, K_obs=1):
    """Four-robot teams (BASELINE config 4): leader as cfg 3, the others at the reference's Pstart offsets from it, every member
    with 3 neighbour rows per grid point (+ K_obs obstacle rows).  The members of a team share the contact pattern and the
    obstacle field.  Returns a flat, team-major list of 4 * n_teams instances."""
    ss = np.random.SeedSequence(seed)
    goal = np.array([10.0, 0.0])
    out = []
    for child in ss.spawn(n_teams):
        rng = np.random.Generator(np.random.PCG64(child))
        lead = np.array([rng.uniform(1, 9), rng.uniform(-2, 2)])
        u = rng.uniform()
        contact = [1, 1, 1, 1] if u < 0.5 else ([0, 1, 1, 0] if u < 0.75 else [1, 0, 0, 1])
        ox = rng.uniform(0, 9, 20) + rng.uniform(-0.6, 0.6, 20); oy = rng.uniform(-2, 2, 20) + rng.uniform(-0.6, 0.6, 20)
        obs_all = np.vstack([np.clip(ox, 1, 9), np.clip(oy, -3, 3)])
        for a in range(4):
            out.append(_ext_agent(rng, lead + TEAM_OFFSETS[a], rng.uniform(-0.3, 0.3, 2), goal, N, K_obs, contact, obs_all))
    return out


def cfg5_instances(B, seed=20251018 + 5000, N=8, K_obs=16, team=True):
    """Long-horizon stress (BASELINE config 5): N = 8 grid points (two gait domains), the 16 nearest obstacles as rows at every
    grid point, plus (team=True) the 3 neighbour rows of a four-robot team: B must then be a multiple of 4."""
    if team:
        return cfg4_teams(B // 4, seed=seed, N=N, K_obs=K_obs)
    ss = np.random.SeedSequence(seed)
    goal = np.array([10.0, 0.0])
    out = []
    for child in ss.spawn(B):
        rng = np.random.Generator(np.random.PCG64(child))
        p0 = np.array([rng.uniform(0, 9), rng.uniform(-2, 2)])
        u = rng.uniform()
        contact = [1, 1, 1, 1] if u < 0.5 else ([0, 1, 1, 0] if u < 0.75 else [1, 0, 0, 1])
        ox = rng.uniform(0, 9, 20) + rng.uniform(-0.6, 0.6, 20); oy = rng.uniform(-2, 2, 20) + rng.uniform(-0.6, 0.6, 20)
        obs_all = np.vstack([np.clip(ox, 1, 9), np.clip(oy, -3, 3)])
        out.append(_ext_agent(rng, p0, rng.uniform(-0.3, 0.3, 2), goal, N, K_obs, contact, obs_all))
    return out
