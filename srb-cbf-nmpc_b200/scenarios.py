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
