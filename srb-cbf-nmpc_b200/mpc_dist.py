"""Host-side mirror of the reference's per-agent planner class `MPC_dist` (include/MPC_dist.hpp:141-189,
src/MPC_dist.cpp), for the hot path only: same method names, argument meaning and hidden state
(gaitDomain_, mpc_state_alpha_buffer_, agent_Initial_), with every solve executed by the CUDA library.

`MPC_dist_team` steps many agents in one batched launch (one instance per agent)."""
import numpy as np
from . import capi
from .api import BatchSolver

NDOMAIN = 4            # include/global_loco_opts.h:9
NUMBER_OF_AGENTS = 4   # :14
NUMBER_OF_OBS = 20     # :15
TS = 43 * 0.001         # TSOPTTICK * 0.001, evaluated like the reference does (not the literal 0.043) (include/global_loco_opts.h:10; MPC_dist.cpp:104)
INIT_FOOTPRINT = np.array([[0.2188, -0.1320], [0.2188, 0.1320], [-0.1472, -0.1320], [-0.1472, 0.1320]])  # MPC_dist.cpp:1206-1209


class MPC_dist:
    """One agent.  Construct with a shared BatchSolver (or let it create one with max_batch=1)."""

    def __init__(self, solver=None):
        self.solver = solver or BatchSolver(max_batch=1, max_obs=NUMBER_OF_OBS + NUMBER_OF_AGENTS - 1)
        self.use_snopt = False                      # include/MPC_dist.hpp:139
        # the reference's disabled inter-agent rule (MPC_dist.cpp:391): the other agent, extrapolated over the horizon, competes
        # with the real obstacles for the distance rows.  False = the reference as shipped (state_other carried, unread).
        self.use_state_other = False
        self.agent_id_ = 0                          # MPC_dist.cpp:58
        self.gaitDomain_ = 0                        # :55
        self.domain_ = 0                            # :54
        self.mpc_state_alpha_buffer_ = np.zeros(4)  # :53
        self.agent_Initial_ = np.zeros(2)
        self.Pstart_ = np.zeros(2 * NUMBER_OF_AGENTS)
        self.Pobs = np.zeros((2, NUMBER_OF_OBS)); self.Pobs_real = np.zeros((2, NUMBER_OF_OBS))
        self.Pr_refined_ = None; self.Prd_refined_ = None
        self.q = np.zeros(18); self.dq = np.zeros(18); self.contactInd = np.array([1, 1, 1, 1], np.int32)
        self.toePos_ = np.zeros((3, 4)); self.state_other = np.zeros(4)
        self.alpha_COM_traj_e_ = np.zeros((4, 5)); self.mpc_state_e_x_eventbased_ = np.zeros(16)
        self.mpc_state_e_u_eventbased_ = np.zeros(4); self.qp_solution_eventbased_ = np.zeros(24)
        self.status = None; self.kkt = None

    # ---- setup (MPC_dist.cpp:1106-1137)
    def setAgentID(self, agent_id): self.agent_id_ = int(agent_id)

    def setPstart(self, Pstart):
        self.Pstart_ = np.asarray(Pstart, float).reshape(-1).copy()
        self.agent_Initial_ = self.Pstart_[2 * self.agent_id_:2 * self.agent_id_ + 2].copy()
        self.mpc_state_alpha_buffer_ = np.array([self.agent_Initial_[0], 0.0, self.agent_Initial_[1], 0.0])  # :1111

    def setPobs(self, Pobs): self.Pobs = np.asarray(Pobs, float).copy()
    def setPobs_real(self, Pobs): self.Pobs_real = np.asarray(Pobs, float).copy()

    def setReferenceTrajectory(self, Pr_refined, Prd_refined):
        """Output of generateReferenceTrajectory() (MPC_dist.cpp:930-1104, 8 x 2500 each); that generator is outside the
        hot path, so the arrays are supplied by the caller."""
        self.Pr_refined_ = np.asarray(Pr_refined, float); self.Prd_refined_ = np.asarray(Prd_refined, float)

    # ---- per cycle
    def updateState(self, qin, dqin, ind, toePos, state_vec):  # MPC_dist.cpp:1195-1202
        self.q[:] = np.asarray(qin, float)[:18]; self.dq[:] = np.asarray(dqin, float)[:18]
        self.contactInd[:] = np.asarray(ind, np.int32)[:4]
        self.toePos_ = np.asarray(toePos, float).reshape(3, 4).copy()
        self.state_other = np.asarray(state_vec, float).reshape(-1).copy()   # 4 numbers per neighbour

    def copPlanner_eventbase(self):  # MPC_dist.cpp:702-782: columns [N*gaitDomain_, +N) of this agent's rows
        N, a, g = 4, self.agent_id_, self.gaitDomain_
        cols = slice(N * g, N * g + N)
        qref = np.vstack([self.Pr_refined_[2 * a, cols], self.Prd_refined_[2 * a, cols],
                          self.Pr_refined_[2 * a + 1, cols], self.Prd_refined_[2 * a + 1, cols]])  # 4 x N
        return qref.T.reshape(-1)  # column-major vectorisation ([x,xd,y,yd] x 4), :780

    def footholdsPlanner(self):  # MPC_dist.cpp:1204-1266 (nominal stance during the first two domains)
        if self.gaitDomain_ <= 1:
            self.toePos_[0, :] = INIT_FOOTPRINT[:, 0] + self.agent_Initial_[0]
            self.toePos_[1, :] = INIT_FOOTPRINT[:, 1] + self.agent_Initial_[1]
            self.toePos_[2, :] = 0.0
        return self.toePos_[:2, :].copy()

    def _instance(self):
        xdes = self.copPlanner_eventbase()
        feet = self.footholdsPlanner()
        x0 = np.array([self.q[0], self.dq[0], self.q[1], self.dq[1]])  # :226-229
        obs = self.Pobs_real.copy()
        if self.use_state_other:
            so = self.state_other.reshape(-1, 4)   # one row (x, y, vx, vy) per neighbour; the reference has one neighbour
            ext = np.stack([so[:, 0] + so[:, 2] * 4 * TS, so[:, 1] + so[:, 3] * 4 * TS])
            obs = np.hstack([obs, ext])
        return (x0, xdes, feet, self.contactInd.copy(), obs)

    def _consume(self, out, b):
        self.mpc_state_e_x_eventbased_ = out["X"][:, b].copy()                 # :431-436
        self.mpc_state_e_u_eventbased_ = out["U"][0::2, b].copy()              # :438-440 (u_x only)
        self.qp_solution_eventbased_ = out["x_qp"][:, b].copy()
        self.alpha_COM_traj_e_ = out["alpha"][:, b].reshape(4, 5).copy()       # :450
        self.mpc_state_alpha_buffer_ = out["alpha_buf_next"][:, b].copy()      # :798
        self.status = int(out["status"][0, b]); self.kkt = float(out["kkt"][0, b])
        self.gaitDomain_ += 1                                                  # :452

    def run_NMPC(self):  # MPC_dist.cpp:81-454
        d = capi.soa_from_instances([self._instance()], alpha_buf=[self.mpc_state_alpha_buffer_])
        out = self.solver.solve_host(d, mode=capi.MODE_QP_NLP if self.use_snopt else capi.MODE_QP)
        self._consume(out, 0)

    # ---- getters (include/MPC_dist.hpp:178-181, MPC_dist.cpp:1272-1276)
    def get_alphaCOM(self): return self.alpha_COM_traj_e_.copy()
    def get_MPCsol(self): return self.mpc_state_e_x_eventbased_.reshape(16, 1).copy()
    def get_MPCsol_u(self): return self.mpc_state_e_u_eventbased_.reshape(4, 1).copy()   # mpc_state_e_u_eventbased_, hpp:63, cpp:438-440 (u_x only)
    def get_lastState(self): return np.array([self.q[0], self.q[1], self.dq[0], self.dq[1]])
    def getDomain(self): return self.domain_


class MPC_dist_team:
    """B agents stepped by one batched launch (what A1_Sim_4_agents.cpp does sequentially, SIM4:857-1065)."""

    def __init__(self, agents, solver=None):
        self.agents = agents
        self.solver = solver or BatchSolver(max_batch=len(agents), max_obs=NUMBER_OF_OBS + len(agents) - 1)
        for a in agents: a.solver = self.solver

    def exchange_states(self):
        """What the simulation loop of the reference does between cycles (state_vec = the other agent's get_lastState()), for a
        whole team: every agent receives the (x, y, vx, vy) of all the others, in team order."""
        last = [a.get_lastState() for a in self.agents]
        for i, a in enumerate(self.agents):
            a.state_other = np.concatenate([last[k] for k in range(len(self.agents)) if k != i])

    def run_NMPC_extended(self, k_obs=1, **cfg):
        """EXTENDED MODE (no reference implementation exists; include/cbfnmpc.h): the cycle with inter-agent rows -- one distance
        row per grid point against each of an agent's k_obs nearest real obstacles and against every team mate's predicted position
        (threshold eps_nbr, dec_vars_constr_cost.h:402), team mates' trajectories exchanged at every interior-point iteration.
        BASELINE config 2.  Teams of 1, 2 or 4 agents in one stance class."""
        insts = []
        for a in self.agents:
            x0, xdes, feet, contact, obs = a._instance()
            order = np.argsort(np.hypot(obs[0] - x0[0], obs[1] - x0[2]), kind="stable")[:k_obs]
            insts.append((x0, xdes, np.asarray(feet)[None], contact, obs[:, order].copy()))
        d = capi.ext_soa_from_instances(insts, 4)
        for b, a in enumerate(self.agents):
            d["alpha_buf"][:, b] = a.mpc_state_alpha_buffer_
        out = self.solver.ext_solve_host(d, N=4, team_size=len(self.agents), **cfg)
        for b, a in enumerate(self.agents): a._consume(out, b)
        return out

    def run_NMPC(self):
        insts = [a._instance() for a in self.agents]
        bufs = [a.mpc_state_alpha_buffer_ for a in self.agents]
        d = capi.soa_from_instances(insts, alpha_buf=bufs)
        modes = {bool(a.use_snopt) for a in self.agents}
        if len(modes) != 1:
            raise ValueError("all agents of one launch must share use_snopt")
        out = self.solver.solve_host(d, mode=capi.MODE_QP_NLP if modes.pop() else capi.MODE_QP)
        for b, a in enumerate(self.agents): a._consume(out, b)
        return out
