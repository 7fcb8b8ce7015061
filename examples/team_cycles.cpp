// team_cycles.cpp -- host-side C++ driver above the C ABI: the call sequence of the reference's 1 kHz loop
// (/root/reference/src/A1_Sim.cpp:227-244: updateState -> run_NMPC -> get_alphaCOM/get_MPCsol) for a 4-agent team
// (/root/reference/src/A1_Sim_4_agents.cpp:857-873 start positions and obstacles), with the LIP state rolled forward
// between cycles by the planned trajectory (the reference rolls it forward in RaiSim, which is not part of this path).
//
//   build:  g++ -std=c++17 -O2 -Iinclude examples/team_cycles.cpp -Lsrb-cbf-nmpc_b200/lib -lcbfnmpc -Wl,-rpath,'$ORIGIN/../srb-cbf-nmpc_b200/lib' -o examples/team_cycles
//   run  :  examples/team_cycles [cycles] [use_snopt 0/1]      prints one line per agent and cycle:
//           cycle agent status kkt slack X[16] alpha[20]
//           examples/team_cycles --extended [cycles] [k_obs]  the team solved with inter-agent rows (extended mode: BASELINE config 2,
//                                                             no reference implementation exists); same output lines as the plain run
//           examples/team_cycles --nlptuple [cycles]          every cycle also solved through the dense nlpTuple boundary
//                                                             (cbfnmpc_pack_dense + cbfnmpc_solve_dense_nlptuple); one JSON line
//           examples/team_cycles --latency [cycles]           one JSON line: p50 / p90 wall time of MPC_dist_gpu::run_NMPC() (one
//           agent, QP -> NLP) and of MPC_dist_gpu_team::run_NMPC() (4 agents in one launch) over `cycles` receding-horizon cycles
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <memory>
#include "mpc_dist_gpu.hpp"

int main(int argc, char** argv) {
  using namespace cbfnmpc;
  const bool latency = argc > 1 && std::strcmp(argv[1], "--latency") == 0;
  const bool check_tuple = argc > 1 && std::strcmp(argv[1], "--nlptuple") == 0;
  const bool extended = argc > 1 && std::strcmp(argv[1], "--extended") == 0;   // BASELINE config 2: inter-agent rows (extended mode)
  const int cycles = (latency || check_tuple || extended) ? (argc > 2 ? atoi(argv[2]) : (latency ? 300 : 4)) : (argc > 1 ? atoi(argv[1]) : 3);
  const bool nlp = (latency || check_tuple || extended) ? true : (argc > 2 ? atoi(argv[2]) != 0 : true);
  const int k_obs_ext = (extended && argc > 3) ? atoi(argv[3]) : 1;
  double tuple_x[4][41] = {{0}}, tuple_worst = 0.0; int tuple_status[4] = {0}, tuple_status_mismatch = 0;
  const bool tuple_four[2] = {true, false};  // stance: 4 feet during the first two cycles, a diagonal pair afterwards
  const int ncols = 4 * (cycles + 1);
  const double Ts = 0.043, goal[2] = {10.0, 0.0};
  const double Pstart[8] = {0.0, 0.0, 0.0, -0.9, -1.0, 0.0, -1.0, -0.9};            // SIM4:862
  const int K = 9;
  const double obs_x[K] = {2.2, 1, 0.85, 1, 1, 3, 3.0, 3, -100}, obs_y[K] = {0.9, 1.2, -0.65, 2, -1.75, 0.2, -0.70, -2.2, 99.5};  // SIM4:872-873
  double Pobs[2 * K];
  for (int k = 0; k < K; k++) { Pobs[2 * k] = obs_x[k]; Pobs[2 * k + 1] = obs_y[k]; }
  // straight-line reference toward the goal at 0.2 m/s (stand-in for generateReferenceTrajectory, MPC_dist.cpp:930-1104)
  std::vector<double> Pr(8 * (size_t)ncols), Prd(8 * (size_t)ncols);
  for (int a = 0; a < 4; a++) {
    double dx = goal[0] - Pstart[2 * a], dy = goal[1] - Pstart[2 * a + 1], n = std::sqrt(dx * dx + dy * dy);
    double vx = 0.2 * dx / n, vy = 0.2 * dy / n;
    for (int j = 0; j < ncols; j++) {
      Pr[(size_t)j * 8 + 2 * a] = Pstart[2 * a] + vx * (j + 1) * Ts; Pr[(size_t)j * 8 + 2 * a + 1] = Pstart[2 * a + 1] + vy * (j + 1) * Ts;
      Prd[(size_t)j * 8 + 2 * a] = vx; Prd[(size_t)j * 8 + 2 * a + 1] = vy;
    }
  }
  try {
    Context ctx(0, 4, K);
    std::vector<std::unique_ptr<MPC_dist_gpu>> agents;
    std::vector<MPC_dist_gpu*> ptrs;
    double q[4][18] = {{0}}, dq[4][18] = {{0}};
    for (int a = 0; a < 4; a++) {
      agents.emplace_back(new MPC_dist_gpu(&ctx));
      MPC_dist_gpu& m = *agents.back();
      m.setAgentID(a); m.setPstart(Pstart); m.setPobs(Pobs, K); m.setPobs_real(Pobs, K);
      m.setReferenceTrajectory(Pr.data(), Prd.data(), ncols);
      m.use_snopt = nlp;
      q[a][0] = Pstart[2 * a]; q[a][1] = Pstart[2 * a + 1];
      ptrs.push_back(&m);
    }
    MPC_dist_gpu_team team(&ctx, ptrs);
    std::vector<double> t_one, t_team;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    for (int c = 0; c < cycles; c++) {
      for (int a = 0; a < 4; a++) {
        int ind[4] = {1, 1, 1, 1};
        if (c >= 2) { bool even = (c % 2) == 0; ind[0] = even; ind[1] = !even; ind[2] = !even; ind[3] = even; }  // trot after stand-up
        double toe[12];
        const double fx[4] = {0.2188, 0.2188, -0.1472, -0.1472}, fy[4] = {-0.1320, 0.1320, -0.1320, 0.1320};
        for (int i = 0; i < 4; i++) { toe[3 * i] = q[a][0] + fx[i]; toe[3 * i + 1] = q[a][1] + fy[i]; toe[3 * i + 2] = 0.0; }
        double other[4];
        agents[a ^ 1]->get_lastState(other);  // pairwise neighbour exchange, shared_structs.hpp:63,94
        agents[a]->updateState(q[a], dq[a], ind, toe, other);
      }
      if (check_tuple) {  // every agent's cycle once more through the dense nlpTuple boundary, before the team launch advances it
        for (int a = 0; a < 4; a++) {
          double xt[41]; tuple_status[a] = agents[a]->solve_via_nlptuple(xt);
          std::memcpy(tuple_x[a], xt, sizeof xt);
        }
      }
      if (latency) {  // agent 0 alone (B = 1) on a copy of its state, then the team launch that advances everybody
        MPC_dist_gpu solo = *agents[0];
        auto t0 = now(); solo.run_NMPC(); auto t1 = now();
        t_one.push_back(ms(t0, t1));
      }
      auto t0 = now();
      if (extended) team.run_NMPC_extended(k_obs_ext); else team.run_NMPC();
      auto t1 = now();
      t_team.push_back(ms(t0, t1));
      for (int a = 0; a < 4; a++) {
        double X[16], al[20];
        agents[a]->get_MPCsol(X); agents[a]->get_alphaCOM(al);
        if (!latency && !check_tuple) {
          printf("%d %d %d %.3e %.17g", c, a, agents[a]->status(), agents[a]->kkt(), agents[a]->slack());
          for (int i = 0; i < 16; i++) printf(" %.17g", X[i]);
          for (int i = 0; i < 20; i++) printf(" %.17g", al[i]);
          printf("\n");
        }
        if (check_tuple) {
          double Uu[8]; agents[a]->get_MPCinputs(Uu);
          double dmax = std::fabs(tuple_x[a][24 + 4 * (tuple_four[c >= 2 ? 1 : 0] ? 4 : 2)] - agents[a]->slack());
          for (int i = 0; i < 16; i++) dmax = std::max(dmax, std::fabs(tuple_x[a][i] - X[i]));
          for (int i = 0; i < 8; i++) dmax = std::max(dmax, std::fabs(tuple_x[a][16 + i] - Uu[i]));
          tuple_worst = std::max(tuple_worst, dmax);
          if (tuple_status[a] != agents[a]->status()) tuple_status_mismatch++;
        }
        q[a][0] = X[12]; dq[a][0] = X[13]; q[a][1] = X[14]; dq[a][1] = X[15];  // end of the planned domain
      }
    }
    if (check_tuple)
      printf("{\"nlptuple_check\": {\"cycles\": %d, \"agents\": 4, \"max_abs_diff_X_U_slack\": %.3e, \"status_mismatches\": %d}}\n", cycles, tuple_worst,
             tuple_status_mismatch);
    if (latency) {
      auto pct = [](std::vector<double> v, double p) {
        if (v.size() > 20) v.erase(v.begin(), v.begin() + 10);  // warm-up cycles
        std::sort(v.begin(), v.end());
        return v[(size_t)(p * (v.size() - 1))];
      };
      printf("{\"host\": \"C++ (include/mpc_dist_gpu.hpp)\", \"cycles\": %d, \"single_solve_p50_ms\": %.4f, \"single_solve_p90_ms\": %.4f, "
             "\"team4_cycle_p50_ms\": %.4f, \"team4_cycle_p90_ms\": %.4f}\n",
             cycles, pct(t_one, 0.5), pct(t_one, 0.9), pct(t_team, 0.5), pct(t_team, 0.9));
    }
  } catch (const Error& e) {
    fprintf(stderr, "cbfnmpc error %d: %s\n", e.code, e.what());
    return e.code == CBFNMPC_ERR_NODEVICE ? 3 : 1;
  }
  return 0;
}
