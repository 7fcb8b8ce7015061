// mpc_dist_gpu.hpp -- header-only C++17 host side above the C ABI of cbfnmpc.h.
//
// Mirrors the hot-path part of the reference's per-agent planner class `MPC_dist`
// (/root/reference/include/MPC_dist.hpp:136-189, src/MPC_dist.cpp): the same method names, argument meaning and
// hidden state (gaitDomain_, domain_, mpc_state_alpha_buffer_, agent_Initial_), Eigen-free: every matrix argument
// is the `.data()` pointer of the Eigen object the reference passes (Eigen's default storage is column-major).
// Every solve is executed by libcbfnmpc.so on the GPU; nothing here computes a solution on the CPU.
//
//   reference call (file:line)                                     here
//   MPC_dist::setAgentID                 MPC_dist.cpp:1134         MPC_dist_gpu::setAgentID
//   MPC_dist::setPstart                  :1106-1117                setPstart(const double Pstart[2*NUMBER_OF_AGENTS])
//   MPC_dist::setPobs / setPobs_real     :1119-1132                setPobs / setPobs_real(const double* 2xK col-major)
//   MPC_dist::generateReferenceTrajectory :930-1104 (off the path) setReferenceTrajectory(Pr, Prd)  (8 x ncols, col-major)
//   MPC_dist::updateState                :1195-1202                updateState(q, dq, ind, toePos 3x4 col-major, state_other)
//   MPC_dist::run_NMPC                   :81-454                   run_NMPC()  (returns the per-instance status; the reference returns void)
//   get_alphaCOM / get_MPCsol            MPC_dist.hpp:178-179      get_alphaCOM(double out[20] col-major 4x5) / get_MPCsol(double out[16])
//   get_lastState / getDomain            :180-181, cpp:1272-1276   get_lastState(double out[4]) / getDomain()
//   bool use_snopt                       MPC_dist.hpp:139          use_snopt
//
// `MPC_dist_gpu_team` runs run_NMPC() for many agents with ONE batched launch (what src/A1_Sim_4_agents.cpp:1062-1065
// does sequentially).
#pragma once
#include <cstdint>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "cbfnmpc.h"

namespace cbfnmpc {

constexpr int NDOMAIN = 4;           // include/global_loco_opts.h:9
constexpr int NUMBER_OF_AGENTS = 4;  // :14
constexpr int NUMBER_OF_OBS = 20;    // :15

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// RAII owner of one cbfnmpc_ctx (one per thread and device, like one MPC_dist per agent thread in the reference).
class Context {
 public:
  Context(int device, int max_batch, int max_obs = NUMBER_OF_OBS) : max_batch_(max_batch), max_obs_(max_obs) {
    int rc = cbfnmpc_create(&ctx_, device, max_batch, max_obs);
    if (rc != CBFNMPC_OK) throw Error(rc, std::string("cbfnmpc_create: ") + cbfnmpc_last_error());
  }
  ~Context() { cbfnmpc_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  cbfnmpc_ctx* get() const { return ctx_; }
  int max_batch() const { return max_batch_; }
  int max_obs() const { return max_obs_; }

 private:
  cbfnmpc_ctx* ctx_ = nullptr;
  int max_batch_, max_obs_;
};

// Host-side SoA staging block for B agents (field-major, batch index fastest: a[f * B + b]).
struct Batch {
  int B, K;
  std::vector<double> x0, xdes, feet, obs, abuf;
  std::vector<int32_t> contact;
  std::vector<double> X, U, lam, slack, xqp, alpha, abuf_next, kkt;
  std::vector<int32_t> status, qp_iters, nlp_iters, n_fact;
  std::vector<uint64_t> active;
  Batch(int B_, int K_)
      : B(B_), K(K_), x0(4 * B_), xdes(16 * B_), feet(8 * B_), obs(2 * K_ * B_), abuf(4 * B_), contact(4 * B_), X(16 * B_), U(8 * B_),
        lam(16 * B_), slack(B_), xqp(24 * B_), alpha(20 * B_), abuf_next(4 * B_), kkt(B_), status(B_), qp_iters(B_), nlp_iters(B_),
        n_fact(B_), active(2 * B_) {}
  cbfnmpc_in in() const { return cbfnmpc_in{x0.data(), xdes.data(), feet.data(), contact.data(), obs.data(), K, abuf.data()}; }
  cbfnmpc_out out() {
    return cbfnmpc_out{X.data(), U.data(), lam.data(), slack.data(), xqp.data(), alpha.data(), abuf_next.data(), kkt.data(),
                       status.data(), qp_iters.data(), nlp_iters.data(), n_fact.data(), active.data()};
  }
};

class MPC_dist_gpu {
 public:
  bool use_snopt = false;  // MPC_dist.hpp:139
  // The reference's disabled inter-agent rule (MPC_dist.cpp:391): the other agent, extrapolated over the horizon
  // (x + vx*4*Ts, y + vy*4*Ts), competes with the real obstacles for the distance rows.  false = the reference as shipped
  // (state_other is carried and never read).
  bool use_state_other = false;

  // `ctx` may be shared by several agents that are stepped from the same thread (or by a team, below).
  explicit MPC_dist_gpu(Context* ctx) : ctx_(ctx) {
    std::memset(q_, 0, sizeof q_); std::memset(dq_, 0, sizeof dq_);
    for (int i = 0; i < 4; i++) contactInd_[i] = 1;
  }

  void setAgentID(size_t agent_id) { agent_id_ = agent_id; }  // MPC_dist.cpp:1134
  void setPstart(const double* Pstart /* 2*NUMBER_OF_AGENTS */) {  // :1106-1117
    std::memcpy(Pstart_, Pstart, sizeof Pstart_);
    agent_Initial_[0] = Pstart_[2 * agent_id_]; agent_Initial_[1] = Pstart_[2 * agent_id_ + 1];
    alpha_buf_[0] = agent_Initial_[0]; alpha_buf_[1] = 0.0; alpha_buf_[2] = agent_Initial_[1]; alpha_buf_[3] = 0.0;  // :1111
  }
  void setPobs(const double* Pobs /* 2 x n, column-major */, int n = NUMBER_OF_OBS) { Pobs_.assign(Pobs, Pobs + 2 * n); }
  void setPobs_real(const double* Pobs, int n = NUMBER_OF_OBS) { Pobs_real_.assign(Pobs, Pobs + 2 * n); n_obs_ = n; }
  // Pr_refined_ / Prd_refined_ (2*NUMBER_OF_AGENTS x ncols, column-major), the output of generateReferenceTrajectory()
  void setReferenceTrajectory(const double* Pr, const double* Prd, int ncols) {
    Pr_.assign(Pr, Pr + 2 * NUMBER_OF_AGENTS * (size_t)ncols); Prd_.assign(Prd, Prd + 2 * NUMBER_OF_AGENTS * (size_t)ncols); ref_cols_ = ncols;
  }

  void updateState(const double* qin, const double* dqin, const int* ind, const double* toePos /* 3x4 col-major */,
                   const double* state_vec /* 4 */) {  // MPC_dist.cpp:1195-1202
    std::memcpy(q_, qin, 18 * sizeof(double)); std::memcpy(dq_, dqin, 18 * sizeof(double));
    std::memcpy(contactInd_, ind, 4 * sizeof(int));
    std::memcpy(toePos_, toePos, 12 * sizeof(double));
    std::memcpy(state_other_, state_vec, 4 * sizeof(double));
  }

  // One control cycle of this agent alone (B = 1).  Returns the per-instance status (CBFNMPC_ST_*).
  int run_NMPC() {
    Batch b(1, n_obs());
    pack(b, 0);
    cbfnmpc_in in = b.in(); cbfnmpc_out out = b.out();
    int rc = cbfnmpc_solve_batch_host(ctx_->get(), 1, use_snopt ? CBFNMPC_MODE_QP_NLP : CBFNMPC_MODE_QP, &in, &out);
    if (rc != CBFNMPC_OK) throw Error(rc, std::string("cbfnmpc_solve_batch_host: ") + cbfnmpc_last_error());
    consume(b, 0);
    return status_;
  }

  // The same cycle through the reference's dense boundary type (nlpTuple, pack_nlp_data.h:11-29; MPC_dist.cpp:348 and :403-421):
  // dense Q, f, Aeq, beq, G, h built as MPC_dist.cpp:164-321 does, the QP solved from them (iswiftQp_e's role), the closest
  // obstacle chosen (:371-396), then the NLP started at the QP optimum.  x_out = the decision vector [X16 | U8 | Lambda 4c | delta].
  // Does not touch the object's state (no gaitDomain_ increment, no Bezier fit): a cross-check of run_NMPC(), not a second path.
  int solve_via_nlptuple(double* x_out /* 25 + 4c */, double* kkt_out = nullptr) {
    Batch b(1, n_obs());
    const int gd = gaitDomain_;
    double toe_keep[12]; std::memcpy(toe_keep, toePos_, sizeof toe_keep);
    pack(b, 0);
    std::memcpy(toePos_, toe_keep, sizeof toe_keep); gaitDomain_ = gd;
    int csum = 0; for (int i = 0; i < 4; i++) csum += b.contact[i];
    const int c = csum == 4 ? 4 : 2, nv = 25 + 4 * c, p = 28, mi = 60 + 8 * c;
    double F[8] = {0};
    for (int a = 0; a < 2; a++) {
      if (c == 4) { for (int i = 0; i < 4; i++) F[4 * a + i] = b.feet[4 * a + i]; }
      else if (b.contact[0] == 0) { F[4 * a] = b.feet[4 * a + 1]; F[4 * a + 1] = b.feet[4 * a + 2]; }   // MPC_dist.cpp:1238-1252
      else { F[4 * a] = b.feet[4 * a]; F[4 * a + 1] = b.feet[4 * a + 3]; }
    }
    std::vector<double> Q((size_t)nv * nv), f(nv), A((size_t)p * nv), be(p), G((size_t)mi * nv), h(mi), xq(nv), x0s(nv);
    int rc = cbfnmpc_pack_dense(ctx_->get(), b.x0.data(), b.xdes.data(), F, c, Q.data(), f.data(), A.data(), be.data(), G.data(), h.data());
    if (rc != CBFNMPC_OK) throw Error(rc, std::string("cbfnmpc_pack_dense: ") + cbfnmpc_last_error());
    int st = -1; double kk = 0.0;
    rc = cbfnmpc_solve_dense_nlptuple(ctx_->get(), nv, p, mi, Q.data(), f.data(), A.data(), be.data(), G.data(), h.data(), nullptr, nullptr, 0,
                                      xq.data(), &st, &kk);
    if (rc != CBFNMPC_OK) throw Error(rc, std::string("cbfnmpc_solve_dense_nlptuple (QP): ") + cbfnmpc_last_error());
    if (use_snopt) {
      double best = 1000.0, obs[2] = {b.obs[0], b.obs[b.K]};   // first strict minimum of the Euclidean distance, :371-396
      for (int k = 0; k < b.K; k++) {
        const double dx = b.x0[0] - b.obs[k], dy = b.x0[2] - b.obs[b.K + k], d = std::sqrt(dx * dx + dy * dy);
        if (d < best) { best = d; obs[0] = b.obs[k]; obs[1] = b.obs[b.K + k]; }
      }
      x0s = xq;
      rc = cbfnmpc_solve_dense_nlptuple(ctx_->get(), nv, p, mi, Q.data(), f.data(), A.data(), be.data(), G.data(), h.data(), x0s.data(), obs, 1,
                                        xq.data(), &st, &kk);
      if (rc != CBFNMPC_OK) throw Error(rc, std::string("cbfnmpc_solve_dense_nlptuple (NLP): ") + cbfnmpc_last_error());
    }
    std::memcpy(x_out, xq.data(), sizeof(double) * nv);
    if (kkt_out) *kkt_out = kk;
    return st;
  }

  void get_alphaCOM(double out[20]) const {  // 4x5, column-major like the Eigen::MatrixXd the reference returns
    for (int r = 0; r < 4; r++) for (int j = 0; j < 5; j++) out[4 * j + r] = alpha_[5 * r + j];
  }
  void get_MPCsol(double out[16]) const { std::memcpy(out, X_, sizeof X_); }
  void get_MPCinputs(double out[8]) const { std::memcpy(out, U_, sizeof U_); }
  // mpc_state_e_u_eventbased_ (MPC_dist.hpp:63, cpp:438-440): the x-components of the 4 centre-of-pressure inputs only
  void get_MPCsol_u(double out[4]) const { for (int k = 0; k < 4; k++) out[k] = U_[2 * k]; }
  void get_QPsol(double out[24]) const { std::memcpy(out, xqp_, sizeof xqp_); }
  void get_lastState(double out[4]) const { out[0] = q_[0]; out[1] = q_[1]; out[2] = dq_[0]; out[3] = dq_[1]; }  // :1272-1276
  int getDomain() const { return domain_; }
  int getGaitDomain() const { return gaitDomain_; }
  int status() const { return status_; }
  double kkt() const { return kkt_; }
  double slack() const { return slack_; }

  // ---- used by MPC_dist_gpu_team
  int n_obs() const { return n_obs_ + (use_state_other ? 1 : 0); }  // obstacle columns of this agent's instances
  // copPlanner_eventbase (:702-782) + footholdsPlanner (:1204-1266) + state (:226-229) -> column `b` of a batch
  void pack(Batch& bt, int b) {
    const int B = bt.B, N = 4;
    if (Pr_.empty() || 4 * (gaitDomain_ + 1) > ref_cols_) throw Error(CBFNMPC_ERR_ARG, "reference trajectory missing or exhausted");
    if (n_obs() != bt.K || (int)Pobs_real_.size() != 2 * n_obs_) throw Error(CBFNMPC_ERR_ARG, "obstacle count differs from the batch");
    const size_t rows = 2 * NUMBER_OF_AGENTS;
    for (int k = 0; k < N; k++) {
      const size_t col = (size_t)N * gaitDomain_ + k;
      bt.xdes[(4 * k + 0) * B + b] = Pr_[col * rows + 2 * agent_id_];
      bt.xdes[(4 * k + 1) * B + b] = Prd_[col * rows + 2 * agent_id_];
      bt.xdes[(4 * k + 2) * B + b] = Pr_[col * rows + 2 * agent_id_ + 1];
      bt.xdes[(4 * k + 3) * B + b] = Prd_[col * rows + 2 * agent_id_ + 1];
    }
    if (gaitDomain_ <= 1) {  // nominal stance before and during stand-up, :1211-1216
      static const double fx[4] = {0.2188, 0.2188, -0.1472, -0.1472}, fy[4] = {-0.1320, 0.1320, -0.1320, 0.1320};
      for (int i = 0; i < 4; i++) { toePos_[3 * i] = fx[i] + agent_Initial_[0]; toePos_[3 * i + 1] = fy[i] + agent_Initial_[1]; toePos_[3 * i + 2] = 0.0; }
    }
    for (int i = 0; i < 4; i++) {
      bt.feet[(size_t)i * B + b] = toePos_[3 * i];
      bt.feet[(size_t)(4 + i) * B + b] = toePos_[3 * i + 1];
      bt.contact[(size_t)i * B + b] = contactInd_[i];
    }
    bt.x0[0 * B + b] = q_[0]; bt.x0[1 * B + b] = dq_[0]; bt.x0[2 * B + b] = q_[1]; bt.x0[3 * B + b] = dq_[1];  // :226-229
    for (int k = 0; k < n_obs_; k++) { bt.obs[(size_t)k * B + b] = Pobs_real_[2 * k]; bt.obs[(size_t)(bt.K + k) * B + b] = Pobs_real_[2 * k + 1]; }
    if (use_state_other) {  // :391
      const double Ts = 43 * 0.001;  // TSOPTTICK * 0.001
      bt.obs[(size_t)n_obs_ * B + b] = state_other_[0] + (state_other_[2] * 4.0) * Ts;
      bt.obs[(size_t)(bt.K + n_obs_) * B + b] = state_other_[1] + (state_other_[3] * 4.0) * Ts;
    }
    for (int i = 0; i < 4; i++) bt.abuf[(size_t)i * B + b] = alpha_buf_[i];
  }
  void consume(const Batch& bt, int b) {
    const int B = bt.B;
    for (int i = 0; i < 16; i++) X_[i] = bt.X[(size_t)i * B + b];                 // :431-436
    for (int i = 0; i < 8; i++) U_[i] = bt.U[(size_t)i * B + b];                  // :438-440
    for (int i = 0; i < 24; i++) xqp_[i] = bt.xqp[(size_t)i * B + b];             // :350-354
    for (int i = 0; i < 20; i++) alpha_[i] = bt.alpha[(size_t)i * B + b];         // :450
    for (int i = 0; i < 4; i++) alpha_buf_[i] = bt.abuf_next[(size_t)i * B + b];  // :798
    status_ = bt.status[b]; kkt_ = bt.kkt[b]; slack_ = bt.slack[b];
    gaitDomain_++;                                                                // :452
  }

 private:
  Context* ctx_;
  size_t agent_id_ = 0;
  int gaitDomain_ = 0, domain_ = 0, n_obs_ = NUMBER_OF_OBS, ref_cols_ = 0, status_ = -1;
  double q_[18], dq_[18], toePos_[12] = {0}, state_other_[4] = {0};
  int contactInd_[4];
  double Pstart_[2 * NUMBER_OF_AGENTS] = {0}, agent_Initial_[2] = {0}, alpha_buf_[4] = {0};
  std::vector<double> Pobs_, Pobs_real_, Pr_, Prd_;
  double X_[16] = {0}, U_[8] = {0}, xqp_[24] = {0}, alpha_[20] = {0}, kkt_ = 0.0, slack_ = 0.0;
};

// All agents of a team (or of many teams) in one launch.
class MPC_dist_gpu_team {
 public:
  MPC_dist_gpu_team(Context* ctx, std::vector<MPC_dist_gpu*> agents) : ctx_(ctx), agents_(std::move(agents)) {
    if (agents_.empty() || (int)agents_.size() > ctx->max_batch()) throw Error(CBFNMPC_ERR_ARG, "team size");
  }
  void run_NMPC() {
    const int B = (int)agents_.size();
    Batch bt(B, agents_[0]->n_obs());
    for (int b = 0; b < B; b++) {
      if (agents_[b]->use_snopt != agents_[0]->use_snopt) throw Error(CBFNMPC_ERR_ARG, "all agents of one launch must share use_snopt");
      agents_[b]->pack(bt, b);
    }
    cbfnmpc_in in = bt.in(); cbfnmpc_out out = bt.out();
    int rc = cbfnmpc_solve_batch_host(ctx_->get(), B, agents_[0]->use_snopt ? CBFNMPC_MODE_QP_NLP : CBFNMPC_MODE_QP, &in, &out);
    if (rc != CBFNMPC_OK) throw Error(rc, std::string("cbfnmpc_solve_batch_host: ") + cbfnmpc_last_error());
    for (int b = 0; b < B; b++) agents_[b]->consume(bt, b);
  }

  // EXTENDED MODE (include/cbfnmpc.h; SURVEY.md A.6 -- no reference implementation exists): the same cycle with inter-agent rows.
  // Every agent gets one distance row per grid point against each of its k_obs nearest real obstacles and against every team
  // mate's predicted position at that grid point (threshold eps_nbr: dec_vars_constr_cost.h:402 "epsilon = 2.2 // for robot to
  // robot"); the team mates' trajectories are exchanged at every interior-point iteration inside the solve.  This is BASELINE
  // config 2 ("4-agent team with inter-agent CBFs, one cycle per agent").  The team must have 1, 2 or 4 agents in one stance class.
  void run_NMPC_extended(int k_obs = 1, const cbfnmpc_ext_cfg* cfg_in = nullptr) {
    const int B = (int)agents_.size(), N = 4;
    if (B != 1 && B != 2 && B != 4) throw Error(CBFNMPC_ERR_ARG, "extended mode: teams of 1, 2 or 4 agents");
    Batch bt(B, agents_[0]->n_obs());
    for (int b = 0; b < B; b++) agents_[b]->pack(bt, b);
    const int K = bt.K < k_obs ? bt.K : k_obs;
    std::vector<double> obs(2 * (size_t)K * B);
    for (int b = 0; b < B; b++) {   // the K nearest obstacles of agent b, nearest first (ties: lower index)
      std::vector<int> idx(bt.K);
      for (int k = 0; k < bt.K; k++) idx[k] = k;
      const double px = bt.x0[b], py = bt.x0[2 * (size_t)B + b];
      std::stable_sort(idx.begin(), idx.end(), [&](int i, int j) {
        const double di = std::hypot(bt.obs[(size_t)i * B + b] - px, bt.obs[(size_t)(bt.K + i) * B + b] - py);
        const double dj = std::hypot(bt.obs[(size_t)j * B + b] - px, bt.obs[(size_t)(bt.K + j) * B + b] - py);
        return di < dj;
      });
      for (int k = 0; k < K; k++) { obs[(size_t)k * B + b] = bt.obs[(size_t)idx[k] * B + b]; obs[(size_t)(K + k) * B + b] = bt.obs[(size_t)(bt.K + idx[k]) * B + b]; }
    }
    cbfnmpc_ext_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else cbfnmpc_ext_default_cfg(&cfg);
    cfg.N = N; cfg.team_size = B;
    cbfnmpc_ext_in in = {bt.x0.data(), bt.xdes.data(), bt.feet.data(), bt.contact.data(), obs.data(), K, bt.abuf.data(), nullptr, 0};
    std::vector<int32_t> team_iters(B);
    cbfnmpc_ext_out out = {bt.X.data(), bt.U.data(), bt.lam.data(), bt.slack.data(), bt.xqp.data(), bt.alpha.data(), bt.abuf_next.data(), bt.kkt.data(), nullptr,
                           bt.status.data(), bt.qp_iters.data(), bt.nlp_iters.data(), team_iters.data()};
    int rc = cbfnmpc_ext_solve_batch_host(ctx_->get(), B, &cfg, &in, &out);
    if (rc != CBFNMPC_OK) throw Error(rc, std::string("cbfnmpc_ext_solve_batch_host: ") + cbfnmpc_last_error());
    for (int b = 0; b < B; b++) agents_[b]->consume(bt, b);
    team_iters_ = team_iters[0];
  }
  int team_iters() const { return team_iters_; }

 private:
  Context* ctx_;
  std::vector<MPC_dist_gpu*> agents_;
  int team_iters_ = 0;
};

}  // namespace cbfnmpc
