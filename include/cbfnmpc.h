/* cbfnmpc.h -- C ABI of the B200-native batched CBF-NMPC solver (drop-in for the MPC_dist::run_NMPC hot path).
 *
 * Every entry point replaces an in-process C++ call of the reference (there is no FFI in the reference; these
 * are the calls a maintainer would bind instead -- see INTEGRATION.md):
 *
 *   cbfnmpc_create / destroy      <- MPC_dist::MPC_dist(), ~MPC_dist()          include/MPC_dist.hpp:141-157, src/MPC_dist.cpp:42-77
 *   cbfnmpc_set_gains             <- hard-coded gains / constants               src/MPC_dist.cpp:90-104,164-188,315-318; dec_vars_constr_cost.h:306-307,401
 *   cbfnmpc_set_options           <- iSWIFT GlobalOptions.h:23-26 (QP stage); NLP-stage IPM options (no reference counterpart:
 *                                    the reference never sets an ifopt/SNOPT option, src/MPC_dist.cpp:363-413)
 *   cbfnmpc_solve_batch           <- MPC_dist::updateState + run_NMPC + get_alphaCOM/get_MPCsol, B agents at once, device buffers
 *                                    src/MPC_dist.cpp:1195-1202, 81-454; include/MPC_dist.hpp:178-181
 *   cbfnmpc_solve_batch_host      <- same, host buffers (copies inside): the call LocoWrapper-side code makes   src/A1_Sim.cpp:227-244
 *   cbfnmpc_solve_dense_nlptuple  <- the nlpTuple boundary type                 include/pack_nlp_data.h:11-29, src/MPC_dist.cpp:403-421
 *   cbfnmpc_pack_inputs           <- MPC_dist::updateState + copPlanner_eventbase + footholdsPlanner, B agents at once
 *                                    src/MPC_dist.cpp:1195-1202, 702-782, 1204-1266
 *   cbfnmpc_run_cycle             <- updateState + run_NMPC from the raw per-cycle inputs (pack, then solve)
 *   cbfnmpc_qp_dense_batch        <- iswiftQp_e on a QP without stage structure (the low-level controller's, src/LowLevelCtrl.cpp:18-52),
 *                                    B problems at once                                   optimization/iSWIFT/src/iswift_qp.cpp:44-210
 *   cbfnmpc_generate_reference    <- MPC_dist::generateReferenceTrajectory (offline, once per scenario), T teams at once   src/MPC_dist.cpp:930-1062
 *   cbfnmpc_closed_loop           <- the receding-horizon loop of src/A1_Sim.cpp:983-1266 on the LIP template (no RaiSim): run_NMPC once
 *                                    per gait domain, state rolled to the end of the domain, MPC_dist::updateDistance_to_fail (:21-40)
 *
 * Plain pointers and sizes only.  All floating point is IEEE FP64.  No CPU fallback exists: every solve runs in
 * the CUDA kernels of libcbfnmpc.so, and every call returns a negative code when no CUDA device is usable.
 */
#ifndef CBFNMPC_H
#define CBFNMPC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CBFNMPC_N 4          /* grid points per domain (src/MPC_dist.cpp:92, global_loco_opts.h:9) */
#define CBFNMPC_MAX_ROWS 128 /* bits in the active-set mask (112 one-sided rows at 4 stance feet)     */

/* return codes */
#define CBFNMPC_OK 0
#define CBFNMPC_ERR_ARG (-1)
#define CBFNMPC_ERR_CUDA (-2)
#define CBFNMPC_ERR_NODEVICE (-3)
#define CBFNMPC_ERR_STRUCTURE (-4) /* dense nlpTuple does not have the MPC_dist structure */

/* per-instance status */
#define CBFNMPC_ST_CONVERGED 0
#define CBFNMPC_ST_MAXITER 1
#define CBFNMPC_ST_INFEASIBLE 2  /* line search / multiplier divergence: NLP has no interior */
#define CBFNMPC_ST_NUMERICAL 3   /* inertia correction exhausted */
#define CBFNMPC_ST_BADINPUT 4    /* number of stance feet other than 4 or 2 (src/MPC_dist.cpp:1238-1252; with 2 the reference picks
                                    FL,RR when contactInd[0] == 0 and FR,RL otherwise, whatever the other flags say -- so does this),
                                    or non-finite data (measured state, reference columns, stance feet, closest obstacle; the
                                    reference hands those to its solvers unchecked).  X, U, ... are not written for status 4. */

#define CBFNMPC_MODE_QP 0      /* use_snopt == false: the iSWIFT-equivalent QP optimum is the answer (MPC_dist.hpp:139) */
#define CBFNMPC_MODE_QP_NLP 1  /* use_snopt == true : QP optimum -> NLP (src/MPC_dist.cpp:348 -> 402-427)              */

typedef struct cbfnmpc_ctx cbfnmpc_ctx;

typedef struct {
  double Qw, Rw, Pw, slack_w;        /* src/MPC_dist.cpp:172-175 */
  double mu_fric, h_com, g_acc, Ts;  /* :90, :99-100, :104       */
  double eps_dist, v_max, box;       /* dec_vars_constr_cost.h:401, :306-307; MPC_dist.cpp:317-318 */
} cbfnmpc_gains;

typedef struct {
  double qp_reltol, qp_abstol; int qp_maxit;     /* iSWIFT GlobalOptions.h:23-26 */
  double tol, mu_init, kappa_eps, kappa_mu, tau_min, push;
  double gamma_theta, gamma_phi, eta_phi, kappa_sigma, alpha_min;
  double dw_first, dw_min, kw_minus, kw_plus, kw_plus_first, dw_max, rho_lambda, z_div;
  double sigma_cap, refine_rel, refine_abs;
  int max_iter, filter_cap, max_refine;
  double acc_tol;  /* acceptable KKT error once the barrier parameter has reached its floor tol/10 (Ipopt's acceptable_tol) */
} cbfnmpc_options;

/* Structure-of-arrays batch inputs: field-major, batch index fastest ( a[f * B + b] ).
 *   x0      [4][B]   measured LIP state (q[0], dq[0], q[1], dq[1])                    src/MPC_dist.cpp:226-229
 *   xdes    [16][B]  com_desired_Traj_vec_eventbased_ ([x,xd,y,yd] x 4)               :702-782
 *   feet_xy [8][B]   toePos_(0:1, 0:3): x of FR,FL,RR,RL then y of FR,FL,RR,RL        :1195-1252
 *   contact [4][B]   contactInd (int32)                                               :1198
 *   obs_xy  [2*K][B] Pobs_real: x of the K obstacles then y (closest one is used)     :371-396
 *   alpha_buf [4][B] mpc_state_alpha_buffer_                                          :792-798
 */
typedef struct {
  const double *x0, *xdes, *feet_xy;
  const int32_t *contact;
  const double *obs_xy;
  int n_obs;
  const double *alpha_buf;
} cbfnmpc_in;

/* Outputs (any pointer may be NULL to skip it):
 *   X [16][B]  mpc_state_e_x_eventbased_          U [8][B] COP inputs (u_x,u_y) x 4     src/MPC_dist.cpp:431-440
 *   lam [16][B] support weights (4c used)         slack [B] distance-row slack variable
 *   x_qp [24][B] X,U of the QP stage (qp_solution_eventbased_)                          :350-354
 *   alpha [20][B] alpha_COM_traj_e_ (row-major 4x5)   alpha_buf_next [4][B]              :784-855
 *   status, qp_iters, nlp_iters, n_fact [B] (int32)   kkt [B] final unscaled KKT error
 *   active [2][B] (uint64) bit r set <=> one-sided row r (reference row order) is active (z > s)
 */
typedef struct {
  double *X, *U, *lam, *slack, *x_qp, *alpha, *alpha_buf_next, *kkt;
  int32_t *status, *qp_iters, *nlp_iters, *n_fact;
  uint64_t *active;
} cbfnmpc_out;

int cbfnmpc_version(void);
const char *cbfnmpc_last_error(void);
void cbfnmpc_default_gains(cbfnmpc_gains *g);
void cbfnmpc_default_options(cbfnmpc_options *o);

int cbfnmpc_create(cbfnmpc_ctx **ctx, int device, int max_batch, int max_obs);
void cbfnmpc_destroy(cbfnmpc_ctx *ctx);
int cbfnmpc_set_gains(cbfnmpc_ctx *ctx, const cbfnmpc_gains *g);
int cbfnmpc_set_options(cbfnmpc_ctx *ctx, const cbfnmpc_options *o);

/* device pointers; asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream) */
int cbfnmpc_solve_batch(cbfnmpc_ctx *ctx, int B, int mode, const cbfnmpc_in *in, const cbfnmpc_out *out, void *cuda_stream);
/* host pointers; copies in, solves, copies out, synchronises.  Batches of >= 16384 instances travel as 2 column ranges on
 * separate streams so that copies overlap solves; pinned buffers laid out back to back need one copy per range and direction. */
int cbfnmpc_solve_batch_host(cbfnmpc_ctx *ctx, int B, int mode, const cbfnmpc_in *in, const cbfnmpc_out *out);
/* number of kernel launches issued by this context so far (bench.py's gpu_launches) */
long long cbfnmpc_launch_count(const cbfnmpc_ctx *ctx);
/* last solve kernel's duration in ms measured with CUDA events on the launch stream (synchronises).  A host-buffer call on a
 * large batch launches one kernel per column range (see cbfnmpc_solve_batch_host); the first range's kernel is the one timed. */
int cbfnmpc_last_kernel_ms(cbfnmpc_ctx *ctx, float *ms);
/* on = 0: do not record the event pair around the solve kernel (cbfnmpc_last_kernel_ms then answers for the last launch that was
 * timed).  Without the event between them the solve kernel is launched as a programmatic dependent of the classify pre-pass, which
 * takes its launch latency out of the call (a few microseconds per launch).  Default: on. */
int cbfnmpc_set_kernel_timing(cbfnmpc_ctx *ctx, int on);

/* One instance in the reference's own dense layout (column-major Eigen matrices of pack_nlp_data.h:11-29).
 * The compact form is recovered from the matrices; CBFNMPC_ERR_STRUCTURE if they are not MPC_dist-shaped. */
int cbfnmpc_solve_dense_nlptuple(cbfnmpc_ctx *ctx, int nv, int p, int mi, const double *Q, const double *f,
                                 const double *Aeq, const double *beq, const double *G, const double *h,
                                 const double *x0_start, const double obs[2], int mode_nlp, double *x_out, int *status,
                                 double *kkt);

/* The dense matrices MPC_dist::run_NMPC builds for one agent (src/MPC_dist.cpp:164-321), from the compact inputs and the
 * context's gains: host code, no device work.  F = stance matrix 2 x c (row-major, leading dimension 4), c in {2, 4};
 * outputs column-major: Q nv x nv, f nv, Aeq 28 x nv, beq 28, G (60+8c) x nv, h 60+8c, nv = 25 + 4c.  Lets a host program
 * (or a test) produce an nlpTuple for cbfnmpc_solve_dense_nlptuple without Eigen. */
int cbfnmpc_pack_dense(const cbfnmpc_ctx *ctx, const double x0[4], const double xdes[16], const double *F, int c, double *Q,
                       double *f, double *Aeq, double *beq, double *G, double *h);

/* Raw per-cycle inputs, exactly what MPC_dist::updateState receives plus the object's own state, for B agents.
 * Structure-of-arrays, batch index fastest; device pointers.
 *   q, dq        [nq][B]   generalised coordinates / velocities (nq = 18 in the reference; rows 0 and 1 are read)   MPC_dist.cpp:226-229
 *   contact      [4][B]    contactInd                                                                               :1198
 *   toe_pos      [12][B]   toePos as Eigen::Matrix<double,3,4> column-major: element (r, foot) at row r + 3*foot    :1199
 *   gait_domain  [B]       gaitDomain_ (selects the reference columns and the nominal-stance rule)                   :716, :1211
 *   agent_id     [B]       agent_id_ (rows 2*id, 2*id+1 of the reference trajectory)                                :715-718
 *   team         [B]       index into the per-team arrays below
 *   agent_initial[2][B]    agent_Initial_ (set by setPstart)                                                        :1106-1112
 *   Pr_refined, Prd_refined  [n_teams] matrices 8 x ref_cols, column-major (Eigen::Matrix<double,8,2500>)            MPC_dist.hpp, :930-1104
 *   Pobs_real    [n_teams] matrices 2 x n_obs, column-major                                                         :1124, :371-396
 *   state_other  [4*n_other][B]  optional (NULL / 0 = the reference as shipped).  The reference carries the other agent's
 *                (x, y, vx, vy) into run_NMPC (updateState, :1200; nlpTuple) and its disabled rule (:391) makes that agent's
 *                position extrapolated over the horizon, (x + vx*4*Ts, y + vy*4*Ts), THE obstacle of the distance rows.  With
 *                n_other > 0 the extrapolated positions of the n_other neighbours (rows 4k .. 4k+3 = neighbour k) compete with
 *                the real obstacles in the closest-obstacle rule: they are appended as obstacle columns n_obs .. n_obs+n_other-1.
 * The reference reads columns [4*gait_domain, 4*gait_domain+4) of the trajectory without a bound check; here a column
 * index past ref_cols-1 is clamped to the last column. */
typedef struct {
  const double *q, *dq;
  int nq;
  const int32_t *contact;
  const double *toe_pos;
  const int32_t *gait_domain, *agent_id, *team;
  const double *agent_initial;
  const double *Pr_refined, *Prd_refined;
  int ref_cols, n_teams;
  const double *Pobs_real;
  int n_obs;
  const double *state_other;
  int n_other;
} cbfnmpc_raw_in;

/* fills packed->x0, xdes, feet_xy, contact, obs_xy (device arrays of the sizes cbfnmpc_in documents; packed->n_obs is ignored,
 * raw->n_obs + raw->n_other columns are written); asynchronous on `cuda_stream` */
int cbfnmpc_pack_inputs(cbfnmpc_ctx *ctx, int B, const cbfnmpc_raw_in *raw, const cbfnmpc_in *packed, void *cuda_stream);
/* state_other for agents stored team by team (team_size consecutive agents form a team, B a multiple of team_size): neighbour k
 * of agent a is the k-th OTHER member of its team, rows (q[0], q[1], dq[0], dq[1]) of that member.  state_other
 * [4*(team_size-1)][B], device.  (Agents of a team spread over several GPUs exchange the same rows with an all-gather, see
 * srb-cbf-nmpc_b200/sharding.py.) */
int cbfnmpc_team_states(cbfnmpc_ctx *ctx, int B, int team_size, const double *q, const double *dq, double *state_other,
                        void *cuda_stream);
/* pack into the context's own staging buffers (B <= max_batch, raw->n_obs + raw->n_other <= max_obs), then solve; alpha_buf
 * [4][B] device */
int cbfnmpc_run_cycle(cbfnmpc_ctx *ctx, int B, int mode, const cbfnmpc_raw_in *raw, const double *alpha_buf,
                      const cbfnmpc_out *out, void *cuda_stream);

/* Closed-loop worlds.  One world = one agent whose raw inputs are the arrays of a cbfnmpc_raw_in; after every
 * cycle the world is advanced ON THE DEVICE: state <- x_4 of the solution (the end of the gait domain under perfect tracking
 * of the template), gait_domain += 1, alpha_buf <- alpha_buf_next (MPC_dist.cpp:798, :452), stance feet = nominal footprint
 * about the new position, contact pattern = 4 feet during the first two domains, then the alternating diagonal pairs of
 * the trot; success/distance_to_fail follow MPC_dist::updateDistance_to_fail (MPC_dist.cpp:21-40) evaluated at the 4 grid
 * points of the domain: the first time an obstacle is closer than fail_radius (0.5 in the reference) the world fails and
 * records its distance from the origin.  The solver status is recorded but, as in the reference, not acted upon.
 * All pointers are device pointers; q, dq, contact, toe_pos, gait_domain are the SAME arrays as in `raw` (updated in place). */
typedef struct {
  double *q, *dq;              /* [nq][B] */
  int32_t *contact;            /* [4][B] */
  double *toe_pos;             /* [12][B] */
  int32_t *gait_domain;        /* [B] */
  double *alpha_buf;           /* [4][B] */
  int32_t *success;            /* [B]  1 = no obstacle came within fail_radius so far (initialise to 1) */
  double *distance_to_fail;    /* [B]  (initialise to the reference's default 10) */
  int32_t *n_not_converged;    /* [B]  cycles whose solver status was not CBFNMPC_ST_CONVERGED (initialise to 0) */
  double *traj;                /* optional [n_cycles][4][B]: (x, xd, y, yd) at the end of every cycle */
} cbfnmpc_world_state;
int cbfnmpc_closed_loop(cbfnmpc_ctx *ctx, int B, int n_cycles, int mode, const cbfnmpc_raw_in *raw, const cbfnmpc_world_state *w,
                        double fail_radius, void *cuda_stream);

/* generateReferenceTrajectory for n_teams teams of 4 agents: potential field toward (goal_x, goal_y), obstacle
 * repulsion (the PLANNING obstacles of setPobs), Lennard-Jones forces between team mates; `loop` steps (100000 in the reference),
 * decimated by 40.  Device pointers:
 *   Pstart [n_teams][8]   (x, y of the 4 agents, as setPstart receives it)
 *   Pobs   [n_teams] matrices 2 x n_obs column-major
 *   Pr_refined, Prd_refined (out) [n_teams] matrices 8 x (loop / 40) column-major -- the layout cbfnmpc_raw_in expects.
 * The reference's column quirk is kept: the last two output columns are the undecimated steps loop/40 and loop/40 + 1. */
int cbfnmpc_generate_reference(cbfnmpc_ctx *ctx, int n_teams, const double *Pstart, const double *Pobs, int n_obs, double goal_x,
                               double goal_y, int loop, double *Pr_refined, double *Prd_refined, void *cuda_stream);

/* B dense convex QPs of identical dimensions, iSWIFT's algorithm and the context's qp_* options, one warp per problem:
 *     min 1/2 x'Px + c'x   s.t.  A x = b,  G x <= h          n <= 32, p <= 24 (may be 0), 1 <= m <= 64
 * Device pointers; problem q's data is contiguous: P + q n n (n x n), A + q p n (p x n), G + q m n (m x n), all column-major
 * like the Eigen matrices iswiftQp_e receives; c + q n, b + q p, h + q m.  Outputs x [B][n]; y [B][p], z [B][m] optional (NULL);
 * status [B]: 0 converged, 1 factorisation failed (P + G'WG not positive definite), 2 iteration limit; iters [B]. */
int cbfnmpc_qp_dense_batch(cbfnmpc_ctx *ctx, int B, int n, int p, int m, const double *P, const double *c, const double *A,
                           const double *b, const double *G, const double *h, double *x, double *y, double *z, int32_t *status,
                           int32_t *iters, void *cuda_stream);

/* ---------------------------------------------------------------------------------------------------------------------
 * EXTENDED MODE (BASELINE configs 4 and 5).  NO REFERENCE IMPLEMENTATION EXISTS: the reference's NLP has
 * N = 4 grid points, one obstacle and no inter-agent rows (src/MPC_dist.cpp:81-454).  Generalised along the reference's own hints:
 *   N = 4 or 8 grid points (N / 4 gait domains, one stance matrix per domain);
 *   n_obs obstacle rows per grid point   |p_k - o_i|^2   + delta >= eps_obs   (the row of dec_vars_constr_cost.h:262-265, for EVERY
 *     obstacle handed in -- the caller pre-selects, e.g. the 16 nearest);
 *   neighbour rows per grid point        |p_k - p^j_k|^2 + delta >= eps_nbr   (DV:402 "epsilon = 2.2 // for robot to robot"; the
 *     reference's disabled rule MPC_dist.cpp:391 extrapolates the neighbour's current state instead);
 *   teams: team_size consecutive instances form a team whose members solve together and read each other's published trajectories
 *     at the top of every interior-point iteration (under-relaxed Jacobi: pub <- pub + relax (iterate - pub); finished when nobody
 *     stepped and every |iterate - pub| <= pub_tol).  The members of a team must share the stance class (4 feet / diagonal pair).
 * With N = 4, n_obs = 1, no neighbours the answers are those of cbfnmpc_solve_batch given the closest obstacle.
 * Arrays are structure-of-arrays, batch index fastest, like cbfnmpc_in / cbfnmpc_out:
 *   x0 [4][B]; xdes [4N][B]; feet_xy [8 N/4][B] (per gait domain: x of the 4 feet, then y); contact [4][B];
 *   obs_xy [2 n_obs][B] (x of all rows, then y); alpha_buf [4][B] optional;
 *   nbr_traj [n_nbr][2N][B] optional, team_size 1 only: FIXED neighbour trajectories (rows 2k, 2k+1 = x, y at grid point k)
 * Outputs (any may be NULL): X [4N][B], U [2N][B], lam [4N][B], slack [B], x_qp [6N][B], alpha [20][B] (Bezier fit of the first gait
 * domain), alpha_buf_next [4][B], kkt [B], pub [2N][B] (published trajectory at the end), status / qp_iters / nlp_iters / team_iters [B]. */
typedef struct {
  int N, team_size;
  double eps_obs, eps_nbr; /* defaults (double)1.9f, (double)2.2f */
  double relax, pub_tol;   /* defaults 0.7, 1e-10 */
  int max_team_iters;      /* default 64: iterations of a team's loop (99 % of the config-4 teams settle within 38, 99.9 % within 60) */
  int mu_stall_iters;      /* default 6: team mode, iterations without a barrier decrease after which it is lowered anyway */
} cbfnmpc_ext_cfg;
typedef struct {
  const double *x0, *xdes, *feet_xy;
  const int32_t *contact;
  const double *obs_xy;
  int n_obs;
  const double *alpha_buf;
  const double *nbr_traj;
  int n_nbr;
} cbfnmpc_ext_in;
typedef struct {
  double *X, *U, *lam, *slack, *x_qp, *alpha, *alpha_buf_next, *kkt, *pub;
  int32_t *status, *qp_iters, *nlp_iters, *team_iters;
} cbfnmpc_ext_out;
void cbfnmpc_ext_default_cfg(cbfnmpc_ext_cfg *cfg);
/* device pointers, asynchronous on `cuda_stream`; a team (team_size x N lanes) lives in one warp and exchanges by shuffles */
int cbfnmpc_ext_solve_batch(cbfnmpc_ctx *ctx, int B, const cbfnmpc_ext_cfg *cfg, const cbfnmpc_ext_in *in, const cbfnmpc_ext_out *out,
                            void *cuda_stream);
/* host pointers: copies in, solves, copies out, synchronises */
int cbfnmpc_ext_solve_batch_host(cbfnmpc_ctx *ctx, int B, const cbfnmpc_ext_cfg *cfg, const cbfnmpc_ext_in *in, const cbfnmpc_ext_out *out);

/* AGENT-SHARDED teams (BASELINE config 4 as worded: the members of a team live on different GPUs; SURVEY.md 8e collective C1).
 * cbfnmpc_comm_unique_id / cbfnmpc_comm_init create the library's own NCCL communicator over the ranks that share teams (the id,
 * 128 bytes, is made by one rank and handed to the others by the host program's own means).  cbfnmpc_ext_solve_sharded: this rank
 * holds members [rank m, (rank + 1) m) of EVERY one of n_teams teams, m = team_size / nranks, member-major (instance b =
 * local member * n_teams + team, B_local = m n_teams; arrays as above with B = B_local, device pointers, nbr_traj unused).  One kernel
 * launch per interior-point iteration, an ncclAllGather of every member's published trajectory ((2N + 2) n_teams doubles each) in
 * between; without a communicator (nranks = 1) the same path runs on one GPU with a device copy instead of the all-gather.
 * Answers equal those of cbfnmpc_ext_solve_batch on the team-major layout. */
typedef struct {
  int launches, iterations, exchanges, nranks;
  long long bytes_per_exchange; /* contributed by this rank per all-gather */
  float total_ms;               /* device time of the whole call (CUDA events) */
  int fused;                    /* 1: the exchange ran inside one launch over peer-memory mailboxes (cbfnmpc_ext_solve_fused) */
} cbfnmpc_ext_stats;
int cbfnmpc_comm_unique_id(void *id128);
int cbfnmpc_comm_init(cbfnmpc_ctx *ctx, int nranks, int rank, const void *id128);
int cbfnmpc_comm_destroy(cbfnmpc_ctx *ctx);
int cbfnmpc_ext_solve_sharded(cbfnmpc_ctx *ctx, int n_teams, const cbfnmpc_ext_cfg *cfg, const cbfnmpc_ext_in *in,
                              const cbfnmpc_ext_out *out, cbfnmpc_ext_stats *stats, void *cuda_stream);
/* The same teams, same arguments and same answers, with the exchange INSIDE one launch: every member stores its published block
 * straight into mailboxes in the memory of the ranks that hold its team mates (peer memory over NVLink, mapped once through CUDA
 * IPC) and waits on per-(member, team) sequence flags in its own -- no launch and no collective between the interior-point
 * iterations.  Collective over the communicator like cbfnmpc_ext_solve_sharded; synchronises the stream before it returns.  A
 * launch must be resident at once (its members wait for each other): batches beyond that (about 4700 instances of config 4 per
 * B200) run as a few launches over slices of the teams.  Needs peer access between the ranks; where that is missing it runs
 * cbfnmpc_ext_solve_sharded instead (stats->fused tells which). */
int cbfnmpc_ext_solve_fused(cbfnmpc_ctx *ctx, int n_teams, const cbfnmpc_ext_cfg *cfg, const cbfnmpc_ext_in *in,
                            const cbfnmpc_ext_out *out, cbfnmpc_ext_stats *stats, void *cuda_stream);

/* FP64 DFMA peak microbenchmark (TFLOP/s) used as the roofline denominator by bench.py */
int cbfnmpc_measure_fp64_peak(cbfnmpc_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
