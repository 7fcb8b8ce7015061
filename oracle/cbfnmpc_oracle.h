/* TEST INFRASTRUCTURE -- CPU oracle for the CBF-NMPC hot path (MPC_dist::run_NMPC).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this.
 * The product path (srb-cbf-nmpc_b200/) never links, imports or calls it.
 *
 * Pinning status:
 *   - QP stage  (orc_qp_mehrotra): restates iSWIFT (optimization/iSWIFT/src/Prime.c:127-230, Auxilary.c);
 *     pinned against the reference's own C core compiled in place (oracle/_ref/libiswift_ref.so) on the
 *     reference's canned QP (include/Matrices_small.h) and on restated MPC QPs -- tests/test_oracle_qp.py.
 *   - NLP stage (orc_nlp_ipm): **weakly pinned** (round 2; "parity unpinned" in the strict sense: no runnable reference
 *     solver).  The reference solves it with ifopt + SNOPT 7.7.7 (src/MPC_dist.cpp:363-421), neither of which is in
 *     /root/reference or installable offline.  The one NLP output the reference tree holds, the SNOPT print file
 *     /root/reference/print_file.out, is turned into a fixture by tests/golden/make_snopt_pin.py (inputs reconstructed from
 *     the logged start point and gradients): on it the QP stage equals SNOPT's logged start point (= iSWIFT's optimum) to
 *     5e-10 and the NLP stage reaches SNOPT's active set (distance rows 1 and 4) with slack 0.7549 against 0.76247 and
 *     objective 773.25 against SNOPT's unfinished 803.70 (tests/test_oracle_nlp.py, tests/test_gpu_boundary.py).  The NLP itself (rows, bounds, Jacobian, cost:
 *     include/dec_vars_constr_cost.h:245-438) is restated exactly; the solver is this repo's primal-dual
 *     filter interior-point method (DESIGN.md "IPM-F"), cross-checked against scipy SLSQP/trust-constr in
 *     tests/test_oracle_nlp.py and against the golden vectors minted by tests/golden/make_golden.py.
 */
#ifndef CBFNMPC_ORACLE_H
#define CBFNMPC_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NV_MAX 41
#define ORC_NEQ 28
#define ORC_ND_MAX 76     /* distance rows: 4 in the reference (one per stage); up to 19 per stage in the extended form */
#define ORC_MROWS_MAX 184 /* 92 linear + ORC_ND_MAX distance + 16 one-sided velocity rows at c = 4 */

typedef struct {
  /* QP stage (iSWIFT GlobalOptions.h:23-26) */
  double qp_reltol, qp_abstol;
  int qp_maxit;
  /* NLP stage (IPM-F) */
  double tol, mu_init, kappa_eps, kappa_mu, theta_mu, tau_min, push;
  double gamma_theta, gamma_phi, delta, s_theta, s_phi, eta_phi, kappa_sigma, alpha_min;
  double dw_first, dw_min, kw_minus, kw_plus, kw_plus_first, dw_max, rho_lambda, z_div;
  int max_iter, filter_cap;
  double acc_tol; /* acceptable KKT error once mu has reached tol/10 */
} orc_options;

typedef struct {
  double Qw, Rw, Pw, slack_w, mu_fric, h_com, g_acc, Ts, eps_dist, v_max, box;
} orc_gains;

typedef struct {
  double x[ORC_NV_MAX];    /* final decision vector, reference layout [X16 | U8 | Lambda 4c | delta]          */
  double x_qp[ORC_NV_MAX]; /* QP-stage solution (what MPC_dist logs as qp_solution_eventbased_)            */
  double y[ORC_NEQ];
  double z[ORC_MROWS_MAX];
  double s[ORC_MROWS_MAX];
  double alpha[20];        /* 4x5 Bezier coefficients, row-major (state row, control point)                */
  double alpha_buf_next[4];
  double obs[2];
  double kkt;              /* final unscaled KKT error (max-norm of stationarity, equality residual, s*z)  */
  double mu;
  int nv, c, m;
  int status;              /* 0 converged, 1 max-iter, 2 line-search failure (infeasible), 3 inertia, 4 bad input */
  int qp_flag, qp_iters, nlp_iters, n_fact, n_ls;
} orc_result;

void orc_default_options(orc_options *o);
void orc_default_gains(orc_gains *g);

/* src/MPC_dist.cpp:99-127.  Ad row-major 4x4, Bd row-major 4x2 */
void orc_lip_discrete(const orc_gains *g, double *Ad, double *Bd);

/* src/MPC_dist.cpp:1230-1252: F (2 x c, row-major) from toe xy (2x4 row-major) and contact flags; returns c or -1 */
int orc_stance_matrix(const double *feet_xy, const int *contact, double *F);

/* src/MPC_dist.cpp:132-321.  All matrices COLUMN-major (Eigen default).  nv = 25+4c, p = 28, mi = 60+8c */
void orc_pack_dense(const orc_gains *g, const double *x0, const double *xdes, const double *F, int c,
                    double *Q, double *f, double *Aeq, double *beq, double *G, double *h);

/* src/MPC_dist.cpp:371-396.  obs_xy is 2 x K row-major */
void orc_generate_reference(const double *Pstart, const double *Pobs, int n_obs, double goal_x, double goal_y, int loop,
                            double *Pr, double *Prd);
void orc_cop_planner(const double *Pr, const double *Prd, int ref_cols, int agent_id, int gait_domain, double *xdes16);
void orc_footholds(const double *toe_pos12, int gait_domain, const double *agent_initial, double *feet_xy8);
int orc_closest_obstacle(const double *pos, const double *obs_xy, int K, double *out);

/* src/MPC_dist.cpp:784-855 (effective semantics) */
void orc_bezier_fit(const double *alpha_buf, const double *X16, double *alpha20, double *alpha_buf_next);

/* iSWIFT restated, dense, column-major inputs.  Returns flag (0 optimal, 2 max-iter, 1 KKT failure). */
typedef int (*orc_qp_backend)(int n, int m, int p, const double *P, const double *c, const double *A, const double *b, const double *G,
                              const double *h, double *x, double *y, double *z, double *s, int *iters, double *t_setup, double *t_solve);
void orc_set_qp_backend(orc_qp_backend f); /* NULL = orc_qp_mehrotra */
int orc_qp_mehrotra(int n, int m, int p, const double *P, const double *c, const double *A, const double *b,
                    const double *G, const double *h, double reltol, double abstol, int maxit, double sigma_d,
                    double *x, double *y, double *z, double *s, int *iters);

/* NLP stage on dense nlpTuple data (include/pack_nlp_data.h:11-29), column-major. with_nl = 0 drops the
 * 4 distance + 8 velocity rows (then it is the QP solved by IPM-F; used in tests only). */
int orc_nlp_ipm(int nv, int mi, const double *Q, const double *f, const double *Aeq, const double *beq,
                const double *G, const double *h, const double *obs, const double *x_start, int with_nl,
                const orc_gains *g, const orc_options *o, orc_result *res);

/* The same NLP with a general set of distance rows ("extended mode": K obstacles and neighbour rows per
 * stage; no reference implementation exists, this is the checker the extended kernel will be held against):
 *   row r:  |p_{dstage[r]+1} - dctr[r]|^2 + delta >= deps[r],   r = 0 .. nd-1,  dstage in 0..3, dctr nd x 2 row-major.
 * Row order [G rows | nd distance rows | 8 + 8 velocity rows].  orc_nlp_ipm is the case nd = 4, dstage = 0..3, one centre, eps_dist. */
int orc_nlp_ipm_rows(int nv, int mi, const double *Q, const double *f, const double *Aeq, const double *beq,
                     const double *G, const double *h, int nd, const int *dstage, const double *dctr, const double *deps,
                     const double *x_start, int with_vel, const orc_gains *g, const orc_options *o, orc_result *res);

/* One full run_NMPC: pack -> QP -> (mode 1) closest obstacle -> NLP -> Bezier fit. */
int orc_solve_instance(const orc_gains *g, const orc_options *o, const double *x0, const double *xdes,
                       const double *feet_xy, const int *contact, const double *obs_xy, int K,
                       const double *alpha_buf, int mode, orc_result *res);

/* Batch over B instances, AoS inputs (instance-major), nthreads host threads (one solve per thread at a time). */
int orc_solve_batch(const orc_gains *g, const orc_options *o, int B, const double *x0, const double *xdes,
                    const double *feet_xy, const int *contact, const double *obs_xy, int K,
                    const double *alpha_buf, int mode, int nthreads, orc_result *res);

#ifdef __cplusplus
}
#endif
#endif
