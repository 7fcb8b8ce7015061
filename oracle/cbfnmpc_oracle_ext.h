/* TEST INFRASTRUCTURE -- CPU oracle of the EXTENDED mode (BASELINE configs 4 and 5).
 *
 * NO REFERENCE IMPLEMENTATION EXISTS for this mode: the reference's NLP has N = 4 grid points, one obstacle and no inter-agent
 * rows (src/MPC_dist.cpp:81-454; include/dec_vars_constr_cost.h).  What is generalised, following the reference's own hints:
 *   - horizon N = 4 ND grid points (ND gait domains; the stance matrix may differ per domain),
 *   - K_obs obstacle rows per grid point   |p_k - o_i|^2   + delta >= eps_obs   (the reference's row, DV:262-265, per obstacle),
 *   - n_nbr neighbour rows per grid point  |p_k - p^j_k|^2 + delta >= eps_nbr   (DV:402 "epsilon = 2.2 // robot to robot";
 *     p^j_k = neighbour j's predicted position at grid point k; the reference's disabled rule extrapolates it, MPC_dist.cpp:391),
 *   - team mode ("IPM-F-J"): the members of a team iterate together and read each other's CURRENT iterate at the top of every
 *     interior-point iteration (Jacobi) -- the per-iteration exchange of BASELINE config 4.
 * With N = 4, K_obs = 1, n_nbr = 0 this is orc_nlp_ipm of cbfnmpc_oracle.c, iteration by iteration (tests/test_oracle_ext.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file; "parity unpinned" by construction. */
#ifndef CBFNMPC_ORACLE_EXT_H
#define CBFNMPC_ORACLE_EXT_H
#include "cbfnmpc_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EXT_NMAX 8
#define ORC_EXT_KMAX 19 /* distance rows per grid point (obstacles + neighbours) */
#define ORC_EXT_NV_MAX (ORC_EXT_NMAX * 10 + 1)

typedef struct {
  double x[ORC_EXT_NV_MAX];    /* [X 4N | U 2N | Lambda cN | delta] */
  double x_qp[ORC_EXT_NV_MAX];
  double kkt, mu;
  int N, c, nv, n_rows;        /* n_rows = distance rows per grid point */
  int status;                  /* as orc_result */
  int qp_flag, qp_iters, nlp_iters, n_fact, n_ls, n_bump;
  double zd[ORC_EXT_NMAX * ORC_EXT_KMAX], sd[ORC_EXT_NMAX * ORC_EXT_KMAX]; /* multipliers / slacks of the distance rows, [row i][grid point k] */
} orc_ext_result;

typedef struct orc_ext orc_ext;

/* src/MPC_dist.cpp:132-321 for N grid points: F = ND stance matrices (2 x c each, row-major, one after the other).
 * nv = N (6 + c) + 1, p = 7 N, mi = 16 N - 4 + 2 c N; column-major outputs. */
void orc_ext_pack_dense(const orc_gains *g, int N, const double *x0, const double *xdes, const double *F, int c, double *Q, double *f,
                        double *Aeq, double *beq, double *G, double *h);

/* One agent.  feet_xy: ND blocks of 8 (2 x 4 row-major); obs_xy: 2 x K_obs row-major.  Returns NULL on bad input. */
orc_ext *orc_ext_create(const orc_gains *g, const orc_options *o, int N, const double *x0, const double *xdes, const double *feet_xy,
                        const int *contact, int K_obs, const double *obs_xy, double eps_obs, int n_nbr, double eps_nbr);
void orc_ext_free(orc_ext *e);
int orc_ext_qp(orc_ext *e);                        /* QP stage (iSWIFT restated) -> start point; returns the flag */
void orc_ext_set_start(orc_ext *e, const double *x); /* instead of orc_ext_qp: start the NLP here */
/* nbr: n_nbr x N x 2 (neighbour j, grid point k, xy); NULL when n_nbr = 0 */
void orc_ext_begin(orc_ext *e, const double *nbr);
/* one interior-point iteration with the given neighbour positions.  Returns 0: stepped, 1: converged for these neighbour positions
 * (no step taken; a later call with other positions may step again), 2: finished for good (status set: failure or iteration limit) */
int orc_ext_iter(orc_ext *e, const double *nbr);
void orc_ext_positions(const orc_ext *e, double *pos /* N x 2 */);
void orc_ext_result_get(const orc_ext *e, orc_ext_result *r);

/* drivers: fixed neighbour trajectories (any n_nbr) ... */
int orc_ext_solve(const orc_gains *g, const orc_options *o, int N, const double *x0, const double *xdes, const double *feet_xy,
                  const int *contact, int K_obs, const double *obs_xy, double eps_obs, int n_nbr, double eps_nbr, const double *nbr,
                  orc_ext_result *res);
/* ... and a team of T agents (n_nbr = T - 1; neighbour j of agent a = the j-th OTHER member, in team order), under-relaxed Jacobi:
 * every member publishes pos <- pos + relax (iterate - pos) at the top of an iteration (relax = 0.7: plain Jacobi, relax = 1,
 * overshoots when two members push each other away and oscillates on ~5 % of the config-4 teams); the team has finished when
 * nobody stepped and every published trajectory is within pub_tol of its iterate, or after max_team_iters iterations.  In team
 * mode a member whose barrier parameter has not decreased for mu_stall iterations lowers it anyway (see orc_ext_iter).  Arrays are agent-major: x0 T x 4, xdes T x 4N, feet_xy T x 8 ND, contact T x 4, obs_xy T x 2 x K_obs.  *team_iters =
 * iterations of the team loop.  Returns 0 when every member converged. */
int orc_ext_solve_team(const orc_gains *g, const orc_options *o, int T, int N, const double *x0, const double *xdes, const double *feet_xy,
                       const int *contact, int K_obs, const double *obs_xy, double eps_obs, double eps_nbr, double relax, double pub_tol,
                       int max_team_iters, int mu_stall, orc_ext_result *res, int *team_iters);
/* nteams teams (or, with T = 1, independent agents) on nthreads host threads; arrays as above with a leading team index */
int orc_ext_solve_teams(const orc_gains *g, const orc_options *o, int nteams, int T, int N, const double *x0, const double *xdes,
                        const double *feet_xy, const int *contact, int K_obs, const double *obs_xy, double eps_obs, double eps_nbr,
                        double relax, double pub_tol, int max_team_iters, int mu_stall, int nthreads, orc_ext_result *res, int *team_iters);

int orc_bk_factor(int n, double *A, int lda, int *ipiv);
void orc_bk_inertia(int n, const double *A, int lda, const int *ipiv, int *npos, int *nneg, int *nzero);
void orc_bk_solve(int n, const double *A, int lda, const int *ipiv, double *b);
#ifdef __cplusplus
}
#endif
#endif
