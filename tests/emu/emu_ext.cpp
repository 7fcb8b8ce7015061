// tests/emu/emu_ext.cpp -- DEV/TEST TOOL ONLY.  Compiles the *product kernel source* of the extended mode
// (srb-cbf-nmpc_b200/csrc/cbfnmpc_ext.h) for the host and runs one warp on 32 cooperative fibers with mask-based rendezvous
// (what __shfl_sync / __ballot_sync with a group or team mask do on the device), so that the solver logic -- including the
// team exchange -- can be checked against the oracle on a machine without a GPU.  Not linked into libcbfnmpc.so, not reachable
// from the C ABI, never a fallback.
#define CBF_HOST_EMU
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>
#include "../../srb-cbf-nmpc_b200/csrc/cbfnmpc_ext.h"
#include "../../srb-cbf-nmpc_b200/csrc/cbfnmpc_host.h"

namespace xemu {
WarpCtx g_warp;
static ucontext_t sched_ctx, lane_ctx[32];
static char* stacks[32];
static int state[32];          // 0 runnable, 1 waiting at a rendezvous, 2 finished
static unsigned wmask[32];
void rendezvous(unsigned mask) {
  const int me = g_warp.cur;
  state[me] = 1; wmask[me] = mask;
  swapcontext(&lane_ctx[me], &sched_ctx);
}
void yield() {   // stays runnable; the scheduler comes back to it after the other lanes have had their turn
  const int me = g_warp.cur;
  swapcontext(&lane_ctx[me], &sched_ctx);
}
}  // namespace xemu

static const XParams* g_XP; static const ExtIO* g_io; static int g_C, g_W, g_nlanes; static bool g_fused;
static int g_b[32]; static bool g_valid[32];
static double* g_priv;
static void lane_main() {
  const int l = xemu::g_warp.cur;
  XPriv pr; pr.base = g_priv + l;
  if (g_fused) {   // the fused exchange (xmode 4): mailbox and sequence flags in host memory, the polling lanes yield
    if (g_W == 4) { if (g_C == 4) ext_run_group<4, 4, true>(*g_XP, *g_io, g_b[l], g_valid[l], pr); else ext_run_group<2, 4, true>(*g_XP, *g_io, g_b[l], g_valid[l], pr); }
    else { if (g_C == 4) ext_run_group<4, 8, true>(*g_XP, *g_io, g_b[l], g_valid[l], pr); else ext_run_group<2, 8, true>(*g_XP, *g_io, g_b[l], g_valid[l], pr); }
  } else
  if (g_W == 4) { if (g_C == 4) ext_run_group<4, 4>(*g_XP, *g_io, g_b[l], g_valid[l], pr); else ext_run_group<2, 4>(*g_XP, *g_io, g_b[l], g_valid[l], pr); }
  else { if (g_C == 4) ext_run_group<4, 8>(*g_XP, *g_io, g_b[l], g_valid[l], pr); else ext_run_group<2, 8>(*g_XP, *g_io, g_b[l], g_valid[l], pr); }
  xemu::state[l] = 2;
  swapcontext(&xemu::lane_ctx[l], &xemu::sched_ctx);
}

// one warp: lanes [0, 32), lane l runs instance g_b[l]
static int run_warp() {
  using namespace xemu;
  for (int l = 0; l < 32; l++) {
    if (!stacks[l]) stacks[l] = (char*)malloc(1 << 19);
    getcontext(&lane_ctx[l]);
    lane_ctx[l].uc_stack.ss_sp = stacks[l]; lane_ctx[l].uc_stack.ss_size = 1 << 19; lane_ctx[l].uc_link = &sched_ctx;
    makecontext(&lane_ctx[l], lane_main, 0);
    state[l] = 0;
  }
  for (;;) {
    int nrun = 0, nfin = 0;
    for (int l = 0; l < 32; l++) {
      if (state[l] == 2) { nfin++; continue; }
      if (state[l] == 1) continue;
      g_warp.cur = l;
      swapcontext(&sched_ctx, &lane_ctx[l]);
      nrun++;
    }
    if (nfin == 32) return 0;
    // release every rendezvous whose lanes have all arrived with the same mask
    int released = 0;
    for (int l = 0; l < 32; l++) {
      if (state[l] != 1) continue;
      const unsigned m = wmask[l];
      bool all = true;
      for (int k = 0; k < 32; k++) if (((m >> k) & 1u) && !(state[k] == 1 && wmask[k] == m)) { all = false; break; }
      if (all) { for (int k = 0; k < 32; k++) if ((m >> k) & 1u) state[k] = 0; released++; }
    }
    if (nrun == 0 && released == 0) {
      fprintf(stderr, "emu_ext: deadlock (lanes wait with incompatible masks):");
      for (int l = 0; l < 32; l++) fprintf(stderr, " %d:%d:%08x", l, state[l], wmask[l]);
      fprintf(stderr, "\n");
      return -1;
    }
  }
}

struct emu_ext_args {
  int B, N, T, K_obs, n_nbr, xmode;
  double eps_obs, eps_nbr, relax, pub_tol;
  int max_team_iters, mu_stall;
  const double *x0, *xdes, *feet, *obs, *abuf, *nbr;   // nbr: [n_nbr][2N + 2][B] (xmode 2)
  const int* contact;
  double *X, *U, *lam, *slack, *xqp, *alpha, *abuf_next, *kkt, *pub;
  int *status, *qp_iters, *nlp_iters, *team_iters;
};

extern "C" int emu_ext_solve(const cbfnmpc_gains* g, const cbfnmpc_options* o, const emu_ext_args* a) {
  XParams XP;
  host_fill_kparams(g, o, &XP.k);
  XP.eps_obs = a->eps_obs; XP.eps_nbr = a->eps_nbr; XP.relax = a->relax; XP.pub_tol = a->pub_tol; XP.max_team_it = a->max_team_iters; XP.mu_stall = a->mu_stall;
  ExtIO io;
  memset(&io, 0, sizeof(io));
  io.x0 = a->x0; io.xdes = a->xdes; io.feet = a->feet; io.contact = a->contact; io.obs = a->obs; io.abuf = a->abuf;
  for (int i = 0; i < a->n_nbr && i < 3; i++) io.nbr[i] = a->nbr ? a->nbr + (size_t)i * (2 * a->N + 2) * a->B : nullptr;
  io.pub = a->pub;
  io.X = a->X; io.U = a->U; io.lam = a->lam; io.slack = a->slack; io.xqp = a->xqp; io.alpha = a->alpha; io.abuf_next = a->abuf_next; io.kkt = a->kkt;
  io.status = a->status; io.qp_iters = a->qp_iters; io.nlp_iters = a->nlp_iters; io.team_iters = a->team_iters;
  io.B = a->B; io.ld = a->B; io.N = a->N; io.K_obs = a->K_obs; io.n_nbr = a->n_nbr; io.T = a->T; io.xmode = a->xmode; io.phase = 0; io.n_run = 0;
  // xmode 4 (the fused exchange of cbfnmpc_ext_solve_fused on ONE emulated warp): member-major batch of ONE group of ranks, a->T == 1,
  // a->n_nbr + 1 members per team, B / (n_nbr + 1) teams; every instance of the batch must sit in the warp that runs (its team mates
  // poll for it), i.e. B <= 32 / N instances of one stance class
  double* mbox = nullptr; int fctl[4] = {0, 0, 0, 0};
  g_fused = a->xmode == 4;
  if (g_fused) {
    const int t_all = a->n_nbr + 1, n_teams = a->B / t_all;
    const size_t blk = (size_t)(2 * a->N + 2) * n_teams, need = 2 * (size_t)t_all * blk + (size_t)t_all * n_teams + 2;
    mbox = (double*)calloc(need, sizeof(double));
    io.mbox[0] = mbox; io.n_grp = 1; io.n_teams = n_teams; io.member0 = 0; io.fused_err = fctl; io.fused_its = fctl + 1;
  }
  const int W = a->N, K = a->K_obs + a->n_nbr;
  const int npriv = ext_npriv(K);
  g_priv = (double*)malloc(sizeof(double) * (size_t)npriv * 32);
  g_XP = &XP; g_io = &io; g_W = W;
  const int lanes_per_team = a->T * W, teams_per_warp = 32 / lanes_per_team, n_teams = a->B / a->T;
  int rc = 0;
  for (int cls = 4; cls >= 2 && rc == 0; cls -= 2) {
    // teams of this stance class, in batch order
    int* list = (int*)malloc(sizeof(int) * (n_teams + 1)); int n = 0;
    for (int t = 0; t < n_teams; t++) {
      const int c = ext_classify_team(io, t);
      if (c == cls) list[n++] = t;
      else if (c < 0 && cls == 4) {
        for (int m = 0; m < a->T; m++) ext_store_bad<4, 4>(io, t * a->T + m, 0);
        if (io.xmode == 4) ext_fused_publish_bad(io, t);
      }
    }
    g_C = cls;
    for (int w0 = 0; w0 < n && rc == 0; w0 += teams_per_warp) {
      for (int t = 0; t < npriv * 32; t++) g_priv[t] = __builtin_nan("");   // shared memory is not initialised on the device
      for (int l = 0; l < 32; l++) {
        const int ts = w0 + l / lanes_per_team;
        g_valid[l] = ts < n;
        g_b[l] = g_valid[l] ? list[ts] * a->T + (l % lanes_per_team) / W : 0;
      }
      rc = run_warp();
    }
    free(list);
  }
  free(g_priv);
  if (mbox) free(mbox);
  if (g_fused && fctl[0] != 0) rc = -3;   // a member waited in vain
  g_fused = false;
  return rc;
}
