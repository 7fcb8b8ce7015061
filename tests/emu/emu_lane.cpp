// tests/emu/emu_lane.cpp -- DEV/TEST TOOL ONLY.  Compiles the *product kernel source*
// (srb-cbf-nmpc_b200/csrc/cbfnmpc_lane.h) for the host and runs each 4-lane group on 4 cooperative fibers, so the
// kernel logic can be checked against the oracle on a machine without a GPU (pytest -m "not gpu").
// It is not linked into libcbfnmpc.so, not reachable from the C ABI, and never used as a fallback.
#define CBF_HOST_EMU
#define LANE_NT 4
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>
#include <dlfcn.h>
#include "../../srb-cbf-nmpc_b200/csrc/cbfnmpc_lane.h"
#include "../../srb-cbf-nmpc_b200/csrc/cbfnmpc_host.h"

#ifdef CBF_DBGBUF
double g_dbg[8192]; int g_dbg_n;
extern "C" int emu_debug_read(double* out, int reset) { int n = g_dbg_n; memcpy(out, g_dbg, sizeof(g_dbg)); if (reset) g_dbg_n = 0; return n; }
#endif
namespace lemu {
GroupCtx g_grp;
static ucontext_t sched_ctx, lane_ctx[4];
static char* stacks[4];
static int state[4];  // 0 running, 1 at barrier, 2 finished
static void* site[4];  // call address of the barrier each lane is waiting at
__attribute__((noinline)) void barrier() {
  int me = g_grp.cur;
  state[me] = 1;
  site[me] = __builtin_return_address(0);
  swapcontext(&lane_ctx[me], &sched_ctx);
}
}  // namespace lemu

static const KParams* g_P; static const BatchIO* g_io; static int g_b, g_C;
static double g_priv[LANE_NPRIV * 4];
static void lane_main() {
  Priv pr; pr.base = g_priv + lemu::g_grp.cur;
  if (g_C == 4) lane_run_group<4>(*g_P, *g_io, g_b, true, pr);
  else lane_run_group<2>(*g_P, *g_io, g_b, true, pr);
  lemu::state[lemu::g_grp.cur] = 2;
  swapcontext(&lemu::lane_ctx[lemu::g_grp.cur], &lemu::sched_ctx);
}

static int run_group(const KParams* P, const BatchIO* io, int b, int C) {
  g_P = P; g_io = io; g_b = b; g_C = C;
  // shared memory is not initialised on the device: poison the thread-private columns so that a read-before-write shows up
  for (int t = 0; t < LANE_NPRIV * 4; t++) g_priv[t] = __builtin_nan("");
  for (int l = 0; l < 4; l++) {
    if (!lemu::stacks[l]) lemu::stacks[l] = (char*)malloc(1 << 20);
    getcontext(&lemu::lane_ctx[l]);
    lemu::lane_ctx[l].uc_stack.ss_sp = lemu::stacks[l];
    lemu::lane_ctx[l].uc_stack.ss_size = 1 << 20;
    lemu::lane_ctx[l].uc_link = &lemu::sched_ctx;
    makecontext(&lemu::lane_ctx[l], lane_main, 0);
    lemu::state[l] = 0;
  }
  for (;;) {
    int nb = 0, nf = 0;
    for (int l = 0; l < 4; l++) {
      if (lemu::state[l] == 2) { nf++; continue; }
      lemu::g_grp.cur = l;
      lemu::state[l] = 0;
      swapcontext(&lemu::sched_ctx, &lemu::lane_ctx[l]);
      if (lemu::state[l] == 2) nf++; else nb++;
    }
    if (nf == 4) return 0;
    // CBF_EMU_SITECHECK builds (-O0: one call address per source-level shuffle site) verify that the 4 lanes always meet at the same site
#ifdef CBF_EMU_SITECHECK
    if (nb == 4 && !(lemu::site[0] == lemu::site[1] && lemu::site[1] == lemu::site[2] && lemu::site[2] == lemu::site[3])) {
      Dl_info di; dladdr((void*)&lemu::barrier, &di);
      fprintf(stderr, "emu_lane: lanes of instance %d meet at different shuffle sites: %p %p %p %p (library base %p)\n", b, lemu::site[0],
              lemu::site[1], lemu::site[2], lemu::site[3], di.dli_fbase);
      return -1;
    }
#endif
    if (nf != 0 && nb != 0) { fprintf(stderr, "emu_lane: divergent barrier (instance %d: %d lanes finished, %d waiting)\n", b, nf, nb); return -1; }
  }
}

extern "C" {
void cbfnmpc_default_gains(cbfnmpc_gains* g) { host_default_gains(g); }
void cbfnmpc_default_options(cbfnmpc_options* o) { host_default_options(o); }
// same argument meaning as cbfnmpc_solve_batch_host, plus explicit gains/options; host pointers
int emu_solve_batch_host(const cbfnmpc_gains* g, const cbfnmpc_options* o, int B, int mode, const cbfnmpc_in* in,
                         const cbfnmpc_out* out, const double* xstart) {
  KParams P;
  host_fill_kparams(g, o, &P);
  BatchIO io;
  io.x0 = in->x0; io.xdes = in->xdes; io.feet = in->feet_xy; io.contact = in->contact; io.obs = in->obs_xy; io.n_obs = in->n_obs;
  io.abuf = in->alpha_buf; io.xstart = xstart;
  io.X = out->X; io.U = out->U; io.lam = out->lam; io.slack = out->slack; io.xqp = out->x_qp; io.alpha = out->alpha;
  io.abuf_next = out->alpha_buf_next; io.kkt = out->kkt; io.status = out->status; io.qp_iters = out->qp_iters;
  io.nlp_iters = out->nlp_iters; io.n_fact = out->n_fact; io.active = (unsigned long long*)out->active;
  io.B = B; io.ld = B; io.mode = mode;
  for (int b = 0; b < B; b++) {
    lemu::g_grp.cur = 0;
    const int C = lane_classify(io, b);
    if (C < 0) { lane_store_bad(io, b); continue; }
    if (run_group(&P, &io, b, C)) return -2;
  }
  return 0;
}
}
