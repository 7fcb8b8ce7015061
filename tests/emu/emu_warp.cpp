// tests/emu/emu_warp.cpp -- DEV/TEST TOOL ONLY.  Compiles the *product kernel source*
// (srb-cbf-nmpc_b200/csrc/cbfnmpc_instance.h) for the host and runs each 32-lane warp on 32 cooperative fibers,
// so the kernel logic can be checked against the oracle on a machine without a GPU (pytest -m "not gpu").
// It is not linked into libcbfnmpc.so, not reachable from the C ABI, and never used as a fallback.
#define CBF_HOST_EMU
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>
#include "../../srb-cbf-nmpc_b200/csrc/cbfnmpc_host.h"
#include "../../srb-cbf-nmpc_b200/csrc/cbfnmpc_instance.h"

namespace emu {
WarpCtx g_warp;
static ucontext_t sched_ctx, lane_ctx[32];
static char* stacks[32];
static int state[32];  // 0 running, 1 at barrier, 2 finished
void barrier() {
  int me = g_warp.cur;
  state[me] = 1;
  swapcontext(&lane_ctx[me], &sched_ctx);
}
}  // namespace emu

static WSB* g_ws; static const KParams* g_P; static const BatchIO* g_io; static int g_b;
static void lane_main() {
  solve_warp(*g_ws, *g_P, *g_io, g_b);
  emu::state[emu::g_warp.cur] = 2;
  swapcontext(&emu::lane_ctx[emu::g_warp.cur], &emu::sched_ctx);
}

static int run_warp(const KParams* P, const BatchIO* io, int b) {
  static WSB ws;
  g_ws = &ws; g_P = P; g_io = io; g_b = b;
  for (int l = 0; l < 32; l++) {
    if (!emu::stacks[l]) emu::stacks[l] = (char*)malloc(1 << 18);
    getcontext(&emu::lane_ctx[l]);
    emu::lane_ctx[l].uc_stack.ss_sp = emu::stacks[l];
    emu::lane_ctx[l].uc_stack.ss_size = 1 << 18;
    emu::lane_ctx[l].uc_link = &emu::sched_ctx;
    makecontext(&emu::lane_ctx[l], lane_main, 0);
    emu::state[l] = 0;
  }
  for (;;) {
    int nb = 0, nf = 0;
    for (int l = 0; l < 32; l++) {
      if (emu::state[l] == 2) { nf++; continue; }
      emu::g_warp.cur = l;
      emu::state[l] = 0;
      swapcontext(&emu::sched_ctx, &emu::lane_ctx[l]);
      if (emu::state[l] == 2) nf++; else nb++;
    }
    if (nf == 32) return 0;
    if (nf != 0 && nb != 0) { fprintf(stderr, "emu: divergent barrier (instance %d: %d lanes finished, %d waiting)\n", b, nf, nb); return -1; }
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
  io.B = B; io.mode = mode;
  for (int b = 0; b < B; b++)
    if (run_warp(&P, &io, b)) return -2;
  return 0;
}
}
