// tests/emu/emu_lane.cpp -- DEV/TEST TOOL ONLY.  Compiles the *product kernel source*
// (srb-cbf-nmpc_b200/csrc/cbfnmpc_lane.h) for the host and runs one warp (8 four-lane groups = 8 instances of one stance class)
// on 32 cooperative fibers with mask-based rendezvous -- what __shfl_sync / __ballot_sync with the warp-uniform member mask of
// lane_prims.h do on the device -- so the kernel logic AND its mask discipline can be checked against the oracle on a machine
// without a GPU (pytest -m "not gpu").  A rendezvous is released only when every lane named in its mask has arrived with the
// same mask (sitecheck build: and from the same source-level site); lanes that wait with incompatible masks are a deadlock on
// the device and an error here.  It is not linked into libcbfnmpc.so, not reachable from the C ABI, and never used as a fallback.
#define CBF_HOST_EMU
#define LANE_NT 32
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
WarpCtx g_warp;
static ucontext_t sched_ctx, lane_ctx[32];
static char* stacks[32];
static int state[32];       // 0 runnable, 1 waiting at a rendezvous, 2 finished
static unsigned wmask[32];  // mask each lane waits with
static void* site[32];      // call address of the rendezvous each lane is waiting at
__attribute__((noinline)) void rendezvous(unsigned mask) {
  const int me = g_warp.cur;
  state[me] = 1; wmask[me] = mask;
  site[me] = __builtin_return_address(0);
  swapcontext(&lane_ctx[me], &sched_ctx);
}
}  // namespace lemu

static const KParams* g_P; static const BatchIO* g_io; static int g_C;
static int g_b[32]; static bool g_valid[32];
static double g_priv[LANE_NPRIV * 32];
static bool g_late = false;
extern "C" void emu_set_late_obstacles(int on) { g_late = on != 0; }
static void lane_main() {
  const int l = lemu::g_warp.cur;
  Priv pr; pr.base = g_priv + l;
  if (g_late) {   // the instantiation of cbfnmpc_lane_kernel_late: closest-obstacle step after the QP stage
    if (g_C == 4) lane_run_group<4, true>(*g_P, *g_io, g_b[l], g_valid[l], pr);
    else lane_run_group<2, true>(*g_P, *g_io, g_b[l], g_valid[l], pr);
  } else if (g_C == 4) lane_run_group<4>(*g_P, *g_io, g_b[l], g_valid[l], pr);
  else lane_run_group<2>(*g_P, *g_io, g_b[l], g_valid[l], pr);
  lemu::state[l] = 2;
  swapcontext(&lemu::lane_ctx[l], &lemu::sched_ctx);
}

// one warp: group k (lanes 4k .. 4k+3) runs instance list[k] (k < n), the other groups are idle (valid = false)
static int run_warp(const KParams* P, const BatchIO* io, const int* list, int n, int C) {
  using namespace lemu;
  g_P = P; g_io = io; g_C = C;
  // shared memory is not initialised on the device: poison the thread-private columns so that a read-before-write shows up
  for (int t = 0; t < LANE_NPRIV * 32; t++) g_priv[t] = __builtin_nan("");
  for (int l = 0; l < 32; l++) {
    g_valid[l] = (l >> 2) < n; g_b[l] = g_valid[l] ? list[l >> 2] : 0;
    if (!stacks[l]) stacks[l] = (char*)malloc(1 << 20);
    getcontext(&lane_ctx[l]);
    lane_ctx[l].uc_stack.ss_sp = stacks[l];
    lane_ctx[l].uc_stack.ss_size = 1 << 20;
    lane_ctx[l].uc_link = &sched_ctx;
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
    // release every rendezvous whose lanes have all arrived with the same mask (sitecheck build: at the same site)
    int released = 0;
    for (int l = 0; l < 32; l++) {
      if (state[l] != 1) continue;
      const unsigned m = wmask[l];
      if (!((m >> l) & 1u)) { fprintf(stderr, "emu_lane: lane %d waits with a mask that does not name it (%08x)\n", l, m); return -1; }
      bool all = true;
      for (int k = 0; k < 32; k++) if (((m >> k) & 1u) && !(state[k] == 1 && wmask[k] == m)) { all = false; break; }
      if (!all) continue;
#ifdef CBF_EMU_SITECHECK  // -O0: one call address per source-level shuffle / vote site
      for (int k = 0; k < 32; k++) if (((m >> k) & 1u) && site[k] != site[l]) {
        Dl_info di; dladdr((void*)&lemu::rendezvous, &di);
        fprintf(stderr, "emu_lane: lanes %d and %d (instances %d, %d) meet with mask %08x at different sites: %p %p (library base %p)\n",
                l, k, g_b[l], g_b[k], m, site[l], site[k], di.dli_fbase);
        return -1;
      }
#endif
      for (int k = 0; k < 32; k++) if ((m >> k) & 1u) state[k] = 0;
      released++;
    }
    if (nrun == 0 && released == 0) {
      fprintf(stderr, "emu_lane: deadlock (lanes wait with incompatible masks, or for lanes that have finished):");
      for (int l = 0; l < 32; l++) fprintf(stderr, " %d:%d:%08x", l, state[l], wmask[l]);
      fprintf(stderr, "\n");
      return -1;
    }
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
  io.B = B; io.ld = B; io.mode = mode; io.obs_flag = nullptr; io.obs_seq = 0;
  // stance classes as the device sorts them (cbfnmpc_classify_kernel): all 4-foot instances, then the diagonal pairs, 8 per warp
  for (int cls = 4; cls >= 2; cls -= 2) {
    int list[8], n = 0;
    for (int b = 0; b <= B; b++) {
      if (b < B) {
        lemu::g_warp.cur = 0;
        const int C = lane_classify(io, b);
        if (C < 0) { if (cls == 4) lane_store_bad(io, b); continue; }
        if (C != cls) continue;
        list[n++] = b;
      }
      if (n == 8 || (b == B && n > 0)) { if (run_warp(&P, &io, list, n, cls)) return -2; n = 0; }
    }
  }
  return 0;
}
}
