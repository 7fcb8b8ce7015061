// cbfnmpc_ext_api.cuh -- C ABI of the extended mode (include/cbfnmpc.h, "extended mode").  Included by cbfnmpc_api.cu inside its
// extern "C" block.  NCCL is reached through dlopen (libnccl.so.2): the library loads, and every other entry point works, on
// machines without NCCL; only the agent-sharded call with more than one rank needs it.

// ---- NCCL by name (the five entry points used; types as in nccl.h)
typedef struct { char internal[128]; } xnccl_uid;
typedef void* xnccl_comm;
struct XNccl {
  void* h;
  int (*GetUniqueId)(xnccl_uid*);
  int (*CommInitRank)(xnccl_comm*, int, xnccl_uid, int);
  int (*AllGather)(const void*, void*, size_t, int, xnccl_comm, cudaStream_t);
  int (*CommDestroy)(xnccl_comm);
  const char* (*GetErrorString)(int);
};
static XNccl g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
static int xnccl_load() {
  if (g_nccl.h) return CBFNMPC_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(CBFNMPC_ERR_ARG, std::string("NCCL is not available: ") + dlerror());
  g_nccl.GetUniqueId = (int (*)(xnccl_uid*))dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (int (*)(xnccl_comm*, int, xnccl_uid, int))dlsym(h, "ncclCommInitRank");
  g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, xnccl_comm, cudaStream_t))dlsym(h, "ncclAllGather");
  g_nccl.CommDestroy = (int (*)(xnccl_comm))dlsym(h, "ncclCommDestroy");
  g_nccl.GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
    return fail(CBFNMPC_ERR_ARG, "NCCL library lacks an entry point");
  g_nccl.h = h;
  return CBFNMPC_OK;
}
#define XNCCL_DOUBLE 8 /* ncclFloat64 */
#define XNCCL_INT32 2  /* ncclInt32 */
#define NK(call)                                                                                                        \
  do {                                                                                                                  \
    int r_ = (call);                                                                                                    \
    if (r_ != 0) return fail(CBFNMPC_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(r_));                  \
  } while (0)

void cbfnmpc_ext_default_cfg(cbfnmpc_ext_cfg* c) {
  if (!c) return;
  c->N = 4; c->team_size = 1;
  c->eps_obs = (double)1.9f;    // dec_vars_constr_cost.h:401
  c->eps_nbr = (double)2.2f;    // :402 "for robot to robot"
  c->relax = 0.7; c->pub_tol = 1e-10;
  c->mu_stall_iters = 6;        // team mode: 0.2 % of the config-4 teams otherwise stall with the barrier parameter stuck
  c->max_team_iters = 64;       // 99 % of the config-4 teams settle within 38 iterations, 99.9 % within 60; the rest crawl (a pair that
                                // pushes each other: Jacobi contraction 0.9997 per iteration) and would set the time of the whole launch
}

static int ext_check(const cbfnmpc_ctx* c, int B, const cbfnmpc_ext_cfg* cfg, const cbfnmpc_ext_in* in, const cbfnmpc_ext_out* out, int T) {
  if (!c || !cfg || !in || !out || B < 1) return fail(CBFNMPC_ERR_ARG, "extended mode: bad argument");
  if (cfg->N != 4 && cfg->N != 8) return fail(CBFNMPC_ERR_ARG, "extended mode: N must be 4 or 8");
  if (T != 1 && T != 2 && T != 4) return fail(CBFNMPC_ERR_ARG, "extended mode: team_size must be 1, 2 or 4");
  if (T * cfg->N > 32) return fail(CBFNMPC_ERR_ARG, "extended mode: a team (team_size x N lanes) must fit in one warp");
  if (B % T != 0) return fail(CBFNMPC_ERR_ARG, "extended mode: batch must be a multiple of team_size");
  if (!in->x0 || !in->xdes || !in->feet_xy || !in->contact) return fail(CBFNMPC_ERR_ARG, "extended mode: missing input array");
  if (in->n_obs < 0 || (in->n_obs > 0 && !in->obs_xy)) return fail(CBFNMPC_ERR_ARG, "extended mode: obstacle array missing");
  if (!(cfg->relax > 0.0 && cfg->relax <= 1.0) || !(cfg->pub_tol >= 0.0) || cfg->max_team_iters < 1 || cfg->mu_stall_iters < 1) return fail(CBFNMPC_ERR_ARG, "extended mode: relax in (0, 1], pub_tol >= 0, max_team_iters >= 1");
  return CBFNMPC_OK;
}

static void ext_fill_io(ExtIO* io, int B, int ld, const cbfnmpc_ext_cfg* cfg, const cbfnmpc_ext_in* in, const cbfnmpc_ext_out* out) {
  memset(io, 0, sizeof(*io));
  io->x0 = in->x0; io->xdes = in->xdes; io->feet = in->feet_xy; io->contact = in->contact; io->obs = in->obs_xy; io->abuf = in->alpha_buf;
  io->X = out->X; io->U = out->U; io->lam = out->lam; io->slack = out->slack; io->xqp = out->x_qp; io->alpha = out->alpha;
  io->abuf_next = out->alpha_buf_next; io->kkt = out->kkt; io->status = out->status; io->qp_iters = out->qp_iters; io->nlp_iters = out->nlp_iters;
  io->team_iters = out->team_iters; io->pub = out->pub;
  io->B = B; io->ld = ld; io->N = cfg->N; io->K_obs = in->n_obs;
}

static int ext_launch(cbfnmpc_ctx* c, const XParams& XP, const ExtIO& io, int n_teams, int* perm, cudaStream_t st, bool classify) {
  const int K = io.K_obs + io.n_nbr;
  const size_t smem = sizeof(double) * (size_t)ext_npriv(K) * XNT;
  if (smem > 227 * 1024) return fail(CBFNMPC_ERR_ARG, "extended mode: too many rows per grid point for the shared memory of one SM");
  int* counts = perm + 2 * (size_t)n_teams;
  if (classify) {
    cbfnmpc_ext_classify_kernel<<<1, 1024, 0, st>>>(io, n_teams, perm, counts);
    CK(cudaGetLastError());
    c->launches++;
  }
  const int tpw = XNT / (io.T * io.N);
  const int grid = (n_teams + tpw - 1) / tpw + 1;
  if (io.N == 4) {
    CK(cudaFuncSetAttribute(cbfnmpc_ext_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cbfnmpc_ext_kernel<4><<<grid, XNT, smem, st>>>(XP, io, n_teams, perm, counts);
  } else {
    CK(cudaFuncSetAttribute(cbfnmpc_ext_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cbfnmpc_ext_kernel<8><<<grid, XNT, smem, st>>>(XP, io, n_teams, perm, counts);
  }
  CK(cudaGetLastError());
  c->launches++;
  return CBFNMPC_OK;
}

static int ext_perm(cbfnmpc_ctx* c, int n_teams, cudaStream_t st) {
  if (n_teams > c->x_perm_cap) {
    CK(cudaStreamSynchronize(st));
    if (c->x_perm) CK(cudaFree(c->x_perm));
    CK(cudaMalloc(&c->x_perm, sizeof(int) * (2 * (size_t)n_teams + 2)));
    c->x_perm_cap = n_teams;
  }
  return CBFNMPC_OK;
}

static void ext_xparams(const cbfnmpc_ctx* c, const cbfnmpc_ext_cfg* cfg, XParams* XP) {
  XP->k = c->kp; XP->eps_obs = cfg->eps_obs; XP->eps_nbr = cfg->eps_nbr; XP->relax = cfg->relax; XP->pub_tol = cfg->pub_tol; XP->max_team_it = cfg->max_team_iters; XP->mu_stall = cfg->mu_stall_iters;
}

int cbfnmpc_ext_solve_batch(cbfnmpc_ctx* c, int B, const cbfnmpc_ext_cfg* cfg, const cbfnmpc_ext_in* in, const cbfnmpc_ext_out* out, void* cuda_stream) {
  const int T = cfg ? cfg->team_size : 0;
  int rc = ext_check(c, B, cfg, in, out, T);
  if (rc) return rc;
  if (T > 1 && in->nbr_traj) return fail(CBFNMPC_ERR_ARG, "extended mode: fixed neighbour trajectories need team_size 1");
  if (in->n_nbr < 0 || in->n_nbr > 3 || (in->n_nbr > 0 && !in->nbr_traj && T == 1)) return fail(CBFNMPC_ERR_ARG, "extended mode: bad neighbour data");
  const int n_nbr = T > 1 ? T - 1 : in->n_nbr;
  if (in->n_obs + n_nbr < 1) return fail(CBFNMPC_ERR_ARG, "extended mode: at least one distance row per grid point");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  ExtIO io;
  ext_fill_io(&io, B, B, cfg, in, out);
  io.T = T; io.n_nbr = n_nbr; io.xmode = T > 1 ? 1 : (n_nbr > 0 ? 2 : 0); io.phase = 0;
  for (int i = 0; i < n_nbr && T == 1; i++) io.nbr[i] = in->nbr_traj + (size_t)i * (2 * cfg->N) * B;
  XParams XP; ext_xparams(c, cfg, &XP);
  const int n_teams = B / T;
  rc = ext_perm(c, n_teams, st);
  if (rc) return rc;
  rc = ctx_enter(c, st);
  if (rc) return rc;
  CK(cudaEventRecord(c->ev0, st));
  rc = ext_launch(c, XP, io, n_teams, c->x_perm, st, true);
  if (rc) return rc;
  CK(cudaEventRecord(c->ev1, st)); c->timed = true;
  return ctx_leave(c, st);
}

// host staging of the extended mode (grow-only)
static int ext_staging(cbfnmpc_ctx* c, size_t n_in, size_t n_out, size_t n_int) {
  if (n_in > c->x_in_cap) { if (c->x_in) CK(cudaFree(c->x_in)); CK(cudaMalloc(&c->x_in, sizeof(double) * n_in)); c->x_in_cap = n_in; }
  if (n_out > c->x_out_cap) { if (c->x_out) CK(cudaFree(c->x_out)); CK(cudaMalloc(&c->x_out, sizeof(double) * n_out)); c->x_out_cap = n_out; }
  if (n_int > c->x_int_cap) { if (c->x_int) CK(cudaFree(c->x_int)); CK(cudaMalloc(&c->x_int, sizeof(int) * n_int)); c->x_int_cap = n_int; }
  return CBFNMPC_OK;
}

int cbfnmpc_ext_solve_batch_host(cbfnmpc_ctx* c, int B, const cbfnmpc_ext_cfg* cfg, const cbfnmpc_ext_in* in, const cbfnmpc_ext_out* out) {
  const int T = cfg ? cfg->team_size : 0;
  int rc = ext_check(c, B, cfg, in, out, T);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  if (c->have_last) { CK(cudaEventSynchronize(c->ev_last)); c->have_last = false; }
  const int N = cfg->N, ND = N / 4, K = in->n_obs, nn = (T == 1 && in->nbr_traj) ? in->n_nbr : 0;
  const size_t Bs = (size_t)B;
  // inputs: x0 4 | xdes 4N | feet 8 ND | abuf 4 | obs 2K | nbr nn 2N ; outputs: X 4N | U 2N | lam 4N | slack 1 | xqp 6N | alpha 20 | abuf_next 4 | kkt 1 | pub 2N
  const size_t rin = 4 + 4 * N + 8 * ND + 4 + 2 * K + (size_t)nn * 2 * N, rout = 4 * N + 2 * N + 4 * N + 1 + 6 * N + 20 + 4 + 1 + 2 * N;
  rc = ext_staging(c, rin * Bs, rout * Bs, 8 * Bs);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  double* d = c->x_in; size_t o = 0;
  cbfnmpc_ext_in din = *in;
  auto up = [&](const double* h, size_t rows, const double** dst) -> int {
    if (rows == 0) { *dst = nullptr; return 0; }
    if (h) CK(cudaMemcpyAsync(d + o * Bs, h, sizeof(double) * rows * Bs, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(d + o * Bs, 0, sizeof(double) * rows * Bs, st));
    *dst = d + o * Bs; o += rows; return 0;
  };
  if ((rc = up(in->x0, 4, &din.x0)) || (rc = up(in->xdes, 4 * N, &din.xdes)) || (rc = up(in->feet_xy, 8 * ND, &din.feet_xy)) ||
      (rc = up(in->alpha_buf, 4, &din.alpha_buf)) || (rc = up(in->obs_xy, 2 * K, &din.obs_xy)) || (rc = up(nn ? in->nbr_traj : nullptr, (size_t)nn * 2 * N, &din.nbr_traj)))
    return rc;
  if (!nn) din.nbr_traj = nullptr;
  din.n_nbr = nn;
  CK(cudaMemcpyAsync(c->x_int, in->contact, sizeof(int) * 4 * Bs, cudaMemcpyHostToDevice, st));
  din.contact = c->x_int;
  double* q = c->x_out; size_t p = 0;
  cbfnmpc_ext_out dout;
  memset(&dout, 0, sizeof(dout));
  struct OS { double** dev; double* host; size_t rows; };
  OS os[9] = {{&dout.X, out->X, (size_t)4 * N}, {&dout.U, out->U, (size_t)2 * N}, {&dout.lam, out->lam, (size_t)4 * N}, {&dout.slack, out->slack, 1},
              {&dout.x_qp, out->x_qp, (size_t)6 * N}, {&dout.alpha, out->alpha, 20}, {&dout.alpha_buf_next, out->alpha_buf_next, 4}, {&dout.kkt, out->kkt, 1},
              {&dout.pub, out->pub, (size_t)2 * N}};
  for (int i = 0; i < 9; i++) { *os[i].dev = q + p * Bs; p += os[i].rows; }
  int* di = c->x_int + 4 * Bs;
  dout.status = di; dout.qp_iters = di + Bs; dout.nlp_iters = di + 2 * Bs; dout.team_iters = di + 3 * Bs;
  CK(cudaMemsetAsync(di, 0, sizeof(int) * 4 * Bs, st));
  rc = cbfnmpc_ext_solve_batch(c, B, cfg, &din, &dout, st);
  if (rc) return rc;
  for (int i = 0; i < 9; i++) if (os[i].host) CK(cudaMemcpyAsync(os[i].host, *os[i].dev, sizeof(double) * os[i].rows * Bs, cudaMemcpyDeviceToHost, st));
  int32_t* const oi[4] = {out->status, out->qp_iters, out->nlp_iters, out->team_iters};
  for (int i = 0; i < 4; i++) if (oi[i]) CK(cudaMemcpyAsync(oi[i], di + (size_t)i * Bs, sizeof(int) * Bs, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  c->have_last = false;
  return CBFNMPC_OK;
}

// mailboxes of cbfnmpc_ext_solve_fused: unmap the peers', free the own one (the next fused call sets them up again)
// (An exporter must not free memory that an importer still has open, and the ranks of a communicator do not leave it at the same
//  moment: when a communicator goes, only the mappings of the PEERS' mailboxes are closed; the own mailbox stays allocated and is
//  exported again to the next communicator.  It is freed when it has to grow -- after a barrier over the ranks, ext_fused_setup --
//  and with the context.)
static void ext_fused_unmap(cbfnmpc_ctx* c) {
  for (int r = 0; r < 8; r++) { if (c->f_peer[r] && c->f_peer[r] != c->f_mbox) cudaIpcCloseMemHandle(c->f_peer[r]); c->f_peer[r] = nullptr; }
  c->f_mapped = false; c->f_disabled = false;
}
static void ext_fused_release(cbfnmpc_ctx* c) {
  ext_fused_unmap(c);
  if (c->f_mbox) cudaFree(c->f_mbox);
  c->f_mbox = nullptr; c->f_mbox_cap = 0;
}

// ---- agent-sharded teams: the path's one collective
int cbfnmpc_comm_unique_id(void* id128) {
  if (!id128) return fail(CBFNMPC_ERR_ARG, "null argument");
  int rc = xnccl_load();
  if (rc) return rc;
  xnccl_uid u;
  NK(g_nccl.GetUniqueId(&u));
  memcpy(id128, u.internal, 128);
  return CBFNMPC_OK;
}
int cbfnmpc_comm_init(cbfnmpc_ctx* c, int nranks, int rank, const void* id128) {
  if (!c || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_comm_init: bad argument");
  int rc = xnccl_load();
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  if (c->comm) { g_nccl.CommDestroy((xnccl_comm)c->comm); c->comm = nullptr; }
  CK(cudaStreamSynchronize(c->stream)); ext_fused_unmap(c);
  xnccl_uid u; memcpy(u.internal, id128, 128);
  xnccl_comm cm = nullptr;
  NK(g_nccl.CommInitRank(&cm, nranks, u, rank));
  c->comm = cm; c->comm_size = nranks; c->comm_rank = rank;
  return CBFNMPC_OK;
}
int cbfnmpc_comm_destroy(cbfnmpc_ctx* c) {
  if (!c) return fail(CBFNMPC_ERR_ARG, "null argument");
  if (c->comm && g_nccl.h) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); g_nccl.CommDestroy((xnccl_comm)c->comm); }
  cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); ext_fused_unmap(c);
  c->comm = nullptr; c->comm_size = 1; c->comm_rank = 0;
  return CBFNMPC_OK;
}

// This rank holds members [rank * mpr, (rank + 1) * mpr) of EVERY one of n_teams teams (mpr = team_size / nranks), member-major:
// instance b = (member - rank * mpr) * n_teams + team.  One launch per interior-point iteration; between the launches the published
// trajectories of all members are all-gathered (NCCL over NVLink; a device copy when nranks = 1).  The launch of iteration i + 1 is
// enqueued before the host looks at the "teams still iterating" counter of iteration i (two pinned slots), so the stream never
// drains; iterations of teams that have finished are no-ops.
int cbfnmpc_ext_solve_sharded(cbfnmpc_ctx* c, int n_teams, const cbfnmpc_ext_cfg* cfg, const cbfnmpc_ext_in* in, const cbfnmpc_ext_out* out,
                              cbfnmpc_ext_stats* stats, void* cuda_stream) {
  if (!c || !cfg) return fail(CBFNMPC_ERR_ARG, "null argument");
  const int T = cfg->team_size, G = c->comm ? c->comm_size : 1, rank = c->comm ? c->comm_rank : 0;
  if (T < 2 || T % G != 0 || G > 8) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_ext_solve_sharded: team_size must be a multiple (>= 2) of the communicator's size (<= 8)");
  const int mpr = T / G, B = mpr * n_teams;
  int rc = ext_check(c, B, cfg, in, out, 1);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  const int N = cfg->N, n_nbr = T - 1, K = in->n_obs + n_nbr;
  const size_t blk = (size_t)(2 * N + 2) * n_teams;   // one member's published block
  const size_t n_state = (size_t)ext_nstate(K) * (size_t)B * N, n_pub = blk * mpr, n_gath = blk * T;
  if (n_state > c->x_state_cap) { CK(cudaStreamSynchronize(st)); if (c->x_state) CK(cudaFree(c->x_state)); CK(cudaMalloc(&c->x_state, sizeof(double) * n_state)); c->x_state_cap = n_state; }
  if (n_gath + n_pub > c->x_pub_cap) { CK(cudaStreamSynchronize(st)); if (c->x_pub) CK(cudaFree(c->x_pub)); CK(cudaMalloc(&c->x_pub, sizeof(double) * (n_gath + n_pub))); c->x_pub_cap = n_gath + n_pub; }
  if (!c->x_active) { CK(cudaMalloc(&c->x_active, sizeof(int) * 18)); CK(cudaHostAlloc(&c->x_active_h, sizeof(int) * 18, cudaHostAllocDefault)); CK(cudaEventCreateWithFlags(&c->x_ev[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->x_ev[1], cudaEventDisableTiming)); CK(cudaEventCreate(&c->x_t0)); CK(cudaEventCreate(&c->x_t1)); }
  double* pub = c->x_pub; double* gathered = c->x_pub + n_pub;
  ExtIO io;
  ext_fill_io(&io, B, B, cfg, in, out);
  io.pub = pub; io.gathered = gathered; io.n_teams = n_teams; io.member0 = rank * mpr;
  io.T = 1; io.n_nbr = n_nbr; io.xmode = 3; io.state = c->x_state;
  XParams XP; ext_xparams(c, cfg, &XP);
  rc = ext_perm(c, B, st);
  if (rc) return rc;
  rc = ctx_enter(c, st);
  if (rc) return rc;
  auto gather = [&]() -> int {
    if (G == 1) CK(cudaMemcpyAsync(gathered, pub, sizeof(double) * n_pub, cudaMemcpyDeviceToDevice, st));
    else NK(g_nccl.AllGather(pub, gathered, n_pub, XNCCL_DOUBLE, (xnccl_comm)c->comm, st));
    return CBFNMPC_OK;
  };
  CK(cudaEventRecord(c->x_t0, st));
  io.phase = 1; io.n_active = nullptr;
  if ((rc = ext_launch(c, XP, io, B, c->x_perm, st, true)) || (rc = gather())) return rc;
  io.phase = 2;
  if ((rc = ext_launch(c, XP, io, B, c->x_perm, st, false)) || (rc = gather())) return rc;
  const int cap = cfg->max_team_iters + 2;
  int it = 0, done_at = -1;
  for (; it < cap; it++) {
    const int s = it & 1;
    CK(cudaMemsetAsync(c->x_active + s, 0, sizeof(int), st));
    io.phase = 3; io.n_active = c->x_active + s;
    if ((rc = ext_launch(c, XP, io, B, c->x_perm, st, false)) || (rc = gather())) return rc;
    // The counter that ends the loop must be the same on every rank (a rank that stopped calling the all-gather would leave the
    // others waiting in it): the ranks' local counts are gathered and summed.  Local counts alone differ when a rank holds no
    // RUNNING member of a team that still iterates elsewhere -- a member with a bad contact pattern and no team mate on its rank.
    if (G > 1) {
      NK(g_nccl.AllGather(c->x_active + s, c->x_active + 2 + 8 * s, 1, XNCCL_INT32, (xnccl_comm)c->comm, st));
      CK(cudaMemcpyAsync(c->x_active_h + 2 + 8 * s, c->x_active + 2 + 8 * s, sizeof(int) * G, cudaMemcpyDeviceToHost, st));
    } else {
      CK(cudaMemcpyAsync(c->x_active_h + s, c->x_active + s, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaEventRecord(c->x_ev[s], st));
    if (it >= 1) {   // look at the previous iteration's counter while this one runs
      CK(cudaEventSynchronize(c->x_ev[s ^ 1]));
      int active = 0;
      if (G > 1) { for (int r = 0; r < G; r++) active += c->x_active_h[2 + 8 * (s ^ 1) + r]; }
      else active = c->x_active_h[s ^ 1];
      if (active == 0) { done_at = it - 1; it++; break; }
    }
  }
  io.phase = 4; io.n_active = nullptr;
  if ((rc = ext_launch(c, XP, io, B, c->x_perm, st, false))) return rc;
  CK(cudaEventRecord(c->x_t1, st));
  if (stats) {
    CK(cudaEventSynchronize(c->x_t1));
    float ms = 0.0f; CK(cudaEventElapsedTime(&ms, c->x_t0, c->x_t1));
    stats->launches = it + 3; stats->iterations = done_at >= 0 ? done_at + 1 : it; stats->exchanges = it + 2;
    stats->bytes_per_exchange = (long long)(sizeof(double) * n_pub); stats->total_ms = ms; stats->nranks = G; stats->fused = 0;
  }
  return ctx_leave(c, st);
}

// ---- the same teams with the exchange INSIDE one launch: peer-memory mailboxes instead of a collective per iteration
// Collective over the communicator (every rank calls it with the same sizes).  Mailbox of a rank: [2 parities][T members][2N + 2
// rows][n_teams] doubles + [2][T][n_teams] int flags; the ranks exchange CUDA IPC handles of their mailboxes once (through the
// library's own ncclAllGather) and map each other's.  Returns with c->f_disabled set -- on EVERY rank -- when some rank could not map
// a peer (the caller then uses the per-iteration path).
static int ext_fused_setup(cbfnmpc_ctx* c, size_t need, cudaStream_t st) {
  const int G = c->comm ? c->comm_size : 1, rank = c->comm ? c->comm_rank : 0;
  if (G > 8) { c->f_disabled = true; return CBFNMPC_OK; }
  if (!c->f_ctl) { CK(cudaMalloc(&c->f_ctl, sizeof(int) * 4)); CK(cudaMalloc(&c->f_bar, sizeof(double) * 16 * 9)); }
  if (c->f_mbox && c->f_mapped && need <= c->f_mbox_cap) return CBFNMPC_OK;
  CK(cudaStreamSynchronize(st));
  // (re)mapping: every rank first closes its mappings of the peers' mailboxes; when the own mailbox has to grow all ranks meet
  // before anybody frees its own (an exporter must not free memory that an importer still has open)
  ext_fused_unmap(c);
  if (!c->f_mbox || need > c->f_mbox_cap) {   // (same outcome on every rank: same call history)
    if (G > 1) { NK(g_nccl.AllGather(c->f_bar, c->f_bar + 8, 1, XNCCL_DOUBLE, (xnccl_comm)c->comm, st)); CK(cudaStreamSynchronize(st)); }
    if (c->f_mbox) CK(cudaFree(c->f_mbox));
    c->f_mbox = nullptr; c->f_mbox_cap = 0;
    CK(cudaMalloc(&c->f_mbox, sizeof(double) * need)); c->f_mbox_cap = need;
    CK(cudaMemset(c->f_mbox, 0, sizeof(double) * need));
  }
  c->f_peer[rank] = c->f_mbox;
  if (G == 1) { c->f_mapped = true; return CBFNMPC_OK; }
  // handles: 64 bytes = 8 doubles per rank, slot 0 = own, slots 1 .. G = gathered
  cudaIpcMemHandle_t h, all[8];
  bool ok = cudaIpcGetMemHandle(&h, c->f_mbox) == cudaSuccess && getenv("CBFNMPC_NO_IPC") == nullptr;   // (the variable: to test the fallback)
  if (!ok) { cudaGetLastError(); memset(&h, 0, sizeof(h)); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  CK(cudaMemcpyAsync(c->f_bar, &h, 64, cudaMemcpyHostToDevice, st));
  NK(g_nccl.AllGather(c->f_bar, c->f_bar + 8, 8, XNCCL_DOUBLE, (xnccl_comm)c->comm, st));
  CK(cudaMemcpyAsync(all, c->f_bar + 8, 64 * (size_t)G, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int r = 0; r < G && ok; r++) {
    if (r == rank) continue;
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
    c->f_peer[r] = (double*)p;
  }
  // did everybody get there?
  const double mine = ok ? 1.0 : 0.0; double oks[8];
  CK(cudaMemcpyAsync(c->f_bar, &mine, sizeof(double), cudaMemcpyHostToDevice, st));
  NK(g_nccl.AllGather(c->f_bar, c->f_bar + 8, 1, XNCCL_DOUBLE, (xnccl_comm)c->comm, st));
  CK(cudaMemcpyAsync(oks, c->f_bar + 8, sizeof(double) * G, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int r = 0; r < G; r++) if (oks[r] == 0.0) ok = false;
  if (!ok) { ext_fused_unmap(c); c->f_disabled = true; }
  else c->f_mapped = true;
  return CBFNMPC_OK;
}

int cbfnmpc_ext_solve_fused(cbfnmpc_ctx* c, int n_teams, const cbfnmpc_ext_cfg* cfg, const cbfnmpc_ext_in* in, const cbfnmpc_ext_out* out,
                            cbfnmpc_ext_stats* stats, void* cuda_stream) {
  if (!c || !cfg) return fail(CBFNMPC_ERR_ARG, "null argument");
  const int T = cfg->team_size, G = c->comm ? c->comm_size : 1, rank = c->comm ? c->comm_rank : 0;
  if (T < 2 || T > 4 || T % G != 0) return fail(CBFNMPC_ERR_ARG, "cbfnmpc_ext_solve_fused: team_size (2 .. 4) must be a multiple of the communicator's size");
  const int mpr = T / G, B = mpr * n_teams;
  int rc = ext_check(c, B, cfg, in, out, 1);
  if (rc) return rc;
  CK(cudaSetDevice(c->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
  const int N = cfg->N, n_nbr = T - 1, K = in->n_obs + n_nbr;
  // every CTA must be resident at once: a member waits for members that run in other CTAs, here and on the peers
  const size_t smem = sizeof(double) * (size_t)ext_npriv(K) * XNT;
  const int tpw = XNT / N;
  int per_sm = 0;
  if (smem <= 227 * 1024) {
    if (N == 4) { CK(cudaFuncSetAttribute(cbfnmpc_ext_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cbfnmpc_ext_fused_kernel<4>, XNT, smem)); }
    else { CK(cudaFuncSetAttribute(cbfnmpc_ext_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cbfnmpc_ext_fused_kernel<8>, XNT, smem)); }
  }
  const size_t blk = (size_t)(2 * N + 2) * n_teams;
  const size_t need = 2 * (size_t)T * blk + ((size_t)2 * T * n_teams + 1) / 2 + 2;
  // teams per launch: what is resident at once (the class lists pad two classes: one spare CTA); larger batches run in slices of teams
  const int cap_inst = (per_sm * c->sm_count - 2) * tpw;
  const int ns_max = cap_inst / mpr;
  bool fused = ns_max >= 1 && !c->f_disabled;
  if (fused) {   // (the capacity test has the same outcome on every rank: same sizes, same devices)
    rc = ext_fused_setup(c, need, st);
    if (rc) return rc;
    fused = !c->f_disabled;
  }
  if (!fused) {
    rc = cbfnmpc_ext_solve_sharded(c, n_teams, cfg, in, out, stats, cuda_stream);
    if (stats) stats->fused = 0;
    return rc;
  }
  if (!c->x_active) { CK(cudaMalloc(&c->x_active, sizeof(int) * 18)); CK(cudaHostAlloc(&c->x_active_h, sizeof(int) * 18, cudaHostAllocDefault)); CK(cudaEventCreateWithFlags(&c->x_ev[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->x_ev[1], cudaEventDisableTiming)); CK(cudaEventCreate(&c->x_t0)); CK(cudaEventCreate(&c->x_t1)); }
  ExtIO io;
  ext_fill_io(&io, B, B, cfg, in, out);
  io.n_teams = n_teams; io.member0 = rank * mpr;
  io.T = 1; io.n_nbr = n_nbr; io.xmode = 4; io.phase = 0;
  for (int r = 0; r < G; r++) io.mbox[r] = c->f_peer[r];
  io.n_grp = G; io.fused_err = c->f_ctl; io.fused_its = c->f_ctl + 1;
  XParams XP; ext_xparams(c, cfg, &XP);
  rc = ext_perm(c, B, st);
  if (rc) return rc;
  rc = ctx_enter(c, st);
  if (rc) return rc;
  // own flags and control words to zero, then a barrier over the communicator: nobody writes into a mailbox before its owner has
  // cleared it, and nobody starts before everybody has finished the previous call
  CK(cudaMemsetAsync(c->f_mbox + 2 * (size_t)T * blk, 0, sizeof(int) * 2 * (size_t)T * n_teams, st));
  CK(cudaMemsetAsync(c->f_ctl, 0, sizeof(int) * 4, st));
  if (G > 1) NK(g_nccl.AllGather(c->f_bar, c->f_bar + 8, 1, XNCCL_DOUBLE, (xnccl_comm)c->comm, st));
  CK(cudaEventRecord(c->x_t0, st));
  int n_launch = 0;
  const int n_slices = (n_teams + ns_max - 1) / ns_max;
  for (int k = 0; k < n_slices; k++) {   // (equal slices; a slice's teams are independent of every other slice's)
    const int t0 = (int)((long long)n_teams * k / n_slices), t1 = (int)((long long)n_teams * (k + 1) / n_slices), ns = t1 - t0;
    const int Bk = mpr * ns, items = (Bk + tpw - 1) / tpw + 1;
    io.slice_t0 = t0; io.slice_ns = ns;
    int* perm = c->x_perm; int* counts = perm + 2 * (size_t)Bk;
    cbfnmpc_ext_classify_kernel<<<1, 1024, 0, st>>>(io, Bk, perm, counts);
    CK(cudaGetLastError());
    if (N == 4) cbfnmpc_ext_fused_kernel<4><<<items, XNT, smem, st>>>(XP, io, Bk, perm, counts);
    else cbfnmpc_ext_fused_kernel<8><<<items, XNT, smem, st>>>(XP, io, Bk, perm, counts);
    CK(cudaGetLastError());
    c->launches += 2; n_launch += 2;
  }
  CK(cudaEventRecord(c->x_t1, st));
  int ctl[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(ctl, c->f_ctl, sizeof(int) * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (stats) {
    float ms = 0.0f; CK(cudaEventElapsedTime(&ms, c->x_t0, c->x_t1));
    stats->launches = n_launch; stats->iterations = ctl[1] + 1; stats->exchanges = ctl[1] + 3;
    stats->bytes_per_exchange = (long long)(sizeof(double) * blk * mpr) * (G - 1); stats->total_ms = ms; stats->nranks = G; stats->fused = 1;
  }
  rc = ctx_leave(c, st);
  if (ctl[0] != 0) return fail(CBFNMPC_ERR_CUDA, "cbfnmpc_ext_solve_fused: a member waited in vain for a neighbour's publication (watchdog)");
  return rc;
}

