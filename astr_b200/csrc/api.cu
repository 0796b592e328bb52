// astr_b200/csrc/api.cu -- the C ABI of include/astr_gpu.h: context, line-operator
// tables, halo exchange, and the Runge-Kutta stage built from sweep.cu / pointwise.cu.
//
// Stage order and semantics follow src/mainloop.F90:396-482 (time_integration_rk):
//   filterq -> (qrhs=0) -> [boucon: caller] -> qswap -> gradcal -> rhscal -> RK update ->
//   [spongefilter: caller] -> updatefvar
#include "../../include/astr_gpu.h"
#include "common.cuh"
#include "pointwise.cuh"
#include "geom.cuh"
#include <nccl.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>
#include <mutex>

// -------------------------------------------------------------------------------------
// errors / counters
// -------------------------------------------------------------------------------------
static std::string g_err;
static long long g_launches = 0;

int astr_fail(const char* what, cudaError_t e, const char* file, int line) {
  char buf[1024];
  snprintf(buf, sizeof buf, "astr_gpu: %s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
  g_err = buf;
  return 1;
}
int astr_fail_msg(const char* msg) {
  g_err = std::string("astr_gpu: ") + msg;
  return 1;
}
void astr_count_launch(int n) { g_launches += n; }

// -------------------------------------------------------------------------------------
// NCCL through dlopen: the library loads without NCCL for single-GPU use and shares the
// host process' libnccl.so.2 (torch's bundled copy, or the MPI application's).
// -------------------------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) return false;
#define SYM(field, name) *(void**)(&field) = dlsym(h, name); if (!field) return false;
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return true;
  }
};
static NcclApi g_nccl;

#define NCCL_OK(call)                                                              \
  do {                                                                             \
    ncclResult_t r_ = (call);                                                      \
    if (r_ != ncclSuccess) {                                                       \
      g_err = std::string("astr_gpu: ") + #call + " failed: " + g_nccl.GetErrorString(r_); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

// -------------------------------------------------------------------------------------
// host-side line operators (product code; the reference builds the same tables in
// fd_scheme_initiate src/derivative.F90:63-158 and compact_filter_initiate
// src/filter.F90:31-100, pre-factored by tridiagonal_thomas_proprocess
// src/commfunc.F90:752-774)
// -------------------------------------------------------------------------------------
struct HostOp {
  int first_node = 0, nrows = 0, ntype = 0, n = 0, C = 1, nsf = 0, nsl = 0;
  std::vector<double> a, c, ac1, ac2, ac3, pf, qb;
  double* d_tab = nullptr;  // 5*nrows doubles on the device
  LinePlan plan;            // register-resident engine (sweep2.cu); plan.ok == 0: not applicable
  LineOp dev() const {
    LineOp o;
    o.ac1 = d_tab; o.ac2 = d_tab + nrows; o.ac3 = d_tab + 2 * nrows;
    o.pf = d_tab + 3 * nrows; o.qb = d_tab + 4 * nrows;
    o.first_node = first_node; o.nrows = nrows; o.ntype = ntype; o.n = n; o.C = C;
    o.nsf = nsf; o.nsl = nsl;
    return o;
  }
};

static void factorise(HostOp& h) {
  const int N = h.nrows;
  h.ac1.assign(N, 0.0); h.ac2.assign(N, 0.0); h.ac3.assign(N, 0.0);
  h.ac1[0] = h.c[0];
  h.ac2[0] = 1.0;  // row 0 is not scaled by the forward sweep (commfunc.F90:800-804)
  h.ac3[0] = 0.0;
  for (int i = 1; i < N; ++i) {
    h.ac1[i] = h.c[i] / (1.0 - h.a[i] * h.ac1[i - 1]);
    h.ac2[i] = 1.0 / (1.0 - h.a[i] * h.ac1[i - 1]);
    h.ac3[i] = h.a[i] / (1.0 - h.a[i] * h.ac1[i - 1]);
  }
  // chunk propagation products of the partitioned solve (sweep.cu)
  h.C = astr_sweep_max_chunks(N);
  h.pf.assign(N, 0.0); h.qb.assign(N, 0.0);
  for (int c = 0; c < h.C; ++c) {
    const int ra = chunk_start(c, N, h.C), rb = chunk_start(c + 1, N, h.C) - 1;
    double p = 1.0;
    for (int r = ra; r <= rb; ++r) { p = p * (-h.ac3[r]); h.pf[r] = p; }
    double q = 1.0;
    for (int r = rb; r >= ra; --r) { q = q * (-h.ac1[r]); h.qb[r] = q; }
  }
}

// dir 0 (i lines): the register engine needs 16-byte aligned chunk windows (linecore.h, align_even)
static void build_deriv(HostOp& h, int ntype, int n, int dir) {
  h.ntype = ntype; h.n = n;
  build_lhs(OP_DERIV, ntype, n, 0.0, h.a, h.c, h.first_node, h.nsf, h.nsl);
  h.nrows = (int)h.a.size();
  factorise(h);
  build_line_plan(h.plan, OP_DERIV, ntype, n, h.first_node, h.nsf, h.nsl, h.a, h.c, ASTR_NWMAX, dir == 0);
}

static void build_filter(HostOp& h, int ntype, int n, double alfa, int dir) {
  h.ntype = ntype; h.n = n;
  build_lhs(OP_FILTER, ntype, n, alfa, h.a, h.c, h.first_node, h.nsf, h.nsl);
  h.nrows = (int)h.a.size();
  factorise(h);
  build_line_plan(h.plan, OP_FILTER, ntype, n, h.first_node, h.nsf, h.nsl, h.a, h.c, ASTR_NWMAX, dir == 0);
}

// compact_flux_initiate (src/flux.F90:32-118) for flux_uw ('+', OP_FLUXP) / flux_dw ('-', OP_FLUXM)
static void build_flux(HostOp& h, int optype, int ntype, int n, double bfacmpld, int dir) {
  h.ntype = ntype; h.n = n;
  build_lhs(optype, ntype, n, bfacmpld, h.a, h.c, h.first_node, h.nsf, h.nsl);
  h.nrows = (int)h.a.size();
  factorise(h);
  build_line_plan(h.plan, optype, ntype, n, h.first_node, h.nsf, h.nsl, h.a, h.c, ASTR_NWMAX, dir == 0);
}

// -------------------------------------------------------------------------------------
// context
// -------------------------------------------------------------------------------------
enum ProfCat { PC_FILTER_I, PC_FILTER_J, PC_FILTER_K, PC_HALO, PC_GRAD_I, PC_GRAD_J, PC_GRAD_K, PC_VISC,
               PC_FLUX, PC_DIV_I, PC_DIV_J, PC_DIV_K, PC_RK, PC_FVAR, PC_XPACK, PC_XNCCL, PC_XUNPACK, PC_COUNT };

struct ProfSpan { int cat; cudaEvent_t a, b; };

struct Ctx {
  astr_cfg cfg;
  Layout L;
  Thermo th;
  double* pool = nullptr;        // S_CORE fields
  double* scr = nullptr;         // 15 scratch fields, lazily allocated
  // crash control (src/mainloop.F90:709-1198)
  double* crinod = nullptr;      // critical nodes as 0/1, one field; nullptr until the crash control is first used
  double* bak[2] = {nullptr, nullptr};   // databakup: dat_a%q, dat_b%q (5 fields each)
  int bak_counter[2] = {0, 0};
  char datpnt = 'o';
  unsigned long long* d_count = nullptr; // device counters (2)
  long long* d_list = nullptr;           // crashfix: keys of the flagged nodes
  bool spg_global = false;       // spg_def='circl'
  double* spg_global_coef = nullptr;
  double* rhsav = nullptr;       // rk4: 5 fields (src/mainloop.F90:394), allocated at the first rk4 update
  HostOp fd[3], fl[3];
  HostOp fxp[3], fxm[3];         // flux_uw_* / flux_dw_* (conschm 543 only)
  double* up = nullptr;          // UP_TOTAL fields of the upwind path, lazily allocated
  bool upwind() const { return (cfg.conschm / 100) % 2 == 1; }
  FilterCoef fc;
  double* d_partial = nullptr;   // stats partial sums
  double* d_out2 = nullptr;
  cudaStream_t st = nullptr;
  cudaStream_t cur = nullptr;    // stream the halo exchanges are issued on: st, or xst while an exchange overlaps compute
  cudaStream_t xst = nullptr;
  cudaEvent_t ev_shell = nullptr, ev_xdone = nullptr;
  bool have_metrics = false, have_grad = false;
  bool rhs_in_g = false;         // qrhs lives as three directional derivatives in the G slots
  bool sigma_partial = false;    // sigma/qflux are in memory only on the face shells (fused rhscal)
  double force[3] = {0, 0, 0};
  double* ycoord = nullptr;      // x(:,:,:,2) for src_chan (flowtype channel only)
  struct Sponge { bool on = false; int beg = -1, end = -2; double* coef = nullptr; } spg[6];   // i0, im, j0, jm, k0, km
  bool any_sponge() const { if (spg_global) return true; for (const Sponge& s : spg) if (s.on) return true; return false; }
  double* stage = nullptr;       // dense host-layout staging field of the upload path
  double* dense[2] = {nullptr, nullptr};            // checkpoint staging: dense node arrays
  cudaEvent_t ev_dense[4] = {nullptr, nullptr, nullptr, nullptr};
  double* d_inflow = nullptr;    // vel_in(0:jm,0:km,3) | tmp_in(0:jm,0:km) | tmp_prof(0:jm)  (bctype(1)=11)
  double* d_src = nullptr;       // [0..3] bulk integrals, [4..7] (force, force.ubulk)
  bool src_pending = false;      // src_chan's term is not in the G slots: consumers add d_src+4
  const double* src() const { return src_pending ? d_src + 4 : nullptr; }
  // multi-block
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  double* xbuf[4] = {nullptr, nullptr, nullptr, nullptr};  // send lo, send hi, recv lo, recv hi
  size_t xbuf_doubles = 0;
  // peer-to-peer exchange (fused pack + NVLink stores): one IPC-exported arena per rank
  bool p2p = false;
  char* xarena = nullptr;                 // XHeader + receive windows [direction][side]
  char* peer_arena[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // per neighbour slot (nbr[])
  int peer_dims[6][3];
  std::vector<void*> ipc_opened;
  unsigned long long xseq[3] = {0, 0, 0};
  long long xtimeout_cycles = 0;          // cfg.xchg_timeout_ms in SM clock cycles (<= 0: wait forever)
  bool xfailed = false;                   // sticky: a peer-memory exchange timed out
  // profiling
  bool profile = false;
  std::vector<ProfSpan> spans;
  std::vector<cudaEvent_t> free_events;
  double prof_ms[PC_COUNT] = {0};
  long long prof_n[PC_COUNT] = {0};
  double* slot(int s) const { return (s >= S_SCR ? scr + (size_t)(s - S_SCR) * L.fstride : pool + (size_t)s * L.fstride); }
};
static Ctx* g = nullptr;

#define NEED_CTX()                                           \
  do {                                                       \
    if (!g) return astr_fail_msg("astr_gpu_init not called"); \
  } while (0)
#define TRY(x)                  \
  do {                          \
    int rc_ = (x);              \
    if (rc_) return rc_;        \
  } while (0)

struct ProfScope {
  int cat; bool on; cudaEvent_t a = nullptr, b = nullptr; cudaStream_t st = nullptr;
  static cudaEvent_t get() {
    if (!g->free_events.empty()) { cudaEvent_t e = g->free_events.back(); g->free_events.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  explicit ProfScope(int c) : cat(c), on(g && g->profile) {
    if (on) { a = get(); b = get(); st = g->cur; cudaEventRecord(a, st); }
  }
  ~ProfScope() {
    if (on) { cudaEventRecord(b, st); g->spans.push_back({cat, a, b}); }
  }
};
static void prof_collect() {
  for (auto& s : g->spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { g->prof_ms[s.cat] += ms; g->prof_n[s.cat] += 1; }
    g->free_events.push_back(s.a); g->free_events.push_back(s.b);
  }
  g->spans.clear();
}

static int ensure_scratch() {
  if (g->scr) return 0;
  const size_t bytes = (size_t)(S_TOTAL - S_SCR) * g->L.fstride * sizeof(double);
  CUDA_OK(cudaMalloc(&g->scr, bytes));
  CUDA_OK(cudaMemsetAsync(g->scr, 0, bytes, g->st));
  TRY(astr_sweep2_register_pool(1, g->scr, S_TOTAL - S_SCR, g->L));
  return 0;
}

static FieldList fields(int s0, int n) {
  FieldList fl;
  fl.nf = n;
  for (int i = 0; i < n; ++i) fl.f[i] = g->slot(s0 + i);
  return fl;
}

// -------------------------------------------------------------------------------------
// halo exchange of one direction (dataswap / qswap / datasync, src/parallel.F90)
// -------------------------------------------------------------------------------------
constexpr int XMAX_PLANES = 45;            // 9 fields x 5 planes (sigma+qflux) >= 5 fields x 6 planes (qswap)
static int exchange_dir_p2p(const FieldList& fl, int d, int l0, int l1, int phases = 3);

static int exchange_dir(const FieldList& fl, int d, int mode) {
  const astr_cfg& c = g->cfg;
  if (c.size[d] == 1) {
    if (!c.lhomo[d]) return 0;
    ProfScope ps(PC_HALO);
    return pw_halo_wrap(g->L, fl, d, mode, g->cur);
  }
  if (!g->comm) return astr_fail_msg("multi-block exchange needs astr_gpu_comm_init");
  const Layout& L = g->L;
  const int l0 = (mode == XMODE_SWAP) ? 1 : 0, l1 = (mode == XMODE_SYNC) ? 0 : ASTR_HM;
  if (g->p2p && fl.nf * (l1 - l0 + 1) <= XMAX_PLANES) return exchange_dir_p2p(fl, d, l0, l1);
  const int n1 = (d == 0) ? L.jm + 1 : L.im + 1, n2 = (d == 2) ? L.jm + 1 : L.km + 1;
  const size_t cnt = (size_t)(l1 - l0 + 1) * n1 * n2 * fl.nf;
  if (cnt > g->xbuf_doubles) {
    for (auto& b : g->xbuf) { if (b) cudaFree(b); b = nullptr; }
    const size_t want = cnt;
    for (auto& b : g->xbuf) CUDA_OK(cudaMalloc(&b, want * sizeof(double)));
    g->xbuf_doubles = want;
  }
  const int lo = c.nbr[2 * d], hi = c.nbr[2 * d + 1];
  {
    ProfScope ps(PC_XPACK);
    if (lo >= 0) TRY(pw_pack(L, fl, d, 0, l0, l1, g->xbuf[0], g->cur));
    if (hi >= 0) TRY(pw_pack(L, fl, d, 1, l0, l1, g->xbuf[1], g->cur));
  }
  ProfScope* pn = new ProfScope(PC_XNCCL);
  NCCL_OK(g_nccl.GroupStart());
  if (lo >= 0) NCCL_OK(g_nccl.Send(g->xbuf[0], cnt, ncclDouble, lo, g->comm, g->cur));
  if (hi >= 0) NCCL_OK(g_nccl.Send(g->xbuf[1], cnt, ncclDouble, hi, g->comm, g->cur));
  // when lo == hi (two blocks, periodic) the peer's first message is its low-side buffer,
  // which belongs in my high halo: post that receive first.
  if (hi >= 0) NCCL_OK(g_nccl.Recv(g->xbuf[3], cnt, ncclDouble, hi, g->comm, g->cur));
  if (lo >= 0) NCCL_OK(g_nccl.Recv(g->xbuf[2], cnt, ncclDouble, lo, g->comm, g->cur));
  NCCL_OK(g_nccl.GroupEnd());
  delete pn;
  ProfScope pu(PC_XUNPACK);
  if (hi >= 0) TRY(pw_unpack(L, fl, d, 1, l0, l1, g->xbuf[3], g->cur));
  if (lo >= 0) TRY(pw_unpack(L, fl, d, 0, l0, l1, g->xbuf[2], g->cur));
  return 0;
}

int astr_exchange_slots(const int* slots, int nf, int d, int mode) {
  FieldList fl;
  fl.nf = nf;
  for (int i = 0; i < nf; ++i) fl.f[i] = g->slot(slots[i]);
  return exchange_dir(fl, d, mode);
}
double* astr_slot_ptr(int slot) { return g->slot(slot); }

// gridsendrecv of one direction (src/parallel.F90:2780-3035): x lives in slots S_G+0..2
int astr_xhalo_exchange(int d) {
  const astr_cfg& c = g->cfg;
  const Layout& L = g->L;
  double* x3[3] = {g->slot(S_G), g->slot(S_G + 1), g->slot(S_G + 2)};
  if (c.size[d] == 1) return geom_xhalo_single(L, x3, d, g->st);
  if (!g->comm) return astr_fail_msg("multi-block gridgeom needs astr_gpu_comm_init");
  const int n1 = (d == 0) ? L.jm + 1 : L.im + 1, n2 = (d == 2) ? L.jm + 1 : L.km + 1;
  const size_t cnt = (size_t)3 * ASTR_HM * n1 * n2;
  if (cnt > g->xbuf_doubles) {
    for (auto& b : g->xbuf) { if (b) cudaFree(b); b = nullptr; }
    for (auto& b : g->xbuf) CUDA_OK(cudaMalloc(&b, cnt * sizeof(double)));
    g->xbuf_doubles = cnt;
  }
  const int lo = c.nbr[2 * d], hi = c.nbr[2 * d + 1];
  TRY(geom_xhalo_pack(L, x3, d, g->xbuf[0], g->xbuf[1], g->st));
  NCCL_OK(g_nccl.GroupStart());
  if (lo >= 0) NCCL_OK(g_nccl.Send(g->xbuf[0], cnt, ncclDouble, lo, g->comm, g->st));
  if (hi >= 0) NCCL_OK(g_nccl.Send(g->xbuf[1], cnt, ncclDouble, hi, g->comm, g->st));
  if (hi >= 0) NCCL_OK(g_nccl.Recv(g->xbuf[3], cnt, ncclDouble, hi, g->comm, g->st));
  if (lo >= 0) NCCL_OK(g_nccl.Recv(g->xbuf[2], cnt, ncclDouble, lo, g->comm, g->st));
  NCCL_OK(g_nccl.GroupEnd());
  return geom_xhalo_unpack(L, x3, d, lo >= 0 ? g->xbuf[2] : nullptr, hi >= 0 ? g->xbuf[3] : nullptr, g->st);
}

// communicator, peer mappings and exchange buffers (finalize, or a repeated comm_init)
static void release_comm() {
  if (!g) return;
  if (g->st) cudaStreamSynchronize(g->st);
  for (void* p : g->ipc_opened) cudaIpcCloseMemHandle(p);
  g->ipc_opened.clear();
  for (auto& pa : g->peer_arena) pa = nullptr;
  if (g->xarena) { cudaFree(g->xarena); g->xarena = nullptr; }
  for (auto& b : g->xbuf) { if (b) cudaFree(b); b = nullptr; }
  g->xbuf_doubles = 0;
  if (g->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g->comm);
  g->comm = nullptr;
  g->p2p = false; g->xfailed = false;
  for (auto& q : g->xseq) q = 0;
}

// ---- peer-to-peer exchange: arena layout, IPC set-up ---------------------------------------
constexpr size_t XHEADER_BYTES = 1024;
struct XHeader {
  unsigned long long ready[6];             // [direction*2 + side] written by the neighbour on that side
  unsigned long long ack[6];               // [direction*2 + side] written by the neighbour on that side
  unsigned int csend[6], crecv[6];         // CTA completion counters (local)
  unsigned int err;
};
static size_t xwin_doubles(const int dims[3], int d) {
  const size_t n1 = (d == 0) ? dims[1] + 1 : dims[0] + 1, n2 = (d == 2) ? dims[1] + 1 : dims[2] + 1;
  return (size_t)XMAX_PLANES * n1 * n2;
}
static size_t xwin_offset(const int dims[3], int d, int side) {   // bytes from the arena start
  size_t off = XHEADER_BYTES;
  for (int dd = 0; dd < d; ++dd) off += 2 * xwin_doubles(dims, dd) * sizeof(double);
  return off + (size_t)side * xwin_doubles(dims, d) * sizeof(double);
}
static size_t xarena_bytes(const int dims[3]) { return xwin_offset(dims, 3, 0); }

// Every rank exports one arena; neighbours map it with CUDA IPC.  The decision to use the
// peer-to-peer path is collective (all ranks or none): anything that fails falls back to NCCL.
static int p2p_setup() {
  const astr_cfg& c = g->cfg;
  int ok = !c.xchg_nccl;
  const int dims[3] = {c.im, c.jm, c.km};
  struct Rec { cudaIpcMemHandle_t h; int dims[3]; int ok; char pad[128 - sizeof(cudaIpcMemHandle_t) - 16]; };
  static_assert(sizeof(Rec) == 128, "Rec size");
  Rec mine;
  memset(&mine, 0, sizeof mine);
  if (ok) {
    const size_t bytes = xarena_bytes(dims);
    if (cudaMalloc(&g->xarena, bytes) != cudaSuccess) { g->xarena = nullptr; ok = 0; cudaGetLastError(); }
    else {
      // local failures only clear `ok`: every rank must still reach the two collectives below
      if (cudaMemsetAsync(g->xarena, 0, XHEADER_BYTES, g->st) != cudaSuccess) { ok = 0; cudaGetLastError(); }
      if (ok && cudaIpcGetMemHandle(&mine.h, g->xarena) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    }
  }
  for (int k = 0; k < 3; ++k) mine.dims[k] = dims[k];
  mine.ok = ok;
  std::vector<Rec> all(g->nranks);
  char* dbuf = nullptr;
  CUDA_OK(cudaMalloc(&dbuf, sizeof(Rec) * g->nranks));     // 128 bytes per rank: a failure here is fatal for the context anyway
  CUDA_OK(cudaMemcpyAsync(dbuf + sizeof(Rec) * g->rank, &mine, sizeof(Rec), cudaMemcpyHostToDevice, g->st));
  NCCL_OK(g_nccl.AllGather(dbuf + sizeof(Rec) * g->rank, dbuf, sizeof(Rec), ncclChar, g->comm, g->st));
  CUDA_OK(cudaMemcpyAsync(all.data(), dbuf, sizeof(Rec) * g->nranks, cudaMemcpyDeviceToHost, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  for (const Rec& r : all) ok = ok && r.ok;
  // map the neighbours' arenas (a peer that is both the low and the high neighbour is opened once)
  std::vector<std::pair<int, char*>> opened;
  for (int s = 0; ok && s < 6; ++s) {
    const int r = c.nbr[s];
    if (r < 0 || c.size[s / 2] == 1) continue;
    if (r == g->rank || r >= g->nranks) { ok = 0; break; }
    char* base = nullptr;
    for (auto& o : opened) if (o.first == r) base = o.second;
    if (!base) {
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
      base = (char*)p;
      opened.push_back({r, base});
      g->ipc_opened.push_back(p);
    }
    g->peer_arena[s] = base;
    for (int k = 0; k < 3; ++k) g->peer_dims[s][k] = all[r].dims[k];
  }
  // collective agreement
  int* dflag = (int*)dbuf;
  CUDA_OK(cudaMemcpyAsync(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice, g->st));
  NCCL_OK(g_nccl.AllReduce(dflag, dflag, 1, ncclInt, ncclMin, g->comm, g->st));
  CUDA_OK(cudaMemcpyAsync(&ok, dflag, sizeof(int), cudaMemcpyDeviceToHost, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  cudaFree(dbuf);
  g->p2p = ok != 0;
  if (!g->p2p && g->rank == 0 && !c.xchg_nccl)
    fprintf(stderr, "astr_gpu: peer-to-peer halo exchange unavailable (CUDA IPC), using ncclSend/ncclRecv\n");
  return 0;
}

// one direction through peer memory: SEND (fused pack + remote stores + ready flag), RECV (wait, unpack, ack)
// phases: 1 = SEND, 2 = RECV (of the exchange the last SEND of this direction started), 3 = both
static int exchange_dir_p2p(const FieldList& fl, int d, int l0, int l1, int phases) {
  const astr_cfg& c = g->cfg;
  const Layout& L = g->L;
  const int dims[3] = {c.im, c.jm, c.km};
  XHeader* hdr = reinterpret_cast<XHeader*>(g->xarena);
  if (phases & 1) ++g->xseq[d];
  const unsigned long long seq = g->xseq[d];
  XArgs sa, ra;
  memset(&sa, 0, sizeof sa); memset(&ra, 0, sizeof ra);
  sa.l0 = ra.l0 = l0; sa.l1 = ra.l1 = l1;
  sa.err = ra.err = &hdr->err;
  sa.timeout_cycles = ra.timeout_cycles = g->xtimeout_cycles;
  for (int side = 0; side < 2; ++side) {
    const int slot = 2 * d + side;
    if (c.nbr[slot] < 0) continue;
    char* peer = g->peer_arena[slot];
    XHeader* ph = reinterpret_cast<XHeader*>(peer);
    const int oslot = 2 * d + (1 - side);      // the face of the neighbour that touches mine
    XSide& s = sa.s[side];
    s.active = 1;
    s.remote = reinterpret_cast<double*>(peer + xwin_offset(g->peer_dims[slot], d, 1 - side));
    s.wait_flag = &hdr->ack[slot]; s.wait_val = seq - 1;      // my previous data has been consumed
    s.signal_flag = &ph->ready[oslot]; s.signal_val = seq;
    s.counter = &hdr->csend[slot];
    XSide& r = ra.s[side];
    r.active = 1;
    r.local = reinterpret_cast<const double*>(g->xarena + xwin_offset(dims, d, side));
    r.wait_flag = &hdr->ready[slot]; r.wait_val = seq;
    r.signal_flag = &ph->ack[oslot]; r.signal_val = seq;
    r.counter = &hdr->crecv[slot];
  }
  if (phases & 1) { ProfScope ps(PC_XPACK); TRY(pw_xsend(L, fl, d, sa, g->cur)); }
  if (phases & 2) { ProfScope ps(PC_XUNPACK); TRY(pw_xrecv(L, fl, d, ra, g->cur)); }
  return 0;
}

// Called by every API entry point that synchronises the stream and hands data back to the host: a timed-out
// exchange poisons the state (k_xface neither moves data nor signals), and the error stays until finalize.
static int p2p_check() {
  if (!g->p2p) return 0;
  if (!g->xfailed) {
    unsigned int e = 0;
    CUDA_OK(cudaMemcpyAsync(&e, g->xarena + offsetof(XHeader, err), sizeof e, cudaMemcpyDeviceToHost, g->st));
    CUDA_OK(cudaStreamSynchronize(g->st));
    g->xfailed = (e != 0);
  }
  if (g->xfailed)
    return astr_fail_msg("halo exchange timed out waiting for a neighbour (peer-memory flags): the device state is "
                         "invalid from that exchange on; finalize and restart from the last checkpoint");
  return 0;
}

static int dataswap(const FieldList& fl, int direction /*-1 all*/) {
  // all directions through peer memory: the three exchanges are independent (planes 1..hm of the interior go out,
  // halo planes come in, no corner exchange: src/parallel.F90:4172-4173), so every SEND is started before the first
  // RECV waits -- by then the neighbours' data has long arrived
  if (direction < 0 && g->p2p && g->comm && fl.nf * ASTR_HM <= XMAX_PLANES) {
    for (int d = 0; d < 3; ++d) {
      if (g->cfg.size[d] == 1) TRY(exchange_dir(fl, d, XMODE_SWAP));
      else TRY(exchange_dir_p2p(fl, d, 1, ASTR_HM, 1));
    }
    for (int d = 0; d < 3; ++d)
      if (g->cfg.size[d] > 1) TRY(exchange_dir_p2p(fl, d, 1, ASTR_HM, 2));
    return 0;
  }
  for (int d = 0; d < 3; ++d)
    if (direction < 0 || direction == d) TRY(exchange_dir(fl, d, XMODE_SWAP));
  return 0;
}

// -------------------------------------------------------------------------------------
// sweeps
// -------------------------------------------------------------------------------------
static int sweep(int d, int optype, const double* const* in, double* const* out, int nf, int epi, int o_lo,
                 int o_hi) {
  SweepArgs a;
  memset(&a, 0, sizeof a);
  a.L = g->L;
  const HostOp& hop = optype == OP_DERIV ? g->fd[d] : optype == OP_FILTER ? g->fl[d] : optype == OP_FLUXP ? g->fxp[d] : g->fxm[d];
  a.op = hop.dev();
  a.nf = nf;
  for (int i = 0; i < nf; ++i) { a.in[i] = in[i]; a.out[i] = out[i]; }
  a.epi = epi; a.o_lo = o_lo; a.o_hi = o_hi;
  if (optype == OP_DERIV && !g->cfg.scheme_compact) return pw_diff6e(d, a, g->st);   // difschm '...e'
  if (!g->cfg.legacy_sweep) {
    const int rc = astr_launch_sweep2(d, optype, hop.plan, a, g->st);
    if (rc >= 0) return rc;
  }
  return astr_launch_sweep(d, optype, a, g->st);
}

int astr_sweep_slots(int d, int optype, const int* in_slots, const int* out_slots, int nf, int epi, int o_lo,
                     int o_hi) {
  const double* in[ASTR_MAXF]; double* out[ASTR_MAXF];
  for (int i = 0; i < nf; ++i) { in[i] = g->slot(in_slots[i]); out[i] = g->slot(out_slots[i]); }
  return sweep(d, optype, in, out, nf, epi, o_lo, o_hi);
}

static int dim_of(int d) { return d == 0 ? g->cfg.im : (d == 1 ? g->cfg.jm : g->cfg.km); }

// -------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------
extern "C" {

const char* astr_gpu_last_error(void) { return g_err.c_str(); }
int astr_gpu_sizeof_cfg(void) { return (int)sizeof(astr_cfg); }

static int init_impl(const astr_cfg* cfg) {
  if (!cfg) return astr_fail_msg("null cfg");
  if (cfg->abi_version != ASTR_GPU_ABI_VERSION) return astr_fail_msg("abi_version mismatch");
  if (cfg->hm != ASTR_HM || cfg->numq != ASTR_GPU_NUMQ) return astr_fail_msg("hm must be 5 and numq 5");
  // ndims=3, or ndims=2 with km=0 (the k planes -hm..hm then replicate plane 0, src/parallel.F90:4296-4299)
  if (!((cfg->ndims == 3 && cfg->km >= 1) || (cfg->ndims == 2 && cfg->km == 0 && cfg->size[2] == 1)))
    return astr_fail_msg("ndims must be 3 (km>=1) or 2 (km=0, ksize=1)");
  for (int n = 0; n < 6; ++n)
    if (!(cfg->bctype[n] == 1 || cfg->bctype[n] == 41 || (cfg->bctype[n] == 11 && n == 0) ||
          (cfg->bctype[n] == 21 && (n == 1 || n == 3)) || (cfg->bctype[n] == 51 && n >= 2) ||
          (cfg->bctype[n] == 421 && (n == 2 || n == 3))))
      return astr_fail_msg("bctype on the device: 1 (periodic), 41 (isothermal wall, any face), 11 (inflow, imin), "
                           "21 (outflow, imax or jmax), 51 (farfield, jmin / jmax / kmin / kmax), 421 (slip adiabatic "
                           "wall, jmin / jmax) -- the faces the reference's own routines implement");

  // 643c: compact_central; 642e: explicit_central (diff6ec ignores the scheme digits, derivative.F90:319)
  if (!(cfg->scheme_compact ? cfg->difschm == 643 : (cfg->difschm / 100) == 6))
    return astr_fail_msg("difschm must be 643c (compact_central) or 6xxe (explicit_central)");
  // conschm: central (= difschm, comsolver.F90:94-101) or 543c, the upwind compact scheme (convrsdcmp)
  if (cfg->conschm_explicit) {
    // conschm '<odd>..e': convrsduwd with recons_exp; the reference stops for npdcj/npdck == 4 (solver.F90:810-821)
    if ((cfg->conschm / 100) % 2 != 1) return astr_fail_msg("conschm_explicit needs an odd leading digit (upwind-biased)");
    const int rs = cfg->recon_schem;
    if (!(rs == -1 || rs == 0 || rs == 1 || rs == 2 || rs == 3 || rs == 5 || rs == 6))
      return astr_fail_msg("recon_schem must be -1, 0, 1, 2, 3, 5 or 6 (src/flux.F90:269-350)");
    if (cfg->npdc[1] == 4 || (cfg->ndims == 3 && cfg->npdc[2] == 4))
      return astr_fail_msg("convrsduwd stops for npdcj/npdck == 4 (src/solver.F90:810-821, :1002-1013)");
  } else if (cfg->conschm != cfg->difschm && cfg->conschm != 543)
    return astr_fail_msg("conschm must equal difschm (central), be 543c (upwind compact) or an explicit upwind scheme");
  if (cfg->rkscheme != 3 && cfg->rkscheme != 4) return astr_fail_msg("rkscheme must be 3 (rk3) or 4 (rk4)");
  if (cfg->reserved0 != 0) return astr_fail_msg("cfg.reserved0 must be 0");
  for (int d = 0; d < 3; ++d)
    if (cfg->npdc[d] < 1 || cfg->npdc[d] > 4) return astr_fail_msg("npdc must be 1..4");
  if (g) astr_gpu_finalize();
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  if (ndev < 1) return astr_fail_msg("no CUDA device: this library has no CPU path");
  if (cfg->device >= 0) CUDA_OK(cudaSetDevice(cfg->device));
  g = new Ctx();
  g->cfg = *cfg;
  Layout& L = g->L;
  L.im = cfg->im; L.jm = cfg->jm; L.km = cfg->km;
  L.pitch = ((ASTR_IOFF + cfg->im + 1 + ASTR_HM + 1) + 31) / 32 * 32;
  L.njt = cfg->jm + 1 + 2 * ASTR_HM; L.nkt = cfg->km + 1 + 2 * ASTR_HM;
  L.sj = L.pitch; L.sk = (long long)L.pitch * L.njt;
  L.org = ASTR_IOFF + L.sj * ASTR_HM + L.sk * ASTR_HM;
  L.fstride = (L.sk * L.nkt + 31) / 32 * 32;
  g->th.reynolds = cfg->reynolds; g->th.prandtl = cfg->prandtl; g->th.const1 = cfg->const1;
  g->th.const2 = cfg->const2; g->th.const5 = cfg->const5; g->th.const6 = cfg->const6;
  g->th.tempconst = cfg->tempconst; g->th.tempconst1 = cfg->tempconst1;
  g->th.gamma = cfg->gamma; g->th.mach = cfg->mach;
  g->th.nondimen = cfg->nondimen ? 1 : 0;                     // src/solver.F90:124-128
  g->th.rgas = 287.1;
  g->th.cp = cfg->gamma / (cfg->gamma - 1.0) * g->th.rgas;
  g->th.cv = g->th.rgas / (cfg->gamma - 1.0);
  CUDA_OK(cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&g->xst, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&g->ev_shell, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&g->ev_xdone, cudaEventDisableTiming));
  g->cur = g->st;
  {
    // flag-wait timeout of the peer-memory exchange in SM clock cycles (clock64 runs at the SM clock)
    int dev = 0, khz = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    // default 60 s: ranks of one job skew by seconds around host-side phases (grid generation, uploads, I/O on rank
    // 0); 4 s (round 1) was exceeded by eight ranks finishing their 20 GB uploads at different times
    const long long ms = cfg->xchg_timeout_ms == 0 ? 60000 : cfg->xchg_timeout_ms;
    g->xtimeout_cycles = ms < 0 ? 0 : ms * (long long)khz;
  }
  const size_t bytes = (size_t)S_CORE * L.fstride * sizeof(double);
  CUDA_OK(cudaMalloc(&g->pool, bytes));
  CUDA_OK(cudaMemsetAsync(g->pool, 0, bytes, g->st));
  TRY(astr_sweep2_register_pool(0, g->pool, S_CORE, g->L));
  TRY(astr_sweep2_register_pool(1, nullptr, 0, g->L));
  TRY(astr_sweep2_register_pool(2, nullptr, 0, g->L));
  CUDA_OK(cudaMalloc(&g->d_partial, (size_t)4 * (cfg->jm + 1) * (cfg->km + 1) * sizeof(double)));
  CUDA_OK(cudaMalloc(&g->d_out2, 4 * sizeof(double)));
  CUDA_OK(cudaMalloc(&g->d_src, 8 * sizeof(double)));
  build_filter_coef(g->fc, cfg->alfa_filter, 1.11, 0.98, cfg->bfacmpld);
  TRY(astr_set_filter_coef(g->fc));
  for (int d = 0; d < cfg->ndims; ++d) {
    const int n = dim_of(d);
    build_deriv(g->fd[d], cfg->npdc[d], n, d);
    build_filter(g->fl[d], cfg->npdc[d], n, cfg->alfa_filter, d);
    TRY(astr_sweep2_set_plan(d, OP_DERIV, g->fd[d].plan, g->fc));
    TRY(astr_sweep2_set_plan(d, OP_FILTER, g->fl[d].plan, g->fc));
    if (g->upwind() && !cfg->conschm_explicit) {
      build_flux(g->fxp[d], OP_FLUXP, cfg->npdc[d], n, cfg->bfacmpld, d);
      build_flux(g->fxm[d], OP_FLUXM, cfg->npdc[d], n, cfg->bfacmpld, d);
      TRY(astr_sweep2_set_plan(d, OP_FLUXP, g->fxp[d].plan, g->fc));
      TRY(astr_sweep2_set_plan(d, OP_FLUXM, g->fxm[d].plan, g->fc));
    }
    for (HostOp* h : {&g->fd[d], &g->fl[d], &g->fxp[d], &g->fxm[d]}) {
      if (h->nrows == 0) continue;   // flux operators exist only for conschm 543
      if (h->C < 1) return astr_fail_msg("block too small: every direction needs at least 12 nodes");
      const size_t nb = (size_t)5 * h->nrows * sizeof(double);
      CUDA_OK(cudaMalloc(&h->d_tab, nb));
      std::vector<double> t;
      t.reserve(5 * h->nrows);
      for (auto* v : {&h->ac1, &h->ac2, &h->ac3, &h->pf, &h->qb}) t.insert(t.end(), v->begin(), v->end());
      CUDA_OK(cudaMemcpy(h->d_tab, t.data(), nb, cudaMemcpyHostToDevice));
    }
  }
  if (g->upwind()) {
    TRY(astr_set_flux_coef(cfg->bfacmpld));
    const size_t ub = (size_t)UP_TOTAL * L.fstride * sizeof(double);
    CUDA_OK(cudaMalloc(&g->up, ub));
    CUDA_OK(cudaMemsetAsync(g->up, 0, ub, g->st));
    TRY(astr_sweep2_register_pool(2, g->up, UP_TOTAL, g->L));
  }
  g_launches = 0;
  CUDA_OK(cudaStreamSynchronize(g->st));
  return 0;
}

// A failed init leaves no context behind: later calls report "astr_gpu_init not called" instead of
// dereferencing a half-built one.
int astr_gpu_init(const astr_cfg* cfg) {
  const int rc = init_impl(cfg);
  if (rc && g) {
    const std::string keep = g_err;
    astr_gpu_finalize();
    g_err = keep;
  }
  return rc;
}

// Safe on a half-built context (astr_gpu_init calls it on every error path): every member is checked.
int astr_gpu_finalize(void) {
  if (!g) return 0;
  if (g->st) { cudaStreamSynchronize(g->st); prof_collect(); }
  for (auto e : g->free_events) cudaEventDestroy(e);
  release_comm();
  for (int d = 0; d < 3; ++d)
    for (HostOp* h : {&g->fd[d], &g->fl[d], &g->fxp[d], &g->fxm[d]})
      if (h->d_tab) cudaFree(h->d_tab);
  if (g->up) cudaFree(g->up);
  for (int k = 0; k < 3; ++k) astr_sweep2_register_pool(k, nullptr, 0, g->L);
  if (g->pool) cudaFree(g->pool);
  if (g->scr) cudaFree(g->scr);
  if (g->rhsav) cudaFree(g->rhsav);
  if (g->crinod) cudaFree(g->crinod);
  for (auto& d : g->dense) if (d) cudaFree(d);
  for (auto& e : g->ev_dense) if (e) cudaEventDestroy(e);
  for (auto& b : g->bak) if (b) cudaFree(b);
  if (g->d_count) cudaFree(g->d_count);
  if (g->d_list) cudaFree(g->d_list);
  if (g->spg_global_coef) cudaFree(g->spg_global_coef);
  if (g->d_partial) cudaFree(g->d_partial);
  if (g->d_out2) cudaFree(g->d_out2);
  if (g->d_src) cudaFree(g->d_src);
  if (g->ycoord) cudaFree(g->ycoord);
  if (g->d_inflow) cudaFree(g->d_inflow);
  if (g->stage) cudaFree(g->stage);
  for (auto& sp : g->spg) if (sp.coef) cudaFree(sp.coef);
  if (g->st) cudaStreamDestroy(g->st);
  if (g->xst) cudaStreamDestroy(g->xst);
  if (g->ev_shell) cudaEventDestroy(g->ev_shell);
  if (g->ev_xdone) cudaEventDestroy(g->ev_xdone);
  delete g;
  g = nullptr;
  cudaGetLastError();
  return 0;
}

int astr_gpu_synchronize(void) {
  NEED_CTX();
  CUDA_OK(cudaStreamSynchronize(g->st));
  prof_collect();
  return p2p_check();
}

int astr_gpu_comm_unique_id(char id[128]) {
  if (!g_nccl.load()) return astr_fail_msg("cannot load libnccl.so.2");
  ncclUniqueId uid;
  NCCL_OK(g_nccl.GetUniqueId(&uid));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(id, &uid, 128);
  return 0;
}

int astr_gpu_comm_init(const char id[128], int nranks, int rank) {
  NEED_CTX();
  if (!g_nccl.load()) return astr_fail_msg("cannot load libnccl.so.2");
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  release_comm();                                    // a second comm_init replaces the first communicator
  NCCL_OK(g_nccl.CommInitRank(&g->comm, nranks, uid, rank));
  g->nranks = nranks; g->rank = rank;
  return p2p_setup();
}

// ---- host <-> device field transfer (Fortran halo'd box <-> padded device box) --------
static int copy_dev(double* field, double* host, bool to_host) {
  const Layout& L = g->L;
  const size_t w = (size_t)(L.im + 1 + 2 * ASTR_HM) * sizeof(double);
  const size_t rows = (size_t)L.njt * L.nkt;
  if (!to_host && rows * w >= ((size_t)8 << 20)) {
    // uploads: one contiguous DMA into a dense staging field + a re-pitch kernel (measured on the bench box:
    // 54 GB/s contiguous against 30 GB/s for the pitched 2-D host-to-device copy); stream order protects the buffer
    if (!g->stage) CUDA_OK(cudaMalloc(&g->stage, rows * w));
    CUDA_OK(cudaMemcpyAsync(g->stage, host, rows * w, cudaMemcpyHostToDevice, g->st));
    return pw_repitch(L, g->stage, field, g->st);
  }
  double* dev = field + (ASTR_IOFF - ASTR_HM);
  if (to_host)
    CUDA_OK(cudaMemcpy2DAsync(host, w, dev, (size_t)L.pitch * sizeof(double), w, rows, cudaMemcpyDeviceToHost, g->st));
  else
    CUDA_OK(cudaMemcpy2DAsync(dev, (size_t)L.pitch * sizeof(double), host, w, w, rows, cudaMemcpyHostToDevice, g->st));
  return 0;
}
static int copy_field(int slot, double* host, bool to_host) { return copy_dev(g->slot(slot), host, to_host); }
static size_t host_field_elems() {
  const Layout& L = g->L;
  return (size_t)(L.im + 1 + 2 * ASTR_HM) * L.njt * L.nkt;
}

static int api_slot(int field_id, int* slot) {
  if (field_id < 0 || field_id >= ASTR_F_SSF) return astr_fail_msg("bad field id");
  if (field_id < ASTR_F_QRHS) *slot = S_Q + field_id;                        // q, rho, vel, prs, tmp
  else if (field_id < ASTR_F_JACOB) *slot = S_QRHS + (field_id - ASTR_F_QRHS);
  else if (field_id == ASTR_F_JACOB) *slot = S_JAC;
  else if (field_id < ASTR_F_DVEL) *slot = S_DXI + (field_id - ASTR_F_DXI);
  else if (field_id < ASTR_F_SIGMA) *slot = S_SCR + (field_id - ASTR_F_DVEL);  // dvel 9, dtmp 3
  else if (field_id < ASTR_F_QFLUX) *slot = S_SIGMA + (field_id - ASTR_F_SIGMA);
  else if (field_id < ASTR_F_X) *slot = S_QFLUX + (field_id - ASTR_F_QFLUX);
  else if (field_id < ASTR_F_QSAVE) *slot = S_G + (field_id - ASTR_F_X);      // x lives in G slots during gridgeom
  else if (field_id < ASTR_F_VOR) *slot = S_QSAVE + (field_id - ASTR_F_QSAVE);
  else *slot = S_SCR + 12 + (field_id - ASTR_F_VOR);
  return 0;
}

static int ensure_crinod();

int astr_gpu_get_field(int field_id, double* host) {
  NEED_CTX();
  if (field_id == ASTR_F_CRINOD) {
    TRY(ensure_crinod());
    TRY(copy_dev(g->crinod, host, true));
    CUDA_OK(cudaStreamSynchronize(g->st));
    return 0;
  }
  if (field_id == ASTR_F_SSF || field_id == ASTR_F_LSHOCK) {
    if (!g->up) return astr_fail_msg("ssf/lshock exist only on the upwind convection paths");
    TRY(copy_dev(g->up + (size_t)(field_id == ASTR_F_SSF ? UP_SSF : UP_LSH) * g->L.fstride, host, true));
    CUDA_OK(cudaStreamSynchronize(g->st));
    return 0;
  }
  int s;
  TRY(api_slot(field_id, &s));
  if (s >= S_SCR) {
    TRY(ensure_scratch());
    if (!g->have_grad) return astr_fail_msg("dvel/dtmp/vor requested before gradcal");
    TRY(pw_materialise_grad(g->L, g->pool, g->scr, g->st));
  }
  if (s >= S_QRHS && s < S_QRHS + 5 && g->rhs_in_g) TRY(pw_sum_qrhs(g->L, g->pool, g->src(), g->st));
  if (s >= S_SIGMA && s < S_SIGMA + 9 && g->sigma_partial) {
    // the fused rhscal keeps the interior of sigma/qflux in registers: materialise on request
    const Box all = {{0, 0, 0}, {g->L.im, g->L.jm, g->L.km}};
    TRY(pw_visc(g->L, g->pool, g->th, all, g->st));
    g->sigma_partial = false;
  }
  TRY(copy_field(s, host, true));
  CUDA_OK(cudaStreamSynchronize(g->st));
  return 0;
}

int astr_gpu_set_field(int field_id, const double* host) {
  NEED_CTX();
  if (field_id == ASTR_F_CRINOD) {
    TRY(ensure_crinod());
    TRY(copy_dev(g->crinod, const_cast<double*>(host), false));
    CUDA_OK(cudaStreamSynchronize(g->st));
    return 0;
  }
  if (field_id >= ASTR_F_SSF) return astr_fail_msg("ssf/lshock are derived fields");
  int s;
  TRY(api_slot(field_id, &s));
  if (s >= S_SCR) return astr_fail_msg("dvel/dtmp/vor are derived fields");
  if (s >= S_QRHS && s < S_QRHS + 5) {
    if (g->rhs_in_g) TRY(pw_sum_qrhs(g->L, g->pool, g->src(), g->st));
    g->rhs_in_g = false;   // the caller now owns qrhs as an array
    g->src_pending = false;
  }
  TRY(copy_field(s, const_cast<double*>(host), false));
  CUDA_OK(cudaStreamSynchronize(g->st));
  // metrics given field by field (jacob, dxi): the caller vouches for the whole set, as with set_metrics
  if (s == S_JAC || (s >= S_DXI && s < S_DXI + 9)) g->have_metrics = true;
  return 0;
}

int astr_gpu_device_ptr(int field_id, void** dptr, long long strides[3], long long* origin) {
  NEED_CTX();
  int s;
  TRY(api_slot(field_id, &s));
  if (s >= S_SCR) TRY(ensure_scratch());
  *dptr = g->slot(s);
  strides[0] = 1; strides[1] = g->L.sj; strides[2] = g->L.sk;
  *origin = g->L.org;
  return 0;
}

int astr_gpu_set_metrics(const double* dxi, const double* jacob) {
  NEED_CTX();
  const size_t ne = host_field_elems();
  // dxi(-hm:im+hm,-hm:jm+hm,-hm:km+hm,1:3,1:3): 4th index = xi direction a, 5th = x index b
  for (int b = 0; b < 3; ++b)
    for (int a = 0; a < 3; ++a)
      TRY(copy_field(S_DXI + 3 * a + b, const_cast<double*>(dxi) + (size_t)(a + 3 * b) * ne, false));
  TRY(copy_field(S_JAC, const_cast<double*>(jacob), false));
  CUDA_OK(cudaStreamSynchronize(g->st));
  g->have_metrics = true;
  return 0;
}

// src_chan integrates in y (src/solver.F90:321): the channel case keeps x(:,:,:,2) resident
static int keep_ycoord(const double* dev_field) {
  const size_t bytes = (size_t)g->L.fstride * sizeof(double);
  if (!g->ycoord) CUDA_OK(cudaMalloc(&g->ycoord, bytes));
  CUDA_OK(cudaMemcpyAsync(g->ycoord, dev_field, bytes, cudaMemcpyDeviceToDevice, g->st));
  return 0;
}

int astr_gpu_set_grid(const double* x) {
  NEED_CTX();
  if (g->cfg.flowtype != 1) return 0;   // only src_chan reads coordinates
  const size_t bytes = (size_t)g->L.fstride * sizeof(double);
  if (!g->ycoord) { CUDA_OK(cudaMalloc(&g->ycoord, bytes)); CUDA_OK(cudaMemsetAsync(g->ycoord, 0, bytes, g->st)); }
  const Layout& L = g->L;
  const size_t w = (size_t)(L.im + 1 + 2 * ASTR_HM) * sizeof(double);
  CUDA_OK(cudaMemcpy2DAsync(g->ycoord + (ASTR_IOFF - ASTR_HM), (size_t)L.pitch * sizeof(double),
                            x + host_field_elems(), w, w, (size_t)L.njt * L.nkt, cudaMemcpyHostToDevice, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  return 0;
}

int astr_gpu_gridgeom(const double* x) {
  NEED_CTX();
  TRY(ensure_scratch());
  const size_t ne = host_field_elems();
  for (int m = 0; m < 3; ++m) TRY(copy_field(S_G + m, const_cast<double*>(x) + (size_t)m * ne, false));
  if (g->cfg.flowtype == 1) TRY(keep_ycoord(g->slot(S_G + 1)));
  TRY(geom_gridgeom(g->L, g->cfg, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  g->have_metrics = true;
  return 0;
}

int astr_gpu_upload_state(const double* q, const double* rho, const double* vel, const double* prs,
                          const double* tmp) {
  NEED_CTX();
  const size_t ne = host_field_elems();
  if (q) for (int m = 0; m < 5; ++m) TRY(copy_field(S_Q + m, const_cast<double*>(q) + m * ne, false));
  if (rho) TRY(copy_field(S_RHO, const_cast<double*>(rho), false));
  if (vel) for (int m = 0; m < 3; ++m) TRY(copy_field(S_VEL + m, const_cast<double*>(vel) + m * ne, false));
  if (prs) TRY(copy_field(S_PRS, const_cast<double*>(prs), false));
  if (tmp) TRY(copy_field(S_TMP, const_cast<double*>(tmp), false));
  CUDA_OK(cudaStreamSynchronize(g->st));
  return 0;
}

int astr_gpu_download_state(double* q, double* rho, double* vel, double* prs, double* tmp) {
  NEED_CTX();
  const size_t ne = host_field_elems();
  if (q) for (int m = 0; m < 5; ++m) TRY(copy_field(S_Q + m, q + m * ne, true));
  if (rho) TRY(copy_field(S_RHO, rho, true));
  if (vel) for (int m = 0; m < 3; ++m) TRY(copy_field(S_VEL + m, vel + m * ne, true));
  if (prs) TRY(copy_field(S_PRS, prs, true));
  if (tmp) TRY(copy_field(S_TMP, tmp, true));
  CUDA_OK(cudaStreamSynchronize(g->st));
  prof_collect();
  return p2p_check();
}

// ---- stage operators -------------------------------------------------------------------
int astr_gpu_filterq(void) {
  NEED_CTX();
  const astr_cfg& c = g->cfg;
  const FieldList fq = fields(S_Q, 5);
  for (int d = 0; d < c.ndims; ++d) {              // comsolver.F90:588 `if(ndims==3)` for k
    TRY(dataswap(fq, d));                          // comsolver.F90:535,562,590
    const int n = dim_of(d);
    const int nt = c.npdc[d];
    // physical-boundary nodes stay unfiltered (filter.F90:141-142): in place => not written
    const int o_lo = (nt == 1 || nt == 4) ? 1 : 0, o_hi = (nt == 2 || nt == 4) ? n - 1 : n;
    ProfScope ps(PC_FILTER_I + d);
    TRY(sweep(d, OP_FILTER, fq.f, fq.f, 5, EPI_STORE, o_lo, o_hi));
  }
  return 0;
}

// boucon (src/bc.F90:327-407): faces n=1..6 in order; a face is treated only by the ranks that
// own it (irk==0 / irk==irkm etc.)
int astr_gpu_boucon(void) {
  NEED_CTX();
  const astr_cfg& c = g->cfg;
  for (int n = 0; n < 6; ++n) {
    const int bt = c.bctype[n], d = n / 2, side = n % 2;
    if (bt == 1) continue;
    const bool mine = side ? (c.rank[d] == c.size[d] - 1) : (c.rank[d] == 0);
    if (bt == 41) {                                   // noslip(n,twall(n)), bc.F90:6306
      if (!mine) continue;
      ProfScope ps(PC_HALO);
      TRY(pw_noslip(g->L, g->pool, g->th, d, side, c.twall[n], g->st));
    } else if ((bt == 11 && n == 0) || (bt == 21 && (n == 1 || n == 3)) || (bt == 51 && n >= 2) ||
               (bt == 421 && (n == 2 || n == 3))) {
      // inflow(1) bc.F90:1366, outflow(2|4) :3404, farfield(3..6) :3008, slipadibwall(3|4) :7231
      if (!mine) continue;
      BcArgs a;
      a.kind = bt; a.side = side; a.pinf = c.pinf; a.deltat = c.deltat;
      a.vinf[0] = c.uinf; a.vinf[1] = c.vinf; a.vinf[2] = c.winf; a.roinf = c.roinf;
      a.vel_in = a.tmp_in = a.tmp_prof = nullptr;
      if (bt == 11) {
        if (!g->d_inflow) return astr_fail_msg("boucon: bctype 11 needs astr_gpu_set_inflow");
        const size_t nf = (size_t)(c.jm + 1) * (c.km + 1);
        a.vel_in = g->d_inflow; a.tmp_in = g->d_inflow + 3 * nf; a.tmp_prof = g->d_inflow + 4 * nf;
      }
      ProfScope ps(PC_HALO);
      TRY(pw_bcface(g->L, g->pool, g->th, d, a, g->st));
    } else {
      return astr_fail_msg("boucon: this bctype / face combination does not run on the device");
    }
  }
  return 0;
}

// src_chan (src/solver.F90:295-353): bulk integrals -> psum -> (force, force.ubulk) on the device;
// the term itself is added where qrhs is consumed (RK update / qrhs materialisation)
static int src_chan() {
  if (!g->ycoord) return astr_fail_msg("src_chan needs node coordinates: call astr_gpu_set_grid or astr_gpu_gridgeom");
  ProfScope ps(PC_FLUX);
  TRY(pw_bulk(g->L, g->pool, g->ycoord, g->d_partial, g->d_src, g->st));
  if (g->cfg.size[0] * g->cfg.size[1] * g->cfg.size[2] > 1) {
    if (!g->comm) return astr_fail_msg("src_chan psum needs astr_gpu_comm_init");
    NCCL_OK(g_nccl.AllReduce(g->d_src, g->d_src, 4, ncclDouble, ncclSum, g->comm, g->st));
  }
  TRY(pw_src_coef(g->d_src, g->force, g->d_src + 4, g->st));
  g->src_pending = true;
  return 0;
}

int astr_gpu_qswap(void) {
  NEED_CTX();
  const astr_cfg& c = g->cfg;
  const Layout& L = g->L;
  const FieldList fq = fields(S_Q, 5);
  for (int d = 0; d < 3; ++d) {
    TRY(exchange_dir(fq, d, XMODE_QSWAP));
    bool lo, hi;
    if (c.size[d] == 1) lo = hi = (c.lhomo[d] != 0);
    else { lo = c.nbr[2 * d] >= 0; hi = c.nbr[2 * d + 1] >= 0; }
    const int dm = dim_of(d);
    ProfScope ps(PC_HALO);
    // q2fvar on the slabs dm:dm+hm and -hm:0 (parallel.F90:4904-4915 etc.)
    for (int side = 1; side >= 0; --side) {
      if (!(side ? hi : lo)) continue;
      Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
      b.lo[d] = side ? dm : -ASTR_HM;
      b.hi[d] = side ? dm + ASTR_HM : 0;
      TRY(pw_q2fvar(L, g->pool, g->th, b, g->st));
    }
  }
  return 0;
}

int astr_gpu_gradcal(void) {
  NEED_CTX();
  if (!g->have_metrics) return astr_fail_msg("gradcal before set_metrics/gridgeom");
  const double* in[4] = {g->slot(S_VEL), g->slot(S_VEL + 1), g->slot(S_VEL + 2), g->slot(S_TMP)};
  for (int d = 0; d < g->cfg.ndims; ++d) {         // comsolver.F90:418 `if(ndims==3)` for k: raw k-derivatives stay 0
    double* out[4];
    for (int m = 0; m < 4; ++m) out[m] = g->slot(S_RAW + 4 * d + m);
    ProfScope ps(PC_GRAD_I + d);
    TRY(sweep(d, OP_DERIV, in, out, 4, EPI_STORE, 0, dim_of(d)));
  }
  g->have_grad = true;
  return 0;
}

// ducrossensor (src/commcal.F90:196-357): ssf on 0..im etc., dataswap(ssf), lshock
int astr_gpu_ducrossensor(void) {
  NEED_CTX();
  if (!g->up) return astr_fail_msg("ducrossensor needs conschm 543");
  if (!g->have_grad) return astr_fail_msg("ducrossensor before gradcal");
  {
    ProfScope ps(PC_FLUX);
    TRY(uw_ducros_ssf(g->L, g->pool, g->up, g->cfg.npdc, g->st));
  }
  FieldList fl; fl.nf = 1; fl.f[0] = g->up + (size_t)UP_SSF * g->L.fstride;
  TRY(dataswap(fl, -1));
  ProfScope ps(PC_FLUX);
  return uw_ducros_flag(g->L, g->up, g->cfg.npdc, g->cfg.shkcrt, g->st);
}

// convrsdcmp (src/solver.F90:1271-1937) of direction d; the result -(Fh(i)-Fh(i-1)) is added to the
// five slots starting at dst0 (rmw_mask: components that already hold the viscous derivative)
static int convrsdcmp_dir(int d, int dst0, int rmw_mask) {
  const astr_cfg& c = g->cfg;
  const Layout& L = g->L;
  const int dm = dim_of(d), nt = c.npdc[d];
  UpwindArgs ua;
  const int s[3] = {c.is, c.js, c.ks}, e[3] = {c.ie, c.je, c.ke};
  for (int k = 0; k < 3; ++k) { ua.s[k] = s[k]; ua.e[k] = e[k]; ua.box.lo[k] = s[k]; ua.box.hi[k] = e[k]; }
  ua.box.lo[d] = s[d] - 1;
  ua.lss = (nt == 1 || nt == 4) ? 0 : -ASTR_HM;           // solver.F90:1313-1327
  ua.lee = (nt == 2 || nt == 4) ? dm : dm + ASTR_HM;
  ua.dim = dm; ua.ntype = nt; ua.lchardecomp = c.lchardecomp;
  ua.explicit_recons = c.conschm_explicit; ua.recon_schem = c.recon_schem; ua.bfacmpld = c.bfacmpld;
  ua.crinod = g->crinod;
  // lshock exists (`allocated(lshock)`) exactly when ducrossensor runs (solver.F90:221,225)
  ua.sson = c.conschm_explicit ? (c.recon_schem == 5 || c.lchardecomp) : c.lchardecomp;
  ProfScope ps(PC_DIV_I + d);
  TRY(uw_sw_split(L, g->pool, g->up, g->th, d, ua.lss, ua.lee, g->st));
  if (c.conschm_explicit) {          // convrsduwd: no compact flux solves
    TRY(uw_interface_flux(L, g->pool, g->up, g->th, d, ua, g->st));
    return uw_fhdiff(L, g->pool, g->up, d, ua, dst0, rmw_mask, g->st);
  }
  const double* in[5]; double* out[5];
  for (int m = 0; m < 5; ++m) { in[m] = g->up + (size_t)(UP_FSW + m) * L.fstride; out[m] = g->up + (size_t)(UP_FHC + m) * L.fstride; }
  TRY(sweep(d, OP_FLUXP, in, out, 5, EPI_STORE, -1, dm));
  for (int m = 0; m < 5; ++m) { in[m] = g->up + (size_t)(UP_FSW + 5 + m) * L.fstride; out[m] = g->up + (size_t)(UP_FHC + 5 + m) * L.fstride; }
  TRY(sweep(d, OP_FLUXM, in, out, 5, EPI_STORE, -1, dm));
  TRY(uw_interface_flux(L, g->pool, g->up, g->th, d, ua, g->st));
  return uw_fhdiff(L, g->pool, g->up, d, ua, dst0, rmw_mask, g->st);
}

int astr_gpu_rhscal(void) {
  NEED_CTX();
  if (!g->have_grad) return astr_fail_msg("rhscal before gradcal");
  const astr_cfg& c = g->cfg;
  const Layout& L = g->L;
  FluxRanges fr = {{c.is, c.js, c.ks}, {c.ie, c.je, c.ke}};
  const bool upw = g->upwind();
  if (upw) {
    // the convective part comes from convrsdcmp: the G slots carry the viscous fluxes only
    for (int d = 0; d < 3; ++d) { fr.s[d] = 1; fr.e[d] = 0; }
    // solver.F90:221 (explicit: recon_schem==5 .or. lchardecomp), :225 (compact: lchardecomp)
    if (c.lchardecomp || (c.conschm_explicit && c.recon_schem == 5)) TRY(astr_gpu_ducrossensor());
  }
  if (c.diffterm) {
    const bool multi = c.size[0] * c.size[1] * c.size[2] > 1;
    if (multi && g->cfg.overlap_visc) {
      // multi-block: the sigma/qflux exchange (solver.F90:2604-2606) is hidden behind the interior pass.
      // 1. stresses on the face shells only (what the exchange sends), 2. exchange on the side stream,
      // 3. meanwhile the fused stress + flux pass over the whole block on the main stream, 4. join.
      {
        ProfScope ps(PC_VISC);
        for (int d = 0; d < c.ndims; ++d)
          for (int side = 0; side < 2; ++side) {
            Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
            const int dm = dim_of(d);
            b.lo[d] = side ? std::max(dm - ASTR_HM, 0) : 0;
            b.hi[d] = side ? dm : std::min(ASTR_HM, dm);
            TRY(pw_visc(L, g->pool, g->th, b, g->st));
          }
      }
      CUDA_OK(cudaEventRecord(g->ev_shell, g->st));
      CUDA_OK(cudaStreamWaitEvent(g->xst, g->ev_shell, 0));
      g->cur = g->xst;
      const int rc = dataswap(fields(S_SIGMA, 9), -1);
      g->cur = g->st;
      if (rc) return rc;
      CUDA_OK(cudaEventRecord(g->ev_xdone, g->xst));
      { ProfScope ps(PC_VISC); TRY(pw_visc_flux(L, g->pool, g->th, fr, c.ndims, 0, g->st)); }
      CUDA_OK(cudaStreamWaitEvent(g->st, g->ev_xdone, 0));
    } else {
      // viscous stress + flux assembly of the block in one pass (sigma/qflux stay in registers and
      // reach memory only on the face shells the exchange reads), then the halo exchange
      { ProfScope ps(PC_VISC); TRY(pw_visc_flux(L, g->pool, g->th, fr, c.ndims, 1, g->st)); }
      TRY(dataswap(fields(S_SIGMA, 9), -1));         // sigma(6)+qflux(3), solver.F90:2604-2606
    }
    g->sigma_partial = true;
  }
  {
    ProfScope ps(PC_FLUX);
    Box b = {{0, 0, 0}, {L.im, L.jm, L.km}};
    if (!c.diffterm && !upw) TRY(pw_flux(L, g->pool, b, c.ndims == 3 ? 7 : 3, fr, c.diffterm, g->st));
    for (int d = 0; d < c.ndims && (c.diffterm || !upw); ++d)   // halo slabs of direction d (fluxes on halo nodes,
      for (int side = 0; side < 2; ++side) {       // solver.F90:2200-2206)
        Box h = b;
        h.lo[d] = side ? dim_of(d) + 1 : -ASTR_HM;
        h.hi[d] = side ? dim_of(d) + ASTR_HM : -1;
        TRY(pw_flux(L, g->pool, h, 1 << d, fr, c.diffterm, g->st));
      }
  }
  const int s[3] = {c.is, c.js, c.ks}, e[3] = {c.ie, c.je, c.ke};
  if (!c.scheme_compact) {
    // explicit stencils cannot run in place: accumulate the three directions into qrhs
    for (int d = 0; d < c.ndims; ++d) {
      const double* in[5]; double* out[5];
      for (int m = 0; m < 5; ++m) { in[m] = g->slot(S_G + 5 * d + m); out[m] = g->slot(S_QRHS + m); }
      if (c.diffterm || !upw) {
        ProfScope ps(PC_DIV_I + d);
        TRY(sweep(d, OP_DERIV, in, out, 5, d == 0 ? EPI_STOREZ : EPI_ADD, s[d], e[d]));
      }
      if (upw) TRY(convrsdcmp_dir(d, S_QRHS, (c.diffterm || d > 0) ? 31 : 0));
    }
    g->rhs_in_g = false;
    g->src_pending = false;
    if (c.flowtype == 1) { TRY(src_chan()); TRY(pw_add_force(g->L, g->pool, g->src(), g->st)); g->src_pending = false; }
    return 0;
  }
  // d(G_d)/d(xi_d) in place (zero outside is:ie etc., where the reference does not accumulate);
  // the three directions are summed by the RK update kernel: no read-modify-write of qrhs
  // (2-D blocks: the G slots of the zeta direction are never written and stay zero)
  for (int d = 0; d < c.ndims; ++d) {
    const double* in[5]; double* out[5];
    for (int m = 0; m < 5; ++m) { in[m] = g->slot(S_G + 5 * d + m); out[m] = g->slot(S_G + 5 * d + m); }
    if (!upw) {
      ProfScope ps(PC_DIV_I + d);
      TRY(sweep(d, OP_DERIV, in, out, 5, EPI_STOREZ, s[d], e[d]));
    } else {
      if (c.diffterm) {   // the mass equation has no viscous flux: 4 fields
        ProfScope ps(PC_DIV_I + d);
        TRY(sweep(d, OP_DERIV, in + 1, out + 1, 4, EPI_STOREZ, s[d], e[d]));
      }
      TRY(convrsdcmp_dir(d, S_G + 5 * d, c.diffterm ? 30 : 0));
    }
  }
  g->rhs_in_g = true;
  g->src_pending = false;
  if (c.flowtype == 1) TRY(src_chan());             // solver.F90:262
  return 0;
}

static int rk_coef(int rkstep, double dt, RkCoef& rk) {
  const int nst = g->cfg.rkscheme;
  if (rkstep < 1 || rkstep > nst) return astr_fail_msg(nst == 4 ? "rkstep must be 1..4" : "rkstep must be 1..3");
  rk.scheme = nst; rk.last = (rkstep == nst); rk.rhsav = nullptr;
  if (nst == 4) {
    // src/mainloop.F90:368-376
    const double co4[2][4] = {{0.5, 0.5, 1.0, 1.0 / 6.0}, {1.0, 2.0, 2.0, 1.0}};
    rk.c1 = co4[0][rkstep - 1]; rk.c2 = co4[1][rkstep - 1]; rk.c3 = 0.0;
    if (!g->rhsav) CUDA_OK(cudaMalloc(&g->rhsav, (size_t)5 * g->L.fstride * sizeof(double)));   // mainloop.F90:394
    rk.rhsav = g->rhsav;
  } else {
    // src/mainloop.F90:350-362
    const double co[3][3] = {{1.0, 0.0, 1.0}, {0.75, 0.25, 0.25}, {1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0}};
    rk.c1 = co[rkstep - 1][0]; rk.c2 = co[rkstep - 1][1]; rk.c3 = co[rkstep - 1][2];
  }
  rk.dt = dt; rk.first = (rkstep == 1);
  rk.rhs_in_g = g->rhs_in_g ? 1 : 0;
  return 0;
}

int astr_gpu_rk_update(int rkstep, double deltat) {
  NEED_CTX();
  RkCoef rk;
  TRY(rk_coef(rkstep, deltat, rk));
  rk.with_fvar = 0;
  ProfScope ps(PC_RK);
  return pw_rk_update(g->L, g->pool, g->th, rk, g->src(), g->st);
}

int astr_gpu_updatefvar(void) {
  NEED_CTX();
  Box b = {{0, 0, 0}, {g->L.im, g->L.jm, g->L.km}};
  ProfScope ps(PC_FVAR);
  return pw_q2fvar(g->L, g->pool, g->th, b, g->st);
}

int astr_gpu_rk_stage(int rkstep, double deltat) {
  NEED_CTX();
  if (g->cfg.lfilter) TRY(astr_gpu_filterq());
  TRY(astr_gpu_boucon());
  TRY(astr_gpu_qswap());
  TRY(astr_gpu_gradcal());
  TRY(astr_gpu_rhscal());
  RkCoef rk;
  TRY(rk_coef(rkstep, deltat, rk));
  if (g->any_sponge()) {                             // mainloop.F90:478: spongefilter sits between the update and updatefvar
    rk.with_fvar = 0;
    { ProfScope ps(PC_RK); TRY(pw_rk_update(g->L, g->pool, g->th, rk, g->src(), g->st)); }
    TRY(astr_gpu_spongefilter());
    return astr_gpu_updatefvar();
  }
  rk.with_fvar = 1;
  ProfScope ps(PC_RK);
  return pw_rk_update(g->L, g->pool, g->th, rk, g->src(), g->st);
}

int astr_gpu_rk_steps(int nsteps, double deltat) {
  NEED_CTX();
  for (int s = 0; s < nsteps; ++s)
    for (int rk = 1; rk <= g->cfg.rkscheme; ++rk) TRY(astr_gpu_rk_stage(rk, deltat));
  return 0;
}

int astr_gpu_rk_steps_timed(int nsteps, double deltat, float* ms) {
  NEED_CTX();
  cudaEvent_t a, b;
  CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
  CUDA_OK(cudaEventRecord(a, g->st));
  int rc = astr_gpu_rk_steps(nsteps, deltat);
  CUDA_OK(cudaEventRecord(b, g->st));
  CUDA_OK(cudaEventSynchronize(b));
  CUDA_OK(cudaEventElapsedTime(ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  prof_collect();
  return rc ? rc : p2p_check();
}

int astr_gpu_dataswap(int field_id, int direction) {
  NEED_CTX();
  int s;
  TRY(api_slot(field_id, &s));
  if (s >= S_SCR) return astr_fail_msg("dataswap of a derived field");
  return dataswap(fields(s, 1), direction - 1);
}

// inflow data of alloinflow (src/bc.F90:69-83) as profileinflow / freestreaminflow / inflowintp leave it:
// vel_in(0:jm,0:km,1:3), tmp_in(0:jm,0:km), tmp_prof(0:jm).  Only ranks with irk==0 use it.
int astr_gpu_set_inflow(const double* vel_in, const double* tmp_in, const double* tmp_prof) {
  NEED_CTX();
  const astr_cfg& c = g->cfg;
  const size_t nf = (size_t)(c.jm + 1) * (c.km + 1), tot = 4 * nf + c.jm + 1;
  if (!g->d_inflow) CUDA_OK(cudaMalloc(&g->d_inflow, tot * sizeof(double)));
  CUDA_OK(cudaMemcpyAsync(g->d_inflow, vel_in, 3 * nf * sizeof(double), cudaMemcpyHostToDevice, g->st));
  CUDA_OK(cudaMemcpyAsync(g->d_inflow + 3 * nf, tmp_in, nf * sizeof(double), cudaMemcpyHostToDevice, g->st));
  CUDA_OK(cudaMemcpyAsync(g->d_inflow + 4 * nf, tmp_prof, (c.jm + 1) * sizeof(double), cudaMemcpyHostToDevice, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  return 0;
}

// sponge layer of one face (spongelayer_define_ijk / layer_setup, src/sponge_layer.F90:442-1011, stay on the host):
// face 0 i0, 1 im, 3 jm, 4 k0, 5 km; beg..end = node range of the layer along the face direction on this rank
// (beg<0: the layer exists, lspg_* is set, but not on this rank: the rank still takes part in the exchange);
// coef = sponge_damp_coef over [beg:end] x [s:e] x [s:e] of the other two directions, Fortran order
int astr_gpu_set_sponge(int face, int beg, int end, const double* coef) {
  NEED_CTX();
  if (face < 0 || face > 5 || face == 2) return astr_fail_msg("set_sponge: face must be 0 (i0), 1 (im), 3 (jm), 4 (k0) or 5 (km)");
  Ctx::Sponge& sp = g->spg[face];
  sp.on = true; sp.beg = beg; sp.end = end;
  if (sp.coef) { cudaFree(sp.coef); sp.coef = nullptr; }
  if (beg >= 0) {
    const astr_cfg& c = g->cfg;
    const int s[3] = {c.is, c.js, c.ks}, e[3] = {c.ie, c.je, c.ke};
    size_t cnt = (size_t)(end - beg + 1);
    for (int o = 0; o < 3; ++o) if (o != face / 2) cnt *= (size_t)(e[o] - s[o] + 1);
    CUDA_OK(cudaMalloc(&sp.coef, cnt * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(sp.coef, coef, cnt * sizeof(double), cudaMemcpyHostToDevice, g->st));
    CUDA_OK(cudaStreamSynchronize(g->st));
  }
  return 0;
}

// spg_def='circl' (spongelayer_define_circle stays on the host, src/sponge_layer.F90:369-440): coef =
// sponge_damp_coef(is:ie,js:je,ks:ke) of this rank, or NULL when the rank has no damped node (lsponge_loc false:
// it still takes part in the exchange, lsponge being the global flag).  Switches spongefilter to spongefilter_global.
int astr_gpu_set_sponge_global(const double* coef) {
  NEED_CTX();
  g->spg_global = true;
  if (g->spg_global_coef) { cudaFree(g->spg_global_coef); g->spg_global_coef = nullptr; }
  if (coef) {
    const astr_cfg& c = g->cfg;
    const size_t cnt = (size_t)(c.ie - c.is + 1) * (size_t)(c.je - c.js + 1) * (size_t)(c.ke - c.ks + 1);
    CUDA_OK(cudaMalloc(&g->spg_global_coef, cnt * sizeof(double)));
    CUDA_OK(cudaMemcpyAsync(g->spg_global_coef, coef, cnt * sizeof(double), cudaMemcpyHostToDevice, g->st));
    CUDA_OK(cudaStreamSynchronize(g->st));
  }
  return 0;
}

// spongefilter -> spongefilter_layer / spongefilter_global (src/sponge_layer.F90:55-364)
int astr_gpu_spongefilter(void) {
  NEED_CTX();
  const astr_cfg& c = g->cfg;
  if (g->spg_global) {
    // src/sponge_layer.F90:333: dataswap(q) in every direction, then the damped 7-point average over is:ie x js:je x ks:ke
    TRY(dataswap(fields(S_Q, 5), -1));
    if (!g->spg_global_coef) return 0;
    Box b = {{c.is, c.js, c.ks}, {c.ie, c.je, c.ke}};
    ProfScope ps(PC_FVAR);
    return pw_sponge(g->L, g->pool, b, g->spg_global_coef, g->st);
  }
  static const int faces[5] = {0, 1, 3, 4, 5};
  const FieldList fq = fields(S_Q, 5);
  for (int f : faces) {
    const Ctx::Sponge& sp = g->spg[f];
    if (!sp.on) continue;
    const int d = f / 2;
    TRY(dataswap(fq, d));
    if (sp.beg < 0) continue;
    Box b = {{c.is, c.js, c.ks}, {c.ie, c.je, c.ke}};
    b.lo[d] = sp.beg; b.hi[d] = sp.end;
    ProfScope ps(PC_FVAR);
    TRY(pw_sponge(g->L, g->pool, b, sp.coef, g->st));
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// crash control (lcracon), src/mainloop.F90:709-1198
// ---------------------------------------------------------------------------------
constexpr long long CRASHFIX_CAP = 1 << 20;       // flagged nodes one crashfix call can repair
static int ensure_crinod() {
  if (!g->crinod) {
    CUDA_OK(cudaMalloc(&g->crinod, (size_t)g->L.fstride * sizeof(double)));
    CUDA_OK(cudaMemsetAsync(g->crinod, 0, (size_t)g->L.fstride * sizeof(double), g->st));   // crinod=.false., :81
  }
  if (!g->d_count) CUDA_OK(cudaMalloc(&g->d_count, 2 * sizeof(unsigned long long)));
  return 0;
}
static int read_count(int which, long long* out) {
  unsigned long long v = 0;
  CUDA_OK(cudaMemcpyAsync(&v, g->d_count + which, sizeof v, cudaMemcpyDeviceToHost, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  *out = (long long)v;
  return 0;
}

// crashcheck (:709-768), the detection: *nbad = nodes of this rank whose density is not >= 0 (now critical nodes).
// The decision (`por`, recovery or stop, :770-812) stays with the caller.
int astr_gpu_crashcheck(long long* nbad) {
  NEED_CTX();
  TRY(ensure_crinod());
  CUDA_OK(cudaMemsetAsync(g->d_count, 0, 2 * sizeof(unsigned long long), g->st));
  TRY(pw_crashcheck(g->L, g->pool, g->crinod, g->d_count, g->st));
  TRY(read_count(0, nbad));
  return p2p_check();
}

// crinod_expansion (:985-1033): *counter as the reference counts it (27 per flagged node of -1..dim+1), this rank
int astr_gpu_crinod_expansion(long long* counter) {
  NEED_CTX();
  TRY(ensure_crinod());
  CUDA_OK(cudaMemsetAsync(g->d_count, 0, 2 * sizeof(unsigned long long), g->st));
  TRY(pw_crinod_dilate(g->L, g->crinod, g->slot(S_QRHS), g->d_count, g->st));      // qrhs is dead between steps
  FieldList fl; fl.nf = 1; fl.f[0] = g->crinod;
  TRY(dataswap(fl, -1));
  if (counter) TRY(read_count(0, counter));
  return 0;
}

// databakup (:826-972).  mode 0 'backup', 1 'recovery': two alternating device copies of q.  *slot = copy used
// (0 dat_a, 1 dat_b): the caller keeps / restores its own scalars (nstep, time, massflux, force ...) per copy;
// *recover_counter = recover_counter of that copy after this call.  A recovery also runs updatefvar and, from the
// second recovery of a copy on, crinod_expansion.
int astr_gpu_databakup(int mode, int* slot, int* recover_counter) {
  NEED_CTX();
  const size_t bytes = (size_t)5 * g->L.fstride * sizeof(double);
  if (mode == 0) {
    if (g->datpnt == 'o') g->datpnt = 'a';
    const int s = g->datpnt == 'a' ? 0 : 1;
    if (!g->bak[s]) CUDA_OK(cudaMalloc(&g->bak[s], bytes));
    CUDA_OK(cudaMemcpyAsync(g->bak[s], g->slot(S_Q), bytes, cudaMemcpyDeviceToDevice, g->st));
    g->bak_counter[s] = 0;
    g->datpnt = s == 0 ? 'b' : 'a';
    if (slot) *slot = s;
    if (recover_counter) *recover_counter = 0;
    return 0;
  }
  if (mode != 1) return astr_fail_msg("databakup: mode must be 0 (backup) or 1 (recovery)");
  if (g->datpnt == 'o') return astr_fail_msg("databakup: no backup data available");
  const int s = g->datpnt == 'a' ? 0 : 1;
  // after a single backup datpnt points at the copy that was never written (the reference then reads an
  // unallocated array, :941)
  if (!g->bak[s]) return astr_fail_msg("databakup: no backup data available");
  Box b = {{0, 0, 0}, {g->L.im, g->L.jm, g->L.km}};
  TRY(pw_copy_box(g->L, g->slot(S_Q), g->bak[s], 5, b, g->st));       // q(0:im,0:jm,0:km,:) only
  g->bak_counter[s] += 1;
  g->datpnt = s == 0 ? 'b' : 'a';
  if (slot) *slot = s;
  if (recover_counter) *recover_counter = g->bak_counter[s];
  TRY(astr_gpu_updatefvar());
  if (g->bak_counter[s] > 1) TRY(astr_gpu_crinod_expansion(nullptr));
  return 0;
}

// crashfix (:1045-1198): *nfixed = nodes of this rank that were wiped (the psum and the report stay with the caller);
// ig0, jg0: global index of the block's node 0 (module parallel), for the domain test of the neighbours (:1108-1110)
int astr_gpu_crashfix(int ig0, int jg0, long long* nfixed) {
  NEED_CTX();
  TRY(ensure_crinod());
  if (!g->d_list) CUDA_OK(cudaMalloc(&g->d_list, CRASHFIX_CAP * sizeof(long long)));
  CUDA_OK(cudaMemsetAsync(g->d_count, 0, 2 * sizeof(unsigned long long), g->st));
  const double eps_rho = 1.0e-5, eps_tmp = 1.0e-5;
  // thermal(density=eps_rho,temperature=eps_tmp), src/fludyna.F90:45-88
  const double eps_prs = g->th.nondimen ? eps_rho * eps_tmp / g->th.const2 : eps_rho * eps_tmp * g->th.rgas;
  TRY(pw_crashfix_flag(g->L, g->pool, g->crinod, eps_rho, eps_prs, eps_tmp, g->d_count, g->d_list, CRASHFIX_CAP, g->st));
  long long n = 0;
  TRY(read_count(0, &n));
  *nfixed = 0;
  if (n == 0) return 0;
  if (n > CRASHFIX_CAP) return astr_fail_msg("crashfix: more than 2^20 nodes of this rank have non-positive density, pressure or temperature");
  // storage order (i fastest), as the reference's loop visits them
  std::vector<long long> keys((size_t)n);
  CUDA_OK(cudaMemcpyAsync(keys.data(), g->d_list, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  std::sort(keys.begin(), keys.end());
  CUDA_OK(cudaMemcpyAsync(g->d_list, keys.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, g->st));
  CrashFixArgs a;
  const astr_cfg& c = g->cfg;
  a.g0[0] = ig0; a.g0[1] = jg0; a.g0[2] = 0; a.ia = c.ia; a.ja = c.ja; a.n = n;
  TRY(pw_crashfix_apply(g->L, g->pool, g->th, g->d_list, a, g->d_count + 1, g->st));
  TRY(read_count(1, nfixed));
  return 0;
}

// ---------------------------------------------------------------------------------
// checkpoint staging (writeflfed / readcheckpoint, src/readwrite.F90:1723-1984, :1381-1470): the six datasets
// ro, u1, u2, u3, p, t as dense node arrays (0:im,0:jm,0:km), Fortran order -- what h5write / h5read take.
// Two dense device buffers alternate: the pack of field m+1 runs under the copy of field m (copy stream g->xst).
// ---------------------------------------------------------------------------------
static int dense_buffers() {
  const size_t bytes = (size_t)(g->L.im + 1) * (g->L.jm + 1) * (g->L.km + 1) * sizeof(double);
  for (auto& d : g->dense) if (!d) CUDA_OK(cudaMalloc(&d, bytes));
  for (auto& e : g->ev_dense) if (!e) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return 0;
}
int astr_gpu_stage_checkpoint(double* ro, double* u1, double* u2, double* u3, double* p, double* t) {
  NEED_CTX();
  TRY(dense_buffers());
  double* host[6] = {ro, u1, u2, u3, p, t};
  const int slots[6] = {S_RHO, S_VEL, S_VEL + 1, S_VEL + 2, S_PRS, S_TMP};
  const size_t bytes = (size_t)(g->L.im + 1) * (g->L.jm + 1) * (g->L.km + 1) * sizeof(double);
  for (int m = 0; m < 6; ++m) {
    const int b = m & 1;
    if (m >= 2) CUDA_OK(cudaStreamWaitEvent(g->st, g->ev_dense[2 + b], 0));     // the copy that last used this buffer
    TRY(pw_dense(g->L, g->slot(slots[m]), g->dense[b], true, g->st));
    CUDA_OK(cudaEventRecord(g->ev_dense[b], g->st));
    CUDA_OK(cudaStreamWaitEvent(g->xst, g->ev_dense[b], 0));
    CUDA_OK(cudaMemcpyAsync(host[m], g->dense[b], bytes, cudaMemcpyDeviceToHost, g->xst));
    CUDA_OK(cudaEventRecord(g->ev_dense[2 + b], g->xst));
  }
  CUDA_OK(cudaStreamSynchronize(g->xst));
  CUDA_OK(cudaStreamSynchronize(g->st));
  return p2p_check();
}
// readcheckpoint (:1448-1453) + updateq (src/fludyna.F90:254-300: q from density, velocity and temperature)
int astr_gpu_restore_checkpoint(const double* ro, const double* u1, const double* u2, const double* u3, const double* p,
                                const double* t) {
  NEED_CTX();
  TRY(dense_buffers());
  const double* host[6] = {ro, u1, u2, u3, p, t};
  const int slots[6] = {S_RHO, S_VEL, S_VEL + 1, S_VEL + 2, S_PRS, S_TMP};
  const size_t bytes = (size_t)(g->L.im + 1) * (g->L.jm + 1) * (g->L.km + 1) * sizeof(double);
  for (int m = 0; m < 6; ++m) {
    const int b = m & 1;
    if (m >= 2) CUDA_OK(cudaStreamWaitEvent(g->xst, g->ev_dense[2 + b], 0));    // the unpack that last read this buffer
    CUDA_OK(cudaMemcpyAsync(g->dense[b], host[m], bytes, cudaMemcpyHostToDevice, g->xst));
    CUDA_OK(cudaEventRecord(g->ev_dense[b], g->xst));
    CUDA_OK(cudaStreamWaitEvent(g->st, g->ev_dense[b], 0));
    TRY(pw_dense(g->L, g->slot(slots[m]), g->dense[b], false, g->st));
    CUDA_OK(cudaEventRecord(g->ev_dense[2 + b], g->st));
  }
  TRY(pw_updateq(g->L, g->pool, g->th, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  return 0;
}

int astr_gpu_set_force(const double force[3]) {
  NEED_CTX();
  for (int i = 0; i < 3; ++i) g->force[i] = force[i];
  return 0;
}

static int reduce3(int what, int nout, double* out) {
  const astr_cfg& c = g->cfg;
  TRY(pw_reduce(g->L, g->pool, g->ycoord, g->th, what, c.rank[1] == 0, c.rank[1] == c.size[1] - 1, g->d_partial, g->d_out2,
                g->st));
  double h[3];
  CUDA_OK(cudaMemcpyAsync(h, g->d_out2, 3 * sizeof(double), cudaMemcpyDeviceToHost, g->st));
  CUDA_OK(cudaStreamSynchronize(g->st));
  for (int i = 0; i < nout; ++i) out[i] = h[i];
  return p2p_check();
}

int astr_gpu_reduce_tgv(double out[3]) {
  NEED_CTX();
  if (g->cfg.ndims != 3) return astr_fail_msg("reduce_tgv covers ndims=3 only");
  if (!g->have_grad) return astr_fail_msg("reduce_tgv before gradcal");
  return reduce3(0, 3, out);
}

int astr_gpu_reduce_cfl(double out[3]) {
  NEED_CTX();
  if (!g->have_metrics) return astr_fail_msg("reduce_cfl before set_metrics/gridgeom");
  return reduce3(1, 3, out);
}

int astr_gpu_reduce_channel(double out[2]) {
  NEED_CTX();
  if (!g->ycoord) return astr_fail_msg("reduce_channel needs node coordinates: call astr_gpu_set_grid or astr_gpu_gridgeom");
  if (!g->have_grad) return astr_fail_msg("reduce_channel before gradcal");
  return reduce3(2, 2, out);
}

int astr_gpu_kernel_launches(long long* count) { *count = g_launches; return 0; }

int astr_gpu_set_profile(int on) {
  NEED_CTX();
  g->profile = on != 0;
  if (on) { memset(g->prof_ms, 0, sizeof g->prof_ms); memset(g->prof_n, 0, sizeof g->prof_n); }
  return 0;
}
// ms[PC_COUNT], n[PC_COUNT]: accumulated CUDA-event time and span count per category
int astr_gpu_get_profile(double* ms, long long* n, int cap) {
  NEED_CTX();
  CUDA_OK(cudaStreamSynchronize(g->st));
  prof_collect();
  for (int i = 0; i < PC_COUNT && i < cap; ++i) { ms[i] = g->prof_ms[i]; n[i] = g->prof_n[i]; }
  return PC_COUNT;
}

// Times `iters` launches of one line-solve sweep over nfields scratch-free core fields
// (q -> G slots for the derivative, G in place for the filter) with CUDA events on the
// library stream.  Used by bench.py for the live roofline measurement.
int astr_gpu_bench_sweep(int op, int dir, int nfields, int iters, float* ms_per_launch) {
  NEED_CTX();
  if (nfields < 1 || nfields > 5 || dir < 0 || dir > 2) return astr_fail_msg("bench_sweep: bad arguments");
  // op: 0 derivative (q -> G slots), 1 filter (G in place); +10 swaps the placement (derivative in place, filter
  // out of place) for experiments
  const bool swap = op >= 10;
  op %= 10;
  if (op != OP_DERIV && op != OP_FILTER) return astr_fail_msg("bench_sweep: bad operator");
  const bool inplace = (op == OP_FILTER) != swap;
  const double* in[5]; double* out[5];
  for (int m = 0; m < nfields; ++m) {
    in[m] = inplace ? g->slot(S_G + m) : g->slot(S_Q + m);
    out[m] = g->slot(S_G + m);
  }
  cudaEvent_t a, b;
  CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
  TRY(sweep(dir, op, in, out, nfields, EPI_STORE, 0, dim_of(dir)));
  CUDA_OK(cudaEventRecord(a, g->st));
  for (int it = 0; it < iters; ++it) TRY(sweep(dir, op, in, out, nfields, EPI_STORE, 0, dim_of(dir)));
  CUDA_OK(cudaEventRecord(b, g->st));
  CUDA_OK(cudaEventSynchronize(b));
  float ms = 0.f;
  CUDA_OK(cudaEventElapsedTime(&ms, a, b));
  *ms_per_launch = ms / iters;
  cudaEventDestroy(a); cudaEventDestroy(b);
  return 0;
}

}  // extern "C"
