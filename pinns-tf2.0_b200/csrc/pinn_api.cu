// libpinn_b200.so -- C ABI (include/pinn_b200.h) over the fused sm_100a PINN kernels.
#include "../../include/pinn_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <mutex>
#include <vector>

#include "burgers_fused.cuh"
#include "burgers_fused_v2.cuh"
#include "nls_fused.cuh"
#include "optim_kernels.cuh"
#include "generic_fused.cuh"

namespace {

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return -1; }

// Live handles of this process.  A launch with the in-kernel tail (fused_tail) keeps its last CTAs spinning on their SMs until the
// launch's stragglers arrive; two such launches running CONCURRENTLY on one device (two handles, two streams) could fill every SM
// with spinning CTAs while both still have CTAs waiting for an SM.  So a handle uses the in-kernel tail only while no OTHER handle on
// the same device has such a launch in flight (checked with cudaStreamQuery); otherwise it falls back to the tail kernels.
std::mutex g_live_mu;
std::vector<struct pinn_handle*> g_live;

#define CUDA_TRY(expr)                                                                                         \
  do {                                                                                                         \
    cudaError_t _e = (expr);                                                                                   \
    if (_e != cudaSuccess)                                                                                     \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

// ---- NCCL, resolved at run time (torch's bundled libnccl.so.2 is already in the process under torchrun;
// otherwise the system one is opened).  Only the five entry points the path needs.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& why) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { why = "NCCL symbols missing"; return false; }
    return true;
  }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8;   // ncclDouble
constexpr int NCCL_SUM = 0;       // ncclSum

constexpr int LOSS_RING = 4096;

// buffers of the Gram-matrix L-BFGS formulation (optim_kernels.cuh: lbfgs_dots / lbfgs_solve / lbfgs_apply)
struct LbfgsGram { double *SY = nullptr, *YY = nullptr, *part = nullptr; int n_corr_cap = 0, n_blocks = 0, stride = 0; };

}  // namespace

struct pinn_handle {
  int pde = 0, device = 0, rank = 0, world = 1;
  std::vector<int> layers;
  double lb[2] = {0, 0}, ub[2] = {1, 1};
  int P_net = 0, P = 0;             // net parameters; P = P_net (+2 identification)
  double nu = 0.0, dt = 0.0;
  double* d_irk = nullptr;          // DISC: (q+1) x q stage matrix
  int irk_q = 0;
  std::vector<double> h_x0;         // DISC: data x positions (host copy for re-assembly); IDE_DISC: snapshot 0 (x_1 lives in h_tb)
  std::vector<double> h_u0, h_u1;   // IDE_DISC: targets of the two snapshots
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int n_sm = 0;
  long long launches = 0;
  bool tail_inflight = false;   // a launch with the in-kernel tail may still be running on this handle's stream
  pinn::ReduceMap last_map{};       // reduction map / grid / stride of the most recent fused launch
  int last_grid = 0, last_stride = 0;
  int kernel_kind = 0;              // 0: specialised Burgers DMMA kernel, 1: specialised NLS DMMA kernel, 2: generic
  double *d_gH = nullptr, *d_gA = nullptr, *d_gS = nullptr;   // generic kernel scratch
  long long g_pts = 0;
  int burgers_kernel = 2;           // 2: warp-specialised (default); 1: single-role v1 (PINN_BURGERS_KERNEL=v1)

  // parameters and optimiser state
  double* d_w = nullptr;            // WPAD-padded flat weights
  int w_cap = 0;
  double* d_R = nullptr;            // [P grad | 3 loss parts | pad]
  double* d_partials = nullptr;
  int n_cta = 0, pstride = 0;
  double *d_m = nullptr, *d_v = nullptr;
  int* d_step = nullptr;
  double* d_loss_ring = nullptr;
  long long adam_steps = 0;

  // point sets (device SoA x[], t[]).  Layout: [data region, capacity dcap | collocation region, capacity ccap];
  // data points are stored right-aligned in their region so that [data | collocation] is contiguous:
  // kernel point 0 is at offset dcap - n_d.  A per-step collocation upload touches only its own region.
  double *d_x = nullptr, *d_t = nullptr, *d_u = nullptr;
  long long dcap = 0, ccap = 0, u_cap = 0;
  long long n_c = 0, n_c_global = 0, n_d = 0, n_b = 0;
  int d_out_dim = 1;
  double data_weight = 1.0;
  std::vector<double> h_tb;                 // NLS boundary times
  std::vector<double> h_icx, h_ict;         // NLS initial-condition points (host copy for re-assembly)
  const double *map_x = nullptr, *map_t = nullptr;   // zero-copy collocation block (device aliases of pinned host memory)
  long long n_aux = 0;                      // points stored in the data region (Burgers: n_d; NLS: n0p + 2 n_b)
  double *d_scrH = nullptr, *d_scrA = nullptr, *d_scrS = nullptr;   // NLS activation / adjoint / seed scratch
  int scr_pts = 0;
  std::vector<double> stage_x, stage_t;     // host staging for set_data (de-interleave)

  // measurement
  std::vector<cudaEvent_t> events;
  double* d_flush = nullptr;
  size_t flush_bytes = 0;

  // L-BFGS
  pinn::LbfgsState* d_lb = nullptr;
  double *d_gold = nullptr, *d_d = nullptr, *d_S = nullptr, *d_Y = nullptr, *d_xfinal = nullptr, *d_fhist = nullptr;
  int* d_logged = nullptr;
  int lb_corr_cap = 0, lb_iter_cap = 0;
  LbfgsGram lb_gram;
  std::vector<double> lb_fhist;             // f of every evaluation of the last pinn_lbfgs run (custom_lbfgs.py f_hist)

  // scratch for predict / derivatives
  double *d_px = nullptr, *d_pout = nullptr;
  long long px_cap = 0, pout_cap = 0;

  ncclComm_t comm = nullptr;
  // fused NVLink push exchange (optim_kernels.cuh: reduce_exchange)
  double* d_xchg = nullptr;            // [2][world][slot_len] doubles, then [2][world][n_blocks] u64 flags (IPC-exported)
  bool p2p_ready = false, p2p_mapped = false;
  pinn::XchgPeers peers{};
  void* peer_base[pinn::P2P_MAX] = {nullptr};
  int* d_xseq = nullptr;               // [0] published evaluations, [1] block counter
  int* d_p2p_err = nullptr;
};

namespace {

bool is_burgers_net(const std::vector<int>& L) {
  if (L.size() != 10 || L[0] != 2 || L[9] != 1) return false;
  for (int i = 1; i <= 8; i++) if (L[i] != 20) return false;
  return true;
}
bool is_nls_net(const std::vector<int>& L) {
  if (L.size() != 6 || L[0] != 2 || L[5] != 2) return false;
  for (int i = 1; i <= 4; i++) if (L[i] != 100) return false;
  return true;
}

int nls_upload_points(pinn_t* h);
int nls_launch_eval(pinn_t* h, const int* run_flag);
int generic_launch_eval(pinn_t* h, const int* run_flag);
int disc_upload_points(pinn_t* h);

// One L-BFGS iteration on the device (stop tests of the pending evaluation, history update, direction, step).
// Default: the single-CTA kernel that follows the reference's loop literally (lbfgs_iterate).
// PINN_LBFGS=gram selects the Gram-matrix formulation (lbfgs_dots -> lbfgs_solve -> lbfgs_apply: every SM reads its slice of
// the history; ~12 us less per iteration at n_corr = 50).  It is NOT the default: in 1 of 8 long fixed-step runs
// (profiles/lbfgs_stability_r02.jsonl) its coefficient-space recursion produced a bad direction at iteration ~400 where the
// literal loop did not -- the triangular recurrence over s_i.y_m amplifies rounding when the stored pairs are nearly
// dependent, whereas the literal loop measures every s_i.q against the actual q -- and the reference has no line search to
// recover from one bad step.
bool lbfgs_serial() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINN_LBFGS"); v = (e && !strcmp(e, "gram")) ? 0 : 1; }
  return v == 1;
}
int lbfgs_gram_ensure(LbfgsGram* gm, int n_corr, int P) {
  const int nb = (P + pinn::LB_CHUNK - 1) / pinn::LB_CHUNK;
  if (n_corr <= gm->n_corr_cap && nb <= gm->n_blocks) return 0;
  if (gm->SY) cudaFree(gm->SY);
  if (gm->YY) cudaFree(gm->YY);
  if (gm->part) cudaFree(gm->part);
  gm->SY = gm->YY = gm->part = nullptr;
  const size_t NS = (size_t)n_corr + 1;
  gm->stride = pinn::LB_NSCAL + 4 * n_corr + 1;
  if (cudaMalloc((void**)&gm->SY, NS * NS * 8) != cudaSuccess || cudaMalloc((void**)&gm->YY, NS * NS * 8) != cudaSuccess ||
      cudaMalloc((void**)&gm->part, (size_t)nb * gm->stride * 8) != cudaSuccess)
    return fail("cudaMalloc failed (L-BFGS Gram buffers)");
  gm->n_corr_cap = n_corr; gm->n_blocks = nb;
  if ((size_t)n_corr * n_corr * 8 > 48 * 1024 &&
      cudaFuncSetAttribute(pinn::lbfgs_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, n_corr * n_corr * 8) != cudaSuccess)
    return fail("cudaFuncSetAttribute(lbfgs_solve) failed");
  return 0;
}
void lbfgs_gram_free(LbfgsGram* gm) {
  if (gm->SY) cudaFree(gm->SY);
  if (gm->YY) cudaFree(gm->YY);
  if (gm->part) cudaFree(gm->part);
  *gm = LbfgsGram{};
}
int lbfgs_launch_iteration(cudaStream_t stream, pinn::LbfgsState* st, double* w, const double* R, int P, double* gold, double* d,
                           double* S, double* Y, double* xfinal, double* fhist, int* logged, const LbfgsGram& gm, long long* launches) {
  if (lbfgs_serial()) {
#define LB_LAUNCH(E, NT) pinn::lbfgs_iterate<E, NT><<<1, NT, 0, stream>>>(st, w, R, P, gold, d, S, Y, xfinal, fhist, logged)
    if (P <= 12 * 256) LB_LAUNCH(12, 256);        // Burgers-size vectors: 8 warps, 12 entries per thread
    else if (P <= 8 * 1024) LB_LAUNCH(8, 1024);
    else LB_LAUNCH(32, 1024);
#undef LB_LAUNCH
    if (cudaGetLastError() != cudaSuccess) return fail("lbfgs_iterate launch failed");
    if (launches) *launches += 1;
    return 0;
  }
  const int nb = (P + pinn::LB_CHUNK - 1) / pinn::LB_CHUNK;
  pinn::lbfgs_dots<<<nb, pinn::LB_DOT_THREADS, 0, stream>>>(st, R, P, gold, d, S, Y, gm.part, gm.stride);
  pinn::lbfgs_solve<<<1, 128, (size_t)gm.n_corr_cap * gm.n_corr_cap * 8, stream>>>(st, R, P, gm.part, nb, gm.stride, gm.SY, gm.YY, fhist,
                                                                                   logged);
  pinn::lbfgs_apply<<<nb, pinn::LB_CHUNK, 0, stream>>>(st, w, R, P, gold, d, S, Y, xfinal);
  if (cudaGetLastError() != cudaSuccess) return fail("L-BFGS iteration launch failed");
  if (launches) *launches += 3;
  return 0;
}

// after a stream synchronisation: has the fused P2P exchange reported a peer that never published?
int check_p2p(pinn_t* h) {
  if (!h->p2p_ready) return 0;
  int e = 0;
  if (cudaMemcpy(&e, h->d_p2p_err, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("reading the P2P error flag failed");
  if (e) return fail("P2P exchange timed out waiting for a peer rank (a rank died or the evaluation counts diverged); "
                     "weights and optimiser state were left untouched by that step");
  return 0;
}

pinn::NetDesc net_desc(const pinn_t* h) {
  pinn::NetDesc nd{};
  nd.n_layers = (int)h->layers.size() - 1;
  int o = 0;
  for (int l = 0; l < nd.n_layers; l++) {
    nd.dims[l] = h->layers[l];
    nd.woff[l] = o; o += h->layers[l] * h->layers[l + 1];
    nd.boff[l] = o; o += h->layers[l + 1];
  }
  nd.dims[nd.n_layers] = h->layers.back();
  nd.lb0 = h->lb[0]; nd.lb1 = h->lb[1];
  nd.dx0 = h->ub[0] - h->lb[0]; nd.dx1 = h->ub[1] - h->lb[1];
  return nd;
}

int ensure(double** p, long long* cap, long long need) {
  if (need <= *cap) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr;
  long long c = need + need / 8 + 64;
  if (cudaMalloc((void**)p, (size_t)c * sizeof(double)) != cudaSuccess) return fail("cudaMalloc failed");
  *cap = c;
  return 0;
}

// grow the point arrays, preserving both regions
int ensure_points(pinn_t* h, long long need_d, long long need_c) {
  if (need_d <= h->dcap && need_c <= h->ccap && h->d_x) return 0;
  long long nd = h->dcap, nc = h->ccap;
  if (need_d > nd) nd = need_d + need_d / 4 + 1024;
  if (need_c > nc) nc = need_c + need_c / 8 + 1024;
  if (nd < 1024) nd = 1024;
  double *nx = nullptr, *nt = nullptr;
  CUDA_TRY(cudaMalloc((void**)&nx, (size_t)(nd + nc) * 8));
  CUDA_TRY(cudaMalloc((void**)&nt, (size_t)(nd + nc) * 8));
  if (h->d_x) {
    if (h->n_aux) {
      CUDA_TRY(cudaMemcpyAsync(nx + nd - h->n_aux, h->d_x + h->dcap - h->n_aux, h->n_aux * 8, cudaMemcpyDeviceToDevice, h->stream));
      CUDA_TRY(cudaMemcpyAsync(nt + nd - h->n_aux, h->d_t + h->dcap - h->n_aux, h->n_aux * 8, cudaMemcpyDeviceToDevice, h->stream));
    }
    // the device-resident collocation block: nothing to preserve while the batch is a zero-copy mapping of pinned host
    // memory (n_c then counts the HOST batch and ccap may be 0), and never more than the old region held
    const long long keep_c = h->map_x ? 0 : (h->n_c < h->ccap ? h->n_c : h->ccap);
    if (keep_c) {
      CUDA_TRY(cudaMemcpyAsync(nx + nd, h->d_x + h->dcap, keep_c * 8, cudaMemcpyDeviceToDevice, h->stream));
      CUDA_TRY(cudaMemcpyAsync(nt + nd, h->d_t + h->dcap, keep_c * 8, cudaMemcpyDeviceToDevice, h->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_x); cudaFree(h->d_t);
  }
  h->d_x = nx; h->d_t = nt; h->dcap = nd; h->ccap = nc;
  return 0;
}

struct AdamArgs { double lr, b1, b2, eps; };

// Launch a tail kernel, optionally with programmatic stream serialization (PDL, PINN_PDL=1): its blocks may then be placed while
// the fused kernel that precedes it in the stream drains, and wait in pdl_wait() for its completion.  OFF by default: measured
// on 1xB200 the step got 5 us SLOWER with it (0.4094 vs 0.4044 ms: the early-placed tail blocks sit on the SMs whose fused CTA
// finished first and the launch latency they were meant to hide is smaller than what their parking costs).
bool use_pdl() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINN_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_tail_kernel(void (*kernel)(KArgs...), int grid, int block, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = use_pdl() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// The tail of an evaluation: fixed-order reduction of the per-CTA partials into R = [grad | loss parts] on every rank,
// optionally fused with the Adam update.
// (Evaluations of the specialised Burgers kernel do all of this in their own last CTAs -- fused_tail -- and never get here.)
//   world == 1      : reduce_partials, or reduce_adam (one kernel)
//   world > 1, P2P  : reduce_exchange -- reduction + NVLink all-to-all push + rank-ordered sum (+ Adam) in ONE kernel
//   world > 1, NCCL : reduce_partials + ncclAllReduce(P+3 doubles) (+ adam_update)
int launch_tail(pinn_t* h, const int* run_flag, const AdamArgs* ad) {
  const pinn::ReduceMap& map = h->last_map;
  const int nb = (map.n_out + 31) / 32;
  if (h->world == 1) {
    if (ad)
      CUDA_TRY(launch_tail_kernel(pinn::reduce_adam, nb, 256, h->stream, h->d_partials, h->last_grid, h->last_stride, h->d_R, map, h->d_w,
                                  h->d_m, h->d_v, h->P, h->d_step, ad->lr, ad->b1, ad->b2, ad->eps, h->d_loss_ring, LOSS_RING));
    else
      CUDA_TRY(launch_tail_kernel(pinn::reduce_partials, nb, 256, h->stream, h->d_partials, h->last_grid, h->last_stride, h->d_R, map,
                                  run_flag));
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return 0;
  }
  if (h->p2p_ready) {
    if (nb != h->peers.n_blocks || map.n_out > h->peers.slot_len) return fail("internal: exchange buffer geometry mismatch");
    const int xgrid = nb < 2 * h->n_sm ? nb : 2 * h->n_sm;      // certainly co-resident: <= 2 blocks of 256 threads per SM
    CUDA_TRY(launch_tail_kernel(pinn::reduce_exchange, xgrid, 256, h->stream, h->d_partials, h->last_grid, h->last_stride, map, run_flag,
                                h->peers, h->d_xseq, h->d_R, h->d_p2p_err, ad ? 1 : 0, h->d_w, h->d_m, h->d_v, h->P, h->d_step,
                                ad ? ad->lr : 0.0, ad ? ad->b1 : 0.0, ad ? ad->b2 : 0.0, ad ? ad->eps : 0.0, h->d_loss_ring,
                                LOSS_RING));
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return 0;
  }
  CUDA_TRY(launch_tail_kernel(pinn::reduce_partials, nb, 256, h->stream, h->d_partials, h->last_grid, h->last_stride, h->d_R, map, run_flag));
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  // one exchange over [gradient | loss parts].  A skipped evaluation (L-BFGS stopped) still takes part with
  // stale but rank-identical participation so that ranks never diverge.
  int rc = g_nccl.AllReduce(h->d_R, h->d_R, (size_t)h->P + 3, NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
  if (rc != 0) return fail(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
  if (ad) {
    pinn::adam_update<<<(h->P + 127) / 128, 128, 0, h->stream>>>(h->d_w, h->d_m, h->d_v, h->d_R, h->P, h->d_step, ad->lr, ad->b1,
                                                                   ad->b2, ad->eps, h->d_loss_ring, LOSS_RING);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
  }
  return 0;
}

// true when no other live handle on h's device may still be running a launch with the in-kernel tail
bool tail_slot_free(pinn_t* h) {
  std::lock_guard<std::mutex> lk(g_live_mu);
  for (pinn_t* o : g_live) {
    if (o == h || o->device != h->device || !o->tail_inflight) continue;
    if (cudaStreamQuery(o->stream) == cudaSuccess) o->tail_inflight = false;
    else return false;
  }
  (void)cudaGetLastError();               // cudaErrorNotReady from the query is not an error of ours
  return true;
}

// v2 kernel with tail_done != nullptr: the launch also reduces the partials (+ exchanges them, + applies Adam when ad != nullptr) in
// its last CTAs (fused_tail); *tail_done tells the caller that no tail kernel is needed.  PINN_FUSED_TAIL=0 keeps the tail kernels.
int burgers_launch_eval(pinn_t* h, const int* run_flag, const AdamArgs* ad = nullptr, bool* tail_done = nullptr) {
  namespace B = pinn::burgers;
  const bool ide = h->pde == PINN_BURGERS_IDE;
  const long long n_total = ide ? h->n_d : h->n_d + h->n_c;
  if (n_total <= 0) return fail("no points set (pinn_set_collocation / pinn_set_data)");
  B::Args a{};
  a.w = h->d_w;
  a.x = h->d_x + (h->dcap - h->n_d); a.t = h->d_t + (h->dcap - h->n_d); a.utgt = h->d_u;
  a.xc = h->map_x; a.tc = h->map_t;
  if (h->map_x && h->burgers_kernel != 2) return fail("zero-copy collocation needs the v2 Burgers kernel");
  a.n_total = n_total;
  a.c0 = ide ? 0 : h->n_d;
  a.n_c = ide ? h->n_d : h->n_c;
  a.d0 = 0;
  a.n_d = h->n_d;
  const long long nfg = ide ? h->n_d : h->n_c_global;
  a.wf = nfg > 0 ? 1.0 / (double)nfg : 0.0;
  a.wd = h->n_d > 0 ? h->data_weight / (double)h->n_d : 0.0;
  a.lb0 = h->lb[0]; a.lb1 = h->lb[1];
  a.dx0 = h->ub[0] - h->lb[0]; a.dx1 = h->ub[1] - h->lb[1];
  a.nu = h->nu; a.ide = ide ? 1 : 0;
  a.partials = h->d_partials;
  a.run_flag = run_flag;
  // v2: tile-granular distribution inside the kernel (any grid <= tiles); v1: rounds of 32 points striped over the CTAs
  a.chains = 4;
  const long long n_tiles = (n_total + B::TILE - 1) / B::TILE;
  const long long rounds = (n_total + B::ROUND - 1) / B::ROUND;
  const long long units = h->burgers_kernel == 2 ? n_tiles : rounds;
  int grid = (int)(units < h->n_cta ? units : h->n_cta);
  pinn::ReduceMap map{};
  map.p_net = B::P_NET;
  map.n_extra = 0;
  if (ide) { map.extra_src[map.n_extra++] = B::IDX_DL1; map.extra_src[map.n_extra++] = B::IDX_DL2; }
  map.extra_src[map.n_extra++] = B::IDX_LD;
  map.extra_src[map.n_extra++] = 3023;        // (boundary part: always 0 for Burgers)
  map.extra_src[map.n_extra++] = B::IDX_LF;
  map.n_out = map.p_net + map.n_extra;
  static const bool fused_tail_on = [] { const char* e = getenv("PINN_FUSED_TAIL"); return !(e && e[0] == '0'); }();
  const bool xchg_ok = h->world > 1 && h->p2p_ready && (map.n_out + 31) / 32 == h->peers.n_blocks && map.n_out <= h->peers.slot_len;
  // every evaluation of the v2 kernel carries its own tail: Adam steps (ad != nullptr) and plain evaluations (L-BFGS, loss/gradient
  // queries; a launch skipped through run_flag skips its tail with it, exactly like the stand-alone tail kernels)
  if (tail_done && fused_tail_on && (h->world == 1 || xchg_ok) && h->burgers_kernel == 2 && grid <= h->n_sm && tail_slot_free(h)) {
    h->tail_inflight = true;
    pinn::FusedTail& ft = a.tail;
    ft.enabled = 1; ft.adam = ad ? 1 : 0; ft.ctr = h->d_step + 2; ft.R = h->d_R; ft.map = map;
    if (h->world > 1) { ft.xchg = 1; ft.peers = h->peers; ft.xseq = h->d_xseq; ft.err = h->d_p2p_err; }
    ft.w = h->d_w; ft.m = h->d_m; ft.v = h->d_v; ft.P = h->P; ft.step = h->d_step;
    if (ad) { ft.lr = ad->lr; ft.b1 = ad->b1; ft.b2 = ad->b2; ft.eps = ad->eps; }
    ft.loss_ring = h->d_loss_ring; ft.ring = LOSS_RING;
    *tail_done = true;
  }
  if (h->burgers_kernel == 2)
    pinn::burgers2::fused_loss_grad<<<grid, pinn::burgers2::THREADS, pinn::burgers2::SMEM_BYTES, h->stream>>>(a);
  else
    B::fused_loss_grad<<<grid, B::THREADS, B::SMEM_BYTES, h->stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  h->last_map = map; h->last_grid = grid; h->last_stride = B::PSTRIDE;
  return 0;
}

// one evaluation: the fused loss/gradient kernel of the handle's PDE, then (unless fused_only) the tail
int launch_eval(pinn_t* h, const int* run_flag, bool fused_only = false, const AdamArgs* ad = nullptr) {
  int rc;
  bool tail_done = false;
  if (h->kernel_kind == 2) rc = generic_launch_eval(h, run_flag);
  else if (h->pde == PINN_BURGERS_INF || h->pde == PINN_BURGERS_IDE) rc = burgers_launch_eval(h, run_flag, ad, fused_only ? nullptr : &tail_done);
  else rc = nls_launch_eval(h, run_flag);
  if (rc) return -1;
  if (fused_only || tail_done) return 0;
  return launch_tail(h, run_flag, ad);
}

// DISC: data region = [x_0 data points | x_1 boundary points], right-aligned like every data block (t is unused)
int disc_upload_points(pinn_t* h) {
  const long long n_aux = h->n_d + h->n_b;
  if (ensure_points(h, n_aux > h->n_aux ? n_aux : h->n_aux, 0)) return -1;
  h->stage_x.assign(n_aux, 0.0); h->stage_t.assign(n_aux, 0.0);
  for (long long i = 0; i < h->n_d; i++) h->stage_x[i] = h->h_x0[i];
  for (long long i = 0; i < h->n_b; i++) h->stage_x[h->n_d + i] = h->h_tb[i];
  if (n_aux) {
    CUDA_TRY(cudaMemcpyAsync(h->d_x + h->dcap - n_aux, h->stage_x.data(), n_aux * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_t + h->dcap - n_aux, h->stage_t.data(), n_aux * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  h->n_aux = n_aux;
  return 0;
}

int generic_launch_eval(pinn_t* h, const int* run_flag) {
  namespace G = pinn::generic;
  const bool idd = h->pde == PINN_BURGERS_IDE_DISC;
  const bool disc = h->pde == PINN_BURGERS_DISC || idd;
  if (disc && (!h->d_irk || h->irk_q + (idd ? 0 : 1) != h->layers.back())) return fail("discrete-time model: call pinn_set_irk first");
  const bool nls = h->pde == PINN_NLS_INF, ide = h->pde == PINN_BURGERS_IDE;
  const long long n_total = disc ? h->n_aux : (nls ? h->n_aux + h->n_c : (ide ? h->n_d : h->n_d + h->n_c));
  if (n_total <= 0) return fail("no points set (pinn_set_collocation / pinn_set_data)");
  const long long per = (n_total + h->n_cta - 1) / h->n_cta;
  // points per CTA: a multiple of 16 for large sets; small sets (the discrete-time models have ~250 points) are spread over as
  // many CTAs as possible -- an even count, because NLS boundary pairs (lb_k, ub_k) are adjacent and must not straddle a CTA
  const long long pts = per >= 16 ? (per + 15) / 16 * 16 : (per + 1) / 2 * 2;
  // only CTAs that own points are launched: every CTA writes (and the tail kernel reads) a full partial vector
  const int grid = (int)((n_total + pts - 1) / pts);
  pinn::NetDesc nd = net_desc(h);
  int maxw = 0; long long hsum = 0;
  for (int l = 0; l < nd.n_layers - 1; l++) { hsum += 4 * pts * nd.dims[l + 1]; if (nd.dims[l + 1] > maxw) maxw = nd.dims[l + 1]; }
  const int out_w = nd.dims[nd.n_layers];
  if (out_w > maxw) maxw = out_w;                 // the IRK head reuses the adjoint scratch for N, N-bar and 2(U_0 - u_0)
  const long long a_per = 4 * pts * maxw;
  if (pts > h->g_pts) {
    if (h->d_gH) cudaFree(h->d_gH);
    if (h->d_gA) cudaFree(h->d_gA);
    if (h->d_gS) cudaFree(h->d_gS);
    h->d_gH = h->d_gA = h->d_gS = nullptr;
    CUDA_TRY(cudaMalloc((void**)&h->d_gH, (size_t)h->n_cta * hsum * 8));          // sized for the largest grid (all CTAs)
    CUDA_TRY(cudaMalloc((void**)&h->d_gA, (size_t)h->n_cta * 2 * a_per * 8));
    CUDA_TRY(cudaMalloc((void**)&h->d_gS, (size_t)h->n_cta * 2 * pts * 4 * out_w * 8));
    h->g_pts = pts;
  }
  G::Args a{};
  a.w = h->d_w; a.nd = nd;
  const long long n_front = (nls || disc) ? h->n_aux : h->n_d;
  a.x = h->d_x + (h->dcap - n_front); a.t = h->d_t + (h->dcap - n_front); a.tgt = h->d_u;
  a.n_total = n_total; a.pde = h->pde;
  a.in_dim = disc ? 1 : 2; a.dt = h->dt; a.irk = h->d_irk;
  a.c0 = ide ? 0 : h->n_d; a.n_c = ide ? h->n_d : h->n_c; a.d0 = 0; a.n_d = h->n_d;
  const long long nfg = ide ? h->n_d : h->n_c_global;
  a.wf = nfg > 0 ? 1.0 / (double)nfg : 0.0;
  a.wd = h->n_d > 0 ? h->data_weight / (double)h->n_d : 0.0;
  a.nu = h->nu;
  a.n0 = h->n_d; a.n0p = (h->n_d + 1) & ~1LL; a.nb = h->n_b;
  a.w0 = h->n_d > 0 ? h->data_weight / (double)h->n_d : 0.0;
  a.wb = h->n_b > 0 ? h->data_weight / (double)h->n_b : 0.0;
  if (nls && h->n_aux != a.n0p + 2 * a.nb) return fail("internal: NLS auxiliary block out of date");
  a.p_net = h->P_net;
  a.scrH = h->d_gH; a.scrA = h->d_gA; a.scrS = h->d_gS;
  a.h_per_cta = hsum; a.a_per_cta = a_per;
  a.pts = (int)pts; a.maxw = maxw;
  static const bool generic_dfma = [] { const char* e = getenv("PINN_GENERIC_DFMA"); return e && e[0] == '1'; }();
  a.dmma = generic_dfma ? 0 : 1;
  a.partials = h->d_partials; a.pstride = h->pstride; a.run_flag = run_flag;
  G::fused_loss_grad<<<grid, G::THREADS, 0, h->stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  pinn::ReduceMap map{};
  map.p_net = h->P_net;
  map.n_extra = 0;
  if (ide || idd) { map.extra_src[map.n_extra++] = h->P_net + 0; map.extra_src[map.n_extra++] = h->P_net + 1; }
  map.extra_src[map.n_extra++] = h->P_net + 3;
  map.extra_src[map.n_extra++] = h->P_net + 4;
  map.extra_src[map.n_extra++] = h->P_net + 5;
  map.n_out = map.p_net + map.n_extra;
  h->last_map = map; h->last_grid = grid; h->last_stride = h->pstride;
  return 0;
}

// NLS: (re)assemble the auxiliary block [initial-condition points | pad to even | (lb_k, ub_k) pairs] right-aligned in
// the data region (inf_cont_schrodinger.py:50-53 builds X_lb=(lb0,tb), X_ub=(ub0,tb)).
int nls_upload_points(pinn_t* h) {
  const long long n0 = h->n_d, n0p = (n0 + 1) & ~1LL, nb = h->n_b;
  const long long n_aux = n0p + 2 * nb;
  // the right-aligned start moves with n_aux: grow first with the OLD count so nothing is lost, then rewrite
  if (ensure_points(h, n_aux > h->n_aux ? n_aux : h->n_aux, h->n_c)) return -1;
  h->stage_x.assign(n_aux, 0.0); h->stage_t.assign(n_aux, 0.0);
  for (long long i = 0; i < n0; i++) { h->stage_x[i] = h->h_icx[i]; h->stage_t[i] = h->h_ict[i]; }
  if (n0p > n0) { h->stage_x[n0] = h->lb[0]; h->stage_t[n0] = h->lb[1]; }     // inert alignment point
  for (long long k = 0; k < nb; k++) {
    h->stage_x[n0p + 2 * k] = h->lb[0];     h->stage_t[n0p + 2 * k] = h->h_tb[k];
    h->stage_x[n0p + 2 * k + 1] = h->ub[0]; h->stage_t[n0p + 2 * k + 1] = h->h_tb[k];
  }
  if (n_aux) {
    CUDA_TRY(cudaMemcpyAsync(h->d_x + h->dcap - n_aux, h->stage_x.data(), n_aux * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_t + h->dcap - n_aux, h->stage_t.data(), n_aux * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  h->n_aux = n_aux;
  return 0;
}

int nls_launch_eval(pinn_t* h, const int* run_flag) {
  namespace N = pinn::nls;
  const long long n0 = h->n_d, n0p = (n0 + 1) & ~1LL, nb = h->n_b;
  const long long n_total = h->n_aux + h->n_c;
  if (n_total <= 0) return fail("no points set (pinn_set_collocation / pinn_set_data / pinn_set_boundary)");
  if (h->n_aux != n0p + 2 * nb) return fail("internal: NLS auxiliary block out of date");
  const int grid = h->n_cta;
  long long per = (n_total + grid - 1) / grid;
  const int pts = (int)((per + N::RPTS - 1) / N::RPTS * N::RPTS);
  if (pts > h->scr_pts) {
    if (h->d_scrH) cudaFree(h->d_scrH);
    if (h->d_scrA) cudaFree(h->d_scrA);
    if (h->d_scrS) cudaFree(h->d_scrS);
    h->d_scrH = h->d_scrA = h->d_scrS = nullptr;
    CUDA_TRY(cudaMalloc((void**)&h->d_scrH, (size_t)grid * 16 * pts * N::W * 8));
    CUDA_TRY(cudaMalloc((void**)&h->d_scrA, (size_t)grid * 8 * pts * N::W * 8));
    CUDA_TRY(cudaMalloc((void**)&h->d_scrS, (size_t)grid * pts * 8 * 8));
    h->scr_pts = pts;
  }
  N::Args a{};
  a.w = h->d_w;
  a.x = h->d_x + (h->dcap - h->n_aux); a.t = h->d_t + (h->dcap - h->n_aux); a.uv0 = h->d_u;
  a.n_total = n_total; a.n0 = n0; a.n0p = n0p; a.nb = nb; a.nc = h->n_c;
  a.w0 = n0 > 0 ? h->data_weight / (double)n0 : 0.0;
  a.wb = nb > 0 ? h->data_weight / (double)nb : 0.0;
  a.wf = h->n_c_global > 0 ? 1.0 / (double)h->n_c_global : 0.0;
  a.lb0 = h->lb[0]; a.lb1 = h->lb[1]; a.dx0 = h->ub[0] - h->lb[0]; a.dx1 = h->ub[1] - h->lb[1];
  a.scratchH = h->d_scrH; a.scratchA = h->d_scrA; a.scratchS = h->d_scrS;
  a.pts = h->scr_pts;
  // the kernel derives its point range from a.pts, so keep the distribution even when the scratch is larger
  a.pts = pts;
  a.partials = h->d_partials; a.run_flag = run_flag;
  N::fused_loss_grad<<<grid, N::THREADS, N::SMEM_BYTES, h->stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  pinn::ReduceMap map{};
  map.p_net = N::P_NET;
  map.n_extra = 3;
  map.extra_src[0] = N::IDX_L0; map.extra_src[1] = N::IDX_LB; map.extra_src[2] = N::IDX_LF;
  map.n_out = map.p_net + map.n_extra;
  h->last_map = map; h->last_grid = grid; h->last_stride = N::PSTRIDE;
  return 0;
}

}  // namespace

extern "C" {

const char* pinn_last_error(void) { return g_err.c_str(); }
const char* pinn_version(void) { return "pinn_b200 0.1 (sm_100a, fp64 DMMA)"; }

int pinn_p2p_export(pinn_t* h, void* out64) {
  if (!h || !out64) return fail("pinn_p2p_export: null argument");
  if (h->world < 2 || h->world > pinn::P2P_MAX) return fail("pinn_p2p_export: needs 2..8 ranks");
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_xchg) {
    const int n = h->P + 3;
    h->peers.world = h->world; h->peers.rank = h->rank;
    h->peers.slot_len = ((n + 15) / 16) * 16;
    h->peers.n_blocks = (n + 31) / 32;
    const size_t data_bytes = (size_t)2 * h->world * h->peers.slot_len * 8;
    const size_t flag_bytes = (size_t)2 * h->world * h->peers.n_blocks * 8;
    CUDA_TRY(cudaMalloc((void**)&h->d_xchg, data_bytes + flag_bytes));
    CUDA_TRY(cudaMemset(h->d_xchg, 0, data_bytes + flag_bytes));
    CUDA_TRY(cudaMalloc((void**)&h->d_p2p_err, 4));
    CUDA_TRY(cudaMemset(h->d_p2p_err, 0, 4));
    CUDA_TRY(cudaMalloc((void**)&h->d_xseq, 8));
    CUDA_TRY(cudaMemset(h->d_xseq, 0, 8));
    CUDA_TRY(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t hd;
  CUDA_TRY(cudaIpcGetMemHandle(&hd, h->d_xchg));
  static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &hd, 64);
  return 0;
}

int pinn_p2p_connect(pinn_t* h, const void* handles, int world) {
  if (!h || !handles) return fail("pinn_p2p_connect: null argument");
  if (world != h->world || world < 2 || world > pinn::P2P_MAX) return fail("pinn_p2p_connect: world must match pinn_create (2..8)");
  if (!h->d_xchg) return fail("pinn_p2p_connect: call pinn_p2p_export first");
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t data_doubles = (size_t)2 * world * h->peers.slot_len;
  for (int r = 0; r < world; r++) {
    void* base = nullptr;
    if (r == h->rank) {
      base = h->d_xchg;
    } else {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, (const char*)handles + (size_t)r * 64, 64);
      cudaError_t e = cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) { cudaGetLastError(); return fail(std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e)); }
      h->peer_base[r] = base;
    }
    h->peers.data[r] = (double*)base;
    h->peers.flag[r] = reinterpret_cast<unsigned long long*>((double*)base + data_doubles);
  }
  h->p2p_mapped = true;
  h->p2p_ready = true;
  return 0;
}

int pinn_p2p_enable(pinn_t* h, int on) {
  if (!h) return fail("null handle");
  if (on && !h->p2p_mapped) return fail("pinn_p2p_enable: call pinn_p2p_connect first");
  h->p2p_ready = on != 0;
  return 0;
}

int pinn_nccl_unique_id(void* out128) {
  std::string why;
  if (!g_nccl.load(why)) return fail(why);
  ncclUniqueId id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail("ncclGetUniqueId failed");
  memcpy(out128, &id, 128);
  return 0;
}

int pinn_create(pinn_t** out, int pde_id, int n_layers, const int* layers, const double lb[2], const double ub[2],
                int device, int rank, int world, const void* nccl_uid) {
  if (!out || !layers || !lb || !ub) return fail("pinn_create: null argument");
  if (n_layers < 3 || n_layers > pinn::MAXL) return fail("pinn_create: need 3..16 layer sizes");
  if (pde_id < 0 || pde_id > 4) return fail("pinn_create: unknown pde_id");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("pinn_create: no CUDA device -- this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail("pinn_create: bad device index");
  pinn_t* h = new pinn_t();
  h->pde = pde_id; h->device = device; h->rank = rank; h->world = world < 1 ? 1 : world;
  h->layers.assign(layers, layers + n_layers);
  const bool disc = pde_id == PINN_BURGERS_DISC || pde_id == PINN_BURGERS_IDE_DISC;
  for (int i = 0; i < n_layers; i++) {
    const int v = h->layers[i];
    const int cap = (disc && i == n_layers - 1) ? 512 : pinn::MAXW;     // the IRK head may be q+1 = 501 wide
    if (v < 1 || v > cap) { delete h; return fail("pinn_create: layer width out of range (1..128; IRK head <= 512)"); }
  }
  if (h->layers[0] != (disc ? 1 : 2)) {
    delete h;
    return fail(disc ? "pinn_create: the discrete-time model takes a 1-D input (x)" : "pinn_create: input dimension must be 2 (x,t)");
  }
  h->lb[0] = lb[0]; h->ub[0] = ub[0];
  h->lb[1] = disc ? 0.0 : lb[1]; h->ub[1] = disc ? 1.0 : ub[1];            // the t axis does not exist for 1-D nets
  if (!(h->ub[0] > h->lb[0]) || !(h->ub[1] > h->lb[1])) { delete h; return fail("pinn_create: need ub > lb"); }
  const int want_out = pde_id == PINN_NLS_INF ? 2 : 1;
  if (pde_id == PINN_BURGERS_IDE_DISC) {
    if (h->layers.back() < 1) { delete h; return fail("pinn_create: the discrete-time identification model needs q >= 1 outputs"); }
  } else if (disc) {
    if (h->layers.back() < 2) { delete h; return fail("pinn_create: the discrete-time model needs q+1 >= 2 outputs"); }
  } else if (h->layers.back() != want_out) {
    delete h;
    return fail(pde_id == PINN_NLS_INF ? "pinn_create: the Schrodinger problem needs 2 network outputs (u, v)"
                                        : "pinn_create: the Burgers problems need 1 network output (u)");
  }
  {
    // specialised DMMA kernels for the two BASELINE nets, generic DFMA kernel for any other layer list
    const char* fg = getenv("PINN_FORCE_GENERIC");
    const bool force = fg && fg[0] == '1';
    const bool special = disc ? false : ((pde_id == PINN_NLS_INF) ? is_nls_net(h->layers) : is_burgers_net(h->layers));
    h->kernel_kind = (special && !force) ? (pde_id == PINN_NLS_INF ? 1 : 0) : 2;
  }
  h->P_net = 0;
  for (int l = 0; l + 1 < n_layers; l++) h->P_net += layers[l] * layers[l + 1] + layers[l + 1];
  h->P = h->P_net + ((pde_id == PINN_BURGERS_IDE || pde_id == PINN_BURGERS_IDE_DISC) ? 2 : 0);

#define CREATE_TRY(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      fail(std::string(#expr) + ": " + cudaGetErrorString(_e));                                   \
      pinn_destroy(h);                                                                            \
      return -1;                                                                                  \
    }                                                                                             \
  } while (0)
  CREATE_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CREATE_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    fail(std::string("pinn_create: device '") + prop.name + "' is not sm_100 class; this library is built for sm_100a only");
    pinn_destroy(h);
    return -1;
  }
  h->n_sm = prop.multiProcessorCount;
  CREATE_TRY(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  CREATE_TRY(cudaEventCreate(&h->ev0));
  CREATE_TRY(cudaEventCreate(&h->ev1));
  h->w_cap = ((h->P + 3 + 63) / 64) * 64 + 64;
  if (h->w_cap < pinn::burgers::WPAD) h->w_cap = pinn::burgers::WPAD;
  if (h->w_cap < pinn::nls::WPAD + 64) h->w_cap = pinn::nls::WPAD + 64;
  CREATE_TRY(cudaMalloc((void**)&h->d_w, h->w_cap * 8));
  CREATE_TRY(cudaMemset(h->d_w, 0, h->w_cap * 8));
  CREATE_TRY(cudaMalloc((void**)&h->d_R, h->w_cap * 8));
  CREATE_TRY(cudaMemset(h->d_R, 0, h->w_cap * 8));
  CREATE_TRY(cudaMalloc((void**)&h->d_m, h->w_cap * 8));
  CREATE_TRY(cudaMalloc((void**)&h->d_v, h->w_cap * 8));
  CREATE_TRY(cudaMemset(h->d_m, 0, h->w_cap * 8));
  CREATE_TRY(cudaMemset(h->d_v, 0, h->w_cap * 8));
  CREATE_TRY(cudaMalloc((void**)&h->d_step, 16));
  CREATE_TRY(cudaMemset(h->d_step, 0, 16));
  CREATE_TRY(cudaMalloc((void**)&h->d_loss_ring, LOSS_RING * 8));
  CREATE_TRY(cudaMemset(h->d_loss_ring, 0, LOSS_RING * 8));
  CREATE_TRY(cudaMalloc((void**)&h->d_lb, sizeof(pinn::LbfgsState)));
  CREATE_TRY(cudaMemset(h->d_lb, 0, sizeof(pinn::LbfgsState)));
  if (h->kernel_kind == 2) {
    h->n_cta = 2 * h->n_sm;
    h->pstride = ((h->P_net + 8 + 15) / 16) * 16;
  } else if (pde_id == PINN_NLS_INF) {
    h->n_cta = pinn::nls::grid_size(h->n_sm);
    h->pstride = pinn::nls::PSTRIDE;
    CREATE_TRY(cudaFuncSetAttribute(pinn::nls::fused_loss_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, pinn::nls::SMEM_BYTES));
  } else {
    h->n_cta = h->n_sm;
    h->pstride = pinn::burgers::PSTRIDE;
    CREATE_TRY(cudaFuncSetAttribute(pinn::burgers::fused_loss_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, pinn::burgers::SMEM_BYTES));
    CREATE_TRY(cudaFuncSetAttribute(pinn::burgers2::fused_loss_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, pinn::burgers2::SMEM_BYTES));
    const char* kv = getenv("PINN_BURGERS_KERNEL");
    if (kv && (!strcmp(kv, "v1") || !strcmp(kv, "1"))) h->burgers_kernel = 1;
  }
  CREATE_TRY(cudaMalloc((void**)&h->d_partials, (size_t)h->n_cta * h->pstride * 8));
  CREATE_TRY(cudaMemset(h->d_partials, 0, (size_t)h->n_cta * h->pstride * 8));
  if (h->world > 1) {
    std::string why;
    if (!nccl_uid) { fail("pinn_create: world > 1 needs an ncclUniqueId"); pinn_destroy(h); return -1; }
    if (!g_nccl.load(why)) { fail(why); pinn_destroy(h); return -1; }
    ncclUniqueId id;
    memcpy(&id, nccl_uid, 128);
    int rc = g_nccl.CommInitRank(&h->comm, h->world, id, h->rank);
    if (rc != 0) { fail(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error")); pinn_destroy(h); return -1; }
  }
  { std::lock_guard<std::mutex> lk(g_live_mu); g_live.push_back(h); }
  *out = h;
  return 0;
}

int pinn_destroy(pinn_t* h) {
  if (!h) return 0;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    for (size_t i = 0; i < g_live.size(); i++)
      if (g_live[i] == h) { g_live.erase(g_live.begin() + i); break; }
  }
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int r = 0; r < pinn::P2P_MAX; r++) if (h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
  if (h->d_xchg) cudaFree(h->d_xchg);
  if (h->d_p2p_err) cudaFree(h->d_p2p_err);
  if (h->d_xseq) cudaFree(h->d_xseq);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  double* bufs[] = {h->d_w, h->d_R, h->d_partials, h->d_m, h->d_v, h->d_loss_ring, h->d_x, h->d_t, h->d_u, h->d_scrH, h->d_scrA, h->d_scrS, h->d_gH, h->d_gA, h->d_gS, h->d_irk, h->d_gold,
                    h->d_d, h->d_S, h->d_Y, h->d_xfinal, h->d_fhist, h->d_px, h->d_pout};
  for (double* b : bufs) if (b) cudaFree(b);
  if (h->d_step) cudaFree(h->d_step);
  if (h->d_lb) cudaFree(h->d_lb);
  if (h->d_logged) cudaFree(h->d_logged);
  lbfgs_gram_free(&h->lb_gram);
  for (cudaEvent_t e : h->events) cudaEventDestroy(e);
  if (h->d_flush) cudaFree(h->d_flush);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int64_t pinn_num_params(const pinn_t* h) { return h ? h->P : -1; }

int pinn_set_pde_params(pinn_t* h, const double* p, int n) {
  if (!h) return fail("null handle");
  if (h->pde == PINN_BURGERS_INF) {
    if (n != 1 || !p) return fail("pinn_set_pde_params: BURGERS_INF takes exactly one parameter (nu)");
    h->nu = p[0];
    return 0;
  }
  if (h->pde == PINN_BURGERS_DISC) {
    if (n != 2 || !p) return fail("pinn_set_pde_params: BURGERS_DISC takes [nu, dt]");
    h->nu = p[0]; h->dt = p[1];
    return 0;
  }
  if (h->pde == PINN_BURGERS_IDE_DISC) {
    if (n != 1 || !p) return fail("pinn_set_pde_params: BURGERS_IDE_DISC takes [dt]");
    h->dt = p[0];
    return 0;
  }
  if (n != 0) return fail("pinn_set_pde_params: this PDE takes no constants");
  return 0;
}

int pinn_get_params(pinn_t* h, double* p, int n) {
  if (!h || !p) return fail("pinn_get_params: null argument");
  if (h->pde == PINN_BURGERS_INF) {
    if (n != 1) return fail("pinn_get_params: BURGERS_INF has one parameter (nu)");
    p[0] = h->nu;                                   // inf_cont_burgers.py:92-93
    return 0;
  }
  if (h->pde == PINN_BURGERS_IDE || h->pde == PINN_BURGERS_IDE_DISC) {
    if (n != 2) return fail("pinn_get_params: identification has two parameters (lambda_1, exp(lambda_2))");
    CUDA_TRY(cudaSetDevice(h->device));
    double lam[2];
    CUDA_TRY(cudaMemcpyAsync(lam, h->d_w + h->P_net, 16, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    p[0] = lam[0]; p[1] = std::exp(lam[1]);         // ide_cont_burgers.py:109-114, ide_disc_burgers.py:138-143
    return 0;
  }
  if (h->pde == PINN_BURGERS_DISC) {
    if (n != 2) return fail("pinn_get_params: BURGERS_DISC has two constants (nu, dt)");
    p[0] = h->nu; p[1] = h->dt;
    return 0;
  }
  if (n != 0) return fail("pinn_get_params: this PDE has no parameters");
  return 0;
}

int pinn_set_irk(pinn_t* h, const double* irk, int q) {
  if (!h || !irk) return fail("pinn_set_irk: null argument");
  if (h->pde != PINN_BURGERS_DISC && h->pde != PINN_BURGERS_IDE_DISC) return fail("pinn_set_irk: only the discrete-time models have stage matrices");
  const bool idd = h->pde == PINN_BURGERS_IDE_DISC;
  if (!idd && q + 1 != h->layers.back()) return fail("pinn_set_irk: q+1 must equal the network's output width");
  if (idd && q != h->layers.back()) return fail("pinn_set_irk: q must equal the network's output width");
  const size_t n_rows = idd ? (size_t)2 * q : (size_t)q + 1;      // identification: [M_0 ; M_1], each q x q
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->d_irk) cudaFree(h->d_irk);
  h->d_irk = nullptr;
  CUDA_TRY(cudaMalloc((void**)&h->d_irk, n_rows * q * 8));
  CUDA_TRY(cudaMemcpy(h->d_irk, irk, n_rows * q * 8, cudaMemcpyHostToDevice));
  h->irk_q = q;
  return 0;
}

int pinn_set_collocation(pinn_t* h, const double* x, const double* t, int64_t n, int64_t n_global) {
  if (!h) return fail("null handle");
  if (h->pde == PINN_BURGERS_IDE) return fail("pinn_set_collocation: identification uses the data points as residual points");
  if (h->pde == PINN_BURGERS_DISC || h->pde == PINN_BURGERS_IDE_DISC)
    return fail("pinn_set_collocation: the discrete-time models have no collocation set (snapshot / boundary points only)");
  if (n < 0 || (n > 0 && (!x || !t))) return fail("pinn_set_collocation: bad arguments");
  if (n_global < n) return fail("pinn_set_collocation: n_global < n");
  CUDA_TRY(cudaSetDevice(h->device));
  if (ensure_points(h, h->n_aux, n)) return -1;
  if (n) {
    // straight from the caller's buffer into the collocation region (truly asynchronous when it is pinned)
    CUDA_TRY(cudaMemcpyAsync(h->d_x + h->dcap, x, n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_t + h->dcap, t, n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // the buffers are only borrowed for the call
  }
  h->n_c = n; h->n_c_global = n_global;
  h->map_x = h->map_t = nullptr;
  return 0;
}

int pinn_set_collocation_mapped(pinn_t* h, const double* x_pinned, const double* t_pinned, int64_t n, int64_t n_global) {
  if (!h) return fail("null handle");
  if (h->kernel_kind != 0 || h->pde != PINN_BURGERS_INF)
    return fail("pinn_set_collocation_mapped: zero-copy collocation is implemented for the fused Burgers inference kernel only");
  if (n <= 0 || !x_pinned || !t_pinned || n_global < n) return fail("pinn_set_collocation_mapped: bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  void *dx = nullptr, *dt = nullptr;
  if (cudaHostGetDevicePointer(&dx, (void*)x_pinned, 0) != cudaSuccess || cudaHostGetDevicePointer(&dt, (void*)t_pinned, 0) != cudaSuccess) {
    cudaGetLastError();
    return fail("pinn_set_collocation_mapped: the buffers must be pinned host memory (pinn_host_alloc / cudaHostAlloc)");
  }
  if (ensure_points(h, h->n_aux, 0)) return -1;
  // the kernel that is about to be enqueued reads the caller's buffers: they must stay valid and unchanged until the
  // next synchronising call (pinn_adam_step with a loss pointer, pinn_loss_grad, pinn_sync)
  h->map_x = (const double*)dx; h->map_t = (const double*)dt;
  h->n_c = n; h->n_c_global = n_global;
  return 0;
}

int pinn_set_data(pinn_t* h, const double* X, int64_t n, int in_dim, const double* u, int out_dim, double weight) {
  if (!h) return fail("null handle");
  if (n < 0 || (n > 0 && (!X || !u))) return fail("pinn_set_data: bad arguments");
  if (in_dim != 1 && in_dim != 2) return fail("pinn_set_data: in_dim must be 1 (broadcast quirk) or 2");
  if (h->pde == PINN_BURGERS_IDE_DISC) return pinn_set_snapshot(h, 0, X, n, u);
  if (h->pde == PINN_BURGERS_DISC) {
    // x_0 (n,1) and u_0 (n,1): the snapshot the q stages are fitted to (inf_disc_burgers.py:98-101; u_0 broadcasts)
    if (in_dim != 1 || out_dim != 1) return fail("pinn_set_data: the discrete-time model takes x_0 (n,1) and u_0 (n,1)");
    CUDA_TRY(cudaSetDevice(h->device));
    h->h_x0.assign(X, X + n);
    if (n) {
      if (ensure(&h->d_u, &h->u_cap, n)) return -1;
      CUDA_TRY(cudaMemcpyAsync(h->d_u, u, n * 8, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    h->n_d = n; h->d_out_dim = 1; h->data_weight = weight;
    return disc_upload_points(h);
  }
  if (out_dim != h->layers.back()) return fail("pinn_set_data: out_dim does not match the network head");
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->pde == PINN_NLS_INF) {
    h->h_icx.resize(n); h->h_ict.resize(n);
    for (int64_t i = 0; i < n; i++) {
      h->h_icx[i] = X[i * in_dim];
      h->h_ict[i] = in_dim == 1 ? X[i] : X[i * in_dim + 1];   // quirk Q1: (N,1) input broadcast, t := x
    }
    if (n) {
      if (ensure(&h->d_u, &h->u_cap, n * out_dim)) return -1;
      CUDA_TRY(cudaMemcpyAsync(h->d_u, u, n * out_dim * 8, cudaMemcpyHostToDevice, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    h->n_d = n; h->d_out_dim = out_dim; h->data_weight = weight;
    return nls_upload_points(h);
  }
  // shrinking/growing the data set moves its right-aligned start; the collocation region is untouched
  if (ensure_points(h, n, h->n_c)) return -1;
  h->stage_x.resize(n); h->stage_t.resize(n);
  for (int64_t i = 0; i < n; i++) {
    h->stage_x[i] = X[i * in_dim];
    h->stage_t[i] = in_dim == 1 ? X[i] : X[i * in_dim + 1];   // (N,1) input broadcast by the Lambda: t := x
  }
  if (n) {
    if (ensure(&h->d_u, &h->u_cap, n * out_dim)) return -1;
    CUDA_TRY(cudaMemcpyAsync(h->d_x + h->dcap - n, h->stage_x.data(), n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_t + h->dcap - n, h->stage_t.data(), n * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_u, u, n * out_dim * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  h->n_d = n; h->n_aux = n; h->d_out_dim = out_dim; h->data_weight = weight;
  return 0;
}

int pinn_set_snapshot(pinn_t* h, int which, const double* x, int64_t n, const double* u) {
  if (!h) return fail("null handle");
  if (h->pde != PINN_BURGERS_IDE_DISC) return fail("pinn_set_snapshot: only the discrete-time identification model has two snapshots");
  if ((which != 0 && which != 1) || n < 0 || (n > 0 && (!x || !u))) return fail("pinn_set_snapshot: bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  // host copies of both snapshots: the device blocks [x_0 | x_1] and [u_0 | u_1] are re-assembled whenever one changes
  (which == 0 ? h->h_x0 : h->h_tb).assign(x, x + n);
  (which == 0 ? h->h_u0 : h->h_u1).assign(u, u + n);
  h->n_d = (long long)h->h_x0.size();
  h->n_b = (long long)h->h_tb.size();
  h->d_out_dim = 1; h->data_weight = 1.0;
  const long long nu = h->n_d + h->n_b;
  if (nu) {
    if (ensure(&h->d_u, &h->u_cap, nu)) return -1;
    if (h->n_d) CUDA_TRY(cudaMemcpyAsync(h->d_u, h->h_u0.data(), h->n_d * 8, cudaMemcpyHostToDevice, h->stream));
    if (h->n_b) CUDA_TRY(cudaMemcpyAsync(h->d_u + h->n_d, h->h_u1.data(), h->n_b * 8, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return disc_upload_points(h);
}

int pinn_set_boundary(pinn_t* h, const double* tb, int64_t n_b) {
  if (!h) return fail("null handle");
  if (h->pde == PINN_BURGERS_DISC) {
    if (n_b < 0 || (n_b > 0 && !tb)) return fail("pinn_set_boundary: bad arguments");
    h->h_tb.assign(tb, tb + n_b);
    h->n_b = n_b;
    CUDA_TRY(cudaSetDevice(h->device));
    return disc_upload_points(h);
  }
  if (h->pde != PINN_NLS_INF) return fail("pinn_set_boundary: only the NLS and discrete-time inference problems have a boundary term");
  if (n_b < 0 || (n_b > 0 && !tb)) return fail("pinn_set_boundary: bad arguments");
  h->h_tb.assign(tb, tb + n_b);
  h->n_b = n_b;
  CUDA_TRY(cudaSetDevice(h->device));
  return nls_upload_points(h);
}

int pinn_set_weights(pinn_t* h, const double* w, int64_t n) {
  if (!h || !w) return fail("pinn_set_weights: null argument");
  if (n != h->P) return fail("pinn_set_weights: expected " + std::to_string(h->P) + " values, got " + std::to_string(n));
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(h->d_w, w, n * 8, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int pinn_get_weights(pinn_t* h, double* w, int64_t n) {
  if (!h || !w) return fail("pinn_get_weights: null argument");
  if (n != h->P) return fail("pinn_get_weights: expected " + std::to_string(h->P) + " values, got " + std::to_string(n));
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(w, h->d_w, n * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int pinn_loss_grad(pinn_t* h, const double* w_or_null, double* loss_out, double* grad_out_or_null,
                   double* parts_out_or_null) {
  if (!h) return fail("null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (w_or_null) CUDA_TRY(cudaMemcpyAsync(h->d_w, w_or_null, h->P * 8, cudaMemcpyHostToDevice, h->stream));
  if (launch_eval(h, nullptr)) return -1;
  double parts[3];
  CUDA_TRY(cudaMemcpyAsync(parts, h->d_R + h->P, 24, cudaMemcpyDeviceToHost, h->stream));
  if (grad_out_or_null) CUDA_TRY(cudaMemcpyAsync(grad_out_or_null, h->d_R, h->P * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (check_p2p(h)) return -1;
  if (loss_out) *loss_out = parts[0] + parts[1] + parts[2];
  if (parts_out_or_null) { parts_out_or_null[0] = parts[0]; parts_out_or_null[1] = parts[1]; parts_out_or_null[2] = parts[2]; }
  return 0;
}

int pinn_adam_step(pinn_t* h, double lr, double b1, double b2, double eps, double* loss_out_or_null) {
  if (!h) return fail("null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  // fused kernel, then the tail: ONE more kernel on a single GPU (reduce + Adam) and with the NVLink push exchange
  // (reduce + exchange + Adam); reduce + ncclAllReduce + Adam on the NCCL path
  const AdamArgs ad{lr, b1, b2, eps};
  if (launch_eval(h, nullptr, false, &ad)) return -1;
  h->adam_steps++;
  if (loss_out_or_null) {
    CUDA_TRY(cudaMemcpyAsync(loss_out_or_null, h->d_loss_ring + ((h->adam_steps - 1) % LOSS_RING), 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (check_p2p(h)) return -1;
  }
  return 0;
}

int pinn_adam_steps(pinn_t* h, int n, double lr, double b1, double b2, double eps) {
  if (!h) return fail("null handle");
  if (n < 0) return fail("pinn_adam_steps: n < 0");
  CUDA_TRY(cudaSetDevice(h->device));
  const AdamArgs ad{lr, b1, b2, eps};
  for (int i = 0; i < n; i++) {
    if (launch_eval(h, nullptr, false, &ad)) return -1;
    h->adam_steps++;
  }
  return 0;
}

int pinn_adam_reset(pinn_t* h) {
  if (!h) return fail("null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_m, 0, h->w_cap * 8, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->d_v, 0, h->w_cap * 8, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->d_step, 0, 16, h->stream));
  h->adam_steps = 0;
  return 0;
}

int pinn_last_loss(pinn_t* h, double* loss_out) {
  if (!h || !loss_out) return fail("null argument");
  if (h->adam_steps == 0) return fail("pinn_last_loss: no Adam step has run");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemcpyAsync(loss_out, h->d_loss_ring + ((h->adam_steps - 1) % LOSS_RING), 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (check_p2p(h)) return -1;
  return 0;
}

int pinn_lbfgs(pinn_t* h, int max_iter, double learning_rate, int n_correction, double tol_fun, double tol_x,
               int sync_every, pinn_log_cb log_cb, void* user, int* n_iter_out, int* n_eval_out, int* reason_out,
               double* x_final_or_null) {
  if (!h) return fail("null handle");
  if (n_iter_out) *n_iter_out = 0;
  if (n_eval_out) *n_eval_out = 0;
  if (reason_out) *reason_out = PINN_LBFGS_RUNNING;
  if (max_iter == 0) return 0;                                      // custom_lbfgs.py:43-44
  if (max_iter < 0) return fail("pinn_lbfgs: max_iter < 0");
  if (n_correction <= 0) n_correction = 100;                        // :52 (`or 100`)
  if (n_correction > 128) return fail("pinn_lbfgs: n_correction > 128 not supported");
  if (learning_rate == 0.0) learning_rate = 1.0;                    // :55
  if (tol_fun == 0.0) tol_fun = 1e-5;                               // :50
  if (tol_x == 0.0) tol_x = 1e-19;                                  // :51
  if (sync_every < 1) sync_every = 1;
  const int P = h->P;
  if (lbfgs_serial() && P > 32 * 1024) return fail("pinn_lbfgs: parameter vector too large for the single-CTA L-BFGS kernel");
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_gold) {
    CUDA_TRY(cudaMalloc((void**)&h->d_gold, h->w_cap * 8));
    CUDA_TRY(cudaMalloc((void**)&h->d_d, h->w_cap * 8));
    CUDA_TRY(cudaMalloc((void**)&h->d_xfinal, h->w_cap * 8));
  }
  if (n_correction > h->lb_corr_cap) {
    if (h->d_S) cudaFree(h->d_S);
    if (h->d_Y) cudaFree(h->d_Y);
    h->d_S = h->d_Y = nullptr;
    CUDA_TRY(cudaMalloc((void**)&h->d_S, (size_t)(n_correction + 1) * P * 8));     // + the spare slot of the Gram formulation
    CUDA_TRY(cudaMalloc((void**)&h->d_Y, (size_t)(n_correction + 1) * P * 8));
    h->lb_corr_cap = n_correction;
  }
  if (lbfgs_gram_ensure(&h->lb_gram, n_correction, P)) return -1;
  if (max_iter + 2 > h->lb_iter_cap) {
    if (h->d_fhist) cudaFree(h->d_fhist);
    if (h->d_logged) cudaFree(h->d_logged);
    h->d_fhist = nullptr; h->d_logged = nullptr;
    CUDA_TRY(cudaMalloc((void**)&h->d_fhist, (size_t)(max_iter + 2) * 8));
    CUDA_TRY(cudaMalloc((void**)&h->d_logged, (size_t)(max_iter + 2) * 4));
    h->lb_iter_cap = max_iter + 2;
  }
  CUDA_TRY(cudaMemsetAsync(h->d_logged, 0, (size_t)(max_iter + 2) * 4, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->d_fhist, 0, (size_t)(max_iter + 2) * 8, h->stream));
  pinn::LbfgsState st{};
  st.status = 0; st.n_iter = 0; st.n_eval = 0; st.k = 0; st.head = 0; st.pending = 1;
  st.max_iter = max_iter; st.n_corr = n_correction;
  st.max_eval = max_iter * 1.25;                                    // :49
  st.lr = learning_rate; st.tol_fun = tol_fun; st.tol_x = tol_x;
  st.h_diag = 1.0; st.t = 0.0; st.f = 0.0; st.f_old = 0.0;
  CUDA_TRY(cudaMemcpyAsync(h->d_lb, &st, sizeof(st), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));   // st is a stack object
  const int* run_flag = &h->d_lb->status;
  if (launch_eval(h, run_flag)) return -1;      // initial f, g (:65)

  std::vector<double> fh(max_iter + 2);
  std::vector<int> lg(max_iter + 2);
  int reported = 0;
  auto iterate = [&]() -> int {
    return lbfgs_launch_iteration(h->stream, h->d_lb, h->d_w, h->d_R, P, h->d_gold, h->d_d, h->d_S, h->d_Y, h->d_xfinal, h->d_fhist,
                                  h->d_logged, h->lb_gram, &h->launches);
  };
  while (true) {
    for (int b = 0; b < sync_every; b++) {
      if (iterate()) return -1;
      if (launch_eval(h, run_flag)) return -1;
    }
    CUDA_TRY(cudaMemcpyAsync(&st, h->d_lb, sizeof(st), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(fh.data(), h->d_fhist, (size_t)(max_iter + 2) * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(lg.data(), h->d_logged, (size_t)(max_iter + 2) * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (check_p2p(h)) return -1;
    // iterations whose stop tests have run: all < n_iter, plus n_iter itself unless its evaluation is pending
    const int upto = st.pending ? st.n_iter - 1 : st.n_iter;
    for (int it = reported + 1; it <= upto; it++)
      if (lg[it] && log_cb) log_cb(it, fh[it], user);
    if (upto > reported) reported = upto;
    if (st.status != 0) break;
  }
  h->lb_fhist.assign(fh.begin(), fh.begin() + (st.n_eval < (int)fh.size() ? st.n_eval : (int)fh.size()));
  if (n_iter_out) *n_iter_out = st.n_iter;
  if (n_eval_out) *n_eval_out = st.n_eval;
  if (reason_out) *reason_out = st.status;
  if (x_final_or_null) {
    // x after the loop: differs from the model weights only when the loop ended on max_iter (the last update
    // is not evaluated, utils/custom_lbfgs.py:176-182; utils/neuralnetwork.py:131-136 ignores the return).
    const double* src = (st.status == PINN_LBFGS_MAX_ITER) ? h->d_xfinal : h->d_w;
    CUDA_TRY(cudaMemcpyAsync(x_final_or_null, src, (size_t)P * 8, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  return 0;
}

int pinn_lbfgs_history(pinn_t* h, double* f_hist_out, int capacity, int* n_out) {
  if (!h || !n_out) return fail("pinn_lbfgs_history: null argument");
  const int n = (int)h->lb_fhist.size();
  *n_out = n;
  if (f_hist_out) {
    if (capacity < n) return fail("pinn_lbfgs_history: buffer too small");
    for (int i = 0; i < n; i++) f_hist_out[i] = h->lb_fhist[i];
  }
  return 0;
}

// generic thread-per-point forward on host points (X != nullptr) or on a device SoA range (dx, dt)
static int forward_generic(pinn_t* h, const double* X, const double* dx, const double* dt, int64_t n, int in_dim, double* out,
                           int ns) {
  if (!h || !out || (!X && !dx)) return fail("null argument");
  if (n <= 0) return 0;
  if (in_dim != 1 && in_dim != 2) return fail("in_dim must be 1 or 2");
  CUDA_TRY(cudaSetDevice(h->device));
  const int no = h->layers.back();
  if (ensure(&h->d_pout, &h->pout_cap, n * no * ns)) return -1;
  if (X) {
    if (ensure(&h->d_px, &h->px_cap, n * in_dim)) return -1;
    CUDA_TRY(cudaMemcpyAsync(h->d_px, X, n * in_dim * 8, cudaMemcpyHostToDevice, h->stream));
    dx = h->d_px; dt = nullptr;
  }
  pinn::NetDesc nd = net_desc(h);
  const int threads = 64;
  const int blocks = (int)((n + threads - 1) / threads);
  if (ns == 1) pinn::mlp_forward_generic<1><<<blocks, threads, 0, h->stream>>>(h->d_w, nd, dx, dt, n, in_dim, h->d_pout);
  else pinn::mlp_forward_generic<4><<<blocks, threads, 0, h->stream>>>(h->d_w, nd, dx, dt, n, in_dim, h->d_pout);
  CUDA_TRY(cudaGetLastError());
  h->launches++;
  CUDA_TRY(cudaMemcpyAsync(out, h->d_pout, n * no * ns * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int pinn_predict(pinn_t* h, const double* X, int64_t n, int in_dim, double* out) {
  return forward_generic(h, X, nullptr, nullptr, n, in_dim, out, 1);
}

int pinn_derivatives(pinn_t* h, const double* X, int64_t n, double* out) {
  return forward_generic(h, X, nullptr, nullptr, n, h ? h->layers[0] : 2, out, 4);
}

int64_t pinn_num_residual_points(const pinn_t* h) {
  if (!h) return -1;
  if (h->pde == PINN_BURGERS_DISC || h->pde == PINN_BURGERS_IDE_DISC) return 0;
  return h->pde == PINN_BURGERS_IDE ? h->n_d : h->n_c;
}

int pinn_residual(pinn_t* h, double* f_out, int64_t n_rows) {
  if (!h || !f_out) return fail("null argument");
  if (h->pde == PINN_BURGERS_DISC || h->pde == PINN_BURGERS_IDE_DISC) return fail("pinn_residual: not defined for the discrete-time models");
  const bool ide = h->pde == PINN_BURGERS_IDE;
  const int64_t n = ide ? h->n_d : h->n_c;
  if (n_rows != n)
    return fail("pinn_residual: the output buffer holds " + std::to_string(n_rows) + " rows but " + std::to_string(n) +
                " residual points are stored (pinn_num_residual_points)");
  if (n == 0) return 0;
  const long long off = ide ? h->dcap - h->n_d : h->dcap;
  std::vector<double> D(n * 4 * h->layers.back());
  const double* rx = (!ide && h->map_x) ? h->map_x : h->d_x + off;
  const double* rt = (!ide && h->map_t) ? h->map_t : h->d_t + off;
  if (forward_generic(h, nullptr, rx, rt, n, 2, D.data(), 4)) return -1;
  if (h->pde == PINN_NLS_INF) {
    for (int64_t i = 0; i < n; i++) {
      const double* d = &D[i * 8];
      const double u = d[0], v = d[1], ut = d[4], vt = d[5], uxx = d[6], vxx = d[7];
      const double h2 = u * u + v * v;
      f_out[2 * i] = ut + 0.5 * vxx + h2 * v;
      f_out[2 * i + 1] = vt - 0.5 * uxx - h2 * u;
    }
    return 0;
  }
  double l1 = 1.0, kap = h->nu;
  if (ide) {
    double lam[2];
    CUDA_TRY(cudaMemcpy(lam, h->d_w + h->P_net, 16, cudaMemcpyDeviceToHost));
    l1 = lam[0]; kap = std::exp(lam[1]);
  }
  for (int64_t i = 0; i < n; i++) {
    const double* d = &D[i * 4];
    f_out[i] = d[2] + l1 * d[0] * d[1] - kap * d[3];
  }
  return 0;
}

int pinn_sync(pinn_t* h) {
  if (!h) return fail("null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return check_p2p(h);
}

int pinn_host_alloc(void** out, int64_t bytes) {
  if (!out || bytes <= 0) return fail("pinn_host_alloc: bad arguments");
  CUDA_TRY(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
  return 0;
}
int pinn_host_free(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return 0;
}

int pinn_time_loss_grad_kernel(pinn_t* h, int iters, float* ms_total_out) {
  if (!h || !ms_total_out || iters < 1) return fail("pinn_time_loss_grad_kernel: bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  const int world = h->world;
  h->world = 1;   // kernel only: no collective inside the timed region
  CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
  int rc = 0;
  for (int i = 0; i < iters && rc == 0; i++) rc = launch_eval(h, nullptr, true);
  h->world = world;
  if (rc) return -1;
  CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  CUDA_TRY(cudaEventSynchronize(h->ev1));
  CUDA_TRY(cudaEventElapsedTime(ms_total_out, h->ev0, h->ev1));
  return 0;
}

int pinn_event_record(pinn_t* h, int idx) {
  if (!h || idx < 0 || idx >= 65536) return fail("pinn_event_record: bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  while ((int)h->events.size() <= idx) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreate(&e));
    h->events.push_back(e);
  }
  CUDA_TRY(cudaEventRecord(h->events[idx], h->stream));
  return 0;
}

int pinn_event_elapsed_ms(pinn_t* h, int i, int j, float* ms_out) {
  if (!h || !ms_out || i < 0 || j < 0 || i >= (int)h->events.size() || j >= (int)h->events.size())
    return fail("pinn_event_elapsed_ms: bad arguments");
  CUDA_TRY(cudaEventSynchronize(h->events[j]));
  CUDA_TRY(cudaEventElapsedTime(ms_out, h->events[i], h->events[j]));
  return 0;
}

int pinn_flush_l2(pinn_t* h) {
  if (!h) return fail("null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_flush) {
    h->flush_bytes = (size_t)256 << 20;   // > 126 MB L2
    CUDA_TRY(cudaMalloc((void**)&h->d_flush, h->flush_bytes));
  }
  CUDA_TRY(cudaMemsetAsync(h->d_flush, 0, h->flush_bytes, h->stream));
  return 0;
}

int pinn_test_tanh(const double* x, int n, double* y) {
  if (!x || !y || n < 1) return fail("pinn_test_tanh: bad arguments");
  double *dx = nullptr, *dy = nullptr;
  CUDA_TRY(cudaMalloc((void**)&dx, (size_t)n * 8));
  CUDA_TRY(cudaMalloc((void**)&dy, (size_t)n * 8));
  CUDA_TRY(cudaMemcpy(dx, x, (size_t)n * 8, cudaMemcpyHostToDevice));
  pinn::tanh_fast_kernel<<<(n + 255) / 256, 256>>>(dx, n, dy);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(y, dy, (size_t)n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dy);
  return 0;
}

int64_t pinn_launch_count(const pinn_t* h) { return h ? h->launches : -1; }

int pinn_kernel_info(pinn_t* h, char* buf, int buflen) {
  if (!h || !buf || buflen < 1) return fail("bad arguments");
  cudaFuncAttributes fa{};
  int smem = 0, threads = 0;
  if (h->kernel_kind == 2) {
    CUDA_TRY(cudaFuncGetAttributes(&fa, pinn::generic::fused_loss_grad));
    smem = 0; threads = pinn::generic::THREADS;
  } else if (h->pde == PINN_NLS_INF) {
    CUDA_TRY(cudaFuncGetAttributes(&fa, pinn::nls::fused_loss_grad));
    smem = pinn::nls::SMEM_BYTES; threads = pinn::nls::THREADS;
  } else {
    if (h->burgers_kernel == 2) {
      CUDA_TRY(cudaFuncGetAttributes(&fa, pinn::burgers2::fused_loss_grad));
      smem = pinn::burgers2::SMEM_BYTES; threads = pinn::burgers2::THREADS;
    } else {
      CUDA_TRY(cudaFuncGetAttributes(&fa, pinn::burgers::fused_loss_grad));
      smem = pinn::burgers::SMEM_BYTES; threads = pinn::burgers::THREADS;
    }
  }
  snprintf(buf, buflen, "{\"grid\": %d, \"block\": %d, \"dyn_smem\": %d, \"regs\": %d, \"local_bytes\": %zu, \"sms\": %d}",
           h->n_cta, threads, smem, fa.numRegs, fa.localSizeBytes, h->n_sm);
  return 0;
}

// -----------------------------------------------------------------------------------------------------------------
// Stand-alone device-resident L-BFGS for ANY objective (utils/custom_lbfgs.py:39 takes an arbitrary `opfunc`): the caller
// evaluates f and g wherever it likes and feeds them in; the vectors, the (s, y) history ring and the two-loop recursion
// stay on the device and run through the SAME kernel (lbfgs_iterate) the PINN handles use.
// -----------------------------------------------------------------------------------------------------------------
struct pinn_lbfgs_handle {
  int device = 0, n = 0, max_iter = 0, n_corr = 0;
  cudaStream_t stream = nullptr;
  pinn::LbfgsState* d_st = nullptr;
  double *d_x = nullptr, *d_R = nullptr, *d_gold = nullptr, *d_d = nullptr, *d_S = nullptr, *d_Y = nullptr, *d_xfinal = nullptr,
         *d_fhist = nullptr;
  int* d_logged = nullptr;
  LbfgsGram gram;
  int reported = 0;
};

int pinn_lbfgs_destroy(pinn_lbfgs_t* s) {
  if (!s) return 0;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  double* bufs[] = {s->d_x, s->d_R, s->d_gold, s->d_d, s->d_S, s->d_Y, s->d_xfinal, s->d_fhist};
  for (double* b : bufs) if (b) cudaFree(b);
  if (s->d_logged) cudaFree(s->d_logged);
  lbfgs_gram_free(&s->gram);
  if (s->d_st) cudaFree(s->d_st);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return 0;
}

int pinn_lbfgs_create(pinn_lbfgs_t** out, int device, int64_t n, const double* x0, int max_iter, double learning_rate,
                      int n_correction, double tol_fun, double tol_x, double max_eval) {
  if (!out || !x0 || n < 1) return fail("pinn_lbfgs_create: bad arguments");
  if (max_iter < 1) return fail("pinn_lbfgs_create: max_iter must be >= 1 (lbfgs() returns before anything for maxIter == 0)");
  if (lbfgs_serial() && n > 32 * 1024) return fail("pinn_lbfgs_create: vector too large for the single-CTA L-BFGS kernel (32768 entries)");
  if (n > (1 << 30)) return fail("pinn_lbfgs_create: vector too large");
  if (n_correction <= 0) n_correction = 100;                        // custom_lbfgs.py:52
  if (n_correction > 128) return fail("pinn_lbfgs_create: n_correction > 128 not supported");
  if (learning_rate == 0.0) learning_rate = 1.0;                    // :55
  if (tol_fun == 0.0) tol_fun = 1e-5;                               // :50
  if (tol_x == 0.0) tol_x = 1e-19;                                  // :51
  if (max_eval == 0.0) max_eval = max_iter * 1.25;                  // :49
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("pinn_lbfgs_create: no CUDA device -- this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail("pinn_lbfgs_create: bad device index");
  pinn_lbfgs_t* s = new pinn_lbfgs_t();
  s->device = device; s->n = (int)n; s->max_iter = max_iter; s->n_corr = n_correction;
#define LB_TRY(expr)                                                                                     \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess) { fail(std::string(#expr) + ": " + cudaGetErrorString(_e)); pinn_lbfgs_destroy(s); return -1; } \
  } while (0)
  LB_TRY(cudaSetDevice(device));
  LB_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  const size_t nb = ((size_t)n + 8) * 8;
  LB_TRY(cudaMalloc((void**)&s->d_x, nb));
  LB_TRY(cudaMalloc((void**)&s->d_R, nb));
  LB_TRY(cudaMalloc((void**)&s->d_gold, nb));
  LB_TRY(cudaMalloc((void**)&s->d_d, nb));
  LB_TRY(cudaMalloc((void**)&s->d_xfinal, nb));
  LB_TRY(cudaMalloc((void**)&s->d_S, (size_t)(n_correction + 1) * n * 8));
  LB_TRY(cudaMalloc((void**)&s->d_Y, (size_t)(n_correction + 1) * n * 8));
  if (lbfgs_gram_ensure(&s->gram, n_correction, (int)n)) { pinn_lbfgs_destroy(s); return -1; }
  LB_TRY(cudaMalloc((void**)&s->d_fhist, (size_t)(max_iter + 2) * 8));
  LB_TRY(cudaMalloc((void**)&s->d_logged, (size_t)(max_iter + 2) * 4));
  LB_TRY(cudaMalloc((void**)&s->d_st, sizeof(pinn::LbfgsState)));
  LB_TRY(cudaMemsetAsync(s->d_fhist, 0, (size_t)(max_iter + 2) * 8, s->stream));
  LB_TRY(cudaMemsetAsync(s->d_logged, 0, (size_t)(max_iter + 2) * 4, s->stream));
  LB_TRY(cudaMemsetAsync(s->d_R, 0, nb, s->stream));
  LB_TRY(cudaMemcpyAsync(s->d_x, x0, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
  pinn::LbfgsState st{};
  st.pending = 1; st.max_iter = max_iter; st.n_corr = n_correction; st.max_eval = max_eval;
  st.lr = learning_rate; st.tol_fun = tol_fun; st.tol_x = tol_x; st.h_diag = 1.0;
  LB_TRY(cudaMemcpyAsync(s->d_st, &st, sizeof(st), cudaMemcpyHostToDevice, s->stream));
  LB_TRY(cudaStreamSynchronize(s->stream));
#undef LB_TRY
  *out = s;
  return 0;
}

int pinn_lbfgs_feed(pinn_lbfgs_t* s, double f, const double* g, double* x_next, int* status_out, int* n_iter_out, int* n_eval_out,
                    int* logged_iter_out, double* logged_f_out) {
  if (!s || !g || !x_next || !status_out) return fail("pinn_lbfgs_feed: null argument");
  CUDA_TRY(cudaSetDevice(s->device));
  const int n = s->n;
  const double tail[3] = {f, 0.0, 0.0};          // the kernel reads f as the sum of three loss parts behind the gradient
  CUDA_TRY(cudaMemcpyAsync(s->d_R, g, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
  CUDA_TRY(cudaMemcpyAsync(s->d_R + n, tail, 24, cudaMemcpyHostToDevice, s->stream));
  if (lbfgs_launch_iteration(s->stream, s->d_st, s->d_x, s->d_R, n, s->d_gold, s->d_d, s->d_S, s->d_Y, s->d_xfinal, s->d_fhist,
                             s->d_logged, s->gram, nullptr)) return -1;
  pinn::LbfgsState st{};
  CUDA_TRY(cudaMemcpyAsync(&st, s->d_st, sizeof(st), cudaMemcpyDeviceToHost, s->stream));
  CUDA_TRY(cudaStreamSynchronize(s->stream));
  *status_out = st.status;
  if (n_iter_out) *n_iter_out = st.n_iter;
  if (n_eval_out) *n_eval_out = st.n_eval;
  // the iteration whose stop tests have just run (custom_lbfgs.py:217-218 logs it only if none of them fired)
  const int tested = st.pending ? st.n_iter - 1 : st.n_iter;
  int logged_it = -1;
  double logged_f = 0.0;
  // (at most one iteration is newly logged per evaluation; when the loop ends on max_iter, `tested` also covers the final,
  // never-evaluated and never-logged iteration)
  for (int it = s->reported + 1; it <= tested && it >= 1; it++) {
    int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, s->d_logged + it, 4, cudaMemcpyDeviceToHost));
    if (flag) { logged_it = it; logged_f = f; }
  }
  if (tested > s->reported) s->reported = tested;
  if (logged_iter_out) *logged_iter_out = logged_it;
  if (logged_f_out) *logged_f_out = logged_f;
  // where to evaluate next (status 0), or the vector lbfgs() returns: the last update when the loop ended on max_iter
  const double* src = (st.status == PINN_LBFGS_MAX_ITER) ? s->d_xfinal : s->d_x;
  CUDA_TRY(cudaMemcpy(x_next, src, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int pinn_lbfgs_f_hist(pinn_lbfgs_t* s, double* f_hist_out, int capacity, int* n_out) {
  if (!s || !n_out) return fail("pinn_lbfgs_f_hist: null argument");
  CUDA_TRY(cudaSetDevice(s->device));
  pinn::LbfgsState st{};
  CUDA_TRY(cudaMemcpy(&st, s->d_st, sizeof(st), cudaMemcpyDeviceToHost));
  *n_out = st.n_eval;
  if (f_hist_out) {
    if (capacity < st.n_eval) return fail("pinn_lbfgs_f_hist: buffer too small");
    if (st.n_eval) CUDA_TRY(cudaMemcpy(f_hist_out, s->d_fhist, (size_t)st.n_eval * 8, cudaMemcpyDeviceToHost));
  }
  return 0;
}

}  // extern "C"
