// Cross-CTA reduction, on-device Adam (TF-2.0 semantics) and the device-resident two-loop L-BFGS.
// Replaces: utils/neuralnetwork.py:112-116 (+ Keras Adam), utils/custom_lbfgs.py:39-236.
#pragma once
#include "pinn_common.cuh"

namespace pinn {

// Layout of the reduced vector R: [0, P) gradient; [P, P+3) loss parts (data|ic, boundary, residual).
// It is contiguous so that one ncclAllReduce(P+3) covers gradient and loss.

struct LbfgsState {
  int status;       // PINN_LBFGS_*; 0 = running
  int n_iter;
  int n_eval;
  int k;            // history length
  int head;         // ring slot of the oldest pair
  int pending;      // 1: an evaluation has been enqueued whose stop tests have not run yet
  int max_iter;
  int n_corr;
  double max_eval;
  double lr, tol_fun, tol_x;
  double h_diag, t, f, f_old;
  double ro[128];   // 1 / (y_i . s_i) per ring slot, cached when the pair is pushed (same dot product the reference
                    // recomputes every iteration, utils/custom_lbfgs.py:121-123)
};

// ------------------------------------------------------------------------------------------------
// reduce the per-CTA partial vectors (fixed order -> deterministic).  src index map: gradient entries
// [0,P_net) map 1:1; extra entries are gathered through `extra_src` (parameter and loss-part slots).
// ------------------------------------------------------------------------------------------------
struct ReduceMap {
  int n_out;          // entries of R to produce
  int p_net;          // first p_net entries copy 1:1
  int extra_src[8];   // source column for R[p_net + i]
  int n_extra;
};

// 8 lanes per output entry: each lane sums every 8th CTA partial, then a 3-step shuffle tree in fixed order
// (deterministic).  blockDim = 256 -> 32 entries per block.
__global__ void reduce_partials(const double* __restrict__ partials, int n_cta, int stride, double* __restrict__ R,
                                ReduceMap map, const int* __restrict__ run_flag) {
  if (run_flag && *run_flag != 0) return;
  const int sub = threadIdx.x & 7;
  const int i = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const bool ok = i < map.n_out;
  const int src = ok ? (i < map.p_net ? i : map.extra_src[i - map.p_net]) : 0;
  double s = 0.0;
  if (ok)
    for (int b = sub; b < n_cta; b += 8) s += partials[(size_t)b * stride + src];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) R[i] = s;
}

// ------------------------------------------------------------------------------------------------
// Adam (ResourceApplyAdam of TF 2.0):  alpha_t = lr sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1);
// v += (g^2-v)(1-b2); w -= alpha_t m / (sqrt(v)+eps).   step[0] = t (completed steps).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_entry(double* w, double* m, double* v, double g, int i, int t, double lr, double b1,
                                           double b2, double eps) {
  const double alpha = lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t));
  const double mi = m[i] + (g - m[i]) * (1.0 - b1);
  const double vi = v[i] + (g * g - v[i]) * (1.0 - b2);
  m[i] = mi;
  v[i] = vi;
  w[i] -= (mi * alpha) / (sqrt(vi) + eps);
}

// step[0] = t (completed steps), step[1] = blocks finished in the current launch.  The LAST block to finish advances t,
// so no block can observe the incremented counter (replaces a separate one-thread kernel).
__device__ __forceinline__ void adam_finish(int* step, int nblocks) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(step + 1, 1);
    if (done == nblocks - 1) { step[1] = 0; step[0] += 1; __threadfence(); }
  }
}

__global__ void adam_update(double* __restrict__ w, double* __restrict__ m, double* __restrict__ v,
                            const double* __restrict__ R, int P, int* __restrict__ step, double lr, double b1, double b2,
                            double eps, double* __restrict__ loss_ring, int ring) {
  const int t = step[0] + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P) adam_entry(w, m, v, R[i], i, t, lr, b1, b2, eps);
  if (i == 0) loss_ring[(t - 1) % ring] = R[P] + R[P + 1] + R[P + 2];
  adam_finish(step, gridDim.x);
}

// Single-GPU step: fixed-order reduction of the per-CTA partials fused with the Adam update (no collective between).
__global__ void reduce_adam(const double* __restrict__ partials, int n_cta, int stride, double* __restrict__ R, ReduceMap map,
                            double* __restrict__ w, double* __restrict__ m, double* __restrict__ v, int P,
                            int* __restrict__ step, double lr, double b1, double b2, double eps,
                            double* __restrict__ loss_ring, int ring) {
  const int t = step[0] + 1;
  const int sub = threadIdx.x & 7;
  const int i = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const bool ok = i < map.n_out;
  const int src = ok ? (i < map.p_net ? i : map.extra_src[i - map.p_net]) : 0;
  double s = 0.0;
  if (ok)
    for (int b = sub; b < n_cta; b += 8) s += partials[(size_t)b * stride + src];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (ok && sub == 0) {
    R[i] = s;
    if (i < P) adam_entry(w, m, v, s, i, t, lr, b1, b2, eps);
  }
  // the loss needs the three part sums, which live in (possibly) different blocks: the last block writes it
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(step + 1, 1);
    if (done == (int)gridDim.x - 1) {
      __threadfence();
      const volatile double* Rv = R;
      loss_ring[(t - 1) % ring] = Rv[P] + Rv[P + 1] + Rv[P + 2];
      step[1] = 0;
      step[0] += 1;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused reduce -> all-to-all PUSH over NVLink peer memory -> (Adam) : ONE kernel per rank and evaluation (multi-GPU, one
// process per GPU).  Replaces reduce_partials + ncclAllReduce + adam_update (3 launches + the NCCL kernel).
//
// Every rank owns an IPC-exported exchange buffer   data[2 parities][world][slot_len] doubles, flag[2][world][n_blocks] u64.
// Block b of rank r reduces its 32 entries of the per-CTA partials (fixed order), STORES them into slot [parity][r] of EVERY
// rank's buffer (sub-lane k of an entry writes to peer k: remote stores are fire-and-forget, nobody pulls), fences, and
// raises flag[parity][r][b] = seq on every peer.  It then waits on its LOCAL flags of all ranks for the same block --
// a local spin, no NVLink round trips -- sums the world slots in a fixed butterfly order (bitwise identical on all ranks,
// so replicated optimiser state never diverges) and applies Adam to its entries in the same pass.
// seq / parity come from a device-side counter of published evaluations (identical on all ranks), so an evaluation that
// is skipped on the device (L-BFGS stopped: *run_flag != 0 on every rank alike) consumes neither.  Two parities suffice: a
// rank publishes evaluation k+1 only after it has seen every peer's evaluation k, and a peer publishes k only after it
// has finished consuming k-1 (stream order), so slot parity (k+1)&1 is no longer being read anywhere.
// The wait is bounded (~20 s by the global timer): a dead peer raises *err instead of hanging the GPU, and the update is
// skipped; the host reports it at the next synchronising call.
// ------------------------------------------------------------------------------------------------
constexpr int P2P_MAX = 8;
struct XchgPeers {
  double* data[P2P_MAX];                 // peer exchange buffers
  unsigned long long* flag[P2P_MAX];
  int world, rank;
  int slot_len;                          // doubles per (parity, rank) slot
  int n_blocks;                          // flags per (parity, rank)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// xseq[0] = published evaluations so far, xseq[1] = blocks finished in the current launch
__global__ void __launch_bounds__(256)
reduce_exchange(const double* __restrict__ partials, int n_cta, int stride, ReduceMap map, const int* __restrict__ run_flag,
                XchgPeers peers, int* __restrict__ xseq, double* __restrict__ R, int* __restrict__ err, int adam,
                double* __restrict__ w, double* __restrict__ m, double* __restrict__ v, int P, int* __restrict__ step,
                double lr, double b1, double b2, double eps, double* __restrict__ loss_ring, int ring) {
  if (run_flag && *run_flag != 0) return;          // identical on every rank (replicated L-BFGS state)
  const unsigned long long seq = (unsigned long long)(*(volatile int*)xseq) + 1;
  const int parity = (int)(seq & 1);
  const int sub = threadIdx.x & 7;
  const int i = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3);
  const bool ok = i < map.n_out;
  const int src = ok ? (i < map.p_net ? i : map.extra_src[i - map.p_net]) : 0;
  double s = 0.0;
  if (ok)
    for (int b = sub; b < n_cta; b += 8) s += partials[(size_t)b * stride + src];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);          // every sub-lane holds the entry's local sum
  // ---- push: sub-lane k stores the entry into rank k's buffer, slot [parity][my rank]
  const size_t slot = ((size_t)parity * peers.world + peers.rank) * peers.slot_len;
  if (ok && sub < peers.world) peers.data[sub][slot + i] = s;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < peers.world)
    st_release_sys(peers.flag[threadIdx.x] + ((size_t)parity * peers.world + peers.rank) * peers.n_blocks + blockIdx.x, seq);
  // ---- wait for every rank's copy of this block's entries (local flags)
  if (threadIdx.x < peers.world) {
    const unsigned long long* f = peers.flag[peers.rank] + ((size_t)parity * peers.world + threadIdx.x) * peers.n_blocks + blockIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(f) < seq) {
      if (globaltimer_ns() - t0 > 20000000000ULL) { atomicExch(err, 1); break; }
      __nanosleep(32);
    }
  }
  __syncthreads();
  const bool dead = *(volatile int*)err != 0;       // a peer never published: leave R, weights and optimiser state untouched
  // ---- sum over ranks: sub-lane k loads rank k's value from the LOCAL buffer, butterfly in a fixed order
  double tot = 0.0;
  if (ok && sub < peers.world)
    tot = __ldcv(peers.data[peers.rank] + ((size_t)parity * peers.world + sub) * peers.slot_len + i);
  tot += __shfl_xor_sync(0xffffffffu, tot, 1);
  tot += __shfl_xor_sync(0xffffffffu, tot, 2);
  tot += __shfl_xor_sync(0xffffffffu, tot, 4);
  const int t = adam ? step[0] + 1 : 0;
  if (ok && sub == 0 && !dead) {
    R[i] = tot;
    if (adam && i < P) adam_entry(w, m, v, tot, i, t, lr, b1, b2, eps);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(xseq + 1, 1);
    if (done == (int)gridDim.x - 1) {
      __threadfence();
      if (adam && !dead) {
        const volatile double* Rv = R;
        loss_ring[(t - 1) % ring] = Rv[P] + Rv[P + 1] + Rv[P + 2];
        step[0] += 1;
      }
      xseq[1] = 0;
      xseq[0] += 1;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// L-BFGS iteration kernel: ONE CTA; each thread keeps EPT entries of the working vector in registers.
// Runs (a) the stop tests of the previous iteration's evaluation, (b) the memory update and two-loop
// recursion, (c) the step, exactly in the reference order (utils/custom_lbfgs.py:81-221).
// ------------------------------------------------------------------------------------------------
// Block-wide sum in a fixed order, ONE barrier per call: the per-warp partials go to alternating halves of `red`, so a
// call never overwrites values another warp may still be reading (a warp can be at most one call ahead).
template <int NT>
__device__ __forceinline__ double block_sum_t(double v, double* red, int& phase) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* buf = red + (phase & 1) * 32;
  phase++;
  if (lane == 0) buf[warp] = v;
  __syncthreads();
  constexpr int NW = NT / 32;
  double s = (lane < NW) ? buf[lane] : 0.0;
#pragma unroll
  for (int m = NW / 2; m > 0; m >>= 1) s += shfl_xor_d(s, m);
  return __shfl_sync(0xffffffffu, s, 0);   // every thread holds the total
}

template <int EPT, int LB_THREADS>
__global__ void __launch_bounds__(LB_THREADS, 1)
lbfgs_iterate(LbfgsState* __restrict__ st, double* __restrict__ w, const double* __restrict__ R, int P,
              double* __restrict__ g_old, double* __restrict__ d, double* __restrict__ S, double* __restrict__ Y,
              double* __restrict__ x_final, double* __restrict__ f_hist, int* __restrict__ logged) {
  __shared__ double red[64];
  __shared__ double ro[128], al[128];
  int red_phase = 0;
  auto block_sum = [&](double v, double*) { return block_sum_t<LB_THREADS>(v, red, red_phase); };
  const int tid = threadIdx.x;
  if (st->status != 0) return;

  double gv[EPT];
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    const int i = tid + e * LB_THREADS;
    gv[e] = i < P ? R[i] : 0.0;
  }
  const double f_new = R[P] + R[P + 1] + R[P + 2];

  // ---- (a) bookkeeping + stop tests for the evaluation that has just completed
  if (st->pending) {
    const int n_iter = st->n_iter;
    const int n_eval = st->n_eval + 1;
    double a1 = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; e++) a1 += fabs(gv[e]);
    a1 = block_sum(a1, red);
    int status = 0;
    if (n_iter == 0) {
      // initial evaluation (custom_lbfgs.py:65-76)
      if (a1 <= st->tol_fun) status = 7;
    } else {
      double a2 = 0.0;
      const double t = st->t;
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int i = tid + e * LB_THREADS;
        a2 += i < P ? fabs(d[i] * t) : 0.0;
      }
      a2 = block_sum(a2, red);
      if ((double)n_eval >= st->max_eval) status = 2;                    // :195
      else if (a1 <= st->tol_fun) status = 3;                            // :200-204
      else if (a2 <= st->tol_x) status = 4;                              // :206-210
      else if (fabs(f_new - st->f_old) < st->tol_x) status = 5;          // :212-215
    }
    __syncthreads();
    if (tid == 0) {
      st->n_eval = n_eval;
      st->f = f_new;
      st->pending = 0;
      f_hist[n_eval - 1] = f_new;
      if (n_iter > 0 && status == 0) logged[n_iter] = 1;                 // :217-218 (log after the stop tests)
      st->status = status;
    }
    __syncthreads();
    if (status != 0) return;
  }

  // ---- (b) direction
  const int n_iter = st->n_iter + 1;
  const int n_corr = st->n_corr;
  double dv[EPT];
  if (n_iter == 1) {                                                     // :91-95
#pragma unroll
    for (int e = 0; e < EPT; e++) dv[e] = -gv[e];
  } else {
    const double t_prev = st->t;
    double yv[EPT], sv[EPT];
    double ys = 0.0, yy = 0.0;
#pragma unroll
    for (int e = 0; e < EPT; e++) {
      const int i = tid + e * LB_THREADS;
      yv[e] = i < P ? gv[e] - g_old[i] : 0.0;                            // :98
      sv[e] = i < P ? d[i] * t_prev : 0.0;                               // :99
      ys += yv[e] * sv[e];
      yy += yv[e] * yv[e];
    }
    ys = block_sum(ys, red);
    int k = st->k, head = st->head;
    double h_diag = st->h_diag;
    if (ys > 1e-10) {                                                    // :102-114
      yy = block_sum(yy, red);
      int slot;
      if (k == n_corr) { slot = head; head = (head + 1) % n_corr; }      // drop the oldest pair
      else { slot = (head + k) % n_corr; k++; }
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int i = tid + e * LB_THREADS;
        if (i < P) { S[(size_t)slot * P + i] = sv[e]; Y[(size_t)slot * P + i] = yv[e]; }
      }
      h_diag = ys / yy;
      if (tid == 0) st->ro[slot] = 1.0 / ys;
      __syncthreads();
    }
    // ro_i (i = 0 oldest) from the per-slot cache
    for (int i = tid; i < k; i += LB_THREADS) ro[i] = st->ro[(head + i) % n_corr];
    __syncthreads();
    double qv[EPT];
#pragma unroll
    for (int e = 0; e < EPT; e++) qv[e] = -gv[e];                        // :130
    // first loop (:131-133), newest to oldest; the next pair is prefetched while the current reduction runs
    double sn[EPT], yn[EPT];
    if (k > 0) {
      const size_t base = (size_t)((head + k - 1) % n_corr) * P;
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int j = tid + e * LB_THREADS;
        sn[e] = j < P ? S[base + j] : 0.0;
        yn[e] = j < P ? Y[base + j] : 0.0;
      }
    }
    for (int i = k - 1; i >= 0; i--) {
      double sc[EPT], yc[EPT];
#pragma unroll
      for (int e = 0; e < EPT; e++) { sc[e] = sn[e]; yc[e] = yn[e]; }
      if (i > 0) {
        const size_t base = (size_t)((head + i - 1) % n_corr) * P;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
          const int j = tid + e * LB_THREADS;
          sn[e] = j < P ? S[base + j] : 0.0;
          yn[e] = j < P ? Y[base + j] : 0.0;
        }
      }
      double a = 0.0;
#pragma unroll
      for (int e = 0; e < EPT; e++) a += sc[e] * qv[e];
      a = block_sum(a, red) * ro[i];
      if (tid == 0) al[i] = a;
#pragma unroll
      for (int e = 0; e < EPT; e++) qv[e] -= a * yc[e];
    }
    __syncthreads();
#pragma unroll
    for (int e = 0; e < EPT; e++) qv[e] *= h_diag;                       // :136
    // second loop (:137-139), oldest to newest
    if (k > 0) {
      const size_t base = (size_t)(head % n_corr) * P;
#pragma unroll
      for (int e = 0; e < EPT; e++) {
        const int j = tid + e * LB_THREADS;
        sn[e] = j < P ? S[base + j] : 0.0;
        yn[e] = j < P ? Y[base + j] : 0.0;
      }
    }
    for (int i = 0; i < k; i++) {
      double sc[EPT], yc[EPT];
#pragma unroll
      for (int e = 0; e < EPT; e++) { sc[e] = sn[e]; yc[e] = yn[e]; }
      if (i + 1 < k) {
        const size_t base = (size_t)((head + i + 1) % n_corr) * P;
#pragma unroll
        for (int e = 0; e < EPT; e++) {
          const int j = tid + e * LB_THREADS;
          sn[e] = j < P ? S[base + j] : 0.0;
          yn[e] = j < P ? Y[base + j] : 0.0;
        }
      }
      double a = 0.0;
#pragma unroll
      for (int e = 0; e < EPT; e++) a += yc[e] * qv[e];
      const double be = block_sum(a, red) * ro[i];
      const double c = al[i] - be;
#pragma unroll
      for (int e = 0; e < EPT; e++) qv[e] += c * sc[e];
    }
#pragma unroll
    for (int e = 0; e < EPT; e++) dv[e] = qv[e];
    if (tid == 0) { st->k = k; st->head = head; st->h_diag = h_diag; }
  }

  // ---- (c) step
  double gtd = 0.0, a1 = 0.0;
#pragma unroll
  for (int e = 0; e < EPT; e++) { gtd += gv[e] * dv[e]; a1 += fabs(gv[e]); }
  gtd = block_sum(gtd, red);                                             // :151
  a1 = block_sum(a1, red);
  // g_old = g, f_old = f happen before the progress test (:144-145)
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    const int i = tid + e * LB_THREADS;
    if (i < P) { g_old[i] = gv[e]; d[i] = dv[e]; }
  }
  const double f_cur = st->f;
  if (gtd > -st->tol_x) {                                                // :154-156
    __syncthreads();
    if (tid == 0) { st->n_iter = n_iter; st->f_old = f_cur; st->status = 6; }
    return;
  }
  const double t = (n_iter == 1) ? fmin(1.0, 1.0 / a1) : st->lr;         // :159-163
  const bool last = (n_iter == st->max_iter);
#pragma unroll
  for (int e = 0; e < EPT; e++) {
    const int i = tid + e * LB_THREADS;
    if (i < P) {
      const double xn = w[i] + t * dv[e];                                // :174
      if (last) x_final[i] = xn;       // no evaluation follows (:176-182): the model keeps the previous weights
      else w[i] = xn;
    }
  }
  __syncthreads();
  if (tid == 0) {
    st->n_iter = n_iter;
    st->f_old = f_cur;
    st->t = t;
    if (last) st->status = 1;                                            // :192
    else st->pending = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Generic (any layer sizes, width <= MAXW) thread-per-point kernels, off the per-step path:
// forward (predict, utils/neuralnetwork.py:151-153) and forward Taylor-mode derivatives (parity probes,
// f_model on the stored points).  Weights are read through the read-only path (warp-uniform addresses).
// ------------------------------------------------------------------------------------------------
constexpr int MAXW = 128;
constexpr int MAXL = 16;

struct NetDesc {
  int n_layers;            // number of Dense layers
  int dims[MAXL + 1];      // layer sizes
  int woff[MAXL], boff[MAXL];
  double lb0, lb1, dx0, dx1;
};

// NS = 1: values only; NS = 4: (h, h_x, h_t, h_xx)
template <int NS>
__global__ void mlp_forward_generic(const double* __restrict__ w, NetDesc nd, const double* __restrict__ X,
                                    const double* __restrict__ Tsoa, long long n, int in_dim, double* __restrict__ out) {
  // points either interleaved X[n][in_dim] (Tsoa == nullptr) or SoA (X = x[n], Tsoa = t[n])
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  double h[NS][MAXW], z[NS][MAXW];
  const double x = Tsoa ? X[p] : X[p * in_dim];
  const double t = Tsoa ? Tsoa[p] : (in_dim == 1 ? x : X[p * in_dim + 1]);
  h[0][0] = 2.0 * (x - nd.lb0) / nd.dx0 - 1.0;
  h[0][1] = nd.dims[0] == 2 ? 2.0 * (t - nd.lb1) / nd.dx1 - 1.0 : 0.0;   // 1-D nets (discrete-time models) ignore t
  if (NS == 4) {
    h[1][0] = 2.0 / nd.dx0; h[1][1] = 0.0;
    h[2][0] = 0.0;          h[2][1] = 2.0 / nd.dx1;
    h[3][0] = 0.0;          h[3][1] = 0.0;
  }
  for (int l = 0; l < nd.n_layers; l++) {
    const int fi = nd.dims[l], fo = nd.dims[l + 1];
    const double* Wl = w + nd.woff[l];
    const double* bl = w + nd.boff[l];
    if (l == nd.n_layers - 1) {
      // linear head: written straight to the output (it may be wider than MAXW, e.g. q+1 = 501 IRK stages)
      for (int j = 0; j < fo; j++) {
        double acc[NS];
        acc[0] = __ldg(bl + j);
#pragma unroll
        for (int s = 1; s < NS; s++) acc[s] = 0.0;
        for (int i = 0; i < fi; i++) {
          const double wv = __ldg(Wl + i * fo + j);
#pragma unroll
          for (int s = 0; s < NS; s++) acc[s] = fma(h[s][i], wv, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < NS; s++) out[p * (NS * fo) + s * fo + j] = acc[s];
      }
      return;
    }
    for (int j = 0; j < fo; j++) {
      double acc[NS];
      acc[0] = __ldg(bl + j);
#pragma unroll
      for (int s = 1; s < NS; s++) acc[s] = 0.0;
      for (int i = 0; i < fi; i++) {
        const double wv = __ldg(Wl + i * fo + j);
#pragma unroll
        for (int s = 0; s < NS; s++) acc[s] = fma(h[s][i], wv, acc[s]);
      }
#pragma unroll
      for (int s = 0; s < NS; s++) z[s][j] = acc[s];
    }
    const bool last = (l == nd.n_layers - 1);
    for (int j = 0; j < fo; j++) {
      if (last) {
#pragma unroll
        for (int s = 0; s < NS; s++) h[s][j] = z[s][j];
      } else {
        const double a = tanh(z[0][j]);
        h[0][j] = a;
        if (NS == 4) {
          const double sd = fma(-a, a, 1.0);
          const double zx = z[1][j];
          h[1][j] = sd * zx;
          h[2][j] = sd * z[2][j];
          h[3][j] = sd * fma(-2.0 * a * zx, zx, z[3][j]);
        }
      }
    }
  }
  const int no = nd.dims[nd.n_layers];
  for (int s = 0; s < NS; s++)
    for (int j = 0; j < no; j++) out[p * (NS * no) + s * no + j] = h[s][j];
}


__global__ void tanh_fast_kernel(const double* __restrict__ x, int n, double* __restrict__ y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = tanh_fast(x[i]);
}

}  // namespace pinn
