// Fused Burgers PINN loss + gradient kernel, v2: warp-specialised (sm_100a, fp64 DMMA).
//
// Same mathematics and fragment scheme as burgers_fused.cuh (v1, kept as a cross-check), re-organised so that
// the FP64 pipe of every SM sub-partition always has two warps to draw from and nobody waits at a block barrier:
//
//   warps 0-3  "chain" warps  : own an 8-point tile each; forward DMMA chain through the 8 tanh layers, residual,
//                               seeds, then the backward chain (activation adjoints + input-adjoint DMMAs).  At
//                               every backward layer they stage Z-bar (the adjoint of the pre-activations, 4
//                               streams x 8 points x 20 units) into a 2-slot ring in shared memory and signal an
//                               mbarrier.  The stash written during the forward pass holds the layer OUTPUTS
//                               (a, a_x, a_t, a_xx) in [row][unit] layout, which is both what the division-free
//                               backward formula needs and the A operand of the weight-gradient GEMM.
//   warps 4-7  "wgrad" warps  : consume (stash[l-1], Z-bar_l) pairs of all four chain warps and run the
//                               weight-gradient GEMM  G_l[i][j] += sum_rows H_{l-1}[row][i] Zbar_l[row][j]  with DMMA.
//                               The 3x3 output tiles of a layer are OWNED by warps (3/2/2/2, the heavy role rotates
//                               with the layer so the four sub-partitions stay balanced), so accumulators live in
//                               registers for the whole kernel: no shared-memory accumulator, no atomics, no
//                               cross-warp reduction, and a fixed summation order (deterministic results).
//
// Backward through a tanh layer in terms of its outputs (s = 1 - a^2; A* = adjoints of the outputs):
//   Z_xx = s A_xx;  Z_t = s A_t;  Z_x = s A_x - 4 a a_x A_xx;
//   Z    = s A - 2 a (a_x A_x + a_t A_t) - 2 A_xx (a a_xx + a_x^2)
// (algebraically identical to SURVEY Appendix A; no division by s, so saturated units are safe).
#pragma once
#include "burgers_fused.cuh"

namespace pinn {
namespace burgers2 {

using burgers::W;
using burgers::NHID;
using burgers::P_NET;
using burgers::WPAD;
using burgers::TILE;
using burgers::PSTRIDE;
using burgers::IDX_DL1;
using burgers::IDX_DL2;
using burgers::IDX_LD;
using burgers::IDX_LF;
using burgers::woff;
using burgers::boff;
using burgers::prow;
using burgers::Args;

constexpr int CHAINS = 4;
constexpr int THREADS = 256;              // 4 chain warps + 4 wgrad warps
constexpr int ROUND = CHAINS * TILE;      // 32 points per CTA round
#ifndef PINN_RING
#define PINN_RING 2
#endif
#ifndef PINN_PHASE_OFFSET
#define PINN_PHASE_OFFSET 0
#endif
#ifndef PINN_XT_PREFETCH
#define PINN_XT_PREFETCH 1
#endif
#ifndef PINN_WG_ROLLED
#define PINN_WG_ROLLED 0
#endif
#ifndef PINN_WG_FAST
#define PINN_WG_FAST 0                    // experiment (unmeasured): unconditional weight-gradient operand loads, see wgrad_task
#endif
#ifndef PINN_WG_ROWS
#define PINN_WG_ROWS 1                    // weight-gradient tiles owned by ROWS: warp position k < 3 owns (mt = k; nt = 0,1,2), k == 3 rests;
#endif                                    // one A fragment feeds three DMMAs, operand pointers chosen once (no predicates in the k loop)
#ifndef PINN_ABL_NOACT
#define PINN_ABL_NOACT 0                  // timing ablation: no activation arithmetic in the chain warps
#endif
#ifndef PINN_ABL_NODMMA
#define PINN_ABL_NODMMA 0                 // timing ablation: no chain DMMAs
#endif
#ifndef PINN_ABL_NOSTAGE
#define PINN_ABL_NOSTAGE 0                // timing ablation: chain warps neither stage nor synchronise (use with PINN_ABL_NOWG)
#endif
#ifndef PINN_ABL_NOWG
#define PINN_ABL_NOWG 0                   // timing ablation (wrong gradients): weight-gradient warps only do the ring handshake
#endif
constexpr int RING = PINN_RING;           // Z-bar ring slots per chain warp

// shared memory carve-up (doubles)
constexpr int STASH0 = 160;               // layer 0: a only, [8 rows][20]
constexpr int STASHL = 640;               // layers 1..6: outputs (a, a_x, a_t, a_xx) as [32 rows][20]
constexpr int STASH_PER_WARP = STASH0 + 6 * STASHL;   // 4000
constexpr int SM_W = 0;
constexpr int SM_STASH = SM_W + WPAD;
constexpr int SM_RING = SM_STASH + CHAINS * STASH_PER_WARP;
constexpr int SM_XT = SM_RING + CHAINS * RING * 640;
constexpr int SM_RED = SM_XT + CHAINS * 2 * 16;
constexpr int SM_BAR = SM_RED + 256;      // 1 + 2*CHAINS*RING mbarriers
constexpr int SM_SPECIAL = SM_BAR + 1 + 2 * CHAINS * RING + 2;   // + phase-offset barrier; then (PINN_WG_FAST) a [32 rows][W] page:
                                                                 // column 0 = the bias ones-row (1 on the value stream's rows 0..7), column 1 = 0
constexpr int SM_DOUBLES = SM_SPECIAL + ((PINN_WG_FAST || PINN_WG_ROWS) ? 32 * W : 0);
constexpr int SMEM_BYTES = SM_DOUBLES * 8;   // ~195 KB

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Ring bookkeeping: chain warp c numbers its weight-gradient tasks T = 8*tile + J (J = 7 - layer).  Task T lives in ring
// slot T % RING; it is the (T / RING)-th use of that slot, i.e. phase T / RING of the slot's full and empty barriers.
__device__ __forceinline__ uint64_t* bar_full(uint64_t* bars, int c, int slot) { return bars + 1 + (c * RING + slot) * 2; }
__device__ __forceinline__ uint64_t* bar_empty(uint64_t* bars, int c, int slot) { return bars + 2 + (c * RING + slot) * 2; }
__device__ __forceinline__ void wait_produced(uint64_t* bars, int c, int T) { mbar_wait(bar_full(bars, c, T % RING), (T / RING) & 1); }
__device__ __forceinline__ void wait_consumed(uint64_t* bars, int c, int T) { mbar_wait(bar_empty(bars, c, T % RING), (T / RING) & 1); }

// C-layout read of a [rows][20] staged tile: stream block s, this lane's point row and columns
__device__ __forceinline__ void load_rows(double (&V)[4][3][2], const double* T, int lane) {
  const int pg = prow(lane >> 2), q = lane & 3;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const double* row = T + (8 * s + pg) * W;
    const double2 c0 = *reinterpret_cast<const double2*>(row + 2 * q);
    const double2 c1 = *reinterpret_cast<const double2*>(row + 8 + 2 * q);
    double2 c2 = make_double2(0.0, 0.0);
    if (q < 2) c2 = *reinterpret_cast<const double2*>(row + 16 + 2 * q);
    V[s][0][0] = c0.x; V[s][0][1] = c0.y;
    V[s][1][0] = c1.x; V[s][1][1] = c1.y;
    V[s][2][0] = c2.x; V[s][2][1] = c2.y;
  }
}

// outputs of a tanh layer from its pre-activation streams (in place): Z -> (a, s z_x, s z_t, s (z_xx - 2 a z_x^2))
__device__ __forceinline__ void act_forward(double (&Z)[4][3][2]) {
  if (PINN_ABL_NOACT) return;
#pragma unroll
  for (int nt = 0; nt < 3; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const double a = tanh_fast(Z[0][nt][e]);
      const double zx = Z[1][nt][e];
      const double s = fma(-a, a, 1.0);
      Z[0][nt][e] = a;
      Z[1][nt][e] = s * zx;
      Z[2][nt][e] = s * Z[2][nt][e];
      Z[3][nt][e] = s * fma(-2.0 * a * zx, zx, Z[3][nt][e]);
    }
}

// adjoints of the outputs (A, overwritten with the adjoints of the pre-activations) given the outputs H
__device__ __forceinline__ void act_backward_out(double (&A)[4][3][2], const double (&H)[4][3][2]) {
  if (PINN_ABL_NOACT) return;
#pragma unroll
  for (int nt = 0; nt < 3; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const double a = H[0][nt][e], ax = H[1][nt][e], at = H[2][nt][e], axx = H[3][nt][e];
      const double A0 = A[0][nt][e], Ax = A[1][nt][e], At = A[2][nt][e], Axx = A[3][nt][e];
      const double s = fma(-a, a, 1.0);
      const double u1 = fma(ax, Ax, at * At);
      const double u2 = fma(a, axx, ax * ax);
      double z = fma(-2.0 * a, u1, s * A0);
      z = fma(-2.0 * Axx, u2, z);
      A[0][nt][e] = z;
      A[1][nt][e] = fma(-4.0 * a * ax, Axx, s * Ax);
      A[2][nt][e] = s * At;
      A[3][nt][e] = s * Axx;
    }
}

// ---- wgrad-warp tile ownership.  Hidden layer l: warp position k = (wg - l) & 3; k == 0 owns tiles 0,1,2, k owns
// tiles 2k+1, 2k+2 (tile t = 3*mt + nt).  Layer 0 (MT = 1): warp 0 owns nt 0,1; warp 1 owns nt 2.
struct Own { int n; int mt[3]; int nt[3]; };
__device__ __forceinline__ Own ownership(int l, int wg) {
  Own o;
  if (PINN_WG_ROWS) {
    const int k = l == 0 ? wg : ((wg - l) & 3);
    o.n = (l == 0 ? k == 0 : k < 3) ? 3 : 0;
    o.mt[0] = o.mt[1] = o.mt[2] = l == 0 ? 0 : (k < 3 ? k : 0);
    o.nt[0] = 0; o.nt[1] = 1; o.nt[2] = 2;
    return o;
  }
  if (l == 0) {
    o.n = wg == 0 ? 2 : (wg == 1 ? 1 : 0);
    o.mt[0] = o.mt[1] = o.mt[2] = 0;
    o.nt[0] = wg == 0 ? 0 : 2; o.nt[1] = 1; o.nt[2] = 0;
    return o;
  }
  const int k = (wg - l) & 3;
  o.n = k == 0 ? 3 : 2;
  const int ft = k == 0 ? 0 : 2 * k + 1;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const int t = ft + s;
    o.mt[s] = t / 3;
    o.nt[s] = t - 3 * (t / 3);
  }
  return o;
}

// one weight-gradient task: accumulate this warp's owned tiles of layer L for one chain warp's 8-point tile
template <int L>
__device__ __forceinline__ void wgrad_task(double (&acc)[3][2], const Own& o, const double* Aop, const double* ZB,
                                           const double* Wsm, double sc0, double sc1, int lane) {
  const int g = lane >> 2, q = lane & 3;
  // layer-1 inputs are synthesised from the a-only stash of layer 0: per-lane constants of the units this lane reads
  double w0x[3], w0t[3];
  if (L == 1) {
#pragma unroll
    for (int s = 0; s < 3; s++) {
      const int i = 8 * o.mt[s] + g;
      w0x[s] = i < W ? sc0 * Wsm[i] : 0.0;
      w0t[s] = i < W ? sc1 * Wsm[W + i] : 0.0;
    }
  }
  if (PINN_WG_ROWS) {
    // row ownership: tiles (mt; nt = 0,1,2).  B column pointers (padded columns read the zero column of the special page) and,
    // for regular layers, the A column pointer (bias unit -> ones column, padding -> zero column) are chosen once.
    const double* SP = Wsm - SM_W + SM_SPECIAL;
    const int i = 8 * o.mt[0] + g;
    const double* bp0 = ZB + g + q * W;
    const double* bp1 = ZB + 8 + g + q * W;
    const double* bp2 = (16 + g < W ? ZB + 16 + g : SP + 1) + q * W;
    const double* ap = (i < W ? Aop + i : (i == W ? SP : SP + 1)) + q * W;          // L >= 2: [32 rows][W] outputs of layer L-1
    const double* ap1 = (i < W ? Aop + i : SP + 1) + q * W;                          // L == 1: a-only stash [8 rows][W]
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
      const int st = ks >> 1;
      const int off = (8 * st + 4 * (ks & 1)) * W;              // physical row 8*stream + 4*(ks&1) (+ q, folded into the pointers)
      double a;
      if (L >= 2) {
        a = ap[off];
      } else if (L == 1) {
        const double av = ap1[4 * (ks & 1) * W];
        const double sd = fma(-av, av, 1.0);
        const double v = st == 0 ? av : (st == 1 ? sd * w0x[0] : (st == 2 ? sd * w0t[0] : -2.0 * av * sd * w0x[0] * w0x[0]));
        a = (i == W && st == 0) ? 1.0 : v;                      // i >= W: av = 0 and w0x = w0t = 0, so v = 0
      } else {
        const int pr = 4 * (ks & 1) + q;
        const double xv = Aop[pr * 2 + 0], tv = Aop[pr * 2 + 1];
        a = i == 0 ? (st == 0 ? xv : (st == 1 ? sc0 : 0.0))
                   : (i == 1 ? (st == 0 ? tv : (st == 2 ? sc1 : 0.0)) : ((i == 2 && st == 0) ? 1.0 : 0.0));
      }
      const double b0 = bp0[off], b1 = bp1[off], b2 = bp2[off];
      dmma(acc[0], a, b0);
      dmma(acc[1], a, b1);
      dmma(acc[2], a, b2);
    }
    return;
  }
  if (L >= 2 && PINN_WG_FAST) {
    // Experiment for the operand-fetch overhead seen in profiles/ncu_burgers_v2_r01_lines.md (2.2 instructions per operand per
    // DMMA): choose each owned tile's A and B column pointer ONCE -- a padded lane points into the special page (ones-row or
    // zeros) instead of being predicated -- so that the unrolled k loop is one LDS per operand with an immediate offset.
    const double* SP = Wsm - SM_W + SM_SPECIAL;
    const double* ap[3];
    const double* bp[3];
#pragma unroll
    for (int s = 0; s < 3; s++) {
      const int i = 8 * o.mt[s] + g, j = 8 * o.nt[s] + g;
      ap[s] = (i < W ? Aop + i : (i == W ? SP : SP + 1)) + q * W;
      bp[s] = (j < W ? ZB + j : SP + 1) + q * W;
    }
#pragma unroll
    for (int ks = 0; ks < 8; ks++) {
      const int off = (8 * (ks >> 1) + 4 * (ks & 1)) * W;       // physical row 8*stream + 4*(ks&1) (+ q, folded into the pointers)
#pragma unroll
      for (int s = 0; s < 3; s++)
        if (s < o.n) dmma(acc[s], ap[s][off], bp[s][off]);      // warp-uniform
    }
    return;
  }
  if (L >= 2 && PINN_WG_ROLLED) {
    // regular hidden layers: rolled k loop (keeps the kernel within the instruction cache); slots 0,1 always exist
    const int i0 = 8 * o.mt[0] + g, i1 = 8 * o.mt[1] + g, i2 = 8 * o.mt[2] + g;
    const int j0 = 8 * o.nt[0] + g, j1 = 8 * o.nt[1] + g, j2 = 8 * o.nt[2] + g;
    const bool three = o.n == 3;
#pragma unroll 2
    for (int ks = 0; ks < 8; ks++) {
      const int R = 4 * ks + q;
      const double one = ks < 2 ? 1.0 : 0.0;       // ones-row (bias) on the value stream: rows 0..7
      const double* ar = Aop + R * W;
      const double* br = ZB + R * W;
      const double a0 = i0 < W ? ar[i0] : (i0 == W ? one : 0.0);
      const double a1 = i1 < W ? ar[i1] : (i1 == W ? one : 0.0);
      const double b0 = j0 < W ? br[j0] : 0.0;
      const double b1 = j1 < W ? br[j1] : 0.0;
      dmma(acc[0], a0, b0);
      dmma(acc[1], a1, b1);
      if (three) {
        const double a2 = i2 < W ? ar[i2] : (i2 == W ? one : 0.0);
        const double b2 = j2 < W ? br[j2] : 0.0;
        dmma(acc[2], a2, b2);
      }
    }
    return;
  }
#pragma unroll
  for (int ks = 0; ks < 8; ks++) {
    const int st = ks >> 1;                 // stream of this k-step's rows
    const int pr = 4 * (ks & 1) + q;        // physical point row
    const int R = 8 * st + pr;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      if (s < o.n) {                        // warp-uniform
        const int i = 8 * o.mt[s] + g;
        const int j = 8 * o.nt[s] + g;
        double a;
        if (L >= 2) {
          a = i < W ? Aop[R * W + i] : ((i == W && st == 0) ? 1.0 : 0.0);
        } else if (L == 1) {
          const double av = i < W ? Aop[pr * W + i] : 0.0;
          const double sd = fma(-av, av, 1.0);
          const double v = st == 0 ? av : (st == 1 ? sd * w0x[s] : (st == 2 ? sd * w0t[s] : -2.0 * av * sd * w0x[s] * w0x[s]));
          a = i < W ? v : ((i == W && st == 0) ? 1.0 : 0.0);
        } else {
          // layer 0: inputs (x^, t^) on the value stream, (sc0, 0) on the x stream, (0, sc1) on the t stream; unit 2 = ones
          const double xv = Aop[pr * 2 + 0], tv = Aop[pr * 2 + 1];
          a = i == 0 ? (st == 0 ? xv : (st == 1 ? sc0 : 0.0))
                     : (i == 1 ? (st == 0 ? tv : (st == 2 ? sc1 : 0.0)) : ((i == 2 && st == 0) ? 1.0 : 0.0));
        }
        const double b = j < W ? ZB[R * W + j] : 0.0;
        dmma(acc[s], a, b);
      }
    }
  }
}

template <int L>
__device__ __forceinline__ void wgrad_layer(double (&acc)[3][2], int wg, int it, int half, const double* sm, uint64_t* bars,
                                            double sc0, double sc1, int lane) {
  constexpr int J = 7 - L;                  // task index within a tile (layers 7..0)
  const int T = 8 * it + J;
  const int slot = T % RING;
  const Own o = ownership(L, wg);
#pragma unroll 1
  for (int c = 2 * half; c < 2 * half + 2; c++) {
    wait_produced(bars, c, T);
    const double* stash = sm + SM_STASH + c * STASH_PER_WARP;
    const double* Aop = L >= 2 ? stash + STASH0 + (L - 2) * STASHL : (L == 1 ? stash : sm + SM_XT + (c * 2 + (it & 1)) * 16);
    const double* ZB = sm + SM_RING + (c * RING + slot) * 640;
    if (!PINN_ABL_NOWG && o.n > 0) wgrad_task<L>(acc, o, Aop, ZB, sm + SM_W, sc0, sc1, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty(bars, c, slot));
  }
}

template <int L>
__device__ __forceinline__ void wgrad_flush(const double (&acc)[3][2], int wg, double* outp, int lane) {
  const int g = lane >> 2, q = lane & 3;
  const Own o = ownership(L, wg);
  const int in_dim = L == 0 ? 2 : W;
#pragma unroll
  for (int s = 0; s < 3; s++) {
    if (s < o.n) {
      const int i = 8 * o.mt[s] + g;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int j = 8 * o.nt[s] + 2 * q + e;
        if (j < W) {
          if (i < in_dim) outp[woff(L) + i * W + j] = acc[s][e];
          else if (i == in_dim) outp[boff(L) + j] = acc[s][e];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) fused_loss_grad(const Args p) {
  extern __shared__ __align__(16) double sm[];
  if (p.run_flag && *p.run_flag != 0) return;
  double* Wsm = sm + SM_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BAR);   // [0] weights TMA; then (full, empty) per (chain, slot)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;

  if (PINN_WG_FAST || PINN_WG_ROWS)
    for (int i = threadIdx.x; i < 32 * W; i += THREADS) sm[SM_SPECIAL + i] = (i % W == 0 && i / W < 8) ? 1.0 : 0.0;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    for (int i = 0; i < CHAINS * RING; i++) {
      mbar_init(bars + 1 + 2 * i, 1);        // full: the producing chain warp's lane 0
      mbar_init(bars + 2 + 2 * i, 4);        // empty: lane 0 of each of the 4 wgrad warps
    }
    mbar_init(bars + 1 + 2 * CHAINS * RING, 2);   // phase offset: the two chain warps of half A
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bars, WPAD * 8);
    tma_bulk_g2s(Wsm, p.w, WPAD * 8, bars);
  }
  mbar_wait(bars, 0);

  const double sc0 = 2.0 / p.dx0, sc1 = 2.0 / p.dx1;
  const long long n_rounds = (p.n_total + ROUND - 1) / ROUND;
  const int my_rounds = (int)((n_rounds - blockIdx.x + gridDim.x - 1) / gridDim.x);
  double* outp = p.partials + (size_t)blockIdx.x * PSTRIDE;

  // Roles: chain warps 0-3, weight-gradient warps 4-7.  PINN_CHAIN_HIGH=1 swaps them (the issue arbiter is said to prefer
  // the highest warp id); measured: no gain (0.452 vs 0.445 ms), so the plain order stays.
#ifndef PINN_CHAIN_HIGH
#define PINN_CHAIN_HIGH 0
#endif
  const bool is_wgrad = PINN_CHAIN_HIGH ? (warp < CHAINS) : (warp >= CHAINS);
  if (is_wgrad) {
    // =========================================== wgrad warps ===========================================
    const int wg = warp & 3;
    double a0[3][2], a1[3][2], a2[3][2], a3[3][2], a4[3][2], a5[3][2], a6[3][2], a7[3][2];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
      for (int e = 0; e < 2; e++)
        a0[s][e] = a1[s][e] = a2[s][e] = a3[s][e] = a4[s][e] = a5[s][e] = a6[s][e] = a7[s][e] = 0.0;
#pragma unroll 1
    // The chain warps are consumed in two halves (chains 0,1 then chains 2,3).  Ring back-pressure then shifts the
    // halves by one phase: while one half runs its forward pass (which produces no weight-gradient work), the other
    // half runs its backward pass and keeps these warps -- and the FP64 pipe of every sub-partition -- busy.
    for (int it2 = 0; it2 < (PINN_ABL_NOSTAGE ? 0 : 2 * my_rounds); it2++) {
      const int it = it2 >> 1, half = it2 & 1;
      wgrad_layer<7>(a7, wg, it, half, sm, bars, sc0, sc1, lane);
      wgrad_layer<6>(a6, wg, it, half, sm, bars, sc0, sc1, lane);
      wgrad_layer<5>(a5, wg, it, half, sm, bars, sc0, sc1, lane);
      wgrad_layer<4>(a4, wg, it, half, sm, bars, sc0, sc1, lane);
      wgrad_layer<3>(a3, wg, it, half, sm, bars, sc0, sc1, lane);
      wgrad_layer<2>(a2, wg, it, half, sm, bars, sc0, sc1, lane);
      wgrad_layer<1>(a1, wg, it, half, sm, bars, sc0, sc1, lane);
      wgrad_layer<0>(a0, wg, it, half, sm, bars, sc0, sc1, lane);
    }
    wgrad_flush<0>(a0, wg, outp, lane);
    wgrad_flush<1>(a1, wg, outp, lane);
    wgrad_flush<2>(a2, wg, outp, lane);
    wgrad_flush<3>(a3, wg, outp, lane);
    wgrad_flush<4>(a4, wg, outp, lane);
    wgrad_flush<5>(a5, wg, outp, lane);
    wgrad_flush<6>(a6, wg, outp, lane);
    wgrad_flush<7>(a7, wg, outp, lane);
  } else {
    // =========================================== chain warps ===========================================
    const int c = warp & 3;
    double* stash = sm + SM_STASH + c * STASH_PER_WARP;
    const double l1 = p.ide ? Wsm[P_NET] : 1.0;
    const double kap = p.ide ? exp(Wsm[P_NET + 1]) : p.nu;
    double loss_d = 0.0, loss_f = 0.0, gl1 = 0.0, gl2 = 0.0;
    double g8[3][2], gb8 = 0.0;             // output-layer weight gradient: per-lane partials over this lane's points
#pragma unroll
    for (int nt = 0; nt < 3; nt++) g8[nt][0] = g8[nt][1] = 0.0;
    const int pg = prow(g);
    if (PINN_PHASE_OFFSET && c >= 2) mbar_wait(bars + 1 + 2 * CHAINS * RING, 0);   // start half a tile behind half A
    // coordinates of this lane's point in tile `it`; the next tile's pair is prefetched one tile ahead so that neither
    // the L2/HBM latency nor (zero-copy mode: p.xc in pinned host memory) the PCIe latency is exposed
    auto load_xt = [&](int it, double& xo, double& to) {
      const long long rnd = blockIdx.x + (long long)it * gridDim.x;
      long long pt = rnd * ROUND + c * TILE + g;
      if (pt >= p.n_total) pt = p.n_total - 1;
      if (p.xc && pt >= p.c0 && pt < p.c0 + p.n_c) {
        xo = __ldg(p.xc + (pt - p.c0)); to = __ldg(p.tc + (pt - p.c0));
      } else {
        xo = __ldg(p.x + pt); to = __ldg(p.t + pt);
      }
    };
    double xr_next = 0.0, tr_next = 0.0;
    if (PINN_XT_PREFETCH && my_rounds > 0) load_xt(0, xr_next, tr_next);

#pragma unroll 1
    for (int it = 0; it < my_rounds; it++) {
      const long long rnd = blockIdx.x + (long long)it * gridDim.x;
      const long long pt = rnd * ROUND + c * TILE + g;
      const bool in_set = pt < p.n_total;
      if (!PINN_XT_PREFETCH) load_xt(it, xr_next, tr_next);
      const double xr = xr_next, tr = tr_next;
      if (PINN_XT_PREFETCH && it + 1 < my_rounds) load_xt(it + 1, xr_next, tr_next);
      const double wf = (in_set && pt >= p.c0 && pt < p.c0 + p.n_c) ? p.wf : 0.0;
      const bool has_d = in_set && pt >= p.d0 && pt < p.d0 + p.n_d;
      const double wd = has_d ? p.wd : 0.0;
      const double ut = has_d ? __ldg(p.utgt + (pt - p.d0)) : 0.0;
      const double xh = 2.0 * (xr - p.lb0) / p.dx0 - 1.0;   // utils/neuralnetwork.py:29-30
      const double th = 2.0 * (tr - p.lb1) / p.dx1 - 1.0;
      double* XT = sm + SM_XT + (c * 2 + (it & 1)) * 16;
      if (q == 0) { XT[pg * 2 + 0] = xh; XT[pg * 2 + 1] = th; }

      double H[4][3][2];
      // ---------------- layer 0 (2 -> 20): direct
#pragma unroll
      for (int nt = 0; nt < 3; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int col = 8 * nt + 2 * q + e;
          const bool ok = col < W;
          const double w0 = ok ? Wsm[col] : 0.0, w1 = ok ? Wsm[W + col] : 0.0, b = ok ? Wsm[2 * W + col] : 0.0;
          H[0][nt][e] = fma(xh, w0, fma(th, w1, b));
          H[1][nt][e] = sc0 * w0;
          H[2][nt][e] = sc1 * w1;
          H[3][nt][e] = 0.0;
        }
      act_forward(H);
      // the a-only stash of layer 0 (and, with RING = 3, the layer-1 stash) is still the A operand of the previous
      // tile's last tasks: wait until the layer-1 task (J = 6) has been consumed; consumers retire tasks in order
      if (!PINN_ABL_NOSTAGE && it > 0) wait_consumed(bars, c, 8 * (it - 1) + 6);
      if (!PINN_ABL_NOSTAGE) {
        double* r0 = stash + pg * W;
        *reinterpret_cast<double2*>(r0 + 2 * q) = make_double2(H[0][0][0], H[0][0][1]);
        *reinterpret_cast<double2*>(r0 + 8 + 2 * q) = make_double2(H[0][1][0], H[0][1][1]);
        if (q < 2) *reinterpret_cast<double2*>(r0 + 16 + 2 * q) = make_double2(H[0][2][0], H[0][2][1]);
      }
      // ---------------- hidden layers 1..7: DMMA chain
#pragma unroll 1
      for (int l = 1; l < NHID; l++) {
        const double* Wl = Wsm + woff(l);
        double Z[4][3][2];
#pragma unroll
        for (int nt = 0; nt < 3; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int col = 8 * nt + 2 * q + e;
            Z[0][nt][e] = (col < W) ? Wl[W * W + col] : 0.0;
            Z[1][nt][e] = Z[2][nt][e] = Z[3][nt][e] = 0.0;
          }
        if (!PINN_ABL_NODMMA) burgers::mma_layer(Z, H, Wl, W, 1, lane);
        act_forward(Z);
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
          for (int nt = 0; nt < 3; nt++) { H[s][nt][0] = Z[s][nt][0]; H[s][nt][1] = Z[s][nt][1]; }
        if (!PINN_ABL_NOSTAGE && l < NHID - 1) burgers::stage_rows(stash + STASH0 + (l - 1) * STASHL, H, lane);
      }
      if (PINN_PHASE_OFFSET && it == 0 && c < 2) { __syncwarp(); if (lane == 0) mbar_arrive(bars + 1 + 2 * CHAINS * RING); }
      // ---------------- output layer (20 -> 1), residual, seeds
      double seed[4];
      double A[4][3][2];
      {
        const double* W8 = Wsm + woff(8);
        double w8[3][2];
#pragma unroll
        for (int nt = 0; nt < 3; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int col = 8 * nt + 2 * q + e;
            w8[nt][e] = (col < W) ? W8[col] : 0.0;
          }
        double out[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
          double acc = 0.0;
#pragma unroll
          for (int nt = 0; nt < 3; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) acc = fma(H[s][nt][e], w8[nt][e], acc);
          acc += shfl_xor_d(acc, 1);
          acc += shfl_xor_d(acc, 2);
          out[s] = acc;
        }
        out[0] += Wsm[boff(8)];
        const double u = out[0], ux = out[1], utt = out[2], uxx = out[3];
        const double f = utt + l1 * u * ux - kap * uxx;      // inf_cont_burgers.py:90 / ide_cont_burgers.py:85
        const double r = u - ut;
        const double cc = 2.0 * wf * f;
        seed[0] = fma(cc * l1, ux, 2.0 * wd * r);
        seed[1] = cc * l1 * u;
        seed[2] = cc;
        seed[3] = -cc * kap;
        if (q == 0) {
          loss_d = fma(wd * r, r, loss_d);
          loss_f = fma(wf * f, f, loss_f);
          gl1 = fma(cc * u, ux, gl1);
          gl2 = fma(-cc * kap, uxx, gl2);
          gb8 += seed[0];
        }
        // output-layer weight gradient (per-lane partial over this lane's point) and adjoint of layer-7 outputs
#pragma unroll
        for (int nt = 0; nt < 3; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            double acc = g8[nt][e];
#pragma unroll
            for (int s = 0; s < 4; s++) {
              acc = fma(H[s][nt][e], seed[s], acc);
              A[s][nt][e] = seed[s] * w8[nt][e];
            }
            g8[nt][e] = acc;
          }
      }
      // ---------------- backward: layers 7..1 (task J = 7-l, ring slot J&1)
#pragma unroll 1
      for (int l = NHID - 1; l >= 1; l--) {
        const int T = 8 * it + (7 - l), slot = T % RING;
        if (!PINN_ABL_NOSTAGE && l < NHID - 1) load_rows(H, stash + STASH0 + (l - 1) * STASHL, lane);   // outputs of layer l (l=7: registers)
        act_backward_out(A, H);                                                      // A := Z-bar
        if (!PINN_ABL_NOSTAGE) {
          if (T >= RING) wait_consumed(bars, c, T - RING);                           // the slot's previous task is done
          burgers::stage_rows(sm + SM_RING + (c * RING + slot) * 640, A, lane);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full(bars, c, slot));
        }
        // adjoint of the layer inputs: A_new = Z-bar * W_l^T
        {
          const double* Wl = Wsm + woff(l);
          double An[4][3][2];
#pragma unroll
          for (int s = 0; s < 4; s++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) An[s][nt][0] = An[s][nt][1] = 0.0;
          if (!PINN_ABL_NODMMA) burgers::mma_layer(An, A, Wl, 1, W, lane);
#pragma unroll
          for (int s = 0; s < 4; s++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) { A[s][nt][0] = An[s][nt][0]; A[s][nt][1] = An[s][nt][1]; }
        }
      }
      // ---------------- backward: layer 0 (task 7, slot 1): outputs rebuilt from the a-only stash
      {
        const double* r0 = stash + pg * W;
        const double2 c0 = *reinterpret_cast<const double2*>(r0 + 2 * q);
        const double2 c1 = *reinterpret_cast<const double2*>(r0 + 8 + 2 * q);
        double2 c2 = make_double2(0.0, 0.0);
        if (q < 2) c2 = *reinterpret_cast<const double2*>(r0 + 16 + 2 * q);
        const double av[3][2] = {{c0.x, c0.y}, {c1.x, c1.y}, {c2.x, c2.y}};
#pragma unroll
        for (int nt = 0; nt < 3; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int col = 8 * nt + 2 * q + e;
            const bool ok = col < W;
            const double a = av[nt][e], s = fma(-a, a, 1.0);
            const double zx = ok ? sc0 * Wsm[col] : 0.0, zt = ok ? sc1 * Wsm[W + col] : 0.0;
            H[0][nt][e] = a;
            H[1][nt][e] = s * zx;
            H[2][nt][e] = s * zt;
            H[3][nt][e] = -2.0 * a * s * zx * zx;
          }
        act_backward_out(A, H);
        const int T = 8 * it + 7, slot = T % RING;
        if (!PINN_ABL_NOSTAGE) {
          if (T >= RING) wait_consumed(bars, c, T - RING);
          burgers::stage_rows(sm + SM_RING + (c * RING + slot) * 640, A, lane);
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_full(bars, c, slot));
        }
      }
    }

    // ---------------- chain-warp partials: losses, identification gradients, output-layer gradient
    double* red = sm + SM_RED + c * 64;
    loss_d = warp_sum(loss_d); loss_f = warp_sum(loss_f); gl1 = warp_sum(gl1); gl2 = warp_sum(gl2); gb8 = warp_sum(gb8);
#pragma unroll
    for (int nt = 0; nt < 3; nt++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        double v = g8[nt][e];
        v += shfl_xor_d(v, 4); v += shfl_xor_d(v, 8); v += shfl_xor_d(v, 16);   // sum over the 8 points (g)
        const int col = 8 * nt + 2 * q + e;
        if (g == 0 && col < W) red[8 + col] = v;
      }
    if (lane == 0) { red[0] = loss_d; red[1] = loss_f; red[2] = gl1; red[3] = gl2; red[4] = gb8; }
  }

  __syncthreads();
  {
    const double* red = sm + SM_RED;
    const int t = threadIdx.x;
    if (t < W) outp[woff(8) + t] = red[8 + t] + red[64 + 8 + t] + red[128 + 8 + t] + red[192 + 8 + t];
    if (t == 32) outp[boff(8)] = red[4] + red[64 + 4] + red[128 + 4] + red[192 + 4];
    if (t == 33) outp[IDX_LD] = red[0] + red[64] + red[128] + red[192];
    if (t == 34) outp[IDX_LF] = red[1] + red[65] + red[129] + red[193];
    if (t == 35) outp[IDX_DL1] = red[2] + red[66] + red[130] + red[194];
    if (t == 36) outp[IDX_DL2] = red[3] + red[67] + red[131] + red[195];
    if (t == 37) outp[3023] = 0.0;
  }
}

}  // namespace burgers2
}  // namespace pinn
