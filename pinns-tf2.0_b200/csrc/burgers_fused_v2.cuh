// Fused Burgers PINN loss + gradient kernel, v2: warp-specialised (sm_100a, fp64 DMMA).
//
// Same mathematics and DMMA fragment scheme as burgers_fused.cuh (v1, kept as a cross-check), re-organised so that
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
//                               The 3x3 output tiles of a layer are OWNED by warps, a ROW of tiles each: position
//                               k = (warp - layer) & 3 < 3 owns tiles (mt = k; nt = 0,1,2), position 3 takes one m-tile off
//                               the busiest position at layers <= 3 and rests otherwise (ownership masks: 120/128/128/128
//                               DMMAs per tile and warp; the roles rotate with the layer).  One A fragment
//                               feeds up to three DMMAs, operand column pointers are chosen once per task (padded lanes point
//                               into a ones/zeros page), so the unrolled k loop is 4 LDS + 3 DMMA without predicates.
//                               Accumulators live in registers for the whole kernel: no shared-memory accumulator, no
//                               atomics, no cross-warp reduction, and a fixed summation order (deterministic results).
//
// Register layout of a chain warp ("V5").  A DMMA accumulator fragment (lane = 4g+q) holds point g, units 8nt+2q+e; hidden
// width 20 makes the third N tile half padding (units 20..23 in lanes q >= 2).  Between the GEMMs the 20 units of a point
// are therefore kept as FIVE values per lane and stream: j = 2nt+e for the first two tiles and, for j = 4, unit
// {16,18,17,19}[q] -- one quad shuffle moves the odd units of tile 2 from lanes q < 2 to lanes q >= 2.  This is exactly
// the A-operand layout of k-step j of the next GEMM (the contraction index is permuted: k-step j contracts the units the
// four lanes of a quad hold in slot j), so accumulators chain from layer to layer in registers, and the activation
// arithmetic (tanh and the Taylor/adjoint formulas) runs on 5 instead of 6 values per lane -- no lane computes padding.
//
// Backward through a tanh layer in terms of its outputs (s = 1 - a^2; A* = adjoints of the outputs):
//   Z_xx = s A_xx;  Z_t = s A_t;  Z_x = s A_x - 4 a a_x A_xx;
//   Z    = s A - 2 a (a_x A_x + a_t A_t) - 2 A_xx (a a_xx + a_x^2)
// (algebraically identical to SURVEY Appendix A; no division by s, so saturated units are safe).
//
// Work distribution is tile-granular (8 points): the tiles are dealt evenly to the CTAs and round-robin to a CTA's four chain
// warps, so the slowest CTA carries at most one tile more than the others and small sets spread over all SMs.
//
// The tail of the evaluation runs in this launch too (optim_kernels.cuh: fused_tail): the last CTAs to finish reduce the per-CTA
// partial vectors in a fixed order, exchange them with the other ranks over NVLink (multi-GPU) and apply Adam (Adam steps).
//
// Measured and dropped (profiles/kernel_variants_r02.md): 3-slot ring, per-tile-pair ownership 3/2/2/2 (0.451 -> 0.416 ms
// with row ownership), holding the weight-gradient DMMAs back while the co-resident chain warp is in a DMMA phase
// (slower: the weight-gradient warps starve), publishing Z-bar after the chain warp's own input-adjoint GEMM (no change),
// output units 16..19 of the chain GEMMs as 80 FMAs + a quad reduce-scatter instead of the half-padded third DMMA tile
// (fewer pipe cycles on paper, 0.389 -> 0.416 ms in practice: DFMA interleaved with DMMA costs more than it saves); the same
// for columns 16..19 of the weight-gradient tile rows (0.390 -> 0.410 ms); spreading the weight-gradient DMMAs with __nanosleep
// between k-steps (0.443 ms at 20 ns and worse); and a THREE-warps-per-sub-partition variant that splits every tile's chain
// between two half-warps (N tiles {0,2} / {1}, A operands from shared memory, roles alternating per layer; 384 threads x 168
// registers): correct on the first run, but 0.439 ms -- the FP64 pipe was 70 % busy against 75 % here (pair barriers and the
// A-operand LDS cost more than the added latency hiding bought).
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
constexpr int ROUND = CHAINS * TILE;      // 32 points per full CTA round
#ifndef PINN_RING
#define PINN_RING 2
#endif
// timing ablations (they compute WRONG gradients on purpose; profiles/kernel_variants.py)
#ifndef PINN_ABL_NOWG
#define PINN_ABL_NOWG 0                   // weight-gradient warps only do the ring handshake
#endif
#ifndef PINN_ABL_NOACT
#define PINN_ABL_NOACT 0                  // no activation arithmetic in the chain warps
#endif
#ifndef PINN_ABL_NODMMA
#define PINN_ABL_NODMMA 0                 // no chain DMMAs
#endif
constexpr int RING = PINN_RING;           // Z-bar ring slots per chain warp

// shared memory carve-up (doubles)
constexpr int STASH0 = 640;               // layer 0: outputs (a, a_x, a_t, a_xx) as [32 rows][20], like the other layers
constexpr int STASHL = 640;               // layers 1..6: outputs (a, a_x, a_t, a_xx) as [32 rows][20]
constexpr int STASH_PER_WARP = STASH0 + 6 * STASHL;   // 4480
constexpr int SM_W = 0;
constexpr int SM_STASH = SM_W + WPAD;
constexpr int SM_RING = SM_STASH + CHAINS * STASH_PER_WARP;
constexpr int RED_PER_WARP = 128;         // chain-warp partials: [0..4] scalars, [8..27] dW8, [32..91] dW0 (x row, t row), db0
constexpr int SM_RED = SM_RING + CHAINS * RING * 640;
constexpr int SM_BAR = SM_RED + CHAINS * RED_PER_WARP;      // 1 + 2*CHAINS*RING mbarriers
constexpr int SM_SPECIAL = SM_BAR + 1 + 2 * CHAINS * RING + 1;   // [32 rows][W] page: column 0 = the bias ones-row (1 on the value
                                                                 // stream's rows 0..7), column 1 = 0
constexpr int SM_DOUBLES = SM_SPECIAL + 32 * W;
constexpr int SMEM_BYTES = SM_DOUBLES * 8;   // ~217 KB

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Ring bookkeeping: chain warp c numbers its weight-gradient tasks T = 7*tile + J (J = 7 - layer, layers 7..1; the
// gradients of layer 0 and of the output layer are plain FMAs in the chain warp itself).  Task T lives in ring
// slot T % RING; it is the (T / RING)-th use of that slot, i.e. phase T / RING of the slot's full and empty barriers.
__device__ __forceinline__ uint64_t* bar_full(uint64_t* bars, int c, int slot) { return bars + 1 + (c * RING + slot) * 2; }
__device__ __forceinline__ uint64_t* bar_empty(uint64_t* bars, int c, int slot) { return bars + 2 + (c * RING + slot) * 2; }
__device__ __forceinline__ void wait_produced(uint64_t* bars, int c, int T) { mbar_wait(bar_full(bars, c, T % RING), (T / RING) & 1); }
__device__ __forceinline__ void wait_consumed(uint64_t* bars, int c, int T) { mbar_wait(bar_empty(bars, c, T % RING), (T / RING) & 1); }

// ---------------------------------------------------------------------------------------------------
// chain warps: the V5 register layout
// ---------------------------------------------------------------------------------------------------
// hidden unit held in slot j of lane-quad position q (== the unit k-step j contracts for this lane)
__device__ __forceinline__ int unit5(int j, int q) { return j < 4 ? 8 * (j >> 1) + 2 * q + (j & 1) : (q < 2 ? 16 + 2 * q : 13 + 2 * q); }

// accumulator fragments (C layout) -> V5: slots 0..3 are the first two tiles, slot 4 gathers units 16..19 across the quad
__device__ __forceinline__ void c_to_v5(double (&V)[4][5], const double (&Z)[4][3][2], int lane) {
#pragma unroll
  for (int s = 0; s < 4; s++) {
    V[s][0] = Z[s][0][0]; V[s][1] = Z[s][0][1]; V[s][2] = Z[s][1][0]; V[s][3] = Z[s][1][1];
    const double odd = __shfl_sync(0xffffffffu, Z[s][2][1], (lane & ~3) | (lane & 1));   // units 17, 19 live in lanes q = 0, 1
    V[s][4] = (lane & 2) ? odd : Z[s][2][0];
  }
}

// Z[s][nt][:] += V_s * Bmat, Bmat[k][n] = Wl[k*ldk + n*ldn]  (forward: ldk = 20, ldn = 1; input adjoint: ldk = 1, ldn = 20,
// i.e. the transposed weight).  N columns >= 20 read as zero.  5 k-steps x 4 streams x 3 N tiles = 60 DMMA.
__device__ __forceinline__ void mma_layer5(double (&Z)[4][3][2], const double (&V)[4][5], const double* Wl, int ldk, int ldn,
                                           int lane) {
  const int g = lane >> 2, q = lane & 3;
#pragma unroll
  for (int ks = 0; ks < 5; ks++) {
    const int k = unit5(ks, q);
    double b[3];
#pragma unroll
    for (int nt = 0; nt < 3; nt++) {
      const int n = 8 * nt + g;
      b[nt] = (n < W) ? Wl[k * ldk + n * ldn] : 0.0;
    }
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int nt = 0; nt < 3; nt++) dmma(Z[s][nt], V[s][ks], b[nt]);
  }
}

// outputs of a tanh layer from its pre-activation streams (in place): (z, z_x, z_t, z_xx) -> (a, s z_x, s z_t, s (z_xx - 2 a z_x^2))
__device__ __forceinline__ void act_forward5(double (&V)[4][5]) {
  if (PINN_ABL_NOACT) return;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const double a = tanh_fast(V[0][j]);
    const double zx = V[1][j];
    const double s = fma(-a, a, 1.0);
    V[0][j] = a;
    V[1][j] = s * zx;
    V[2][j] = s * V[2][j];
    V[3][j] = s * fma(-2.0 * a * zx, zx, V[3][j]);
  }
}

// adjoints of the outputs (A, overwritten with the adjoints of the pre-activations) given the outputs H
__device__ __forceinline__ void act_backward5(double (&A)[4][5], const double (&H)[4][5]) {
  if (PINN_ABL_NOACT) return;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const double a = H[0][j], ax = H[1][j], at = H[2][j], axx = H[3][j];
    const double A0 = A[0][j], Ax = A[1][j], At = A[2][j], Axx = A[3][j];
    const double s = fma(-a, a, 1.0);
    const double u1 = fma(ax, Ax, at * At);
    const double u2 = fma(a, axx, ax * ax);
    double z = fma(-2.0 * a, u1, s * A0);
    z = fma(-2.0 * Axx, u2, z);
    A[0][j] = z;
    A[1][j] = fma(-4.0 * a * ax, Axx, s * Ax);
    A[2][j] = s * At;
    A[3][j] = s * Axx;
  }
}

// V5 <-> a staged [32 rows][20 units] tile T[row = 8 s + prow(g)][unit] (ld = 20): two 16-byte accesses for the first two
// N tiles and one 8-byte access for this lane's unit of 16..19 per stream; all bank-conflict-free with the prow() row order.
__device__ __forceinline__ void stage5(double* T, const double (&V)[4][5], int lane) {
  const int pg = prow(lane >> 2), q = lane & 3;
  const int u4 = unit5(4, q);
#pragma unroll
  for (int s = 0; s < 4; s++) {
    double* row = T + (8 * s + pg) * W;
    *reinterpret_cast<double2*>(row + 2 * q) = make_double2(V[s][0], V[s][1]);
    *reinterpret_cast<double2*>(row + 8 + 2 * q) = make_double2(V[s][2], V[s][3]);
    row[u4] = V[s][4];
  }
}
__device__ __forceinline__ void load5(double (&V)[4][5], const double* T, int lane) {
  const int pg = prow(lane >> 2), q = lane & 3;
  const int u4 = unit5(4, q);
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const double* row = T + (8 * s + pg) * W;
    const double2 c0 = *reinterpret_cast<const double2*>(row + 2 * q);
    const double2 c1 = *reinterpret_cast<const double2*>(row + 8 + 2 * q);
    V[s][0] = c0.x; V[s][1] = c0.y; V[s][2] = c1.x; V[s][3] = c1.y;
    V[s][4] = row[u4];
  }
}

// ---------------------------------------------------------------------------------------------------
// wgrad warps
// ---------------------------------------------------------------------------------------------------
// tile-row ownership, hidden layers 1..7: position k = (wg - l) & 3; k < 3 owns row mt = k, k == 3 rests.  7 layers x 3 rows over
// 4 warps is 6/5/5/5 rows (warp 3 rests only at layer 4); to level that, warp 3 hands the third N tile of its row at layers
// 1, 2, 3 to the warp that rests there (warps 0, 1, 2): 120/128/128/128 DMMA per chain tile instead of 120/120/120/144.
// mask: bit nt set = this warp accumulates tile (mt, nt).
struct Own { int mt, mask; };
__device__ __forceinline__ Own ownership(int l, int wg) {
  const int k = (wg - l) & 3;
  Own o;
  if (k < 3) { o.mt = k; o.mask = (wg == 3 && l <= 3) ? 3 : 7; }
  else if (l <= 3) { o.mt = (3 - l) & 3; o.mask = 4; }
  else { o.mt = 0; o.mask = 0; }
  return o;
}

// one weight-gradient task: accumulate this warp's tile row of layer L for one chain warp's 8-point tile.
// Rows of the staged operands: R = 8*stream + 4*(ks&1) + q for k-step ks (K = 32 rows = 8 k-steps); A[row][i] = input unit i
// of layer L (i == 20: the bias ones-row on the value stream), B[row][j] = Z-bar.
template <int L, int MASK>
__device__ __forceinline__ void wgrad_task(double (&acc)[3][2], int mt, const double* Aop, const double* ZB, const double* sm,
                                           double sc0, double sc1, int lane) {
  const int g = lane >> 2, q = lane & 3;
  const double* SP = sm + SM_SPECIAL;
  const double* Wsm = sm + SM_W;
  const int i = 8 * mt + g;
  const double* bp0 = ZB + g + q * W;
  const double* bp1 = ZB + 8 + g + q * W;
  const double* bp2 = (16 + g < W ? ZB + 16 + g : SP + 1) + q * W;
  const double* ap = (i < W ? Aop + i : (i == W ? SP : SP + 1)) + q * W;          // [32 rows][W] outputs of layer L-1
#pragma unroll
  for (int ks = 0; ks < 8; ks++) {
    const int st = ks >> 1;
    const int off = (8 * st + 4 * (ks & 1)) * W;              // physical row 8*stream + 4*(ks&1) (+ q, folded into the pointers)
    const double a = ap[off];
    if (MASK & 1) dmma(acc[0], a, bp0[off]);
    if (MASK & 2) dmma(acc[1], a, bp1[off]);
    if (MASK & 4) dmma(acc[2], a, bp2[off]);
  }
}

template <int L>
__device__ __forceinline__ void wgrad_layer(double (&acc)[3][2], int wg, int it, int half, int cnt, const double* sm,
                                            uint64_t* bars, double sc0, double sc1, int lane) {
  constexpr int J = 7 - L;                  // task index within a tile (layers 7..1)
  const int T = 7 * it + J;
  const int slot = T % RING;
  const Own o = ownership(L, wg);
#pragma unroll 1
  for (int c = 2 * half; c < 2 * half + 2; c++) {
    if (4 * it + c >= cnt) break;           // chain warp c has no tile in this round (the CTA's last round may be partial)
    wait_produced(bars, c, T);
    const double* stash = sm + SM_STASH + c * STASH_PER_WARP;
    const double* Aop = stash + (L - 1) * STASHL;         // outputs of layer L-1 (STASH0 == STASHL)
    const double* ZB = sm + SM_RING + (c * RING + slot) * 640;
    if (!PINN_ABL_NOWG) {                   // warp-uniform: the mask is a function of (layer, warp)
      if (o.mask == 7) wgrad_task<L, 7>(acc, o.mt, Aop, ZB, sm, sc0, sc1, lane);
      else if (o.mask == 3) wgrad_task<L, 3>(acc, o.mt, Aop, ZB, sm, sc0, sc1, lane);
      else if (o.mask == 4) wgrad_task<L, 4>(acc, o.mt, Aop, ZB, sm, sc0, sc1, lane);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty(bars, c, slot));
  }
}

template <int L>
__device__ __forceinline__ void wgrad_flush(const double (&acc)[3][2], int wg, double* outp, int lane) {
  const int g = lane >> 2, q = lane & 3;
  const Own o = ownership(L, wg);
  const int in_dim = W;
  const int i = 8 * o.mt + g;
#pragma unroll
  for (int nt = 0; nt < 3; nt++)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int j = 8 * nt + 2 * q + e;
      if (j < W && ((o.mask >> nt) & 1)) {
        if (i < in_dim) outp[woff(L) + i * W + j] = acc[nt][e];
        else if (i == in_dim) outp[boff(L) + j] = acc[nt][e];
      }
    }
}

__global__ void __launch_bounds__(THREADS, 1) fused_loss_grad(const Args p) {
  extern __shared__ __align__(16) double sm[];
  pdl_launch_dependents();                 // the tail kernel's blocks may be placed as SMs free up; they park in pdl_wait()
  if (p.run_flag && *p.run_flag != 0) return;
  double* Wsm = sm + SM_W;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BAR);   // [0] weights TMA; then (full, empty) per (chain, slot)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;

  for (int i = threadIdx.x; i < 32 * W; i += THREADS) sm[SM_SPECIAL + i] = (i % W == 0 && i / W < 8) ? 1.0 : 0.0;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1);
    for (int i = 0; i < CHAINS * RING; i++) {
      mbar_init(bars + 1 + 2 * i, 1);        // full: the producing chain warp's lane 0
      mbar_init(bars + 2 + 2 * i, 4);        // empty: lane 0 of each of the 4 wgrad warps
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bars, WPAD * 8);
    tma_bulk_g2s(Wsm, p.w, WPAD * 8, bars);
  }
  mbar_wait(bars, 0);

  const double sc0 = 2.0 / p.dx0, sc1 = 2.0 / p.dx1;
  // Tile-granular distribution: the ceil(n/8) tiles are dealt to the CTAs as evenly as possible (the first `rem` CTAs take one
  // more), a CTA's tiles j = 0..cnt-1 go to chain warp j & 3 in round j >> 2.  No CTA runs a whole extra round for a few left-over
  // tiles (N_f = 100 000: 84.5 tiles per CTA = 21.1 rounds instead of 22 for the slowest CTA), and small sets spread over all SMs
  // (2000 points: 1-2 tiles per CTA).
  const long long tiles_total = (p.n_total + TILE - 1) / TILE;
  const long long tbase = tiles_total / gridDim.x, trem = tiles_total - tbase * gridDim.x;
  const int cnt = (int)(tbase + (blockIdx.x < trem ? 1 : 0));                                  // tiles of this CTA
  const long long first_tile = tbase * blockIdx.x + (blockIdx.x < trem ? blockIdx.x : trem);
  const int my_rounds = (cnt + CHAINS - 1) / CHAINS;
  double* outp = p.partials + (size_t)blockIdx.x * PSTRIDE;

  if (warp >= CHAINS) {
    // =========================================== wgrad warps ===========================================
    const int wg = warp & 3;
    double a1[3][2], a2[3][2], a3[3][2], a4[3][2], a5[3][2], a6[3][2], a7[3][2];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
      for (int e = 0; e < 2; e++)
        a1[s][e] = a2[s][e] = a3[s][e] = a4[s][e] = a5[s][e] = a6[s][e] = a7[s][e] = 0.0;
#pragma unroll 1
    // The chain warps are consumed in two halves (chains 0,1 then chains 2,3).  Ring back-pressure then shifts the
    // halves by one phase: while one half runs its forward pass (which produces no weight-gradient work), the other
    // half runs its backward pass and keeps these warps -- and the FP64 pipe of every sub-partition -- busy.
    for (int it2 = 0; it2 < 2 * my_rounds; it2++) {
      const int it = it2 >> 1, half = it2 & 1;
      wgrad_layer<7>(a7, wg, it, half, cnt, sm, bars, sc0, sc1, lane);
      wgrad_layer<6>(a6, wg, it, half, cnt, sm, bars, sc0, sc1, lane);
      wgrad_layer<5>(a5, wg, it, half, cnt, sm, bars, sc0, sc1, lane);
      wgrad_layer<4>(a4, wg, it, half, cnt, sm, bars, sc0, sc1, lane);
      wgrad_layer<3>(a3, wg, it, half, cnt, sm, bars, sc0, sc1, lane);
      wgrad_layer<2>(a2, wg, it, half, cnt, sm, bars, sc0, sc1, lane);
      wgrad_layer<1>(a1, wg, it, half, cnt, sm, bars, sc0, sc1, lane);
    }
    wgrad_flush<1>(a1, wg, outp, lane);
    wgrad_flush<2>(a2, wg, outp, lane);
    wgrad_flush<3>(a3, wg, outp, lane);
    wgrad_flush<4>(a4, wg, outp, lane);
    wgrad_flush<5>(a5, wg, outp, lane);
    wgrad_flush<6>(a6, wg, outp, lane);
    wgrad_flush<7>(a7, wg, outp, lane);
  } else {
    // =========================================== chain warps ===========================================
    const int c = warp;
    const int n_tiles = cnt > c ? (cnt - c + CHAINS - 1) / CHAINS : 0;      // tiles j = c, c + 4, ... of this CTA
    double* stash = sm + SM_STASH + c * STASH_PER_WARP;
    const double l1 = p.ide ? Wsm[P_NET] : 1.0;
    const double kap = p.ide ? exp(Wsm[P_NET + 1]) : p.nu;
    double loss_d = 0.0, loss_f = 0.0, gl1 = 0.0, gl2 = 0.0;
    double g8[5], gb8 = 0.0;                // output-layer weight gradient: per-lane partials over this lane's points
    double g0x[5], g0t[5], g0b[5];          // layer-0 weight gradient (rows x^, t^ of W_0 and b_0): likewise
#pragma unroll
    for (int j = 0; j < 5; j++) g8[j] = g0x[j] = g0t[j] = g0b[j] = 0.0;
    const int pg = prow(g);
    int u5[5];                              // the five hidden units of this lane
#pragma unroll
    for (int j = 0; j < 5; j++) u5[j] = unit5(j, q);
    // coordinates of this lane's point in tile `it`; the next tile's pair is prefetched one tile ahead so that neither
    // the L2/HBM latency nor (zero-copy mode: p.xc in pinned host memory) the PCIe latency is exposed
    auto load_xt = [&](int it, double& xo, double& to) {
      long long pt = (first_tile + 4 * it + c) * TILE + g;
      if (pt >= p.n_total) pt = p.n_total - 1;
      if (p.xc && pt >= p.c0 && pt < p.c0 + p.n_c) {
        xo = __ldg(p.xc + (pt - p.c0)); to = __ldg(p.tc + (pt - p.c0));
      } else {
        xo = __ldg(p.x + pt); to = __ldg(p.t + pt);
      }
    };
    double xr_next = 0.0, tr_next = 0.0;
    if (n_tiles > 0) load_xt(0, xr_next, tr_next);

#pragma unroll 1
    for (int it = 0; it < n_tiles; it++) {
      const long long pt = (first_tile + 4 * it + c) * TILE + g;
      const bool in_set = pt < p.n_total;
      const double xr = xr_next, tr = tr_next;
      if (it + 1 < n_tiles) load_xt(it + 1, xr_next, tr_next);
      const double wf = (in_set && pt >= p.c0 && pt < p.c0 + p.n_c) ? p.wf : 0.0;
      const bool has_d = in_set && pt >= p.d0 && pt < p.d0 + p.n_d;
      const double wd = has_d ? p.wd : 0.0;
      const double ut = has_d ? __ldg(p.utgt + (pt - p.d0)) : 0.0;
      const double xh = 2.0 * (xr - p.lb0) / p.dx0 - 1.0;   // utils/neuralnetwork.py:29-30
      const double th = 2.0 * (tr - p.lb1) / p.dx1 - 1.0;

      double V[4][5];
      // ---------------- layer 0 (2 -> 20): direct
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const double w0 = Wsm[u5[j]], w1 = Wsm[W + u5[j]], b = Wsm[2 * W + u5[j]];
        V[0][j] = fma(xh, w0, fma(th, w1, b));
        V[1][j] = sc0 * w0;
        V[2][j] = sc1 * w1;
        V[3][j] = 0.0;
      }
      act_forward5(V);
      // the stash is still the A operand of the previous tile's weight-gradient tasks: wait until its last one, the layer-1
      // task (J = 6), has been consumed; consumers retire tasks in order
      if (it > 0) wait_consumed(bars, c, 7 * (it - 1) + 6);
      stage5(stash, V, lane);
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
        if (!PINN_ABL_NODMMA) mma_layer5(Z, V, Wl, W, 1, lane);
        c_to_v5(V, Z, lane);
        act_forward5(V);
        if (l < NHID - 1) stage5(stash + STASH0 + (l - 1) * STASHL, V, lane);
      }
      // ---------------- output layer (20 -> 1), residual, seeds
      double A[4][5];
      {
        const double* W8 = Wsm + woff(8);
        double w8[5];
#pragma unroll
        for (int j = 0; j < 5; j++) w8[j] = W8[u5[j]];
        double out[4];
#pragma unroll
        for (int s = 0; s < 4; s++) {
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < 5; j++) acc = fma(V[s][j], w8[j], acc);
          acc += shfl_xor_d(acc, 1);
          acc += shfl_xor_d(acc, 2);
          out[s] = acc;
        }
        out[0] += Wsm[boff(8)];
        const double u = out[0], ux = out[1], utt = out[2], uxx = out[3];
        const double f = utt + l1 * u * ux - kap * uxx;      // inf_cont_burgers.py:90 / ide_cont_burgers.py:85
        const double r = u - ut;
        const double cc = 2.0 * wf * f;
        double seed[4];
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
        for (int j = 0; j < 5; j++) {
          double acc = g8[j];
#pragma unroll
          for (int s = 0; s < 4; s++) {
            acc = fma(V[s][j], seed[s], acc);
            A[s][j] = seed[s] * w8[j];
          }
          g8[j] = acc;
        }
      }
      // ---------------- backward: layers 7..1 (task J = 7-l, ring slot J % RING)
#pragma unroll 1
      for (int l = NHID - 1; l >= 1; l--) {
        const int T = 7 * it + (7 - l), slot = T % RING;
        if (l < NHID - 1) load5(V, stash + STASH0 + (l - 1) * STASHL, lane);        // outputs of layer l (l = 7: registers)
        act_backward5(A, V);                                                         // A := Z-bar
        if (T >= RING) wait_consumed(bars, c, T - RING);                             // the slot's previous task is done
        stage5(sm + SM_RING + (c * RING + slot) * 640, A, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full(bars, c, slot));
        // adjoint of the layer inputs: A_new = Z-bar * W_l^T
        {
          const double* Wl = Wsm + woff(l);
          double An[4][3][2];
#pragma unroll
          for (int s = 0; s < 4; s++)
#pragma unroll
            for (int nt = 0; nt < 3; nt++) An[s][nt][0] = An[s][nt][1] = 0.0;
          if (!PINN_ABL_NODMMA) mma_layer5(An, A, Wl, 1, W, lane);
          c_to_v5(A, An, lane);
        }
      }
      // ---------------- backward: layer 0: its weight gradient is 3 rows (x^, t^, bias) -- 25 FMAs per lane here instead of 24
      // DMMAs in a weight-gradient warp
      {
        load5(V, stash, lane);
        act_backward5(A, V);
#pragma unroll
        for (int j = 0; j < 5; j++) {
          g0x[j] = fma(xh, A[0][j], fma(sc0, A[1][j], g0x[j]));   // inputs: (x^, t^) on the value stream, (sc0, 0) on the x
          g0t[j] = fma(th, A[0][j], fma(sc1, A[2][j], g0t[j]));   // stream, (0, sc1) on the t stream, 0 on the xx stream
          g0b[j] += A[0][j];
        }
      }
    }

    // ---------------- chain-warp partials: losses, identification gradients, output-layer gradient
    double* red = sm + SM_RED + c * RED_PER_WARP;
    loss_d = warp_sum(loss_d); loss_f = warp_sum(loss_f); gl1 = warp_sum(gl1); gl2 = warp_sum(gl2); gb8 = warp_sum(gb8);
#pragma unroll
    for (int j = 0; j < 5; j++) {
      double v = g8[j], vx = g0x[j], vt = g0t[j], vb = g0b[j];
      v += shfl_xor_d(v, 4); v += shfl_xor_d(v, 8); v += shfl_xor_d(v, 16);   // sum over the 8 points (g)
      vx += shfl_xor_d(vx, 4); vx += shfl_xor_d(vx, 8); vx += shfl_xor_d(vx, 16);
      vt += shfl_xor_d(vt, 4); vt += shfl_xor_d(vt, 8); vt += shfl_xor_d(vt, 16);
      vb += shfl_xor_d(vb, 4); vb += shfl_xor_d(vb, 8); vb += shfl_xor_d(vb, 16);
      if (g == 0) { red[8 + u5[j]] = v; red[32 + u5[j]] = vx; red[32 + W + u5[j]] = vt; red[32 + 2 * W + u5[j]] = vb; }
    }
    if (lane == 0) { red[0] = loss_d; red[1] = loss_f; red[2] = gl1; red[3] = gl2; red[4] = gb8; }
  }

  __syncthreads();
  {
    const double* red = sm + SM_RED;
    constexpr int R = RED_PER_WARP;
    const int t = threadIdx.x;
    auto sum4 = [&](int i) { return red[i] + red[R + i] + red[2 * R + i] + red[3 * R + i]; };   // fixed order: chains 0..3
    if (t < W) outp[woff(8) + t] = sum4(8 + t);
    if (t >= 64 && t < 64 + 3 * W) outp[woff(0) + (t - 64)] = sum4(32 + (t - 64));   // W_0 (2 x 20, row-major) then b_0: flat 0..59
    if (t == 32) outp[boff(8)] = sum4(4);
    if (t == 33) outp[IDX_LD] = sum4(0);
    if (t == 34) outp[IDX_LF] = sum4(1);
    if (t == 35) outp[IDX_DL1] = sum4(2);
    if (t == 36) outp[IDX_DL2] = sum4(3);
    if (t == 37) outp[3023] = 0.0;
  }
  if (p.tail.enabled) fused_tail(p.tail, p.partials, PSTRIDE, sm + SM_STASH);     // the stash is dead; all 256 threads arrive
}

}  // namespace burgers2
}  // namespace pinn
