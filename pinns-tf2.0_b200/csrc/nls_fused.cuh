// Fused NLS (Schrodinger) kernel -- placeholder until the [2,100x4,2] kernel lands.
#pragma once
#include "pinn_common.cuh"
namespace pinn { namespace nls {
constexpr int THREADS = 128;
constexpr int SMEM_BYTES = 1024;
constexpr int PSTRIDE = 30816;
inline int grid_size(int n_sm) { return n_sm; }
__global__ void fused_loss_grad() {}
}}
