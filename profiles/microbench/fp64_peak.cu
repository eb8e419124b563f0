// FP64 pipe microbenchmark for B200 (sm_100a): DFMA peak, DMMA.8x8x4 peak, dependent-chain
// latencies, and whether DFMA and DMMA issue concurrently.  Gives the roofline denominator for
// the fused PINN kernels (MEASURED_PEAKS.json carries no FP64 figure).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double x) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9 + i;
  double m = x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], m, 1e-30);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void dmma_kernel(double* out, int iters, double x) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = 0; c1[i] = 0; }
  double a = x * 1e-3 + threadIdx.x * 1e-12, b = x * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: per iteration NM DMMAs and NF DFMAs (independent chains)
template <int NM, int NF>
__global__ void mixed_kernel(double* out, int iters, double x) {
  double c0[NM + 1], c1[NM + 1], f[NF + 1];
#pragma unroll
  for (int i = 0; i < NM; i++) { c0[i] = 0; c1[i] = 0; }
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = i + threadIdx.x * 1e-9;
  double a = x * 1e-3 + threadIdx.x * 1e-12, b = x * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NM; i++) dmma(c0[i], c1[i], a, b);
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = fma(f[i], x, 1e-30);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; i++) s += c0[i] + c1[i];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP32 FFMA for context
template <int ILP>
__global__ void ffma_kernel(float* out, int iters, float x) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-9f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fmaf(acc[i], x, 1e-30f);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f}\n", prop.name, sms, clk_khz / 1e3);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 64 * 1024));
  const int iters = 4096;
  // sweep warps per SM for DFMA and DMMA at ILP 8
  int wps[] = {1, 2, 4, 8, 16, 32};
  for (int w : wps) {
    int threads = 32 * (w < 4 ? 1 : w / 4) ; int blocks_per_sm = w < 4 ? w : 4;
    int grid = sms * blocks_per_sm;
    float ms = time_ms([&] { dfma_kernel<8><<<grid, threads>>>(out, iters, 0.999); });
    double flops = 2.0 * grid * threads * 8.0 * iters;
    float ms2 = time_ms([&] { dmma_kernel<8><<<grid, threads>>>(out, iters, 0.999); });
    double flops2 = 2.0 * 256.0 * grid * (threads / 32) * 8.0 * iters;
    printf("{\"warps_per_sm\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f}\n", w, flops / ms / 1e9, flops2 / ms2 / 1e9);
  }
  // ILP sweep at 4 warps/SM (1 per SMSP): exposes latency
  {
    int grid = sms * 4, threads = 32;
    float a1 = time_ms([&] { dfma_kernel<1><<<grid, threads>>>(out, iters, 0.999); });
    float a2 = time_ms([&] { dfma_kernel<2><<<grid, threads>>>(out, iters, 0.999); });
    float a4 = time_ms([&] { dfma_kernel<4><<<grid, threads>>>(out, iters, 0.999); });
    float a16 = time_ms([&] { dfma_kernel<16><<<grid, threads>>>(out, iters, 0.999); });
    float m1 = time_ms([&] { dmma_kernel<1><<<grid, threads>>>(out, iters, 0.999); });
    float m2 = time_ms([&] { dmma_kernel<2><<<grid, threads>>>(out, iters, 0.999); });
    float m4 = time_ms([&] { dmma_kernel<4><<<grid, threads>>>(out, iters, 0.999); });
    float m12 = time_ms([&] { dmma_kernel<12><<<grid, threads>>>(out, iters, 0.999); });
    // ns per dependent op = ms*1e6/iters/ILP... report cycles at max clock
    double cyc = clk_khz * 1e3 * 1e-3;  // cycles per ms
    printf("{\"one_warp_per_smsp\": true, \"dfma_cyc_per_op\": {\"ilp1\": %.2f, \"ilp2\": %.2f, \"ilp4\": %.2f, \"ilp16\": %.2f}, "
           "\"dmma_cyc_per_op\": {\"ilp1\": %.2f, \"ilp2\": %.2f, \"ilp4\": %.2f, \"ilp12\": %.2f}}\n",
           a1 * cyc / iters / 1, a2 * cyc / iters / 2, a4 * cyc / iters / 4, a16 * cyc / iters / 16,
           m1 * cyc / iters / 1, m2 * cyc / iters / 2, m4 * cyc / iters / 4, m12 * cyc / iters / 12);
  }
  // mixed at 4 and 8 warps/SM
  for (int w : {4, 8, 16}) {
    int threads = 32 * (w / 4), grid = sms * 4;
    float mm = time_ms([&] { mixed_kernel<8, 0><<<grid, threads>>>(out, iters, 0.999); });
    float ff = time_ms([&] { mixed_kernel<0, 16><<<grid, threads>>>(out, iters, 0.999); });
    float mf = time_ms([&] { mixed_kernel<8, 16><<<grid, threads>>>(out, iters, 0.999); });
    printf("{\"warps_per_sm\": %d, \"ms_dmma8\": %.4f, \"ms_dfma16\": %.4f, \"ms_both\": %.4f, \"overlap\": %.2f}\n", w, mm, ff, mf,
           (mm + ff - mf) / (mm < ff ? mm : ff));
  }
  {
    int grid = sms * 4, threads = 256;
    float ms = time_ms([&] { ffma_kernel<8><<<grid, threads>>>((float*)out, iters, 0.999f); });
    printf("{\"ffma_tflops\": %.2f}\n", 2.0 * grid * threads * 8.0 * iters / ms / 1e9);
  }
  return 0;
}
