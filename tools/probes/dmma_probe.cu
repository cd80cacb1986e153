// dmma_probe.cu -- micro-benchmark behind DESIGN.md's "dense KKT tiles on the FP64 tensor pipe" decision.
// Compares, per SM and clock, (a) the dense tile step of admm_multi_kernel as it is today (LDS.64 coefficient +
// LDS.128 broadcast operand + 2 DFMA for two instances per warp) with (b) mma.sync.m8n8k4.f64 with instances as the N
// dimension (8 instances per warp share each coefficient fragment), operands from shared memory, and (c) the bare DMMA
// issue rate / dependent-issue latency from registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// (a) today's dense step: per step one coefficient per lane, one 16-byte operand pair (same address for the whole warp)
template <int STEPS>
__global__ void k_dfma_dense(const double* __restrict__ gcoef, double* out, int iters, long long* clk) {
  extern __shared__ double sm[];
  double* coef = sm;                       // STEPS * 32
  double* w2 = sm + STEPS * 32 + (threadIdx.x >> 5) * 2 * 512;   // per warp 512 pairs
  for (int i = threadIdx.x; i < STEPS * 32; i += blockDim.x) coef[i] = gcoef[i];
  for (int i = threadIdx.x & 31; i < 1024; i += 32) w2[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s = 0; s < STEPS; s += 2) {
      const double v0 = coef[s * 32 + lane], v1 = coef[(s + 1) * 32 + lane];
      const double2 p0 = *reinterpret_cast<const double2*>(w2 + 2 * (s & 511));
      const double2 p1 = *reinterpret_cast<const double2*>(w2 + 2 * ((s + 1) & 511));
      a0 = fma(v0, p0.x, a0); a1 = fma(v0, p0.y, a1);
      b0 = fma(v1, p1.x, b0); b1 = fma(v1, p1.y, b1);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + b0 + b1;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// (b) DMMA, fragments from shared memory: A = 8 rows x 4 cols of coefficients (32 consecutive doubles),
//     B = 4 operand positions x 8 instances (32 consecutive doubles: w8[pos][inst]); RB row blocks share each B fragment
template <int KB, int RB>
__global__ void k_dmma_smem(const double* __restrict__ gcoef, double* out, int iters, long long* clk) {
  extern __shared__ double sm[];
  double* coef = sm;                       // KB * RB * 32
  double* w8 = sm + KB * RB * 32 + (threadIdx.x >> 5) * 8 * 352;  // per warp 352 positions x 8 instances
  for (int i = threadIdx.x; i < KB * RB * 32; i += blockDim.x) coef[i] = gcoef[i];
  for (int i = threadIdx.x & 31; i < 8 * 352; i += 32) w8[i] = 1e-3 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int boff = (lane & 3) * 8 + (lane >> 2);      // B[k = lane%4][n = lane/4] of w8[pos0 + k][n]
  double c[RB][2];
#pragma unroll
  for (int r = 0; r < RB; ++r) c[r][0] = c[r][1] = 0.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      const double b = w8[(k % 88) * 32 + boff];
#pragma unroll
      for (int r = 0; r < RB; ++r) dmma(c[r][0], c[r][1], coef[(k * RB + r) * 32 + lane], b);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int r = 0; r < RB; ++r) s += c[r][0] + c[r][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// (c) bare DMMA from registers with CH independent accumulator chains
template <int CH>
__global__ void k_dmma_reg(double* out, int iters, long long* clk) {
  double c[CH][2];
#pragma unroll
  for (int r = 0; r < CH; ++r) c[r][0] = c[r][1] = 0.0;
  const double a = 1e-3 * threadIdx.x, b = 1e-3 * (threadIdx.x & 7);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int r = 0; r < CH; ++r) dmma(c[r][0], c[r][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int r = 0; r < CH; ++r) s += c[r][0] + c[r][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// (d) bare DFMA from registers (FP64 vector pipe issue rate, for the same table)
template <int CH>
__global__ void k_dfma_reg(double* out, int iters, long long* clk) {
  double c[CH];
#pragma unroll
  for (int r = 0; r < CH; ++r) c[r] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int r = 0; r < CH; ++r) c[r] = fma(c[r], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int r = 0; r < CH; ++r) s += c[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static double avg_clk(long long* d, int n) {
  static long long h[1024];
  cudaMemcpy(h, d, n * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0; for (int i = 0; i < n; ++i) s += (double)h[i];
  return s / n;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, nsm);
  double *gcoef, *out; long long* clk;
  cudaMalloc(&gcoef, 8 * 1024 * 32 * 8); cudaMalloc(&out, 8 * 1024 * 1024); cudaMalloc(&clk, 1024 * 8);
  { static double h[8 * 1024 * 32]; for (int i = 0; i < 8 * 1024 * 32; ++i) h[i] = 1e-4 * (i % 97); cudaMemcpy(gcoef, h, sizeof(h), cudaMemcpyHostToDevice); }
  const int iters = 2000;
  const int warps_list[] = {1, 2, 4, 8, 12, 16};
  for (int w : warps_list) {
    { constexpr int S = 256; size_t sm = (S * 32 + w * 1024) * 8;
      cudaFuncSetAttribute(k_dfma_dense<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      k_dfma_dense<S><<<nsm, w * 32, sm>>>(gcoef, out, 10, clk); cudaDeviceSynchronize();
      k_dfma_dense<S><<<nsm, w * 32, sm>>>(gcoef, out, iters, clk);
      cudaError_t e = cudaDeviceSynchronize(); double c = avg_clk(clk, nsm);
      printf("{\"kernel\": \"dfma_dense_today\", \"warps\": %d, \"err\": %d, \"clk_per_step_per_sm\": %.3f, \"fma_per_clk_per_sm\": %.2f}\n", w, (int)e,
             c / ((double)iters * S * w), (double)iters * S * w * 64 / c); }
    { constexpr int KB = 64, RB = 4; size_t sm = (KB * RB * 32 + w * 8 * 352) * 8;
      cudaFuncSetAttribute(k_dmma_smem<KB, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      k_dmma_smem<KB, RB><<<nsm, w * 32, sm>>>(gcoef, out, 10, clk); cudaDeviceSynchronize();
      k_dmma_smem<KB, RB><<<nsm, w * 32, sm>>>(gcoef, out, iters, clk);
      cudaError_t e = cudaDeviceSynchronize(); double c = avg_clk(clk, nsm);
      printf("{\"kernel\": \"dmma_smem_rb4\", \"warps\": %d, \"err\": %d, \"clk_per_dmma_per_sm\": %.3f, \"fma_per_clk_per_sm\": %.2f}\n", w, (int)e,
             c / ((double)iters * KB * RB * w), (double)iters * KB * RB * w * 256 / c); }
    { constexpr int KB = 128, RB = 1; size_t sm = (KB * RB * 32 + w * 8 * 352) * 8;
      cudaFuncSetAttribute(k_dmma_smem<KB, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
      k_dmma_smem<KB, RB><<<nsm, w * 32, sm>>>(gcoef, out, 10, clk); cudaDeviceSynchronize();
      k_dmma_smem<KB, RB><<<nsm, w * 32, sm>>>(gcoef, out, iters, clk);
      cudaError_t e = cudaDeviceSynchronize(); double c = avg_clk(clk, nsm);
      printf("{\"kernel\": \"dmma_smem_rb1_chain\", \"warps\": %d, \"err\": %d, \"clk_per_dmma_per_sm\": %.3f, \"fma_per_clk_per_sm\": %.2f}\n", w, (int)e,
             c / ((double)iters * KB * RB * w), (double)iters * KB * RB * w * 256 / c); }
    { k_dmma_reg<1><<<nsm, w * 32>>>(out, iters, clk); cudaError_t e = cudaDeviceSynchronize(); double c = avg_clk(clk, nsm);
      printf("{\"kernel\": \"dmma_reg_chain1\", \"warps\": %d, \"err\": %d, \"clk_per_dmma_per_warp\": %.3f, \"fma_per_clk_per_sm\": %.2f}\n", w, (int)e,
             c / ((double)iters * 16), (double)iters * 16 * w * 256 / c); }
    { k_dmma_reg<8><<<nsm, w * 32>>>(out, iters, clk); cudaError_t e = cudaDeviceSynchronize(); double c = avg_clk(clk, nsm);
      printf("{\"kernel\": \"dmma_reg_chain8\", \"warps\": %d, \"err\": %d, \"clk_per_dmma_per_warp\": %.3f, \"fma_per_clk_per_sm\": %.2f}\n", w, (int)e,
             c / ((double)iters * 16 * 8), (double)iters * 16 * 8 * w * 256 / c); }
    { k_dfma_reg<1><<<nsm, w * 32>>>(out, iters, clk); cudaError_t e = cudaDeviceSynchronize(); double c = avg_clk(clk, nsm);
      printf("{\"kernel\": \"dfma_reg_chain1\", \"warps\": %d, \"err\": %d, \"clk_per_dfma_per_warp\": %.3f, \"fma_per_clk_per_sm\": %.2f}\n", w, (int)e,
             c / ((double)iters * 16), (double)iters * 16 * w * 32 / c); }
    { k_dfma_reg<8><<<nsm, w * 32>>>(out, iters, clk); cudaError_t e = cudaDeviceSynchronize(); double c = avg_clk(clk, nsm);
      printf("{\"kernel\": \"dfma_reg_chain8\", \"warps\": %d, \"err\": %d, \"clk_per_dfma_per_warp\": %.3f, \"fma_per_clk_per_sm\": %.2f}\n", w, (int)e,
             c / ((double)iters * 16 * 8), (double)iters * 16 * 8 * w * 32 / c); }
  }
  return 0;
}
