// DMMA (mma.sync.m8n8k4.f64) and DFMA issue rate on this GPU: `warps` warps per SM, 4096 dependent-free instructions each.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dmma_kernel(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  double c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
  for (int i = 0; i < iters; ++i) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
}
__global__ void dfma_kernel(double *out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-9;
  double c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) c[j] = fma(a, b, c[j]);
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out;
  cudaMalloc(&out, sizeof(double) * 148 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps : {4, 8, 16, 32}) {
    for (int which = 0; which < 2; ++which) {
      const int iters = 4096;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) dmma_kernel<<<148, warps * 32>>>(out, iters);
        else dfma_kernel<<<148, warps * 32>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = which == 0 ? 148.0 * warps * iters * 4 * 256 : 148.0 * warps * 32 * iters * 8.0;
      printf("{\"kernel\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"TFLOPs\": %.2f}\n", which == 0 ? "dmma m8n8k4" : "dfma", warps, ms, 2 * fma / (ms * 1e-3) / 1e12);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
