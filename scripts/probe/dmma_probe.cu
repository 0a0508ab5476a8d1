// FP64 throughput on B200: scalar DFMA against the DMMA shapes (m8n8k4, m16n8k4, m16n8k8, m16n8k16).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
template <int KIND>
__global__ void __launch_bounds__(256) probe(double* out, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  double b0 = seed * 0.5, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
  double c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // four independent accumulator chains
      if constexpr (KIND == 0) {
        c[i][0] = fma(a0, b0, c[i][0]); c[i][1] = fma(a1, b1, c[i][1]); c[i][2] = fma(a2, b2, c[i][2]); c[i][3] = fma(a3, b3, c[i][3]);
      } else if constexpr (KIND == 1) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a0), "d"(b0));
      } else if constexpr (KIND == 2) {
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b0));
      } else if constexpr (KIND == 3) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                     : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(b0), "d"(b1), "d"(b2), "d"(b3));
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int KIND>
void run(const char* name, double flops_per_warp_iter, double* out) {
  const int grid = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<KIND><<<grid, 256>>>(out, 1.0);
  cudaEventRecord(e0);
  probe<KIND><<<grid, 256>>>(out, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)grid * 8, insts = warps * ITERS * 4 * (KIND == 0 ? 4 : 1);
  printf("%-10s %8.3f ms  %7.2f TFLOP/s  %6.2f cycles per warp instruction per SM sub-partition (at 1.965 GHz)\n", name, ms,
         warps * ITERS * flops_per_warp_iter / (ms * 1e-3) * 1e-12, ms * 1e-3 * 1.965e9 * 148 * 4 / insts);
}
int main() {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  run<0>("DFMA", 4 * 4 * 64.0, out);
  run<1>("m8n8k4", 4 * 512.0, out);
  run<2>("m16n8k4", 4 * 1024.0, out);
  run<3>("m16n8k8", 4 * 2048.0, out);
  run<4>("m16n8k16", 4 * 4096.0, out);
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
