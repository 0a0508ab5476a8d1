// Memory-system probe for the BP sweep's access pattern on B200: streaming 64-B reads + 64-B writes that are
// (a) contiguous, (b) scattered through a random permutation (what rev[e] does on a random graph),
// (c) gathered reads + contiguous writes, (d) scattered inside 1-MB windows.  No arithmetic.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/scatter_probe scripts/probe/scatter_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
}

// mode 0: dst[i] = src[i]; 1: dst[perm[i]] = src[i]; 2: dst[i] = src[perm[i]]
template <int MODE>
__global__ void k(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ perm, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long r = i, w = i;
    if (MODE == 1) w = perm[i];
    if (MODE == 2) r = perm[i];
    double a0, a1, a2, a3, b0, b1, b2, b3;
    ld256(src + r * 8, a0, a1, a2, a3);
    ld256(src + r * 8 + 4, b0, b1, b2, b3);
    st256(dst + w * 8, a0, a1, a2, a3);
    st256(dst + w * 8 + 4, b0, b1, b2, b3);
  }
}

int main() {
  const long long n = 16000000;  // 64-B records: 1.024 GB per array
  double *src, *dst;
  int* perm;
  cudaMalloc(&src, n * 64);
  cudaMalloc(&dst, n * 64);
  cudaMalloc(&perm, n * 4);
  cudaMemset(src, 0, n * 64);
  std::vector<int> p(n);
  std::mt19937_64 rng(1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[] = {"copy contiguous", "scatter 64B writes (random perm)", "gather 64B reads (random perm)",
                         "scatter within 1 MB windows", "scatter within 64 KB windows"};
  for (int test = 0; test < 5; ++test) {
    std::iota(p.begin(), p.end(), 0);
    if (test == 1 || test == 2) std::shuffle(p.begin(), p.end(), rng);
    if (test >= 3) {
      const long long win = (test == 3 ? (1 << 20) : (1 << 16)) / 64;
      for (long long b = 0; b < n; b += win) std::shuffle(p.begin() + b, p.begin() + std::min(n, b + win), rng);
    }
    cudaMemcpy(perm, p.data(), n * 4, cudaMemcpyHostToDevice);
    for (int grid : {148 * 8, 148 * 16}) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        if (test == 0) k<0><<<grid, 256>>>(src, dst, perm, n);
        else if (test == 2) k<2><<<grid, 256>>>(src, dst, perm, n);
        else k<1><<<grid, 256>>>(src, dst, perm, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::min(best, ms);
      }
      const double bytes = (double)n * (64 + 64 + (test ? 4 : 0));
      printf("%-36s grid=%5d  %.3f ms  %.0f GB/s\n", names[test], grid, best, bytes / best * 1e-6);
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
