// Probe of a "bucketed transpose" message layout for the BP sweep on a random graph:
// processing is in S steps; at step k the kernel READS records gathered at random inside window k (W records)
// and WRITES each record to a random one of the S chunks (s, k) of size W/S -- chunk (s,k) belongs to window s
// and is filled completely during step k.  Compare with the fully random permutation of scatter_probe.cu.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/bucket_probe scripts/probe/bucket_probe.cu
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
__global__ void k(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ rperm,
                  const int* __restrict__ wperm, long long n) {
  // grid-stride in index order: all CTAs work on the same neighbourhood of i at the same time (like the tile loop)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = rperm[i], w = wperm[i];
    double a0, a1, a2, a3, b0, b1, b2, b3;
    ld256(src + r * 8, a0, a1, a2, a3);
    ld256(src + r * 8 + 4, b0, b1, b2, b3);
    st256(dst + w * 8, a0, a1, a2, a3);
    st256(dst + w * 8 + 4, b0, b1, b2, b3);
  }
}

int main() {
  const long long n = 16000000;
  double *src, *dst;
  int *rp, *wp;
  cudaMalloc(&src, n * 64);
  cudaMalloc(&dst, n * 64);
  cudaMalloc(&rp, n * 4);
  cudaMalloc(&wp, n * 4);
  cudaMemset(src, 0, n * 64);
  std::vector<int> r(n), w(n);
  std::mt19937_64 rng(1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int S : {32, 141, 512, 2048}) {
    const long long W = (n + S - 1) / S;       // records per window
    const long long C = (W + S - 1) / S;       // records per chunk
    // read side: random inside window k
    std::iota(r.begin(), r.end(), 0);
    for (long long b = 0; b < n; b += W) std::shuffle(r.begin() + b, r.begin() + std::min(n, b + W), rng);
    // write side: record i (step k = i / W) goes to window s = random, chunk k, random free position in the chunk
    std::vector<long long> fill((size_t)S * S, 0);
    std::vector<int> order(n);
    for (long long k = 0; k < S; ++k) {
      const long long lo = k * W, hi = std::min(n, (k + 1) * W);
      // deal the step's records to the S chunks evenly, in random order
      std::vector<int> idx(hi - lo);
      std::iota(idx.begin(), idx.end(), 0);
      std::shuffle(idx.begin(), idx.end(), rng);
      for (long long t = 0; t < hi - lo; ++t) {
        const long long s = t % S;                          // destination window
        const long long pos = s * W + k * C + fill[s * S + k]++;
        w[lo + idx[t]] = (int)std::min(pos, n - 1);
      }
    }
    // make w a permutation (clamp collisions at the tail are rare; fix by checking)
    std::vector<char> used(n, 0);
    long long coll = 0;
    for (long long i = 0; i < n; ++i) { if (used[w[i]]) ++coll; used[w[i]] = 1; }
    cudaMemcpy(rp, r.data(), n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(wp, w.data(), n * 4, cudaMemcpyHostToDevice);
    for (int grid : {148 * 6, 148 * 16}) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, 128>>>(src, dst, rp, wp, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::min(best, ms);
      }
      printf("S=%5d window=%7.2f MB chunk=%7.1f KB collisions=%lld grid=%5d  %.3f ms  %.0f GB/s\n", S, W * 64 / 1e6,
             C * 64 / 1e3, coll, grid, best, (double)n * 136 / best * 1e-6);
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
