// Memory-system probe for the "pull" formulation of the BP sweep on B200: every directed-edge slot streams its
// own 64-B message row (read, then rewritten in place) and gathers a per-vertex record (64 / 80 / 96 B) of its
// neighbour from an array small enough to live in the 126 MB L2.  No arithmetic.  Compared with the random
// 64-B scatter of the "push" formulation (scatter_probe.cu: 3.3 TB/s).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/pull_probe scripts/probe/pull_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <bool HINT>
__device__ __forceinline__ double2 ld16(const double* p, unsigned long long pol) {
  double2 v;
  if (HINT) asm volatile("ld.global.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
  else asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
template <bool HINT>
__device__ __forceinline__ void ld32(const double* p, double& a, double& b, double& c, double& d, unsigned long long pol) {
  if (HINT)
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p), "l"(pol));
  else asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
template <bool HINT>
__device__ __forceinline__ void st32(double* p, double a, double b, double c, double d, unsigned long long pol) {
  if (HINT)
    asm volatile("st.global.L2::cache_hint.v4.f64 [%4], {%0,%1,%2,%3}, %5;" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p), "l"(pol) : "memory");
  else asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
}

// RD = doubles per gathered record (8, 10 or 12); messages are 8 doubles, updated in place.
template <int RD, bool HINT>
__global__ void __launch_bounds__(256) k_pull(double* __restrict__ msg, const double* __restrict__ rec, const int* __restrict__ col,
                                              long long K) {
  const unsigned long long pf = policy_evict_first(), pl = policy_evict_last();
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const long long c = col[e];
    double a0, a1, a2, a3, b0, b1, b2, b3;
    ld32<HINT>(msg + e * 8, a0, a1, a2, a3, pf);
    ld32<HINT>(msg + e * 8 + 4, b0, b1, b2, b3, pf);
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < RD / 2; ++u) {
      const double2 v = ld16<HINT>(rec + c * RD + 2 * u, pl);
      s += v.x + v.y;
    }
    st32<HINT>(msg + e * 8, a0 + s, a1, a2, a3, pf);
    st32<HINT>(msg + e * 8 + 4, b0, b1, b2, b3, pf);
  }
}

// the push pattern for reference: contiguous read, random 64-B write
__global__ void __launch_bounds__(256) k_push(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ perm, long long K) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const long long w = perm[e];
    double a0, a1, a2, a3, b0, b1, b2, b3;
    ld32<false>(src + e * 8, a0, a1, a2, a3, 0ull);
    ld32<false>(src + e * 8 + 4, b0, b1, b2, b3, 0ull);
    st32<false>(dst + w * 8, a0, a1, a2, a3, 0ull);
    st32<false>(dst + w * 8 + 4, b0, b1, b2, b3, 0ull);
  }
}

template <int RD, bool HINT>
float run(double* msg, const double* rec, const int* col, long long K, int grid) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k_pull<RD, HINT><<<grid, 256>>>(msg, rec, col, K);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}

int main() {
  const long long K = 16000000;  // directed-edge slots: 1.024 GB of messages
  const long long NMAX = 8000000;
  double *msg, *msg2, *rec;
  int* col;
  cudaMalloc(&msg, K * 64);
  cudaMalloc(&msg2, K * 64);
  cudaMalloc(&rec, NMAX * 96);
  cudaMalloc(&col, K * 4);
  cudaMemset(msg, 0, K * 64);
  cudaMemset(msg2, 0, K * 64);
  cudaMemset(rec, 0, NMAX * 96);
  std::vector<int> c(K);
  std::mt19937_64 rng(1);
  {
    std::iota(c.begin(), c.end(), 0);
    std::shuffle(c.begin(), c.end(), rng);
    cudaMemcpy(col, c.data(), K * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
      cudaEventRecord(e0);
      k_push<<<148 * 16, 256>>>(msg, msg2, col, K);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0) best = std::min(best, ms);
    }
    printf("push: contiguous read + random 64B write            %.3f ms  %.0f GB/s (132 B/slot)\n", best, K * 132.0 / best * 1e-6);
  }
  for (long long N : {1000000LL, 2000000LL, 4000000LL, 8000000LL}) {
    for (int locality = 0; locality < 2; ++locality) {
      // locality 0: neighbours uniform over all vertices; 1: 7 of 8 neighbours inside the vertex's own eighth of the ids
      // (a planted partition with vertices numbered by group)
      std::uniform_int_distribution<long long> any(0, N - 1);
      const long long deg = K / N, blockn = N / 8;
      for (long long e = 0; e < K; ++e) {
        const long long v = std::min(e / deg, N - 1);
        if (locality && (rng() & 7)) c[e] = (int)((v / blockn) * blockn + (long long)(rng() % (unsigned long long)blockn));
        else c[e] = (int)any(rng);
      }
      cudaMemcpy(col, c.data(), K * 4, cudaMemcpyHostToDevice);
      for (int grid : {148 * 8, 148 * 16}) {
        const float t8 = run<8, false>(msg, rec, col, K, grid), t8h = run<8, true>(msg, rec, col, K, grid);
        const float t10 = run<10, false>(msg, rec, col, K, grid), t10h = run<10, true>(msg, rec, col, K, grid);
        const float t12h = run<12, true>(msg, rec, col, K, grid);
        printf("pull N=%lldM loc=%d grid=%d | rec64 %.3f ms, hints %.3f ms | rec80 %.3f ms, hints %.3f ms (%.0f GB/s of 132 B/slot) | rec96 hints %.3f ms\n",
               N / 1000000, locality, grid, t8, t8h, t10, t10h, K * 132.0 / t10h * 1e-6, t12h);
      }
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
