// NVLink probe for the partitioned sweep on a 2 x B200 box: how fast can GPU 0 put 64-B messages into GPU 1's HBM?
//   (a) cudaMemcpyPeerAsync of one contiguous block            (copy engine)
//   (b) kernel, contiguous 16 B / thread stores to the peer
//   (c) kernel, 64-B rows at random peer positions, one thread per row (2 x st.v4.f64 = what stage 3 does)
//   (d) kernel, 64-B rows at random peer positions, 4 threads x 16 B per row
//   (e) kernel, 128-B rows at random positions, 8 threads x 16 B
//   (f) pack locally (random gather -> contiguous local buffer), for the pack + bulk copy + unpack alternative
//   (g) unpack locally (contiguous read -> random 64-B scatter)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/probe/p2p_probe scripts/probe/p2p_probe.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void ld256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
}

__global__ void k_contig16(const double2* __restrict__ src, double2* __restrict__ dst, long long n16) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
// one thread per 64-B row
__global__ void k_row1(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ perm, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long w = perm[i];
    double a0, a1, a2, a3, b0, b1, b2, b3;
    ld256(src + i * 8, a0, a1, a2, a3);
    ld256(src + i * 8 + 4, b0, b1, b2, b3);
    st256(dst + w * 8, a0, a1, a2, a3);
    st256(dst + w * 8 + 4, b0, b1, b2, b3);
  }
}
// T threads x 16 B per row of T*16 bytes
template <int T>
__global__ void k_rowT(const double2* __restrict__ src, double2* __restrict__ dst, const int* __restrict__ perm, long long nrows) {
  const long long total = nrows * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / T;
    const int c = (int)(i % T);
    dst[(long long)perm[r] * T + c] = src[i];
  }
}
// gather rows perm[i] -> contiguous (pack)
__global__ void k_pack(const double2* __restrict__ src, double2* __restrict__ dst, const int* __restrict__ perm, long long nrows) {
  const long long total = nrows * 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[(long long)perm[i >> 2] * 4 + (i & 3)];
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  printf("peer access 0->1: %d\n", can);
  const long long n = 8000000;  // 64-B rows: 512 MB (the boundary traffic of a 16M-slot rank at a 50%% cut)
  double *src, *loc, *peer;
  int* perm;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&peer, n * 64));
  CK(cudaMemset(peer, 0, n * 64));
  CK(cudaDeviceEnablePeerAccess(0, 0));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&src, n * 64));
  CK(cudaMalloc(&loc, n * 64));
  CK(cudaMalloc(&perm, n * 4));
  CK(cudaMemset(src, 0, n * 64));
  std::vector<int> p(n);
  std::iota(p.begin(), p.end(), 0);
  std::mt19937_64 rng(1);
  std::shuffle(p.begin(), p.end(), rng);
  CK(cudaMemcpy(perm, p.data(), n * 4, cudaMemcpyHostToDevice));
  // random 128-B rows: permutation of n/2
  std::vector<int> p2(n / 2);
  std::iota(p2.begin(), p2.end(), 0);
  std::shuffle(p2.begin(), p2.end(), rng);
  int* perm2;
  CK(cudaMalloc(&perm2, n / 2 * 4));
  CK(cudaMemcpy(perm2, p2.data(), n / 2 * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double bytes = (double)n * 64;
  auto report = [&](const char* name, float ms) { printf("%-58s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms * 1e-6); };
  auto timeit = [&](auto&& f) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      f();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
    }
    return best;
  };
  report("(a) cudaMemcpyPeerAsync contiguous", timeit([&] { cudaMemcpyPeerAsync(peer, 1, src, 0, n * 64); }));
  for (int grid : {148 * 4, 148 * 16}) {
    printf("grid %d x 256\n", grid);
    report("(b) kernel contiguous 16 B/thread -> peer", timeit([&] { k_contig16<<<grid, 256>>>((double2*)src, (double2*)peer, n * 4); }));
    report("(b') kernel contiguous 16 B/thread -> local", timeit([&] { k_contig16<<<grid, 256>>>((double2*)src, (double2*)loc, n * 4); }));
    report("(c) random 64-B rows, 1 thread/row (2 x v4.f64) -> peer", timeit([&] { k_row1<<<grid, 256>>>(src, peer, perm, n); }));
    report("(c') same -> local", timeit([&] { k_row1<<<grid, 256>>>(src, loc, perm, n); }));
    report("(d) random 64-B rows, 4 threads x 16 B -> peer", timeit([&] { k_rowT<4><<<grid, 256>>>((double2*)src, (double2*)peer, perm, n); }));
    report("(e) random 128-B rows, 8 threads x 16 B -> peer", timeit([&] { k_rowT<8><<<grid, 256>>>((double2*)src, (double2*)peer, perm2, n / 2); }));
    report("(f) pack: random 64-B gather -> contiguous local", timeit([&] { k_pack<<<grid, 256>>>((double2*)src, (double2*)loc, perm, n); }));
    report("(g) unpack: contiguous -> random 64-B scatter local", timeit([&] { k_rowT<4><<<grid, 256>>>((double2*)src, (double2*)loc, perm, n); }));
  }
  {
    // (k) rows in order, (l) runs of 64 rows (4 KB, one tile's boundary messages) at random places, rows in order inside
    std::vector<int> pk(n), pl(n);
    std::iota(pk.begin(), pk.end(), 0);
    const int RUN = 64;
    std::vector<int> runs(n / RUN);
    std::iota(runs.begin(), runs.end(), 0);
    std::shuffle(runs.begin(), runs.end(), rng);
    for (long long i = 0; i < n; ++i) pl[i] = runs[i / RUN] * RUN + (int)(i % RUN);
    int *permk, *perml;
    CK(cudaMalloc(&permk, n * 4));
    CK(cudaMalloc(&perml, n * 4));
    CK(cudaMemcpy(permk, pk.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(perml, pl.data(), n * 4, cudaMemcpyHostToDevice));
    report("(k) rows in order, 1 thread/row (2 x v4.f64) -> peer", timeit([&] { k_row1<<<148 * 16, 256>>>(src, peer, permk, n); }));
    report("(l) 4-KB runs at random places, 1 thread/row -> peer", timeit([&] { k_row1<<<148 * 16, 256>>>(src, peer, perml, n); }));
    report("(k') rows in order, 1 thread/row -> local", timeit([&] { k_row1<<<148 * 16, 256>>>(src, loc, permk, n); }));
  }
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  // both directions at once (what a sweep does): GPU1 -> GPU0 concurrently
  double *src1, *peer0;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&src1, n * 64));
  CK(cudaMemset(src1, 0, n * 64));
  int* perm1;
  CK(cudaMalloc(&perm1, n * 4));
  CK(cudaMemcpy(perm1, p.data(), n * 4, cudaMemcpyHostToDevice));
  peer0 = loc;  // GPU 0's buffer
  cudaStream_t s1;
  CK(cudaStreamCreate(&s1));
  CK(cudaSetDevice(0));
  for (int variant = 0; variant < 3; ++variant) {
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaSetDevice(1));
      CK(cudaDeviceSynchronize());
      CK(cudaSetDevice(0));
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      if (variant == 0) k_row1<<<148 * 16, 256>>>(src, peer, perm, n);
      if (variant == 1) k_contig16<<<148 * 16, 256>>>((double2*)src, (double2*)peer, n * 4);
      if (variant == 2) cudaMemcpyPeerAsync(peer, 1, src, 0, n * 64);
      cudaEventRecord(e1);
      CK(cudaSetDevice(1));
      if (variant == 0) k_row1<<<148 * 16, 256, 0, s1>>>(src1, peer0, perm1, n);
      if (variant == 1) k_contig16<<<148 * 16, 256, 0, s1>>>((double2*)src1, (double2*)peer0, n * 4);
      if (variant == 2) cudaMemcpyPeerAsync(peer0, 0, src1, 1, n * 64, s1);
      CK(cudaSetDevice(0));
      cudaEventSynchronize(e1);
      CK(cudaSetDevice(1));
      CK(cudaDeviceSynchronize());
      CK(cudaSetDevice(0));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
    }
    const char* names[] = {"(h) bidirectional: random 64-B rows 1 thread/row", "(i) bidirectional: contiguous kernel",
                           "(j) bidirectional: cudaMemcpyPeerAsync"};
    report(names[variant], best);
  }
  return 0;
}
