// Shared device/host helpers for libgraphbix_cuda.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <string>

namespace gbx {

#ifndef GBX_NT
#define GBX_NT 128
#endif
constexpr int NT = GBX_NT;         // threads per CTA == directed-edge slots per tile
constexpr int TILE_E = NT;         // a tile is a run of consecutive vertices with <= TILE_E incoming edges
constexpr int TILE_V = NT / 2;     // ... and <= TILE_V vertices
constexpr int MAXQ = 16;

// ---- error plumbing -------------------------------------------------------------------------
std::string& last_error();
#define GBX_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      char _buf[512];                                                                        \
      snprintf(_buf, sizeof _buf, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,              \
               cudaGetErrorString(_e));                                                      \
      gbx::last_error() = _buf;                                                              \
      return GBX_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

// ---- 64 B messages: 256-bit global accesses (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a) -----
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%4], {%0,%1,%2,%3};" ::"d"(a), "d"(b), "d"(c), "d"(d), "l"(p) : "memory");
}

template <int Q>
__device__ __forceinline__ void load_vec(const double* __restrict__ base, long long idx, double (&m)[Q]) {
  const double* p = base + idx * Q;
  if constexpr (Q % 4 == 0) {
#pragma unroll
    for (int t = 0; t < Q; t += 4) ldg256(p + t, m[t], m[t + 1], m[t + 2], m[t + 3]);
  } else if constexpr (Q % 2 == 0) {
#pragma unroll
    for (int t = 0; t < Q; t += 2) {
      double2 v = *reinterpret_cast<const double2*>(p + t);
      m[t] = v.x;
      m[t + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int t = 0; t < Q; ++t) m[t] = p[t];
  }
}

template <int Q>
__device__ __forceinline__ void store_vec(double* __restrict__ base, long long idx, const double (&m)[Q]) {
  double* p = base + idx * Q;
  if constexpr (Q % 4 == 0) {
#pragma unroll
    for (int t = 0; t < Q; t += 4) stg256(p + t, m[t], m[t + 1], m[t + 2], m[t + 3]);
  } else if constexpr (Q % 2 == 0) {
#pragma unroll
    for (int t = 0; t < Q; t += 2) *reinterpret_cast<double2*>(p + t) = make_double2(m[t], m[t + 1]);
  } else {
#pragma unroll
    for (int t = 0; t < Q; ++t) p[t] = m[t];
  }
}

// Message rows: Q components in a row of QM >= Q doubles (pads are stored as zeros, ignored on load).
template <int Q, int QM>
__device__ __forceinline__ void load_msg(const double* __restrict__ base, long long idx, double (&m)[Q]) {
  if constexpr (QM == Q) {
    load_vec<Q>(base, idx, m);
  } else {
    double t[QM];
    load_vec<QM>(base, idx, t);
#pragma unroll
    for (int k = 0; k < Q; ++k) m[k] = t[k];
  }
}
template <int Q, int QM>
__device__ __forceinline__ void store_msg(double* __restrict__ base, long long idx, const double (&m)[Q]) {
  if constexpr (QM == Q) {
    store_vec<Q>(base, idx, m);
  } else {
    double t[QM];
#pragma unroll
    for (int k = 0; k < QM; ++k) t[k] = (k < Q) ? m[k] : 0.0;
    store_vec<QM>(base, idx, t);
  }
}

// ---- power-of-two scaling (exact) -------------------------------------------------------------
// Unbiased exponent of a positive normal double; `ok` is cleared for 0, subnormal, Inf, NaN, negative.
__device__ __forceinline__ int exponent_of(double m, bool& ok) {
  const int hi = __double2hiint(m);
  const int ex = (hi >> 20) & 0x7ff;
  ok = (hi >= 0) && (ex != 0) && (ex != 0x7ff);
  return ex - 1023;
}
// 2^n for n in [-1022, 1023].
__device__ __forceinline__ double pow2i(int n) { return __hiloint2double((1023 + n) << 20, 0); }
// m * 2^n with m in [1,2): 0 below the normal range, capped at 2^1000 above (keeps sums of <= 16 finite).
__device__ __forceinline__ double scale2(double m, int n) {
  if (n < -1022) return 0.0;
  if (n > 1000) n = 1000;
  return m * pow2i(n);
}

// ---- deterministic block reduction of NV doubles per thread ----------------------------------
// Result (sum over the CTA, fixed order) is left in out[0..NV) of thread 0's view via smem `red`.
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* red /* [NT/32][NV] smem */,
                                                   double* dst /* global, NV */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp * NV + k] = x;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < NV; k += NT) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) s += red[w * NV + k];
    dst[k] = s;
  }
  __syncthreads();
}

}  // namespace gbx
