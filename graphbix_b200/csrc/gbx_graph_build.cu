// Device-side construction of the graph structures every gbx_* handle needs, from the reference's `links`
// matrix (K x 2 Int64, column-major, 1-based, both directions of every edge; what `simplegraph` returns,
// sbm.jl:414-425).  The reference builds `A`, `nb` and `inds` for the same purpose (sbm.jl:265-275).
//
//   slot e of row i  <->  message col[e] -> i          (destination-major CSR, rows sorted by source)
//   rev[e]            =   slot of the opposite message
//   erow[e]           =   i (row of the slot)
//   slot_of_link[k]   =   slot of links[k]              (import / export of messages in `links` order)
//   und[u]            =   (e, rev[e]) with rev[e] < e   (every undirected edge once, in slot order)
//
// One 64-bit radix sort (CUB) of (destination, source) keys; everything else is a map or a scan.
#include <cub/cub.cuh>

#include <vector>

#include "gbx_sbm_internal.h"

namespace gbx {
namespace {

enum { ERR_RANGE = 1, ERR_SELF = 2, ERR_DUP = 4, ERR_NOREV = 8 };

__global__ void k_keys(const int64_t* __restrict__ links, long long K, long long n, unsigned long long* keys,
                       int* vals, int* deg, int* err) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < K; k += (long long)gridDim.x * blockDim.x) {
    const long long s = links[k], d = links[K + k];
    if (s < 1 || s > n || d < 1 || d > n) {
      atomicOr(err, ERR_RANGE);
      keys[k] = ~0ull;
      vals[k] = (int)k;
      continue;
    }
    if (s == d) atomicOr(err, ERR_SELF);
    keys[k] = ((unsigned long long)(d - 1) << 32) | (unsigned long long)(s - 1);
    vals[k] = (int)k;
    atomicAdd(&deg[d - 1], 1);
  }
}

__global__ void k_slots(const unsigned long long* __restrict__ keys, const int* __restrict__ vals, long long K,
                        int* col, int* erow, int* slot_of_link, int* err) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const unsigned long long key = keys[e];
    col[e] = (int)(key & 0xffffffffull);
    erow[e] = (int)(key >> 32);
    slot_of_link[vals[e]] = (int)e;
    if (e > 0 && keys[e - 1] == key) atomicOr(err, ERR_DUP);
  }
}

__global__ void k_rev(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ erow,
                      long long K, int* rev, int* err) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const int i = erow[e], j = col[e];
    int lo = rowptr[j], hi = rowptr[j + 1];
    const int end = hi;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (col[mid] < i) lo = mid + 1; else hi = mid;
    }
    if (lo >= end || col[lo] != i) {
      atomicOr(err, ERR_NOREV);
      rev[e] = (int)e;
    } else {
      rev[e] = lo;
    }
  }
}

__global__ void k_flags(const int* __restrict__ rev, long long K, char* flags) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x)
    flags[e] = rev[e] < e;
}

__global__ void k_und(const int* __restrict__ sel, const int* __restrict__ rev, long long L, int2* und) {
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < L; u += (long long)gridDim.x * blockDim.x) {
    const int e = sel[u];
    und[u] = make_int2(e, rev[e]);
  }
}

}  // namespace

// Fills h->rowptr/col/rev/erow/slot_of_link/und (device) and rowptr_host.  Returns GBX_OK or an error code.
int build_graph_device(gbx_sbm* h, const int64_t* links_host, std::vector<int>* rowptr_host) {
  const long long n = h->n, K = h->K;
  cudaStream_t st = h->stream;
  auto alloc = [&](void** p, size_t bytes, bool keep) -> int {
    GBX_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 1));
    if (keep) h->dev_bytes += (int64_t)bytes;
    return GBX_OK;
  };
#define GBX_A(ptr, count, keep)                                                          \
  do {                                                                                   \
    int _rc = alloc((void**)&(ptr), (size_t)(count) * sizeof(*(ptr)), keep);             \
    if (_rc != GBX_OK) return _rc;                                                       \
  } while (0)
  GBX_A(h->rowptr, n + 1, true);
  GBX_A(h->col, K, true);
  GBX_A(h->rev, K, true);
  GBX_A(h->erow, K, true);
  GBX_A(h->slot_of_link, K, true);
  GBX_A(h->und, K / 2 + 1, true);

  int64_t* d_links = nullptr;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  int *vals = nullptr, *vals2 = nullptr, *deg = nullptr, *err = nullptr, *sel = nullptr, *nsel = nullptr;
  char* flags = nullptr;
  void* tmp = nullptr;
  GBX_A(d_links, 2 * K, false);
  GBX_A(keys, K, false);
  GBX_A(keys2, K, false);
  GBX_A(vals, K, false);
  GBX_A(vals2, K, false);
  GBX_A(deg, n + 1, false);
  GBX_A(err, 1, false);
  GBX_A(nsel, 1, false);
  auto cleanup = [&]() {
    cudaFree(d_links); cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(deg);
    cudaFree(err); cudaFree(sel); cudaFree(nsel); cudaFree(flags); cudaFree(tmp);
  };
  auto fail = [&](int rc) { cleanup(); return rc; };
#define GBX_C(expr)                                                                      \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      last_error() = std::string(#expr) + " -> " + cudaGetErrorString(_e);               \
      return fail(GBX_ERR_CUDA);                                                         \
    }                                                                                    \
  } while (0)
  const int grid = h->num_sms * 8;
  GBX_C(cudaMemcpyAsync(d_links, links_host, (size_t)2 * K * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  GBX_C(cudaMemsetAsync(deg, 0, (size_t)(n + 1) * sizeof(int), st));
  GBX_C(cudaMemsetAsync(err, 0, sizeof(int), st));
  if (K > 0) k_keys<<<grid, 256, 0, st>>>(d_links, K, n, keys, vals, deg, err);
  // rowptr = exclusive scan of the in-degrees
  size_t tb1 = 0, tb2 = 0, tb3 = 0;
  int bits = 32;
  while ((1ll << (bits - 32)) < n && bits < 64) ++bits;
  cub::DeviceScan::ExclusiveSum(nullptr, tb1, deg, h->rowptr, (int)(n + 1), st);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, keys, keys2, vals, vals2, (int)K, 0, bits, st);
  GBX_C(cudaMalloc((void**)&flags, (size_t)K + 1));
  GBX_C(cudaMalloc((void**)&sel, ((size_t)K / 2 + 1) * sizeof(int)));
  cub::DeviceSelect::Flagged(nullptr, tb3, cub::CountingInputIterator<int>(0), flags, sel, nsel, (int)K, st);
  const size_t tb = std::max(tb1, std::max(tb2, tb3));
  GBX_C(cudaMalloc(&tmp, tb ? tb : 1));
  size_t t = tb;
  GBX_C(cub::DeviceScan::ExclusiveSum(tmp, t, deg, h->rowptr, (int)(n + 1), st));
  if (K > 0) {
    t = tb;
    GBX_C(cub::DeviceRadixSort::SortPairs(tmp, t, keys, keys2, vals, vals2, (int)K, 0, bits, st));
    k_slots<<<grid, 256, 0, st>>>(keys2, vals2, K, h->col, h->erow, h->slot_of_link, err);
    k_rev<<<grid, 256, 0, st>>>(h->rowptr, h->col, h->erow, K, h->rev, err);
    k_flags<<<grid, 256, 0, st>>>(h->rev, K, flags);
    t = tb;
    GBX_C(cub::DeviceSelect::Flagged(tmp, t, cub::CountingInputIterator<int>(0), flags, sel, nsel, (int)K, st));
  }
  int herr = 0, hn = 0;
  rowptr_host->resize((size_t)n + 1);
  GBX_C(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (K > 0) GBX_C(cudaMemcpyAsync(&hn, nsel, sizeof(int), cudaMemcpyDeviceToHost, st));
  GBX_C(cudaMemcpyAsync(rowptr_host->data(), h->rowptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  GBX_C(cudaStreamSynchronize(st));
  GBX_C(cudaGetLastError());
  if (herr & ERR_RANGE) { last_error() = "links: vertex id out of 1..n"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_SELF) { last_error() = "links: self loop (run simplegraph first)"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_DUP) { last_error() = "links: duplicate link (run simplegraph first)"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_NOREV) {
    last_error() = "links: reverse link missing (graph must be bidirected)";
    return fail(GBX_ERR_INVALID);
  }
  if (2ll * hn != K) { last_error() = "links: odd number of directed links"; return fail(GBX_ERR_INVALID); }
  if (K > 0) k_und<<<grid, 256, 0, st>>>(sel, h->rev, K / 2, h->und);
  GBX_C(cudaStreamSynchronize(st));
  cleanup();
  h->launches += 6;
  return GBX_OK;
#undef GBX_A
#undef GBX_C
}

}  // namespace gbx
