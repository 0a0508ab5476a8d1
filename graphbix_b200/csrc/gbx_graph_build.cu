// Device-side construction of the graph structures every gbx_* handle needs, from the reference's `links`
// matrix (K x 2 Int64, column-major, 1-based, both directions of every edge; what `simplegraph` returns,
// sbm.jl:414-425).  The reference builds `A`, `nb` and `inds` for the same purpose (sbm.jl:265-275).
//
//   slot e of row i  <->  message col[e] -> i          (destination-major CSR, rows sorted by source)
//   rev[e]            =   slot of the opposite message
//   erow[e]           =   i (row of the slot)
//   slot_of_link[k]   =   slot of links[k]              (import / export of messages in `links` order)
//
// A partitioned run (`world` ranks, one GPU each) splits the vertices into `world` equal contiguous ranges; because
// the CSR is destination-major, a rank's rows are one contiguous range of the global slot space, so its local slot
// is the global slot minus a base and rev[e] becomes (owner rank << 28 | slot in the owner's buffer).
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

struct PartBases {
  long long sbeg[MAXPEER + 1];
};

// Local slice of the global structure: slots [sbeg, sbeg + Kl) of rank `rank`.
__global__ void k_slice(const int* __restrict__ gcol, const int* __restrict__ grev, const int* __restrict__ gerow,
                        long long sbeg, long long Kl, long long chunk, PartBases pb, int* col, int* rev, int* erow) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < Kl; e += (long long)gridDim.x * blockDim.x) {
    const long long ge = sbeg + e;
    const int j = gcol[ge];
    const int owner = (int)(j / chunk);
    col[e] = j;
    erow[e] = gerow[ge];
    rev[e] = (int)(((unsigned)owner << REV_SHIFT) | (unsigned)(grev[ge] - pb.sbeg[owner]));
  }
}

}  // namespace

// Builds the global structure on the device, then keeps this rank's slice in the handle: h->rowptr (local, rebased),
// h->col / h->erow (global vertex ids), h->rev (packed owner/slot), h->slot_of_link (global slot of every link) and
// the partition bases.  rowptr_global receives the global row pointers.  Sets h->n, h->K, h->vbeg, h->sbeg.
int build_graph_device(gbx_sbm* h, const int64_t* links_host, std::vector<int>* rowptr_global) {
  const long long n = h->n_global, K = h->K_global;
  cudaStream_t st = h->stream;
  int *g_rowptr = nullptr, *g_col = nullptr, *g_rev = nullptr, *g_erow = nullptr;
  int64_t* d_links = nullptr;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  int *vals = nullptr, *vals2 = nullptr, *deg = nullptr, *err = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() {
    cudaFree(g_rowptr); cudaFree(g_col); cudaFree(g_rev); cudaFree(g_erow);
    cudaFree(d_links); cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(deg);
    cudaFree(err); cudaFree(tmp);
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
#define GBX_T(ptr, count) GBX_C(cudaMalloc((void**)&(ptr), std::max<size_t>((size_t)(count), 1) * sizeof(*(ptr))))
  GBX_T(g_rowptr, n + 1);
  GBX_T(g_col, K);
  GBX_T(g_rev, K);
  GBX_T(g_erow, K);
  GBX_C(cudaMalloc((void**)&h->slot_of_link, std::max<size_t>((size_t)K, 1) * sizeof(int)));
  h->dev_bytes += (int64_t)K * sizeof(int);
  // `links` may already live on this device (a partitioned run whose edge list was generated or received there)
  const int64_t* links = nullptr;
  {
    cudaPointerAttributes at;
    if (links_host && cudaPointerGetAttributes(&at, links_host) == cudaSuccess && at.type == cudaMemoryTypeDevice &&
        at.device == h->device)
      links = links_host;
    cudaGetLastError();
  }
  if (!links) {
    GBX_T(d_links, 2 * K);
    links = d_links;
  }
  GBX_T(keys, K);
  GBX_T(keys2, K);
  GBX_T(vals, K);
  GBX_T(vals2, K);
  GBX_T(deg, n + 1);
  GBX_T(err, 1);
  const int grid = h->num_sms * 8;
  if (links == d_links)
    GBX_C(cudaMemcpyAsync(d_links, links_host, (size_t)2 * K * sizeof(int64_t), cudaMemcpyDefault, st));
  GBX_C(cudaMemsetAsync(deg, 0, (size_t)(n + 1) * sizeof(int), st));
  GBX_C(cudaMemsetAsync(err, 0, sizeof(int), st));
  if (K > 0) k_keys<<<grid, 256, 0, st>>>(links, K, n, keys, vals, deg, err);
  size_t tb1 = 0, tb2 = 0;
  int bits = 32;
  while ((1ll << (bits - 32)) < n && bits < 64) ++bits;
  cub::DeviceScan::ExclusiveSum(nullptr, tb1, deg, g_rowptr, (int)(n + 1), st);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, keys, keys2, vals, vals2, (int)K, 0, bits, st);
  const size_t tb = std::max(tb1, tb2);
  GBX_C(cudaMalloc(&tmp, tb ? tb : 1));
  size_t t = tb;
  GBX_C(cub::DeviceScan::ExclusiveSum(tmp, t, deg, g_rowptr, (int)(n + 1), st));
  if (K > 0) {
    t = tb;
    GBX_C(cub::DeviceRadixSort::SortPairs(tmp, t, keys, keys2, vals, vals2, (int)K, 0, bits, st));
    k_slots<<<grid, 256, 0, st>>>(keys2, vals2, K, g_col, g_erow, h->slot_of_link, err);
    k_rev<<<grid, 256, 0, st>>>(g_rowptr, g_col, g_erow, K, g_rev, err);
  }
  int herr = 0;
  rowptr_global->resize((size_t)n + 1);
  GBX_C(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
  GBX_C(cudaMemcpyAsync(rowptr_global->data(), g_rowptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  GBX_C(cudaStreamSynchronize(st));
  GBX_C(cudaGetLastError());
  // the sort workspace is the peak of the build: give it back before the per-rank arrays are allocated
  cudaFree(d_links); cudaFree(keys); cudaFree(keys2); cudaFree(vals); cudaFree(vals2); cudaFree(deg); cudaFree(tmp);
  d_links = nullptr; keys = keys2 = nullptr; vals = vals2 = deg = nullptr; tmp = nullptr;
  if (herr & ERR_RANGE) { last_error() = "links: vertex id out of 1..n"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_SELF) { last_error() = "links: self loop (run simplegraph first)"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_DUP) { last_error() = "links: duplicate link (run simplegraph first)"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_NOREV) {
    last_error() = "links: reverse link missing (graph must be bidirected)";
    return fail(GBX_ERR_INVALID);
  }
  // ---- this rank's slice ----
  const long long chunk = (n + h->world - 1) / h->world;
  PartBases pb;
  for (int r = 0; r <= MAXPEER; ++r) {
    const long long v = std::min<long long>(n, (long long)std::min(r, h->world) * chunk);
    h->vbeg_all[r] = v;
    h->sbeg_all[r] = (*rowptr_global)[(size_t)v];
    pb.sbeg[r] = h->sbeg_all[r];
  }
  h->vbeg = h->vbeg_all[h->rank];
  h->sbeg = h->sbeg_all[h->rank];
  h->n = h->vbeg_all[h->rank + 1] - h->vbeg;
  h->K = h->sbeg_all[h->rank + 1] - h->sbeg;
  for (int r = 0; r < h->world; ++r)
    if (h->sbeg_all[r + 1] - h->sbeg_all[r] > (long long)REV_MASK) {
      last_error() = "too many edge slots on one rank (limit 2^29)";
      return fail(GBX_ERR_INVALID);
    }
  auto keep = [&](void** p, size_t bytes) -> cudaError_t {
    h->dev_bytes += (int64_t)bytes;
    return cudaMalloc(p, bytes ? bytes : 1);
  };
  GBX_C(keep((void**)&h->rowptr, (size_t)(h->n + 1) * sizeof(int)));
  GBX_C(keep((void**)&h->col, (size_t)h->K * sizeof(int)));
  GBX_C(keep((void**)&h->rev, (size_t)h->K * sizeof(int)));
  GBX_C(keep((void**)&h->erow, (size_t)h->K * sizeof(int)));
  {
    std::vector<int> rl((size_t)h->n + 1);
    for (long long i = 0; i <= h->n; ++i) rl[(size_t)i] = (*rowptr_global)[(size_t)(h->vbeg + i)] - (int)h->sbeg;
    GBX_C(cudaMemcpyAsync(h->rowptr, rl.data(), rl.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    if (h->K > 0) k_slice<<<grid, 256, 0, st>>>(g_col, g_rev, g_erow, h->sbeg, h->K, chunk, pb, h->col, h->rev, h->erow);
    GBX_C(cudaStreamSynchronize(st));
  }
  GBX_C(cudaGetLastError());
  cleanup();
  h->launches += 5;
  return GBX_OK;
#undef GBX_T
#undef GBX_C
}

}  // namespace gbx
