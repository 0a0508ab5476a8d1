// Device-memory block cache.  The hosts this library serves call EM() once per (q, initial condition, sample) on the
// same graph (sbm.jl:912-943), i.e. they create and destroy handles of the same sizes over and over; cudaMalloc /
// cudaFree of the ~40 buffers of a handle cost 4-25 ms per call.  Freed blocks are kept (up to a cap) and handed to
// the next handle.  Blocks of partitioned handles are never cached (their buffers are mapped by other processes).
//   GBX_POOL=0      disable          GBX_POOL_MB=<n>   cap of cached bytes (default: a quarter of the device memory)
#include <map>
#include <mutex>
#include <unordered_map>

#include "gbx_sbm_internal.h"

namespace gbx {
namespace {

constexpr int MAXDEV = 16;
struct Pool {
  std::mutex m;
  std::multimap<size_t, void*> free_blocks[MAXDEV];
  std::unordered_map<void*, size_t> live[MAXDEV];
  size_t cached[MAXDEV] = {0};
  size_t cap[MAXDEV] = {0};
  bool cap_known[MAXDEV] = {false};
  bool enabled = true;
  Pool() {
    if (const char* e = getenv("GBX_POOL")) enabled = atoi(e) != 0;
  }
};
Pool& pool() {
  static Pool* p = new Pool();  // leaked on purpose: no CUDA calls during static destruction
  return *p;
}
size_t round_size(size_t b) {
  if (b < 512) return 512;
  return (b + 511) / 512 * 512;
}
size_t cap_of(Pool& P, int device) {
  if (!P.cap_known[device]) {
    P.cap_known[device] = true;
    if (const char* e = getenv("GBX_POOL_MB")) {
      P.cap[device] = (size_t)atoll(e) << 20;
    } else {
      size_t fr = 0, tot = 0;
      if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) P.cap[device] = tot / 4;
    }
  }
  return P.cap[device];
}

}  // namespace

cudaError_t dev_malloc(int device, void** p, size_t bytes, bool pooled) {
  Pool& P = pool();
  *p = nullptr;
  if (!pooled || !P.enabled || device < 0 || device >= MAXDEV) return cudaMalloc(p, bytes ? bytes : 1);
  const size_t want = round_size(bytes);
  {
    std::lock_guard<std::mutex> g(P.m);
    auto it = P.free_blocks[device].lower_bound(want);
    if (it != P.free_blocks[device].end() && it->first <= want + want / 8 + (64u << 10)) {
      *p = it->second;
      P.live[device][*p] = it->first;
      P.cached[device] -= it->first;
      P.free_blocks[device].erase(it);
      return cudaSuccess;
    }
  }
  cudaError_t e = cudaMalloc(p, want);
  if (e == cudaErrorMemoryAllocation) {  // give the cached blocks back and try once more
    cudaGetLastError();
    pool_trim(device);
    e = cudaMalloc(p, want);
  }
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> g(P.m);
    P.live[device][*p] = want;
  }
  return e;
}

// The caller guarantees that no work that touches the block is still in flight.
void dev_free(int device, void* p) {
  if (!p) return;
  Pool& P = pool();
  if (device >= 0 && device < MAXDEV) {
    std::lock_guard<std::mutex> g(P.m);
    auto it = P.live[device].find(p);
    if (it != P.live[device].end()) {
      const size_t sz = it->second;
      P.live[device].erase(it);
      if (P.enabled && P.cached[device] + sz <= cap_of(P, device)) {
        P.free_blocks[device].emplace(sz, p);
        P.cached[device] += sz;
        return;
      }
    }
  }
  cudaFree(p);
}

void pool_trim(int device) {
  Pool& P = pool();
  std::lock_guard<std::mutex> g(P.m);
  for (int d = 0; d < MAXDEV; ++d) {
    if (device >= 0 && d != device) continue;
    if (P.free_blocks[d].empty()) continue;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(d);
    for (auto& kv : P.free_blocks[d]) cudaFree(kv.second);
    P.free_blocks[d].clear();
    P.cached[d] = 0;
    cudaSetDevice(cur);
  }
}

int64_t pool_cached_bytes(int device) {
  Pool& P = pool();
  std::lock_guard<std::mutex> g(P.m);
  return (device >= 0 && device < MAXDEV) ? (int64_t)P.cached[device] : 0;
}

}  // namespace gbx
