// Device-memory block cache.  The hosts this library serves call EM() once per (q, initial condition, sample) on the
// same graph (sbm.jl:912-943), i.e. they create and destroy handles of the same sizes over and over; cudaMalloc /
// cudaFree of the ~40 buffers of a handle cost 4-25 ms per call.  Freed blocks are kept (up to a cap) and handed to
// the next handle.  Blocks of partitioned handles are never cached (their buffers are mapped by other processes).
//   GBX_POOL=0      disable          GBX_POOL_MB=<n>   cap of cached bytes (default: a quarter of the device memory)
#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

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

// ---------------------------------------------------------------------------------------------
// Upload of a large PAGEABLE host array.  A Julia `ccall` hands over ordinary Arrays: cudaMemcpyAsync from pageable
// memory goes through the driver's single staging buffer at ~6 GB/s (creating a handle from the 256 MB Int64 edge list
// of N = 1M, degree 16: 28 ms against 9-10 ms from pinned memory).  Here a few host threads copy alternate chunks into their own
// pinned buffers and queue the DMA of each chunk on their own streams, so the host copies run in parallel with each
// other and with the transfers; `st` then waits for all of them.  Pinned or small sources take one cudaMemcpyAsync.
// On return the source has been read completely (as with cudaMemcpyAsync from pageable memory).
//   GBX_STAGE_THREADS=<n>  host threads (default 4, 0: always the plain copy)   GBX_STAGE_MIN_MB=<n>  smallest staged upload (16)
namespace {
constexpr size_t STAGE_CHUNK = size_t(2) << 20;   // 2 buffers per thread: 16 MB of pinned memory with 4 threads
constexpr size_t STAGE_MIN = size_t(16) << 20;
constexpr int STAGE_MAXT = 8;
struct Stager {
  std::mutex m;  // one staged upload at a time per process (the buffers are shared)
  int device = -1;
  int nt = 0;
  char* buf[STAGE_MAXT][2] = {};
  cudaEvent_t ev[STAGE_MAXT][2] = {};
  cudaEvent_t done[STAGE_MAXT] = {};
  cudaStream_t s[STAGE_MAXT] = {};
  cudaEvent_t start = nullptr;
};
Stager& stager() {
  static Stager* p = new Stager();  // leaked on purpose, like the pool
  return *p;
}
int stage_threads() {  // read per call: tests switch it
  int t = 4;
  if (const char* e = getenv("GBX_STAGE_THREADS")) t = atoi(e);
  const int hc = (int)std::thread::hardware_concurrency();
  if (hc > 0) t = std::min(t, std::max(1, hc - 1));
  return std::max(0, std::min(t, STAGE_MAXT));
}
size_t stage_min_bytes() {
  if (const char* e = getenv("GBX_STAGE_MIN_MB")) return (size_t)std::max(0ll, atoll(e)) << 20;
  return STAGE_MIN;
}
void stager_release(Stager& S) {
  for (int t = 0; t < STAGE_MAXT; ++t) {
    for (int b = 0; b < 2; ++b) {
      if (S.buf[t][b]) cudaFreeHost(S.buf[t][b]);
      if (S.ev[t][b]) cudaEventDestroy(S.ev[t][b]);
      S.buf[t][b] = nullptr;
      S.ev[t][b] = nullptr;
    }
    if (S.done[t]) cudaEventDestroy(S.done[t]);
    if (S.s[t]) cudaStreamDestroy(S.s[t]);
    S.done[t] = nullptr;
    S.s[t] = nullptr;
  }
  if (S.start) cudaEventDestroy(S.start);
  S.start = nullptr;
  S.device = -1;
  S.nt = 0;
}
cudaError_t stager_prepare(Stager& S, int device, int nt) {
  if (S.device == device && S.nt == nt) return cudaSuccess;
  stager_release(S);  // (buffers of another device: pinned memory is portable, streams and events are not)
  cudaError_t e = cudaSuccess;
  for (int t = 0; t < nt && e == cudaSuccess; ++t) {
    e = cudaStreamCreateWithFlags(&S.s[t], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.done[t], cudaEventDisableTiming);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
      e = cudaHostAlloc((void**)&S.buf[t][b], STAGE_CHUNK, cudaHostAllocDefault);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.ev[t][b], cudaEventDisableTiming);
    }
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.start, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    stager_release(S);
    return e;
  }
  S.device = device;
  S.nt = nt;
  return cudaSuccess;
}
}  // namespace

cudaError_t upload_segments_async(int device, const UploadSeg* segs, size_t nseg, cudaStream_t st) {
  size_t total = 0;
  const void* first = nullptr;
  for (size_t i = 0; i < nseg; ++i) {
    if (segs[i].bytes && !first) first = segs[i].src;
    total += segs[i].bytes;
  }
  if (total == 0) return cudaSuccess;
  const int nt = stage_threads();
  bool pageable = false;
  if (nt > 0 && total >= stage_min_bytes()) {  // (the segments of a call are of one kind: the first one is asked)
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, first);
    pageable = (e == cudaSuccess && at.type == cudaMemoryTypeUnregistered);
    if (e != cudaSuccess) cudaGetLastError();
  }
  auto plain = [&]() {
    cudaError_t e = cudaSuccess;
    for (size_t i = 0; i < nseg && e == cudaSuccess; ++i)
      if (segs[i].bytes) e = cudaMemcpyAsync(segs[i].dst, segs[i].src, segs[i].bytes, cudaMemcpyDefault, st);
    return e;
  };
  if (!pageable) return plain();
  Stager& S = stager();
  std::lock_guard<std::mutex> g(S.m);
  cudaError_t e = stager_prepare(S, device, nt);
  if (e != cudaSuccess) {  // no pinned memory to be had: the plain copies still work
    cudaGetLastError();
    return plain();
  }
  // pieces of at most one chunk, dealt to the threads in turn
  std::vector<UploadSeg> pieces;
  pieces.reserve(nseg + total / STAGE_CHUNK + 1);
  for (size_t i = 0; i < nseg; ++i)
    for (size_t off = 0; off < segs[i].bytes; off += STAGE_CHUNK)
      pieces.push_back(UploadSeg{(char*)segs[i].dst + off, (const char*)segs[i].src + off, std::min(STAGE_CHUNK, segs[i].bytes - off)});
  // the pieces land only after what `st` holds so far (a destination may be a recycled block)
  e = cudaEventRecord(S.start, st);
  if (e != cudaSuccess) return e;
  std::atomic<int> first_err{(int)cudaSuccess};
  auto work = [&](int t) {
    cudaError_t r = cudaSetDevice(device);
    if (r == cudaSuccess) r = cudaStreamWaitEvent(S.s[t], S.start, 0);
    int b = 0;
    for (size_t c = (size_t)t; c < pieces.size() && r == cudaSuccess; c += (size_t)nt, b ^= 1) {
      r = cudaEventSynchronize(S.ev[t][b]);  // the transfer that last used this buffer (a fresh event is complete)
      if (r != cudaSuccess) break;
      memcpy(S.buf[t][b], pieces[c].src, pieces[c].bytes);
      r = cudaMemcpyAsync(pieces[c].dst, S.buf[t][b], pieces[c].bytes, cudaMemcpyHostToDevice, S.s[t]);
      if (r == cudaSuccess) r = cudaEventRecord(S.ev[t][b], S.s[t]);
    }
    if (r == cudaSuccess) r = cudaEventRecord(S.done[t], S.s[t]);
    if (r != cudaSuccess) {
      int ok = (int)cudaSuccess;
      first_err.compare_exchange_strong(ok, (int)r);
    }
  };
  std::vector<std::thread> th;
  th.reserve(nt > 0 ? nt - 1 : 0);
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  e = (cudaError_t)first_err.load();
  for (int t = 0; t < nt && e == cudaSuccess; ++t) e = cudaStreamWaitEvent(st, S.done[t], 0);
  if (e != cudaSuccess) {  // leave nothing of ours in flight behind an error
    for (int t = 0; t < nt; ++t) cudaStreamSynchronize(S.s[t]);
  }
  return e;
}

cudaError_t upload_async(int device, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  const UploadSeg seg{dst, src, bytes};
  return upload_segments_async(device, &seg, 1, st);
}

// The other direction, for results read into pageable arrays (PSI is n x q doubles): the threads' DMAs land in their
// pinned buffers and are copied out while the next chunk is in flight.  Returns when dst holds the data.
cudaError_t download_sync(int device, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaStreamSynchronize(st);
  const int nt = stage_threads();
  bool pageable = false;
  if (nt > 0 && bytes >= stage_min_bytes()) {
    cudaPointerAttributes at;
    const cudaError_t e = cudaPointerGetAttributes(&at, dst);
    pageable = (e == cudaSuccess && at.type == cudaMemoryTypeUnregistered);
    if (e != cudaSuccess) cudaGetLastError();
  }
  cudaError_t e = cudaSuccess;
  if (pageable) {
    Stager& S = stager();
    std::lock_guard<std::mutex> g(S.m);
    e = stager_prepare(S, device, nt);
    if (e == cudaSuccess) {
      e = cudaEventRecord(S.start, st);  // the source is final once what `st` holds so far has run
      if (e != cudaSuccess) return e;
      const size_t nchunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
      std::atomic<int> first_err{(int)cudaSuccess};
      auto work = [&](int t) {
        cudaError_t r = cudaSetDevice(device);
        if (r == cudaSuccess) r = cudaStreamWaitEvent(S.s[t], S.start, 0);
        int b = 0, pb = -1;
        size_t poff = 0, plen = 0;
        for (size_t c = (size_t)t; c < nchunks && r == cudaSuccess; c += (size_t)nt, b ^= 1) {
          const size_t off = c * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
          r = cudaMemcpyAsync(S.buf[t][b], (const char*)src + off, len, cudaMemcpyDeviceToHost, S.s[t]);
          if (r == cudaSuccess) r = cudaEventRecord(S.ev[t][b], S.s[t]);
          if (r == cudaSuccess && pb >= 0) {
            r = cudaEventSynchronize(S.ev[t][pb]);
            if (r == cudaSuccess) memcpy((char*)dst + poff, S.buf[t][pb], plen);
          }
          pb = b;
          poff = off;
          plen = len;
        }
        if (r == cudaSuccess && pb >= 0) {
          r = cudaEventSynchronize(S.ev[t][pb]);
          if (r == cudaSuccess) memcpy((char*)dst + poff, S.buf[t][pb], plen);
        }
        if (r != cudaSuccess) {
          cudaStreamSynchronize(S.s[t]);
          int ok = (int)cudaSuccess;
          first_err.compare_exchange_strong(ok, (int)r);
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
      work(0);
      for (auto& x : th) x.join();
      e = (cudaError_t)first_err.load();
      if (e != cudaSuccess) return e;
      return cudaStreamSynchronize(st);
    }
    cudaGetLastError();  // no pinned memory: the plain copy
  }
  e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  return e;
}

}  // namespace gbx
