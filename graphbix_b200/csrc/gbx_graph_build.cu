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
// is the global slot minus a base and rev[e] becomes (owner rank << REV_SHIFT (29) | slot in the owner's buffer).
//
// One 64-bit radix sort (CUB) of (destination, source) keys; everything else is a map or a scan.
#include <cub/cub.cuh>

#include <functional>
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

// Pull layout (gbx_sbm_pull.cuh): local slot e (row i, column j) keeps the OUTGOING message i -> j and, before the first
// sweep, the incoming one j -> i.  in_link[e] / out_link[e] = index in `links` of (j -> i) / (i -> j); link_of_slot is the
// sorted value array of the build (global slot -> link).
__global__ void k_link_maps(const int* __restrict__ link_of_slot, const int* __restrict__ grev, long long sbeg, long long Kl,
                            int* __restrict__ in_link, int* __restrict__ out_link) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < Kl; e += (long long)gridDim.x * blockDim.x) {
    in_link[e] = link_of_slot[sbeg + e];
    out_link[e] = link_of_slot[grev[sbeg + e]];
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
// `while_sorting` (optional) is called once the global row pointers and this rank's bases (h->vbeg, h->sbeg, h->n, h->K)
// are on the host, while the GPU is still sorting the links: host work that needs only the degrees (the tile schedule)
// overlaps the sort.  It must not touch the device.
int build_graph_device(gbx_sbm* h, const int64_t* links_host, std::vector<int>* rowptr_global, GlobalGraph* keep_global,
                       const std::function<void()>* while_sorting) {
  const long long n = h->n_global, K = h->K_global;
  cudaStream_t st = h->stream;
  int *g_rowptr = nullptr, *g_col = nullptr, *g_rev = nullptr, *g_erow = nullptr;
  int64_t* d_links = nullptr;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  int *vals = nullptr, *vals2 = nullptr, *deg = nullptr, *err = nullptr;
  void* tmp = nullptr;
  cudaEvent_t ev_rowptr = nullptr;
  const int dv = h->device;
  const bool pooled = h->world == 1;
  auto cleanup = [&]() {
    dev_free(dv, g_rowptr); dev_free(dv, g_col); dev_free(dv, g_rev); dev_free(dv, g_erow);
    dev_free(dv, d_links); dev_free(dv, keys); dev_free(dv, keys2); dev_free(dv, vals); dev_free(dv, vals2);
    dev_free(dv, deg); dev_free(dv, err); dev_free(dv, tmp);
    if (ev_rowptr) cudaEventDestroy(ev_rowptr);
    ev_rowptr = nullptr;
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
#define GBX_T(ptr, count) GBX_C(dev_malloc(dv, (void**)&(ptr), std::max<size_t>((size_t)(count), 1) * sizeof(*(ptr)), pooled))
  GBX_T(g_rowptr, n + 1);
  GBX_T(g_col, K);
  GBX_T(g_rev, K);
  GBX_T(g_erow, K);
  GBX_C(dev_malloc(dv, (void**)&h->slot_of_link, std::max<size_t>((size_t)K, 1) * sizeof(int), pooled));
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
    GBX_C(upload_async(h->device, d_links, links_host, (size_t)2 * K * sizeof(int64_t), st));
  GBX_C(cudaMemsetAsync(deg, 0, (size_t)(n + 1) * sizeof(int), st));
  GBX_C(cudaMemsetAsync(err, 0, sizeof(int), st));
  if (K > 0) k_keys<<<grid, 256, 0, st>>>(links, K, n, keys, vals, deg, err);
  size_t tb1 = 0, tb2 = 0;
  int bits = 32;
  while ((1ll << (bits - 32)) < n && bits < 64) ++bits;
  cub::DeviceScan::ExclusiveSum(nullptr, tb1, deg, g_rowptr, (int)(n + 1), st);
  cub::DeviceRadixSort::SortPairs(nullptr, tb2, keys, keys2, vals, vals2, (int)K, 0, bits, st);
  const size_t tb = std::max(tb1, tb2);
  GBX_C(dev_malloc(dv, &tmp, tb ? tb : 1, pooled));
  size_t t = tb;
  GBX_C(cub::DeviceScan::ExclusiveSum(tmp, t, deg, g_rowptr, (int)(n + 1), st));
  // the row pointers go to the host now: everything the host derives from the degrees overlaps the sort below
  rowptr_global->resize((size_t)n + 1);
  GBX_C(cudaEventCreateWithFlags(&ev_rowptr, cudaEventDisableTiming));
  GBX_C(cudaMemcpyAsync(rowptr_global->data(), g_rowptr, (size_t)(n + 1) * sizeof(int), cudaMemcpyDeviceToHost, st));
  GBX_C(cudaEventRecord(ev_rowptr, st));
  if (K > 0) {
    t = tb;
    GBX_C(cub::DeviceRadixSort::SortPairs(tmp, t, keys, keys2, vals, vals2, (int)K, 0, bits, st));
    k_slots<<<grid, 256, 0, st>>>(keys2, vals2, K, g_col, g_erow, h->slot_of_link, err);
    k_rev<<<grid, 256, 0, st>>>(g_rowptr, g_col, g_erow, K, g_rev, err);
  }
  int herr = 0;
  GBX_C(cudaEventSynchronize(ev_rowptr));
  // ---- this rank's slice (host arithmetic on the row pointers), then the caller's host work, then wait for the sort ----
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
  if (while_sorting && *while_sorting) (*while_sorting)();
  GBX_C(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));   // (a copy into pageable memory waits for the stream)
  GBX_C(cudaStreamSynchronize(st));
  GBX_C(cudaGetLastError());
  if (herr & ERR_RANGE) { last_error() = "links: vertex id out of 1..n"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_SELF) { last_error() = "links: self loop (run simplegraph first)"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_DUP) { last_error() = "links: duplicate link (run simplegraph first)"; return fail(GBX_ERR_INVALID); }
  if (herr & ERR_NOREV) {
    last_error() = "links: reverse link missing (graph must be bidirected)";
    return fail(GBX_ERR_INVALID);
  }
  for (int r = 0; r < h->world; ++r)
    if (h->sbeg_all[r + 1] - h->sbeg_all[r] > (long long)REV_MASK) {
      last_error() = "too many edge slots on one rank (limit 2^29)";
      return fail(GBX_ERR_INVALID);
    }
  if (want_pull_layout(h->ops, h->world)) {  // link <-> slot maps of the pull layout, while the sorted values are still around
    GBX_C(dev_malloc(dv, (void**)&h->in_link, (size_t)std::max<long long>(h->K, 1) * sizeof(int), pooled));
    GBX_C(dev_malloc(dv, (void**)&h->out_link, (size_t)std::max<long long>(h->K, 1) * sizeof(int), pooled));
    h->dev_bytes += 2 * (int64_t)h->K * sizeof(int);
    if (h->K > 0) k_link_maps<<<grid, 256, 0, st>>>(vals2, g_rev, h->sbeg, h->K, h->in_link, h->out_link);
    GBX_C(cudaStreamSynchronize(st));
  }
  // the sort workspace is the peak of the build: give it back before the per-rank arrays are allocated
  dev_free(dv, d_links); dev_free(dv, keys); dev_free(dv, keys2); dev_free(dv, vals); dev_free(dv, vals2);
  dev_free(dv, deg); dev_free(dv, tmp);
  d_links = nullptr; keys = keys2 = nullptr; vals = vals2 = deg = nullptr; tmp = nullptr;
  auto keep = [&](void** p, size_t bytes) -> cudaError_t {
    h->dev_bytes += (int64_t)bytes;
    return dev_malloc(dv, p, bytes + 64, pooled);   // slack: the pull kernel's bulk copies round their windows outwards
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
  if (keep_global && h->world > 1) {  // the exchange plan (below) reads the global columns and reverse slots
    keep_global->col = g_col;
    keep_global->rev = g_rev;
    g_col = g_rev = nullptr;
  }
  cleanup();
  h->launches += 5;
  return GBX_OK;
#undef GBX_T
#undef GBX_C
}

void free_global(GlobalGraph* g) {
  if (!g) return;
  cudaFree(g->col);  // only kept for partitioned handles, which never use the block cache
  cudaFree(g->rev);
  g->col = g->rev = nullptr;
}

// ---- plan of the packed boundary exchange (see gbx_sbm_internal.h) -------------------------------------------------
namespace {

struct ColOwnedBy {  // local slot e: is its column (the vertex the new message goes to) owned by `owner`?
  const int* col;
  long long chunk;
  int owner;
  __device__ bool operator()(int e) const { return col[e] / chunk == owner; }
};
struct IncomingSlot {  // global slot ge outside my rows whose column is mine: its message lands in one of my rows
  const int* gcol;
  long long chunk, my_beg, my_end;
  int me;
  __device__ bool operator()(int ge) const { return (ge < my_beg || ge >= my_end) && gcol[ge] / chunk == me; }
};

__global__ void k_set_rev_pack(const int* __restrict__ sel, long long cnt, int owner, int* rev_pack) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += (long long)gridDim.x * blockDim.x)
    rev_pack[sel[k]] = (int)(((unsigned)owner << REV_SHIFT) | (unsigned)k);
}
__global__ void k_unpack_slots(const int* __restrict__ selg, long long cnt, const int* __restrict__ grev, long long sbeg_me,
                               int* unpack_slot) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < cnt; k += (long long)gridDim.x * blockDim.x)
    unpack_slot[k] = (int)(grev[selg[k]] - sbeg_me);
}
struct Bounds {
  long long b[MAXPEER + 2];
  int nb;
};
// out[g][o] = number of slots in [b[g], b[g+1]) whose column is owned by o
__global__ void k_group_owner_hist(const int* __restrict__ col, Bounds bd, long long chunk, unsigned long long* out) {
  __shared__ unsigned int sh[(MAXPEER + 1) * MAXPEER];
  for (int i = threadIdx.x; i < (MAXPEER + 1) * MAXPEER; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const long long n = bd.b[bd.nb];
  for (long long e = bd.b[0] + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    int g = 0;
    while (g + 1 < bd.nb && e >= bd.b[g + 1]) ++g;
    atomicAdd(&sh[g * MAXPEER + (int)(col[e] / chunk)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (MAXPEER + 1) * MAXPEER; i += blockDim.x)
    if (sh[i]) atomicAdd(&out[i], (unsigned long long)sh[i]);
}

}  // namespace

int build_exchange_plan(gbx_sbm* h, const GlobalGraph& g, const std::vector<int4>& tiles) {
  if (h->world <= 1) return GBX_OK;
  const long long n = h->n_global, chunk = (n + h->world - 1) / h->world;
  const long long Kl = h->K, Kg = h->K_global;
  const int me = h->rank, P = h->world;
  cudaStream_t st = h->stream;
  int *sel = nullptr, *dnum = nullptr;
  unsigned long long* dh = nullptr;
  void* tmp = nullptr;
  auto cleanup = [&]() { cudaFree(sel); cudaFree(dnum); cudaFree(dh); cudaFree(tmp); };
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
  const size_t HB = (size_t)(MAXPEER + 1) * MAXPEER;
  GBX_C(cudaMalloc((void**)&dh, HB * sizeof(unsigned long long)));
  GBX_C(cudaMalloc((void**)&dnum, sizeof(int)));
  std::vector<unsigned long long> hh(HB);
  // counts between every pair of ranks: cnt[s][o] = slots in rows of s whose column is owned by o
  long long cnt[MAXPEER][MAXPEER] = {};
  {
    Bounds bd;
    bd.nb = P;
    for (int r = 0; r <= P; ++r) bd.b[r] = h->sbeg_all[r];
    GBX_C(cudaMemsetAsync(dh, 0, HB * sizeof(unsigned long long), st));
    if (Kg > 0) k_group_owner_hist<<<grid, 256, 0, st>>>(g.col, bd, chunk, dh);
    GBX_C(cudaMemcpyAsync(hh.data(), dh, HB * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    GBX_C(cudaStreamSynchronize(st));
    for (int s = 0; s < P; ++s)
      for (int o = 0; o < P; ++o) cnt[s][o] = (long long)hh[(size_t)s * MAXPEER + o];
  }
  h->send_off[0] = 0;
  for (int o = 0; o < P; ++o) h->send_off[o + 1] = h->send_off[o] + (o == me ? 0 : cnt[me][o]);
  for (int o = P; o < MAXPEER; ++o) h->send_off[o + 1] = h->send_off[o];
  h->send_total = h->send_off[P];
  for (int o = 0; o < P; ++o) {
    long long off = 0;
    for (int s = 0; s < me; ++s)
      if (s != o) off += cnt[s][o];
    h->my_off_at[o] = off;
  }
  // pieces of the sweep: equal runs of tiles, at least one full grid each
  const int ntiles = (int)tiles.size();
  int nch = h->nchunks < 1 ? 1 : (h->nchunks > MAXCHUNK ? MAXCHUNK : h->nchunks);
  while (nch > 1 && ntiles / nch < h->num_sms * 8) --nch;
  h->nchunks = nch;
  Bounds cb;
  cb.nb = nch;
  for (int c = 0; c <= nch; ++c) {
    h->chunk_tile[c] = (int)((long long)ntiles * c / nch);
    cb.b[c] = (c == 0) ? 0 : (c == nch ? Kl : (long long)tiles[(size_t)h->chunk_tile[c]].z);
  }
  {
    GBX_C(cudaMemsetAsync(dh, 0, HB * sizeof(unsigned long long), st));
    if (Kl > 0) k_group_owner_hist<<<grid, 256, 0, st>>>(h->col, cb, chunk, dh);
    GBX_C(cudaMemcpyAsync(hh.data(), dh, HB * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    GBX_C(cudaStreamSynchronize(st));
    for (int o = 0; o < MAXPEER; ++o) {
      long long acc = 0;
      for (int c = 0; c <= nch; ++c) {
        h->chunk_send[c][o] = acc;
        if (c < nch) acc += (long long)hh[(size_t)c * MAXPEER + o];
      }
    }
  }
  auto keep = [&](void** p, size_t bytes) -> cudaError_t {
    h->dev_bytes += (int64_t)bytes;
    return cudaMalloc(p, bytes ? bytes : 1);
  };
  const size_t rowb = (size_t)h->qm * sizeof(double);
  GBX_C(keep((void**)&h->rev_pack, (size_t)Kl * sizeof(int)));
  GBX_C(keep((void**)&h->unpack_slot, (size_t)h->send_total * sizeof(int)));
  GBX_C(keep((void**)&h->sendbuf, (size_t)h->send_total * rowb));
  GBX_C(keep((void**)&h->recvbuf[0], (size_t)h->send_total * rowb));
  GBX_C(keep((void**)&h->recvbuf[1], (size_t)h->send_total * rowb));
  h->recvp[0][me] = h->recvbuf[0];
  h->recvp[1][me] = h->recvbuf[1];
  GBX_C(cudaMemcpyAsync(h->rev_pack, h->rev, (size_t)Kl * sizeof(int), cudaMemcpyDeviceToDevice, st));
  // send positions: the k-th local slot (in slot order) whose column belongs to o goes to row k of the run for o
  long long maxsel = h->send_total;
  for (int o = 0; o < P; ++o)
    if (o != me) maxsel = std::max(maxsel, cnt[me][o]);
  GBX_C(cudaMalloc((void**)&sel, std::max<size_t>((size_t)maxsel, 1) * sizeof(int)));
  size_t tb = 0, tb2 = 0;
  cub::CountingInputIterator<int> ids(0);
  cub::DeviceSelect::If(nullptr, tb, ids, sel, dnum, (int)Kl, ColOwnedBy{h->col, chunk, 0}, st);
  cub::DeviceSelect::If(nullptr, tb2, ids, sel, dnum, (int)Kg, IncomingSlot{g.col, chunk, h->sbeg, h->sbeg + Kl, me}, st);
  tb = std::max(tb, tb2);
  GBX_C(cudaMalloc(&tmp, tb ? tb : 1));
  for (int o = 0; o < P; ++o) {
    if (o == me || cnt[me][o] == 0) continue;
    size_t t = tb;
    GBX_C(cub::DeviceSelect::If(tmp, t, ids, sel, dnum, (int)Kl, ColOwnedBy{h->col, chunk, o}, st));
    k_set_rev_pack<<<grid, 256, 0, st>>>(sel, cnt[me][o], o, h->rev_pack);
  }
  // receive side: rows arrive grouped by sender (ascending rank), each group in the sender's slot order; the global
  // slots outside my rows whose column is mine, in slot order, are exactly that sequence
  if (h->send_total > 0) {
    size_t t = tb;
    GBX_C(cub::DeviceSelect::If(tmp, t, ids, sel, dnum, (int)Kg, IncomingSlot{g.col, chunk, h->sbeg, h->sbeg + Kl, me}, st));
    int got = 0;
    GBX_C(cudaMemcpyAsync(&got, dnum, sizeof(int), cudaMemcpyDeviceToHost, st));
    GBX_C(cudaStreamSynchronize(st));
    if ((long long)got != h->send_total) {
      last_error() = "exchange plan: incoming and outgoing boundary counts differ (graph not symmetric?)";
      return fail(GBX_ERR_INVALID);
    }
    k_unpack_slots<<<grid, 256, 0, st>>>(sel, h->send_total, g.rev, h->sbeg, h->unpack_slot);
  }
  GBX_C(cudaStreamSynchronize(st));
  GBX_C(cudaGetLastError());
  GBX_C(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  h->peer_stream[1] = h->copy_stream;
  for (int k = 2; k < P; ++k) GBX_C(cudaStreamCreateWithFlags(&h->peer_stream[k], cudaStreamNonBlocking));
  for (int k = 1; k < P; ++k) GBX_C(cudaEventCreateWithFlags(&h->ev_peer[k], cudaEventDisableTiming));
  for (int c = 0; c < MAXCHUNK; ++c) GBX_C(cudaEventCreateWithFlags(&h->ev_piece[c], cudaEventDisableTiming));
  GBX_C(cudaEventCreateWithFlags(&h->ev_copied, cudaEventDisableTiming));
  GBX_C(cudaEventCreateWithFlags(&h->ev_unpacked, cudaEventDisableTiming));
  cleanup();
  h->launches += 2 + P;
  return GBX_OK;
#undef GBX_C
}

}  // namespace gbx
