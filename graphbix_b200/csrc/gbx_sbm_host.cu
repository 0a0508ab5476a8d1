// libgraphbix_cuda.so -- host side of the sbm.jl EM() replacement: C ABI (include/graphbix_cuda.h), graph
// construction (destination-major CSR of the 2E directed edges, reverse-edge index, tile schedule) and the
// EM loop of sbm.jl:337-367 driving the per-q kernels of gbx_sbm_q.cu.  sm_100a only, no CPU path.
//
// Data layout in HBM (DESIGN.md has the full story)
//   * slot e of row i holds the cavity message col[e] -> i, q doubles (64 B at q = 8), so a vertex reads its
//     incoming messages as one contiguous run; rev[e] = slot of the opposite message i -> col[e];
//   * two message buffers (ping-pong): a sweep reads `cur` and scatters the new messages to `nxt` at rev[e]
//     with 256-bit stores -- a synchronous (Jacobi) sweep, deterministic by construction;
//   * PSI (N x q), theta (N); tiles = runs of consecutive vertices with <= TILE_E = 128 edge slots (one CTA
//     iteration each, one thread per slot); vertices with more than 128 edges ("hubs") get a CTA each.
#include <algorithm>
#include <chrono>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <functional>
#include <vector>

#include "gbx_sbm_internal.h"

namespace gbx {

std::string& last_error() {
  static thread_local std::string s;
  return s;
}

#ifndef GBX_FOR_EACH_Q
#define GBX_FOR_EACH_Q(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)
#endif
#define GBX_DECL(qq)              \
  const SbmOps* sbm_ops_q##qq(); \
  const SbmOps* mod_ops_q##qq();
GBX_FOR_EACH_Q(GBX_DECL)
#undef GBX_DECL

static const SbmOps* ops_for(int model, int q) {
#define GBX_CASE(qq) \
  if (q == qq) return model == 1 ? mod_ops_q##qq() : sbm_ops_q##qq();
  GBX_FOR_EACH_Q(GBX_CASE)
#undef GBX_CASE
  return nullptr;
}

}  // namespace gbx

namespace gbx {
int build_graph_device(gbx_sbm* h, const int64_t* links_host, std::vector<int>* rowptr_host, GlobalGraph* keep_global,
                       const std::function<void()>* while_sorting = nullptr);
void free_global(GlobalGraph* g);
int build_exchange_plan(gbx_sbm* h, const GlobalGraph& g, const std::vector<int4>& tiles);
}

using namespace gbx;

namespace {

#define GBX_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != GBX_OK) return _r; \
  } while (0)

int fail_invalid(const char* msg) {
  last_error() = msg;
  return GBX_ERR_INVALID;
}

int fail_state() {
  last_error() = "gbx_sbm_init has not been called";
  return GBX_ERR_STATE;
}

template <typename T>
int dev_alloc(gbx_sbm* h, T** p, size_t count) {
  // 64 bytes of slack: the pull kernel's bulk copies round their (16-byte aligned) windows outwards
  const size_t bytes = std::max<size_t>(count, 1) * sizeof(T) + 64;
  GBX_CUDA_TRY(dev_malloc(h->device, (void**)p, bytes, h->world == 1));
  h->dev_bytes += (int64_t)bytes;
  return GBX_OK;
}

// Every C-ABI entry starts here.  `join` (default): the main stream first waits for an unpack of boundary rows that
// may still run on the copy stream; the steps that never touch the messages skip that so they overlap with it.
int set_device(gbx_sbm* h, bool join = true) {
  GBX_CUDA_TRY(cudaSetDevice(h->device));
  if (join && h->unpack_inflight) {
    GBX_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_unpacked, 0));
    h->unpack_inflight = false;
  }
  return GBX_OK;
}

int read_state(gbx_sbm* h) {
  GBX_CUDA_TRY(cudaMemcpyAsync(h->h_state, h->st, sizeof(SbmState), cudaMemcpyDeviceToHost, h->stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GBX_OK;
}

// Initial messages: from the caller's K_global x q array in `links` order, or drawn on the device.  Push layout:
// buffer 0 holds the incoming messages of every row; pull layout: buffer 0 outgoing, buffer 1 incoming.
int load_messages(gbx_sbm* h, const double* psicav, uint64_t seed) {
  if (!psicav) return h->pull ? h->ops->pull_load(h, nullptr, seed) : h->ops->random_messages(h, seed);
  const size_t bytes = (size_t)h->K_global * h->q * sizeof(double);
  double* tmp = h->msg[1];  // scratch before the first sweep (large enough on a single rank)
  bool own = false;
  if (h->world > 1 || h->pull) {
    GBX_CUDA_TRY(dev_malloc(h->device, (void**)&tmp, bytes ? bytes : 1, h->world == 1));
    own = true;
  }
  cudaError_t e = upload_async(h->device, tmp, psicav, bytes, h->stream);   // initial messages: K x q doubles
  int rc = GBX_OK;
  if (e == cudaSuccess) rc = h->pull ? h->ops->pull_load(h, tmp, 0) : h->ops->permute(h, tmp, h->msg[0], 1);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (own) dev_free(h->device, tmp);
  GBX_CUDA_TRY(e);
  return rc;
}

// Which sweep kernels this EM run uses (decided at init, when the options are known).
//   push: new messages scattered into the receiver's rows (sbm_vertex_kernel); the faster sweep on one GPU today.
//   pull: a vertex keeps its outgoing messages and rebuilds the incoming ones from per-vertex records
//         (gbx_sbm_pull.cuh); needs message rows of 32 / 64 / 128 bytes and no damping (a damped message cannot be
//         rebuilt from the sender's record).  What crosses NVLink in a partitioned run is then one record per vertex
//         instead of one message per cut edge, so partitioned handles (world > 1) sweep with it by default.
// GBX_PULL=1 / GBX_PULL=0 in the environment force one or the other where possible.
int choose_layout(gbx_sbm* h, double damping) {
  int pull = h->ops->pull_ok && damping == 1.0 && h->rec[0] != nullptr && h->theta_pp[0] != nullptr;
  for (int r = 0; r < h->world; ++r) pull = pull && h->recp[0][r] != nullptr && h->recp[1][r] != nullptr;   // peers mapped
  if (const char* e = getenv("GBX_PULL")) pull = pull && atoi(e) != 0;
  else pull = pull && h->world > 1;
  h->pull = pull;
  h->gen = 0;
  h->thcur = h->th_used = 0;
  if (pull && !h->tmap_ok) GBX_TRY(h->ops->pull_setup(h));
  return GBX_OK;
}

// freeenergy(): sum_i log Z_i, then two passes over the undirected edges (sums, then squared deviations).
int free_energy_steps(gbx_sbm* h) {
  const SbmOps* op = h->ops;
  GBX_TRY(op->vertex_pass(h, MODE_LOGZ));
  GBX_TRY(op->local_totals(h, MODE_LOGZ));
  GBX_TRY(op->apply_totals(h, MODE_LOGZ));
  for (int pass = 0; pass < 2; ++pass) {
    GBX_TRY(op->fe_local(h, pass));
    GBX_TRY(op->fe_apply(h, pass));
  }
  GBX_TRY(op->fe_final(h));
  return GBX_OK;
}

// One EM iteration, sbm.jl:343-347, with a synchronous sweep in place of BP().
// The residual is final after apply_totals: its copy to the pinned mirror is queued there, before update_theta, so
// that it has long arrived when the host asks for it (wait_residual) and the GPU never idles on the convergence test.
int queue_residual(gbx_sbm* h) {
  GBX_CUDA_TRY(cudaMemcpyAsync(&h->h_state->conv, &h->st->conv, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  GBX_CUDA_TRY(cudaEventRecord(h->ev_conv, h->stream));
  return GBX_OK;
}
int wait_residual(gbx_sbm* h, double* conv) {
  GBX_CUDA_TRY(cudaEventSynchronize(h->ev_conv));
  *conv = h->h_state->conv;
  return GBX_OK;
}

int em_iteration(gbx_sbm* h, bool want_residual = false) {
  const SbmOps* op = h->ops;
  GBX_TRY(op->prepare(h, 0, 0));              // gr clamp, h = update_h()                        sbm.jl:102
  GBX_TRY(op->vertex_pass(h, MODE_SWEEP));    // messages, marginals, residual, M-step statistics sbm.jl:103-136
  h->cur ^= 1;
  if (h->opt.mstep_exact) GBX_TRY(op->mstep_exact(h));   // statistics from the new messages of both directions
  GBX_TRY(op->local_totals(h, MODE_SWEEP));
  GBX_TRY(op->apply_totals(h, MODE_SWEEP));   // BPconv, gr, Crs = update_Crs()                  sbm.jl:120-123, 137-147
  if (want_residual) GBX_TRY(queue_residual(h));
  if (h->opt.degreecorrection) {              // theta = update_theta()                          sbm.jl:345-347
    GBX_TRY(op->theta_local(h));
    GBX_TRY(op->theta_apply(h));
    GBX_TRY(op->escale(h));
  }
  h->itr++;
  return GBX_OK;
}

int mod_iteration(gbx_sbm* h, bool want_residual);

// The loop of EM() (sbm.jl:342-362, mod.jl:219-235).  Large graphs: one iteration at a time, the residual read back
// early (queue_residual).  Small graphs are launch bound, so `batch` iterations are queued before the host looks:
// sbm_apply_totals records the first iteration whose residual is below the threshold in SbmState::done_at and every
// kernel of a later iteration returns at once, which leaves the state exactly as the reference's `break` would.
int em_loop(gbx_sbm* h, int itrmax, int check_every, double thr, bool mod, int* cnv, int* itrnum) {
  int batch = h->K < (int64_t)2000000 ? 8 : 1;
  if (const char* e = getenv("GBX_EM_BATCH")) batch = std::max(1, atoi(e));
  batch = std::max(batch, std::max(1, check_every));
  GBX_CUDA_TRY(cudaMemsetAsync(&h->st->done_at, 0, sizeof(int), h->stream));
  const int itr0 = h->itr;  // iterations run before this call (gbx_sbm_iterate): stamps keep growing
  h->cur_thr = thr;
  int rc = GBX_OK, itr = 0;
  // Small graphs (batch > 1) are launch bound: after the first batch (update_theta has run, every launch argument is
  // settled) two consecutive iterations -- one per message buffer parity -- are captured into a CUDA graph and replayed.
  // The replayed kernels carry the stamp -1 and read their iteration number from SbmState::it_now.
  cudaGraphExec_t gexec = nullptr;
  int64_t graph_kernels = 0;   // kernel launches recorded in the graph (two iterations)
  bool graph_ok = batch > 1 && (batch % 2) == 0 && h->world == 1 && !h->pull && !h->opt.mstep_exact &&
                  !(getenv("GBX_GRAPH") && atoi(getenv("GBX_GRAPH")) == 0);
  auto drop_graph = [&]() {
    if (gexec) cudaGraphExecDestroy(gexec);
    gexec = nullptr;
  };
  while (itr < itrmax && rc == GBX_OK) {
    const int nb = std::min(batch, itrmax - itr);
    if (graph_ok && itr > 0 && nb == batch) {
      if (!gexec) {   // record two iterations (nothing runs during a capture)
        cudaGraph_t graph = nullptr;
        const int itr_host = h->itr;
        const int64_t launches_host = h->launches;
        // The recording happens on a stream of the handle's own (the caller's stream may be the legacy default stream,
        // which cannot be captured); the graph is then launched into the caller's stream.  The shared parameter block
        // is released first so that taking it inside the capture involves no event of another stream.
        cudaStream_t user = h->stream;
        cudaError_t ce = cudaSuccess;
        if (!h->cap_stream) ce = cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking);
        if (ce == cudaSuccess) {
          h->ops->release(h);
          h->stream = h->cap_stream;
          ce = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
          if (ce == cudaSuccess) {
            h->cur_it = -1;
            for (int b = 0; b < 2 && rc == GBX_OK; ++b) rc = mod ? mod_iteration(h, false) : em_iteration(h, false);
            ce = cudaStreamEndCapture(h->stream, &graph);
            if (ce == cudaSuccess && rc == GBX_OK) ce = cudaGraphInstantiate(&gexec, graph, 0);
            if (graph) cudaGraphDestroy(graph);
          }
          h->ops->release(h);
          h->stream = user;
        }
        graph_kernels = h->launches - launches_host;
        h->itr = itr_host;             // nothing ran yet
        h->launches = launches_host;
        if (ce != cudaSuccess || rc != GBX_OK || !gexec) {   // no graphs on this system / for this handle: plain launches
          cudaGetLastError();
          drop_graph();
          graph_ok = false;
          rc = GBX_OK;
        }
      }
    }
    if (graph_ok && gexec && itr > 0 && nb == batch) {
      const int start = itr0 + itr;    // the replayed iterations are start + 1 ... start + nb
      h->h_state->it_now = start;
      cudaError_t ce = cudaMemcpyAsync(&h->st->it_now, &h->h_state->it_now, sizeof(int), cudaMemcpyHostToDevice, h->stream);
      for (int b = 0; b < nb / 2 && ce == cudaSuccess; ++b) ce = cudaGraphLaunch(gexec, h->stream);
      if (ce != cudaSuccess) {
        last_error() = std::string("gbx em_loop (graph replay): ") + cudaGetErrorString(ce);
        rc = GBX_ERR_CUDA;
        break;
      }
      h->itr += nb;
      h->launches += (nb / 2) * graph_kernels;   // the kernels the replays ran
    } else {
      for (int b = 1; b <= nb && rc == GBX_OK; ++b) {
        h->cur_it = itr0 + itr + b;
        rc = mod ? mod_iteration(h, nb == 1) : em_iteration(h, nb == 1);
      }
    }
    if (rc != GBX_OK) break;
    itr += nb;
    double conv;
    int done_at = 0;
    if (nb == 1) {
      rc = wait_residual(h, &conv);
      if (rc == GBX_OK && conv < thr) done_at = itr0 + itr;
    } else {
      cudaError_t ce = cudaMemcpyAsync(&h->h_state->done_at, &h->st->done_at, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream);
      if (ce != cudaSuccess) {
        last_error() = std::string("em_loop: ") + cudaGetErrorString(ce);
        rc = GBX_ERR_CUDA;
        break;
      }
      done_at = h->h_state->done_at;
    }
    if (done_at != 0) {
      const int extra = itr0 + itr - done_at;   // queued iterations that found the flag set and did nothing
      if (extra & 1) h->cur ^= 1;               // ... but the host flipped the message buffers for each of them
      if (h->pull) {
        h->gen -= extra;
        const bool dc = !mod && h->opt.degreecorrection;
        if (dc && extra > 0) {                  // ... and the theta generations (update_theta ran once per iteration)
          if (extra & 1) h->thcur ^= 1;
          h->th_used = h->thcur ^ 1;
        }
      }
      h->itr = done_at;
      *cnv = 1;
      *itrnum = done_at - itr0;
      break;
    }
  }
  drop_graph();
  h->cur_it = 0;
  h->cur_thr = -1.0;
  return rc;
}

}  // namespace

extern "C" {

int gbx_version(void) { return 100; }

const char* gbx_last_error(void) { return last_error().c_str(); }

int gbx_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    last_error() = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
    return GBX_ERR_CUDA;
  }
  return n;
}

void gbx_sbm_default_options(gbx_sbm_options* opt) {
  opt->degreecorrection = 1;
  opt->priorlearning = 1;
  opt->learningrate = 0.3;
  opt->maxCrs = 0.0;
  opt->itrmax = 256;
  opt->bp_conv_threshold = 0.000001;
  opt->damping = 1.0;
  opt->check_every = 1;
  opt->mstep_exact = 0;
}

// Pinned mirrors of SbmState are recycled too (cudaMallocHost costs about a millisecond).
static std::mutex g_pin_mutex;
static std::vector<SbmState*> g_pin_free;
static cudaError_t pinned_state_alloc(SbmState** p) {
  {
    std::lock_guard<std::mutex> g(g_pin_mutex);
    if (!g_pin_free.empty()) {
      *p = g_pin_free.back();
      g_pin_free.pop_back();
      return cudaSuccess;
    }
  }
  return cudaMallocHost((void**)p, sizeof(SbmState));
}
static void pinned_state_free(SbmState* p) {
  if (!p) return;
  std::lock_guard<std::mutex> g(g_pin_mutex);
  if (g_pin_free.size() < 64) g_pin_free.push_back(p);
  else cudaFreeHost(p);
}

__global__ void k_fill(double* __restrict__ p, long long n, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
static int fill_async(gbx_sbm* h, double* p, int64_t n, double v) {
  if (n <= 0) return GBX_OK;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)h->num_sms * 8);
  k_fill<<<grid, 256, 0, h->stream>>>(p, (long long)n, v);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

// GBX_TIMING=1: host wall time of the phases of gbx_*_create on stderr.
// slot -> vertex inside its tile (one block per tile, a thread per vertex of the tile)
__global__ void k_slot_lv(const int4* __restrict__ tiles, const int* __restrict__ rowptr, unsigned char* __restrict__ lv) {
  const int4 t = tiles[blockIdx.x];
  for (int v = threadIdx.x; v < t.y; v += blockDim.x)
    for (int e = rowptr[t.x + v]; e < rowptr[t.x + v + 1]; ++e) lv[e] = (unsigned char)v;
}
// degree of the neighbour of every slot: from the local row pointers (one rank) or the degrees by global id
__global__ void k_neighbour_degree(const int* __restrict__ col, const int* __restrict__ rowptr, const double* __restrict__ deg_global,
                                   long long K, float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const int j = col[e];
    out[e] = rowptr ? (float)(rowptr[j + 1] - rowptr[j]) : (float)deg_global[j];
  }
}

struct PhaseTimer {
  bool on = getenv("GBX_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[gbx create] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

static int graph_create(int model, gbx_sbm** out, int device, int64_t n, int64_t K, const int64_t* links, int q,
                        int rank = 0, int world = 1, bool stepped = false) {
  PhaseTimer pt;
  if (!out) return fail_invalid("out is NULL");
  *out = nullptr;
  if (q < 1 || q > MAXQ) return fail_invalid("q must be in 1..16");
  const SbmOps* ops = ops_for(model, q);
  if (!ops) return fail_invalid("this build of libgraphbix_cuda was compiled without that q");
  if (n < 1 || K < 0 || (K > 0 && !links)) return fail_invalid("bad n_vertices / n_links / links");
  if (n >= (int64_t)1 << 31 || K >= (int64_t)1 << 31) return fail_invalid("graph too large for 32-bit slots");
  if (world < 1 || world > MAXPEER || rank < 0 || rank >= world) return fail_invalid("bad rank / world (at most 8 ranks)");
  int ndev = gbx_device_count();
  if (ndev < 0) return ndev;
  if (ndev == 0) {
    last_error() = "no CUDA device: libgraphbix_cuda has no CPU fallback";
    return GBX_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return fail_invalid("device out of range");

  // ---- handle + device-side graph construction (gbx_graph_build.cu) ----
  gbx_sbm* h = new (std::nothrow) gbx_sbm();
  if (!h) return fail_invalid("out of host memory");
  h->device = device;
  h->q = q;
  h->qm = msg_stride(q);
  h->n = n;
  h->K = K;
  h->n_global = n;
  h->K_global = K;
  h->rank = rank;
  h->world = world;
  h->stepped = stepped;
  h->ops = ops;
  std::vector<int> rowptr_g;   // global row pointers
  std::vector<int> rowptr;     // this rank's (rebased)
  GlobalGraph gg;              // global columns / reverse slots, alive until the exchange plan is built
  // ---- tile schedule: runs of consecutive vertices with <= TILE_E slots and <= TV vertices; hubs apart.  Host work on
  // the row pointers alone: build_graph_device runs it while the GPU sorts the links. ----
  const int TV = tile_vertices(q);
  std::vector<int4> tiles;
  std::vector<int> hubs;
  const std::function<void()> schedule = [&]() {
    const int64_t n = h->n, K = h->K;   // this rank's vertices and slots
    if (h->vbeg == 0 && h->sbeg == 0 && (int64_t)rowptr_g.size() == n + 1) {
      rowptr = rowptr_g;   // one rank: the global row pointers are the local ones (rowptr_g is still read below)
    } else {
      rowptr.resize((size_t)n + 1);
      for (int64_t i = 0; i <= n; ++i) rowptr[(size_t)i] = rowptr_g[(size_t)(h->vbeg + i)] - (int)h->sbeg;
    }
    tiles.reserve((size_t)(K / (TILE_E / 2) + n / TV + 16));
    int64_t v = 0;
    while (v < n) {
      const int deg = rowptr[(size_t)v + 1] - rowptr[(size_t)v];
      if (deg > TILE_E) {
        hubs.push_back((int)v);
        ++v;
        continue;
      }
      int64_t w = v;
      int ne = 0;
      while (w < n && (w - v) < TV) {
        const int d = rowptr[(size_t)w + 1] - rowptr[(size_t)w];
        if (d > TILE_E || ne + d > TILE_E) break;
        ne += d;
        ++w;
      }
      tiles.push_back(make_int4((int)v, (int)(w - v), rowptr[(size_t)v], ne));
      v = w;
    }
    // degree statistics of the whole graph: the largest degree; for mod.jl handles constshift = sum_i d_i log d_i / L
    // (mod.jl:155) and the mean excess degree excm (mod.jl:185)
    double sdl = 0.0, sd2 = 0.0, sd = 0.0;
    for (int64_t u = 0; u < h->n_global; ++u) {
      const int di = rowptr_g[(size_t)u + 1] - rowptr_g[(size_t)u];
      h->max_degree = std::max(h->max_degree, di);
      if (model == 1) {
        const double d = (double)di;
        if (d > 0) sdl += d * std::log(d);
        sd2 += d * d;
        sd += d;
      }
    }
    h->constshift = h->K_global > 0 ? sdl / (0.5 * (double)h->K_global) : 0.0;
    h->excm = sd > 0 ? sd2 / sd - 1.0 : 0.0;  // (d'd)/(N mean(d)) - 1
  };
  {
    int major = 0, sms = 0;
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (ce != cudaSuccess) {
      last_error() = std::string("cudaSetDevice/cudaDeviceGetAttribute: ") + cudaGetErrorString(ce);
      delete h;
      return GBX_ERR_CUDA;
    }
    if (major < 10) {
      last_error() = "libgraphbix_cuda is built for sm_100a (B200) only";
      delete h;
      return GBX_ERR_CUDA;
    }
    h->num_sms = sms;
    pt.lap("device properties");
    const int rcb = build_graph_device(h, links, &rowptr_g, &gg, &schedule);
    pt.lap("build_graph_device");
    if (rcb != GBX_OK) {
      const std::string keep = last_error();
      gbx_sbm_destroy(h);
      last_error() = keep;
      return rcb;
    }
  }
  n = h->n;  // from here on: this rank's vertices and slots
  K = h->K;

  h->ntiles = (int)tiles.size();
  pt.lap("(tile schedule ran during the sort)");
  if (world > 1) {
    // pieces of a packed sweep: the exposed copy tail is 1/pieces of the boundary traffic, but many small all-to-all
    // copies run slower through the switch than a few large ones (measured: 2 GPUs best with 8, 8 GPUs with 2)
    h->nchunks = std::max(2, std::min((int)MAXCHUNK, 16 / world));
    if (const char* e = getenv("GBX_EXCHANGE_PIECES")) h->nchunks = atoi(e);
    // Pull sweeps exchange per-vertex records, not messages: no send / receive runs are needed for them.  The packed
    // message exchange of the push sweep is planned only where push is what will run (no pull kernels for this q, or
    // GBX_PULL=0); a damped run on a pull-capable q falls back to direct peer stores (exchange mode 0).
    const char* ep = getenv("GBX_PULL");
    const bool want_push_plan = !ops->pull_ok || (ep && atoi(ep) == 0);   // (world > 1 here: pull is the default)
    int rcp = GBX_OK;
    if (want_push_plan) rcp = build_exchange_plan(h, gg, tiles);
    else h->nchunks = 1;
    free_global(&gg);
    if (rcp != GBX_OK) {
      const std::string keep = last_error();
      gbx_sbm_destroy(h);
      last_error() = keep;
      return rcp;
    }
    h->exchange = want_push_plan ? 1 : 0;
  }
  h->model = model;
  gbx_mod_default_options(&h->mopt);
  h->nhubs = (int)hubs.size();
  gbx_sbm_default_options(&h->opt);
  auto body = [&]() -> int {
    // static per-slot data of the sweeps: the vertex (inside its tile) of every slot, and for sbm.jl the degree of every
    // slot's neighbour (theta_neighbour = degree / <degree of its MAP block>: the block arrives with the message)
    GBX_TRY(dev_alloc(h, &h->slot_lv, (size_t)K + 16));
    if (model == 0) GBX_TRY(dev_alloc(h, &h->coldeg, (size_t)K + 4));
    GBX_TRY(dev_alloc(h, &h->tiles, tiles.size()));
    GBX_TRY(dev_alloc(h, &h->hubs, hubs.size()));
    GBX_TRY(dev_alloc(h, &h->msg[0], (size_t)K * h->qm));
    GBX_TRY(dev_alloc(h, &h->msg[1], (size_t)K * h->qm));
    GBX_TRY(dev_alloc(h, &h->PSI, (size_t)n * q));
    GBX_TRY(dev_alloc(h, &h->theta, (size_t)h->n_global + MAXPEER));
    if (want_pull_layout(ops, world)) {  // pull sweeps: per-vertex records of ALL vertices (two generations; in a partitioned run every
                         // rank stores its vertices' records into every rank's copy), two theta generations of this rank's
                         // vertices, the degree of every slot's neighbour
      GBX_TRY(dev_alloc(h, &h->theta_pp[0], (size_t)n + 2));
      GBX_TRY(dev_alloc(h, &h->theta_pp[1], (size_t)n + 2));
      GBX_TRY(dev_alloc(h, &h->rec[0], (size_t)h->n_global * h->qm));
      GBX_TRY(dev_alloc(h, &h->rec[1], (size_t)h->n_global * h->qm));
      if (!h->coldeg) GBX_TRY(dev_alloc(h, &h->coldeg, (size_t)K + 4));
      GBX_CUDA_TRY(cudaMemsetAsync(h->rec[0], 0, (size_t)h->n_global * h->qm * sizeof(double), h->stream));
      GBX_CUDA_TRY(cudaMemsetAsync(h->rec[1], 0, (size_t)h->n_global * h->qm * sizeof(double), h->stream));
      h->recp[0][h->rank] = h->rec[0];
      h->recp[1][h->rank] = h->rec[1];
    }
    GBX_TRY(dev_alloc(h, &h->totals, (size_t)TOT_MAX));
    GBX_TRY(dev_alloc(h, &h->maxd_dev, 1));
    if (model == 1 || world > 1) {  // degree of every vertex by global id: the field of mod.jl's sweep and free energy;
                                    // the neighbour degrees of a partitioned run
      GBX_TRY(dev_alloc(h, &h->deg_global, (size_t)h->n_global));
      std::vector<double> dg((size_t)h->n_global);
      for (int64_t v = 0; v < h->n_global; ++v) dg[(size_t)v] = (double)(rowptr_g[(size_t)v + 1] - rowptr_g[(size_t)v]);
      GBX_CUDA_TRY(cudaMemcpy(h->deg_global, dg.data(), dg.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    h->msgp[0][h->rank] = h->msg[0];
    h->msgp[1][h->rank] = h->msg[1];
    GBX_TRY(dev_alloc(h, &h->pidx, (size_t)n));
    GBX_TRY(dev_alloc(h, &h->prior_tab, (size_t)(1 + (TILE_E + 1) * q) * (q + 4)));
    GBX_CUDA_TRY(cudaMemset(h->pidx, 0, (size_t)n * sizeof(int)));
    GBX_TRY(dev_alloc(h, &h->st, 1));
    GBX_TRY(dev_alloc(h, &h->params, 1));
    GBX_CUDA_TRY(cudaMemcpy(h->tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
    if (!tiles.empty()) k_slot_lv<<<(unsigned)tiles.size(), 64, 0, h->stream>>>(h->tiles, h->rowptr, h->slot_lv);
    if (h->coldeg && K > 0)
      k_neighbour_degree<<<(unsigned)std::min<int64_t>((K + 255) / 256, (int64_t)h->num_sms * 8), 256, 0, h->stream>>>(
          h->col, world > 1 ? nullptr : h->rowptr, h->deg_global, (long long)K, h->coldeg);
    GBX_CUDA_TRY(cudaGetLastError());
    GBX_CUDA_TRY(cudaMemcpy(h->hubs, hubs.data(), hubs.size() * sizeof(int), cudaMemcpyHostToDevice));
    GBX_CUDA_TRY(cudaMemset(h->st, 0, sizeof(SbmState)));
    GBX_CUDA_TRY(pinned_state_alloc(&h->h_state));
    GBX_CUDA_TRY(cudaEventCreateWithFlags(&h->ev_conv, cudaEventDisableTiming));
    // persistent grids: a whole number of CTAs per SM
    int occ = 1, occm = 1;
    GBX_TRY(ops->occupancy(1, &occ, &occm));
    occ = std::max(occ, 1);
    occm = std::max(occm, 1);
    h->grid_vertex = std::max(1, std::min(h->ntiles, h->num_sms * occ));
    if (const char* e = getenv("GBX_PULL_OCC")) occm = std::max(1, std::min(occm, atoi(e)));   // CTAs per SM (tuning)
    h->grid_pull = std::max(1, std::min(h->ntiles, h->num_sms * std::max(occm, 1)));
    const int64_t eblocks = std::max<int64_t>(1, (K / 2 + NT - 1) / NT);
    h->grid_mstep = h->grid_vertex + h->nhubs;  // M-step statistics: one row per sweep CTA
    h->grid_fe = (int)std::min<int64_t>(eblocks, (int64_t)h->num_sms * 2);
    h->grid_theta = (int)std::min<int64_t>(std::max<int64_t>(1, (n + NT - 1) / NT), (int64_t)h->num_sms * 4);
    h->grid_mstep = std::max(h->grid_vertex * h->nchunks, h->grid_pull) + h->nhubs;  // one row per sweep CTA of every piece
    GBX_TRY(dev_alloc(h, &h->part_vertex, (size_t)h->grid_mstep * PSTRIDE));
    GBX_TRY(dev_alloc(h, &h->part_mstep, (size_t)h->grid_mstep * q * q));
    GBX_TRY(dev_alloc(h, &h->part_theta, (size_t)h->grid_theta * MAXQ));
    GBX_TRY(dev_alloc(h, &h->part_fe, (size_t)h->grid_fe * 8));
    // everything above went into this handle's stream (the legacy stream at this point): wait for that, not for the
    // whole device -- another host thread may be recording a CUDA graph (em_loop), during which a device-wide
    // synchronisation is an error
    GBX_CUDA_TRY(cudaStreamSynchronize(h->stream));
    return GBX_OK;
  };
  const int rc = body();
  pt.lap("allocations + uploads");
  if (rc != GBX_OK) {
    const std::string keep = last_error();
    gbx_sbm_destroy(h);
    last_error() = keep;
    return rc;
  }
  *out = h;
  return GBX_OK;
}

int gbx_sbm_create(gbx_sbm** out, int device, int64_t n, int64_t K, const int64_t* links, int q) {
  return graph_create(0, out, device, n, K, links, q);
}

int gbx_trim_cache(void) {
  pool_trim(-1);
  return GBX_OK;
}
int64_t gbx_cached_bytes(int device) { return pool_cached_bytes(device); }

int gbx_sbm_destroy(gbx_sbm* h) {
  if (!h) return GBX_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);  // blocks go back to the cache: nothing of this handle may still be running
  if (h->ops && h->ops->release) h->ops->release(h);
  for (int k = 1; k < MAXPEER; ++k)
    if (h->peer_stream[k]) cudaStreamSynchronize(h->peer_stream[k]);
  for (int b = 0; b < 2; ++b)
    for (int r = 0; r < MAXPEER; ++r) {
      if (h->ipc_open[b][r]) cudaIpcCloseMemHandle(h->msgp[b][r]);
      if (h->ipc_open_recv[b][r]) cudaIpcCloseMemHandle(h->recvp[b][r]);
      if (h->ipc_open_rec[b][r]) cudaIpcCloseMemHandle(h->recp[b][r]);
    }
  dev_free(h->device, h->rev_pack);
  dev_free(h->device, h->unpack_slot);
  dev_free(h->device, h->sendbuf);
  dev_free(h->device, h->recvbuf[0]);
  dev_free(h->device, h->recvbuf[1]);
  for (int k = 2; k < MAXPEER; ++k)
    if (h->peer_stream[k]) cudaStreamDestroy(h->peer_stream[k]);
  for (int k = 0; k < MAXPEER; ++k)
    if (h->ev_peer[k]) cudaEventDestroy(h->ev_peer[k]);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (int c = 0; c < MAXCHUNK; ++c)
    if (h->ev_piece[c]) cudaEventDestroy(h->ev_piece[c]);
  if (h->ev_copied) cudaEventDestroy(h->ev_copied);
  if (h->ev_conv) cudaEventDestroy(h->ev_conv);
  if (h->ev_unpacked) cudaEventDestroy(h->ev_unpacked);
  dev_free(h->device, h->rowptr);
  dev_free(h->device, h->col);
  dev_free(h->device, h->rev);
  dev_free(h->device, h->erow);
  dev_free(h->device, h->slot_of_link);
  dev_free(h->device, h->tiles);
  dev_free(h->device, h->hubs);
  dev_free(h->device, h->msg[0]);
  dev_free(h->device, h->msg[1]);
  dev_free(h->device, h->PSI);
  if (h->theta_owned) dev_free(h->device, h->theta);
  dev_free(h->device, h->theta_pp[0]);
  dev_free(h->device, h->theta_pp[1]);
  dev_free(h->device, h->in_link);
  dev_free(h->device, h->out_link);
  dev_free(h->device, h->rec[0]);
  dev_free(h->device, h->rec[1]);
  dev_free(h->device, h->coldeg);
  dev_free(h->device, h->slot_lv);
  if (h->totals_owned) dev_free(h->device, h->totals);
  dev_free(h->device, h->maxd_dev);
  dev_free(h->device, h->deg_global);
  dev_free(h->device, h->pidx);
  dev_free(h->device, h->prior_tab);
  dev_free(h->device, h->st);
  dev_free(h->device, h->params);
  dev_free(h->device, h->part_vertex);
  dev_free(h->device, h->part_mstep);
  dev_free(h->device, h->part_theta);
  dev_free(h->device, h->part_fe);
  pinned_state_free(h->h_state);
  delete h;
  return GBX_OK;
}

int gbx_sbm_set_stream(gbx_sbm* h, void* s) {
  if (!h) return fail_invalid("handle is NULL");
  h->stream = (cudaStream_t)s;
  return GBX_OK;
}

int gbx_sbm_set_options(gbx_sbm* h, const gbx_sbm_options* opt) {
  if (!h || !opt) return fail_invalid("handle/options is NULL");
  if (!(opt->damping > 0.0 && opt->damping <= 1.0)) return fail_invalid("damping must be in (0,1]");
  if (!(opt->learningrate >= 0.0 && opt->learningrate <= 1.0)) return fail_invalid("learningrate must be in [0,1]");
  if (opt->itrmax < 0) return fail_invalid("itrmax must be >= 0");
  if (opt->mstep_exact && h->stepped) return fail_invalid("mstep_exact is not available on partitioned handles");
  h->opt = *opt;
  if (h->opt.check_every < 1) h->opt.check_every = 1;
  return GBX_OK;
}

int gbx_sbm_init(gbx_sbm* h, const double* gr_init, const double* Crs_init, const double* psicav, uint64_t seed) {
  if (!h || !Crs_init) return fail_invalid("handle/Crs_init is NULL");
  if (h->model != 0) return fail_invalid("not a gbx_sbm handle");
  if (h->opt.priorlearning && !gr_init) return fail_invalid("gr_init is NULL with priorlearning on");
  GBX_TRY(set_device(h));
  const int q = h->q;
  SbmState s;
  memset(&s, 0, sizeof s);
  h->fail = 0;
  for (int r = 0; r < q; ++r)
    for (int t = 0; t < q; ++t) {
      // sbm.jl:308-317; the matrix is symmetrised (every initial condition of the reference is symmetric)
      double c = 0.5 * (Crs_init[r * q + t] + Crs_init[t * q + r]);
      if (std::isnan(c) || std::isinf(c)) h->fail = 1;
      else if (c < 1e-6) c = 1e-6;
      s.C[r * q + t] = c;
      s.Cused[r * q + t] = c;
    }
  for (int r = 0; r < q; ++r) s.gr[r] = h->opt.priorlearning ? gr_init[r] : 1.0 / q;  // sbm.jl:304-306
  GBX_CUDA_TRY(cudaMemcpyAsync(h->st, &s, sizeof s, cudaMemcpyHostToDevice, h->stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->stream));  // `s` is a stack object
  h->cur = 0;
  h->itr = 0;
  h->theta_valid = false;
  GBX_TRY(choose_layout(h, h->opt.damping));
  GBX_TRY(h->ops->init_model(h));  // theta = 1: row 0 of the prior table
  GBX_TRY(load_messages(h, psicav, seed));
  {
    GBX_TRY(fill_async(h, h->theta, h->n_global, 1.0));  // theta = ones(Ntot,1), sbm.jl:277
    if (h->theta_pp[0]) GBX_TRY(fill_async(h, h->theta_pp[0], h->n, 1.0));
    if (h->theta_pp[1]) GBX_TRY(fill_async(h, h->theta_pp[1], h->n, 1.0));
  }
  if (!h->fail && !h->stepped) {
    // PSI = updatePSI() with h = 0, then gr = update_gr(PSI): sbm.jl:340-341
    // (a partitioned run does this through gbx_sbm_step so that the ranks can all-reduce the totals in between)
    GBX_TRY(h->ops->prepare(h, 1, 0));
    GBX_TRY(h->ops->vertex_pass(h, MODE_MARGINAL));
    GBX_TRY(h->ops->local_totals(h, MODE_MARGINAL));
    GBX_TRY(h->ops->apply_totals(h, MODE_MARGINAL));
  }
  GBX_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->initialised = true;
  return GBX_OK;
}

int gbx_sbm_iterate(gbx_sbm* h, int n_iters, double* bpconv_out) {
  if (!h) return fail_invalid("handle is NULL");
  if (!h->initialised) return fail_state();
  if (h->stepped && n_iters > 0) return fail_invalid("partitioned handle: drive the loop with gbx_sbm_step");
  GBX_TRY(set_device(h));
  for (int i = 0; i < n_iters; ++i) GBX_TRY(em_iteration(h));
  if (bpconv_out) {
    GBX_TRY(read_state(h));
    *bpconv_out = h->h_state->conv;
  }
  return GBX_OK;
}

int gbx_sbm_bp_sweeps(gbx_sbm* h, int n_sweeps) {
  if (!h) return fail_invalid("handle is NULL");
  if (!h->initialised) return fail_state();
  GBX_TRY(set_device(h));
  GBX_TRY(h->ops->prepare(h, 0, 0));
  for (int i = 0; i < n_sweeps; ++i) {
    if (h->pull && i > 0) GBX_TRY(h->ops->prepare(h, 0, 0));   // refresh "the parameters of the previous sweep"
    GBX_TRY(h->ops->vertex_pass(h, MODE_SWEEP));
    h->cur ^= 1;
    if (h->pull) {  // no M-step follows that would note what this sweep ran with (sbm_apply_totals / mod_apply_totals)
      GBX_CUDA_TRY(cudaMemcpyAsync(h->st->Cused, h->st->C, sizeof(h->st->C), cudaMemcpyDeviceToDevice, h->stream));
      GBX_CUDA_TRY(cudaMemcpyAsync(&h->st->eb1_used, &h->st->eb1_cur, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
  }
  return GBX_OK;
}

int gbx_sbm_finalize(gbx_sbm* h, gbx_sbm_result* res) {
  if (!h || !res) return fail_invalid("handle/result is NULL");
  if (!h->initialised) return fail_state();
  GBX_TRY(set_device(h));
  if (!h->fail && !h->stepped) {
    GBX_TRY(h->ops->prepare(h, 0, 1));  // h = update_h(); gr = update_gr(PSI)   sbm.jl:364-365
    GBX_TRY(free_energy_steps(h));      // freeenergy()                          sbm.jl:366
  }
  // (a partitioned run has driven the same steps through gbx_sbm_step, with the ranks' all-reduces in between)
  memset(res, 0, sizeof *res);
  res->fail = h->fail;
  GBX_TRY(read_state(h));
  const SbmState* s = h->h_state;
  res->FE = s->result[0];
  res->CVBayes = s->result[1];
  res->CVGP = s->result[2];
  res->CVGT = s->result[3];
  res->CVMAP = s->result[4];
  res->varCVBayes = s->result[5];
  res->varCVGP = s->result[6];
  res->varCVGT = s->result[7];
  res->varCVMAP = s->result[8];
  res->BPconv = s->conv;
  res->max_dpsi = s->maxd;
  res->nonfinite = s->status & 1;
  res->itrs_done = h->itr;
  return GBX_OK;
}

int gbx_sbm_em(gbx_sbm* h, gbx_sbm_result* res) {
  if (!h || !res) return fail_invalid("handle/result is NULL");
  if (!h->initialised) return fail_state();
  if (h->stepped) return fail_invalid("partitioned handle: drive the loop with gbx_sbm_step (see engine_dist.py)");
  GBX_TRY(set_device(h));
  int cnv = 0, itrnum = 0;
  if (!h->fail) GBX_TRY(em_loop(h, h->opt.itrmax, h->opt.check_every, h->opt.bp_conv_threshold, false, &cnv, &itrnum));
  GBX_TRY(gbx_sbm_finalize(h, res));
  res->cnv = cnv;
  res->itrnum = itrnum;
  return GBX_OK;
}

int gbx_sbm_get_state(gbx_sbm* h, double* gr, double* Crs, double* PSI, double* theta, double* hfield) {
  if (!h) return fail_invalid("handle is NULL");
  GBX_TRY(set_device(h));
  GBX_TRY(read_state(h));
  const int q = h->q;
  if (gr) memcpy(gr, h->h_state->gr, q * sizeof(double));
  if (Crs) memcpy(Crs, h->h_state->C, (size_t)q * q * sizeof(double));
  if (hfield) memcpy(hfield, h->h_state->h, q * sizeof(double));
  if (theta) {
    const double* src = h->pull ? h->theta_pp[h->thcur] : h->theta + h->vbeg;
    GBX_CUDA_TRY(cudaMemcpyAsync(theta, src, (size_t)h->n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  // the big one last: a pageable PSI (a Julia Array) is filled through pinned chunks by several host threads
  if (PSI) GBX_CUDA_TRY(download_sync(h->device, PSI, h->PSI, (size_t)h->n * q * sizeof(double), h->stream));
  else GBX_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return GBX_OK;
}

int gbx_sbm_get_messages(gbx_sbm* h, double* psicav) {
  if (!h || !psicav) return fail_invalid("handle/psicav is NULL");
  GBX_TRY(set_device(h));
  // link order, K_global rows; on a partitioned run the rows of links owned by other ranks come back as zeros
  const size_t bytes = (size_t)h->K_global * h->q * sizeof(double);
  double* tmp = h->msg[h->cur ^ 1];  // the other buffer is scratch between sweeps (single rank, push layout)
  bool own = false;
  if (h->world > 1 || h->pull) {
    GBX_CUDA_TRY(dev_malloc(h->device, (void**)&tmp, bytes ? bytes : 1, h->world == 1));
    own = true;
  }
  int rc = h->pull ? h->ops->pull_export(h, tmp) : h->ops->permute(h, h->msg[h->cur], tmp, 0);
  cudaError_t e = cudaSuccess;
  if (rc == GBX_OK) e = download_sync(h->device, psicav, tmp, bytes, h->stream);
  else e = cudaStreamSynchronize(h->stream);
  if (own) dev_free(h->device, tmp);
  if (rc != GBX_OK) return rc;
  GBX_CUDA_TRY(e);
  return GBX_OK;
}

int gbx_sbm_get_assignment(gbx_sbm* h, int32_t* block) {
  if (!h || !block) return fail_invalid("handle/block is NULL");
  GBX_TRY(set_device(h));
  int* d = nullptr;
  GBX_CUDA_TRY(cudaMalloc((void**)&d, (size_t)h->n * sizeof(int)));
  int rc = h->ops->assignment(h, d);
  cudaError_t e = cudaSuccess;
  if (rc == GBX_OK) {
    e = cudaMemcpyAsync(block, d, (size_t)h->n * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  }
  cudaFree(d);
  if (rc != GBX_OK) return rc;
  GBX_CUDA_TRY(e);
  return GBX_OK;
}

int64_t gbx_sbm_launch_count(const gbx_sbm* h) { return h ? h->launches : 0; }
int64_t gbx_sbm_num_links(const gbx_sbm* h) { return h ? h->K : 0; }
int64_t gbx_sbm_device_bytes(const gbx_sbm* h) { return h ? h->dev_bytes : 0; }

int gbx_sbm_sweep_kind(const gbx_sbm* h) { return h ? h->pull : 0; }

double gbx_sbm_sweep_bytes(const gbx_sbm* h) {
  if (!h) return 0.0;
  const double q = h->q, K = (double)h->K, n = (double)h->n;
  // what the sweep kernel streams: per directed edge the message read and the message written, rev and,
  // with degree correction, one staged theta_row * theta_col scale; per vertex the marginal read and written, the
  // row pointer and the row of the prior table
  const bool dc = h->model == 1 ? false : h->opt.degreecorrection != 0;
  const double damping = h->model == 1 ? h->mopt.damping : h->opt.damping;
  if (h->pull) {
    // pull sweep: per directed edge its own message read and rewritten in place, the neighbour's id and (degree
    // correction) degree; per vertex the marginal read and written, its record written once and fetched from HBM once
    // (the other deg - 1 gathers of it are L2 hits), row pointer, prior-table row, two theta generations
    const double pe = 2.0 * q * 8.0 + 4.0 + (dc ? 4.0 : 0.0);
    const double pv = 2.0 * q * 8.0 + 2.0 * q * 8.0 + 4.0 + 4.0 + (dc ? 16.0 : 0.0);
    return K * pe + n * pv;
  }
  double per_edge = 2.0 * q * 8.0 + 4.0 + 1.0;          // message read, message written, rev, slot -> vertex byte
  if (dc) per_edge += 4.0;                              // degree of the neighbour
  if (damping != 1.0) per_edge += q * 8.0;             // old outgoing message
  const double per_vertex = 2.0 * q * 8.0 + 4.0 + 4.0 + (dc ? 8.0 : 0.0);   // marginal r/w, row pointer, prior row, theta
  return K * per_edge + n * per_vertex;
}

// ---------------------------------------------------------------------------------------------
// mod.jl
// ---------------------------------------------------------------------------------------------
void gbx_mod_default_options(gbx_mod_options* opt) {
  opt->dc = 1;
  opt->learning = 1;
  opt->priorlearning = 0;
  opt->learningrate = 0.3;
  opt->maxomega = 1.0;
  opt->itrmax = 256;
  opt->betafrac = 0.0;
  opt->bp_conv_threshold = 0.000001;
  opt->damping = 1.0;
  opt->check_every = 1;
}

int gbx_mod_create(gbx_mod** out, int device, int64_t n, int64_t K, const int64_t* links, int q) {
  return graph_create(1, out, device, n, K, links, q);
}
int gbx_mod_destroy(gbx_mod* h) { return gbx_sbm_destroy(h); }
int gbx_mod_set_stream(gbx_mod* h, void* s) { return gbx_sbm_set_stream(h, s); }

int gbx_mod_set_options(gbx_mod* h, const gbx_mod_options* opt) {
  if (!h || !opt) return fail_invalid("handle/options is NULL");
  if (h->model != 1) return fail_invalid("not a gbx_mod handle");
  if (!(opt->damping > 0.0 && opt->damping <= 1.0)) return fail_invalid("damping must be in (0,1]");
  if (!(opt->learningrate >= 0.0 && opt->learningrate <= 1.0)) return fail_invalid("learningrate must be in [0,1]");
  if (opt->itrmax < 0) return fail_invalid("itrmax must be >= 0");
  h->mopt = *opt;
  if (h->mopt.check_every < 1) h->mopt.check_every = 1;
  return GBX_OK;
}

int gbx_mod_init(gbx_mod* h, const double* gr, const double* psicav, uint64_t seed) {
  if (!h || !gr) return fail_invalid("handle/gr is NULL");
  if (h->model != 1) return fail_invalid("not a gbx_mod handle");
  GBX_TRY(set_device(h));
  const int q = h->q;
  SbmState s;
  memset(&s, 0, sizeof s);
  for (int r = 0; r < q; ++r) s.gr[r] = gr[r];
  // mod.jl:185-196
  const double excm = h->excm, B = (double)q;
  const double beta0 = std::log(1.0 + B / (excm - 1.0));
  const double betaast = std::log(1.0 + B / (std::sqrt(excm) - 1.0));
  s.alpha = 1.0 / (double)h->K_global;
  s.beta = betaast - h->mopt.betafrac * (betaast - beta0);
  s.omegain = std::exp(s.beta) * s.alpha * s.beta / (std::exp(s.beta) - 1.0);
  s.omegaout = s.alpha * s.beta / (std::exp(s.beta) - 1.0);
  GBX_CUDA_TRY(cudaMemcpyAsync(h->st, &s, sizeof s, cudaMemcpyHostToDevice, h->stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->cur = 0;
  h->itr = 0;
  h->fail = 0;
  GBX_TRY(choose_layout(h, h->mopt.damping));
  GBX_TRY(h->ops->init_model(h));  // prior-table row per vertex: by degree when degree corrected
  GBX_TRY(load_messages(h, psicav, seed));
  if (!h->stepped) {
    // h = zeros(1,B); PSI = updatePSI(): mod.jl:204-206
    GBX_TRY(h->ops->prepare(h, 1, 0));
    GBX_TRY(h->ops->vertex_pass(h, MODE_MARGINAL));
    GBX_TRY(h->ops->local_totals(h, MODE_MARGINAL));
    GBX_TRY(h->ops->apply_totals(h, MODE_MARGINAL));
  }
  GBX_CUDA_TRY(cudaStreamSynchronize(h->stream));
  h->initialised = true;
  return GBX_OK;
}

}  // extern "C"

namespace {
int mod_iteration(gbx_sbm* h, bool want_residual) {
  const SbmOps* op = h->ops;
  GBX_TRY(op->prepare(h, 0, 0));              // h = update_h()                                    mod.jl:56
  GBX_TRY(op->vertex_pass(h, MODE_SWEEP));    // messages, marginals, residual, omega statistic     mod.jl:57-88
  h->cur ^= 1;
  GBX_TRY(op->local_totals(h, MODE_SWEEP));
  GBX_TRY(op->apply_totals(h, MODE_SWEEP));   // conv; gr = update_gr(PSI); omegas, alpha, beta     mod.jl:221-229
  if (want_residual) GBX_TRY(queue_residual(h));
  h->itr++;
  return GBX_OK;
}
}  // namespace

extern "C" {

int gbx_mod_iterate(gbx_mod* h, int n_iters, double* bpconv_out) {
  if (!h) return fail_invalid("handle is NULL");
  if (!h->initialised || h->model != 1) return fail_state();
  if (h->stepped && n_iters > 0) return fail_invalid("partitioned handle: drive the loop with gbx_sbm_step");
  GBX_TRY(set_device(h));
  for (int i = 0; i < n_iters; ++i) GBX_TRY(mod_iteration(h, false));
  if (bpconv_out) {
    GBX_TRY(read_state(h));
    *bpconv_out = h->h_state->conv;
  }
  return GBX_OK;
}

int gbx_mod_finalize(gbx_mod* h, gbx_mod_result* res) {
  if (!h || !res) return fail_invalid("handle/result is NULL");
  if (!h->initialised || h->model != 1) return fail_state();
  GBX_TRY(set_device(h));
  memset(res, 0, sizeof *res);
  if (!h->stepped) {
    GBX_TRY(h->ops->prepare(h, 0, 0));   // h = update_h()   mod.jl:240
    GBX_TRY(free_energy_steps(h));       // freeenergy()     mod.jl:241
  }
  GBX_TRY(read_state(h));
  const SbmState* s = h->h_state;
  res->excm = h->excm;
  res->alpha = s->alpha;
  res->beta = s->beta;
  res->omegain = s->omegain;
  res->omegaout = s->omegaout;
  res->FE = s->result[0];
  res->CVBayes = s->result[1];
  res->CVGP = s->result[2];
  res->CVGT = s->result[3];
  res->CVMAP = s->result[4];
  res->varCVBayes = s->result[5];
  res->varCVGP = s->result[6];
  res->varCVGT = s->result[7];
  res->varCVMAP = s->result[8];
  res->BPconv = s->conv;
  res->max_dpsi = s->maxd;
  res->nonfinite = s->status & 1;
  res->itrs_done = h->itr;
  return GBX_OK;
}

int gbx_mod_em(gbx_mod* h, gbx_mod_result* res) {
  if (!h || !res) return fail_invalid("handle/result is NULL");
  if (!h->initialised || h->model != 1) return fail_state();
  if (h->stepped) return fail_invalid("partitioned handle: drive the loop with gbx_sbm_step");
  GBX_TRY(set_device(h));
  int cnv = 0, itrnum = 0;
  GBX_TRY(em_loop(h, h->mopt.itrmax, h->mopt.check_every, h->mopt.bp_conv_threshold, true, &cnv, &itrnum));
  GBX_TRY(gbx_mod_finalize(h, res));
  res->cnv = cnv;
  res->itrnum = itrnum;
  return GBX_OK;
}

int gbx_mod_get_state(gbx_mod* h, double* gr, double* PSI, double* hfield, double* params) {
  if (!h) return fail_invalid("handle is NULL");
  GBX_TRY(gbx_sbm_get_state(h, gr, nullptr, PSI, nullptr, hfield));
  if (params) {
    params[0] = h->h_state->alpha;
    params[1] = h->h_state->beta;
    params[2] = h->h_state->omegain;
    params[3] = h->h_state->omegaout;
  }
  return GBX_OK;
}
int gbx_mod_get_messages(gbx_mod* h, double* psicav) { return gbx_sbm_get_messages(h, psicav); }
int gbx_mod_get_assignment(gbx_mod* h, int32_t* block) { return gbx_sbm_get_assignment(h, block); }
int64_t gbx_mod_launch_count(const gbx_mod* h) { return gbx_sbm_launch_count(h); }

// ---------------------------------------------------------------------------------------------
// Vertex-partitioned runs: one process per GPU, rank r owns the vertices [r*ceil(n/world), ...) and their rows.
// ---------------------------------------------------------------------------------------------
int gbx_sbm_create_part(gbx_sbm** out, int device, int64_t n, int64_t K, const int64_t* links, int q, int rank,
                        int world) {
  return graph_create(0, out, device, n, K, links, q, rank, world, true);
}
int gbx_mod_create_part(gbx_mod** out, int device, int64_t n, int64_t K, const int64_t* links, int q, int rank,
                        int world) {
  return graph_create(1, out, device, n, K, links, q, rank, world, true);
}

int gbx_sbm_part_info(const gbx_sbm* h, int64_t* info) {
  if (!h || !info) return fail_invalid("handle/info is NULL");
  info[0] = h->vbeg;
  info[1] = h->n;
  info[2] = h->sbeg;
  info[3] = h->K;
  info[4] = TOT_MAX;
  info[5] = (h->n_global + h->world - 1) / h->world;  // vertices per rank (the last rank may own fewer)
  info[6] = h->rank;
  info[7] = h->world;
  return GBX_OK;
}

int gbx_sbm_ipc_export(gbx_sbm* h, int which, unsigned char* handle64) {
  if (!h || !handle64 || which < 0 || which > 5) return fail_invalid("bad arguments");
  if (which >= 2 && which < 4 && !h->recvbuf[which - 2]) return fail_invalid("no receive buffers (single-rank handle)");
  if (which >= 4 && !h->rec[which - 4]) return fail_invalid("no record buffers (this q has no pull kernels)");
  GBX_TRY(set_device(h));
  cudaIpcMemHandle_t mh;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  GBX_CUDA_TRY(cudaIpcGetMemHandle(&mh, which < 2 ? (void*)h->msg[which] : which < 4 ? (void*)h->recvbuf[which - 2] : (void*)h->rec[which - 4]));
  memcpy(handle64, &mh, 64);
  return GBX_OK;
}

int gbx_sbm_ipc_import(gbx_sbm* h, int which, int rank, const unsigned char* handle64) {
  if (!h || !handle64 || which < 0 || which > 5 || rank < 0 || rank >= h->world) return fail_invalid("bad arguments");
  if (rank == h->rank) return GBX_OK;
  GBX_TRY(set_device(h));
  cudaIpcMemHandle_t mh;
  memcpy(&mh, handle64, 64);
  void* p = nullptr;
  GBX_CUDA_TRY(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
  if (which < 2) {
    h->msgp[which][rank] = (double*)p;
    h->ipc_open[which][rank] = true;
  } else if (which < 4) {
    h->recvp[which - 2][rank] = (double*)p;
    h->ipc_open_recv[which - 2][rank] = true;
  } else {
    h->recp[which - 4][rank] = (double*)p;
    h->ipc_open_rec[which - 4][rank] = true;
  }
  return GBX_OK;
}

int gbx_sbm_part_set_buffers(gbx_sbm* h, void* totals_dev, void* theta_dev) {
  if (!h) return fail_invalid("handle is NULL");
  GBX_TRY(set_device(h));
  if (totals_dev) {
    if (h->totals_owned) dev_free(h->device, h->totals);   // came from dev_malloc (possibly the block cache)
    h->totals = (double*)totals_dev;
    h->totals_owned = false;
  }
  if (theta_dev) {
    if (h->theta_owned) dev_free(h->device, h->theta);
    h->theta = (double*)theta_dev;
    h->theta_owned = false;
  }
  return GBX_OK;
}

int gbx_sbm_part_set_exchange(gbx_sbm* h, int mode) {
  if (!h) return fail_invalid("handle is NULL");
  if (mode != 0 && mode != 1) return fail_invalid("exchange mode must be 0 (direct peer stores) or 1 (packed)");
  if (mode == 1 && h->world > 1 && !h->rev_pack) {
    if (!h->ops->pull_ok) return fail_invalid("no exchange plan");
    mode = 0;  // the packed exchange was not planned because this q sweeps with the pull kernels (records, not messages)
  }
  if (h->pending_unpack) return fail_state();
  h->exchange = h->world > 1 ? mode : 0;
  return GBX_OK;
}

int gbx_sbm_step(gbx_sbm* h, int step, int arg) {
  if (!h) return fail_invalid("handle is NULL");
  if (!h->initialised) return fail_state();
  const bool no_messages = step == GBX_STEP_PREPARE || step == GBX_STEP_APPLY_TOTALS || step == GBX_STEP_THETA_LOCAL ||
                           step == GBX_STEP_THETA_APPLY || step == GBX_STEP_ESCALE || step == GBX_STEP_COUNT_ITERATION ||
                           step == GBX_STEP_UNPACK;
  GBX_TRY(set_device(h, !no_messages));
  const SbmOps* op = h->ops;
  switch (step) {
    case GBX_STEP_PREPARE: return op->prepare(h, arg & 1, (arg >> 1) & 1);
    case GBX_STEP_VERTEX_PASS: {
      GBX_TRY(op->vertex_pass(h, arg));
      if (arg == MODE_SWEEP) h->cur ^= 1;
      return GBX_OK;
    }
    case GBX_STEP_LOCAL_TOTALS: return op->local_totals(h, arg);
    case GBX_STEP_APPLY_TOTALS: return op->apply_totals(h, arg);
    case GBX_STEP_THETA_LOCAL: return op->theta_local(h);
    case GBX_STEP_THETA_APPLY: return op->theta_apply(h);
    case GBX_STEP_ESCALE: return op->escale(h);
    case GBX_STEP_FE_LOCAL: return op->fe_local(h, arg);
    case GBX_STEP_FE_APPLY: return op->fe_apply(h, arg);
    case GBX_STEP_FE_FINAL: return op->fe_final(h);
    case GBX_STEP_COUNT_ITERATION: h->itr++; return GBX_OK;
    case GBX_STEP_UNPACK: return op->unpack(h);
    default: return fail_invalid("unknown step");
  }
}

int gbx_sbm_read_residual(gbx_sbm* h, double* bpconv) {
  if (!h || !bpconv) return fail_invalid("handle/bpconv is NULL");
  GBX_TRY(set_device(h));
  GBX_TRY(read_state(h));
  *bpconv = h->h_state->conv;
  return GBX_OK;
}

}  // extern "C"
