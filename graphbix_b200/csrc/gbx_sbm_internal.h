// Internal: the host-side handle behind `gbx_sbm*` and the per-q launcher table.
// One translation unit per q (gbx_sbm_q.cu, -DGBX_Q=<q>) fills a table; gbx_sbm_host.cu dispatches.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/graphbix_cuda.h"
#include "gbx_common.cuh"

namespace gbx {

// Parameters every edge/vertex kernel reads; refreshed once per EM iteration (device -> __constant__).
struct SbmParams {
  double C[MAXQ * MAXQ];     // Crs, row-major with stride q
  double logC[MAXQ * MAXQ];  // log(Crs)
  double lgr[MAXQ];          // log(gr) after the clamp of sbm.jl:89-96
  double h[MAXQ];            // external field, sbm.jl:70-73
  double logN;
  double eb1;                // e^beta - 1 (mod.jl)
  double oin, oout;          // omega_in, omega_out (mod.jl)
  // pull sweeps (gbx_sbm_pull.cuh) rebuild the message a neighbour emitted in the PREVIOUS sweep: the parameters that
  // sweep ran with, and the reciprocal block degree means that turn a neighbour's (degree, MAP block) into its theta
  double Cprev[MAXQ * MAXQ];
  double eb1_prev;
  double invdr[MAXQ];
  int safe_deg;              // largest degree whose plain product of edge factors cannot leave the double range
  int pad_;
};

// Device-resident EM state (small; one per handle).
struct SbmState {
  double gr[MAXQ];
  double C[MAXQ * MAXQ];
  double h[MAXQ];
  double sumPSI[MAXQ];   // sum_i PSI_i of the marginals currently in PSI
  double Stheta[MAXQ];   // sum_i theta_i PSI_i
  double drmean[MAXQ];   // mean degree of each MAP block (update_theta, sbm.jl:47-58)
  double invdr[MAXQ];    // 1 / drmean: theta_i = degree_i * invdr[MAP block of i] wherever a theta is formed
  double Cused[MAXQ * MAXQ];  // Crs the last sweep ran with (pull sweeps rebuild that sweep's messages)
  double eb1_cur, eb1_used;   // e^beta - 1 of the current parameters / of the last sweep (mod.jl)
  double conv, maxd, sumlogZi;
  double cvsum[4], cvmean[4], cvss[4];
  double result[12];
  // mod.jl (community-restricted SBM) scalars
  double omegain, omegaout, alpha, beta, hn2, omegainnew2;
  double cv5sum[5], cv5mean[5], cv5ss[5];
  int status;  // bit 0: non-finite / non-positive value met in a kernel
  int tot_count; // blocks of sbm_local_totals that have delivered their total (the last one applies the totals)
  int it_now;  // iteration counter of launches replayed from a captured CUDA graph (stamp -1), see skip_iteration
  int done_at; // EM iteration whose residual fell below the threshold (0: not yet).  The kernels of LATER iterations
               // return at once, so the host may queue several iterations before it reads the residual back and the
               // state still is exactly that of the converged iteration (sbm.jl:348-353 breaks after update_theta).
};

constexpr int PSTRIDE = 4 * MAXQ + 4;
constexpr int MAXPEER = 8;                       // ranks of one partitioned run (one NVLink domain)
constexpr int REV_SHIFT = 29;                    // rev[e] = owner << 29 | slot in the owner's buffer (3 + 29 bits)
constexpr int REV_MASK = (1 << REV_SHIFT) - 1;
constexpr int MAXCHUNK = 8;                      // a partitioned sweep is launched in <= 8 pieces so that copies overlap it
constexpr int TOT_MAX = 4 + 4 * MAXQ + MAXQ * MAXQ;  // doubles in a totals vector  // per-CTA partial row of the vertex kernels
enum { MODE_SWEEP = 0, MODE_MARGINAL = 1, MODE_LOGZ = 2 };

// Doubles between consecutive messages in the message / send / receive buffers.  Rows are padded to a whole number of
// 32-byte sectors (q = 1..3 -> 4, 5..7 -> 8, 9..11 -> 12, 13..15 -> 16; pads are written as zeros): a 40-byte row
// stored at a random place touches sectors it shares with its neighbours, and partial-sector writes cost a
// read-modify-write in HBM (unpadded, q = 5 ran 1.7x slower than q = 8 and q = 2 slower than q = 4).
constexpr int msg_stride(int q) { return (q + 3) / 4 * 4; }

constexpr int tile_vertices(int q) { return (q <= 8) ? TILE_V / 2 : TILE_V / 4; }

struct SbmOps;

// gbx_pool.cu: device-memory block cache shared by the handles of a process.
cudaError_t dev_malloc(int device, void** p, size_t bytes, bool pooled);
void dev_free(int device, void* p);
void pool_trim(int device);              // device < 0: every device
int64_t pool_cached_bytes(int device);
// Host -> device copy ordered on `st`; large pageable sources go through pinned chunks filled by several host threads.
cudaError_t upload_async(int device, void* dst, const void* src, size_t bytes, cudaStream_t st);
// ... and many pieces at once (the edge lists of a batch of graphs: each too small to be worth staging on its own).
struct UploadSeg {
  void* dst;
  const void* src;
  size_t bytes;
};
cudaError_t upload_segments_async(int device, const UploadSeg* segs, size_t nseg, cudaStream_t st);
// Device -> host copy after what `st` holds; returns when dst is filled (and `st` is idle).  Same staging for pageable dst.
cudaError_t download_sync(int device, void* dst, const void* src, size_t bytes, cudaStream_t st);

// Device copies of the global columns / reverse slots, kept between the graph build and the exchange plan.
struct GlobalGraph {
  int* col = nullptr;
  int* rev = nullptr;
};

}  // namespace gbx

struct gbx_sbm {
  int device = 0, q = 0;
  int qm = 0;                     // gbx::msg_stride(q): doubles per message row in msg / sendbuf / recvbuf
  int64_t n = 0, K = 0;           // vertices / directed-edge slots owned by this handle (all of them when world == 1)
  int64_t n_global = 0, K_global = 0;
  int rank = 0, world = 1;
  bool stepped = false;           // created by gbx_*_create_part: the host drives every step (gbx_sbm_step)
  int64_t vbeg = 0, sbeg = 0;     // first global vertex / first global slot of this rank
  int64_t vbeg_all[gbx::MAXPEER + 1] = {0}, sbeg_all[gbx::MAXPEER + 1] = {0};
  cudaStream_t stream = nullptr;
  gbx_sbm_options opt{};
  int model = 0;                  // 0: sbm.jl, 1: mod.jl (same handle type behind gbx_mod*)
  gbx_mod_options mopt{};
  double constshift = 0.0;        // sum_i d_i log d_i / L (mod.jl:155), 0 without degree correction
  double excm = 0.0;              // mean excess degree (mod.jl:185)
  bool initialised = false;
  int fail = 0;
  int64_t launches = 0, dev_bytes = 0;
  int num_sms = 148;
  const gbx::SbmOps* ops = nullptr;
  // graph (device)
  int *rowptr = nullptr, *col = nullptr, *rev = nullptr, *slot_of_link = nullptr, *hubs = nullptr;
  int* erow = nullptr;            // per edge slot: its row (destination vertex)
  int4* tiles = nullptr;          // (first vertex, #vertices, first edge slot, #edge slots)
  int ntiles = 0, nhubs = 0, max_degree = 1;
  double* deg_global = nullptr;   // degree of every vertex, by global id
  // state (device)
  double *msg[2] = {nullptr, nullptr}, *PSI = nullptr;
  double* msgp[2][gbx::MAXPEER] = {};   // both message buffers of every rank (peer-mapped; [b][rank] == msg[b])
  bool ipc_open[2][gbx::MAXPEER] = {};
  double* theta = nullptr;              // indexed by GLOBAL vertex id (all ranks' values after the all-gather)
  bool theta_owned = true;              // false: the caller provided the buffer (partitioned runs)
  double* totals = nullptr;             // per-pass totals vector (device); caller-provided in partitioned runs
  bool totals_owned = true;
  double* maxd_dev = nullptr;
  int* pidx = nullptr;            // per vertex: row of the prior table (1 + degree * q + MAP block; 0: theta = 1)
  double* prior_tab = nullptr;    // prior factors per (degree, block), rebuilt every EM iteration
  bool theta_valid = false;       // update_theta has run since init
  bool totals_applied = false;    // local_totals has applied the totals itself (one rank): apply_totals has nothing to do
  bool theta_applied = false;     // theta_local has written S_theta itself (one rank): theta_apply has nothing to do
  int cur = 0;
  gbx::SbmState* st = nullptr;
  gbx::SbmParams* params = nullptr;
  double *part_vertex = nullptr, *part_mstep = nullptr, *part_theta = nullptr, *part_fe = nullptr;
  int grid_vertex = 0, grid_mstep = 0, grid_theta = 0, grid_fe = 0, grid_pull = 0;
  gbx::SbmState* h_state = nullptr;  // pinned mirror
  int cur_it = 0;                    // iteration the launchers stamp their kernels with (0: never skip)
  double cur_thr = -1.0;             // residual threshold sbm_apply_totals tests (< 0: never converged)
  cudaEvent_t ev_conv = nullptr;     // the residual of the last iteration has reached h_state->conv
  int itr = 0;
  // ---- packed boundary exchange of a partitioned run (exchange == 1) -------------------------------------------
  // The sweep stores a message for a row of rank o at the next position of this rank's send run for o (positions
  // follow this rank's slot order, so the stores are nearly sequential); the runs are pushed to o's receive buffer
  // by the copy engine while later pieces of the sweep run; after the all-reduce of the totals (the barrier that
  // says every rank's copies have landed) o scatters the received rows into its message buffer.  Random 64-B peer
  // stores reach 300-420 GB/s on NVLink 5, contiguous copies 700-770 GB/s (profiles/r01_p2p_probe.txt).
  int exchange = 0;                    // 0: direct peer stores from the sweep, 1: pack + bulk copy + unpack
  int* rev_pack = nullptr;             // rev[] with remote slots replaced by (owner, position in the send run)
  double* sendbuf = nullptr;           // [send_total][q]
  double* recvbuf[2] = {nullptr, nullptr};   // [send_total][q] each: one per sweep parity (run of rank s at recv_off[s])
  double* recvp[2][gbx::MAXPEER] = {};       // every rank's receive buffers (peer-mapped)
  bool ipc_open_recv[2][gbx::MAXPEER] = {};
  int* unpack_slot = nullptr;          // [send_total]: local slot of every received row
  int64_t send_total = 0;
  int64_t send_off[gbx::MAXPEER + 1] = {0};  // start of the run for rank o in sendbuf, and of the run from o in recvbuf
  int64_t my_off_at[gbx::MAXPEER] = {0};     // start of this rank's run in rank o's recvbuf
  int nchunks = 1;
  int chunk_tile[gbx::MAXCHUNK + 1] = {0};
  int64_t chunk_send[gbx::MAXCHUNK + 1][gbx::MAXPEER] = {};  // rows of the run for o written by pieces < c
  cudaStream_t cap_stream = nullptr;                  // recording stream of the EM loop's CUDA graph (small graphs)
  cudaStream_t copy_stream = nullptr;                 // unpack + the copies to rank (me + 1) % world
  cudaStream_t peer_stream[gbx::MAXPEER] = {};        // [k]: copies to rank (me + k) % world, k >= 2 (own copy engine queue)
  cudaEvent_t ev_peer[gbx::MAXPEER] = {};
  cudaEvent_t ev_piece[gbx::MAXCHUNK] = {};
  cudaEvent_t ev_copied = nullptr, ev_unpacked = nullptr;
  bool unpack_inflight = false;        // k_unpack is running on the copy stream; join before touching messages
  int sweep_rows = 0, sweep_row0 = 0;  // partial rows written by the last vertex pass (count, first row)
  int xparity = 0;                     // receive buffer the next sweep pushes into
  bool pending_unpack = false;         // the last sweep was packed and its rows have not been scattered yet
  // ---- pull sweeps (gbx_sbm_pull.cuh): a vertex keeps its OUTGOING messages and rebuilds the incoming ones ------
  int pull = 0;                        // this EM run sweeps with the pull kernels (chosen by gbx_*_init)
  int gen = 0;                         // sweeps since init: the outgoing messages of generation g live in msg[g & 1];
                                       // at g = 0 msg[1] holds the incoming messages (the layout of the push path)
  double* rec[2] = {nullptr, nullptr}; // per-vertex records [n_global][qm] of generation g in rec[g & 1]
  float* coldeg = nullptr;             // degree of col[e]
  unsigned char* slot_lv = nullptr;    // vertex (inside its tile) of every slot
  int *in_link = nullptr, *out_link = nullptr;   // per local slot: index in `links` of the incoming / outgoing message
  double* recp[2][gbx::MAXPEER] = {};  // rec[b] of every rank (peer-mapped): a sweep stores a vertex's record everywhere
  bool ipc_open_rec[2][gbx::MAXPEER] = {};
  double* theta_pp[2] = {nullptr, nullptr};  // [n] theta of this rank's vertices, two generations
  int thcur = 0, th_used = 0;          // theta_pp[thcur]: current; theta_pp[th_used]: what the last sweep ran with
  alignas(64) unsigned char tmap[2][128] = {};  // CUtensorMap of msg[0] / msg[1] (rows of qm doubles, box of 32 rows)
  bool tmap_ok = false;
};

namespace gbx {

// Launchers of one compiled q.  Each returns GBX_OK or an error code and bumps h->launches.
struct SbmOps {
  int q;
  int (*vertex_pass)(gbx_sbm* h, int mode);      // sweep / marginals / log Z_i over this rank's vertices
  int (*local_totals)(gbx_sbm* h, int mode);     // CTA partial rows -> h->totals (this rank)
  int (*apply_totals)(gbx_sbm* h, int mode);     // (all-reduced) totals -> residual, gr, Crs / omegas, ...
  int (*prepare)(gbx_sbm* h, int zero_h, int gr_from_sum);
  int (*theta_local)(gbx_sbm* h);                // update_theta on this rank's vertices; S_theta partial -> totals
  int (*theta_apply)(gbx_sbm* h);                // totals -> S_theta
  int (*escale)(gbx_sbm* h);                     // theta_row * theta_col per slot (needs every rank's theta)
  int (*fe_local)(gbx_sbm* h, int pass);         // per-edge free-energy terms -> totals
  int (*fe_apply)(gbx_sbm* h, int pass);
  int (*fe_final)(gbx_sbm* h);
  int (*permute)(gbx_sbm* h, const double* src, double* dst, int to_slots);
  int (*random_messages)(gbx_sbm* h, uint64_t seed);
  int (*assignment)(gbx_sbm* h, int* dev_block);
  int (*occupancy)(int variant, int* vertex_occ, int* mstep_occ);
  int (*init_model)(gbx_sbm* h);  // per-vertex prior-table rows of the initial state
  int (*unpack)(gbx_sbm* h);      // packed exchange: received rows -> message buffer
  int (*release)(gbx_sbm* h);     // the handle is going away: give up the shared __constant__ parameter block
  int (*mstep_exact)(gbx_sbm* h); // M-step statistics from the current messages of both directions (after a sweep)
  int pull_ok;                    // this q has pull kernels (message rows of 32, 64 or 128 bytes)
  int (*pull_setup)(gbx_sbm* h);  // tensor maps, neighbour degrees
  int (*pull_load)(gbx_sbm* h, const double* src_dev, uint64_t seed);   // initial messages into both layouts
  int (*pull_export)(gbx_sbm* h, double* dst_dev);                      // current messages in `links` order
};

// The pull layout (records, neighbour degrees, link <-> slot maps) is built only for handles that may sweep with the pull
// kernels: partitioned runs (their default) and single-GPU handles created under GBX_PULL=1.
inline bool want_pull_layout(const SbmOps* ops, int world);

}  // namespace gbx

#include <cstdlib>
inline bool gbx::want_pull_layout(const gbx::SbmOps* ops, int world) {
  if (!ops || !ops->pull_ok) return false;
  if (world > 1) return true;
  const char* e = getenv("GBX_PULL");
  return e && atoi(e) != 0;
}
