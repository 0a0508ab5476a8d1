/*
 * libgraphbix_cuda.so -- C ABI of the B200 (sm_100a) implementation of graphBIX's inner loop.
 *
 * What it replaces in the reference (tatsuro-kawamoto/graphBIX):
 *   gbx_sbm_*  <->  function EM(Ntot,B,links,maxCrs,initial,degreecorrection,itrmax,priorlearning,
 *                               learningrate)                              src/sbm.jl:28-370
 *   gbx_mod_*  <->  function EM(gr,Ntot,B,links,maxomega,itrmax,betafrac,learning,priorlearning,dc,
 *                               learningrate)                              src/mod.jl:3-244
 * Everything outside EM() (DocOpt CLI, edge-list parsing, simplegraph, giant component, spectral
 * initialisation, q sweep, summary/assessment/assignment/.smap writers, PyPlot) stays in the host
 * program; INTEGRATION.md shows the `ccall` lines a maintainer adds to sbm.jl / mod.jl.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays are HOST memory unless a name ends in `_dev`;
 *   - `links` is the reference's K x 2 Int64 matrix exactly as Julia stores it (column-major:
 *     links[0..K) = first column, links[K..2K) = second column), 1-based vertex ids, holding BOTH
 *     directions of every undirected edge, no self loops, no duplicates (output of `simplegraph`,
 *     sbm.jl:414-425);
 *   - a "message" array is double[K][q]: row k is PSIcav[sub2ind((N,N), links[k,1], links[k,2])],
 *     the cavity message links[k,1] -> links[k,2] (sbm.jl:106-109);
 *   - PSI is double[N][q] (row i = marginal of vertex i+1), Crs is double[q][q] (symmetric),
 *     gr / h are double[q], theta is double[N];
 *   - every function returns GBX_OK (0) or a negative error code; gbx_last_error() gives the text.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GRAPHBIX_CUDA_H
#define GRAPHBIX_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GBX_OK 0
#define GBX_ERR_INVALID (-1)   /* bad argument / malformed links            */
#define GBX_ERR_CUDA (-2)      /* CUDA runtime error (no device, OOM, ...)   */
#define GBX_ERR_STATE (-3)     /* call order violated (e.g. run before init) */

#define GBX_MAX_Q 16

typedef struct gbx_sbm gbx_sbm; /* opaque handle: one graph + one EM state on one GPU */

/* Options of EM(); defaults (gbx_sbm_default_options) are the reference's. */
typedef struct gbx_sbm_options {
  int degreecorrection;     /* sbm.jl:345  --dc            (default 1)                            */
  int priorlearning;        /* sbm.jl:113  --prior         (default 1)                            */
  double learningrate;      /* sbm.jl:137  --learning_rate (default 0.3)                          */
  double maxCrs;            /* sbm.jl:910  maxCrs = Ntot   (<=0 means Ntot)                       */
  int itrmax;               /* sbm.jl:342  --itrmax        (default 256)                          */
  double bp_conv_threshold; /* sbm.jl:339  BPconvthreshold (default 1e-6)                         */
  double damping;           /* weight of the new message in the synchronous update, (0,1];
                               1 = undamped (default). Not in the reference (asynchronous sweep). */
  int check_every;          /* read the residual back every n iterations (default 1)              */
  int mstep_exact;          /* update_Crs pairing (sbm.jl:128-136).  0 (default): the statistics are accumulated
                               inside the sweep, pairing each NEW outgoing message with the incoming message it
                               was computed from -- identical at the fixed point, 24 % faster per iteration.
                               1: a separate pass after the sweep pairs the new messages of both directions, as
                               update_Crs does when it runs after BP(); use it when a run may stop at itrmax
                               (sbm.jl:354-359) and its parameters are to be compared with the reference's.
                               Single-GPU handles only.                                                         */
} gbx_sbm_options;

/* What EM() returns besides the arrays (sbm.jl:369). */
typedef struct gbx_sbm_result {
  double FE, CVBayes, CVGP, CVGT, CVMAP;
  double varCVBayes, varCVGP, varCVGT, varCVMAP;
  int fail;        /* sbm.jl:310-312: NaN/Inf in the initial Crs                       */
  int cnv;         /* sbm.jl:348-353                                                   */
  int itrnum;      /* iteration at which BPconv < threshold (0 if never)               */
  int itrs_done;   /* iterations actually run                                          */
  int nonfinite;   /* a NaN/Inf/zero appeared in the messages ("overflow", sbm.jl:355) */
  double BPconv;   /* last residual, sum_i sum_s |dPSI_i(s)| / N (sbm.jl:123)          */
  double max_dpsi; /* last max_i sum_s |dPSI_i(s)|                                     */
} gbx_sbm_result;

int gbx_version(void);
const char* gbx_last_error(void);
int gbx_device_count(void); /* number of visible CUDA devices, or GBX_ERR_CUDA */

/* Device buffers of destroyed handles are kept for the next handle (the hosts call EM() once per q, initial condition
 * and sample on the same graph, sbm.jl:912-943); gbx_trim_cache() gives them back to the driver.  Environment:
 * GBX_POOL=0 disables the cache, GBX_POOL_MB=<n> caps it (default: a quarter of the device memory). */
int gbx_trim_cache(void);
int64_t gbx_cached_bytes(int device);

void gbx_sbm_default_options(gbx_sbm_options* opt);

/* Build the device-resident graph: destination-major CSR of the K = 2E directed edges, the
 * reverse-edge index and the tile schedule (sbm.jl:265-275 builds `nb`, `inds` for the same use). */
int gbx_sbm_create(gbx_sbm** out, int device, int64_t n_vertices, int64_t n_links,
                   const int64_t* links, int q);
int gbx_sbm_destroy(gbx_sbm* h);

/* Run every later launch of this handle on `cuda_stream` (a cudaStream_t; NULL = default stream). */
int gbx_sbm_set_stream(gbx_sbm* h, void* cuda_stream);
int gbx_sbm_set_options(gbx_sbm* h, const gbx_sbm_options* opt);

/* Initial state, sbm.jl:277-341: theta = 1, h = 0, Crs clamped below at 1e-6, messages as given
 * (psicav == NULL: uniform random rows normalised to 1 from `seed`, sbm.jl:319-323), then
 * PSI = updatePSI() and gr = mean(PSI).  gr_init is ignored when priorlearning == 0 (sbm.jl:304). */
int gbx_sbm_init(gbx_sbm* h, const double* gr_init, const double* Crs_init, const double* psicav,
                 uint64_t seed);

/* The EM loop, sbm.jl:342-362, followed by sbm.jl:364-366 (h, gr, free energy and CV errors). */
int gbx_sbm_em(gbx_sbm* h, gbx_sbm_result* res);

/* Pieces of the loop, for tests and benchmarks. */
int gbx_sbm_iterate(gbx_sbm* h, int n_iters, double* bpconv_out); /* n x (BP sweep, M-step, theta)  */
int gbx_sbm_bp_sweeps(gbx_sbm* h, int n_sweeps);                  /* n x BP sweep kernel only        */
int gbx_sbm_finalize(gbx_sbm* h, gbx_sbm_result* res);            /* sbm.jl:364-366                  */

/* Read the state back (any pointer may be NULL). */
int gbx_sbm_get_state(gbx_sbm* h, double* gr, double* Crs, double* PSI, double* theta, double* hfield);
int gbx_sbm_get_messages(gbx_sbm* h, double* psicav);
/* argmax_s PSI[i][s] + 1 (sbm.jl:983-986), computed on the device. */
int gbx_sbm_get_assignment(gbx_sbm* h, int32_t* block);

/* Introspection for benchmarks: kernels launched so far, directed edges, bytes of device memory. */
int64_t gbx_sbm_launch_count(const gbx_sbm* h);
int64_t gbx_sbm_num_links(const gbx_sbm* h);
int64_t gbx_sbm_device_bytes(const gbx_sbm* h);
/* Algorithmic HBM bytes one BP sweep moves (DESIGN.md "bytes per unit"). */
double gbx_sbm_sweep_bytes(const gbx_sbm* h);
/* Which sweep kernels the current EM run uses (decided by gbx_*_init): 0 = push (new messages scattered into the
 * receiver's rows; the default on one GPU), 1 = pull (a vertex keeps its outgoing messages and rebuilds the incoming
 * ones from per-vertex records; needs message rows of 32, 64 or 128 bytes and damping == 1; the default of partitioned
 * handles, where it sends one record per vertex over NVLink instead of one message per cut edge).  GBX_PULL=1 / 0 in
 * the environment forces pull / push where possible.  Results agree to rounding; DESIGN.md "Sweep kernels". */
int gbx_sbm_sweep_kind(const gbx_sbm* h);

/* ------------------------------------------------------------------------------------------------
 * mod.jl: SBM restricted to community structure (omega_in / omega_out).
 *   gbx_mod_*  <->  function EM(gr,Ntot,B,links,maxomega,itrmax,betafrac,learning,priorlearning,dc,learningrate)
 *                                                                                src/mod.jl:3-244
 * Same graph conventions as above (`links`, message rows, PSI layout).
 * ------------------------------------------------------------------------------------------------ */
typedef struct gbx_sbm gbx_mod; /* opaque handle (the same device-resident graph object) */

typedef struct gbx_mod_options {
  int dc;                   /* mod.jl:189  --dc            (default 1)                          */
  int learning;             /* mod.jl:224  --learning      (default 1)                          */
  int priorlearning;        /* mod.jl:221  --prior         (default 0)                          */
  double learningrate;      /* mod.jl:226  --learningrate  (default 0.3).  No effect, as in the reference:
                               update_omega() rebinds EM()'s omegain/omegaout (mod.jl:90-94) before the
                               blend of mod.jl:226-227 reads them, so new is mixed with new.        */
  double maxomega;          /* mod.jl:810  maxomega = 1                                         */
  int itrmax;               /* mod.jl:219  --itrmax        (default 256)                        */
  double betafrac;          /* mod.jl:194  position between beta* and beta0 (driver: 0, +0.1 per retry) */
  double bp_conv_threshold; /* mod.jl:218  BPconvthreshold (default 1e-6)                       */
  double damping;           /* synchronous-update damping, (0,1]; 1 = undamped (default)        */
  int check_every;          /* read the residual back every n iterations (default 1)            */
} gbx_mod_options;

/* What EM() returns besides PSI (mod.jl:243). */
typedef struct gbx_mod_result {
  double excm, alpha, beta, omegain, omegaout;
  double FE, CVBayes, CVGP, CVGT, CVMAP;
  double varCVBayes, varCVGP, varCVGT, varCVMAP;
  int cnv, itrnum, itrs_done, nonfinite;
  double BPconv, max_dpsi;
} gbx_mod_result;

void gbx_mod_default_options(gbx_mod_options* opt);
int gbx_mod_create(gbx_mod** out, int device, int64_t n_vertices, int64_t n_links, const int64_t* links, int q);
int gbx_mod_destroy(gbx_mod* h);
int gbx_mod_set_stream(gbx_mod* h, void* cuda_stream);
int gbx_mod_set_options(gbx_mod* h, const gbx_mod_options* opt);
/* Initial state, mod.jl:171-207: excm, beta from betafrac, alpha = 1/K, omegas, messages as given (NULL: random
 * from `seed`), h = 0, PSI = updatePSI().  `gr` is the prior the driver passes in (mod.jl:807: ones(1,B)/B). */
int gbx_mod_init(gbx_mod* h, const double* gr, const double* psicav, uint64_t seed);
int gbx_mod_em(gbx_mod* h, gbx_mod_result* res);                  /* mod.jl:219-241 */
int gbx_mod_iterate(gbx_mod* h, int n_iters, double* bpconv_out); /* n x (BP sweep, gr, omega update) */
int gbx_mod_finalize(gbx_mod* h, gbx_mod_result* res);            /* mod.jl:240-241 */
/* gr[q], PSI[N][q], hfield[q] (= degrees' * PSI), params[4] = {alpha, beta, omegain, omegaout}; any may be NULL */
int gbx_mod_get_state(gbx_mod* h, double* gr, double* PSI, double* hfield, double* params);
int gbx_mod_get_messages(gbx_mod* h, double* psicav);
int gbx_mod_get_assignment(gbx_mod* h, int32_t* block);
int64_t gbx_mod_launch_count(const gbx_mod* h);

/* ------------------------------------------------------------------------------------------------
 * Labeled SBM (edge labels alpha = 1..n_labels, one affinity matrix and one degree sequence per label):
 *   gbx_lsbm_*  <->  function EM(Ntot,Larray,B,G,links,itrmax,BPconvthreshold,learning,priorlearning,dc,learningrate,
 *                                REDfrac,assortativity,Omega)   src/labeled_sbm/labeled-sbmBIX_0.2.3.8_Julia0.6.2.ipynb
 * `links3` is the notebook's K x 3 Int64 matrix (i, j, label) in column-major order, both directions of every edge with
 * the same label (`labeledsimplegraph`).  The initial affinity matrices (notebook EM, "initial state": from Larray,
 * assortativity and Omega) and the random initial messages are computed by the host and handed over; affinity is
 * double[n_labels][q][q].
 * A handle holds one graph (gbx_lsbm_create) or many (gbx_lsbm_create_batch: BASELINE configs[4], a detectability sweep
 * over thousands of N = 10k graphs).  The graphs of a batch share q, the number of labels and the options; every kernel
 * launch serves all of them (grid y index = graph), each graph keeps its own parameters, residual and convergence
 * stamp, and what gbx_lsbm_em returns for graph g is what a handle of its own would return (same kernels; the launch
 * shapes differ, so to rounding -- a given handle is bit-reproducible from run to run).
 * ------------------------------------------------------------------------------------------------ */
#define GBX_MAX_LABELS 4
typedef struct gbx_lsbm gbx_lsbm;

typedef struct gbx_lsbm_options {
  int dc;                   /* degree correction: degrees[i,alpha] = alpha-labelled links at i, else all ones (default 1) */
  int learning;             /* update the affinity matrices (default 1)                                                   */
  int priorlearning;        /* update gr (default 1)                                                                      */
  double learningrate;      /* default 0.3; as in the notebook the blends mix the new gr / affinity with themselves (no effect) */
  int itrmax;               /* default 512                                                                                */
  double bp_conv_threshold; /* default 1e-6                                                                               */
  double REDfrac;           /* != 0: the last label holds the randomly discarded edges, its matrix stays flat at 1/K      */
  double damping;           /* synchronous-update damping, (0,1]; 1 = undamped (default)                                  */
} gbx_lsbm_options;

typedef struct gbx_lsbm_result {
  double FE, CVBayes, CVGP, CVGT, CVMAP;
  double varCVBayes, varCVGP, varCVGT, varCVMAP;
  int cnv, itrnum, itrs_done, nonfinite;
  double BPconv;
} gbx_lsbm_result;

void gbx_lsbm_default_options(gbx_lsbm_options* opt);
int gbx_lsbm_create(gbx_lsbm** out, int device, int64_t n_vertices, int64_t n_links, const int64_t* links3, int q,
                    int n_labels);
/* `count` graphs in one handle: n_vertices[g], n_links[g], links3[g] (K_g x 3, column-major, vertices 1..n_vertices[g]) */
int gbx_lsbm_create_batch(gbx_lsbm** out, int device, int count, const int64_t* n_vertices, const int64_t* n_links,
                          const int64_t* const* links3, int q, int n_labels);
int gbx_lsbm_count(const gbx_lsbm* h);                               /* graphs in the handle                  */
int gbx_lsbm_destroy(gbx_lsbm* h);
int gbx_lsbm_set_stream(gbx_lsbm* h, void* cuda_stream);
int gbx_lsbm_set_options(gbx_lsbm* h, const gbx_lsbm_options* opt);
/* gr[q], affinity[n_labels][q][q], psicav[K][q] (rows in `links3` order); then PSI = updatePSI(), h = update_h(). */
int gbx_lsbm_init(gbx_lsbm* h, const double* gr, const double* affinity, const double* psicav);
/* the same for every graph of a batch: gr[count][q], affinity[count][n_labels][q][q], psicav[g] -> K_g x q */
int gbx_lsbm_init_batch(gbx_lsbm* h, const double* gr, const double* affinity, const double* const* psicav);
/* the same with the initial messages (rand(1,B) per link, normalised) drawn on the device: row k of graph g from
 * (seeds[g], k), so a graph starts from the same messages alone and inside a batch; seeds[gbx_lsbm_count(h)] */
int gbx_lsbm_init_random(gbx_lsbm* h, const double* gr, const double* affinity, const uint64_t* seeds);
/* the main loop + freeenergy() of every graph of the handle; res[gbx_lsbm_count(h)] */
int gbx_lsbm_em(gbx_lsbm* h, gbx_lsbm_result* res);
/* gbx_lsbm_em on `count` initialised handles one after the other (res: one entry per graph, handle by handle); many
 * graphs run faster in ONE handle (gbx_lsbm_create_batch).  `max_concurrent` is ignored.                          */
int gbx_lsbm_em_batch(gbx_lsbm** hs, int count, gbx_lsbm_result* res, int max_concurrent);
int gbx_lsbm_iterate(gbx_lsbm* h, int n_iters, double* bpconv_out);  /* n x (update_h, BP sweep, gr, affinity); residual of graph 0 */
int gbx_lsbm_finalize(gbx_lsbm* h, gbx_lsbm_result* res);            /* freeenergy() of every graph           */
/* gr[q], affinity[n_labels][q][q], PSI[N][q], hfield[n_labels][q]; any may be NULL.  _of: graph g of a batch */
int gbx_lsbm_get_state(gbx_lsbm* h, double* gr, double* affinity, double* PSI, double* hfield);
int gbx_lsbm_get_state_of(gbx_lsbm* h, int g, double* gr, double* affinity, double* PSI, double* hfield);
int gbx_lsbm_get_messages(gbx_lsbm* h, double* psicav);
int gbx_lsbm_get_messages_of(gbx_lsbm* h, int g, double* psicav);
int64_t gbx_lsbm_launch_count(const gbx_lsbm* h);

/* ------------------------------------------------------------------------------------------------
 * Vertex-partitioned runs (BASELINE.json configs[3]): one process per GPU of one NVLink domain, <= 8 ranks.
 * Rank r owns the vertices [r*c, min(n,(r+1)*c)), c = ceil(n/world), and the rows of their incoming messages.
 * New messages into rows of another rank are stored straight into that rank's buffer over NVLink (peer pointers
 * exchanged once with gbx_sbm_ipc_*); the only per-iteration collectives are the all-reduce of the totals vector
 * and, with degree correction, the all-gather of theta -- issued by the host (torch.distributed / NCCL) between
 * gbx_sbm_step calls.  graphbix_b200/engine_dist.py is the reference driver.  gbx_mod handles work the same way.
 * Every rank passes the same `links`; it may be a host pointer or a pointer into this device's memory (an edge list
 * generated on, or broadcast to, the GPU is used in place).  At most 2^29 directed links per rank.
 * ------------------------------------------------------------------------------------------------ */
enum {
  GBX_STEP_PREPARE = 0,          /* arg bit 0: h = 0 (initial marginals); bit 1: gr = update_gr(PSI) first   */
  GBX_STEP_VERTEX_PASS = 1,      /* arg 0: BP sweep (+ M-step statistics), 1: marginals only, 2: sum log Z_i  */
  GBX_STEP_LOCAL_TOTALS = 2,     /* arg = mode of the pass; fills this rank's totals vector                   */
  GBX_STEP_APPLY_TOTALS = 3,     /* after the all-reduce (sum) of the totals vector                           */
  GBX_STEP_THETA_LOCAL = 4,      /* update_theta on this rank's vertices; S_theta partial -> totals[0..q)      */
  GBX_STEP_THETA_APPLY = 5,      /* after the all-reduce of totals[0..q)                                      */
  GBX_STEP_ESCALE = 6,           /* no-op since round 2 (theta_i theta_j is formed inside the sweeps); kept for callers */
  GBX_STEP_FE_LOCAL = 7,         /* arg = pass (0 sums, 1 squared deviations) -> totals[0..5)                 */
  GBX_STEP_FE_APPLY = 8,
  GBX_STEP_FE_FINAL = 9,
  GBX_STEP_COUNT_ITERATION = 10,
  GBX_STEP_UNPACK = 11           /* packed exchange: after the all-reduce that follows a sweep, scatter the rows  */
                                 /* received from the other ranks into the message buffer (no-op otherwise)       */
};
int gbx_sbm_create_part(gbx_sbm** out, int device, int64_t n_vertices, int64_t n_links, const int64_t* links, int q,
                        int rank, int world);
int gbx_mod_create_part(gbx_mod** out, int device, int64_t n_vertices, int64_t n_links, const int64_t* links, int q,
                        int rank, int world);
/* info[0..8) = first owned vertex, owned vertices, first owned slot, owned slots, length of the totals vector,
 * vertices per rank c, rank, world */
int gbx_sbm_part_info(const gbx_sbm* h, int64_t* info);
/* cudaIpcMemHandle_t (64 bytes) of message buffer `which` (0/1), receive buffer `which - 2` (2/3) or per-vertex record
 * buffer `which - 4` (4/5; pull sweeps: a rank stores the records of its vertices into every rank's copy, so only these
 * two have to be mapped for a pull run); every rank imports every other rank's. */
int gbx_sbm_ipc_export(gbx_sbm* h, int which, unsigned char* handle64);
int gbx_sbm_ipc_import(gbx_sbm* h, int which, int rank, const unsigned char* handle64);
/* Device buffers owned by the caller that the collectives run on: totals (info[4] doubles) and theta
 * (world * c doubles, indexed by global vertex id; rank r's update_theta writes [r*c, ...)). */
int gbx_sbm_part_set_buffers(gbx_sbm* h, void* totals_dev, void* theta_dev);
/* How boundary messages travel.  1 (default for world > 1): the sweep packs them into per-destination send runs in
 * this rank's slot order, the copy engine pushes the runs over NVLink while later pieces of the sweep run, and
 * GBX_STEP_UNPACK scatters the received rows after the all-reduce.  0: the sweep stores each message straight into
 * the owner's buffer (64-B peer stores; slower on NVLink, see profiles/r01_p2p_probe.txt).  Damped sweeps
 * (damping != 1) always use 0 because the previous outgoing message lives on the owner. */
int gbx_sbm_part_set_exchange(gbx_sbm* h, int mode);
int gbx_sbm_step(gbx_sbm* h, int step, int arg);
int gbx_sbm_read_residual(gbx_sbm* h, double* bpconv);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHBIX_CUDA_H */
