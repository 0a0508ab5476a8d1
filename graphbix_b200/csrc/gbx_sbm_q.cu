// Kernels of the full-affinity SBM path (sbm.jl EM()) for ONE compile-time number of groups Q = GBX_Q.
// The build compiles this file once per q in 1..16 (see build.py) so the q x q contraction is fully
// unrolled with the affinity matrix as uniform-register operands from __constant__ memory.
//
// Kernel map (reference lines each one replaces):
//   sbm_vertex_kernel<MODE>  BP sweep over tiled vertices + M-step statistics   sbm.jl:83-136 (mod.jl:38-88)
//   sbm_hub_kernel<MODE>     the same for vertices of degree > 128 (one CTA per vertex)
//   sbm_local_totals         per-CTA partial rows -> this rank's totals vector (one block per column, fixed order)
//   sbm_apply_totals         residual, gr, block degree means, Crs update + clamp sbm.jl:120-123, 47-58, 137-147
//   sbm_theta_kernel         update_theta, S_theta for update_h                  sbm.jl:47-64, 70-73
//   (theta_row * theta_col per edge slot is formed inside the sweeps: sender block in the message sign bits)
//   sbm_prepare / sbm_prior_table   gr clamp, h, logs -> __constant__; prior factors per (degree, block) sbm.jl:89-96, 70-73
//   sbm_fe_kernel<PASS>      log Z_ij, CV errors per edge                        sbm.jl:162-200
//   sbm_finalize             FE, CV errors, variances                            sbm.jl:201-212
//   k_unpack                 partitioned runs: received boundary rows -> message buffer
//   mod_*                    the scalar bookkeeping and free energy of mod.jl's model (GBX_MODEL == 1)
#ifndef GBX_Q
#error "compile with -DGBX_Q=<number of groups>"
#endif
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>

#include "gbx_sbm_internal.h"

namespace gbx {
namespace {

#ifndef GBX_MODEL
#define GBX_MODEL 0
#endif
constexpr int MODEL = GBX_MODEL;  // 0: full-affinity SBM (sbm.jl), 1: community-restricted SBM (mod.jl)
constexpr int MODEL_SBM = 0, MODEL_MOD = 1;
constexpr int Q = GBX_Q;
constexpr int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }
constexpr int QM = msg_stride(Q);    // doubles per message row (padded to whole 32-B sectors)
constexpr int QP = next_pow2(Q);     // lanes per vertex in the aggregation stage (padding lanes idle when Q != 2^k)
constexpr int LV = QP;
constexpr int TV = tile_vertices(Q);
constexpr int VPR = NT / LV;         // vertices per aggregation round
constexpr int QS = Q | 1;            // row stride (doubles) of the per-edge factors in smem: odd -> no bank conflicts

constexpr double LOG_1EM8 = -18.420680743952367;  // log(10^(-8.0)), sbm.jl:42
constexpr double INV_LN2 = 1.4426950408889634;
constexpr double LN2 = 0.6931471805599453;
constexpr double EXP_M500 = 7.124576406741286e-218;  // e^-500, sbm.jl:33-35

__constant__ SbmParams c_P;

// c_P is ONE object per (q, model, device): the handles that share it take turns.  A handle that launches a kernel
// reading c_P first makes sure the object holds ITS parameters: if another handle (or this handle on another stream)
// used it last, this handle's stream waits for that user's queued kernels and reloads h->params.  With one handle
// per (q, model, device) -- the q sweep -- this is a pointer compare.  Host threads: handles of equal (q, model) on
// one device must be driven from one thread at a time (include/graphbix_cuda.h).
struct ParamOwner {
  const gbx_sbm* h = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
};
ParamOwner g_owner[64];
std::mutex g_owner_mu;

int acquire_params(gbx_sbm* h, bool reload) {
  std::lock_guard<std::mutex> lock(g_owner_mu);
  ParamOwner& o = g_owner[h->device & 63];
  if (o.h == h && o.stream == h->stream) return GBX_OK;
  if (o.h != nullptr && o.stream != h->stream) {  // the previous user's kernels may still be reading c_P
    if (!o.ev) GBX_CUDA_TRY(cudaEventCreateWithFlags(&o.ev, cudaEventDisableTiming));
    GBX_CUDA_TRY(cudaEventRecord(o.ev, o.stream));
    GBX_CUDA_TRY(cudaStreamWaitEvent(h->stream, o.ev, 0));
  }
  if (reload)
    GBX_CUDA_TRY(cudaMemcpyToSymbolAsync(c_P, h->params, sizeof(SbmParams), 0, cudaMemcpyDeviceToDevice, h->stream));
  o.h = h;
  o.stream = h->stream;
  return GBX_OK;
}

int op_release(gbx_sbm* h) {  // gbx_sbm_destroy: the handle's stream has been synchronised
  std::lock_guard<std::mutex> lock(g_owner_mu);
  ParamOwner& o = g_owner[h->device & 63];
  if (o.h == h) {
    o.h = nullptr;
    o.stream = nullptr;
  }
  return GBX_OK;
}

// rev[e] packs (owner rank, slot inside the owner's buffer); one rank: owner 0.
__device__ __forceinline__ int rev_owner(int rv) { return (int)((unsigned)rv >> REV_SHIFT); }
__device__ __forceinline__ int rev_slot(int rv) { return rv & REV_MASK; }

struct VertexArgs {
  const int* rowptr;
  const int* col;
  const int* rev;
  const float* coldeg;             // per edge slot: degree of the slot's neighbour (static; degree-corrected sbm.jl runs)
  const unsigned char* slot_lv;    // per edge slot: its vertex inside its tile (static)
  const int4* tiles;     // [ntiles] (first vertex, #vertices, first edge slot, #edge slots)
  int ntiles;
  const int* hubs;    // vertex ids with degree > TILE_E
  const double* theta;
  long long vbeg;             // global id of this rank's first vertex (theta, col and erow use global ids)
  const int* pidx;            // per vertex: row of the prior table, 1 + degree * Q + MAP block (0: theta = 1)
  const double* prior_tab;    // [1 + (TILE_E + 1) * Q][QT]: prior factors per (degree, block); row 0: theta = 1
  const double* cur;            // this rank's current message buffer (the tile loop streams it)
  const double* curp[MAXPEER];  // ... and every rank's (peer-mapped over NVLink): old outgoing messages for damping
  double* nxtp[MAXPEER];        // every rank's next buffer: a new message is stored straight into its owner's rows
  double* PSI;
  double* partials;   // [(grid + nhubs)][PSTRIDE]
  double* mpart;      // [(grid + nhubs)][Q * Q]: M-step statistics (sbm.jl) per CTA
  int* status;
  const SbmState* st;   // for skip_iteration
  int it;
  double damping;
  int dc;         // sbm.jl: degree correction with live thetas (update_theta has run); mod.jl: degree correction
  int dc_model;   // degree correction switched on at all
};

// ---- small device helpers ---------------------------------------------------------------------
// True when an earlier EM iteration has already converged (SbmState::done_at): this launch belongs to an iteration the
// host queued ahead of reading the residual and must not touch anything.  it == 0: unconditional launch.
// it < 0: the launch sits in a captured CUDA graph that is replayed for many iterations; its iteration number is the
// device counter SbmState::it_now, advanced by the first kernel of every iteration (next_iteration in the prepare kernels).
__device__ __forceinline__ int stamp_of(const SbmState* st, int it) {
  return it < 0 ? *reinterpret_cast<const volatile int*>(&st->it_now) : it;
}
__device__ __forceinline__ bool skip_iteration(const SbmState* st, int it) {
  if (it == 0) return false;
  const int d = *reinterpret_cast<const volatile int*>(&st->done_at);
  return d != 0 && d < stamp_of(st, it);
}
// First kernel of an EM iteration (one block): under graph replay advance the device's iteration counter, then decide.
__device__ __forceinline__ bool next_iteration_skips(SbmState* st, int it) {
  if (it >= 0) return skip_iteration(st, it);
  const int now = *reinterpret_cast<const volatile int*>(&st->it_now) + 1;
  __syncthreads();   // every thread has read the old value
  if (threadIdx.x == 0) st->it_now = now;
  const int d = *reinterpret_cast<const volatile int*>(&st->done_at);
  return d != 0 && d < now;
}

// 2^n with saturation: 0 below the normal range, +inf above.
__device__ __forceinline__ double pow2c(int n) {
  if (n < -1022) return 0.0;
  if (n > 1023) return INFINITY;
  return pow2i(n);
}
// 1/x for positive normal x to ~1 ulp: MUFU.RCP64H seed + two Newton steps (no special cases needed here).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
// Degree-corrected sbm.jl runs: a message carries the MAP block of its SENDER in the sign bits of its first
// ceil(log2 q) components (messages are positive).  The receiver needs theta_sender = degree_sender / <degree of that
// block> for the edge factor (sbm.jl:87); the degree is a static per-slot value and the block means are q parameters, so
// nothing per edge has to be recomputed when update_theta moves the thetas (no theta_row * theta_col pass over the slots).
constexpr int NB_SIGN = (MODEL != 0 || Q <= 1) ? 0 : (Q <= 2 ? 1 : (Q <= 4 ? 2 : (Q <= 8 ? 3 : 4)));
__device__ __forceinline__ int take_block_bits(double (&m)[Q]) {   // strips the bits, returns the block
  int blk = 0;
#pragma unroll
  for (int b = 0; b < NB_SIGN; ++b) blk |= (int)((unsigned)__double2hiint(m[b]) >> 31) << b;
#pragma unroll
  for (int b = 0; b < NB_SIGN; ++b) m[b] = fabs(m[b]);
  return blk;
}
__device__ __forceinline__ void put_block_bits(double (&m)[Q], int blk) {
#pragma unroll
  for (int b = 0; b < NB_SIGN; ++b)
    if ((blk >> b) & 1) m[b] = __hiloint2double(__double2hiint(m[b]) | (int)0x80000000, __double2loint(m[b]));
}

// Reductions over the QP lanes that share a vertex.  They name all 32 lanes, so every call site must be reached by
// the whole warp (the aggregation stage keeps its branches warp-uniform with a vote for that reason).
template <typename T>
__device__ __forceinline__ T group_max(T x) {
#pragma unroll
  for (int o = QP / 2; o > 0; o >>= 1) {
    const T y = __shfl_xor_sync(0xffffffffu, x, o);
    x = (y > x) ? y : x;
  }
  return x;
}
template <typename T>
__device__ __forceinline__ T group_min(T x) {
#pragma unroll
  for (int o = QP / 2; o > 0; o >>= 1) {
    const T y = __shfl_xor_sync(0xffffffffu, x, o);
    x = (y < x) ? y : x;
  }
  return x;
}
__device__ __forceinline__ double group_sum(double x) {
#pragma unroll
  for (int o = QP / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// Sum (or max) of column `c` of a [nrows][stride] array of per-CTA partials by one warp, in a fixed order.
template <bool MAX = false>
__device__ __forceinline__ double warp_column_reduce(const double* __restrict__ partials, int nrows, int stride, int c) {
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int r = lane; r < nrows; r += 32) {
    const double x = partials[(size_t)r * stride + c];
    s = MAX ? fmax(s, x) : s + x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double y = __shfl_xor_sync(0xffffffffu, s, o);
    s = MAX ? fmax(s, y) : s + y;
  }
  return s;
}

// T[t] = theta_i theta_s (psi * Crs)[t]  (sbm.jl:87), rescaled by 2^-ke so that max_t T[t] is in [1,2).
// A zero / NaN / Inf factor is left as it is (ke = 0): like log(0) in the reference it turns the messages that
// depend on it into NaN, which reaches the residual and is reported as `nonfinite` ("overflow", sbm.jl:355).
__device__ __forceinline__ void edge_transform(const double (&psi)[Q], double sc, double (&T)[Q], int& ke) {
  double m = 0.0;
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    if constexpr (MODEL == MODEL_MOD) {
      T[t] = fma(psi[t], c_P.eb1, 1.0);  // 1 + psi (e^beta - 1), mod.jl:42
    } else {
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < Q; ++s) acc = fma(psi[s], c_P.C[s * Q + t], acc);
      T[t] = acc * sc;
    }
    m = fmax(m, T[t]);
  }
  bool ok;
  ke = exponent_of(m, ok);
  if (!ok) ke = 0;
  const double s2 = pow2i(-ke);
#pragma unroll
  for (int t = 0; t < Q; ++t) T[t] *= s2;
}

// Cavity message i -> j (sbm.jl:108-109) from the vertex product u and this edge's factor T, the absolute
// clamp of normalize_logprob (sbm.jl:42) at 2^(ffrac + n) in the scaled domain, damping, scatter to rev slot.
__device__ __forceinline__ void emit_message_local(const double* __restrict__ u, const double (&T)[Q], double ffrac, int n,
                                                   double (&w)[Q]) {
  double S = 0.0;
  bool low = false;
  const double thr = pow2c(n + 1);
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    w[t] = u[t] * fast_rcp(T[t]);
    S += w[t];
    low = low || (w[t] < thr);
  }
  if (low) {  // rare: some component is at or below the clamp floor
    const double fl = scale2(exp2(ffrac), n);
    S = 0.0;
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      w[t] = (w[t] < fl) ? fl : w[t];
      S += w[t];
    }
  }
  const double inv = fast_rcp(S);
#pragma unroll
  for (int t = 0; t < Q; ++t) w[t] *= inv;
}
__device__ __forceinline__ void emit_message(const VertexArgs& a, const double* __restrict__ u, const double (&T)[Q],
                                             double ffrac, int n, int rv, double (&w)[Q], int blk) {
  emit_message_local(u, T, ffrac, n, w);
  const int owner = rev_owner(rv), slot = rev_slot(rv);
  if (a.damping != 1.0) {
    double old[Q];
    load_msg<Q, QM>(a.curp[owner], slot, old);
    take_block_bits(old);
    const double al = a.damping, be = 1.0 - a.damping;
#pragma unroll
    for (int t = 0; t < Q; ++t) w[t] = be * old[t] + al * w[t];
  }
  if constexpr (NB_SIGN > 0) {
    double ws[Q];
#pragma unroll
    for (int t = 0; t < Q; ++t) ws[t] = w[t];
    put_block_bits(ws, blk);   // the sender's MAP block travels with the message
    store_msg<Q, QM>(a.nxtp[owner], slot, ws);
  } else {
    store_msg<Q, QM>(a.nxtp[owner], slot, w);
  }
}

// ---- M-step statistics fused into the sweep ------------------------------------------------------------
// update_Crs (sbm.jl:128-136) needs G[s][t] = sum over directed links (i,j) of a_s b_t / Z with a = psi_{i->j},
// b = psi_{j->i}, Z = sum_st a_s Crs_st b_t.  The thread that emits psi_{i->j} holds both (its new output and the
// message it consumed), so the statistics are accumulated right there -- the new outgoing message paired with the
// previous incoming one; at the fixed point that is the reference's pairing -- and the 1 GB re-read of a separate
// M-step pass disappears.  The accumulation sum_e a(e) (x) b(e) is a (q x E) x (E x q) product: it runs as FP64
// DMMA m8n8k4 over groups of 4 edge slots, operands exchanged through the warp's own rows of shared memory, the
// q x q accumulator living in two registers per lane per 8x8 block for the whole kernel.
constexpr int MB = (Q + 7) / 8;  // 8x8 blocks per dimension
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
struct MstepAcc {
  double c[MB * MB * 2];
};

// Called by a full warp.  arow: the warp's 32 rows (stride QS doubles) holding a/Z of its 32 edge slots;
// bget(grp, k, comp): component `comp` of the message consumed by the warp's slot 4*grp + k; nrows = real rows.
template <bool FULL, typename BGet>
__device__ __forceinline__ void mstep_accumulate_impl(MstepAcc& g, const double* arow, BGet bget, int nrows) {
  const int lane = threadIdx.x & 31;
  const int k = lane & 3, cidx = lane >> 2;
  const double* pa = arow + k * QS + cidx;
#pragma unroll
  for (int grp = 0; grp < 8; ++grp) {
    const bool live = FULL || (4 * grp + k < nrows);
    double av[MB], bv[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const int comp = m * 8 + cidx;
      const bool ok = live && (Q % 8 == 0 || comp < Q);
      av[m] = ok ? pa[grp * 4 * QS + m * 8] : 0.0;
      bv[m] = ok ? bget(grp, k, comp) : 0.0;
    }
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
      for (int n = 0; n < MB; ++n) dmma884(g.c[(m * MB + n) * 2], g.c[(m * MB + n) * 2 + 1], av[m], bv[n]);
  }
}
// The same with both operands fetched through callables (the pull kernel keeps them in swizzled tiles).
template <bool FULL, typename AGet, typename BGet>
__device__ __forceinline__ void mstep_accumulate_ab_impl(MstepAcc& g, AGet aget, BGet bget, int nrows) {
  const int lane = threadIdx.x & 31;
  const int k = lane & 3, cidx = lane >> 2;
#pragma unroll
  for (int grp = 0; grp < 8; ++grp) {
    const bool live = FULL || (4 * grp + k < nrows);
    double av[MB], bv[MB];
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      const int comp = m * 8 + cidx;
      const bool ok = live && (Q % 8 == 0 || comp < Q);
      av[m] = ok ? aget(grp, k, comp) : 0.0;
      bv[m] = ok ? bget(grp, k, comp) : 0.0;
    }
#pragma unroll
    for (int m = 0; m < MB; ++m)
#pragma unroll
      for (int n = 0; n < MB; ++n) dmma884(g.c[(m * MB + n) * 2], g.c[(m * MB + n) * 2 + 1], av[m], bv[n]);
  }
}
template <typename AGet, typename BGet>
__device__ __forceinline__ void mstep_accumulate_ab(MstepAcc& g, AGet aget, BGet bget, int nrows) {
  if (nrows >= 32) mstep_accumulate_ab_impl<true>(g, aget, bget, 32);
  else if (nrows > 0) mstep_accumulate_ab_impl<false>(g, aget, bget, nrows);
}
template <typename BGet>
__device__ __forceinline__ void mstep_accumulate(MstepAcc& g, const double* arow, BGet bget, int nrows) {
  if (nrows >= 32) mstep_accumulate_impl<true>(g, arow, bget, 32);
  else if (nrows > 0) mstep_accumulate_impl<false>(g, arow, bget, nrows);
}

// Barrier over the NT threads that carry vertex / edge work: the whole CTA in the push kernels, the consumer warps
// (named barrier 1) in the pull kernel, whose producer warp has left by then.
template <bool NAMED>
__device__ __forceinline__ void cta_sync() {
  if constexpr (NAMED) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
  else __syncthreads();
}

// Fixed-order reduction of the per-warp q x q accumulators into this CTA's row of the M-step partials.
template <bool NAMED = false>
__device__ __forceinline__ void store_mstep_partials(const MstepAcc& g, double* scratch /* smem, >= 64 * MB * MB */,
                                                     double* mrow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = 0; w < NT / 32; ++w) {
    cta_sync<NAMED>();
    if (warp == w) {
#pragma unroll
      for (int m = 0; m < MB; ++m)
#pragma unroll
        for (int n = 0; n < MB; ++n)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = m * 8 + (lane >> 2), col = n * 8 + (lane & 3) * 2 + j;
            const int idx = row * (8 * MB) + col;
            scratch[idx] = (w == 0 ? 0.0 : scratch[idx]) + g.c[(m * MB + n) * 2 + j];
          }
    }
  }
  cta_sync<NAMED>();
  for (int i = threadIdx.x; i < Q * Q; i += NT) mrow[i] = scratch[(i / Q) * (8 * MB) + (i % Q)];
}

// Per-thread accumulators of the aggregation lanes (lane t of a group owns component t).
struct LaneAcc {
  double psi = 0.0;   // sum of PSI_i(t)
  double dpsi = 0.0;  // sum of degree_i PSI_i(t)  (h of mod.jl:22-24)
  double mod = 0.0;   // sum over this thread's edge slots of the omega_in statistic (mod.jl:86-87)
  double res = 0.0;   // sum of per-vertex L1 changes (lane t = 0)
  double logz = 0.0;  // sum of log Z_i            (lane t = 0)
  double vmax = 0.0;  // max per-vertex L1 change  (lane t = 0)
};

// Bits of this lane's QP-wide group inside a warp ballot.
__device__ __forceinline__ unsigned group_bits(unsigned ballot) {
  if constexpr (QP == 32) return ballot;
  const unsigned lane = threadIdx.x & 31u;
  return (ballot >> (lane & ~(unsigned)(QP - 1))) & ((1u << QP) - 1u);
}

// Prior factor of a vertex for the lane's component t, and the clamp floor of sbm.jl:42 in the scaled domain:
//   x_t = log gr_t - theta h_t, xmax = max_t x_t, E = exp(x_t - xmax), 2^(fi + ffrac) = 1e-8 / e^xmax.
// It depends on the vertex only through theta = degree / <degree of its MAP block>, so it is tabulated once per
// EM iteration per (degree, block) by sbm_prior_table and read here as one coalesced row.
constexpr int QT = Q + 4;
struct Prior {
  double E, ffrac, xmax, fi;  // fi is an integer kept as a double so that loading it does not stall on a conversion
  double ifrac;               // 2^-ffrac (pull sweeps fold the floor into the vertex's record)
};

__device__ __forceinline__ void prior_row(double th, double* row /* QT */) {
  double x[Q], xmax = -INFINITY;
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    x[t] = c_P.lgr[t] - th * c_P.h[t];
    xmax = fmax(xmax, x[t]);
  }
#pragma unroll
  for (int t = 0; t < Q; ++t) row[t] = exp(x[t] - xmax);
  const double f = (LOG_1EM8 - xmax) * INV_LN2;
  const double fi = fmin(fmax(floor(f), -100000.0), 100000.0);
  row[Q] = f - floor(f);
  row[Q + 1] = fi;
  row[Q + 2] = xmax;
  row[Q + 3] = exp2(-row[Q]);
}

__global__ void sbm_prior_table(const SbmState* st, double* tab, int dc, int it) {
  if (skip_iteration(st, it)) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int nrows = dc ? 1 + (TILE_E + 1) * Q : 1;
  if (r >= nrows) return;
  double th = 1.0;
  if (r > 0) {
    const int d = (r - 1) / Q, blk = (r - 1) % Q;
    if constexpr (MODEL == MODEL_MOD) th = (double)d;  // mod.jl:48: the field is alpha beta d_i h
    else th = (double)d * st->invdr[blk];              // the expression of sbm_theta_kernel: bit-identical theta
  }
  double row[QT];
  prior_row(th, row);
  for (int k = 0; k < QT; ++k) tab[(size_t)r * QT + k] = row[k];
}

template <int MODE, typename Args>
__device__ __forceinline__ Prior load_prior(const Args& a, int r, int t) {
  Prior pr;
  const double* row = a.prior_tab + (size_t)r * QT;
  pr.E = (t < Q) ? row[t] : 0.0;
  pr.ffrac = row[Q];
  pr.fi = row[Q + 1];
  pr.xmax = (MODE == MODE_LOGZ) ? row[Q + 2] : 0.0;
  pr.ifrac = row[Q + 3];
  return pr;
}

// Everything a vertex needs once lane t holds prod_t = product of its (scaled) edge factors and K = the exponent
// taken out of it:  true unnormalised value v_t = u_t 2^K e^xmax,  u_t = E_t prod_t;  clamp floor 2^(fn + ffrac).
// Runs on the QP lanes of a group (all lanes of the warp execute it; `valid` gates the side effects).
template <int MODE, typename Args>
__device__ __forceinline__ void vertex_finish(const Args& a, bool valid, int gv, int deg, int t, const Prior& pr,
                                              double old, double prod, int K, double& u_out, int& fn,
                                              LaneAcc& acc, unsigned long long* s_dr, int* s_nr, int* best_out = nullptr) {
  const bool tl = t < Q;
  const double u = tl ? pr.E * prod : 0.0;
  u_out = u;
  fn = (int)pr.fi - K;
  if constexpr (MODE == MODE_LOGZ) {
    // logsumexp(logunnormalizedMessage(i)), sbm.jl:153, with the 500 cut-off of sbm.jl:33-35
    const double umax = group_max(u);
    const double cut = umax * EXP_M500;
    const double s = group_sum(tl ? ((u < cut) ? cut : u) : 0.0);
    if (valid && t == 0) acc.logz += log(s) + (double)K * LN2 + pr.xmax;
  } else {
    double w = u;
    const unsigned low = group_bits(__ballot_sync(0xffffffffu, tl && (u < pow2c(fn + 1))));
    if (low) {  // rare: a component is at or below the clamp floor
      const double fl = scale2(exp2(pr.ffrac), fn);
      w = (u < fl) ? fl : u;
    }
    const double S = group_sum(tl ? w : 0.0);
    const double p = w * fast_rcp(S);
    double res = 0.0;
    if constexpr (MODE == MODE_SWEEP) res = group_sum(tl ? fabs(p - old) : 0.0);
    const double pm = group_max(tl ? p : -1.0);
    const unsigned eq = group_bits(__ballot_sync(0xffffffffu, tl && (p == pm)));
    if (best_out) *best_out = eq ? __ffs(eq) - 1 : 0;   // MAP block: findmax, first maximum (sbm.jl:49)
    if (valid) {
      if (tl) {
        a.PSI[(size_t)gv * Q + t] = p;
        acc.psi += p;
        if constexpr (MODEL == MODEL_MOD) acc.dpsi += (double)deg * p;
      }
      if (t == 0) {
        acc.res += res;
        acc.vmax = fmax(acc.vmax, res);
        if (MODEL == MODEL_SBM && a.dc_model && eq) {  // update_theta statistics: findmax, first maximum (sbm.jl:49)
          const int best = __ffs(eq) - 1;
          atomicAdd(&s_dr[best], (unsigned long long)deg);
          atomicAdd(&s_nr[best], 1);
        }
      }
    }
  }
}

// Deterministic CTA reduction of the lane accumulators into this CTA's partial row.
template <bool NAMED = false>
__device__ __forceinline__ void store_partials(const LaneAcc& acc, double* scratch /* smem, NT doubles */,
                                               double* prow, const unsigned long long* s_dr, const int* s_nr) {
  const int tid = threadIdx.x;
  cta_sync<NAMED>();
  scratch[tid] = acc.psi;
  cta_sync<NAMED>();
  if (tid < Q) {  // sumPSI[t]: lanes holding component t, in thread order
    double s = 0.0;
    for (int k = tid; k < NT; k += LV) s += scratch[k];
    prow[2 + tid] = s;
  }
  cta_sync<NAMED>();
  scratch[tid] = acc.res;
  cta_sync<NAMED>();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < NT; k += LV) s += scratch[k];
    prow[0] = s;
  }
  cta_sync<NAMED>();
  scratch[tid] = acc.logz;
  cta_sync<NAMED>();
  if (tid == 0) {
    double s = 0.0;
    for (int k = 0; k < NT; k += LV) s += scratch[k];
    prow[1] = s;
  }
  cta_sync<NAMED>();
  scratch[tid] = acc.vmax;
  cta_sync<NAMED>();
  if (tid == 0) {
    double m = 0.0;
    for (int k = 0; k < NT; k += LV) m = fmax(m, scratch[k]);
    prow[2 + 3 * Q] = m;
  }
  if (tid < Q) {
    prow[2 + Q + tid] = (double)s_dr[tid];
    prow[2 + 2 * Q + tid] = (double)s_nr[tid];
  }
  if constexpr (MODEL == MODEL_MOD) {
    cta_sync<NAMED>();
    scratch[tid] = acc.dpsi;
    cta_sync<NAMED>();
    if (tid < Q) {
      double s = 0.0;
      for (int k = tid; k < NT; k += LV) s += scratch[k];
      prow[3 + 3 * Q + tid] = s;
    }
    cta_sync<NAMED>();
    scratch[tid] = acc.mod;
    cta_sync<NAMED>();
    if (tid == 0) {
      double s = 0.0;
      for (int k = 0; k < NT; ++k) s += scratch[k];
      prow[3 + 4 * Q] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// BP sweep over the tiled (degree <= TILE_E = 128) vertices.  Persistent CTAs; the messages of tile k+1 are staged into
// shared memory with cp.async (LDGSTS) while tile k is computed, so the HBM latency is covered by bytes in
// flight in shared memory rather than by registers or occupancy (double buffer, 2 CTA barriers per tile).
//   stage 1  one thread per incoming edge slot: psi row from smem, T = theta theta psi*Crs, rescaled   -> smem
//   stage 2  QP lanes per vertex: product of the T's, prior, clamp floor, marginal, residual            -> smem
//   stage 3  one thread per edge slot: cavity = product / own factor, clamp, normalise, damp, scatter to rev[e]
constexpr int ROWB = QM * 8;                                 // bytes of one message row (padded)
constexpr int CHB = (QM % 2 == 0) ? 16 : 8;                  // cp.async unit
constexpr int CH = ROWB / CHB;                               // units per message row
constexpr bool SWZ = (QM == 4 || QM == 8 || QM == 16);       // power-of-two rows: rotate units to dodge bank conflicts
constexpr int PROWB = Q * 8;                                 // bytes of one row of PSI (not padded)
constexpr int PCHB = (Q % 2 == 0) ? 16 : 8;
constexpr int PCH = PROWB / PCHB;
constexpr int SWZ_DIV = (CH >= 8) ? 1 : 8 / CH;

__device__ __forceinline__ int swz_unit(int row, int c) {
  if constexpr (SWZ) return (c + row / SWZ_DIV) % CH;
  return c;
}

// T[t] = sc * (psi * Crs)[t], no rescaling.
__device__ __forceinline__ void edge_factor(const double (&psi)[Q], double sc, double (&T)[Q]) {
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    if constexpr (MODEL == MODEL_MOD) {
      T[t] = fma(psi[t], c_P.eb1, 1.0);  // 1 + psi (e^beta - 1), mod.jl:42
    } else {
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < Q; ++s) acc = fma(psi[s], c_P.C[s * Q + t], acc);
      T[t] = acc * sc;
    }
  }
}

// Per-tile inputs staged by cp.async, double buffered: nothing in the tile loop waits on a global load.
struct StageBuf {
  static constexpr size_t psi_off = 0;                                   // ne messages (rotated units)
  static constexpr size_t th_off = psi_off + (size_t)TILE_E * ROWB;      // nv thetas of the tile's vertices
  static constexpr size_t cd_off = th_off + (size_t)TV * 8;              // ne neighbour degrees (float)
  static constexpr size_t rev_off = cd_off + (size_t)TILE_E * 4;         // ne reverse slots
  static constexpr size_t old_off = rev_off + (size_t)TILE_E * 4;        // nv old marginals
  static constexpr size_t rp_off = old_off + (size_t)TV * ROWB;          // nv + 1 row pointers
  static constexpr size_t pi_off = rp_off + (size_t)(TV + 4) * 4;        // nv prior-table rows
  static constexpr size_t lv_off = pi_off + (size_t)TV * 4;              // ne + 3 bytes: vertex (inside the tile) of a slot
  static constexpr size_t bytes = ((lv_off + (size_t)TILE_E + 4 + 15) / 16) * 16;
};
struct SweepSmem {
  static constexpr size_t t_off = 2 * StageBuf::bytes;
  static constexpr size_t u_off = t_off + (size_t)TILE_E * QS * 8;
  static constexpr size_t ff_off = u_off + (size_t)TV * Q * 8;
  static constexpr size_t scratch_off = ff_off + (size_t)TV * 8;
  static constexpr size_t dr_off = scratch_off + (size_t)(NT > 64 * MB * MB ? NT : 64 * MB * MB) * 8;
  static constexpr size_t nr_off = dr_off + (size_t)Q * 8;
  static constexpr size_t fn_off = nr_off + (size_t)Q * 4;
  static constexpr size_t best_off = fn_off + (size_t)TV * 4;            // MAP block of the tile's vertices
  static constexpr size_t inv_off = ((best_off + (size_t)TV * 4 + 7) / 8) * 8;   // 1 / <degree of a block>
  static constexpr size_t total = ((inv_off + (size_t)Q * 8 + 127) / 128) * 128;
};

__device__ __forceinline__ void cp_async_unit(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);  // folds to base + offset after inlining
  if constexpr (CHB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_punit(void* smem_dst, const void* gsrc) {  // a unit of a PSI row
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if constexpr (PCHB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Stage everything tile `td` = (v0, nv, e0, ne) needs from HBM into one buffer, coalesced.  Every warp copies the
// messages of its own 32 edge slots (lane l, round i: unit 32 i + l of the warp's 32 * CH units), because stage 3 of
// a warp still reads them (M-step operands) when a faster warp is already staging the tile after next into the same
// buffer: with warp-private rows that overwrite is ordered by the warp's own program order.  Units sit at a rotated
// position inside their row so that stage 1 (one thread per row) reads without bank conflicts; for CH | 32 the
// rotation does not depend on the round, so the offset is a per-thread constant.
template <int MODE>
__device__ __forceinline__ void stage_tile(const VertexArgs& a, const int4& td, char* buf) {
  const int tid = threadIdx.x;
  const int v0 = td.x, nv = td.y, e0 = td.z, ne = td.w;
  const char* src = reinterpret_cast<const char*>(a.cur) + (size_t)e0 * ROWB;
  const int nunits = ne * CH;
  const int lane = tid & 31, wrow = tid & ~31;  // first row of this warp
  if constexpr (32 % CH == 0) {
    constexpr int RPR = 32 / CH;  // rows per round
    const int row0 = wrow + lane / CH, c = lane % CH;
    const int u0 = wrow * CH + lane;
    const char* g = src + (size_t)u0 * CHB;
    if constexpr (!SWZ || (RPR / SWZ_DIV) % CH == 0) {
      char* d = buf + StageBuf::psi_off + row0 * ROWB + swz_unit(row0, c) * CHB;
#pragma unroll
      for (int i = 0; i < CH; ++i)
        if (u0 + i * 32 < nunits) cp_async_unit(d + i * 32 * CHB, g + (size_t)i * 32 * CHB);
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int row = row0 + i * RPR;
        if (u0 + i * 32 < nunits)
          cp_async_unit(buf + StageBuf::psi_off + row * ROWB + swz_unit(row, c) * CHB, g + (size_t)i * 32 * CHB);
      }
    }
  } else {
    for (int k = lane; k < 32 * CH; k += 32) {
      const int u = wrow * CH + k;
      if (u >= nunits) break;
      const int row = u / CH, c = u - row * CH;
      cp_async_unit(buf + StageBuf::psi_off + row * ROWB + swz_unit(row, c) * CHB, src + (size_t)u * CHB);
    }
  }
  if (tid < ne) {
    if (MODEL == MODEL_SBM && a.dc) cp_async_4(buf + StageBuf::cd_off + tid * 4, a.coldeg + e0 + tid);
    if (MODE == MODE_SWEEP) cp_async_4(buf + StageBuf::rev_off + tid * 4, a.rev + e0 + tid);
  }
  if (ne > 0 && (MODE == MODE_SWEEP || (MODEL == MODEL_SBM && a.dc))) {   // slot -> vertex bytes: a 4-byte window rounded outwards
    const int la0 = e0 & ~3;
    if (tid * 4 < ne + (e0 - la0)) cp_async_4(buf + StageBuf::lv_off + tid * 4, a.slot_lv + la0 + tid * 4);
  }
  if (tid <= nv) cp_async_4(buf + StageBuf::rp_off + tid * 4, a.rowptr + v0 + tid);
  if (tid < nv) cp_async_4(buf + StageBuf::pi_off + tid * 4, a.pidx + v0 + tid);
  if (MODEL == MODEL_SBM && a.dc && tid < nv) cp_async_8(buf + StageBuf::th_off + tid * 8, a.theta + a.vbeg + v0 + tid);
  if (MODE == MODE_SWEEP) {
    const char* osrc = reinterpret_cast<const char*>(a.PSI) + (size_t)v0 * PROWB;
    for (int g = tid; g < nv * PCH; g += NT) cp_async_punit(buf + StageBuf::old_off + g * PCHB, osrc + (size_t)g * PCHB);
  }
}

__device__ __forceinline__ void read_staged(const char* buf, int row, double (&psi)[Q]) {
  const char* r = buf + StageBuf::psi_off + row * ROWB;
  if constexpr (CHB == 16) {
#pragma unroll
    for (int c = 0; c < (Q + 1) / 2; ++c) {  // the units that hold components (the rest of a padded row is zeros)
      const double2 v = *reinterpret_cast<const double2*>(r + swz_unit(row, c) * 16);
      psi[2 * c] = v.x;
      if (2 * c + 1 < Q) psi[2 * c + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < CH; ++c) psi[c] = *reinterpret_cast<const double*>(r + c * 8);
  }
}

// Exponent bookkeeping of a running product: pull the exponent of m into k (m stays in [1,2)); a zero stays zero.
__device__ __forceinline__ void renorm(double& m, int& k) {
  bool ok;
  const int ex = exponent_of(m, ok);
  if (ok) {
    m *= pow2i(-ex);
    k += ex;
  } else if (m == m && !(m > 1.0)) {
    m = 0.0;  // underflow: the component is negligible from here on (NaN / Inf are left alone)
  }
}

template <int MODE>
__global__ void __launch_bounds__(NT, ((Q <= 8) ? 3 : 2) * (256 / NT)) sbm_vertex_kernel(VertexArgs a) {
  extern __shared__ __align__(128) char smem[];
  double* Tsm = reinterpret_cast<double*>(smem + SweepSmem::t_off);
  double* Usm = reinterpret_cast<double*>(smem + SweepSmem::u_off);
  double* ffS = reinterpret_cast<double*>(smem + SweepSmem::ff_off);
  double* scratch = reinterpret_cast<double*>(smem + SweepSmem::scratch_off);
  unsigned long long* s_dr = reinterpret_cast<unsigned long long*>(smem + SweepSmem::dr_off);
  int* s_nr = reinterpret_cast<int*>(smem + SweepSmem::nr_off);
  int* fnS = reinterpret_cast<int*>(smem + SweepSmem::fn_off);
  int* bestS = reinterpret_cast<int*>(smem + SweepSmem::best_off);
  double* invS = reinterpret_cast<double*>(smem + SweepSmem::inv_off);

  const int tid = threadIdx.x;
  if (MODE == MODE_SWEEP && skip_iteration(a.st, a.it)) return;
  if (tid < Q) {
    s_dr[tid] = 0ull;
    s_nr[tid] = 0;
    invS[tid] = (MODEL == MODEL_SBM) ? c_P.invdr[tid] : 0.0;   // read by the barrier that publishes the first tile
  }
  LaneAcc acc;
  MstepAcc macc;
#pragma unroll
  for (int k = 0; k < MB * MB * 2; ++k) macc.c[k] = 0.0;
  const int t = tid % LV;
  const bool tl = t < Q;
  const int tt = tl ? t : 0;
  const int vlane = tid / LV;                // vertex handled by this lane in an aggregation round
  const int vwarp = (tid & ~31) / LV;        // first vertex of this warp in a round

  const int G = gridDim.x;
  const int last = a.ntiles - 1;
  int4 tdN = a.tiles[min((int)blockIdx.x, last)];   // descriptor of the tile whose copies are issued next
  stage_tile<MODE>(a, tdN, smem);
  cp_async_commit();
  int4 td = tdN;
  tdN = a.tiles[min((int)blockIdx.x + G, last)];
  cp_async_wait_all();
  __syncthreads();

  int buf = 0;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += G) {
    const int v0 = td.x, nv = td.y, e0 = td.z, ne = td.w;
    const char* cb = smem + (size_t)buf * StageBuf::bytes;
    const int* rpS = reinterpret_cast<const int*>(cb + StageBuf::rp_off);
    const int* piS = reinterpret_cast<const int*>(cb + StageBuf::pi_off);
    const double* oldS = reinterpret_cast<const double*>(cb + StageBuf::old_off);
    const unsigned char* lvS = reinterpret_cast<const unsigned char*>(cb + StageBuf::lv_off) + (e0 & 3);
    // ---- copies of tile k+1 into the other buffer, then the descriptor of tile k+2 into the same registers ----
    {
      int4 tn = tdN;
      if (tile + G >= a.ntiles) tn.y = tn.w = -1;  // nothing to stage
      stage_tile<MODE>(a, tn, smem + (size_t)(buf ^ 1) * StageBuf::bytes);
      cp_async_commit();
    }
    const int4 tdK1 = tdN;
    tdN = a.tiles[min(tile + 2 * G, last)];
    // ---- prior factors of this tile's first round of vertices (an L2-resident table row per vertex) ----
    Prior pr = load_prior<MODE>(a, (vlane < nv) ? piS[vlane] : 0, t);

    // ---- stage 1 ----
    double T[Q];
    double sc = 1.0;
    // The slot's vertex is read NOW and kept for stage 3: the bytes were staged by other threads, and a faster warp starts
    // overwriting this buffer (copies of the tile after next) while slower warps are still in stage 3 -- no barrier lies
    // in between.  (Everything else stage 3 reads from the stage buffer was staged by the reading thread or its warp.)
    int lv = 0;
    if ((MODE == MODE_SWEEP || (MODEL == MODEL_SBM && a.dc)) && tid < ne) lv = lvS[tid];
    if (tid < ne) {
      double psi[Q];
      read_staged(cb, tid, psi);
      const int blk = take_block_bits(psi);   // MAP block of the sender (degree-corrected sbm.jl runs)
      if (MODEL == MODEL_SBM && a.dc)         // theta_row theta_col, sbm.jl:87
        sc = reinterpret_cast<const double*>(cb + StageBuf::th_off)[lv] *
             ((double)reinterpret_cast<const float*>(cb + StageBuf::cd_off)[tid] * invS[blk]);
      edge_factor(psi, sc, T);
#pragma unroll
      for (int k = 0; k < Q; ++k) Tsm[tid * QS + k] = T[k];
    }
    __syncthreads();

    // ---- stage 2 ----
    for (int vb = 0; vb < nv; vb += VPR) {
      if (vb + vwarp >= nv) break;  // warp-uniform: no vertex left for this warp
      const int v = vb + vlane;
      const bool valid = v < nv;
      if (vb > 0) pr = load_prior<MODE>(a, valid ? piS[v] : 0, t);  // tiles with more than VPR vertices
      const int b = valid ? rpS[v] - e0 : 0, e = valid ? rpS[v + 1] - e0 : 0;
      const double old = (MODE == MODE_SWEEP && valid && tl) ? oldS[v * Q + t] : 0.0;
      const double* p = &Tsm[b * QS + tt];
      const double* pe = &Tsm[e * QS + tt];
      double m1 = 1.0, m2 = 1.0;
      int K = 0;
      // Warp-uniform choice (the group reductions below use full-warp shuffles): both branches give bit-identical
      // products, the second one only adds exact power-of-two rescaling.
      if (!__any_sync(0xffffffffu, e - b > c_P.safe_deg)) {
        // the factors are bounded so that a product of this many cannot leave the double range: plain product
        for (; p + QS < pe; p += 2 * QS) {
          m1 *= p[0];
          m2 *= p[QS];
        }
        if (p < pe) m1 *= p[0];
        m1 *= m2;
      } else {
        while (p + 7 * QS < pe) {  // 8 slots per trip: two chains of 4, then pull the exponents out
          m1 *= p[0];
          m2 *= p[QS];
          m1 *= p[2 * QS];
          m2 *= p[3 * QS];
          m1 *= p[4 * QS];
          m2 *= p[5 * QS];
          m1 *= p[6 * QS];
          m2 *= p[7 * QS];
          p += 8 * QS;
          renorm(m1, K);
          renorm(m2, K);
        }
        for (; p + QS < pe; p += 2 * QS) {
          m1 *= p[0];
          m2 *= p[QS];
        }
        if (p < pe) m1 *= p[0];
        renorm(m1, K);
        renorm(m2, K);
        m1 *= m2;  // in [1,4)
        const int Kc = group_max((tl && m1 != 0.0) ? K : -(1 << 30));
        const int Kcc = (Kc == -(1 << 30)) ? 0 : Kc;
        const int d = K - Kcc;  // <= 0 for live components
        m1 = (d < -1000 || m1 == 0.0) ? 0.0 : m1 * pow2i(d > 0 ? 0 : d);
        K = Kcc;
      }
      double u;
      int fn, best = 0;
      vertex_finish<MODE>(a, valid, v0 + v, e - b, t, pr, old, m1, K, u, fn, acc, s_dr, s_nr, &best);
      if constexpr (MODE == MODE_SWEEP) {
        if (valid) {
          if (tl) Usm[v * Q + t] = u;
          if (t == 0) {
            ffS[v] = pr.ffrac;
            fnS[v] = fn;
            bestS[v] = best;
          }
        }
      }
    }
    cp_async_wait_all();  // tile k+1 has landed (this thread's part); the barrier publishes everybody's
    __syncthreads();

    // ---- stage 3 ----
    if constexpr (MODE == MODE_SWEEP) {
      double mod_term = 0.0;
      if (tid < ne) {
        const int rv = reinterpret_cast<const int*>(cb + StageBuf::rev_off)[tid];
        double w[Q];
        emit_message(a, &Usm[lv * Q], T, ffS[lv], fnS[lv], rv, w, bestS[lv]);
        if constexpr (MODEL == MODEL_SBM) {
          // a / Z for the M-step: Z = sum_s a_s (Crs b)_s = sum_s w_s T_s / sc
          double z = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) z = fma(w[k], T[k], z);
          const double kz = sc * fast_rcp(z);
#pragma unroll
          for (int k = 0; k < Q; ++k) Tsm[tid * QS + k] = w[k] * kz;   // this thread's own row: free since stage 2
        } else {
          // mod.jl:86-87: s = psi_{i->j} . psi_{j->i};  omega_in s / ((omega_in - omega_out) s + omega_out)
          double psi[Q];
          read_staged(cb, tid, psi);
          double sdot = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) sdot = fma(w[k], psi[k], sdot);
          mod_term = c_P.oin * sdot / ((c_P.oin - c_P.oout) * sdot + c_P.oout);
        }
      }
      if constexpr (MODEL == MODEL_SBM) {
        __syncwarp();
        const int wbase = tid & ~31;
        const char* pb = cb + StageBuf::psi_off;
        const char* pw = pb + wbase * ROWB;  // this warp's 32 staged rows
        if constexpr (SWZ && CH == 4) {
          // rotated layout: unit (c + row/2) % 4 of row wbase + 4 grp + k; wbase/2 and 2 grp only flip bit 1
          const int kk = tid & 3, cc = (tid & 31) >> 2;
          const int p0 = ((cc >> 1) + (kk >> 1)) & 3;
          const int off0 = kk * ROWB + p0 * 16 + (cc & 1) * 8, off1 = kk * ROWB + (p0 ^ 2) * 16 + (cc & 1) * 8;
          mstep_accumulate(macc, &Tsm[wbase * QS],
                           [&](int grp, int, int) {   // fabs: the first components carry the sender's block in their sign
                             return fabs(*reinterpret_cast<const double*>(pw + grp * 4 * ROWB + ((grp & 1) ? off1 : off0)));
                           },
                           ne - wbase);
        } else {
          mstep_accumulate(macc, &Tsm[wbase * QS],
                           [&](int grp, int k, int comp) {
                             const int row = wbase + 4 * grp + k;
                             if constexpr (CHB == 16)
                               return fabs(*reinterpret_cast<const double*>(pb + row * ROWB + swz_unit(row, comp >> 1) * 16 +
                                                                            (comp & 1) * 8));
                             else
                               return fabs(*reinterpret_cast<const double*>(pb + row * ROWB + comp * 8));
                           },
                           ne - wbase);
        }
        __syncwarp();
      } else {
        acc.mod += mod_term;
      }
    }
    td = tdK1;
    buf ^= 1;
  }
  store_partials(acc, scratch, a.partials + (size_t)blockIdx.x * PSTRIDE, s_dr, s_nr);
  if constexpr (MODE == MODE_SWEEP && MODEL == MODEL_SBM)
    store_mstep_partials(macc, scratch, a.mpart + (size_t)blockIdx.x * Q * Q);
}

// ---------------------------------------------------------------------------------------------
// Hubs (degree > TILE_E = 128): one CTA per vertex, two passes over its incoming messages; the per-component
// product is carried as (mantissa in [1,2), exponent) so any degree is safe.
template <int MODE>
__global__ void __launch_bounds__(NT) sbm_hub_kernel(VertexArgs a, int part_base) {
  __shared__ double pmS[(NT / 32) * Q];
  __shared__ int pkS[(NT / 32) * Q];
  __shared__ int ksS[NT / 32];
  __shared__ double Ush[Q];
  __shared__ double scratch[(NT > 64 * MB * MB) ? NT : 64 * MB * MB];
  __shared__ double rowA[(MODEL == MODEL_SBM && MODE == MODE_SWEEP) ? NT * QS : 1];   // a / Z of the warp's 32 slots
  __shared__ double rowB[(MODEL == MODEL_SBM && MODE == MODE_SWEEP) ? NT * QS : 1];   // consumed messages
  __shared__ double ffSh;
  __shared__ int fnSh, bestSh;
  __shared__ unsigned long long s_dr[Q];
  __shared__ int s_nr[Q];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (MODE == MODE_SWEEP && skip_iteration(a.st, a.it)) return;
  const int gv = a.hubs[blockIdx.x];
  const int e0 = a.rowptr[gv], e1 = a.rowptr[gv + 1];
  // theta of the prior: sbm.jl degree correction factor, or the degree itself in mod.jl:48
  const double th = !a.dc ? 1.0 : (MODEL == MODEL_MOD ? (double)(e1 - e0) : a.theta[a.vbeg + gv]);
  if (tid < Q) {
    s_dr[tid] = 0ull;
    s_nr[tid] = 0;
  }

  double macc_mod = 0.0;
  double pm[Q];
  int pk[Q];
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    pm[t] = 1.0;
    pk[t] = 0;
  }
  int ksum = 0;
  for (int e = e0 + tid; e < e1; e += NT) {
    double psi[Q], T[Q];
    int ke;
    load_msg<Q, QM>(a.cur, e, psi);
    take_block_bits(psi);   // hubs gather the neighbour's theta itself (they are rare)
    const double sc = (MODEL == MODEL_SBM && a.dc) ? th * a.theta[a.col[e]] : 1.0;
    edge_transform(psi, sc, T, ke);
    ksum += ke;
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      pm[t] *= T[t];
      renorm(pm[t], pk[t]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ksum += __shfl_xor_sync(0xffffffffu, ksum, o);
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const double m2 = __shfl_xor_sync(0xffffffffu, pm[t], o);
      const int k2 = __shfl_xor_sync(0xffffffffu, pk[t], o);
      pm[t] *= m2;
      pk[t] += k2;
      renorm(pm[t], pk[t]);
    }
  }
  if (lane == 0) {
    ksS[warp] = ksum;
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      pmS[warp * Q + t] = pm[t];
      pkS[warp * Q + t] = pk[t];
    }
  }
  __syncthreads();

  // warp 0 finishes the vertex: lane t < QP holds component t (same code path as the tiled kernel)
  LaneAcc acc;
  if (warp == 0) {
    const int t = lane % QP;
    const bool tl = t < Q;
    int Ks = 0;
    for (int w = 0; w < NT / 32; ++w) Ks += ksS[w];
    double m = 1.0;
    int k = 0;
    for (int w = 0; w < NT / 32; ++w) {
      m *= pmS[w * Q + (tl ? t : 0)];
      k += pkS[w * Q + (tl ? t : 0)];
      renorm(m, k);
    }
    const int kc = group_max((tl && m != 0.0) ? k : -(1 << 30));
    const int kcc = (kc == -(1 << 30)) ? 0 : kc;
    const int d = k - kcc;  // <= 0 for live components
    const double prod = (d < -1022 || m == 0.0) ? 0.0 : m * pow2i(d > 0 ? 0 : d);
    const bool writer = lane < QP;
    const double old = (MODE == MODE_SWEEP && tl) ? a.PSI[(size_t)gv * Q + t] : 0.0;
    double row[QT];
    prior_row(th, row);  // hubs are rare: no table, the prior is evaluated in place
    Prior pr;
    pr.E = 0.0;
#pragma unroll
    for (int k = 0; k < Q; ++k)
      if (k == t) pr.E = row[k];
    pr.ffrac = row[Q];
    pr.fi = row[Q + 1];
    pr.xmax = row[Q + 2];
    pr.ifrac = row[Q + 3];
    double u;
    int fn, best = 0;
    vertex_finish<MODE>(a, writer, gv, e1 - e0, t, pr, old, prod, Ks + kcc, u, fn, acc, s_dr, s_nr, &best);
    const double ffrac = pr.ffrac;
    if (writer) {
      if (tl) Ush[t] = u;
      if (t == 0) {
        ffSh = ffrac;
        fnSh = fn;
        bestSh = best;
      }
    }
  }
  __syncthreads();

  MstepAcc macc;
#pragma unroll
  for (int k = 0; k < MB * MB * 2; ++k) macc.c[k] = 0.0;
  if constexpr (MODE == MODE_SWEEP) {
    for (int base = e0; base < e1; base += NT) {  // uniform trip count: the M-step accumulation is warp-wide
      const int e = base + tid;
      double mod_term = 0.0;
      if (e < e1) {
        double psi[Q], T[Q], w[Q];
        int ke;
        load_msg<Q, QM>(a.cur, e, psi);
        take_block_bits(psi);
        const double sc = (MODEL == MODEL_SBM && a.dc) ? th * a.theta[a.col[e]] : 1.0;
        edge_transform(psi, sc, T, ke);
        emit_message(a, Ush, T, ffSh, fnSh + ke, a.rev[e], w, bestSh);
        if constexpr (MODEL == MODEL_SBM) {
          double z = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) z = fma(w[k], T[k], z);
          const double kz = sc * pow2i(-ke) * fast_rcp(z);  // T carries the extra factor 2^-ke here
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            rowA[tid * QS + k] = w[k] * kz;
            rowB[tid * QS + k] = psi[k];
          }
        } else {
          double sdot = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) sdot = fma(w[k], psi[k], sdot);
          mod_term = c_P.oin * sdot / ((c_P.oin - c_P.oout) * sdot + c_P.oout);
        }
      }
      if constexpr (MODEL == MODEL_SBM) {
        __syncwarp();
        const int wbase = tid & ~31;
        const double* bb = &rowB[wbase * QS];
        mstep_accumulate(macc, &rowA[wbase * QS], [&](int grp, int k, int comp) { return bb[(4 * grp + k) * QS + comp]; },
                         e1 - base - wbase);
        __syncwarp();
      } else {
        macc_mod += mod_term;
      }
    }
  }
  // warp 0 carries the vertex accumulators (LaneAcc of every other thread is zero); its group layout matches LV = QP
  acc.mod = macc_mod;
  store_partials(acc, scratch, a.partials + (size_t)(part_base + blockIdx.x) * PSTRIDE, s_dr, s_nr);
  if constexpr (MODE == MODE_SWEEP && MODEL == MODEL_SBM)
    store_mstep_partials(macc, scratch, a.mpart + (size_t)(part_base + blockIdx.x) * Q * Q);
}

// ---------------------------------------------------------------------------------------------
// Totals of one vertex pass.  Every CTA left a partial row; `sbm_local_totals` reduces them in a fixed order
// into one vector per rank, the ranks of a partitioned run all-reduce that vector (NCCL, host side), and
// `sbm_apply_totals` does the scalar bookkeeping of the sweep from the global sums.
//   [0] residual  [1] sum log Z_i  [2,2+Q) sum PSI  [2+Q,2+2Q) degree sum per MAP block  [2+2Q,2+3Q) block sizes
//   [2+3Q,2+4Q) sum degree*PSI (mod.jl)  [2+4Q] omega_in statistic (mod.jl)  [3+4Q] unused  [4+4Q, +Q*Q) G (M-step)
constexpr int TOT_G = 4 + 4 * Q;
constexpr int TOT_LEN = TOT_G + Q * Q;

// One rank: nothing sits between the totals and their application (a partitioned run all-reduces them), so the block that
// finishes last applies them right away -- one launch less per pass.
struct ApplyArgs {
  int fuse;       // 0: only the totals (partitioned runs; the apply kernel follows the all-reduce)
  int mode, prior, dc, learning;
  double N, K, lr, cap, thr;
};
#if GBX_MODEL == 0
__device__ __forceinline__ void sbm_apply_totals_body(SbmState* st, const double* __restrict__ tot, const double* maxd, int mode,
                                                      int prior, int dc, double N, double lr, double maxCrs, int it, double thr);
#else
__device__ __forceinline__ void mod_apply_totals_body(SbmState* st, const double* __restrict__ tot, const double* maxd, int mode,
                                                      int prior, int dc, int learning, double N, double K, double lr,
                                                      double maxomega, int it, double thr);
#endif

__global__ void __launch_bounds__(256) sbm_local_totals(const double* __restrict__ pv, int nrows,
                                                        const double* __restrict__ pm, int with_g,
                                                        double* __restrict__ totals, double* __restrict__ maxd_out,
                                                        SbmState* st, int it, ApplyArgs aa) {
  // One block per column of the totals vector (the last block: the max column); fixed summation order.
  if (skip_iteration(st, it)) return;
  __shared__ double red[256];
  const int tid = threadIdx.x;
  const int ncol = 4 * Q + 4 + (with_g ? Q * Q : 0);
  const int c = blockIdx.x;
  const bool is_max = (c == ncol);
  const double* src = nullptr;
  int stride = PSTRIDE, col = 0;
  if (is_max) {
    src = pv;
    col = 2 + 3 * Q;
  } else if (c < 4 * Q + 4) {
    // partial-row columns: 0,1 | 2..2+3Q | maxd at 2+3Q | dpsi at 3+3Q.. | mod stat at 3+4Q
    if (c < 2 + 3 * Q) { src = pv; col = c; }
    else if (c < 2 + 4 * Q) { if (MODEL == MODEL_MOD) { src = pv; col = c + 1; } }
    else if (c == 2 + 4 * Q) { if (MODEL == MODEL_MOD) { src = pv; col = 3 + 4 * Q; } }
  } else {
    src = pm;
    stride = Q * Q;
    col = c - (4 * Q + 4);
  }
  double v = 0.0;
  if (src) {
    for (int r = tid; r < nrows; r += 256) {
      const double x = src[(size_t)r * stride + col];
      v = is_max ? fmax(v, x) : v + x;
    }
  }
  red[tid] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] = is_max ? fmax(red[tid], red[tid + o]) : red[tid] + red[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    if (is_max) *maxd_out = red[0];
    else totals[c] = red[0];
  }
  if (!aa.fuse) return;
  // ---- the last block to get here applies the totals ----
  __shared__ int s_last;
  __shared__ double s_tot[TOT_G + Q * Q + 1];
  if (tid == 0) {
    __threadfence();                                                 // this block's total is visible device-wide ...
    s_last = (atomicAdd(&st->tot_count, 1) == (int)gridDim.x - 1);   // ... before it is counted
  }
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) st->tot_count = 0;   // for the next pass (the next kernel starts after this one has ended)
  __threadfence();
  for (int k = tid; k < ncol; k += 256) s_tot[k] = __ldcg(&totals[k]);
  if (tid == 0) s_tot[TOT_G + Q * Q] = __ldcg(maxd_out);
  __syncthreads();
#if GBX_MODEL == 0
  sbm_apply_totals_body(st, s_tot, &s_tot[TOT_G + Q * Q], aa.mode, aa.prior, aa.dc, aa.N, aa.lr, aa.cap, it, aa.thr);
#else
  mod_apply_totals_body(st, s_tot, &s_tot[TOT_G + Q * Q], aa.mode, aa.prior, aa.dc, aa.learning, aa.N, aa.K, aa.lr, aa.cap, it,
                        aa.thr);
#endif
}

#if GBX_MODEL == 0
// sbm.jl:120-123 (residual, gr), 47-58 (block degree means), 137-147 (update_Crs) from the global totals.
// (a __device__ body so that, on one rank, the last block of sbm_local_totals can run it right away: one launch less)
__device__ __forceinline__ void sbm_apply_totals_body(SbmState* st, const double* __restrict__ tot, const double* maxd, int mode,
                                                      int prior, int dc, double N, double lr, double maxCrs, int it, double thr) {
  const int tid = threadIdx.x;
  if (skip_iteration(st, it)) return;
  if (mode == MODE_LOGZ) {
    if (tid == 0) {
      st->sumlogZi = tot[1];
      if (!(tot[1] == tot[1])) st->status |= 1;
    }
    return;
  }
  __shared__ double g[Q];
  if (tid < Q) {
    const double snew = tot[2 + tid];
    double gnew = st->gr[tid];
    if (mode == MODE_MARGINAL) {
      gnew = snew / N;                                          // gr = update_gr(PSI), sbm.jl:341
      st->Stheta[tid] = snew;                                   // theta = 1 (sbm.jl:277)
    } else {
      if (prior) gnew += (snew - st->sumPSI[tid]) / N;          // sbm.jl:120-122 summed over the sweep
      if (!dc) st->Stheta[tid] = snew;
    }
    st->gr[tid] = gnew;
    g[tid] = gnew;
    st->sumPSI[tid] = snew;
    st->drmean[tid] = tot[2 + Q + tid] / tot[2 + 2 * Q + tid];  // sbm.jl:54-58 (NaN for an empty block, unused)
    st->invdr[tid] = 1.0 / st->drmean[tid];
  }
  if (tid == 0 && mode == MODE_SWEEP) {
    st->conv = tot[0] / N;  // sbm.jl:123
    if (it != 0 && st->done_at == 0 && st->conv < thr) st->done_at = stamp_of(st, it);   // sbm.jl:348-353, acted on from it + 1
    st->maxd = *maxd;
    if (!(tot[0] == tot[0]) || !(tot[0] < INFINITY)) st->status |= 1;  // NaN/Inf reached the marginals
  }
  __syncthreads();
  if (mode == MODE_SWEEP && tid < Q * Q) {
    // update_Crs(), sbm.jl:137-147, from the statistics the sweep accumulated (G over all directed links)
    const int s = tid / Q, t = tid - s * Q;
    const double* G = tot + TOT_G;
    const double Cold = st->C[tid];
    st->Cused[tid] = Cold;                                         // what the sweep just ran with (pull sweeps)
    const double Cn = Cold * 0.5 * (G[s * Q + t] + G[t * Q + s]);  // Crsnew is symmetric (both directions of a link)
    double c = lr * Cn / (N * (g[s] * g[t])) + (1.0 - lr) * Cold;  // sbm.jl:137; (g g) first: bitwise symmetric in s, t
    if (c > maxCrs) c = maxCrs; else if (c < 0.0) c = 0.0;          // sbm.jl:138-146
    if (!(c == c)) st->status |= 1;
    st->C[tid] = c;
  }
}
__global__ void sbm_apply_totals(SbmState* st, const double* __restrict__ tot, const double* maxd, int mode, int prior,
                                 int dc, double N, double lr, double maxCrs, int it, double thr) {
  sbm_apply_totals_body(st, tot, maxd, mode, prior, dc, N, lr, maxCrs, it, thr);
}
#endif

// ---------------------------------------------------------------------------------------------
// update_theta(), sbm.jl:47-64, and S_theta = sum_i theta_i PSI_i for the next update_h().
__global__ void __launch_bounds__(NT) sbm_theta_kernel(const SbmState* st, const int* __restrict__ rowptr, int n,
                                                       const double* __restrict__ PSI, double* theta,
                                                       int* __restrict__ pidx, double* partials, int it) {
  __shared__ double red[(NT / 32) * Q];
  if (skip_iteration(st, it)) return;
  double acc[Q];
#pragma unroll
  for (int t = 0; t < Q; ++t) acc[t] = 0.0;
  for (int i = blockIdx.x * NT + threadIdx.x; i < n; i += gridDim.x * NT) {
    double p[Q];
    load_vec<Q>(PSI, i, p);
    int best = 0;
    double bv = -1.0;
#pragma unroll
    for (int t = 0; t < Q; ++t)
      if (p[t] > bv) {
        bv = p[t];
        best = t;
      }
    const double th = (double)(rowptr[i + 1] - rowptr[i]) * st->invdr[best];   // degrees[i] / drmean[block[i]], sbm.jl:61
    theta[i] = th;
    const int deg = rowptr[i + 1] - rowptr[i];
    pidx[i] = (deg <= TILE_E) ? 1 + deg * Q + best : 0;  // row of the prior table (hubs evaluate the prior in place)
#pragma unroll
    for (int t = 0; t < Q; ++t) acc[t] = fma(th, p[t], acc[t]);
  }
  block_reduce_store<Q>(acc, red, partials + (size_t)blockIdx.x * Q);
}

// partial rows -> totals[0..ncol) (local), then apply after the ranks' all-reduce
__global__ void k_local_columns(const double* __restrict__ partials, int nrows, int ncol, double* __restrict__ totals,
                                const SbmState* st = nullptr, int it = 0) {
  if (st && skip_iteration(st, it)) return;
  for (int c = threadIdx.x >> 5; c < ncol; c += blockDim.x >> 5) {
    const double v = warp_column_reduce(partials, nrows, ncol, c);
    if ((threadIdx.x & 31) == 0) totals[c] = v;
  }
}

// One rank: the column sums of the update_theta partials go straight into S_theta (k_local_columns + sbm_apply_theta in
// one launch; a partitioned run all-reduces the totals in between and keeps the two steps).
__global__ void k_theta_columns_apply(const double* __restrict__ partials, int nrows, SbmState* st, int it) {
  if (skip_iteration(st, it)) return;
  for (int c = threadIdx.x >> 5; c < Q; c += blockDim.x >> 5) {
    const double v = warp_column_reduce(partials, nrows, Q, c);
    if ((threadIdx.x & 31) == 0) st->Stheta[c] = v;
  }
}

__global__ void sbm_apply_theta(SbmState* st, const double* __restrict__ totals, int it) {
  if (skip_iteration(st, it)) return;
  if (threadIdx.x < Q) st->Stheta[threadIdx.x] = totals[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// Per-iteration parameters: gr clamp (sbm.jl:89-96), h = update_h() (sbm.jl:70-73), logs; written to a
// global staging struct that the host copies device->constant in stream order.
__global__ void sbm_prepare(SbmState* st, SbmParams* out, double N, int zero_h, int gr_from_sum, int dc_range,
                            double maxdeg, int it) {
  if (next_iteration_skips(st, it)) return;
  __shared__ double g[Q];
  const int tid = threadIdx.x;
  if (tid == 0) {
    double s = 0.0;
    for (int r = 0; r < Q; ++r) {
      double x = gr_from_sum ? st->sumPSI[r] / N : st->gr[r];  // gr = update_gr(PSI), sbm.jl:365
      if (x != x || x < 1e-9) x = 1.0 / N;
      g[r] = x;
      s += x;
    }
    for (int r = 0; r < Q; ++r) {
      g[r] /= s;
      st->gr[r] = g[r];
    }
    out->logN = log(N);
  }
  __syncthreads();
  if (tid < Q) {
    out->lgr[tid] = log(g[tid]);
    double hh = 0.0;
    if (!zero_h)
      for (int s = 0; s < Q; ++s) hh += st->Stheta[s] * st->C[s * Q + tid];
    hh /= N;
    st->h[tid] = hh;
    out->h[tid] = hh;
  }
  for (int k = tid; k < Q * Q; k += blockDim.x) {
    out->C[k] = st->C[k];
    out->logC[k] = log(st->C[k]);
    out->Cprev[k] = st->Cused[k];
  }
  if (tid < Q) out->invdr[tid] = st->invdr[tid];
  if (tid == 0) {
    // Largest degree for which a plain product of edge factors provably stays inside the double range:
    // factor in [thmin^2 Cmin, thmax^2 Cmax] with theta = degree / <degree of the block> in [1/dmax, maxdeg/dmin].
    double cmin = INFINITY, cmax = 0.0;
    for (int k = 0; k < Q * Q; ++k) {
      cmin = fmin(cmin, st->C[k]);
      cmax = fmax(cmax, st->C[k]);
    }
    double tmin = 1.0, tmax = 1.0;
    if (dc_range) {
      double dmin = INFINITY, dmax = 0.0;
      for (int r = 0; r < Q; ++r) {
        const double d = st->drmean[r];
        if (d == d) {
          dmin = fmin(dmin, d);
          dmax = fmax(dmax, d);
        }
      }
      if (dmax > 0.0 && dmin > 0.0) {
        tmin = 1.0 / dmax;
        tmax = maxdeg / dmin;
      }
    }
    const double lo = fabs(log2(tmin * tmin * cmin)), hi = fabs(log2(tmax * tmax * cmax));
    const double worst = fmax(fmax(lo, hi), 1.0);
    int sd = 0;
    if (worst == worst && worst < INFINITY) sd = (int)fmin(900.0 / worst, (double)TILE_E);
    out->safe_deg = sd;
  }
}

// ---------------------------------------------------------------------------------------------
// freeenergy(), sbm.jl:162-200: per undirected edge log Z_ij and the three Gibbs/MAP terms.
// PASS 0 accumulates sums, PASS 1 sums of squared deviations from the means (unbiased var, sbm.jl:208-211).
struct PeerCur {
  const double* p[MAXPEER];
};

// Per undirected edge (a = psi_{i->j}, b = psi_{j->i}, lc = log theta_i + log theta_j - log N): log Z_ij and the three
// Gibbs / MAP terms of freeenergy(), sbm.jl:162-200.
__device__ __forceinline__ void fe_terms_sbm(const double (&a)[Q], const double (&b)[Q], double lc, double* x) {
  double Z0 = 0.0, W = 0.0, L1 = 0.0, G1 = 0.0;
  int sa = 0, sb = 0;
  double ma = -1.0, mb = -1.0;
#pragma unroll
  for (int s = 0; s < Q; ++s) {
    if (a[s] > ma) { ma = a[s]; sa = s; }
    if (b[s] > mb) { mb = b[s]; sb = s; }
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const double w = a[s] * b[t];
      const double wc = w * c_P.C[s * Q + t];
      W += w;
      Z0 += wc;
      L1 += w * c_P.logC[s * Q + t];
      G1 += wc * c_P.logC[s * Q + t];
    }
  }
  x[0] = log(Z0) + lc;                // logZijs, sbm.jl:176
  x[1] = W * lc + L1;                 // CVGPs,   sbm.jl:178-184
  x[2] = (Z0 * lc + G1) / Z0;         // CVGTs,   sbm.jl:186-192
  x[3] = lc + c_P.logC[sa * Q + sb];  // CVMAPs,  sbm.jl:194-198
}

// freeenergy(), mod.jl:127-153, per undirected edge: x0 = logZij, x1 = CVBP, x2 = CVGP, x3 = CVGT, x4 = CVMAP.
__device__ __forceinline__ void fe_terms_mod(const double (&a)[Q], const double (&b)[Q], double di, double dj, double oin,
                                             double oout, double beta, double loin, double loout, double* x) {
  double s = 0.0;
  int sa = 0, sb = 0;
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    s = fma(a[t], b[t], s);
    if (a[t] > a[sa]) sa = t;
    if (b[t] > b[sb]) sb = t;
  }
  x[0] = log(1.0 + c_P.eb1 * s);
  x[1] = x[0] + log(di) + loout + log(dj);
  x[2] = beta * s;
  x[3] = di * dj * (oout * loout + (oin * loin - oout * loout) * s) / exp(x[1]);
  x[4] = (sa == sb) ? loin : loout;
}

template <int PASS>
__global__ void __launch_bounds__(NT) sbm_fe_kernel(const SbmState* st, const int* __restrict__ rev,
                                                    const int* __restrict__ col, const int* __restrict__ erow,
                                                    long long L, const double* __restrict__ psi, PeerCur peers,
                                                    const double* __restrict__ theta, int dc, int rank,
                                                    double* partials) {
  __shared__ double red[(NT / 32) * 4];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  double mean[4] = {0.0, 0.0, 0.0, 0.0};
  if (PASS == 1)
    for (int k = 0; k < 4; ++k) mean[k] = st->cvmean[k];
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < L; i += (long long)gridDim.x * NT) {
    const int e = (int)i, r = rev[i];
    // every undirected edge once: at the larger of its two global slots (the `i < j -> continue` of sbm.jl:166)
    if (rev_owner(r) < rank || (rev_owner(r) == rank && rev_slot(r) < e)) {
      double a[Q], b[Q];
      load_msg<Q, QM>(psi, e, b);                               // j -> i   (i = row of e, j = col[e])
      load_msg<Q, QM>(peers.p[rev_owner(r)], rev_slot(r), a);   // i -> j   (in the row of j, possibly on a peer)
      take_block_bits(a);
      take_block_bits(b);
      double lc = -c_P.logN;
      if (dc) lc += log(theta[erow[e]]) + log(theta[col[e]]);
      double x[4];
      fe_terms_sbm(a, b, lc, x);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (PASS == 0) acc[k] += x[k];
        else acc[k] += (x[k] - mean[k]) * (x[k] - mean[k]);
      }
    }
  }
  block_reduce_store<4>(acc, red, partials + (size_t)blockIdx.x * 4);
}

__global__ void sbm_apply_fe(SbmState* st, const double* __restrict__ totals, int pass, double L) {
  const int c = threadIdx.x;
  if (c < 4) {
    if (pass == 0) {
      st->cvsum[c] = totals[c];
      st->cvmean[c] = totals[c] / L;
    } else {
      st->cvss[c] = totals[c];
    }
  }
}

__global__ void sbm_finalize(SbmState* st, double N, double L) {
  if (threadIdx.x == 0) {
    double ave = 0.0;
    for (int s = 0; s < Q; ++s)
      for (int t = 0; t < Q; ++t) ave += st->gr[s] * st->C[s * Q + t] * st->gr[t];  // sbm.jl:201
    st->result[0] = -((st->sumlogZi - st->cvsum[0]) / N + 0.5 * ave);                 // FE, sbm.jl:203
    for (int k = 0; k < 4; ++k) {
      st->result[1 + k] = 1.0 - st->cvsum[k] / L;                                     // sbm.jl:204-207
      st->result[5 + k] = st->cvss[k] / (L - 1.0);                                    // sbm.jl:208-211
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Messages between `links` order (all K_global links) and this rank's slots [sbeg, sbeg + Kl).
//   to_slots: dst (local slots) <- src (link order);  otherwise dst (link order, zeros for links of other ranks) <- src
__global__ void k_permute_rows(const double* __restrict__ src, double* __restrict__ dst,
                               const int* __restrict__ slot_of_link, long long K, int to_slots, long long sbeg,
                               long long Kl) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < K;
       k += (long long)gridDim.x * blockDim.x) {
    const long long s = (long long)slot_of_link[k] - sbeg;
    const bool local = s >= 0 && s < Kl;
    double m[Q];
    if (to_slots) {
      if (!local) continue;
      load_vec<Q>(src, k, m);
      store_msg<Q, QM>(dst, s, m);
    } else {
      if (local) {
        load_msg<Q, QM>(src, s, m);
        take_block_bits(m);   // messages leave without the sender's block in their signs
      } else {
#pragma unroll
        for (int t = 0; t < Q; ++t) m[t] = 0.0;
      }
      store_vec<Q>(dst, k, m);
    }
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}

// rand(1,B) normalised (sbm.jl:319-323), counter-based so the result does not depend on the launch shape.
__global__ void k_random_messages(double* __restrict__ dst, const int* __restrict__ slot_of_link, long long K,
                                  uint64_t seed, long long sbeg, long long Kl) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < K;
       k += (long long)gridDim.x * blockDim.x) {
    const long long slot = (long long)slot_of_link[k] - sbeg;
    if (slot < 0 || slot >= Kl) continue;  // keyed by the global link index: the same draw whatever the partition
    double m[Q];
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < Q; ++t) {
      const uint64_t r = splitmix64(seed ^ splitmix64((uint64_t)k * Q + t));
      m[t] = ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      s += m[t];
    }
#pragma unroll
    for (int t = 0; t < Q; ++t) m[t] /= s;
    store_msg<Q, QM>(dst, slot, m);
  }
}

__global__ void k_assignment(const double* __restrict__ PSI, int n, int* __restrict__ block) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double p[Q];
    load_vec<Q>(PSI, i, p);
    int best = 0;
    double bv = p[0];
#pragma unroll
    for (int t = 1; t < Q; ++t)
      if (p[t] > bv) {
        bv = p[t];
        best = t;
      }
    block[i] = best + 1;
  }
}

// Row of the prior table per vertex at init: theta = 1 (row 0) for sbm.jl; 1 + degree * Q for mod.jl with
// degree correction (the field alpha beta d_i h depends on the vertex through its degree only, mod.jl:48).
__global__ void k_init_pidx(const int* __restrict__ rowptr, int n, int by_degree, int* __restrict__ pidx) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int deg = rowptr[i + 1] - rowptr[i];
    pidx[i] = (by_degree && deg <= TILE_E) ? 1 + deg * Q : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// update_Crs statistics (sbm.jl:128-136) from the CURRENT messages of both directions, for runs that ask for the
// reference's pairing (gbx_sbm_options::mstep_exact): G += (x_e / Z) (x) x_rev(e) over all directed slots, where
// x_e is the message stored at slot e and Z = x_e' Crs x_rev(e) (the theta factors cancel).  Works on either
// layout: both keep a message and its reverse at slots e and rev[e] of one buffer.
__global__ void __launch_bounds__(NT) sbm_mstep_exact_kernel(const double* __restrict__ msg, const int* __restrict__ rev,
                                                             long long K, double* mpart, const SbmState* st, int it) {
  __shared__ double rowA[NT * QS];
  __shared__ double rowB[NT * QS];
  __shared__ double scratch[(NT > 64 * MB * MB) ? NT : 64 * MB * MB];
  if (skip_iteration(st, it)) return;
  const int tid = threadIdx.x;
  MstepAcc macc;
#pragma unroll
  for (int k = 0; k < MB * MB * 2; ++k) macc.c[k] = 0.0;
  for (long long base = (long long)blockIdx.x * NT; base < K; base += (long long)gridDim.x * NT) {
    const long long e = base + tid;
    if (e < K) {
      double a[Q], b[Q];
      load_msg<Q, QM>(msg, e, a);
      load_msg<Q, QM>(msg, (long long)rev_slot(rev[e]), b);
      take_block_bits(a);   // no-ops on the pull layout, whose messages carry no bits
      take_block_bits(b);
      double z = 0.0;
#pragma unroll
      for (int t = 0; t < Q; ++t) {
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < Q; ++s) acc = fma(a[s], c_P.C[s * Q + t], acc);
        z = fma(acc, b[t], z);
      }
      const double kz = fast_rcp(z);
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        rowA[tid * QS + k] = a[k] * kz;
        rowB[tid * QS + k] = b[k];
      }
    }
    __syncwarp();
    const int wbase = tid & ~31;
    const double* bb = &rowB[wbase * QS];
    const long long left = K - base - wbase;
    mstep_accumulate(macc, &rowA[wbase * QS], [&](int grp, int k, int comp) { return bb[(4 * grp + k) * QS + comp]; },
                     (int)(left > 32 ? 32 : (left < 0 ? 0 : left)));
    __syncwarp();
  }
  store_mstep_partials(macc, scratch, mpart + (size_t)blockIdx.x * Q * Q);
}

#include "gbx_sbm_pull.cuh"

// =============================================================================================
// Launchers (host)
// =============================================================================================
int grid_for(gbx_sbm* h, int64_t items) {
  return (int)std::min<int64_t>(std::max<int64_t>(1, (items + 255) / 256), (int64_t)h->num_sms * 8);
}

VertexArgs vertex_args(gbx_sbm* h) {
  VertexArgs a;
  a.rowptr = h->rowptr;
  a.col = h->col;
  a.rev = h->rev;
  a.coldeg = h->coldeg;
  a.slot_lv = h->slot_lv;
  a.tiles = h->tiles;
  a.ntiles = h->ntiles;
  a.hubs = h->hubs;
  a.theta = h->theta;
  a.vbeg = h->vbeg;
  a.pidx = h->pidx;
  a.prior_tab = h->prior_tab;
  a.cur = h->msg[h->cur];
  for (int r = 0; r < MAXPEER; ++r) {
    a.curp[r] = h->msgp[h->cur][r];
    a.nxtp[r] = h->msgp[h->cur ^ 1][r];
  }
  a.PSI = h->PSI;
  a.partials = h->part_vertex;
  a.mpart = h->part_mstep;
  a.status = &h->st->status;
  a.st = h->st;
  a.it = h->cur_it;
  a.damping = (MODEL == MODEL_MOD) ? h->mopt.damping : h->opt.damping;
  a.dc_model = (MODEL == MODEL_MOD) ? h->mopt.dc : h->opt.degreecorrection;
  a.dc = (MODEL == MODEL_MOD) ? a.dc_model : (a.dc_model && h->theta_valid);   // before update_theta every theta is 1
  return a;
}

template <int MODE>
int launch_vertex_pass_t(gbx_sbm* h) {
  { const int rc = acquire_params(h, true); if (rc != GBX_OK) return rc; }
  VertexArgs a = vertex_args(h);
  if (h->ntiles > 0) {
    sbm_vertex_kernel<MODE><<<h->grid_vertex, NT, SweepSmem::total, h->stream>>>(a);
    h->launches++;
  }
  if (h->nhubs > 0) {
    sbm_hub_kernel<MODE><<<h->nhubs, NT, 0, h->stream>>>(a, h->grid_vertex);
    h->launches++;
  }
  h->sweep_rows = (h->ntiles > 0 ? h->grid_vertex : 0) + h->nhubs;
  h->sweep_row0 = h->ntiles > 0 ? 0 : h->grid_vertex;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

// Partitioned sweep with the packed exchange: the messages for rows of rank o go to this rank's send run for o
// (rev_pack holds the positions), hubs first, then the tiles in `nchunks` pieces; when a piece is done its part of
// every run is pushed to the owner's receive buffer by the copy engine while the next piece computes.
int launch_packed_sweep(gbx_sbm* h) {
  { const int rc = acquire_params(h, true); if (rc != GBX_OK) return rc; }
  VertexArgs a = vertex_args(h);
  a.rev = h->rev_pack;
  for (int o = 0; o < h->world; ++o)
    if (o != h->rank) a.nxtp[o] = h->sendbuf + (size_t)h->send_off[o] * QM;
  const int G = h->grid_vertex, NC = h->nchunks;
  if (h->nhubs > 0) {
    sbm_hub_kernel<MODE_SWEEP><<<h->nhubs, NT, 0, h->stream>>>(a, NC * G);
    h->launches++;
  }
  const size_t rowb = (size_t)QM * sizeof(double);
  double* const* dst = h->recvp[h->xparity];
  for (int c = 0; c < NC; ++c) {
    VertexArgs ac = a;
    ac.tiles = a.tiles + h->chunk_tile[c];
    ac.ntiles = h->chunk_tile[c + 1] - h->chunk_tile[c];
    ac.partials = a.partials + (size_t)c * G * PSTRIDE;
    ac.mpart = a.mpart + (size_t)c * G * Q * Q;
    if (ac.ntiles > 0) {
      sbm_vertex_kernel<MODE_SWEEP><<<G, NT, SweepSmem::total, h->stream>>>(ac);
      h->launches++;
    } else {
      GBX_CUDA_TRY(cudaMemsetAsync(ac.partials, 0, (size_t)G * PSTRIDE * sizeof(double), h->stream));
      GBX_CUDA_TRY(cudaMemsetAsync(ac.mpart, 0, (size_t)G * Q * Q * sizeof(double), h->stream));
    }
    GBX_CUDA_TRY(cudaEventRecord(h->ev_piece[c], h->stream));
    // one queue per destination, destinations in rotated order: at any moment the ranks push to different owners
    for (int k = 1; k < h->world; ++k) {
      const int o = (h->rank + k) % h->world;
      const long long b = h->chunk_send[c][o], e = h->chunk_send[c + 1][o];
      if (e <= b) continue;
      GBX_CUDA_TRY(cudaStreamWaitEvent(h->peer_stream[k], h->ev_piece[c], 0));
      GBX_CUDA_TRY(cudaMemcpyAsync(dst[o] + (size_t)(h->my_off_at[o] + b) * QM,
                                   h->sendbuf + (size_t)(h->send_off[o] + b) * QM, (size_t)(e - b) * rowb,
                                   cudaMemcpyDeviceToDevice, h->peer_stream[k]));
    }
  }
  for (int k = 1; k < h->world; ++k) {  // the totals (and the all-reduce) follow the copies
    GBX_CUDA_TRY(cudaEventRecord(h->ev_peer[k], h->peer_stream[k]));
    GBX_CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_peer[k], 0));
  }
  h->sweep_rows = NC * G + h->nhubs;
  h->sweep_row0 = 0;
  h->pending_unpack = true;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

// Rows received from the other ranks (parity of the last sweep) -> their slots of the (new) current message buffer.
__global__ void k_unpack(const double* __restrict__ recv, const int* __restrict__ slot, long long rows, double* __restrict__ msg) {
  constexpr int U = (QM % 2 == 0) ? QM / 2 : QM;  // 16-B units per (padded) row when it is a whole number of them
  const long long total = rows * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / U;
    const int c = (int)(i - r * U);
    const long long s = slot[r];
    if constexpr (QM % 2 == 0)
      reinterpret_cast<double2*>(msg)[s * U + c] = reinterpret_cast<const double2*>(recv)[i];
    else
      msg[s * U + c] = recv[i];
  }
}

// Runs on the copy stream, next to the small kernels and collectives that follow the sweep on the main stream; every
// entry point that touches the messages joins it first (join_unpack in gbx_sbm_host.cu).
int op_unpack(gbx_sbm* h) {
  if (!h->pending_unpack) return GBX_OK;
  h->pending_unpack = false;
  const int parity = h->xparity;
  h->xparity ^= 1;
  if (h->send_total > 0) {
    // called after the sweep's buffer swap: h->cur is the buffer the sweep wrote
    GBX_CUDA_TRY(cudaEventRecord(h->ev_piece[0], h->stream));  // after the all-reduce: every rank's rows have landed
    GBX_CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_piece[0], 0));
    k_unpack<<<grid_for(h, h->send_total * ((QM % 2 == 0) ? QM / 2 : QM)), 256, 0, h->copy_stream>>>(
        h->recvbuf[parity], h->unpack_slot, (long long)h->send_total, h->msg[h->cur]);
    h->launches++;
    GBX_CUDA_TRY(cudaGetLastError());
    GBX_CUDA_TRY(cudaEventRecord(h->ev_unpacked, h->copy_stream));
    h->unpack_inflight = true;
  }
  return GBX_OK;
}

int pull_vertex_pass(gbx_sbm* h, int mode);
int pull_fe_local(gbx_sbm* h, int pass);

int op_vertex_pass(gbx_sbm* h, int mode) {
  if (h->pull) return pull_vertex_pass(h, mode);
  if (mode == MODE_SWEEP) {
    const double damping = (MODEL == MODEL_MOD) ? h->mopt.damping : h->opt.damping;
    if (h->world > 1 && h->exchange == 1 && damping == 1.0) return launch_packed_sweep(h);
    return launch_vertex_pass_t<MODE_SWEEP>(h);
  }
  if (mode == MODE_MARGINAL) return launch_vertex_pass_t<MODE_MARGINAL>(h);
  return launch_vertex_pass_t<MODE_LOGZ>(h);
}

int op_local_totals(gbx_sbm* h, int mode) {
  const size_t skip = (size_t)h->sweep_row0;  // rows of the last vertex pass (a packed sweep has one set per piece)
  const int with_g = (MODEL == MODEL_SBM && mode == MODE_SWEEP) ? 1 : 0;
  ApplyArgs aa{};
  aa.fuse = (h->world == 1 && !h->stepped) ? 1 : 0;   // one rank: the last block applies the totals (op_apply_totals skips)
  aa.mode = mode;
  aa.thr = h->cur_thr;
  if (MODEL == MODEL_MOD) {
    aa.prior = h->mopt.priorlearning;
    aa.dc = h->mopt.dc;
    aa.learning = h->mopt.learning;
    aa.N = (double)h->n_global;
    aa.K = (double)h->K_global;
    aa.lr = h->mopt.learningrate;
    aa.cap = h->mopt.maxomega;
  } else {
    aa.prior = h->opt.priorlearning;
    aa.dc = h->opt.degreecorrection;
    aa.N = (double)h->n_global;
    aa.lr = h->opt.learningrate;
    aa.cap = h->opt.maxCrs > 0 ? h->opt.maxCrs : (double)h->n_global;
  }
  sbm_local_totals<<<4 * Q + 4 + (with_g ? Q * Q : 0) + 1, 256, 0, h->stream>>>(
      h->part_vertex + skip * PSTRIDE, h->sweep_rows, h->part_mstep + skip * Q * Q, with_g, h->totals, h->maxd_dev, h->st, h->cur_it,
      aa);
  h->totals_applied = aa.fuse != 0;
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

#if GBX_MODEL == 0
int op_apply_totals(gbx_sbm* h, int mode) {
  if (h->totals_applied) {   // the last block of sbm_local_totals has done it
    h->totals_applied = false;
    return GBX_OK;
  }
  const double maxCrs = h->opt.maxCrs > 0 ? h->opt.maxCrs : (double)h->n_global;
  sbm_apply_totals<<<1, 256, 0, h->stream>>>(h->st, h->totals, h->maxd_dev, mode, h->opt.priorlearning,
                                             h->opt.degreecorrection, (double)h->n_global, h->opt.learningrate, maxCrs, h->cur_it,
                                             h->cur_thr);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int op_prepare(gbx_sbm* h, int zero_h, int gr_from_sum) {
  const int dc_tab = h->opt.degreecorrection && h->theta_valid;
  sbm_prepare<<<1, 64, 0, h->stream>>>(h->st, h->params, (double)h->n_global, zero_h, gr_from_sum, dc_tab,
                                       (double)h->max_degree, h->cur_it);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  { const int rc = acquire_params(h, false); if (rc != GBX_OK) return rc; }
  GBX_CUDA_TRY(cudaMemcpyToSymbolAsync(c_P, h->params, sizeof(SbmParams), 0, cudaMemcpyDeviceToDevice, h->stream));
  {
    const int nrows = dc_tab ? 1 + (TILE_E + 1) * Q : 1;
    sbm_prior_table<<<(nrows + 127) / 128, 128, 0, h->stream>>>(h->st, h->prior_tab, dc_tab, h->cur_it);
    h->launches++;
    GBX_CUDA_TRY(cudaGetLastError());
  }
  return GBX_OK;
}

int op_theta_local(gbx_sbm* h) {
  if (h->pull) h->thcur ^= 1;  // keep the thetas the last sweep ran with: the next sweep rebuilds that sweep's messages
  sbm_theta_kernel<<<h->grid_theta, NT, 0, h->stream>>>(h->st, h->rowptr, (int)h->n, h->PSI,
                                                        h->pull ? h->theta_pp[h->thcur] : h->theta + h->vbeg, h->pidx,
                                                        h->part_theta, h->cur_it);
  h->theta_valid = true;
  if (h->world == 1 && !h->stepped) {   // nobody reduces the totals across ranks: sum and apply in one launch
    k_theta_columns_apply<<<1, 256, 0, h->stream>>>(h->part_theta, h->grid_theta, h->st, h->cur_it);
    h->theta_applied = true;
  } else {
    k_local_columns<<<1, 256, 0, h->stream>>>(h->part_theta, h->grid_theta, Q, h->totals, h->st, h->cur_it);
    h->theta_applied = false;
  }
  h->launches += 2;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int op_theta_apply(gbx_sbm* h) {
  if (h->theta_applied) {   // op_theta_local has already written S_theta
    h->theta_applied = false;
    return GBX_OK;
  }
  sbm_apply_theta<<<1, 32, 0, h->stream>>>(h->st, h->totals, h->cur_it);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

// theta_row * theta_col per edge slot used to be refreshed here after every update_theta (one pass over the slots);
// both sweeps now form it per edge from the sender's block (message sign bits / records), its static degree and the
// block degree means, so there is nothing to do.  Kept as a step of the partitioned sequence (GBX_STEP_ESCALE).
int op_escale(gbx_sbm* h) {
  (void)h;
  return GBX_OK;
}

PeerCur peer_cur(const gbx_sbm* h) {
  PeerCur p;
  for (int r = 0; r < MAXPEER; ++r) p.p[r] = h->msgp[h->cur][r];
  return p;
}

int op_fe_local(gbx_sbm* h, int pass) {
  if (h->pull) return pull_fe_local(h, pass);
  { const int rc = acquire_params(h, true); if (rc != GBX_OK) return rc; }
  if (pass == 0)
    sbm_fe_kernel<0><<<h->grid_fe, NT, 0, h->stream>>>(h->st, h->rev, h->col, h->erow, (long long)h->K, h->msg[h->cur],
                                                       peer_cur(h), h->theta, h->opt.degreecorrection, h->rank, h->part_fe);
  else
    sbm_fe_kernel<1><<<h->grid_fe, NT, 0, h->stream>>>(h->st, h->rev, h->col, h->erow, (long long)h->K, h->msg[h->cur],
                                                       peer_cur(h), h->theta, h->opt.degreecorrection, h->rank, h->part_fe);
  k_local_columns<<<1, 128, 0, h->stream>>>(h->part_fe, h->grid_fe, 4, h->totals);
  h->launches += 2;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int op_fe_apply(gbx_sbm* h, int pass) {
  sbm_apply_fe<<<1, 32, 0, h->stream>>>(h->st, h->totals, pass, 0.5 * (double)h->K_global);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int op_fe_final(gbx_sbm* h) {
  sbm_finalize<<<1, 32, 0, h->stream>>>(h->st, (double)h->n_global, 0.5 * (double)h->K_global);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}
#endif

int op_permute(gbx_sbm* h, const double* src, double* dst, int to_slots) {
  k_permute_rows<<<grid_for(h, h->K_global), 256, 0, h->stream>>>(src, dst, h->slot_of_link, (long long)h->K_global, to_slots,
                                                                 (long long)h->sbeg, (long long)h->K);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int op_random_messages(gbx_sbm* h, uint64_t seed) {
  k_random_messages<<<grid_for(h, h->K_global), 256, 0, h->stream>>>(h->msg[0], h->slot_of_link, (long long)h->K_global, seed,
                                                                    (long long)h->sbeg, (long long)h->K);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int op_assignment(gbx_sbm* h, int* dev_block) {
  k_assignment<<<grid_for(h, h->n), 256, 0, h->stream>>>(h->PSI, (int)h->n, dev_block);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int op_init_model(gbx_sbm* h) {
  const int by_degree = (MODEL == MODEL_MOD) && h->mopt.dc;
  k_init_pidx<<<grid_for(h, h->n), 256, 0, h->stream>>>(h->rowptr, (int)h->n, by_degree, h->pidx);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

#define GBX_TRY_RC(expr)         \
  do {                           \
    const int _rc = (expr);      \
    if (_rc != GBX_OK) return _rc; \
  } while (0)

// ---- pull launchers ---------------------------------------------------------------------------------------------------
int pull_set_attributes() {
  if constexpr (PULL_OK) {
    const int sm = (int)PullSmem::total;
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_pull_kernel<MODE_SWEEP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_pull_kernel<MODE_SWEEP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_pull_kernel<MODE_MARGINAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_pull_kernel<MODE_MARGINAL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_pull_kernel<MODE_LOGZ, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_pull_kernel<MODE_LOGZ, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  }
  return GBX_OK;
}

PullArgs pull_args(gbx_sbm* h, int mode) {
  PullArgs a;
  a.tiles = h->tiles;
  a.ntiles = h->ntiles;
  a.rowptr = h->rowptr;
  a.col = h->col;
  a.coldeg = h->coldeg;
  a.slot_lv = h->slot_lv;
  a.pidx = h->pidx;
  a.prior_tab = h->prior_tab;
  a.theta_cur = h->theta_pp[h->thcur];
  a.theta_prev = h->theta_pp[h->th_used];
  a.PSI = h->PSI;
  // Generation g = h->gen.  A sweep reads the outgoing messages of generation g - 1 and overwrites them with those of
  // generation g + 1: buffer (g + 1) & 1; at g = 0 that buffer holds the incoming messages of generation 0 instead.
  // A marginal / log Z pass reads the same buffer without writing.
  a.msg = h->msg[(h->gen + 1) & 1];
  a.rec_old = h->rec[h->gen & 1];
  a.rec_new = h->rec[(h->gen + 1) & 1];
  a.npeer = 0;
  for (int r = 0; r < h->world; ++r)
    if (r != h->rank) a.rec_peer[a.npeer++] = h->recp[(h->gen + 1) & 1][r];
  for (int p = a.npeer; p < MAXPEER; ++p) a.rec_peer[p] = nullptr;
  a.vbeg = h->vbeg;
  a.partials = h->part_vertex;
  a.mpart = h->part_mstep;
  a.st = h->st;
  a.it = h->cur_it;
  a.dc_model = (MODEL == MODEL_MOD) ? h->mopt.dc : h->opt.degreecorrection;
  a.dc = (MODEL == MODEL_MOD) ? a.dc_model : (a.dc_model && h->theta_valid);
  a.hubs = h->hubs;
  (void)mode;
  return a;
}

template <int MODE, bool RECON>
int launch_pull_t(gbx_sbm* h) {
  if constexpr (PULL_OK) {
    GBX_TRY_RC(acquire_params(h, true));
    PullArgs a = pull_args(h, MODE);
    const int G = h->grid_pull;
    if (h->ntiles > 0) {
      const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(h->tmap[(h->gen + 1) & 1]);
      sbm_pull_kernel<MODE, RECON><<<G, P_NT, PullSmem::total, h->stream>>>(*tm, a);
      h->launches++;
    }
    if (h->nhubs > 0) {
      sbm_pull_hub_kernel<MODE, RECON><<<h->nhubs, NT, 0, h->stream>>>(a, G);
      h->launches++;
    }
    h->sweep_rows = (h->ntiles > 0 ? G : 0) + h->nhubs;
    h->sweep_row0 = h->ntiles > 0 ? 0 : G;
    GBX_CUDA_TRY(cudaGetLastError());
    if (MODE == MODE_SWEEP) {
      h->gen++;
      h->th_used = h->thcur;
    }
    return GBX_OK;
  } else {
    last_error() = "no pull kernels for this q";
    return GBX_ERR_STATE;
  }
}

int pull_vertex_pass(gbx_sbm* h, int mode) {
  const bool recon = h->gen > 0;
  if (mode == MODE_SWEEP) return recon ? launch_pull_t<MODE_SWEEP, true>(h) : launch_pull_t<MODE_SWEEP, false>(h);
  if (mode == MODE_MARGINAL) return recon ? launch_pull_t<MODE_MARGINAL, true>(h) : launch_pull_t<MODE_MARGINAL, false>(h);
  return recon ? launch_pull_t<MODE_LOGZ, true>(h) : launch_pull_t<MODE_LOGZ, false>(h);
}

int pull_fe_local(gbx_sbm* h, int pass) {
  if constexpr (PULL_OK) {
    GBX_TRY_RC(acquire_params(h, true));
    PullFeArgs f;
    f.col = h->col;
    f.erow = h->erow;
    f.coldeg = h->coldeg;
    f.msg_out = h->msg[h->gen & 1];
    f.msg_other = h->msg[(h->gen + 1) & 1];
    f.rec = h->rec[h->gen & 1];
    f.theta_cur = h->theta_pp[h->thcur];
    f.theta_prev = h->theta_pp[h->th_used];
    f.degs = h->deg_global;
    f.vbeg = h->vbeg;
    f.K = h->K;
    f.dc_model = (MODEL == MODEL_MOD) ? h->mopt.dc : h->opt.degreecorrection;
    f.dc = (MODEL == MODEL_MOD) ? f.dc_model : (f.dc_model && h->theta_valid);
    const bool recon = h->gen > 0;
    if (pass == 0) {
      if (recon) pull_fe_kernel<0, true><<<h->grid_fe, NT, 0, h->stream>>>(h->st, f, h->part_fe);
      else pull_fe_kernel<0, false><<<h->grid_fe, NT, 0, h->stream>>>(h->st, f, h->part_fe);
    } else {
      if (recon) pull_fe_kernel<1, true><<<h->grid_fe, NT, 0, h->stream>>>(h->st, f, h->part_fe);
      else pull_fe_kernel<1, false><<<h->grid_fe, NT, 0, h->stream>>>(h->st, f, h->part_fe);
    }
    constexpr int NX = (MODEL == MODEL_MOD) ? 5 : 4;
    k_local_columns<<<1, 32 * NX, 0, h->stream>>>(h->part_fe, h->grid_fe, NX, h->totals);
    h->launches += 2;
    GBX_CUDA_TRY(cudaGetLastError());
    return GBX_OK;
  } else {
    (void)pass;
    last_error() = "no pull kernels for this q";
    return GBX_ERR_STATE;
  }
}

// Tensor maps of the two message buffers (rows of QM doubles, boxes of 32 rows, hardware swizzle of the row size) and
// the neighbour degrees.  The driver entry point is looked up at run time: the library links no libcuda.
int op_pull_setup(gbx_sbm* h) {
  if constexpr (PULL_OK) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      GBX_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
      if (!fn || qres != cudaDriverEntryPointSuccess) {
        last_error() = "cuTensorMapEncodeTiled is not available in this driver";
        return GBX_ERR_CUDA;
      }
      encode = reinterpret_cast<EncodeFn>(fn);
    }
    const cuuint64_t rows = (cuuint64_t)std::max<int64_t>(h->K, 1);
    for (int b = 0; b < 2; ++b) {
      const cuuint64_t gdim[2] = {(cuuint64_t)QM, rows};
      const cuuint64_t gstr[1] = {(cuuint64_t)ROWB};
      const cuuint32_t box[2] = {(cuuint32_t)QM, 32u};
      const cuuint32_t estr[2] = {1u, 1u};
      const CUtensorMapSwizzle sw = (ROWB == 32) ? CU_TENSOR_MAP_SWIZZLE_32B
                                                 : (ROWB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
      const CUresult r = encode(reinterpret_cast<CUtensorMap*>(h->tmap[b]), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->msg[b], gdim,
                                gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        last_error() = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")";
        return GBX_ERR_CUDA;
      }
    }
    if (h->K > 0) {
      if (h->world > 1) k_coldeg_global<<<grid_for(h, h->K), 256, 0, h->stream>>>(h->col, h->deg_global, (long long)h->K, h->coldeg);
      else k_coldeg<<<grid_for(h, h->K), 256, 0, h->stream>>>(h->col, h->rowptr, (long long)h->K, h->coldeg);
      h->launches++;
      GBX_CUDA_TRY(cudaGetLastError());
    }
    h->tmap_ok = true;
    return GBX_OK;
  } else {
    return GBX_OK;
  }
}

// Initial messages (K x q in `links` order on the device, or drawn from `seed`) into both layouts: msg[0] outgoing,
// msg[1] incoming.
int op_pull_load(gbx_sbm* h, const double* src_dev, uint64_t seed) {
  if (h->K > 0) {
    if (src_dev)
      k_pull_permute<<<grid_for(h, h->K), 256, 0, h->stream>>>(src_dev, nullptr, h->msg[1], h->msg[0], h->in_link, h->out_link,
                                                              (long long)h->K, 1);
    else
      k_pull_random_messages<<<grid_for(h, h->K), 256, 0, h->stream>>>(h->msg[1], h->msg[0], h->in_link, h->out_link,
                                                                      (long long)h->K, seed);
    h->launches++;
    GBX_CUDA_TRY(cudaGetLastError());
  }
  return GBX_OK;
}

int op_pull_export(gbx_sbm* h, double* dst_dev) {
  // rows of links whose source vertex another rank owns come back as zeros (as on the push layout)
  if (h->world > 1) GBX_CUDA_TRY(cudaMemsetAsync(dst_dev, 0, (size_t)h->K_global * Q * sizeof(double), h->stream));
  if (h->K > 0) {
    k_pull_permute<<<grid_for(h, h->K), 256, 0, h->stream>>>(nullptr, dst_dev, nullptr, h->msg[h->gen & 1], h->in_link,
                                                            h->out_link, (long long)h->K, 0);
    h->launches++;
    GBX_CUDA_TRY(cudaGetLastError());
  }
  return GBX_OK;
}

// Replaces the rows of M-step partials the sweep just wrote (one per sweep CTA and hub) by statistics paired the
// reference's way; sbm_local_totals sums whatever is there.
int op_mstep_exact(gbx_sbm* h) {
  if (MODEL != MODEL_SBM || h->world > 1) return GBX_OK;
  GBX_TRY_RC(acquire_params(h, true));
  const double* cur = h->pull ? h->msg[h->gen & 1] : h->msg[h->cur];
  const int rows = h->sweep_rows;
  if (rows > 0) {
    sbm_mstep_exact_kernel<<<rows, NT, 0, h->stream>>>(cur, h->rev, (long long)h->K,
                                                     h->part_mstep + (size_t)h->sweep_row0 * Q * Q, h->st, h->cur_it);
    h->launches++;
    GBX_CUDA_TRY(cudaGetLastError());
  }
  return GBX_OK;
}

int op_occupancy(int variant, int* vertex_occ, int* mstep_occ) {
  (void)variant;
  {
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_vertex_kernel<MODE_SWEEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepSmem::total));
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_vertex_kernel<MODE_MARGINAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepSmem::total));
    GBX_CUDA_TRY(cudaFuncSetAttribute(sbm_vertex_kernel<MODE_LOGZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepSmem::total));
    GBX_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(vertex_occ, sbm_vertex_kernel<MODE_SWEEP>, NT, SweepSmem::total));
  }
  *mstep_occ = 0;
  if constexpr (PULL_OK) {
    GBX_TRY_RC(pull_set_attributes());
    GBX_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(mstep_occ, sbm_pull_kernel<MODE_SWEEP, true>, P_NT, PullSmem::total));
  }
  return GBX_OK;
}

#if GBX_MODEL == 1
// =============================================================================================
// mod.jl: SBM restricted to community structure (omega_in on the diagonal, omega_out off it), src/mod.jl:3-244.
// The sweep reuses sbm_vertex_kernel / sbm_hub_kernel (edge factor 1 + psi (e^beta - 1), prior field
// alpha beta d_i h); below are the scalar bookkeeping, the omega M-step and the free energy of that model.
// =============================================================================================
// mod.jl:72 (residual), 22-24 (h numerator), 221-223 (gr), 77-112 and 224-229 (omegas, alpha, beta) from the totals.
__device__ __forceinline__ void mod_apply_totals_body(SbmState* st, const double* __restrict__ tot, const double* maxd, int mode,
                                                      int prior, int dc, int learning, double N, double K, double lr,
                                                      double maxomega, int it, double thr) {
  const int tid = threadIdx.x;
  if (skip_iteration(st, it)) return;
  if (mode == MODE_LOGZ) {
    if (tid == 0) {
      st->sumlogZi = tot[1];
      if (!(tot[1] == tot[1])) st->status |= 1;
    }
    return;
  }
  __shared__ double hs[Q];
  if (tid < Q) {
    const double snew = tot[2 + tid];
    st->sumPSI[tid] = snew;
    hs[tid] = dc ? tot[2 + 3 * Q + tid] : snew;                     // degrees' * PSI, mod.jl:22-24 (ones without dc)
    st->Stheta[tid] = hs[tid];
    if (mode == MODE_SWEEP && prior) st->gr[tid] = snew / N;         // gr = update_gr(PSI), mod.jl:221-223
  }
  __syncthreads();
  if (tid == 0 && mode == MODE_SWEEP) {
    st->conv = tot[0] / N;  // mod.jl:72
    if (it != 0 && st->done_at == 0 && st->conv < thr) st->done_at = stamp_of(st, it);   // mod.jl:230-234, acted on from it + 1
    st->maxd = *maxd;
    if (!(tot[0] == tot[0]) || !(tot[0] < INFINITY)) st->status |= 1;
    st->eb1_used = st->eb1_cur;                                     // what the sweep just ran with (pull sweeps)
    if (learning) {
      // The statistic sum_e omega_in s_e / ((omega_in - omega_out) s_e + omega_out) (mod.jl:86-87) was accumulated by
      // the sweep over ALL directed slots (new outgoing message . previous incoming one): twice per undirected edge.
      const double omegainnew = 0.5 * tot[2 + 4 * Q];
      double dn = 0.0;
      for (int t = 0; t < Q; ++t) dn += hs[t] * hs[t];                // norm(update_h())^2 of the new marginals
      double oin = 2.0 * omegainnew / dn;
      double oout = 2.0 * (0.5 * K - omegainnew) / ((dc ? K * K : N * N) - dn);
      if (oin > maxomega) oin = maxomega;                             // mod.jl:97-103
      else if (oin < 0.0) oin = 0.0;
      else if (oout < 0.0) oout = 0.0;
      // mod.jl:226-227.  update_omega() has already assigned EM()'s own omegain / omegaout (closure write-through,
      // mod.jl:90-94), so the reference blends the new value with itself: the learning rate does not act here.
      st->omegain = lr * oin + (1.0 - lr) * oin;
      st->omegaout = lr * oout + (1.0 - lr) * oout;
      st->beta = log(st->omegain / st->omegaout);                     // update_alphabeta, mod.jl:108-112
      st->alpha = (st->omegain - st->omegaout) / st->beta;
      if (!(st->beta == st->beta) || !(st->alpha == st->alpha)) st->status |= 1;
    }
  }
}
__global__ void mod_apply_totals(SbmState* st, const double* __restrict__ tot, const double* maxd, int mode, int prior,
                                 int dc, int learning, double N, double K, double lr, double maxomega, int it, double thr) {
  mod_apply_totals_body(st, tot, maxd, mode, prior, dc, learning, N, K, lr, maxomega, it, thr);
}

__global__ void mod_prepare(SbmState* st, SbmParams* out, double N, int zero_h, int it) {
  if (next_iteration_skips(st, it)) return;
  const int tid = threadIdx.x;
  if (tid < Q) {
    out->lgr[tid] = log(st->gr[tid]);
    const double hh = zero_h ? 0.0 : st->Stheta[tid];  // h = update_h(), mod.jl:56 (zeros before the first sweep)
    st->h[tid] = hh;
    out->h[tid] = st->alpha * st->beta * hh;            // the prior field is alpha beta d_i h, mod.jl:46-48
  }
  if (tid == 0) {
    out->logN = log(N);
    const double eb1 = exp(st->beta) - 1.0;             // written as in mod.jl:42
    out->eb1 = eb1;
    st->eb1_cur = eb1;
    if (zero_h) st->eb1_used = eb1;                     // init: no sweep has run yet
    out->eb1_prev = st->eb1_used;
    out->oin = st->omegain;
    out->oout = st->omegaout;
    double hn2 = 0.0;
    for (int t = 0; t < Q; ++t) {
      const double hh = zero_h ? 0.0 : st->Stheta[t];
      hn2 += hh * hh;
    }
    st->hn2 = hn2;
    // edge factors lie in [1, 1 + eb1] (or [1 + eb1, 1] for beta < 0): bound for the plain-product path
    const double worst = fmax(fabs(log2(1.0 + eb1)), 1e-3);
    out->safe_deg = (worst == worst && worst < INFINITY) ? (int)fmin(900.0 / worst, (double)TILE_E) : 0;
  }
}

// freeenergy(), mod.jl:127-153.  x0 = logZij, x1 = CVBP, x2 = CVGP, x3 = CVGT, x4 = CVMAP per undirected edge.
template <int PASS>
__global__ void __launch_bounds__(NT) mod_fe_kernel(const SbmState* st, const int* __restrict__ rev,
                                                    const int* __restrict__ col, const int* __restrict__ erow,
                                                    const double* __restrict__ degs, long long L,
                                                    const double* __restrict__ psi, PeerCur peers, int dc, int rank,
                                                    double* partials) {
  __shared__ double red[(NT / 32) * 5];
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  double mean[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  if (PASS == 1)
    for (int k = 0; k < 5; ++k) mean[k] = st->cv5mean[k];
  const double oin = st->omegain, oout = st->omegaout, beta = st->beta;
  const double loin = log(oin), loout = log(oout);
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < L; i += (long long)gridDim.x * NT) {
    const int2 er = make_int2((int)i, rev[i]);
    if (!(rev_owner(er.y) < rank || (rev_owner(er.y) == rank && rev_slot(er.y) < er.x))) continue;  // each edge once
    double a[Q], b[Q];
    load_msg<Q, QM>(psi, er.x, b);
    load_msg<Q, QM>(peers.p[rev_owner(er.y)], rev_slot(er.y), a);
    double di = 1.0, dj = 1.0;
    if (dc) {  // degrees of both endpoints: `degs` is indexed by global vertex id (every rank's, like theta)
      di = degs[erow[er.x]];
      dj = degs[col[er.x]];
    }
    double x[5];
    fe_terms_mod(a, b, di, dj, oin, oout, beta, loin, loout, x);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      if (PASS == 0) acc[k] += x[k];
      else acc[k] += (x[k] - mean[k]) * (x[k] - mean[k]);
    }
  }
  block_reduce_store<5>(acc, red, partials + (size_t)blockIdx.x * 5);
}

__global__ void mod_apply_fe(SbmState* st, const double* __restrict__ totals, int pass, double L) {
  const int c = threadIdx.x;
  if (c < 5) {
    if (pass == 0) {
      st->cv5sum[c] = totals[c];
      st->cv5mean[c] = totals[c] / L;
    } else {
      st->cv5ss[c] = totals[c];
    }
  }
}

__global__ void mod_finalize(SbmState* st, double N, double L, double constshift) {
  if (threadIdx.x == 0) {
    const double al = st->alpha, be = st->beta, oout = st->omegaout;
    st->result[0] = -(st->sumlogZi - st->cv5sum[0] + 0.5 * al * be * st->hn2) / (be * N) -
                    (L / N) * (2.0 * L * oout + log(oout) + constshift) / be;          // FE, mod.jl:158
    st->result[1] = -st->cv5sum[1] / L + 1.0;                                        // CVBayes, mod.jl:159
    st->result[2] = -st->cv5sum[2] / L - log(oout) + 1.0 - constshift;               // CVGP,    mod.jl:160
    st->result[3] = -st->cv5sum[3] / L + 1.0 - constshift;                           // CVGT,    mod.jl:161
    st->result[4] = -st->cv5sum[4] / L + 1.0 - constshift;                           // CVMAP,   mod.jl:162
    st->result[5] = st->cv5ss[0] / (L - 1.0);                                        // var(logZijs), mod.jl:163
    st->result[6] = st->cv5ss[2] / (L - 1.0);
    st->result[7] = st->cv5ss[3] / (L - 1.0);
    st->result[8] = st->cv5ss[4] / (L - 1.0);
  }
}

int mop_apply_totals(gbx_sbm* h, int mode) {
  if (h->totals_applied) {   // the last block of sbm_local_totals has done it
    h->totals_applied = false;
    return GBX_OK;
  }
  mod_apply_totals<<<1, 64, 0, h->stream>>>(h->st, h->totals, h->maxd_dev, mode, h->mopt.priorlearning, h->mopt.dc,
                                            h->mopt.learning, (double)h->n_global, (double)h->K_global,
                                            h->mopt.learningrate, h->mopt.maxomega, h->cur_it, h->cur_thr);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int mop_prepare(gbx_sbm* h, int zero_h, int) {
  mod_prepare<<<1, 32, 0, h->stream>>>(h->st, h->params, (double)h->n_global, zero_h, h->cur_it);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  { const int rc = acquire_params(h, false); if (rc != GBX_OK) return rc; }
  GBX_CUDA_TRY(cudaMemcpyToSymbolAsync(c_P, h->params, sizeof(SbmParams), 0, cudaMemcpyDeviceToDevice, h->stream));
  const int dc_tab = h->mopt.dc;
  const int nrows = dc_tab ? 1 + (TILE_E + 1) * Q : 1;
  sbm_prior_table<<<(nrows + 127) / 128, 128, 0, h->stream>>>(h->st, h->prior_tab, dc_tab, h->cur_it);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int mop_noop(gbx_sbm*) { return GBX_OK; }

PeerCur peer_cur(const gbx_sbm* h) {
  PeerCur p;
  for (int r = 0; r < MAXPEER; ++r) p.p[r] = h->msgp[h->cur][r];
  return p;
}

int mop_fe_local(gbx_sbm* h, int pass) {
  if (h->pull) return pull_fe_local(h, pass);
  { const int rc = acquire_params(h, true); if (rc != GBX_OK) return rc; }
  if (pass == 0)
    mod_fe_kernel<0><<<h->grid_fe, NT, 0, h->stream>>>(h->st, h->rev, h->col, h->erow, h->deg_global, (long long)h->K,
                                                       h->msg[h->cur], peer_cur(h), h->mopt.dc, h->rank, h->part_fe);
  else
    mod_fe_kernel<1><<<h->grid_fe, NT, 0, h->stream>>>(h->st, h->rev, h->col, h->erow, h->deg_global, (long long)h->K,
                                                       h->msg[h->cur], peer_cur(h), h->mopt.dc, h->rank, h->part_fe);
  k_local_columns<<<1, 160, 0, h->stream>>>(h->part_fe, h->grid_fe, 5, h->totals);
  h->launches += 2;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int mop_fe_apply(gbx_sbm* h, int pass) {
  mod_apply_fe<<<1, 32, 0, h->stream>>>(h->st, h->totals, pass, 0.5 * (double)h->K_global);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int mop_fe_final(gbx_sbm* h) {
  // degrees = ones without dc (mod.jl:189-191), hence constshift = 0 then
  mod_finalize<<<1, 32, 0, h->stream>>>(h->st, (double)h->n_global, 0.5 * (double)h->K_global,
                                        h->mopt.dc ? h->constshift : 0.0);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

const SbmOps kOps = {Q,          op_vertex_pass, op_local_totals, mop_apply_totals,   mop_prepare,   mop_noop,     mop_noop,
                     mop_noop,   mop_fe_local,   mop_fe_apply,    mop_fe_final,       op_permute,    op_random_messages,
                     op_assignment, op_occupancy, op_init_model, op_unpack, op_release, op_mstep_exact,
                     PULL_OK ? 1 : 0, op_pull_setup, op_pull_load, op_pull_export};
#else
const SbmOps kOps = {Q,          op_vertex_pass, op_local_totals, op_apply_totals,    op_prepare,    op_theta_local, op_theta_apply,
                     op_escale,  op_fe_local,    op_fe_apply,     op_fe_final,        op_permute,    op_random_messages,
                     op_assignment, op_occupancy, op_init_model, op_unpack, op_release, op_mstep_exact,
                     PULL_OK ? 1 : 0, op_pull_setup, op_pull_load, op_pull_export};
#endif

}  // namespace

#define GBX_CAT_(a, b) a##b
#define GBX_CAT(a, b) GBX_CAT_(a, b)
#if GBX_MODEL == 1
const SbmOps* GBX_CAT(mod_ops_q, GBX_Q)() { return &kOps; }
#else
const SbmOps* GBX_CAT(sbm_ops_q, GBX_Q)() { return &kOps; }
#endif

}  // namespace gbx
