// Labeled stochastic block model (edge labels alpha = 1..G, one affinity matrix and one degree sequence per label):
// EM + BP of /root/reference/src/labeled_sbm/labeled-sbmBIX_0.2.3.8_Julia0.6.2.ipynb, cell 0, `function EM`.
//
// A handle holds ONE graph or MANY (gbx_lsbm_create_batch: the disjoint union of `count` graphs in one CSR, BASELINE
// configs[4]: thousands of N = 10k graphs).  Every kernel is launched on a 2-D grid whose y index names a graph: a block
// works on the vertex / slot range of its graph with that graph's parameter block (LParams), so one launch set per EM
// iteration serves the whole batch, graphs converge one by one (their blocks return at once, the host drops them from the
// list of live graphs) and a batch of one is the single-graph path -- same kernels, same arithmetic.
//
// Arithmetic: the notebook works in the log domain (one log per message component per update).  Here the product of a
// vertex's edge terms is carried as mantissa x 2^exponent per component and one log per component is taken per VERTEX;
// a cavity message is the vertex's weight divided by the edge term (a reciprocal, no log / exp per edge), with the
// absolute clamp of normalize_logprob applied to the same true magnitudes.  Synchronous (Jacobi) sweeps like the other
// models: same fixed point as the notebook's asynchronous sweep.
//
//   k_lsbm_vertex<MODE>   logunnormalizedMessage / BP / updatePSI / sum log Z_i      (notebook: same names)
//   k_lsbm_denom          denom[alpha,:] = sum_i degrees[i,alpha] PSI[i,:], sum PSI   (update_affinity, update_h, update_gr)
//   k_lsbm_mstep          affinitynew[alpha] += PSIij / sum(PSIij) over i < j          (update_affinity)
//   k_lsbm_pre / k_lsbm_post   h = update_h(); gr and affinity updates with the learning rate
//   k_lsbm_fe<PASS>       log Z^{ij}, Gibbs / MAP errors per edge                     (freeenergy)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "gbx_sbm_internal.h"

namespace gbx {
int build_graph_device(gbx_sbm* h, const int64_t* links_host, std::vector<int>* rowptr_host, GlobalGraph* keep_global,
                       const std::function<void()>* while_sorting = nullptr);
}
using namespace gbx;

namespace {

constexpr int LMAXG = GBX_MAX_LABELS;
constexpr double LOG1EM8 = -18.420680743952367;  // log(10^-8)
enum { LMODE_SWEEP = 0, LMODE_MARGINAL = 1, LMODE_LOGZ = 2 };

struct LParams {
  double gr[MAXQ];
  double lgr[MAXQ];                  // log(gr), refreshed whenever gr changes (k_lsbm_post; gbx_lsbm_init)
  double aff[LMAXG][MAXQ * MAXQ];   // row-major B x B per label
  double h[LMAXG][MAXQ];
  double denom[LMAXG][MAXQ];
  double sumPSI[MAXQ];
  double Gnew[LMAXG][MAXQ * MAXQ];
  double conv, sumlogZi;
  double fesum[4], femean[4], fess[4];
  double result[9];
  int status;                        // bit 0: non-finite value seen; bit 1: the two directions of a link differ in label
  int parity;                        // message buffer that holds this graph's current messages (flipped by every sweep
                                     // that really ran: a graph that has converged keeps its buffer)
  int pad_;
  int done_at;                       // EM iteration whose residual fell below the threshold (0: not yet); kernels stamped
                                     // with a later iteration return at once (the host queues several iterations)
};

// One graph of the handle: its vertices [vbeg, vbeg + n) and slots [sbeg, sbeg + K) inside the union.
struct LGraph {
  int vbeg, n;
  long long sbeg, K;
  double N;
};

#define L_TRY(expr)              \
  do {                           \
    int _r = (expr);             \
    if (_r != GBX_OK) return _r; \
  } while (0)

int l_invalid(const char* msg) {
  last_error() = msg;
  return GBX_ERR_INVALID;
}

// ---- small device helpers -----------------------------------------------------------------------------------------
__device__ __forceinline__ bool l_skip(const LParams* P, int it) {   // see LParams::done_at; it == 0: never skip
  if (it == 0) return false;
  const int d = *reinterpret_cast<const volatile int*>(&P->done_at);
  return d != 0 && d < it;
}
// graph of this block: grid y index, through the list of live graphs when there is one
__device__ __forceinline__ int l_graph(const int* __restrict__ live) { return live ? live[blockIdx.y] : (int)blockIdx.y; }
// exponent bookkeeping of a running product: m stays in [1,2), k collects the exponents (zero / non-finite left alone)
__device__ __forceinline__ void l_renorm(double& m, int& k) {
  bool ok;
  const int ex = exponent_of(m, ok);
  if (ok) {
    m *= pow2i(-ex);
    k += ex;
  }
}
// 1/x for positive normal x: MUFU.RCP64H seed + two Newton steps
__device__ __forceinline__ double l_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
}
__device__ __forceinline__ double logsumexp_dev(const double* x, int q) {
  double m = -INFINITY;
  for (int t = 0; t < q; ++t) m = fmax(m, x[t]);
  double s = 0.0;
  for (int t = 0; t < q; ++t) s += exp(((m - x[t] > 500.0) ? m - 500.0 : x[t]) - m);
  return m + log(s);
}
__device__ __forceinline__ double block_sum(double v, double* red) {  // 256 threads, fixed order
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] += red[tid + o];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// Message rows: 256/128-bit accesses when q is a compile-time constant (gbx_common.cuh), scalar otherwise.
template <int QT, int QA>
__device__ __forceinline__ void row_load(const double* __restrict__ base, long long idx, int q, double (&m)[QA]) {
  if constexpr (QT > 0) {
    load_vec<QT>(base, idx, m);
  } else {
    for (int t = 0; t < q; ++t) m[t] = base[idx * q + t];
  }
}
template <int QT, int QA>
__device__ __forceinline__ void row_store(double* __restrict__ base, long long idx, int q, const double (&m)[QA]) {
  if constexpr (QT > 0) {
    store_vec<QT>(base, idx, m);
  } else {
    for (int t = 0; t < q; ++t) base[idx * q + t] = m[t];
  }
}

__global__ void k_lab_slots(const int64_t* __restrict__ lab_link, const int* __restrict__ slot_of_link, long long K, int G,
                            unsigned char* __restrict__ lab, int* err) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < K; k += (long long)gridDim.x * blockDim.x) {
    const long long a = lab_link[k];
    if (a < 1 || a > G) {
      atomicOr(err, 1);
      continue;
    }
    lab[slot_of_link[k]] = (unsigned char)(a - 1);
  }
}
__global__ void k_lab_check(const unsigned char* __restrict__ lab, const int* __restrict__ rev, long long K, int* err) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x)
    if (lab[rev[e]] != lab[e]) atomicOr(err, 2);
}
// degrees[i, alpha] = number of alpha-labelled links at i (notebook EM, "initial state"); ones without degree correction
__global__ void k_lsbm_degrees(const int* __restrict__ rowptr, const unsigned char* __restrict__ lab, int n, int G, int dc,
                               double* __restrict__ deg) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double c[LMAXG];
    for (int a = 0; a < G; ++a) c[a] = dc ? 0.0 : 1.0;
    if (dc)
      for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) c[lab[e]] += 1.0;
    for (int a = 0; a < G; ++a) deg[(size_t)i * G + a] = c[a];
  }
}

// degrees[i, alpha] degrees[j, alpha] Ntot of every slot (row i, column j, label alpha): fixed for the whole run, so the
// sweep reads one streamed value per slot instead of gathering the neighbour's degree twice per iteration.
__global__ void k_lsbm_dd(const LGraph* __restrict__ gt, int G, const int* __restrict__ erow, const int* __restrict__ col,
                          const unsigned char* __restrict__ lab, const double* __restrict__ deg, double* __restrict__ dd) {
  const LGraph gg = gt[blockIdx.y];
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < gg.K; k += (long long)gridDim.x * blockDim.x) {
    const long long e = gg.sbeg + k;
    const int a = lab[e];
    dd[e] = deg[(size_t)erow[e] * G + a] * deg[(size_t)col[e] * G + a] * gg.N;
  }
}

// ---- the sweep ----------------------------------------------------------------------------------------------------
// LPV lanes per vertex (1 when the launch has enough vertices to fill the GPU, else 8, 16 or 32 from the mean degree), one
// lane per incoming link (rounds of LPV for larger degrees): the lanes read the vertex's contiguous run of messages
// coalesced, each forms the term of its link, degrees[i,a] degrees[s,a] Ntot (psi_{s->i} affinity[a]), and multiplies it
// into a running product kept as mantissa x 2^exponent per component (logunnormalizedMessage sums the logs of these
// terms: one log per component per VERTEX here).  The cavity message of a link (BP: logPSIi - log(term), then
// normalize_logprob) is the vertex's weight divided by the link's term; the clamp at log(1e-8) becomes a floor in the
// same scale, and the 500 cut-off of logsumexp cannot bind above that floor by more than e^-500.
// QT > 0: q known at compile time (loops unrolled, per-lane vectors in registers); QT == 0: any q <= 16.
template <int QT, int QA>
__device__ __forceinline__ void l_term(const double (&psi)[QA], const double* __restrict__ A, double dd, int q, double (&term)[QA]) {
#pragma unroll
  for (int t = 0; t < QA; ++t) {
    if (t < q) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < QA; ++r)
        if (r < q) s += psi[r] * A[r * q + t];
      term[t] = dd * s;
    }
  }
}

template <int MODE, int QT, int LPV>
#ifndef GBX_LSBM_MINB
#define GBX_LSBM_MINB 1
#endif
__global__ void __launch_bounds__(256, GBX_LSBM_MINB) k_lsbm_vertex(const LGraph* __restrict__ gt, const int* __restrict__ live, int qrt, int G,
                                                     const int* __restrict__ rowptr, const int* __restrict__ rev,
                                                     const unsigned char* __restrict__ lab, const double* __restrict__ deg,
                                                     const double* __restrict__ ddN, double* __restrict__ msg0,
                                                     double* __restrict__ msg1, double* __restrict__ PSI, LParams* Pall,
                                                     double damping, int it, double* __restrict__ pv, int xcap) {
  __shared__ double red[256];
  const int gi = l_graph(live);
  LParams* P = Pall + gi;
  if (MODE == LMODE_SWEEP && l_skip(P, it)) return;
  const LGraph gg = gt[gi];
  const int n = gg.n;
  const double N = gg.N;
  const int par = P->parity;
  const double* __restrict__ cur = par ? msg1 : msg0;
  double* __restrict__ nxt = par ? msg0 : msg1;
  const int q = QT > 0 ? QT : qrt;
  constexpr int QA = QT > 0 ? QT : MAXQ;
  constexpr int GPB = 256 / LPV;                        // vertices per block and round
  const int lane = threadIdx.x % LPV, grp = threadIdx.x / LPV;
  double acc = 0.0;  // residual (sweep) or sum of log Z_i, kept by lane 0 of every group
  for (int i0 = blockIdx.x * GPB; i0 < n; i0 += gridDim.x * GPB) {   // block-uniform trip count
    const bool valid = i0 + grp < n;
    const int i = gg.vbeg + i0 + grp;
    const int e0 = valid ? rowptr[i] : 0, e1 = valid ? rowptr[i + 1] : 0;
    const bool single = (e1 - e0) <= LPV;   // every lane keeps the term of its one link in registers
    double x[QA], term[QA], pm[QA];
    int pk[QA];
#pragma unroll
    for (int t = 0; t < QA; ++t) {
      pm[t] = 1.0;
      pk[t] = 0;
      term[t] = 1.0;
    }
    // thread-per-vertex launches with few groups keep the terms of a vertex's first NC links in registers: the second
    // pass (cavities) then reads neither the messages nor the per-slot scales again for vertices of degree <= NC
    constexpr int NC = (LPV == 1 && QT > 0 && QT <= 4 && MODE == LMODE_SWEEP) ? (QT <= 2 ? 12 : (QT == 3 ? 8 : 6)) : 0;
    double tc[NC > 0 ? NC : 1][QA];
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      const int e = e0 + r + lane;   // LPV == 1: lane == 0
      if (e < e1) {
        double psi[QA];
        row_load<QT, QA>(cur, e, q, psi);
        l_term<QT, QA>(psi, P->aff[lab[e]], ddN[e], q, tc[r]);
#pragma unroll
        for (int t = 0; t < QA; ++t) {
          if (t < q) {
            pm[t] *= tc[r][t];
            l_renorm(pm[t], pk[t]);
          }
        }
      }
    }
    for (int base = e0 + NC * LPV; base < e1; base += LPV) {
      const int e = base + lane;
      if (e < e1) {
        double psi[QA];
        row_load<QT, QA>(cur, e, q, psi);
        l_term<QT, QA>(psi, P->aff[lab[e]], ddN[e], q, term);
#pragma unroll
        for (int t = 0; t < QA; ++t) {
          if (t < q) {
            pm[t] *= term[t];
            l_renorm(pm[t], pk[t]);
          }
        }
      }
    }
    __syncwarp();   // groups of one warp may have walked different numbers of rounds
#pragma unroll
    for (int t = 0; t < QA; ++t) {
      if (t < q) {
        double m = pm[t];
        int k = pk[t];
#pragma unroll
        for (int o = LPV / 2; o > 0; o >>= 1) {
          m *= __shfl_xor_sync(0xffffffffu, m, o);
          k += __shfl_xor_sync(0xffffffffu, k, o);
          l_renorm(m, k);
        }
        double ext = 0.0;
        if (valid)
          for (int a = 0; a < G; ++a) ext += deg[(size_t)i * G + a] * P->h[a][t];
        // sum of the log terms, then prior and field as logunnormalizedMessage adds them
        x[t] = (log(m) + (double)k * 0.6931471805599453) + (P->lgr[t] - ext);
      }
    }
    if (MODE == LMODE_LOGZ) {
      if (valid && lane == 0) acc += logsumexp_dev(x, q);
      continue;
    }
    // normalize_logprob(x): weights in the scale of the largest component, the clamp at log(1e-8) as a floor in that
    // scale (the 500 cut-off of logsumexp cannot bind above the floor by more than e^-500)
    double M = -INFINITY;
#pragma unroll
    for (int t = 0; t < QA; ++t)
      if (t < q) M = fmax(M, x[t]);
    const double lfl = LOG1EM8 - M;
    const bool all_floor = lfl > 700.0;          // every component is below 1e-8: flat
    const double fl = all_floor ? 1.0 : exp(lfl);
    double y[QA], w[QA];
    {
      double S = 0.0;
#pragma unroll
      for (int t = 0; t < QA; ++t) {
        if (t < q) {
          y[t] = exp(x[t] - M);
          double c = (y[t] < fl) ? fl : y[t];
          if (all_floor) c = 1.0;
          w[t] = c;
          S += c;
        }
      }
      const double inv = 1.0 / S;
#pragma unroll
      for (int t = 0; t < QA; ++t)
        if (t < q) w[t] *= inv;
    }
    if (valid && lane == 0) {
      for (int t = 0; t < q; ++t) {
        if (MODE == LMODE_SWEEP) acc += fabs(w[t] - PSI[(size_t)i * q + t]) / N;
        if (!(w[t] == w[t])) P->status |= 1;
      }
    }
    __syncwarp();   // lane 0 has read the old marginal
    if (MODE == LMODE_SWEEP) {
      // cavities: the vertex's weights y (scale of its largest component) over the link's term, same floor
      auto emit = [&](int e, const double (&tm)[QA]) {
        double m[QA];
        double S = 0.0;
#pragma unroll
        for (int t = 0; t < QA; ++t) {
          if (t < q) {
            double c = y[t] * l_rcp(tm[t]);
            c = (c < fl) ? fl : c;               // normalize_logprob: array[r] < log(1e-8) -> log(1e-8)
            if (all_floor) c = 1.0;
            m[t] = c;
            S += c;
          }
        }
        const double inv = 1.0 / S;
#pragma unroll
        for (int t = 0; t < QA; ++t)
          if (t < q) m[t] *= inv;
        const long long rv = rev[e];
        if (damping != 1.0) {
          double old[QA];
          row_load<QT, QA>(cur, rv, q, old);
          for (int t = 0; t < q; ++t) m[t] = (1.0 - damping) * old[t] + damping * m[t];
        }
        row_store<QT, QA>(nxt, rv, q, m);
      };
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const int e = e0 + r + lane;
        if (e < e1) emit(e, tc[r]);
      }
      for (int base = e0 + NC * LPV; base < e1; base += LPV) {
        const int e = base + lane;
        if (e < e1) {
          if (!single) {
            double psi[QA];
            row_load<QT, QA>(cur, e, q, psi);
            l_term<QT, QA>(psi, P->aff[lab[e]], ddN[e], q, term);
          }
          emit(e, term);
        }
      }
    }
    if (valid && lane == 0)
      for (int t = 0; t < q; ++t) PSI[(size_t)i * q + t] = w[t];
  }
  if (MODE != LMODE_MARGINAL) {
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) pv[(size_t)gi * xcap + blockIdx.x] = s;   // residual / sum log Z_i of this block (k_lsbm_reduce)
  }
}

// denom[alpha, t] = sum_i degrees[i, alpha] PSI[i, t] and sum_i PSI[i, t].  A thread owns one component t (the first
// (256 / q) q threads of a block walk the graph's PSI elements coalesced), accumulates in registers, and the block adds
// the threads of equal t in a fixed order: one atomic per (alpha, t) and block.
__global__ void __launch_bounds__(256) k_lsbm_denom(const LGraph* __restrict__ gt, const int* __restrict__ live, int q, int G,
                                                    const double* __restrict__ deg, const double* __restrict__ PSI,
                                                    LParams* Pall, int it, double* __restrict__ pd, int xcap) {
  __shared__ double sh[LMAXG + 1][256];
  const int gi = l_graph(live);
  LParams* P = Pall + gi;
  if (l_skip(P, it)) return;
  const LGraph gg = gt[gi];
  const int tid = threadIdx.x;
  const int vpb = 256 / q;                       // vertices per block and round
  const int t = tid % q, vl = tid / q;
  double acc[LMAXG + 1];
#pragma unroll
  for (int a = 0; a <= LMAXG; ++a) acc[a] = 0.0;
  if (vl < vpb) {
    for (int i0 = blockIdx.x * vpb + vl; i0 < gg.n; i0 += gridDim.x * vpb) {
      const int i = gg.vbeg + i0;
      const double p = PSI[(size_t)i * q + t];
      acc[LMAXG] += p;
#pragma unroll
      for (int a = 0; a < LMAXG; ++a)
        if (a < G) acc[a] += deg[(size_t)i * G + a] * p;
    }
  }
#pragma unroll
  for (int a = 0; a <= LMAXG; ++a) sh[a][tid] = acc[a];
  __syncthreads();
  if (tid < (LMAXG + 1) * q) {
    const int a = tid / q, tt = tid % q;
    double s = 0.0;
    if (a == LMAXG || a < G)
      for (int k = tt; k < vpb * q; k += q) s += sh[a][k];
    pd[((size_t)gi * xcap + blockIdx.x) * (LMAXG + 1) * q + a * q + tt] = s;   // summed over the blocks by k_lsbm_reduce
  }
}

// affinitynew[alpha] += PSIij / sum(PSIij), once per undirected link (i < j); the factor degrees * degrees * Ntot of
// the notebook's PSIij cancels in the normalisation, and affinity[alpha][r,s] is the same for every link of a label:
//     affinitynew[alpha][r,s] = affinity[alpha][r,s] * sum_links (psi_{i->j}[r] / Z_link) psi_{j->i}[s],
// a sum of outer products.  A warp takes 32 slots at a time: every lane normalises its link and parks the two vectors
// in shared memory, then every lane accumulates the (r,s) entries it owns over the 32 links in registers.
// Runs after the sweep and before k_lsbm_post flips the graph's buffer: the new messages are in buffer parity ^ 1.
constexpr int MSTEP_THREADS = 128;
template <int QT>
__global__ void __launch_bounds__(MSTEP_THREADS) k_lsbm_mstep(const LGraph* __restrict__ gt, const int* __restrict__ live, int qrt,
                                                              int G, const int* __restrict__ erow, const int* __restrict__ col,
                                                              const int* __restrict__ rev, const unsigned char* __restrict__ lab,
                                                              const double* __restrict__ msg0, const double* __restrict__ msg1,
                                                              LParams* Pall, int it, double* __restrict__ pm, int xcap) {
  extern __shared__ double dyn[];
  const int gi = l_graph(live);
  LParams* P = Pall + gi;
  if (l_skip(P, it)) return;                       // [warps][32][2 q]
  const LGraph gg = gt[gi];
  const double* __restrict__ cur = P->parity ? msg0 : msg1;
  __shared__ double sh[(MSTEP_THREADS / 32) * LMAXG * (QT > 0 ? QT * QT : MAXQ * MAXQ)];   // one slab per warp
  __shared__ unsigned char labS[MSTEP_THREADS];
  constexpr int WSL = LMAXG * (QT > 0 ? QT * QT : MAXQ * MAXQ);
  const int q = QT > 0 ? QT : qrt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, qq = q * q;
  for (int k = threadIdx.x; k < (MSTEP_THREADS / 32) * WSL; k += blockDim.x) sh[k] = 0.0;
  __syncthreads();
  double* pmrow = pm + ((size_t)gi * xcap + blockIdx.x) * G * qq;   // this block's partial sums (k_lsbm_reduce adds the blocks)
  if constexpr (QT > 0 && QT <= 8) {
    // q <= 8: the sum of outer products of a warp's 32 links is a (q x 32)(32 x q) product -> FP64 DMMA m8n8k4 over
    // groups of 4 links, one accumulator (two registers per lane) per label; operands of other labels are zeroed.
    constexpr int RS = 9;                                 // row stride (doubles): conflict-free for rows and fragments
    double* ra = dyn + (size_t)warp * 32 * 2 * RS;        // a / Z of the warp's 32 links
    double* rb = ra + 32 * RS;                            // the message of the other direction
    double c0[LMAXG], c1[LMAXG];
#pragma unroll
    for (int g = 0; g < LMAXG; ++g) c0[g] = c1[g] = 0.0;
    const int kq = lane & 3, cidx = lane >> 2;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long chunks = (gg.K + 31) / 32;
    for (long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < chunks; c += nwarps) {
      const long long k = c * 32 + lane;
      const long long e = gg.sbeg + k;
      unsigned char L = 255;
      if (k < gg.K && erow[e] < col[e]) {                 // row i < column j: every undirected link once
        L = lab[e];
        const double* A = P->aff[L];
        double mij[QT], mji[QT];
        load_vec<QT>(cur, rev[e], mij);                   // i -> j
        load_vec<QT>(cur, e, mji);                        // j -> i
        double Z = 0.0;
#pragma unroll
        for (int r = 0; r < QT; ++r) {
          double t = 0.0;
#pragma unroll
          for (int s2 = 0; s2 < QT; ++s2) t += mji[s2] * A[r * QT + s2];
          Z += mij[r] * t;
        }
        const double inv = 1.0 / Z;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          ra[lane * RS + r] = r < QT ? mij[r] * inv : 0.0;
          rb[lane * RS + r] = r < QT ? mji[r] : 0.0;
        }
      }
      labS[threadIdx.x] = L;
      __syncwarp();
#pragma unroll
      for (int grp = 0; grp < 8; ++grp) {
        const int row = 4 * grp + kq;
        const unsigned char Lr = labS[(warp << 5) + row];
        const double av = ra[row * RS + cidx], bv = rb[row * RS + cidx];
#pragma unroll
        for (int g = 0; g < LMAXG; ++g) {
          if (g < G) {   // warp-uniform
            const bool on = Lr == g;
            const double a = on ? av : 0.0, b = on ? bv : 0.0;
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[g]), "+d"(c1[g])
                         : "d"(a), "d"(b));
          }
        }
      }
      __syncwarp();
    }
    // accumulator layout: lane holds (row = lane / 4, columns 2 (lane % 4), + 1)
#pragma unroll
    for (int g = 0; g < LMAXG; ++g) {
      const int r = lane >> 2, s0 = 2 * (lane & 3);
      if (g < G && r < QT) {   // every entry of a warp's slab has one owner lane
        if (s0 < QT) sh[warp * WSL + g * qq + r * QT + s0] = c0[g];
        if (s0 + 1 < QT) sh[warp * WSL + g * qq + r * QT + s0 + 1] = c1[g];
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < G * qq; k += blockDim.x) {   // the warps in a fixed order
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < MSTEP_THREADS / 32; ++w) v += sh[w * WSL + k];
      pmrow[k] = v;
    }
    return;
  }
  double* rows = dyn + (size_t)warp * 32 * 2 * q;
  constexpr int EPL = ((QT > 0 ? QT * QT : MAXQ * MAXQ) + 31) / 32;         // entries (r,s) a lane may own
  double acc[LMAXG][EPL];
#pragma unroll
  for (int g = 0; g < LMAXG; ++g)
#pragma unroll
    for (int m = 0; m < EPL; ++m) acc[g][m] = 0.0;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long chunks = (gg.K + 31) / 32;
  for (long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < chunks; c += nwarps) {
    const long long k = c * 32 + lane;
    const long long e = gg.sbeg + k;
    unsigned char L = 255;
    if (k < gg.K && erow[e] < col[e]) {                 // row i < column j: every undirected link once
      L = lab[e];
      const double* A = P->aff[L];
      const double* mij = cur + (size_t)rev[e] * q;     // i -> j
      const double* mji = cur + (size_t)e * q;          // j -> i
      double Z = 0.0;
      for (int r = 0; r < q; ++r) {
        double t = 0.0;
        for (int s2 = 0; s2 < q; ++s2) t += mji[s2] * A[r * q + s2];
        Z += mij[r] * t;
      }
      const double inv = 1.0 / Z;
      for (int r = 0; r < q; ++r) {
        rows[(lane * 2) * q + r] = mij[r] * inv;
        rows[(lane * 2 + 1) * q + r] = mji[r];
      }
    }
    labS[threadIdx.x] = L;
    __syncwarp();
    for (int j = 0; j < 32; ++j) {
      const unsigned char Lj = labS[(warp << 5) + j];   // the same for every lane
      if (Lj == 255) continue;
      const double* aj = rows + (size_t)(j * 2) * q;
      const double* bj = aj + q;
#pragma unroll
      for (int m = 0; m < EPL; ++m) {
        const int kk = lane + 32 * m;
        if (kk < qq) {
          const double v = aj[kk / q] * bj[kk % q];
#pragma unroll
          for (int g = 0; g < LMAXG; ++g)
            if (Lj == g) acc[g][m] += v;
        }
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int g = 0; g < LMAXG; ++g)
#pragma unroll
    for (int m = 0; m < EPL; ++m) {
      const int kk = lane + 32 * m;
      if (g < G && kk < qq) sh[warp * WSL + g * qq + kk] = acc[g][m];   // one owner lane per entry of the warp's slab
    }
  __syncthreads();
  for (int k = threadIdx.x; k < G * qq; k += blockDim.x) {   // the warps in a fixed order
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < MSTEP_THREADS / 32; ++w) v += sh[w * WSL + k];
    pmrow[k] = v;
  }
}

// Per-block partial sums -> the graph's parameter block, in a fixed order (no floating-point atomics anywhere: a run is
// bit-reproducible, and a graph gives the same bits alone and inside a batch as long as the launch shapes agree).
// One block per graph; a warp per column, lanes strided over the blocks, then a shuffle tree.
//   what bit 0: residual (pv -> conv)   bit 1: sum log Z_i (pv -> sumlogZi)   bit 2: denom / sumPSI (pd)   bit 3: Gnew (pm)
__device__ __forceinline__ double l_column(const double* __restrict__ base, int nrows, size_t stride) {
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int r = lane; r < nrows; r += 32) s += base[(size_t)r * stride];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}
__global__ void __launch_bounds__(256) k_lsbm_reduce(const int* __restrict__ live, LParams* Pall, int q, int G, int xcap, int what,
                                                     const double* __restrict__ pv, int nv, const double* __restrict__ pd, int nd,
                                                     const double* __restrict__ pm, int nm, int it) {
  const int gi = l_graph(live);
  LParams* P = Pall + gi;
  if (l_skip(P, it)) return;
  // the graph's columns are dealt to the warps of its blocks (grid x: several blocks per graph when few graphs are live)
  const int lane = threadIdx.x & 31, nw = (blockDim.x >> 5) * gridDim.x;
  const int warp = (threadIdx.x >> 5) + (blockDim.x >> 5) * blockIdx.x;
  if ((what & 3) && warp == 0) {
    const double v = l_column(pv + (size_t)gi * xcap, nv, 1);
    if (lane == 0) {
      if (what & 1) P->conv = v;
      else P->sumlogZi = v;
    }
  }
  if (what & 4) {
    const int ncol = (LMAXG + 1) * q;
    for (int c = warp; c < ncol; c += nw) {
      const int a = c / q, t = c % q;
      if (a < LMAXG && a >= G) continue;
      const double v = l_column(pd + (size_t)gi * xcap * ncol + c, nd, (size_t)ncol);
      if (lane == 0) {
        if (a == LMAXG) P->sumPSI[t] = v;
        else P->denom[a][t] = v;
      }
    }
  }
  if (what & 8) {
    const int qq = q * q, ncol = G * qq;
    for (int c = warp; c < ncol; c += nw) {
      const double v = l_column(pm + (size_t)gi * xcap * ncol + c, nm, (size_t)ncol);
      if (lane == 0) P->Gnew[c / qq][c % qq] = v * P->aff[c / qq][c % qq];
    }
  }
}

// h[alpha, :] = sum_i degrees[i, alpha] (PSI * affinity[alpha])[i, :] = denom[alpha, :] * affinity[alpha]   (update_h)
__global__ void k_lsbm_pre(const int* __restrict__ live, LParams* Pall, int q, int G, int it) {
  const int tid = threadIdx.x;
  LParams* P = Pall + l_graph(live);
  if (l_skip(P, it)) return;
  if (tid == 0) P->conv = 0.0;   // the sweep that follows accumulates its residual here
  if (tid < G * q) {
    const int a = tid / q, t = tid % q;
    double s = 0.0;
    for (int r = 0; r < q; ++r) s += P->denom[a][r] * P->aff[a][r * q + t];
    P->h[a][t] = s;
  }
}

// gr = lr * update_gr(PSI) + (1 - lr) * gr; affinity = lr * update_affinity() + (1 - lr) * affinity   (EM main loop).
// In the notebook update_gr rebinds EM()'s `gr` before the blend reads it and update_affinity mutates and returns the
// same Dict, so both blends mix the new value with itself: the learning rate does not act (kept bit for bit).
// Also: the sweep of this iteration has run, so the graph's current messages are now in the other buffer.
__global__ void k_lsbm_post(const LGraph* __restrict__ gt, const int* __restrict__ live, LParams* Pall, int q, int G, double lr,
                            int prior, int learning, int red, int it, double thr) {
  const int tid = threadIdx.x;
  const int gi = l_graph(live);
  LParams* P = Pall + gi;
  if (l_skip(P, it)) return;
  const double N = gt[gi].N, K = (double)gt[gi].K;
  if (tid == 0) {
    P->parity ^= 1;
    if (it != 0 && P->done_at == 0 && P->conv < thr) P->done_at = it;   // acted on from iteration it + 1
  }
  if (prior && tid < q) {
    const double g = P->sumPSI[tid] / N;
    P->gr[tid] = lr * g + (1.0 - lr) * g;
    P->lgr[tid] = log(P->gr[tid]);
  }
  if (learning) {
    for (int k = tid; k < G * q * q; k += blockDim.x) {
      const int a = k / (q * q), r = (k % (q * q)) / q, s = k % q;
      double v = (P->Gnew[a][r * q + s] + P->Gnew[a][s * q + r]) / (P->denom[a][r] * P->denom[a][s]);
      if (v < 1e-8) v = 1e-8;                            // "underflow cut-off"
      if (red && a == G - 1) v = 1.0 / K;                // RED edges: fixed flat matrix
      if (!(v == v)) P->status |= 1;
      // Gnew / denom are read by every thread of this single block before anybody writes aff: aff has its own slot k
      P->aff[a][r * q + s] = lr * v + (1.0 - lr) * v;
    }
  }
}

// freeenergy(): per link with i > j.  PASS 0: sums; PASS 1: squared deviations from the means.
template <int PASS>
__global__ void __launch_bounds__(256) k_lsbm_fe(const LGraph* __restrict__ gt, int q, int G, const int* __restrict__ erow,
                                                 const int* __restrict__ col, const int* __restrict__ rev,
                                                 const unsigned char* __restrict__ lab, const double* __restrict__ deg,
                                                 const double* __restrict__ msg0, const double* __restrict__ msg1,
                                                 LParams* Pall, double* __restrict__ pf, int xcap) {
  __shared__ double red[256];
  LParams* P = Pall + blockIdx.y;
  const LGraph gg = gt[blockIdx.y];
  const double* __restrict__ cur = P->parity ? msg1 : msg0;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < gg.K; k += (long long)gridDim.x * blockDim.x) {
    const long long e = gg.sbeg + k;
    const int i = erow[e], j = col[e];
    if (i <= j) continue;
    const int a = lab[e];
    const double dd = deg[(size_t)i * G + a] * deg[(size_t)j * G + a];
    const double* A = P->aff[a];
    const double* mij = cur + (size_t)rev[e] * q;   // PSIcav[indij]
    const double* mji = cur + (size_t)e * q;        // PSIcav[indji]
    double Z = 0.0, gp = 0.0, gt_ = 0.0;
    int sa = 0, sb = 0;
    for (int r = 0; r < q; ++r) {
      if (mij[r] > mij[sa]) sa = r;
      if (mji[r] > mji[sb]) sb = r;
      for (int s = 0; s < q; ++s) {
        const double w = mij[r] * mji[s], x = dd * A[r * q + s], lx = log(x);
        Z += x * w;
        gp += w * lx;
        gt_ += w * x * lx;
      }
    }
    const double val[4] = {log(Z), gp, gt_ / Z, log(dd * A[sa * q + sb])};
    for (int kk = 0; kk < 4; ++kk) {
      if (PASS == 0) v[kk] += val[kk];
      else { const double d = val[kk] - P->femean[kk]; v[kk] += d * d; }
    }
  }
  for (int k = 0; k < 4; ++k) {
    const double s = block_sum(v[k], red);
    if (threadIdx.x == 0) pf[((size_t)blockIdx.y * xcap + blockIdx.x) * 4 + k] = s;
  }
}
// sums (pass 0) or sums of squared deviations (pass 1) of the four edge terms over the blocks, fixed order; 128 threads
__global__ void k_lsbm_fe_sum(const LGraph* __restrict__ gt, LParams* Pall, const double* __restrict__ pf, int nf, int xcap,
                              int pass) {
  LParams* P = Pall + blockIdx.y;
  const int k = threadIdx.x >> 5;
  if (k >= 4) return;
  const double v = gt[blockIdx.y].K > 0 ? l_column(pf + (size_t)blockIdx.y * xcap * 4 + k, nf, 4) : 0.0;
  if ((threadIdx.x & 31) == 0) {
    if (pass == 0) {
      P->fesum[k] = v;
      P->femean[k] = v / (0.5 * (double)gt[blockIdx.y].K);
    } else {
      P->fess[k] = v;
    }
  }
}
__global__ void k_lsbm_finalize(const LGraph* __restrict__ gt, LParams* Pall) {
  if (threadIdx.x != 0) return;
  LParams* P = Pall + blockIdx.y;
  const double N = gt[blockIdx.y].N, L = 0.5 * (double)gt[blockIdx.y].K;
  P->result[0] = -(P->sumlogZi - P->fesum[0]) / N - L / N;      // FE
  for (int k = 0; k < 4; ++k) {
    P->result[1 + k] = 1.0 - P->fesum[k] / L;                   // CVBayes, CVGP, CVGT, CVMAP
    P->result[5 + k] = P->fess[k] / (L - 1.0);                  // var(...)
  }
  for (int k = 0; k < 9; ++k)
    if (!(P->result[k] == P->result[k])) P->status |= 1;
}

// Union of a batch on the device: the links of graph y (sources, destinations in [k0, k0 + K) of the column-major K x 2
// array `uni`) are shifted by the vertices of the graphs before it; ids outside 1..n are flagged.
struct LSpan {
  long long k0, K;
  int v0, n;
};
__global__ void k_lsbm_shift(const LSpan* __restrict__ sp, long long Ktot, int64_t* __restrict__ uni, int* err) {
  const LSpan g = sp[blockIdx.y];
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < g.K; k += (long long)gridDim.x * blockDim.x) {
    const int64_t a = uni[g.k0 + k], b = uni[Ktot + g.k0 + k];
    if (a < 1 || a > g.n || b < 1 || b > g.n) atomicOr(err, 4);
    uni[g.k0 + k] = a + g.v0;
    uni[Ktot + g.k0 + k] = b + g.v0;
  }
}
// rand(1,B) per link, normalised (notebook EM, "initial state"), drawn on the device: row kk of graph y from
// (seed[y], kk), so a graph gets the same messages alone and inside a batch.
__device__ __forceinline__ uint64_t l_mix(uint64_t x) {   // splitmix64
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void k_lsbm_random(const LSpan* __restrict__ sp, const uint64_t* __restrict__ seed, int q,
                              const int* __restrict__ slot_of_link, double* __restrict__ msg) {
  const LSpan g = sp[blockIdx.y];
  const uint64_t sd = seed[blockIdx.y];
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < g.K; k += (long long)gridDim.x * blockDim.x) {
    double m[MAXQ], sum = 0.0;
    for (int t = 0; t < q; ++t) {
      const uint64_t r = l_mix(sd ^ l_mix((uint64_t)k * (uint64_t)q + (uint64_t)t));
      m[t] = ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
      sum += m[t];
    }
    double* row = msg + (size_t)slot_of_link[g.k0 + k] * q;
    for (int t = 0; t < q; ++t) row[t] = m[t] / sum;
  }
}

// rows between `links` order (K x q) and slot order
__global__ void k_lsbm_permute(const double* __restrict__ src, double* __restrict__ dst, const int* __restrict__ slot_of_link,
                               long long K, int q, int to_slots) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < K * q; k += (long long)gridDim.x * blockDim.x) {
    const long long row = k / q;
    const int t = (int)(k - row * q);
    const long long s = slot_of_link[row];
    if (to_slots) dst[s * q + t] = src[k];
    else dst[k] = src[s * q + t];
  }
}

}  // namespace

struct gbx_lsbm {
  gbx_sbm g;                 // the device-resident graph (rowptr, col, rev, erow, slot_of_link): the union of all graphs
  int q = 0, G = 0;
  int count = 1;             // graphs in this handle
  gbx_lsbm_options opt{};
  std::vector<LGraph> gt_host;
  LGraph* gt = nullptr;      // device copy of the graph table
  int* live = nullptr;       // device: graphs still running (EM of a batch), in ascending order
  std::vector<int> live_host;
  int nlive = 0;             // 0: every graph (live list not in use)
  int64_t nmax = 0, Kmax = 0;   // largest graph
  unsigned char* lab = nullptr;
  double *deg = nullptr, *dd = nullptr, *msg[2] = {nullptr, nullptr}, *PSI = nullptr;
  LParams* P = nullptr;      // [count]
  LParams* hP = nullptr;     // pinned mirror, [count]
  int* h_done = nullptr;     // pinned, [count]: done_at of every graph
  int xcap = 1;              // most blocks (grid x) any launch gives one graph: rows of the partial-sum arrays
  double *pv = nullptr, *pd = nullptr, *pm = nullptr, *pf = nullptr;   // per graph and block: residual / log Z, denominators,
                                                                       // M-step sums, free-energy terms (k_lsbm_reduce)
  int nbv = 0, nbd = 0, nbm = 0;   // blocks per graph of the last vertex / denominator / M-step launch
  LSpan* spans = nullptr;    // device, [count]: link range and vertex shift of every graph
  uint64_t* seeds = nullptr; // device, [count]
  bool initialised = false;
  int64_t launches = 0;
  int itr = 0;               // EM iterations queued since init
  int cur_it = 0;            // stamp of the iteration being queued (0: never skip)
  double cur_thr = -1.0;     // residual threshold k_lsbm_post tests (< 0: never converged)
};

namespace {

int l_sync_state(gbx_lsbm* h) {
  GBX_CUDA_TRY(cudaMemcpyAsync(h->hP, h->P, sizeof(LParams) * (size_t)h->count, cudaMemcpyDeviceToHost, h->g.stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->g.stream));
  return GBX_OK;
}

int l_read_conv(gbx_lsbm* h) {  // the residual of graph 0 alone: 8 bytes instead of the parameter blocks
  GBX_CUDA_TRY(cudaMemcpyAsync(&h->hP->conv, &h->P->conv, sizeof(double), cudaMemcpyDeviceToHost, h->g.stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->g.stream));
  return GBX_OK;
}

// grid of a kernel that walks `items` units per graph, `per_block` of them per block and round: enough blocks to cover
// the largest graph, but no more than ~8 blocks per SM over all the graphs of the launch
dim3 l_grid(const gbx_lsbm* h, int64_t items, int per_block, int ngraphs) {
  const int64_t want = std::max<int64_t>(1, (items + per_block - 1) / per_block);
  const int64_t cap = std::min<int64_t>(h->xcap, std::max<int64_t>(1, ((int64_t)h->g.num_sms * 8 + ngraphs - 1) / ngraphs));
  return dim3((unsigned)std::min(want, cap), (unsigned)ngraphs, 1);
}
int l_ngraphs(const gbx_lsbm* h) { return h->nlive > 0 ? h->nlive : h->count; }
// blocks per graph of k_lsbm_reduce: a warp per column of the partial sums, but no more blocks than the GPU holds
dim3 l_reduce_grid(const gbx_lsbm* h, int ngraphs) {
  const int ncol = 1 + (LMAXG + 1) * h->q + h->G * h->q * h->q;
  const int want = (ncol + 7) / 8;
  const int cap = std::max(1, h->g.num_sms * 8 / std::max(ngraphs, 1));
  return dim3((unsigned)std::max(1, std::min(want, cap)), (unsigned)ngraphs, 1);
}
const int* l_live(const gbx_lsbm* h) { return h->nlive > 0 ? h->live : nullptr; }

int l_denom(gbx_lsbm* h) {
  const int ng = l_ngraphs(h);
  const dim3 gd = l_grid(h, h->nmax, 256 / h->q, ng);
  k_lsbm_denom<<<gd, 256, 0, h->g.stream>>>(h->gt, l_live(h), h->q, h->G, h->deg, h->PSI, h->P, h->cur_it, h->pd, h->xcap);
  h->nbd = (int)gd.x;
  h->launches += 1;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

template <int MODE, int QT, int LPV>
int l_vertex_ql(gbx_lsbm* h) {
  const int ng = l_ngraphs(h);
  const dim3 gv = l_grid(h, h->nmax, 256 / LPV, ng);
  k_lsbm_vertex<MODE, QT, LPV><<<gv, 256, 0, h->g.stream>>>(
      h->gt, l_live(h), h->q, h->G, h->g.rowptr, h->g.rev, h->lab, h->deg, h->dd, h->msg[0], h->msg[1], h->PSI, h->P,
      h->opt.damping, h->cur_it, h->pv, h->xcap);
  h->nbv = (int)gv.x;
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}
template <int MODE, int QT>
int l_vertex_q(gbx_lsbm* h) {
  // A launch with enough vertices fills the GPU with one thread per vertex (no redundant normalisation across lanes);
  // small launches are latency bound and want the links of a vertex spread over lanes.
  const int64_t nlaunch = h->nlive > 0 ? (int64_t)h->nlive * h->nmax : h->g.n;
  static const int force_lpv = [] { const char* e = getenv("GBX_LSBM_LPV"); return e ? atoi(e) : 0; }();
  if (force_lpv == 8) return l_vertex_ql<MODE, QT, 8>(h);
  if (force_lpv == 16) return l_vertex_ql<MODE, QT, 16>(h);
  if (force_lpv == 1 || nlaunch >= 100000) return l_vertex_ql<MODE, QT, 1>(h);
  const double mean_degree = h->g.n > 0 ? (double)h->g.K / (double)h->g.n : 0.0;
  if (mean_degree <= 10.0) return l_vertex_ql<MODE, QT, 8>(h);
  if (mean_degree <= 24.0) return l_vertex_ql<MODE, QT, 16>(h);
  return l_vertex_ql<MODE, QT, 32>(h);
}
template <int MODE>
int l_vertex(gbx_lsbm* h) {
  switch (h->q) {
    case 2: return l_vertex_q<MODE, 2>(h);
    case 3: return l_vertex_q<MODE, 3>(h);
    case 4: return l_vertex_q<MODE, 4>(h);
    case 5: return l_vertex_q<MODE, 5>(h);
    case 6: return l_vertex_q<MODE, 6>(h);
    case 8: return l_vertex_q<MODE, 8>(h);
    default: return l_vertex_q<MODE, 0>(h);
  }
}

// one pass of the notebook's main loop over the live graphs: h = update_h(); BP(); gr and affinity updates.  denom
// (sum_i degrees PSI) of the current marginals is already there: from init or from the end of the previous pass.
int l_iteration(gbx_lsbm* h) {
  cudaStream_t st = h->g.stream;
  const int it = h->cur_it;
  const int ng = l_ngraphs(h);
  k_lsbm_pre<<<dim3(1, ng), 64, 0, st>>>(l_live(h), h->P, h->q, h->G, it);   // h = update_h(); residual := 0
  h->launches++;
  L_TRY(l_vertex<LMODE_SWEEP>(h));                              // (BPconv, PSI) = BP()
  L_TRY(l_denom(h));                                            // from the new marginals
  if (h->opt.learning && h->g.K > 0) {
    const dim3 mg = l_grid(h, h->Kmax, MSTEP_THREADS, ng);
    const size_t msm = (size_t)(MSTEP_THREADS / 32) * 32 * 2 * std::max(h->q, 9) * sizeof(double);
#define GBX_LSBM_MSTEP(QT)                                                                                                  \
  k_lsbm_mstep<QT><<<mg, MSTEP_THREADS, msm, st>>>(h->gt, l_live(h), h->q, h->G, h->g.erow, h->g.col, h->g.rev, h->lab, \
                                                   h->msg[0], h->msg[1], h->P, it, h->pm, h->xcap)
    switch (h->q) {
      case 2: GBX_LSBM_MSTEP(2); break;
      case 3: GBX_LSBM_MSTEP(3); break;
      case 4: GBX_LSBM_MSTEP(4); break;
      case 5: GBX_LSBM_MSTEP(5); break;
      case 6: GBX_LSBM_MSTEP(6); break;
      case 8: GBX_LSBM_MSTEP(8); break;
      default: GBX_LSBM_MSTEP(0); break;
    }
#undef GBX_LSBM_MSTEP
    h->nbm = (int)mg.x;
    h->launches++;
  }
  // the blocks' partial sums -> residual, denominators, M-step statistics of every live graph (fixed order)
  const bool mstep = h->opt.learning && h->g.K > 0;
  k_lsbm_reduce<<<l_reduce_grid(h, ng), 256, 0, st>>>(l_live(h), h->P, h->q, h->G, h->xcap, 1 | 4 | (mstep ? 8 : 0), h->pv, h->nbv,
                                                      h->pd, h->nbd, h->pm, mstep ? h->nbm : 0, it);
  h->launches++;
  k_lsbm_post<<<dim3(1, ng), 256, 0, st>>>(h->gt, l_live(h), h->P, h->q, h->G, h->opt.learningrate, h->opt.priorlearning,
                                           h->opt.learning, h->opt.REDfrac != 0.0, it, h->cur_thr);
  h->launches++;
  h->itr++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

int l_create(gbx_lsbm** out, int device, int count, const int64_t* n_vertices, const int64_t* n_links,
             const int64_t* const* links3, int q, int n_labels) {
  if (!out) return l_invalid("out is NULL");
  *out = nullptr;
  if (q < 1 || q > MAXQ) return l_invalid("q must be in 1..16");
  if (n_labels < 1 || n_labels > LMAXG) return l_invalid("number of labels must be in 1..GBX_MAX_LABELS");
  if (count < 1 || count > 65535 || !n_vertices || !n_links || !links3) return l_invalid("bad count / sizes / links");
  int64_t n = 0, K = 0;
  for (int g = 0; g < count; ++g) {
    if (n_vertices[g] < 1 || n_links[g] < 0 || (n_links[g] > 0 && !links3[g])) return l_invalid("bad n_vertices / n_links / links");
    n += n_vertices[g];
    K += n_links[g];
  }
  if (n >= (int64_t)1 << 31 || K >= (int64_t)1 << 31) return l_invalid("graph too large for 32-bit slots");
  const int ndev = gbx_device_count();
  if (ndev < 0) return ndev;
  if (ndev == 0) {
    last_error() = "no CUDA device: libgraphbix_cuda has no CPU fallback";
    return GBX_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return l_invalid("device out of range");
  gbx_lsbm* h = new (std::nothrow) gbx_lsbm();
  if (!h) return l_invalid("out of host memory");
  auto bail = [&](int rc) {
    const std::string keep = last_error();
    gbx_lsbm_destroy(h);
    last_error() = keep;
    return rc;
  };
  h->q = q;
  h->G = n_labels;
  h->count = count;
  h->g.device = device;
  h->g.q = q;
  h->g.n = h->g.n_global = n;
  h->g.K = h->g.K_global = K;
  h->g.rank = 0;
  h->g.world = 1;
  gbx_lsbm_default_options(&h->opt);
  int major = 0, sms = 0;
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (ce != cudaSuccess || major < 10) {
    last_error() = ce != cudaSuccess ? std::string("cudaSetDevice: ") + cudaGetErrorString(ce)
                                     : std::string("libgraphbix_cuda is built for sm_100a (B200) only");
    delete h;
    return GBX_ERR_CUDA;
  }
  h->g.num_sms = sms;
  // The union of the graphs: vertices of graph g are shifted by the vertices before it.  One graph: `links3` as it is
  // (the first two columns of the column-major K x 3 matrix are the K x 2 `links` the graph build expects); a batch is
  // assembled on the device (three contiguous copies per graph, one kernel for the shifts) and built from there.
  h->gt_host.resize((size_t)count);
  const int64_t* links_all = links3[0];
  const int64_t* labels_all = links3[0] ? links3[0] + 2 * n_links[0] : nullptr;
  int64_t* d_uni = nullptr;   // sources, destinations (column-major K x 2) and labels of the union
  std::vector<LSpan> spans((size_t)count);
  {
    int64_t v0 = 0, k0 = 0;
    for (int g = 0; g < count; ++g) {
      spans[(size_t)g] = LSpan{(long long)k0, (long long)n_links[g], (int)v0, (int)n_vertices[g]};
      v0 += n_vertices[g];
      k0 += n_links[g];
    }
  }
  if (count > 1 && K > 0) {
    cudaError_t e = dev_malloc(device, (void**)&d_uni, (size_t)3 * (size_t)K * sizeof(int64_t), true);
    LSpan* d_sp = nullptr;
    int* d_err = nullptr;
    int herr = 0;
    if (e == cudaSuccess) e = dev_malloc(device, (void**)&d_sp, sizeof(LSpan) * (size_t)count, true);
    if (e == cudaSuccess) e = dev_malloc(device, (void**)&d_err, sizeof(int), true);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_err, 0, sizeof(int), h->g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_sp, spans.data(), sizeof(LSpan) * (size_t)count, cudaMemcpyHostToDevice, h->g.stream);
    if (e == cudaSuccess) {   // three columns per graph; pageable sources are staged by host threads (gbx_pool.cu)
      std::vector<UploadSeg> segs;
      segs.reserve((size_t)3 * (size_t)count);
      for (int g = 0; g < count; ++g) {
        const int64_t Kg = n_links[g], k0 = spans[(size_t)g].k0;
        if (Kg == 0) continue;
        for (int c = 0; c < 3; ++c)
          segs.push_back(UploadSeg{d_uni + (size_t)c * K + k0, links3[g] + (size_t)c * Kg, (size_t)Kg * sizeof(int64_t)});
      }
      e = upload_segments_async(device, segs.data(), segs.size(), h->g.stream);
    }
    if (e == cudaSuccess) {
      int64_t Kmax = 0;
      for (int g = 0; g < count; ++g) Kmax = std::max(Kmax, n_links[g]);
      const int64_t want = std::max<int64_t>(1, (Kmax + 255) / 256), cap = std::max<int64_t>(1, ((int64_t)sms * 8 + count - 1) / count);
      k_lsbm_shift<<<dim3((unsigned)std::min(want, cap), (unsigned)count), 256, 0, h->g.stream>>>(d_sp, (long long)K, d_uni, d_err);
      h->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, h->g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->g.stream);
    dev_free(device, d_err);
    h->spans = d_sp;
    if (e != cudaSuccess || herr) {
      last_error() = e != cudaSuccess ? std::string("gbx_lsbm_create_batch: ") + cudaGetErrorString(e)
                                      : std::string("links: vertex id out of 1..n");
      dev_free(device, d_uni);
      gbx_lsbm_destroy(h);
      return e != cudaSuccess ? GBX_ERR_CUDA : GBX_ERR_INVALID;
    }
    links_all = d_uni;
    labels_all = d_uni + 2 * K;
  }
  std::vector<int> rowptr_g;
  int rc = build_graph_device(&h->g, links_all, &rowptr_g, nullptr);
  if (rc != GBX_OK) {
    dev_free(device, d_uni);
    return bail(rc);
  }
  {
    int64_t v0 = 0;
    for (int g = 0; g < count; ++g) {
      LGraph& t = h->gt_host[(size_t)g];
      t.vbeg = (int)v0;
      t.n = (int)n_vertices[g];
      t.sbeg = rowptr_g[(size_t)v0];
      t.K = (long long)rowptr_g[(size_t)(v0 + n_vertices[g])] - t.sbeg;
      t.N = (double)n_vertices[g];
      h->nmax = std::max<int64_t>(h->nmax, t.n);
      h->Kmax = std::max<int64_t>(h->Kmax, t.K);
      if (t.K != n_links[g]) {   // a link of graph g whose reverse is missing would have failed the build already
        last_error() = "links: a graph of the batch has links that do not come in both directions";
        return bail(GBX_ERR_INVALID);
      }
      v0 += n_vertices[g];
    }
  }
  auto body = [&]() -> int {
    const int dv = device;
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->lab, std::max<size_t>((size_t)K, 1), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->deg, std::max<size_t>((size_t)n * h->G, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->dd, std::max<size_t>((size_t)K, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->msg[0], std::max<size_t>((size_t)K * q, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->msg[1], std::max<size_t>((size_t)K * q, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->PSI, (size_t)n * q * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->P, sizeof(LParams) * (size_t)count, true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->gt, sizeof(LGraph) * (size_t)count, true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->live, sizeof(int) * (size_t)count, true));
    // partial sums per graph and block: a launch gives one graph at most xcap blocks
    // (at least 32, so that the last live graphs of a batch still get a fair share of the GPU, unless the M-step
    // partials of a batch with large q would exceed 256 MB)
    h->xcap = (int)std::max<int64_t>(32, ((int64_t)sms * 8 + count - 1) / count);
    while (h->xcap > 8 && (size_t)count * h->xcap * h->G * q * q * sizeof(double) > ((size_t)256 << 20)) h->xcap /= 2;
    const size_t rows = (size_t)count * (size_t)h->xcap;
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->pv, rows * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->pd, rows * (size_t)(LMAXG + 1) * q * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->pm, rows * (size_t)h->G * q * q * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->pf, rows * 4 * sizeof(double), true));
    GBX_CUDA_TRY(cudaMallocHost((void**)&h->hP, sizeof(LParams) * (size_t)count));
    GBX_CUDA_TRY(cudaMallocHost((void**)&h->h_done, sizeof(int) * (size_t)count));
    GBX_CUDA_TRY(cudaMemset(h->P, 0, sizeof(LParams) * (size_t)count));
    GBX_CUDA_TRY(cudaMemcpy(h->gt, h->gt_host.data(), sizeof(LGraph) * (size_t)count, cudaMemcpyHostToDevice));
    // labels of the links -> slots; both directions of a link must agree
    if (!h->spans) {
      GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->spans, sizeof(LSpan) * (size_t)count, true));
      GBX_CUDA_TRY(cudaMemcpy(h->spans, spans.data(), sizeof(LSpan) * (size_t)count, cudaMemcpyHostToDevice));
    }
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->seeds, sizeof(uint64_t) * (size_t)count, true));
    int64_t* d_lab = nullptr;
    int* err = nullptr;
    if (!d_uni) GBX_CUDA_TRY(dev_malloc(dv, (void**)&d_lab, std::max<size_t>((size_t)K, 1) * sizeof(int64_t), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&err, sizeof(int), true));
    int herr = 0;
    const int grid_e = (int)std::min<int64_t>(std::max<int64_t>(1, (K + 255) / 256), (int64_t)sms * 8);
    cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), h->g.stream);
    if (e == cudaSuccess && K > 0 && !d_uni)
      e = upload_async(dv, d_lab, labels_all, (size_t)K * sizeof(int64_t), h->g.stream);
    if (d_uni) {   // the labels of a batch are already on the device, behind the links
      d_lab = d_uni + 2 * K;
    }
    if (e == cudaSuccess && K > 0) {
      k_lab_slots<<<grid_e, 256, 0, h->g.stream>>>(d_lab, h->g.slot_of_link, (long long)K, h->G, h->lab, err);
      k_lab_check<<<grid_e, 256, 0, h->g.stream>>>(h->lab, h->g.rev, (long long)K, err);
      h->launches += 2;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, h->g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->g.stream);
    if (d_uni) dev_free(dv, d_uni);
    else dev_free(dv, d_lab);
    d_uni = nullptr;
    dev_free(dv, err);
    GBX_CUDA_TRY(e);
    if (herr & 1) return l_invalid("links: label out of 1..n_labels");
    if (herr & 2) return l_invalid("links: the two directions of a link carry different labels");
    return GBX_OK;
  };
  rc = body();
  if (rc != GBX_OK) {
    dev_free(device, d_uni);
    return bail(rc);
  }
  *out = h;
  return GBX_OK;
}

int l_not_ready(const gbx_lsbm* h) {
  if (!h) return l_invalid("handle is NULL");
  if (!h->initialised) {
    last_error() = "gbx_lsbm_init has not been called";
    return GBX_ERR_STATE;
  }
  return GBX_OK;
}

int l_init(gbx_lsbm* h, const double* gr, const double* affinity, const double* const* psicav, const uint64_t* seeds) {
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  const int q = h->q, G = h->G;
  cudaStream_t st = h->g.stream;
  memset(h->hP, 0, sizeof(LParams) * (size_t)h->count);
  for (int g = 0; g < h->count; ++g) {
    LParams* s = h->hP + g;
    for (int t = 0; t < q; ++t) {
      s->gr[t] = gr[(size_t)g * q + t];
      s->lgr[t] = std::log(s->gr[t]);
    }
    for (int a = 0; a < G; ++a)
      for (int k = 0; k < q * q; ++k) s->aff[a][k] = affinity[((size_t)g * G + a) * q * q + k];
  }
  GBX_CUDA_TRY(cudaMemcpyAsync(h->P, h->hP, sizeof(LParams) * (size_t)h->count, cudaMemcpyHostToDevice, st));
  h->itr = 0;
  h->nlive = 0;
  const int64_t K = h->g.K;
  const int sms = h->g.num_sms;
  const int grid_e = (int)std::min<int64_t>(std::max<int64_t>(1, (K + 255) / 256), (int64_t)sms * 8);
  const int grid_v = (int)std::min<int64_t>(std::max<int64_t>(1, (h->g.n + 255) / 256), (int64_t)sms * 8);
  if (K > 0 && !psicav) {
    GBX_CUDA_TRY(cudaMemcpyAsync(h->seeds, seeds, sizeof(uint64_t) * (size_t)h->count, cudaMemcpyHostToDevice, st));
    k_lsbm_random<<<l_grid(h, h->Kmax, 256, h->count), 256, 0, st>>>(h->spans, h->seeds, q, h->g.slot_of_link, h->msg[0]);
    h->launches++;
  } else if (K > 0) {
    // the other message buffer is scratch before the first sweep: the rows arrive there in `links` order, graph by graph
    int64_t k0 = 0;
    for (int g = 0; g < h->count; ++g) {
      const int64_t Kg = h->gt_host[(size_t)g].K;
      if (Kg > 0)
        GBX_CUDA_TRY(cudaMemcpyAsync(h->msg[1] + (size_t)k0 * q, psicav[g], (size_t)Kg * q * sizeof(double),
                                     cudaMemcpyHostToDevice, st));
      k0 += Kg;
    }
    k_lsbm_permute<<<grid_e, 256, 0, st>>>(h->msg[1], h->msg[0], h->g.slot_of_link, (long long)K, q, 1);
    h->launches++;
  }
  k_lsbm_degrees<<<grid_v, 256, 0, st>>>(h->g.rowptr, h->lab, (int)h->g.n, G, h->opt.dc, h->deg);
  if (K > 0)
    k_lsbm_dd<<<l_grid(h, h->Kmax, 256, h->count), 256, 0, st>>>(h->gt, G, h->g.erow, h->g.col, h->lab, h->deg, h->dd);
  h->launches += 2;
  GBX_CUDA_TRY(cudaGetLastError());
  // h = zeros(G,B); PSI = updatePSI(); h = update_h()   (notebook EM, "initial state")
  h->cur_it = 0;
  L_TRY(l_vertex<LMODE_MARGINAL>(h));
  L_TRY(l_denom(h));
  k_lsbm_reduce<<<dim3(1, h->count), 256, 0, st>>>(nullptr, h->P, q, G, h->xcap, 4, h->pv, 0, h->pd, h->nbd, h->pm, 0, 0);
  k_lsbm_pre<<<dim3(1, h->count), 64, 0, st>>>(nullptr, h->P, q, G, 0);
  h->launches += 2;
  GBX_CUDA_TRY(cudaGetLastError());
  GBX_CUDA_TRY(cudaStreamSynchronize(st));
  h->initialised = true;
  return GBX_OK;
}

// freeenergy() of every graph of the handle
int l_finalize(gbx_lsbm* h, gbx_lsbm_result* res) {
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  cudaStream_t st = h->g.stream;
  const int saved_live = h->nlive, saved_it = h->cur_it;
  h->nlive = 0;   // every graph
  h->cur_it = 0;
  const dim3 one(1, (unsigned)h->count);
  int rc = l_vertex<LMODE_LOGZ>(h);
  h->nlive = saved_live;
  h->cur_it = saved_it;
  L_TRY(rc);
  k_lsbm_reduce<<<one, 256, 0, st>>>(nullptr, h->P, h->q, h->G, h->xcap, 2, h->pv, h->nbv, h->pd, 0, h->pm, 0, 0);
  h->launches++;
  if (h->g.K > 0) {
    const dim3 ge = l_grid(h, h->Kmax, 256, h->count);
    k_lsbm_fe<0><<<ge, 256, 0, st>>>(h->gt, h->q, h->G, h->g.erow, h->g.col, h->g.rev, h->lab, h->deg, h->msg[0], h->msg[1], h->P,
                                     h->pf, h->xcap);
    k_lsbm_fe_sum<<<one, 128, 0, st>>>(h->gt, h->P, h->pf, (int)ge.x, h->xcap, 0);
    k_lsbm_fe<1><<<ge, 256, 0, st>>>(h->gt, h->q, h->G, h->g.erow, h->g.col, h->g.rev, h->lab, h->deg, h->msg[0], h->msg[1], h->P,
                                     h->pf, h->xcap);
    k_lsbm_fe_sum<<<one, 128, 0, st>>>(h->gt, h->P, h->pf, (int)ge.x, h->xcap, 1);
    h->launches += 4;
  }
  k_lsbm_finalize<<<one, 32, 0, st>>>(h->gt, h->P);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  L_TRY(l_sync_state(h));
  for (int g = 0; g < h->count; ++g) {
    gbx_lsbm_result* o = res + g;
    memset(o, 0, sizeof *o);
    const double* r = h->hP[g].result;
    o->FE = r[0];
    o->CVBayes = r[1];
    o->CVGP = r[2];
    o->CVGT = r[3];
    o->CVMAP = r[4];
    o->varCVBayes = r[5];
    o->varCVGP = r[6];
    o->varCVGT = r[7];
    o->varCVMAP = r[8];
    o->BPconv = h->hP[g].conv;
    o->nonfinite = h->hP[g].status & 1;
  }
  return GBX_OK;
}

// The main loop of the notebook's EM() over every graph of the handle.  Small graphs are launch bound, so `batch`
// iterations are queued before the host looks at anything: k_lsbm_post stamps the iteration at which a graph's residual
// fell below the threshold into its LParams::done_at and every kernel of a later iteration returns at once for that
// graph, which leaves the state the notebook's `break` leaves.  After every batch the host reads the stamps back (4
// bytes per graph) and rebuilds the list of live graphs the next launches walk.
int l_em(gbx_lsbm* h, gbx_lsbm_result* res) {
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  cudaStream_t st = h->g.stream;
  const int count = h->count;
  int batch = h->g.K < (int64_t)2000000 * count ? 8 : 1;
  if (const char* e = getenv("GBX_EM_BATCH")) batch = std::max(1, atoi(e));
  GBX_CUDA_TRY(cudaMemset2DAsync(&h->P->done_at, sizeof(LParams), 0, sizeof(int), (size_t)count, st));
  const int itr0 = h->itr;
  std::vector<int> itrnum((size_t)count, 0), cnv((size_t)count, 0);
  h->live_host.resize((size_t)count);
  for (int g = 0; g < count; ++g) h->live_host[(size_t)g] = g;
  int nlive = count, done = 0;
  h->cur_thr = h->opt.bp_conv_threshold;
  auto finish = [&](int rc) {
    h->cur_it = 0;
    h->cur_thr = -1.0;
    h->nlive = 0;
    return rc;
  };
  while (nlive > 0 && done < h->opt.itrmax) {
    if (count > 1) {
      GBX_CUDA_TRY(cudaMemcpyAsync(h->live, h->live_host.data(), sizeof(int) * (size_t)nlive, cudaMemcpyHostToDevice, st));
      h->nlive = nlive;
    }
    const int nb = std::min(batch, h->opt.itrmax - done);
    for (int b = 1; b <= nb; ++b) {
      h->cur_it = itr0 + done + b;
      const int rc = l_iteration(h);
      if (rc != GBX_OK) return finish(rc);
    }
    done += nb;
    GBX_CUDA_TRY(cudaMemcpy2DAsync(h->h_done, sizeof(int), &h->P->done_at, sizeof(LParams), sizeof(int), (size_t)count,
                                   cudaMemcpyDeviceToHost, st));
    cudaError_t ce = cudaStreamSynchronize(st);   // also orders the next upload of the live list after these launches
    if (ce != cudaSuccess) {
      last_error() = std::string("gbx_lsbm_em: ") + cudaGetErrorString(ce);
      return finish(GBX_ERR_CUDA);
    }
    int keep = 0;
    for (int k = 0; k < nlive; ++k) {
      const int g = h->live_host[(size_t)k];
      if (h->h_done[g] != 0) {
        cnv[(size_t)g] = 1;
        itrnum[(size_t)g] = h->h_done[g] - itr0;
      } else {
        h->live_host[(size_t)keep++] = g;
      }
    }
    nlive = keep;
  }
  h->itr = itr0 + done;
  finish(GBX_OK);
  L_TRY(l_finalize(h, res));
  for (int g = 0; g < count; ++g) {
    res[g].cnv = cnv[(size_t)g];
    res[g].itrnum = itrnum[(size_t)g];
    res[g].itrs_done = cnv[(size_t)g] ? itrnum[(size_t)g] : done;
  }
  return GBX_OK;
}

}  // namespace

extern "C" {

void gbx_lsbm_default_options(gbx_lsbm_options* opt) {
  if (!opt) return;
  opt->dc = 1;
  opt->learning = 1;
  opt->priorlearning = 1;
  opt->learningrate = 0.3;
  opt->itrmax = 512;
  opt->bp_conv_threshold = 0.000001;
  opt->REDfrac = 0.0;
  opt->damping = 1.0;
}

int gbx_lsbm_create(gbx_lsbm** out, int device, int64_t n, int64_t K, const int64_t* links3, int q, int n_labels) {
  return l_create(out, device, 1, &n, &K, &links3, q, n_labels);
}

int gbx_lsbm_create_batch(gbx_lsbm** out, int device, int count, const int64_t* n_vertices, const int64_t* n_links,
                          const int64_t* const* links3, int q, int n_labels) {
  return l_create(out, device, count, n_vertices, n_links, links3, q, n_labels);
}

int gbx_lsbm_count(const gbx_lsbm* h) { return h ? h->count : 0; }

int gbx_lsbm_destroy(gbx_lsbm* h) {
  if (!h) return GBX_OK;
  const int dv = h->g.device;
  cudaSetDevice(dv);
  cudaStreamSynchronize(h->g.stream);
  dev_free(dv, h->g.rowptr);
  dev_free(dv, h->g.col);
  dev_free(dv, h->g.rev);
  dev_free(dv, h->g.erow);
  dev_free(dv, h->g.slot_of_link);
  dev_free(dv, h->lab);
  dev_free(dv, h->deg);
  dev_free(dv, h->dd);
  dev_free(dv, h->msg[0]);
  dev_free(dv, h->msg[1]);
  dev_free(dv, h->PSI);
  dev_free(dv, h->P);
  dev_free(dv, h->gt);
  dev_free(dv, h->live);
  dev_free(dv, h->spans);
  dev_free(dv, h->seeds);
  dev_free(dv, h->pv);
  dev_free(dv, h->pd);
  dev_free(dv, h->pm);
  dev_free(dv, h->pf);
  if (h->hP) cudaFreeHost(h->hP);
  if (h->h_done) cudaFreeHost(h->h_done);
  delete h;
  return GBX_OK;
}

int gbx_lsbm_set_stream(gbx_lsbm* h, void* cuda_stream) {
  if (!h) return l_invalid("handle is NULL");
  h->g.stream = (cudaStream_t)cuda_stream;
  return GBX_OK;
}

int gbx_lsbm_set_options(gbx_lsbm* h, const gbx_lsbm_options* opt) {
  if (!h || !opt) return l_invalid("handle/options is NULL");
  if (!(opt->damping > 0.0 && opt->damping <= 1.0)) return l_invalid("damping must be in (0, 1]");
  if (!(opt->learningrate >= 0.0 && opt->learningrate <= 1.0)) return l_invalid("learningrate must be in [0, 1]");
  if (opt->itrmax < 0) return l_invalid("itrmax must be >= 0");
  h->opt = *opt;
  return GBX_OK;
}

int gbx_lsbm_init(gbx_lsbm* h, const double* gr, const double* affinity, const double* psicav) {
  if (!h || !gr || !affinity || (!psicav && h->g.K > 0)) return l_invalid("handle/gr/affinity/psicav is NULL");
  if (h->count != 1) return l_invalid("this handle holds a batch: use gbx_lsbm_init_batch");
  return l_init(h, gr, affinity, &psicav, nullptr);
}

int gbx_lsbm_init_batch(gbx_lsbm* h, const double* gr, const double* affinity, const double* const* psicav) {
  if (!h || !gr || !affinity || !psicav) return l_invalid("handle/gr/affinity/psicav is NULL");
  for (int g = 0; g < h->count; ++g)
    if (!psicav[g] && h->gt_host[(size_t)g].K > 0) return l_invalid("psicav of a graph is NULL");
  return l_init(h, gr, affinity, psicav, nullptr);
}

int gbx_lsbm_init_random(gbx_lsbm* h, const double* gr, const double* affinity, const uint64_t* seeds) {
  if (!h || !gr || !affinity || !seeds) return l_invalid("handle/gr/affinity/seeds is NULL");
  return l_init(h, gr, affinity, nullptr, seeds);
}

int gbx_lsbm_iterate(gbx_lsbm* h, int n_iters, double* bpconv_out) {
  L_TRY(l_not_ready(h));
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  for (int i = 0; i < n_iters; ++i) L_TRY(l_iteration(h));
  if (bpconv_out) {
    L_TRY(l_read_conv(h));
    *bpconv_out = h->hP->conv;
  }
  return GBX_OK;
}

int gbx_lsbm_finalize(gbx_lsbm* h, gbx_lsbm_result* res) {
  if (!res) return l_invalid("handle/result is NULL");
  L_TRY(l_not_ready(h));
  return l_finalize(h, res);
}

int gbx_lsbm_em(gbx_lsbm* h, gbx_lsbm_result* res) {
  if (!res) return l_invalid("handle/result is NULL");
  L_TRY(l_not_ready(h));
  return l_em(h, res);
}

// EM() of `count` initialised handles of ONE device, side by side: every handle runs on a host thread-free round robin --
// a batch of iterations of each is queued on a stream of its own before any is waited for.  Kept for handles that
// already exist; gbx_lsbm_create_batch + gbx_lsbm_em (one launch set for all graphs) is the faster way to run many graphs.
int gbx_lsbm_em_batch(gbx_lsbm** hs, int count, gbx_lsbm_result* res, int max_concurrent) {
  if (count < 0 || (count > 0 && (!hs || !res))) return l_invalid("handles/results is NULL");
  (void)max_concurrent;
  int off = 0;
  for (int i = 0; i < count; ++i) {
    if (!hs[i]) return l_invalid("a handle is NULL");
    L_TRY(l_not_ready(hs[i]));
  }
  for (int i = 0; i < count; ++i) {
    L_TRY(l_em(hs[i], res + off));
    off += hs[i]->count;
  }
  return GBX_OK;
}

int gbx_lsbm_get_state_of(gbx_lsbm* h, int g, double* gr, double* affinity, double* PSI, double* hfield) {
  if (!h) return l_invalid("handle is NULL");
  if (g < 0 || g >= h->count) return l_invalid("graph index out of range");
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  const int q = h->q, G = h->G;
  GBX_CUDA_TRY(cudaMemcpyAsync(h->hP + g, h->P + g, sizeof(LParams), cudaMemcpyDeviceToHost, h->g.stream));
  if (PSI)
    GBX_CUDA_TRY(cudaMemcpyAsync(PSI, h->PSI + (size_t)h->gt_host[(size_t)g].vbeg * q,
                                 (size_t)h->gt_host[(size_t)g].n * q * sizeof(double), cudaMemcpyDeviceToHost, h->g.stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->g.stream));
  const LParams* s = h->hP + g;
  if (gr) memcpy(gr, s->gr, q * sizeof(double));
  if (affinity)
    for (int a = 0; a < G; ++a) memcpy(affinity + (size_t)a * q * q, s->aff[a], (size_t)q * q * sizeof(double));
  if (hfield)
    for (int a = 0; a < G; ++a) memcpy(hfield + (size_t)a * q, s->h[a], q * sizeof(double));
  return GBX_OK;
}

int gbx_lsbm_get_state(gbx_lsbm* h, double* gr, double* affinity, double* PSI, double* hfield) {
  return gbx_lsbm_get_state_of(h, 0, gr, affinity, PSI, hfield);
}

// current messages of graph g, rows in the order of its `links3`
int gbx_lsbm_get_messages_of(gbx_lsbm* h, int g, double* psicav) {
  if (!h || !psicav) return l_invalid("handle/psicav is NULL");
  if (g < 0 || g >= h->count) return l_invalid("graph index out of range");
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  const int64_t K = h->g.K;
  const LGraph& t = h->gt_host[(size_t)g];
  if (t.K == 0) return GBX_OK;
  cudaStream_t st = h->g.stream;
  GBX_CUDA_TRY(cudaMemcpyAsync(&h->hP[g].parity, &h->P[g].parity, sizeof(int), cudaMemcpyDeviceToHost, st));
  GBX_CUDA_TRY(cudaStreamSynchronize(st));
  const int par = h->hP[g].parity;
  // Between sweeps the other buffer is scratch for graphs whose parity this is -- but graphs of a batch may differ in
  // parity, so the rows go through a temporary of their own.
  double* tmp = nullptr;
  GBX_CUDA_TRY(dev_malloc(h->g.device, (void**)&tmp, (size_t)K * h->q * sizeof(double), true));
  const int grid_e = (int)std::min<int64_t>(std::max<int64_t>(1, (K + 255) / 256), (int64_t)h->g.num_sms * 8);
  k_lsbm_permute<<<grid_e, 256, 0, st>>>(h->msg[par], tmp, h->g.slot_of_link, (long long)K, h->q, 0);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  int64_t k0 = 0;   // links of graph g start after the links of the graphs before it
  for (int k = 0; k < g; ++k) k0 += h->gt_host[(size_t)k].K;
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(psicav, tmp + (size_t)k0 * h->q, (size_t)t.K * h->q * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dev_free(h->g.device, tmp);
  GBX_CUDA_TRY(e);
  return GBX_OK;
}

int gbx_lsbm_get_messages(gbx_lsbm* h, double* psicav) { return gbx_lsbm_get_messages_of(h, 0, psicav); }

int64_t gbx_lsbm_launch_count(const gbx_lsbm* h) { return h ? h->launches + h->g.launches : 0; }

}  // extern "C"
