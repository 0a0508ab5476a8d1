// Labeled stochastic block model (edge labels alpha = 1..G, one affinity matrix and one degree sequence per label):
// EM + BP of /root/reference/src/labeled_sbm/labeled-sbmBIX_0.2.3.8_Julia0.6.2.ipynb, cell 0, `function EM`.
//
// FIRST CUDA PATH of this model (DESIGN.md row (f')): correct and parity-tested, not yet tuned.  It works in the log
// domain exactly as the notebook does (one log per message component per update), one thread per vertex, runtime q;
// the tuned machinery of gbx_sbm_q.cu (tiles, cp.async staging, linear-domain products, fused M-step) is the next step.
// Synchronous (Jacobi) sweeps like the other models: same fixed point as the notebook's asynchronous sweep.
//
//   k_lsbm_vertex<MODE>   logunnormalizedMessage / BP / updatePSI / sum log Z_i      (notebook: same names)
//   k_lsbm_denom          denom[alpha,:] = sum_i degrees[i,alpha] PSI[i,:], sum PSI   (update_affinity, update_h, update_gr)
//   k_lsbm_mstep          affinitynew[alpha] += PSIij / sum(PSIij) over i < j          (update_affinity)
//   k_lsbm_pre / k_lsbm_post   h = update_h(); gr and affinity updates with the learning rate
//   k_lsbm_fe<PASS>       log Z^{ij}, Gibbs / MAP errors per edge                     (freeenergy)
#include <algorithm>
#include <cstring>
#include <vector>

#include "gbx_sbm_internal.h"

namespace gbx {
int build_graph_device(gbx_sbm* h, const int64_t* links_host, std::vector<int>* rowptr_host, GlobalGraph* keep_global);
}
using namespace gbx;

namespace {

constexpr int LMAXG = GBX_MAX_LABELS;
constexpr double LOG1EM8 = -18.420680743952367;  // log(10^-8)
enum { LMODE_SWEEP = 0, LMODE_MARGINAL = 1, LMODE_LOGZ = 2 };

struct LParams {
  double gr[MAXQ];
  double aff[LMAXG][MAXQ * MAXQ];   // row-major B x B per label
  double h[LMAXG][MAXQ];
  double denom[LMAXG][MAXQ];
  double sumPSI[MAXQ];
  double Gnew[LMAXG][MAXQ * MAXQ];
  double conv, sumlogZi;
  double fesum[4], femean[4], fess[4];
  double result[9];
  int status;                        // bit 0: non-finite value seen; bit 1: the two directions of a link differ in label
  int done_at;                       // EM iteration whose residual fell below the threshold (0: not yet); kernels stamped
                                     // with a later iteration return at once (the host queues several iterations)
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
// normalize_logprob (notebook): clamp at log(1e-8), then exp(x - logsumexp(x)) with logsumexp's 500 cut-off.
__device__ __forceinline__ void normalize_logprob(const double* x, int q, double* p) {
  double m = -INFINITY;
  for (int t = 0; t < q; ++t) {
    const double y = x[t] < LOG1EM8 ? LOG1EM8 : x[t];
    p[t] = y;
    m = fmax(m, y);
  }
  double s = 0.0;
  for (int t = 0; t < q; ++t) {
    const double y = (m - p[t] > 500.0) ? m - 500.0 : p[t];
    p[t] = exp(y - m);
    s += p[t];
  }
  const double inv = 1.0 / s;
  for (int t = 0; t < q; ++t) p[t] *= inv;
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
__global__ void k_lsbm_dd(long long K, int G, const int* __restrict__ erow, const int* __restrict__ col,
                          const unsigned char* __restrict__ lab, const double* __restrict__ deg, double N,
                          double* __restrict__ dd) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const int a = lab[e];
    dd[e] = deg[(size_t)erow[e] * G + a] * deg[(size_t)col[e] * G + a] * N;
  }
}

// ---- the sweep ----------------------------------------------------------------------------------------------------
// LPV lanes per vertex (1 for large graphs, else 8, 16 or 32 from the mean degree), one lane per incoming link (rounds of LPV for
// larger degrees): the lanes read the vertex's contiguous run of messages coalesced, each computes the log term of its
// link once, the terms are summed over the group, and each lane then emits the cavity message of its own link.
// QT > 0: q known at compile time (loops unrolled, per-lane vectors in registers); QT == 0: any q <= 16.
template <int MODE, int QT, int LPV>
__global__ void __launch_bounds__(256) k_lsbm_vertex(int n, int qrt, int G, const int* __restrict__ rowptr,
                                                     const int* __restrict__ col, const int* __restrict__ rev,
                                                     const unsigned char* __restrict__ lab, const double* __restrict__ deg,
                                                     const double* __restrict__ ddN, const double* __restrict__ cur,
                                                     double* __restrict__ nxt,
                                                     double* __restrict__ PSI, LParams* P, double N, double damping,
                                                     int it) {
  __shared__ double red[256];
  if (MODE == LMODE_SWEEP && l_skip(P, it)) return;
  const int q = QT > 0 ? QT : qrt;
  constexpr int QA = QT > 0 ? QT : MAXQ;
  constexpr int GPB = 256 / LPV;                        // vertices per block and round
  const int lane = threadIdx.x % LPV, grp = threadIdx.x / LPV;
  double acc = 0.0;  // residual (sweep) or sum of log Z_i, kept by lane 0 of every group
  for (int i0 = blockIdx.x * GPB; i0 < n; i0 += gridDim.x * GPB) {   // block-uniform trip count
    const int i = i0 + grp;
    const bool valid = i < n;
    const int e0 = valid ? rowptr[i] : 0, e1 = valid ? rowptr[i + 1] : 0;
    const bool single = (e1 - e0) <= LPV;   // every lane keeps the term of its one link in registers
    double x[QA], term[QA];
#pragma unroll
    for (int t = 0; t < QA; ++t) x[t] = 0.0;
    for (int base = e0; base < e1; base += LPV) {
      const int e = base + lane;
      if (e < e1) {
        const int a = lab[e];
        const double dd = ddN[e];
        double psi[QA];
        row_load<QT, QA>(cur, e, q, psi);
        const double* A = P->aff[a];
#pragma unroll
        for (int t = 0; t < QA; ++t) {
          if (t < q) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < QA; ++r)
              if (r < q) s += psi[r] * A[r * q + t];
            term[t] = log(dd * s);
            x[t] += term[t];
          }
        }
      }
    }
    __syncwarp();   // groups of one warp may have walked different numbers of rounds
#pragma unroll
    for (int t = 0; t < QA; ++t) {
      if (t < q) {
        double v = x[t];
#pragma unroll
        for (int o = LPV / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        double ext = 0.0;
        if (valid)
          for (int a = 0; a < G; ++a) ext += deg[(size_t)i * G + a] * P->h[a][t];
        x[t] = v + (log(P->gr[t]) - ext);   // prior and field added after the messages, as logunnormalizedMessage does
      }
    }
    if (MODE == LMODE_LOGZ) {
      if (valid && lane == 0) acc += logsumexp_dev(x, q);
      continue;
    }
    double w[QA];
    normalize_logprob(x, q, w);
    if (valid && lane == 0) {
      for (int t = 0; t < q; ++t) {
        if (MODE == LMODE_SWEEP) acc += fabs(w[t] - PSI[(size_t)i * q + t]) / N;
        if (!(w[t] == w[t])) P->status |= 1;
      }
    }
    __syncwarp();   // lane 0 has read the old marginal
    if (MODE == LMODE_SWEEP) {
      for (int base = e0; base < e1; base += LPV) {
        const int e = base + lane;
        if (e < e1) {
          if (!single) {
            const int a = lab[e];
            const double dd = ddN[e];
            double psi[QA];
            row_load<QT, QA>(cur, e, q, psi);
            const double* A = P->aff[a];
#pragma unroll
            for (int t = 0; t < QA; ++t) {
              if (t < q) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < QA; ++r)
                  if (r < q) s += psi[r] * A[r * q + t];
                term[t] = log(dd * s);
              }
            }
          }
          double cav[QA], m[QA];
#pragma unroll
          for (int t = 0; t < QA; ++t)
            if (t < q) cav[t] = x[t] - term[t];
          normalize_logprob(cav, q, m);
          const long long rv = rev[e];
          if (damping != 1.0) {
            double old[QA];
            row_load<QT, QA>(cur, rv, q, old);
            for (int t = 0; t < q; ++t) m[t] = (1.0 - damping) * old[t] + damping * m[t];
          }
          row_store<QT, QA>(nxt, rv, q, m);
        }
      }
    }
    if (valid && lane == 0)
      for (int t = 0; t < q; ++t) PSI[(size_t)i * q + t] = w[t];
  }
  if (MODE != LMODE_MARGINAL) {
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) atomicAdd(MODE == LMODE_SWEEP ? &P->conv : &P->sumlogZi, s);
  }
}

// denom[alpha, t] = sum_i degrees[i, alpha] PSI[i, t] and sum_i PSI[i, t]
__global__ void __launch_bounds__(256) k_lsbm_denom(int n, int q, int G, const double* __restrict__ deg,
                                                    const double* __restrict__ PSI, LParams* P, int it) {
  __shared__ double sh[(LMAXG + 1) * MAXQ];
  if (l_skip(P, it)) return;
  for (int k = threadIdx.x; k < (LMAXG + 1) * MAXQ; k += blockDim.x) sh[k] = 0.0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    for (int t0 = 0; t0 < q; ++t0) {
      const int t = (t0 + (int)threadIdx.x) % q;   // rotated: the lanes of a warp add to different addresses
      const double p = PSI[(size_t)i * q + t];
      atomicAdd(&sh[LMAXG * MAXQ + t], p);
      for (int a = 0; a < G; ++a) atomicAdd(&sh[a * MAXQ + t], deg[(size_t)i * G + a] * p);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < (LMAXG + 1) * MAXQ; k += blockDim.x) {
    const int a = k / MAXQ, t = k % MAXQ;
    if (t >= q || sh[k] == 0.0) continue;
    if (a < LMAXG) {
      if (a < G) atomicAdd(&P->denom[a][t], sh[k]);
    } else {
      atomicAdd(&P->sumPSI[t], sh[k]);
    }
  }
}

// affinitynew[alpha] += PSIij / sum(PSIij), once per undirected link (i < j); the factor degrees * degrees * Ntot of
// the notebook's PSIij cancels in the normalisation, and affinity[alpha][r,s] is the same for every link of a label:
//     affinitynew[alpha][r,s] = affinity[alpha][r,s] * sum_links (psi_{i->j}[r] / Z_link) psi_{j->i}[s],
// a sum of outer products.  A warp takes 32 slots at a time: every lane normalises its link and parks the two vectors
// in shared memory, then every lane accumulates the (r,s) entries it owns over the 32 links in registers.
constexpr int MSTEP_THREADS = 128;
__global__ void __launch_bounds__(MSTEP_THREADS) k_lsbm_mstep(long long K, int q, int G, const int* __restrict__ erow,
                                                              const int* __restrict__ col, const int* __restrict__ rev,
                                                              const unsigned char* __restrict__ lab,
                                                              const double* __restrict__ cur, LParams* P, int it) {
  extern __shared__ double dyn[];
  if (l_skip(P, it)) return;                       // [warps][32][2 q]
  __shared__ double sh[LMAXG * MAXQ * MAXQ];
  __shared__ unsigned char labS[MSTEP_THREADS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, qq = q * q;
  double* rows = dyn + (size_t)warp * 32 * 2 * q;
  for (int k = threadIdx.x; k < G * qq; k += blockDim.x) sh[k] = 0.0;
  __syncthreads();
  constexpr int EPL = (MAXQ * MAXQ + 31) / 32;         // entries (r,s) a lane may own
  double acc[LMAXG][EPL];
#pragma unroll
  for (int g = 0; g < LMAXG; ++g)
#pragma unroll
    for (int m = 0; m < EPL; ++m) acc[g][m] = 0.0;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long chunks = (K + 31) / 32;
  for (long long c = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < chunks; c += nwarps) {
    const long long e = c * 32 + lane;
    unsigned char L = 255;
    if (e < K && erow[e] < col[e]) {                    // row i < column j: every undirected link once
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
        const int k = lane + 32 * m;
        if (k < qq) {
          const double v = aj[k / q] * bj[k % q];
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
      const int k = lane + 32 * m;
      if (g < G && k < qq && acc[g][m] != 0.0) atomicAdd(&sh[g * qq + k], acc[g][m]);
    }
  __syncthreads();
  for (int k = threadIdx.x; k < G * qq; k += blockDim.x) {
    const int g = k / qq, rs = k % qq;
    if (sh[k] != 0.0) atomicAdd(&P->Gnew[g][rs], sh[k] * P->aff[g][rs]);
  }
}

__global__ void k_lsbm_zero(LParams* P, int what, int it) {  // what bit 0: denom / sumPSI / Gnew, bit 1: conv, bit 2: free energy sums
  const int tid = threadIdx.x;
  if (l_skip(P, it)) return;
  if (what & 1) {
    for (int k = tid; k < LMAXG * MAXQ; k += blockDim.x) P->denom[k / MAXQ][k % MAXQ] = 0.0;
    for (int k = tid; k < MAXQ; k += blockDim.x) P->sumPSI[k] = 0.0;
    for (int k = tid; k < LMAXG * MAXQ * MAXQ; k += blockDim.x) P->Gnew[k / (MAXQ * MAXQ)][k % (MAXQ * MAXQ)] = 0.0;
  }
  if ((what & 2) && tid == 0) P->conv = 0.0;
  if ((what & 4) && tid == 0) {
    P->sumlogZi = 0.0;
    for (int k = 0; k < 4; ++k) P->fesum[k] = P->femean[k] = P->fess[k] = 0.0;
  }
}

// h[alpha, :] = sum_i degrees[i, alpha] (PSI * affinity[alpha])[i, :] = denom[alpha, :] * affinity[alpha]   (update_h)
__global__ void k_lsbm_pre(LParams* P, int q, int G, int it) {
  const int tid = threadIdx.x;
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
__global__ void k_lsbm_post(LParams* P, int q, int G, double N, double K, double lr, int prior, int learning, int red,
                            int it, double thr) {
  const int tid = threadIdx.x;
  if (l_skip(P, it)) return;
  if (tid == 0 && it != 0 && P->done_at == 0 && P->conv < thr) P->done_at = it;   // acted on from iteration it + 1
  if (prior && tid < q) {
    const double g = P->sumPSI[tid] / N;
    P->gr[tid] = lr * g + (1.0 - lr) * g;
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
__global__ void __launch_bounds__(256) k_lsbm_fe(long long K, int q, int G, const int* __restrict__ erow,
                                                 const int* __restrict__ col, const int* __restrict__ rev,
                                                 const unsigned char* __restrict__ lab, const double* __restrict__ deg,
                                                 const double* __restrict__ cur, LParams* P) {
  __shared__ double red[256];
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const int i = erow[e], j = col[e];
    if (i <= j) continue;
    const int a = lab[e];
    const double dd = deg[(size_t)i * G + a] * deg[(size_t)j * G + a];
    const double* A = P->aff[a];
    const double* mij = cur + (size_t)rev[e] * q;   // PSIcav[indij]
    const double* mji = cur + (size_t)e * q;        // PSIcav[indji]
    double Z = 0.0, gp = 0.0, gt = 0.0;
    int sa = 0, sb = 0;
    for (int r = 0; r < q; ++r) {
      if (mij[r] > mij[sa]) sa = r;
      if (mji[r] > mji[sb]) sb = r;
      for (int s = 0; s < q; ++s) {
        const double w = mij[r] * mji[s], x = dd * A[r * q + s], lx = log(x);
        Z += x * w;
        gp += w * lx;
        gt += w * x * lx;
      }
    }
    const double val[4] = {log(Z), gp, gt / Z, log(dd * A[sa * q + sb])};
    for (int k = 0; k < 4; ++k) {
      if (PASS == 0) v[k] += val[k];
      else { const double d = val[k] - P->femean[k]; v[k] += d * d; }
    }
  }
  for (int k = 0; k < 4; ++k) {
    const double s = block_sum(v[k], red);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(PASS == 0 ? &P->fesum[k] : &P->fess[k], s);
  }
}
__global__ void k_lsbm_fe_mean(LParams* P, double L) {
  if (threadIdx.x < 4) P->femean[threadIdx.x] = P->fesum[threadIdx.x] / L;
}
__global__ void k_lsbm_finalize(LParams* P, double N, double L) {
  if (threadIdx.x != 0) return;
  P->result[0] = -(P->sumlogZi - P->fesum[0]) / N - L / N;      // FE
  for (int k = 0; k < 4; ++k) {
    P->result[1 + k] = 1.0 - P->fesum[k] / L;                   // CVBayes, CVGP, CVGT, CVMAP
    P->result[5 + k] = P->fess[k] / (L - 1.0);                  // var(...)
  }
  for (int k = 0; k < 9; ++k)
    if (!(P->result[k] == P->result[k])) P->status |= 1;
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
  gbx_sbm g;                 // the device-resident graph (rowptr, col, rev, erow, slot_of_link)
  int q = 0, G = 0;
  gbx_lsbm_options opt{};
  unsigned char* lab = nullptr;
  double *deg = nullptr, *dd = nullptr, *msg[2] = {nullptr, nullptr}, *PSI = nullptr;
  LParams* P = nullptr;
  LParams* hP = nullptr;     // pinned mirror
  int cur = 0;
  bool initialised = false;
  int64_t launches = 0;
  int grid_v = 1, grid_e = 1, grid_w = 1;
  int itr = 0;               // EM iterations run since init
  int cur_it = 0;            // stamp of the iteration being queued (0: never skip)
  double cur_thr = -1.0;     // residual threshold k_lsbm_post tests (< 0: never converged)
};

namespace {

int l_sync_state(gbx_lsbm* h) {
  GBX_CUDA_TRY(cudaMemcpyAsync(h->hP, h->P, sizeof(LParams), cudaMemcpyDeviceToHost, h->g.stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->g.stream));
  return GBX_OK;
}

int l_read_conv(gbx_lsbm* h) {  // the residual alone: 8 bytes instead of the whole parameter block
  GBX_CUDA_TRY(cudaMemcpyAsync(&h->hP->conv, &h->P->conv, sizeof(double), cudaMemcpyDeviceToHost, h->g.stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->g.stream));
  return GBX_OK;
}

int l_denom(gbx_lsbm* h) {
  k_lsbm_zero<<<1, 256, 0, h->g.stream>>>(h->P, 1, h->cur_it);
  k_lsbm_denom<<<h->grid_v, 256, 0, h->g.stream>>>((int)h->g.n, h->q, h->G, h->deg, h->PSI, h->P, h->cur_it);
  h->launches += 2;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}

template <int MODE, int QT, int LPV>
int l_vertex_ql(gbx_lsbm* h) {
  const int gpb = 256 / LPV;
  const int grid = (int)std::min<int64_t>(std::max<int64_t>(1, (h->g.n + gpb - 1) / gpb), (int64_t)h->g.num_sms * 8);
  k_lsbm_vertex<MODE, QT, LPV><<<grid, 256, 0, h->g.stream>>>((int)h->g.n, h->q, h->G, h->g.rowptr, h->g.col, h->g.rev,
                                                              h->lab, h->deg, h->dd, h->msg[h->cur], h->msg[h->cur ^ 1], h->PSI,
                                                              h->P, (double)h->g.n, h->opt.damping, h->cur_it);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  return GBX_OK;
}
template <int MODE, int QT>
int l_vertex_q(gbx_lsbm* h) {
  // Large graphs fill the GPU with one thread per vertex (no redundant normalisation across lanes: 9.3 against 11.6-15.4
  // ms per iteration at N = 1M); small graphs are latency bound and want the links of a vertex spread over lanes.
  if (h->g.n >= 100000) return l_vertex_ql<MODE, QT, 1>(h);
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

// one pass of the notebook's main loop: h = update_h(); BP(); gr and affinity updates.  denom (sum_i degrees PSI) of the
// current marginals is already there: from init or from the end of the previous pass.
int l_iteration(gbx_lsbm* h) {
  cudaStream_t st = h->g.stream;
  const int it = h->cur_it;
  k_lsbm_pre<<<1, 64, 0, st>>>(h->P, h->q, h->G, it);           // h = update_h(); residual := 0
  h->launches++;
  L_TRY(l_vertex<LMODE_SWEEP>(h));                              // (BPconv, PSI) = BP()
  h->cur ^= 1;
  L_TRY(l_denom(h));                                            // from the new marginals
  if (h->opt.learning && h->g.K > 0) {
    const int mgrid = (int)std::min<int64_t>(std::max<int64_t>(1, (h->g.K + MSTEP_THREADS - 1) / MSTEP_THREADS),
                                             (int64_t)h->g.num_sms * 8);
    k_lsbm_mstep<<<mgrid, MSTEP_THREADS, (size_t)(MSTEP_THREADS / 32) * 32 * 2 * h->q * sizeof(double), st>>>(
        (long long)h->g.K, h->q, h->G, h->g.erow, h->g.col, h->g.rev, h->lab, h->msg[h->cur], h->P, it);
    h->launches++;
  }
  k_lsbm_post<<<1, 256, 0, st>>>(h->P, h->q, h->G, (double)h->g.n, (double)h->g.K, h->opt.learningrate,
                                 h->opt.priorlearning, h->opt.learning, h->opt.REDfrac != 0.0, it, h->cur_thr);
  h->launches++;
  h->itr++;
  GBX_CUDA_TRY(cudaGetLastError());
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
  if (!out) return l_invalid("out is NULL");
  *out = nullptr;
  if (q < 1 || q > MAXQ) return l_invalid("q must be in 1..16");
  if (n_labels < 1 || n_labels > LMAXG) return l_invalid("number of labels must be in 1..GBX_MAX_LABELS");
  if (n < 1 || K < 0 || (K > 0 && !links3)) return l_invalid("bad n_vertices / n_links / links");
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
  std::vector<int> rowptr_g;
  // the first two columns of the column-major K x 3 matrix are the K x 2 `links` the graph build expects
  int rc = build_graph_device(&h->g, links3, &rowptr_g, nullptr);
  if (rc != GBX_OK) return bail(rc);
  auto body = [&]() -> int {
    const int dv = device;
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->lab, std::max<size_t>((size_t)K, 1), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->deg, std::max<size_t>((size_t)n * h->G, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->dd, std::max<size_t>((size_t)K, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->msg[0], std::max<size_t>((size_t)K * q, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->msg[1], std::max<size_t>((size_t)K * q, 1) * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->PSI, (size_t)n * q * sizeof(double), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&h->P, sizeof(LParams), true));
    GBX_CUDA_TRY(cudaMallocHost((void**)&h->hP, sizeof(LParams)));
    GBX_CUDA_TRY(cudaMemset(h->P, 0, sizeof(LParams)));
    h->grid_v = (int)std::min<int64_t>(std::max<int64_t>(1, (n + 255) / 256), (int64_t)sms * 8);
    h->grid_w = (int)std::min<int64_t>(std::max<int64_t>(1, (n + 7) / 8), (int64_t)sms * 8);   // one warp per vertex
    h->grid_e = (int)std::min<int64_t>(std::max<int64_t>(1, (K + 255) / 256), (int64_t)sms * 8);
    // labels of the links -> slots; both directions of a link must agree
    int64_t* d_lab = nullptr;
    int* err = nullptr;
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&d_lab, std::max<size_t>((size_t)K, 1) * sizeof(int64_t), true));
    GBX_CUDA_TRY(dev_malloc(dv, (void**)&err, sizeof(int), true));
    int herr = 0;
    cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), h->g.stream);
    if (e == cudaSuccess && K > 0)
      e = cudaMemcpyAsync(d_lab, links3 + 2 * K, (size_t)K * sizeof(int64_t), cudaMemcpyDefault, h->g.stream);
    if (e == cudaSuccess && K > 0) {
      k_lab_slots<<<h->grid_e, 256, 0, h->g.stream>>>(d_lab, h->g.slot_of_link, (long long)K, h->G, h->lab, err);
      k_lab_check<<<h->grid_e, 256, 0, h->g.stream>>>(h->lab, h->g.rev, (long long)K, err);
      h->launches += 2;
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, h->g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->g.stream);
    dev_free(dv, d_lab);
    dev_free(dv, err);
    GBX_CUDA_TRY(e);
    if (herr & 1) return l_invalid("links: label out of 1..n_labels");
    if (herr & 2) return l_invalid("links: the two directions of a link carry different labels");
    return GBX_OK;
  };
  rc = body();
  if (rc != GBX_OK) return bail(rc);
  *out = h;
  return GBX_OK;
}

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
  if (h->hP) cudaFreeHost(h->hP);
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
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  const int q = h->q, G = h->G;
  cudaStream_t st = h->g.stream;
  LParams* s = h->hP;
  memset(s, 0, sizeof(LParams));
  for (int t = 0; t < q; ++t) s->gr[t] = gr[t];
  for (int a = 0; a < G; ++a)
    for (int k = 0; k < q * q; ++k) s->aff[a][k] = affinity[(size_t)a * q * q + k];
  GBX_CUDA_TRY(cudaMemcpyAsync(h->P, s, sizeof(LParams), cudaMemcpyHostToDevice, st));
  GBX_CUDA_TRY(cudaStreamSynchronize(st));
  h->cur = 0;
  h->itr = 0;
  const int64_t K = h->g.K;
  if (K > 0) {
    // the other message buffer is scratch before the first sweep
    GBX_CUDA_TRY(cudaMemcpyAsync(h->msg[1], psicav, (size_t)K * q * sizeof(double), cudaMemcpyHostToDevice, st));
    k_lsbm_permute<<<h->grid_e, 256, 0, st>>>(h->msg[1], h->msg[0], h->g.slot_of_link, (long long)K, q, 1);
    h->launches++;
  }
  k_lsbm_degrees<<<h->grid_v, 256, 0, st>>>(h->g.rowptr, h->lab, (int)h->g.n, G, h->opt.dc, h->deg);
  if (K > 0) k_lsbm_dd<<<h->grid_e, 256, 0, st>>>((long long)K, G, h->g.erow, h->g.col, h->lab, h->deg, (double)h->g.n, h->dd);
  h->launches += 2;
  GBX_CUDA_TRY(cudaGetLastError());
  // h = zeros(G,B); PSI = updatePSI(); h = update_h()   (notebook EM, "initial state")
  L_TRY(l_vertex<LMODE_MARGINAL>(h));
  L_TRY(l_denom(h));
  k_lsbm_pre<<<1, 64, 0, st>>>(h->P, q, G, 0);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  GBX_CUDA_TRY(cudaStreamSynchronize(st));
  h->initialised = true;
  return GBX_OK;
}

int gbx_lsbm_iterate(gbx_lsbm* h, int n_iters, double* bpconv_out) {
  if (!h) return l_invalid("handle is NULL");
  if (!h->initialised) {
    last_error() = "gbx_lsbm_init has not been called";
    return GBX_ERR_STATE;
  }
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  for (int i = 0; i < n_iters; ++i) L_TRY(l_iteration(h));
  if (bpconv_out) {
    L_TRY(l_read_conv(h));
    *bpconv_out = h->hP->conv;
  }
  return GBX_OK;
}

int gbx_lsbm_finalize(gbx_lsbm* h, gbx_lsbm_result* res) {
  if (!h || !res) return l_invalid("handle/result is NULL");
  if (!h->initialised) {
    last_error() = "gbx_lsbm_init has not been called";
    return GBX_ERR_STATE;
  }
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  cudaStream_t st = h->g.stream;
  const double L = 0.5 * (double)h->g.K, N = (double)h->g.n;
  k_lsbm_zero<<<1, 32, 0, st>>>(h->P, 4, 0);
  h->launches++;
  L_TRY(l_vertex<LMODE_LOGZ>(h));
  if (h->g.K > 0) {
    k_lsbm_fe<0><<<h->grid_e, 256, 0, st>>>((long long)h->g.K, h->q, h->G, h->g.erow, h->g.col, h->g.rev, h->lab, h->deg,
                                            h->msg[h->cur], h->P);
    k_lsbm_fe_mean<<<1, 32, 0, st>>>(h->P, L);
    k_lsbm_fe<1><<<h->grid_e, 256, 0, st>>>((long long)h->g.K, h->q, h->G, h->g.erow, h->g.col, h->g.rev, h->lab, h->deg,
                                            h->msg[h->cur], h->P);
    h->launches += 3;
  }
  k_lsbm_finalize<<<1, 32, 0, st>>>(h->P, N, L);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  L_TRY(l_sync_state(h));
  memset(res, 0, sizeof *res);
  const double* r = h->hP->result;
  res->FE = r[0];
  res->CVBayes = r[1];
  res->CVGP = r[2];
  res->CVGT = r[3];
  res->CVMAP = r[4];
  res->varCVBayes = r[5];
  res->varCVGP = r[6];
  res->varCVGT = r[7];
  res->varCVMAP = r[8];
  res->BPconv = h->hP->conv;
  res->nonfinite = h->hP->status & 1;
  return GBX_OK;
}

}  // extern "C"

namespace {

// The main loop of the notebook's EM() as a small state machine, so that many handles can be driven side by side:
// queue a batch of iterations on the handle's stream (plus the 4-byte read-back of the iteration at which the device
// found BPconv < threshold), later look at it.  Small graphs are launch bound: `batch` iterations are queued before
// the host looks; kernels of iterations after the converged one return at once (l_skip), so the state is the one
// the notebook's `break` leaves.
struct LsbmRun {
  gbx_lsbm* h = nullptr;
  int batch = 1, itr0 = 0, done = 0, cnv = 0, itrnum = 0;
  bool finished = false;
  cudaEvent_t ev = nullptr;
};

int lrun_begin(LsbmRun& r, gbx_lsbm* h) {
  r.h = h;
  r.batch = h->g.K < (int64_t)2000000 ? 8 : 1;
  if (const char* e = getenv("GBX_EM_BATCH")) r.batch = std::max(1, atoi(e));
  GBX_CUDA_TRY(cudaMemsetAsync(&h->P->done_at, 0, sizeof(int), h->g.stream));
  r.itr0 = h->itr;
  r.done = r.cnv = r.itrnum = 0;
  r.finished = h->opt.itrmax <= 0;
  h->cur_thr = h->opt.bp_conv_threshold;
  if (!r.ev) GBX_CUDA_TRY(cudaEventCreateWithFlags(&r.ev, cudaEventDisableTiming));
  return GBX_OK;
}
int lrun_queue(LsbmRun& r) {  // queue the next batch of iterations and the read-back
  gbx_lsbm* h = r.h;
  const int nb = std::min(r.batch, h->opt.itrmax - r.done);
  for (int b = 1; b <= nb; ++b) {
    h->cur_it = r.itr0 + r.done + b;
    L_TRY(l_iteration(h));
  }
  r.done += nb;
  GBX_CUDA_TRY(cudaMemcpyAsync(&h->hP->done_at, &h->P->done_at, sizeof(int), cudaMemcpyDeviceToHost, h->g.stream));
  GBX_CUDA_TRY(cudaEventRecord(r.ev, h->g.stream));
  return GBX_OK;
}
int lrun_poll(LsbmRun& r) {  // wait for the queued batch; decide whether the run is over
  gbx_lsbm* h = r.h;
  cudaError_t ce = cudaEventSynchronize(r.ev);
  if (ce != cudaSuccess) {
    last_error() = std::string("gbx_lsbm_em: ") + cudaGetErrorString(ce);
    return GBX_ERR_CUDA;
  }
  const int done_at = h->hP->done_at;
  if (done_at != 0) {
    if ((r.itr0 + r.done - done_at) & 1) h->cur ^= 1;   // queued passes that did nothing still flipped the buffers
    h->itr = done_at;
    r.cnv = 1;
    r.itrnum = done_at - r.itr0;
    r.done = r.itrnum;
    r.finished = true;
  } else if (r.done >= h->opt.itrmax) {
    r.finished = true;
  }
  return GBX_OK;
}
int lrun_end(LsbmRun& r, gbx_lsbm_result* res) {
  gbx_lsbm* h = r.h;
  h->cur_it = 0;
  h->cur_thr = -1.0;
  L_TRY(gbx_lsbm_finalize(h, res));
  res->cnv = r.cnv;
  res->itrnum = r.itrnum;
  res->itrs_done = r.done;
  return GBX_OK;
}

}  // namespace

extern "C" {

int gbx_lsbm_em(gbx_lsbm* h, gbx_lsbm_result* res) {
  if (!h || !res) return l_invalid("handle/result is NULL");
  if (!h->initialised) {
    last_error() = "gbx_lsbm_init has not been called";
    return GBX_ERR_STATE;
  }
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  LsbmRun r;
  int rc = lrun_begin(r, h);
  while (rc == GBX_OK && !r.finished) {
    rc = lrun_queue(r);
    if (rc == GBX_OK) rc = lrun_poll(r);
  }
  if (rc == GBX_OK) rc = lrun_end(r, res);
  if (r.ev) cudaEventDestroy(r.ev);
  h->cur_it = 0;
  h->cur_thr = -1.0;
  return rc;
}

// EM() of `count` initialised handles of ONE device, side by side: every handle gets a stream of its own for the call
// (up to `max_concurrent` run at a time, 0 = 16), batches of iterations of all running handles are queued before any is
// waited for, and a finished handle hands its stream to the next one.  The graphs of a detectability sweep (BASELINE
// configs[4]: thousands of N = 10k graphs) are launch bound one at a time; side by side their kernels overlap on the GPU
// and the host's launch latency is hidden.  Results are those of gbx_lsbm_em on each handle (same kernels, same order).
int gbx_lsbm_em_batch(gbx_lsbm** hs, int count, gbx_lsbm_result* res, int max_concurrent) {
  if (count < 0 || (count > 0 && (!hs || !res))) return l_invalid("handles/results is NULL");
  if (count == 0) return GBX_OK;
  const int dev = hs[0] ? hs[0]->g.device : 0;
  for (int i = 0; i < count; ++i) {
    if (!hs[i]) return l_invalid("a handle is NULL");
    if (!hs[i]->initialised) {
      last_error() = "gbx_lsbm_init has not been called on every handle";
      return GBX_ERR_STATE;
    }
    if (hs[i]->g.device != dev) return l_invalid("the handles of a batch must live on one device");
  }
  GBX_CUDA_TRY(cudaSetDevice(dev));
  const int W = std::max(1, std::min(count, max_concurrent > 0 ? max_concurrent : 16));
  std::vector<cudaStream_t> streams((size_t)W, nullptr);
  std::vector<cudaStream_t> saved((size_t)count, nullptr);
  std::vector<LsbmRun> run((size_t)W);
  std::vector<int> who((size_t)W, -1);
  int rc = GBX_OK, next = 0, live = 0;
  for (int w = 0; w < W && rc == GBX_OK; ++w)
    if (cudaStreamCreateWithFlags(&streams[(size_t)w], cudaStreamNonBlocking) != cudaSuccess) rc = GBX_ERR_CUDA;
  auto start = [&](int w) -> int {
    const int i = next++;
    who[(size_t)w] = i;
    saved[(size_t)i] = hs[i]->g.stream;
    cudaStreamSynchronize(hs[i]->g.stream);   // whatever init queued on the handle's own stream
    hs[i]->g.stream = streams[(size_t)w];
    ++live;
    return lrun_begin(run[(size_t)w], hs[i]);
  };
  for (int w = 0; w < W && next < count && rc == GBX_OK; ++w) rc = start(w);
  while (live > 0 && rc == GBX_OK) {
    for (int w = 0; w < W && rc == GBX_OK; ++w)
      if (who[(size_t)w] >= 0 && !run[(size_t)w].finished) rc = lrun_queue(run[(size_t)w]);
    for (int w = 0; w < W && rc == GBX_OK; ++w) {
      if (who[(size_t)w] < 0) continue;
      if (!run[(size_t)w].finished) rc = lrun_poll(run[(size_t)w]);
      if (rc == GBX_OK && run[(size_t)w].finished) {
        const int i = who[(size_t)w];
        rc = lrun_end(run[(size_t)w], &res[i]);
        hs[i]->g.stream = saved[(size_t)i];
        who[(size_t)w] = -1;
        --live;
        if (rc == GBX_OK && next < count) rc = start(w);
      }
    }
  }
  for (int w = 0; w < W; ++w) {   // on an error path: give the handles their streams back
    if (who[(size_t)w] >= 0) {
      cudaStreamSynchronize(streams[(size_t)w]);
      hs[who[(size_t)w]]->g.stream = saved[(size_t)who[(size_t)w]];
      hs[who[(size_t)w]]->cur_it = 0;
      hs[who[(size_t)w]]->cur_thr = -1.0;
    }
    if (run[(size_t)w].ev) cudaEventDestroy(run[(size_t)w].ev);
    if (streams[(size_t)w]) {
      cudaStreamSynchronize(streams[(size_t)w]);
      cudaStreamDestroy(streams[(size_t)w]);
    }
  }
  return rc;
}

int gbx_lsbm_get_state(gbx_lsbm* h, double* gr, double* affinity, double* PSI, double* hfield) {
  if (!h) return l_invalid("handle is NULL");
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  L_TRY(l_sync_state(h));
  const int q = h->q, G = h->G;
  if (gr) memcpy(gr, h->hP->gr, q * sizeof(double));
  if (affinity)
    for (int a = 0; a < G; ++a) memcpy(affinity + (size_t)a * q * q, h->hP->aff[a], (size_t)q * q * sizeof(double));
  if (hfield)
    for (int a = 0; a < G; ++a) memcpy(hfield + (size_t)a * q, h->hP->h[a], q * sizeof(double));
  if (PSI) {
    GBX_CUDA_TRY(cudaMemcpyAsync(PSI, h->PSI, (size_t)h->g.n * q * sizeof(double), cudaMemcpyDeviceToHost, h->g.stream));
    GBX_CUDA_TRY(cudaStreamSynchronize(h->g.stream));
  }
  return GBX_OK;
}

int gbx_lsbm_get_messages(gbx_lsbm* h, double* psicav) {
  if (!h || !psicav) return l_invalid("handle/psicav is NULL");
  GBX_CUDA_TRY(cudaSetDevice(h->g.device));
  const int64_t K = h->g.K;
  if (K == 0) return GBX_OK;
  double* tmp = h->msg[h->cur ^ 1];  // scratch between sweeps
  k_lsbm_permute<<<h->grid_e, 256, 0, h->g.stream>>>(h->msg[h->cur], tmp, h->g.slot_of_link, (long long)K, h->q, 0);
  h->launches++;
  GBX_CUDA_TRY(cudaGetLastError());
  GBX_CUDA_TRY(cudaMemcpyAsync(psicav, tmp, (size_t)K * h->q * sizeof(double), cudaMemcpyDeviceToHost, h->g.stream));
  GBX_CUDA_TRY(cudaStreamSynchronize(h->g.stream));
  return GBX_OK;
}

int64_t gbx_lsbm_launch_count(const gbx_lsbm* h) { return h ? h->launches + h->g.launches : 0; }

}  // extern "C"
