// "Pull" BP sweep (included by gbx_sbm_q.cu inside its anonymous namespace; one instance per compiled q and model).
//
// The push sweep (sbm_vertex_kernel) scatters every new message i -> j into the row of j: half of its HBM traffic is a
// random 64-byte store, which B200 serves at half the streaming rate (profiles/r01_scatter_probe.txt).  Here a vertex
// keeps its OUTGOING messages in its own contiguous rows and rebuilds what it receives:
//
//     psi_{i->j}(n)  =  normalize( clamp( U_i(n) / T(psi_{j->i}(n-1)) ) )                  sbm.jl:108-109
//
// where U_i(n) = prior_i * prod_s T(psi_{s->i}(n-1)) is the unnormalised marginal of i after sweep n (sbm.jl:83-98) and
// psi_{j->i}(n-1) is the message j itself emitted one sweep earlier.  So sweep n+1 at vertex j
//   * streams its own rows psi_{j->i}(n-1) (contiguous, TMA tensor copies with the 32/64/128-byte hardware swizzle so
//     that one thread per row reads them without bank conflicts),
//   * gathers one per-vertex record U_i(n) per neighbour from an N x q array that lives in the 126 MB L2 (64 MB at
//     N = 1M, q = 8),
//   * rebuilds psi_{i->j}(n), forms the factors T, the product U_j(n+1), the marginal, and
//   * writes psi_{j->i}(n+1) back IN PLACE over psi_{j->i}(n-1) (same slot, same thread) and its own record U_j(n+1).
// Messages of even and odd generations therefore live in two buffers that are each rewritten in place every other sweep;
// all HBM streams are contiguous, the random access is an L2-resident gather.  The schedule, every formula and the order
// of the floating-point operations inside a vertex are those of the push sweep (tests/sync_model.py states them).
//
// A record holds U''_t = u_t / (floor_i * theta_i): the floor of normalize_logprob (1e-8 in true magnitude, sbm.jl:42) and
// the degree-correction factor theta_i used in the sweep that produced it are folded into the scale, so the consumer's
// clamp is "U''_t / (Crs psi)_t < theta_j" and no per-vertex floor has to travel.  The sign bits of the first log2(q)
// components carry the MAP block of i (U is positive), from which the consumer forms theta_i = degree_i / <degree of the
// block> for the new factor (sbm.jl:47-64, 87).
//
// Warp-specialised, mbarrier-phased: warp 4 is the producer (lane 0 issues the TMA copies of the streamed tile data into
// a 3-stage ring; all 32 lanes issue the dependent 16-byte cp.async gathers of the records into a 2-stage ring, one tile
// behind), warps 0-3 are one thread per edge slot / q lanes per vertex as in the push kernel.
#pragma once

#include <cuda.h>


constexpr bool PULL_OK = (QM == 4 || QM == 8 || QM == 16);     // rows of 32 / 64 / 128 bytes: a hardware swizzle fits
constexpr int SWZ_MASK = (ROWB <= 32) ? 1 : (ROWB <= 64 ? 3 : 7);
constexpr int NB_BLK = (Q <= 1) ? 0 : (Q <= 2 ? 1 : (Q <= 4 ? 2 : (Q <= 8 ? 3 : 4)));   // sign bits that carry the MAP block
#ifndef GBX_PS
#define GBX_PS 3
#endif
#ifndef GBX_PR
#define GBX_PR 2
#endif
constexpr int P_S = GBX_PS;             // stages of the streamed ring: the producer runs P_S - 1 tiles ahead
constexpr int P_R = GBX_PR;             // stages of the gathered ring: records are fetched P_R - 1 tiles ahead
constexpr int P_NT = NT + 32;           // threads of a pull CTA: NT consumers + one producer warp
#ifndef GBX_GATHER_LATE
#define GBX_GATHER_LATE 1
#endif
// The record gathers of the next tile are issued by the upper half of the consumer warps between phase A and phase B:
// with ~8 vertices per tile only the lower warps have vertex work in phase B, so the gathers leave the critical path
// (0: every warp gathers its own rows at the top of a tile, as in the first version).
constexpr bool GATHER_LATE = GBX_GATHER_LATE != 0;
static_assert(P_S > P_R || (GATHER_LATE && P_S == P_R), "a tile's neighbour ids must have been streamed before its records can be gathered");
constexpr int G_T0 = GATHER_LATE ? NT / 2 : 0;        // first gathering thread
constexpr int G_NT = NT - G_T0;                       // gathering threads
constexpr int QTP = Q + 4;              // prior-table row: E_0..E_{Q-1}, ffrac, fi, xmax, 2^-ffrac

__device__ __forceinline__ int swz(int off) { return off ^ (((off >> 7) & SWZ_MASK) << 4); }

// ---- PTX: mbarrier, TMA, cp.async ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Wait for the phase of parity `parity` to complete.  try_wait suspends the thread in hardware for up to the hinted
// time; a thread that still finds the phase open backs off with nanosleep so that its polling does not take issue
// slots from the warps that have work (ncu: an un-throttled poll loop was 30% of the executed instructions).
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_t(unsigned long long* bar, unsigned parity) {
  const unsigned addr = smem_u32(bar);
  for (;;) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) break;
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
  }
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) { mbar_wait_t<20>(bar, parity); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 1-D bulk copy global -> shared (UBLKCP); dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// 2-D tensor copy (UTMALDG): box of the tensor map at (c0, c1) -> shared, hardware swizzle applied
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, unsigned long long* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void cp_async_16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// x * 2^n for any n (two exact steps), n clamped to what a double can express
__device__ __forceinline__ double mul_pow2(double x, int n) {
  n = max(-2040, min(2040, n));
  const int n1 = n / 2, n2 = n - n1;
  return x * pow2i(n1) * pow2i(n2);
}

// ---- shared memory of a pull CTA -------------------------------------------------------------------------------------
// Shared memory of a pull CTA (bytes).  Message stages first (1024-byte aligned: swizzle atoms), then the small per-tile
// arrays of every stage (16-byte aligned pieces), the gathered records (P_R stages; after phase A a slot's row holds its
// factor T, later its M-step operand: the record is dead by then), per-vertex scratch and the mbarriers.
struct PStage {  // the small arrays of one streamed stage
  static constexpr size_t col_off = 0;                                        // TILE_E + 4 neighbour ids
  static constexpr size_t deg_off = col_off + (size_t)(TILE_E + 4) * 4;      // TILE_E + 4 neighbour degrees (float)
  static constexpr size_t rp_off = deg_off + (size_t)(TILE_E + 4) * 4;       // TV + 8 row pointers
  static constexpr size_t pi_off = rp_off + (size_t)(TV + 8) * 4;            // TV + 4 prior-table rows
  static constexpr size_t thc_off = pi_off + (size_t)(TV + 4) * 4;           // TV + 2 current thetas
  static constexpr size_t thp_off = thc_off + (size_t)(TV + 2) * 8;          // TV + 2 thetas of the previous sweep
  static constexpr size_t desc_off = thp_off + (size_t)(TV + 2) * 8;         // the tile descriptor (int4)
  static constexpr size_t lv_off = desc_off + 16;                            // TILE_E + 16: vertex (inside the tile) of a slot
  static constexpr size_t bytes = ((lv_off + (size_t)TILE_E + 16 + 15) / 16) * 16;
};
struct PullSmem {
  static constexpr size_t msg_bytes = (size_t)TILE_E * ROWB;                         // one stage of message rows
  static constexpr size_t msg_off = 0;                                                // P_S stages
  static constexpr size_t gat_off = msg_off + P_S * msg_bytes;                       // P_R x TILE_E gathered records
  static constexpr size_t small_off = gat_off + (size_t)P_R * msg_bytes;             // P_S x PStage
  static constexpr size_t u_off = ((small_off + P_S * PStage::bytes + 15) / 16) * 16;
  static constexpr size_t ff_off = u_off + (size_t)TV * Q * 8;
  static constexpr size_t scratch_off = ff_off + (size_t)TV * 8;
  static constexpr size_t dr_off = scratch_off + (size_t)(NT > 64 * MB * MB ? NT : 64 * MB * MB) * 8;
  static constexpr size_t nr_off = dr_off + (size_t)Q * 8;
  static constexpr size_t fn_off = nr_off + (size_t)((Q + 1) / 2 * 2) * 4;
  static constexpr size_t bar_off = ((fn_off + (size_t)TV * 4 + 7) / 8) * 8;          // 2 P_S + P_R mbarriers
  static constexpr size_t total = ((bar_off + (size_t)(2 * P_S + P_R) * 8 + 127) / 128) * 128 + 1024;  // + alignment slack
};

#ifdef GBX_FORCE_CTAS
constexpr int PULL_MIN_CTAS = GBX_FORCE_CTAS;
#else
constexpr int PULL_MIN_CTAS = (PullSmem::total + 1024 <= 56 * 1024) ? 4 : (PullSmem::total + 1024 <= 75 * 1024) ? 3 : ((PullSmem::total + 1024 <= 113 * 1024) ? 2 : 1);
#endif

struct PullArgs {
  const int4* tiles;   // (first vertex, #vertices, first edge slot, #edge slots)
  int ntiles;
  const int* rowptr;
  const int* col;        // neighbour of every slot (global vertex id)
  const unsigned char* slot_lv;   // vertex (inside its tile) of every slot
  const float* coldeg;   // its degree (degree-corrected sbm.jl runs)
  const int* pidx;
  const double* prior_tab;
  const double* theta_cur;   // per local vertex: theta now / theta the previous sweep ran with
  const double* theta_prev;
  double* PSI;
  double* msg;             // the message buffer this pass streams (a sweep rewrites it in place)
  const double* rec_old;   // records of the previous sweep, by global vertex id (all ranks' vertices)
  double* rec_new;         // this rank's copy of the new records ...
  double* rec_peer[MAXPEER];   // ... and every other rank's (peer-mapped over NVLink): a vertex's record is stored
  int npeer;                   //     into all of them, so the next sweep gathers locally
  long long vbeg;
  double* partials;
  double* mpart;
  const SbmState* st;
  int it;
  int dc;         // degree correction with live thetas (update_theta has run)
  int dc_model;   // degree correction switched on at all
  const int* hubs;
};

// One message row (Q components) of a swizzled shared-memory tile, read by the thread that owns the row.
__device__ __forceinline__ void read_swz(const char* base, int row, double (&m)[Q]) {
#pragma unroll
  for (int c = 0; c < (Q + 1) / 2; ++c) {
    const double2 v = *reinterpret_cast<const double2*>(base + swz(row * ROWB + c * 16));
    m[2 * c] = v.x;
    if (2 * c + 1 < Q) m[2 * c + 1] = v.y;
  }
}
__device__ __forceinline__ void write_swz(char* base, int row, const double (&m)[Q]) {
#pragma unroll
  for (int c = 0; c < (Q + 1) / 2; ++c)
    *reinterpret_cast<double2*>(base + swz(row * ROWB + c * 16)) = make_double2(m[2 * c], (2 * c + 1 < Q) ? m[2 * c + 1] : 0.0);
}

__device__ __forceinline__ double* swz_elem(char* base, int row, int t) {
  return reinterpret_cast<double*>(base + swz(row * ROWB + (t >> 1) * 16) + (t & 1) * 8);
}

// The message i -> j of the previous sweep, rebuilt from the record of i and the message j sent to i one sweep earlier:
// w_t = U''_t / (Cprev psi_old)_t, clamped from below at thr = theta_j of that sweep (the fold of the floor and of theta_i
// into U'' makes this the clamp of sbm.jl:42), normalised.  Also returns the MAP block of i from the sign bits.
__device__ __forceinline__ void rebuild_incoming(const double (&mold)[Q], double (&U)[Q], double thr, double (&psi)[Q], int& blk) {
  blk = 0;
#pragma unroll
  for (int b = 0; b < NB_BLK; ++b) blk |= ((unsigned)__double2hiint(U[b]) >> 31) << b;
  double S = 0.0;
  bool low = false;
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    double To;
    if constexpr (MODEL == MODEL_MOD) {
      To = fma(mold[t], c_P.eb1_prev, 1.0);
    } else {
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < Q; ++s) acc = fma(mold[s], c_P.Cprev[s * Q + t], acc);
      To = acc;
    }
    psi[t] = fabs(U[t]) * fast_rcp(To);
    low = low || (psi[t] < thr);
  }
  if (low) {
#pragma unroll
    for (int t = 0; t < Q; ++t) psi[t] = (psi[t] < thr) ? thr : psi[t];
  }
#pragma unroll
  for (int t = 0; t < Q; ++t) S += psi[t];
  const double inv = fast_rcp(S);
#pragma unroll
  for (int t = 0; t < Q; ++t) psi[t] *= inv;
}

// Component t of the record of a vertex: u_t / (floor * theta), floor = 2^(fn + ffrac) in the scale of u, brought into
// the double range (where the shift has to be cut the clamp cannot bind any more: the largest component is then more
// than 2^900 above the floor).  eumax: largest biased exponent among the vertex's components.
__device__ __forceinline__ double record_component(double u, int fn, double ifrac, double invth, int eumax, bool neg) {
  int sh = -fn;
  const int top = eumax - 1023;
  if (sh + top > 900) sh = 900 - top;
  const double x = mul_pow2(u * ifrac * invth, sh);
  return neg ? -x : x;
}

// ----------------------------------------------------------------------------------------------------------------------
// MODE: MODE_SWEEP (messages, marginals, residual, M-step statistics), MODE_MARGINAL (marginals only), MODE_LOGZ.
// RECON: the incoming messages are rebuilt from records (any pass after the first sweep); otherwise `msg` holds the
// incoming messages themselves (the state right after init) and a sweep overwrites them with the outgoing ones.
template <int MODE, bool RECON>
__global__ void __launch_bounds__(P_NT, PULL_MIN_CTAS)
    sbm_pull_kernel(const __grid_constant__ CUtensorMap tmap, PullArgs a) {
  extern __shared__ __align__(1024) char smem_raw[];
  // 1024-byte alignment (swizzle atoms) by an offset, not through an integer cast: the pointer keeps its shared-memory
  // address space and the accesses below compile to LDS / STS rather than generic loads and stores
  char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  double* Usm = reinterpret_cast<double*>(smem + PullSmem::u_off);
  double* ffS = reinterpret_cast<double*>(smem + PullSmem::ff_off);
  double* scratch = reinterpret_cast<double*>(smem + PullSmem::scratch_off);
  unsigned long long* s_dr = reinterpret_cast<unsigned long long*>(smem + PullSmem::dr_off);
  int* s_nr = reinterpret_cast<int*>(smem + PullSmem::nr_off);
  int* fnS = reinterpret_cast<int*>(smem + PullSmem::fn_off);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + PullSmem::bar_off);
  unsigned long long* full_s = bars;                    // [P_S] streamed stage has landed
  unsigned long long* empty_s = bars + P_S;             // [P_S] consumers are done with it
  unsigned long long* full_g = bars + 2 * P_S;          // [P_R] gathered records have landed

  const int tid = threadIdx.x;
  if (MODE == MODE_SWEEP && skip_iteration(a.st, a.it)) return;
  const int G = gridDim.x;
  const int bid = blockIdx.x;
  const int nt = (bid < a.ntiles) ? (a.ntiles - bid + G - 1) / G : 0;   // tiles of this CTA
  constexpr bool THETA = (MODEL == MODEL_SBM);                            // theta arrays are streamed (when dc is on)

  if (tid == 0) {
    for (int s = 0; s < P_S; ++s) {
      mbar_init(&full_s[s], 1);
      mbar_init(&empty_s[s], NT / 32);
    }
    for (int r = 0; r < P_R; ++r) mbar_init(&full_g[r], G_NT);   // one cp.async arrival per gathering thread
    fence_barrier_init();
  }
  if (tid < Q) {
    s_dr[tid] = 0ull;
    s_nr[tid] = 0;
  }
  __syncthreads();

  if (tid >= NT) {
    // =================================================== producer warp ===================================================
    const int lane = tid & 31;
    auto load_stream = [&](int j) {
      const int s = j % P_S;
      __syncwarp();
      if (lane == 0) {   // one lane waits for the stage and issues its copies; the others meet it at the __syncwarp below
        if (j >= P_S) mbar_wait_t<200>(&empty_s[s], ((j / P_S) & 1) ^ 1);
        const int4 td = a.tiles[bid + j * G];
        char* sb = smem + PullSmem::small_off + (size_t)s * PStage::bytes;
        char* mb = smem + PullSmem::msg_off + (size_t)s * PullSmem::msg_bytes;
        *reinterpret_cast<int4*>(sb + PStage::desc_off) = td;
        const int v0 = td.x, nv = td.y, e0 = td.z, ne = td.w;
        unsigned long long* bar = &full_s[s];
        // sizes of the copies (16-byte windows rounded outwards), announced to the barrier before any copy is issued
        const int ea0 = e0 & ~3, va0 = v0 & ~3, ta0 = v0 & ~1;
        const unsigned nbox = (unsigned)((ne + 31) / 32);
        const unsigned nb_col = ne > 0 ? (unsigned)(((ne + (e0 - ea0)) * 4 + 15) & ~15) : 0u;
        const unsigned nb_deg = (MODEL == MODEL_SBM && a.dc_model) ? nb_col : 0u;
        const unsigned nb_rp = (unsigned)(((nv + 1 + (v0 - va0)) * 4 + 15) & ~15);
        const unsigned nb_pi = (unsigned)(((nv + (v0 - va0)) * 4 + 15) & ~15);
        const unsigned nb_th = (THETA && a.dc_model) ? (unsigned)(((nv + (v0 - ta0)) * 8 + 15) & ~15) : 0u;
        const int la0 = e0 & ~15;
        const unsigned nb_lv = ne > 0 ? (unsigned)((ne + (e0 - la0) + 15) & ~15) : 0u;
        mbar_arrive_expect_tx(bar, nbox * 32u * ROWB + nb_col + nb_deg + nb_rp + nb_pi + nb_th + (RECON ? nb_th : 0u) + nb_lv);
        for (unsigned b = 0; b < nbox; ++b)   // message rows, 32 at a time (rows past the buffer's end read as zeros)
          tma_load_2d(mb + (size_t)b * 32 * ROWB, &tmap, bar, 0, e0 + (int)b * 32);
        if (nb_col) bulk_load(sb + PStage::col_off, a.col + ea0, nb_col, bar);
        if (nb_lv) bulk_load(sb + PStage::lv_off, a.slot_lv + la0, nb_lv, bar);
        if (nb_deg) bulk_load(sb + PStage::deg_off, a.coldeg + ea0, nb_deg, bar);
        bulk_load(sb + PStage::rp_off, a.rowptr + va0, nb_rp, bar);
        bulk_load(sb + PStage::pi_off, a.pidx + va0, nb_pi, bar);
        if (nb_th) {
          bulk_load(sb + PStage::thc_off, a.theta_cur + ta0, nb_th, bar);
          if (RECON) bulk_load(sb + PStage::thp_off, a.theta_prev + ta0, nb_th, bar);
        }
      }
      __syncwarp();
    };
    // stream stages run two tiles ahead of the consumers; the consumers themselves issue the record gathers of the
    // tile after the one they work on (128 threads share that work; one warp could not keep up)
    for (int j = 0; j < P_S - 1 && j < nt; ++j) load_stream(j);
    for (int j = 0; j + P_S - 1 < nt; ++j) load_stream(j + P_S - 1);
    return;
  }

  // ===================================================== consumers =====================================================
  LaneAcc acc;
  MstepAcc macc;
#pragma unroll
  for (int k = 0; k < MB * MB * 2; ++k) macc.c[k] = 0.0;
  const int t = tid % LV;
  const bool tl = t < Q;
  const int tt = tl ? t : 0;
  const int vlane = tid / LV;
  const int vwarp = (tid & ~31) / LV;
  const int lane = tid & 31;

  // Gathers of tile jj's records into gather stage jj % P_R.  Every warp fetches the rows of its own 32 slots (CH
  // consecutive lanes one record): a stage's rows double as the slot's factor / M-step operand until the end of phase C
  // of tile jj - P_R, and with warp-private rows the overwrite is ordered by the warp's own program order (the other
  // warps read a row only in phase B, two consumer barriers earlier).
  auto issue_gather = [&](int jj) {
    if constexpr (RECON) {
      const int sg = jj % P_S, rg = jj % P_R;
      mbar_wait(&full_s[sg], (jj / P_S) & 1);   // its neighbour ids have landed
      const char* sbg = smem + PullSmem::small_off + (size_t)sg * PStage::bytes;
      const int4 tg = *reinterpret_cast<const int4*>(sbg + PStage::desc_off);
      const int* colS = reinterpret_cast<const int*>(sbg + PStage::col_off) + (tg.z & 3);
      char* gbg = smem + PullSmem::gat_off + (size_t)rg * PullSmem::msg_bytes;
      const char* rec = reinterpret_cast<const char*>(a.rec_old);
      if constexpr (GATHER_LATE) {
        // all rows of the tile, CH consecutive threads one record; every consumer warp is past phase C of the tile that
        // used this gather stage last (they met at the barrier after phase A), so any thread may write any row
        const int gt = tid - G_T0;
#pragma unroll
        for (int i = 0; i < TILE_E * CH / G_NT; ++i) {
          const int ul = gt + G_NT * i;
          const int row = ul / CH, c = ul % CH;
          if (row < tg.w) cp_async_16(gbg + swz(row * ROWB + c * 16), rec + ((size_t)colS[row] * ROWB + (size_t)c * 16));
        }
      } else {
        const int wrow = tid & ~31, ln = tid & 31;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
          const int ul = ln + 32 * i;
          const int row = wrow + ul / CH, c = ul % CH;
          if (row < tg.w) cp_async_16(gbg + swz(row * ROWB + c * 16), rec + ((size_t)colS[row] * ROWB + (size_t)c * 16));
        }
      }
      cp_async_mbar_arrive_noinc(&full_g[rg]);
    }
  };
  if (tid >= G_T0)
    for (int j = 0; j < P_R - 1 && j < nt; ++j) issue_gather(j);

  for (int j = 0; j < nt; ++j) {
    const int s = j % P_S, r = j % P_R;
    char* sb = smem + PullSmem::small_off + (size_t)s * PStage::bytes;
    char* gb = smem + PullSmem::gat_off + (size_t)r * PullSmem::msg_bytes;   // records, then factors T, then M-step rows
    if (!GATHER_LATE && j + P_R - 1 < nt) issue_gather(j + P_R - 1);
    mbar_wait(&full_s[s], (j / P_S) & 1);
    if constexpr (RECON) mbar_wait(&full_g[r], (j / P_R) & 1);
    const int4 td = *reinterpret_cast<const int4*>(sb + PStage::desc_off);
    const int v0 = td.x, nv = td.y, e0 = td.z, ne = td.w;
    const int* rpS = reinterpret_cast<const int*>(sb + PStage::rp_off) + (v0 & 3);
    const int* piS = reinterpret_cast<const int*>(sb + PStage::pi_off) + (v0 & 3);
    const double* thcS = reinterpret_cast<const double*>(sb + PStage::thc_off) + (v0 & 1);
    const double* thpS = reinterpret_cast<const double*>(sb + PStage::thp_off) + (v0 & 1);
    const unsigned char* lvS = reinterpret_cast<const unsigned char*>(sb + PStage::lv_off) + (e0 & 15);
    char* msgS = smem + PullSmem::msg_off + (size_t)s * PullSmem::msg_bytes;

    // prior factors of this tile's first round of vertices (an L2-resident table row per vertex) and their old marginals:
    // both loads are in flight through phase A
    Prior pr = load_prior<MODE>(a, (vlane < nv) ? piS[vlane] : 0, t);
    double old0 = 0.0;
    if (MODE == MODE_SWEEP && vlane < nv && tl) old0 = a.PSI[(size_t)(v0 + vlane) * Q + t];

    // ---- phase A: one thread per edge slot -- incoming message, its factor T ----
    double T[Q];
    double sc = 1.0;
    if (tid < ne) {
      double psi[Q];
      read_swz(msgS, tid, psi);
      const int lv = lvS[tid];
      int blk = 0;
      if constexpr (RECON) {
        double U[Q], mold[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) mold[k] = psi[k];
        read_swz(gb, tid, U);
        const double thr = (THETA && a.dc_model) ? thpS[lv] : 1.0;
        rebuild_incoming(mold, U, thr, psi, blk);
        if (MODE == MODE_SWEEP) write_swz(msgS, tid, psi);   // the consumed message, for the M-step pairing
      }
      if constexpr (MODEL == MODEL_SBM) {
        if (a.dc) {
          const float dg = reinterpret_cast<const float*>(sb + PStage::deg_off)[(e0 & 3) + tid];
          sc = thcS[lv] * ((double)dg * c_P.invdr[blk]);     // theta_j theta_i, sbm.jl:87
        }
      }
      edge_factor(psi, sc, T);
      write_swz(gb, tid, T);   // over the record it was rebuilt from
    }
    consumer_sync();
    if (GATHER_LATE && tid >= G_T0 && j + P_R - 1 < nt) issue_gather(j + P_R - 1);

    // ---- phase B: QP lanes per vertex -- product, prior, clamp floor, marginal, residual, the vertex's record ----
    for (int vb = 0; vb < nv; vb += VPR) {
      if (vb + vwarp >= nv) break;
      const int v = vb + vlane;
      const bool valid = v < nv;
      if (vb > 0) pr = load_prior<MODE>(a, valid ? piS[v] : 0, t);
      const int b = valid ? rpS[v] - e0 : 0, e = valid ? rpS[v + 1] - e0 : 0;
      double old = old0;
      if (vb > 0) old = (MODE == MODE_SWEEP && valid && tl) ? a.PSI[(size_t)(v0 + v) * Q + t] : 0.0;
      auto Tat = [&](int k) { return *swz_elem(gb, k, tt); };   // factor of slot k, this lane's component
      int k = b;
      double m1 = 1.0, m2 = 1.0;
      int K = 0;
      if (!__any_sync(0xffffffffu, e - b > c_P.safe_deg)) {
        for (; k + 1 < e; k += 2) {
          m1 *= Tat(k);
          m2 *= Tat(k + 1);
        }
        if (k < e) m1 *= Tat(k);
        m1 *= m2;
      } else {
        while (k + 7 < e) {
          m1 *= Tat(k);
          m2 *= Tat(k + 1);
          m1 *= Tat(k + 2);
          m2 *= Tat(k + 3);
          m1 *= Tat(k + 4);
          m2 *= Tat(k + 5);
          m1 *= Tat(k + 6);
          m2 *= Tat(k + 7);
          k += 8;
          renorm(m1, K);
          renorm(m2, K);
        }
        for (; k + 1 < e; k += 2) {
          m1 *= Tat(k);
          m2 *= Tat(k + 1);
        }
        if (k < e) m1 *= Tat(k);
        renorm(m1, K);
        renorm(m2, K);
        m1 *= m2;
        const int Kc = group_max((tl && m1 != 0.0) ? K : -(1 << 30));
        const int Kcc = (Kc == -(1 << 30)) ? 0 : Kc;
        const int d = K - Kcc;
        m1 = (d < -1000 || m1 == 0.0) ? 0.0 : m1 * pow2i(d > 0 ? 0 : d);
        K = Kcc;
      }
      double u;
      int fn, best = 0;
      vertex_finish<MODE>(a, valid, v0 + v, e - b, t, pr, old, m1, K, u, fn, acc, s_dr, s_nr, &best);
      if constexpr (MODE == MODE_SWEEP) {
        const int eumax = group_max(tl ? ((__double2hiint(u) >> 20) & 0x7ff) : 0);
        if (valid) {
          if (tl) Usm[v * Q + t] = u;
          if (t == 0) {
            ffS[v] = pr.ffrac;
            fnS[v] = fn;
          }
          const double invth = (THETA && a.dc) ? fast_rcp(thcS[v]) : 1.0;
          if (t < QM) {
            const double x = tl ? record_component(u, fn, pr.ifrac, invth, eumax, t < NB_BLK && ((best >> t) & 1)) : 0.0;
            const size_t ri = (size_t)(a.vbeg + v0 + v) * QM + t;
            a.rec_new[ri] = x;
            for (int p = 0; p < a.npeer; ++p) a.rec_peer[p][ri] = x;   // QP lanes write one contiguous row
          }
        }
      }
    }
    consumer_sync();

    // ---- phase C: one thread per edge slot -- outgoing message over the one it replaces, M-step statistics ----
    if constexpr (MODE == MODE_SWEEP) {
      double mod_term = 0.0;
      if (tid < ne) {
        const int lv = lvS[tid];
        double w[Q];
        emit_message_local(&Usm[lv * Q], T, ffS[lv], fnS[lv], w);
        store_msg<Q, QM>(a.msg, (long long)e0 + tid, w);
        if constexpr (MODEL == MODEL_SBM) {
          double z = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) z = fma(w[k], T[k], z);
          const double kz = sc * fast_rcp(z);
          double ar[Q];
#pragma unroll
          for (int k = 0; k < Q; ++k) ar[k] = w[k] * kz;
          write_swz(gb, tid, ar);   // the slot's row once more: a / Z for the M-step
        } else {
          double psi[Q];
          read_swz(msgS, tid, psi);
          double sdot = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) sdot = fma(w[k], psi[k], sdot);
          mod_term = c_P.oin * sdot / ((c_P.oin - c_P.oout) * sdot + c_P.oout);
        }
      }
      if constexpr (MODEL == MODEL_SBM) {
        __syncwarp();
        const int wbase = tid & ~31;
        mstep_accumulate_ab(macc, [&](int grp, int k, int comp) { return *swz_elem(gb, wbase + 4 * grp + k, comp); },
                            [&](int grp, int k, int comp) { return *swz_elem(msgS, wbase + 4 * grp + k, comp); }, ne - wbase);
        __syncwarp();
      } else {
        acc.mod += mod_term;
      }
    }
    // ---- hand the stages back to the producer ----
    fence_proxy_async();   // this thread's plain stores into the message stage precede the next TMA write there
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_s[s]);
  }
  // CTA-wide reductions of the consumers (the producer warp has left): barrier 1 over NT threads
  store_partials<true>(acc, scratch, a.partials + (size_t)blockIdx.x * PSTRIDE, s_dr, s_nr);
  if constexpr (MODE == MODE_SWEEP && MODEL == MODEL_SBM)
    store_mstep_partials<true>(macc, scratch, a.mpart + (size_t)blockIdx.x * Q * Q);
}

// ----------------------------------------------------------------------------------------------------------------------
// Hubs (degree > TILE_E): one CTA per vertex, two passes over its slots (the incoming messages are rebuilt twice rather
// than parked anywhere: hubs are rare); products carried as mantissa + exponent as in sbm_hub_kernel.
template <int MODE, bool RECON>
__global__ void __launch_bounds__(NT) sbm_pull_hub_kernel(PullArgs a, int part_base) {
  __shared__ double pmS[(NT / 32) * Q];
  __shared__ int pkS[(NT / 32) * Q];
  __shared__ int ksS[NT / 32];
  __shared__ double Ush[Q];
  __shared__ double scratch[(NT > 64 * MB * MB) ? NT : 64 * MB * MB];
  __shared__ double rowA[(MODEL == MODEL_SBM && MODE == MODE_SWEEP) ? NT * QS : 1];
  __shared__ double rowB[(MODEL == MODEL_SBM && MODE == MODE_SWEEP) ? NT * QS : 1];
  __shared__ double ffSh;
  __shared__ int fnSh;
  __shared__ unsigned long long s_dr[Q];
  __shared__ int s_nr[Q];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (MODE == MODE_SWEEP && skip_iteration(a.st, a.it)) return;
  const int gv = a.hubs[blockIdx.x];
  const int e0 = a.rowptr[gv], e1 = a.rowptr[gv + 1];
  const double thj = (MODEL == MODEL_SBM && a.dc) ? a.theta_cur[gv] : 1.0;
  const double thr = (MODEL == MODEL_SBM && a.dc_model && RECON) ? a.theta_prev[gv] : 1.0;
  // theta of the prior: the degree-correction factor (sbm.jl) or the degree itself (mod.jl:48)
  const double th = (MODEL == MODEL_MOD) ? (a.dc_model ? (double)(e1 - e0) : 1.0) : thj;
  if (tid < Q) {
    s_dr[tid] = 0ull;
    s_nr[tid] = 0;
  }
  auto incoming = [&](int e, double (&psi)[Q], double& sc) {
    load_msg<Q, QM>(a.msg, e, psi);
    int blk = 0;
    if constexpr (RECON) {
      double U[Q], mold[Q];
#pragma unroll
      for (int k = 0; k < Q; ++k) mold[k] = psi[k];
      load_msg<Q, QM>(a.rec_old, a.col[e], U);
      rebuild_incoming(mold, U, thr, psi, blk);
    }
    sc = 1.0;
    if constexpr (MODEL == MODEL_SBM) {
      if (a.dc) sc = thj * ((double)a.coldeg[e] * c_P.invdr[blk]);
    }
  };

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
    double psi[Q], T[Q], sc;
    int ke;
    incoming(e, psi, sc);
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
    const int d = k - kcc;
    const double prod = (d < -1022 || m == 0.0) ? 0.0 : m * pow2i(d > 0 ? 0 : d);
    const bool writer = lane < QP;
    const double old = (MODE == MODE_SWEEP && tl) ? a.PSI[(size_t)gv * Q + t] : 0.0;
    double row[QT];
    prior_row(th, row);
    Prior pr;
    pr.E = 0.0;
#pragma unroll
    for (int kk = 0; kk < Q; ++kk)
      if (kk == t) pr.E = row[kk];
    pr.ffrac = row[Q];
    pr.fi = row[Q + 1];
    pr.xmax = row[Q + 2];
    pr.ifrac = row[Q + 3];
    double u;
    int fn, best = 0;
    vertex_finish<MODE>(a, writer, gv, e1 - e0, t, pr, old, prod, Ks + kcc, u, fn, acc, s_dr, s_nr, &best);
    if constexpr (MODE == MODE_SWEEP) {
      const int eumax = group_max(tl ? ((__double2hiint(u) >> 20) & 0x7ff) : 0);
      if (writer) {
        if (tl) Ush[t] = u;
        if (t == 0) {
          ffSh = pr.ffrac;
          fnSh = fn;
        }
        const double invth = (MODEL == MODEL_SBM && a.dc) ? fast_rcp(thj) : 1.0;
        if (t < QM) {
          const double x = tl ? record_component(u, fn, pr.ifrac, invth, eumax, t < NB_BLK && ((best >> t) & 1)) : 0.0;
          const size_t ri = (size_t)(a.vbeg + gv) * QM + t;
          a.rec_new[ri] = x;
          for (int p = 0; p < a.npeer; ++p) a.rec_peer[p][ri] = x;
        }
      }
    }
  }
  __syncthreads();

  MstepAcc macc;
#pragma unroll
  for (int k = 0; k < MB * MB * 2; ++k) macc.c[k] = 0.0;
  if constexpr (MODE == MODE_SWEEP) {
    for (int base = e0; base < e1; base += NT) {
      const int e = base + tid;
      double mod_term = 0.0;
      if (e < e1) {
        double psi[Q], T[Q], w[Q], sc;
        int ke;
        incoming(e, psi, sc);
        edge_transform(psi, sc, T, ke);
        emit_message_local(Ush, T, ffSh, fnSh + ke, w);
        store_msg<Q, QM>(a.msg, e, w);
        if constexpr (MODEL == MODEL_SBM) {
          double z = 0.0;
#pragma unroll
          for (int k = 0; k < Q; ++k) z = fma(w[k], T[k], z);
          const double kz = sc * pow2i(-ke) * fast_rcp(z);
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
  acc.mod = macc_mod;
  store_partials(acc, scratch, a.partials + (size_t)(part_base + blockIdx.x) * PSTRIDE, s_dr, s_nr);
  if constexpr (MODE == MODE_SWEEP && MODEL == MODEL_SBM)
    store_mstep_partials(macc, scratch, a.mpart + (size_t)(part_base + blockIdx.x) * Q * Q);
}

// ----------------------------------------------------------------------------------------------------------------------
// freeenergy() on the pull layout: slot e of row i (column j) holds psi_{i->j}; psi_{j->i} is rebuilt from the record of j
// (or read from the incoming-message buffer before the first sweep).  Every undirected edge once, at i > j (sbm.jl:166).
struct PullFeArgs {
  const int* col;
  const int* erow;
  const float* coldeg;
  const double* msg_out;    // current outgoing messages
  const double* msg_other;  // RECON: the outgoing messages one sweep earlier; otherwise the incoming messages
  const double* rec;        // records of the last sweep
  const double* theta_cur;
  const double* theta_prev;
  const double* degs;       // mod.jl: degree by global vertex id
  long long vbeg;
  long long K;
  int dc;        // degree correction with live thetas (sbm.jl) / degree correction (mod.jl)
  int dc_model;
};

template <int PASS, bool RECON>
__global__ void __launch_bounds__(NT) pull_fe_kernel(const SbmState* st, PullFeArgs f, double* partials) {
  constexpr int NX = (MODEL == MODEL_MOD) ? 5 : 4;
  __shared__ double red[(NT / 32) * NX];
  double acc[NX], mean[NX];
#pragma unroll
  for (int k = 0; k < NX; ++k) {
    acc[k] = 0.0;
    mean[k] = (PASS == 1) ? (MODEL == MODEL_MOD ? st->cv5mean[k] : st->cvmean[k]) : 0.0;
  }
  const double oin = st->omegain, oout = st->omegaout, beta = st->beta;
  const double loin = (MODEL == MODEL_MOD) ? log(oin) : 0.0, loout = (MODEL == MODEL_MOD) ? log(oout) : 0.0;
  for (long long e = (long long)blockIdx.x * NT + threadIdx.x; e < f.K; e += (long long)gridDim.x * NT) {
    const int i = f.erow[e], j = f.col[e];
    if (i <= j) continue;
    double a[Q], b[Q];
    load_msg<Q, QM>(f.msg_out, e, a);
    load_msg<Q, QM>(f.msg_other, e, b);
    int blk = 0;
    if constexpr (RECON) {
      double U[Q], mold[Q];
#pragma unroll
      for (int k = 0; k < Q; ++k) mold[k] = b[k];
      load_msg<Q, QM>(f.rec, j, U);
      const double thr = (MODEL == MODEL_SBM && f.dc_model) ? f.theta_prev[i - f.vbeg] : 1.0;
      rebuild_incoming(mold, U, thr, b, blk);
    }
    double x[5];
    if constexpr (MODEL == MODEL_MOD) {
      double di = 1.0, dj = 1.0;
      if (f.dc) {
        di = f.degs[i];
        dj = f.degs[j];
      }
      fe_terms_mod(a, b, di, dj, oin, oout, beta, loin, loout, x);
    } else {
      double lc = -c_P.logN;
      if (f.dc) lc += log(f.theta_cur[i - f.vbeg]) + log((double)f.coldeg[e] * c_P.invdr[blk]);
      fe_terms_sbm(a, b, lc, x);
    }
#pragma unroll
    for (int k = 0; k < NX; ++k) {
      if (PASS == 0) acc[k] += x[k];
      else acc[k] += (x[k] - mean[k]) * (x[k] - mean[k]);
    }
  }
  block_reduce_store<NX>(acc, red, partials + (size_t)blockIdx.x * NX);
}

// ---- layout helpers ----------------------------------------------------------------------------------------------------
// degree of the neighbour of every slot
__global__ void k_coldeg(const int* __restrict__ col, const int* __restrict__ rowptr_global, long long K, float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x) {
    const int j = col[e];
    out[e] = (float)(rowptr_global[j + 1] - rowptr_global[j]);
  }
}
// Messages between `links` order and the pull layout of this rank's slots: slot e keeps the outgoing message
// out_link[e] and (before the first sweep) the incoming message in_link[e].
//   to_slots: in_buf[e] = src[in_link[e]], out_buf[e] = src[out_link[e]];  otherwise dst[out_link[e]] = out_buf[e]
__global__ void k_pull_permute(const double* __restrict__ src, double* __restrict__ dst, double* __restrict__ in_buf,
                               double* __restrict__ out_buf, const int* __restrict__ in_link, const int* __restrict__ out_link,
                               long long Kl, int to_slots) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < Kl; e += (long long)gridDim.x * blockDim.x) {
    double m[Q];
    if (to_slots) {
      load_vec<Q>(src, in_link[e], m);
      store_msg<Q, QM>(in_buf, e, m);
      load_vec<Q>(src, out_link[e], m);
      store_msg<Q, QM>(out_buf, e, m);
    } else {
      load_msg<Q, QM>(out_buf, e, m);
      store_vec<Q>(dst, out_link[e], m);
    }
  }
}
__device__ __forceinline__ void random_message(uint64_t seed, long long k, double (&m)[Q]) {
  double sum = 0.0;
#pragma unroll
  for (int t = 0; t < Q; ++t) {
    const uint64_t r = splitmix64(seed ^ splitmix64((uint64_t)k * Q + t));   // the draw of k_random_messages
    m[t] = ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    sum += m[t];
  }
#pragma unroll
  for (int t = 0; t < Q; ++t) m[t] /= sum;
}
__global__ void k_pull_random_messages(double* __restrict__ in_buf, double* __restrict__ out_buf,
                                       const int* __restrict__ in_link, const int* __restrict__ out_link, long long Kl,
                                       uint64_t seed) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < Kl; e += (long long)gridDim.x * blockDim.x) {
    double m[Q];
    random_message(seed, in_link[e], m);
    store_msg<Q, QM>(in_buf, e, m);
    random_message(seed, out_link[e], m);
    store_msg<Q, QM>(out_buf, e, m);
  }
}
// degree of the neighbour of every slot, from the degrees by global vertex id
__global__ void k_coldeg_global(const int* __restrict__ col, const double* __restrict__ deg_global, long long K,
                                float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < K; e += (long long)gridDim.x * blockDim.x)
    out[e] = (float)deg_global[col[e]];
}
