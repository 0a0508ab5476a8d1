/* CPU ORACLE (test infrastructure, NOT product code): C restatement of `EM()` of the reference,
 * /root/reference/src/sbm.jl:28-370, statement for statement like oracle/sbm_oracle.py but fast enough for
 * 10^5..10^6 directed edges.  PARITY UNPINNED against the Julia reference itself (no Julia here, no golden
 * vectors upstream); it is pinned against oracle/sbm_oracle.py (tests/test_oracle_c.py, same sweep orders ->
 * same numbers) and used as the CPU baseline of bench.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Build:  make -C oracle      (gcc -O2 -shared -fPIC -o oracle/_build/libsbm_oracle.so oracle/sbm_oracle.c -lm)
 *
 * Arithmetic follows the reference (log domain, logsumexp with sorting, the Float32 logbiprob of sbm.jl:170),
 * not the CUDA path. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAXB 32

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void sort_asc(double* a, int n) { /* sortslices(array, dims=2) of a 1 x n matrix, sbm.jl:31 */
  for (int i = 1; i < n; ++i) {
    double x = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
    a[j + 1] = x;
  }
}

/* sbm.jl:30-37 */
static double logsumexp(const double* array, int n) {
  double a[MAXB * MAXB];
  memcpy(a, array, n * sizeof(double));
  sort_asc(a, n);
  double maxval = a[n - 1];
  for (int k = 0; k < n; ++k)
    if (maxval - a[k] > 500) a[k] = maxval - 500;
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += exp(a[k] - a[0]);
  return a[0] + log(s);
}

/* the same on a Float32 vector (sbm.jl:170-176): clamp in Float32, sums in Float64 */
static double logsumexp_f32(const float* array, int n) {
  float a[MAXB * MAXB];
  memcpy(a, array, n * sizeof(float));
  for (int i = 1; i < n; ++i) {
    float x = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > x) { a[j + 1] = a[j]; --j; }
    a[j + 1] = x;
  }
  float maxval = a[n - 1];
  for (int k = 0; k < n; ++k)
    if (maxval - a[k] > 500.0f) a[k] = maxval - 500.0f;
  double s = 0.0;
  for (int k = 0; k < n; ++k) s += exp((double)a[k] - (double)a[0]);
  return (double)a[0] + log(s);
}

/* sbm.jl:39-45 (in place) */
static void normalize_logprob(double* a, int n) {
  const double lo = log(pow(10.0, -8.0));
  for (int r = 0; r < n; ++r)
    if (a[r] < lo) a[r] = lo;
  double lse = logsumexp(a, n);
  for (int r = 0; r < n; ++r) a[r] = exp(a[r] - lse);
}

typedef struct {
  int64_t N, K;
  int B;
  const int64_t* links; /* K x 2 row-major, 1-based */
  int64_t* ptr;         /* N+1: in-links of vertex i are in_link[ptr[i]..ptr[i+1]) sorted by source */
  int64_t* in_link;     /* link id of the message s -> i */
  int64_t* in_src;
  int64_t* rev;         /* rev[k] = link id of (j,i) for k = (i,j) */
  double* degrees;
  double *PSIcav, *PSI, *theta, *gr, *Crs, *h;
} Ctx;

static uint64_t rng_state;
static uint64_t rng_next(void) { /* splitmix64 */
  uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

typedef struct { int64_t key, id; } KeyId;
static int cmp_keyid(const void* a, const void* b) {
  int64_t x = ((const KeyId*)a)->key, y = ((const KeyId*)b)->key;
  return (x > y) - (x < y);
}

static int build(Ctx* c) {
  int64_t N = c->N, K = c->K;
  c->ptr = calloc(N + 1, sizeof(int64_t));
  c->in_link = malloc((K + 1) * sizeof(int64_t));
  c->in_src = malloc((K + 1) * sizeof(int64_t));
  c->rev = malloc((K + 1) * sizeof(int64_t));
  c->degrees = calloc(N, sizeof(double));
  KeyId* ks = malloc((K + 1) * sizeof(KeyId));
  if (!c->ptr || !c->in_link || !c->in_src || !c->rev || !c->degrees || !ks) return -1;
  for (int64_t k = 0; k < K; ++k) {
    int64_t s = c->links[2 * k] - 1, d = c->links[2 * k + 1] - 1;
    if (s < 0 || s >= N || d < 0 || d >= N) return -2;
    ks[k].key = d * N + s; /* destination-major, sources ascending: nb[i], sbm.jl:269-273 */
    ks[k].id = k;
    c->ptr[d + 1]++;
    c->degrees[d] += 1.0; /* sum(A, dims=2) of the symmetric A */
  }
  for (int64_t i = 0; i < N; ++i) c->ptr[i + 1] += c->ptr[i];
  qsort(ks, K, sizeof(KeyId), cmp_keyid);
  for (int64_t p = 0; p < K; ++p) {
    c->in_link[p] = ks[p].id;
    c->in_src[p] = ks[p].key % N;
  }
  /* reverse links: (j,i) sits in row j at the position of source i */
  for (int64_t p = 0; p < K; ++p) {
    int64_t i = ks[p].key / N, j = ks[p].key % N; /* message j -> i */
    int64_t lo = c->ptr[j], hi = c->ptr[j + 1];
    while (lo < hi) {
      int64_t mid = (lo + hi) / 2;
      if (c->in_src[mid] < i) lo = mid + 1; else hi = mid;
    }
    if (lo >= c->ptr[j + 1] || c->in_src[lo] != i) { free(ks); return -3; }
    c->rev[ks[p].id] = c->in_link[lo]; /* link (i -> j) */
  }
  free(ks);
  return 0;
}

/* PSIcav[k] * Crs (row vector times matrix) */
static void msg_times_C(const Ctx* c, int64_t k, double* out) {
  const int B = c->B;
  const double* m = c->PSIcav + k * B;
  for (int t = 0; t < B; ++t) {
    double s = 0.0;
    for (int r = 0; r < B; ++r) s += m[r] * c->Crs[r * B + t];
    out[t] = s;
  }
}

/* sbm.jl:83-98 */
static void logunnormalizedMessage(Ctx* c, int64_t i, double* out) {
  const int B = c->B;
  double lm[MAXB], mc[MAXB];
  for (int t = 0; t < B; ++t) lm[t] = 0.0;
  for (int64_t p = c->ptr[i]; p < c->ptr[i + 1]; ++p) {
    int64_t s = c->in_src[p];
    msg_times_C(c, c->in_link[p], mc);
    for (int t = 0; t < B; ++t) lm[t] += log(c->theta[i] * c->theta[s] * mc[t]);
  }
  double sum = 0.0;
  for (int r = 0; r < B; ++r) {
    if (isnan(c->gr[r]) || c->gr[r] < pow(10.0, -9.0)) c->gr[r] = 1.0 / (double)c->N;
    sum += c->gr[r];
  }
  for (int r = 0; r < B; ++r) c->gr[r] /= sum;
  for (int t = 0; t < B; ++t) out[t] = log(c->gr[t]) - c->theta[i] * c->h[t] + lm[t];
}

/* sbm.jl:70-73 */
static void update_h(Ctx* c) {
  const int B = c->B;
  double s[MAXB];
  for (int t = 0; t < B; ++t) s[t] = 0.0;
  for (int64_t i = 0; i < c->N; ++i)
    for (int t = 0; t < B; ++t) s[t] += c->theta[i] * c->PSI[i * B + t];
  for (int t = 0; t < B; ++t) {
    double x = 0.0;
    for (int r = 0; r < B; ++r) x += s[r] * c->Crs[r * B + t];
    c->h[t] = x / (double)c->N;
  }
}

/* sbm.jl:100-126 */
static double BP(Ctx* c, const int64_t* perm, int priorlearning) {
  const int B = c->B;
  const double N = (double)c->N;
  double conv = 0.0;
  update_h(c);
  double logPSIi[MAXB], mc[MAXB], cav[MAXB], hprev[MAXB], prev[MAXB], hn[MAXB];
  for (int64_t q = 0; q < c->N; ++q) {
    int64_t i = perm[q];
    logunnormalizedMessage(c, i, logPSIi);
    for (int64_t p = c->ptr[i]; p < c->ptr[i + 1]; ++p) {
      int64_t kji = c->in_link[p], kij = c->rev[kji], j = c->in_src[p];
      msg_times_C(c, kji, mc);
      for (int t = 0; t < B; ++t) cav[t] = logPSIi[t] - log(c->theta[j] * c->theta[i] * mc[t]);
      normalize_logprob(cav, B);
      memcpy(c->PSIcav + kij * B, cav, B * sizeof(double));
    }
    double* P = c->PSI + i * B;
    for (int t = 0; t < B; ++t) {
      double x = 0.0;
      for (int r = 0; r < B; ++r) x += c->theta[i] * P[r] * c->Crs[r * B + t];
      hprev[t] = x / N;
      prev[t] = P[t] / N;
    }
    logunnormalizedMessage(c, i, logPSIi);
    normalize_logprob(logPSIi, B);
    memcpy(P, logPSIi, B * sizeof(double));
    for (int t = 0; t < B; ++t) {
      double x = 0.0;
      for (int r = 0; r < B; ++r) x += c->theta[i] * P[r] * c->Crs[r * B + t];
      hn[t] = x / N;
    }
    for (int t = 0; t < B; ++t) {
      c->h[t] += hn[t] - hprev[t];
      if (priorlearning) c->gr[t] += P[t] / N - prev[t];
      conv += fabs(P[t] / N - prev[t]);
    }
  }
  return conv;
}

/* sbm.jl:128-148 */
static void update_Crs(Ctx* c, double maxCrs, double lr) {
  const int B = c->B;
  double Cn[MAXB * MAXB], cc[MAXB * MAXB];
  for (int k = 0; k < B * B; ++k) Cn[k] = 0.0;
  for (int64_t k = 0; k < c->K; ++k) {
    int64_t i = c->links[2 * k] - 1, j = c->links[2 * k + 1] - 1;
    const double *a = c->PSIcav + k * B, *b = c->PSIcav + c->rev[k] * B;
    double s = 0.0;
    for (int r = 0; r < B; ++r)
      for (int t = 0; t < B; ++t) {
        cc[r * B + t] = c->theta[i] * c->theta[j] * a[r] * b[t] * c->Crs[r * B + t];
        s += cc[r * B + t];
      }
    for (int x = 0; x < B * B; ++x) Cn[x] += cc[x] / s;
  }
  for (int r = 0; r < B; ++r)
    for (int t = 0; t < B; ++t) {
      double v = lr * Cn[r * B + t] / ((double)c->N * c->gr[r] * c->gr[t]) + (1 - lr) * c->Crs[r * B + t];
      if (v > maxCrs) v = maxCrs; else if (v < 0) v = 0;
      cc[r * B + t] = v;
    }
  memcpy(c->Crs, cc, B * B * sizeof(double));
}

/* sbm.jl:47-64 */
static void update_theta(Ctx* c) {
  const int B = c->B;
  double dr[MAXB], nr[MAXB];
  for (int t = 0; t < B; ++t) dr[t] = nr[t] = 0.0;
  int* block = malloc(c->N * sizeof(int));
  for (int64_t i = 0; i < c->N; ++i) {
    int best = 0;
    for (int t = 1; t < B; ++t)
      if (c->PSI[i * B + t] > c->PSI[i * B + best]) best = t; /* findmax: first maximum */
    block[i] = best;
    dr[best] += c->degrees[i];
    nr[best] += 1.0;
  }
  for (int t = 0; t < B; ++t) dr[t] /= nr[t];
  for (int64_t i = 0; i < c->N; ++i) c->theta[i] = c->degrees[i] / dr[block[i]];
  free(block);
}

static double variance(const double* x, int64_t n) { /* var(): unbiased */
  if (n < 2) return NAN;
  double m = 0.0;
  for (int64_t i = 0; i < n; ++i) m += x[i];
  m /= (double)n;
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += (x[i] - m) * (x[i] - m);
  return s / (double)(n - 1);
}

/* sbm.jl:150-213; out[9] = FE, CVBayes, CVGP, CVGT, CVMAP, var x4 */
static void freeenergy(Ctx* c, int float32_logbiprob, double* out) {
  const int B = c->B;
  const double N = (double)c->N;
  double v[MAXB];
  double sumlogZi = 0.0;
  for (int64_t i = 0; i < c->N; ++i) {
    logunnormalizedMessage(c, i, v);
    sumlogZi += logsumexp(v, B);
  }
  int64_t Ltot = c->K / 2, cnt = 0;
  double *lz = malloc((Ltot + 1) * sizeof(double)), *gp = malloc((Ltot + 1) * sizeof(double));
  double *gt = malloc((Ltot + 1) * sizeof(double)), *mp = malloc((Ltot + 1) * sizeof(double));
  double lb[MAXB * MAXB];
  float lbf[MAXB * MAXB];
  for (int64_t k = 0; k < c->K; ++k) {
    int64_t i = c->links[2 * k] - 1, j = c->links[2 * k + 1] - 1;
    if (i < j) continue;
    const double *a = c->PSIcav + k * B, *b = c->PSIcav + c->rev[k] * B;
    double lti = log(c->theta[i]), ltj = log(c->theta[j]);
    for (int s = 0; s < B; ++s)
      for (int t = 0; t < B; ++t) {
        lb[s * B + t] = log(a[s]) + lti + log(c->Crs[s * B + t]) + ltj + log(b[t]);
        lbf[s * B + t] = (float)lb[s * B + t];
      }
    lz[cnt] = (float32_logbiprob ? logsumexp_f32(lbf, B * B) : logsumexp(lb, B * B)) - log(N);
    double g1 = 0.0, g2 = 0.0;
    for (int s = 0; s < B; ++s)
      for (int t = 0; t < B; ++t) {
        double lterm = lti + log(c->Crs[s * B + t]) + ltj - log(N);
        g1 += a[s] * lterm * b[t];
        g2 += a[s] * c->theta[i] * c->Crs[s * B + t] * c->theta[j] * lterm * b[t] / (N * exp(lz[cnt]));
      }
    gp[cnt] = g1;
    gt[cnt] = g2;
    int sa = 0, sb = 0;
    for (int t = 1; t < B; ++t) {
      if (a[t] > a[sa]) sa = t;
      if (b[t] > b[sb]) sb = t;
    }
    mp[cnt] = lti + log(c->Crs[sa * B + sb]) + ltj - log(N);
    cnt++;
  }
  double ave = 0.0;
  for (int s = 0; s < B; ++s)
    for (int t = 0; t < B; ++t) ave += c->gr[s] * c->Crs[s * B + t] * c->gr[t];
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int64_t e = 0; e < cnt; ++e) { s0 += lz[e]; s1 += gp[e]; s2 += gt[e]; s3 += mp[e]; }
  out[0] = -((sumlogZi - s0) / N + 0.5 * ave);
  out[1] = 1 - s0 / (double)Ltot;
  out[2] = 1 - s1 / (double)Ltot;
  out[3] = 1 - s2 / (double)Ltot;
  out[4] = 1 - s3 / (double)Ltot;
  out[5] = variance(lz, cnt);
  out[6] = variance(gp, cnt);
  out[7] = variance(gt, cnt);
  out[8] = variance(mp, cnt);
  free(lz); free(gp); free(gt); free(mp);
}

/* EM(), sbm.jl:28-370.  perms: NULL (permutations drawn from `seed`) or itrmax x N 0-based vertex orders.
 * iout[0..3] = fail, cnv, itrnum, itrs_done;  dout[0] = BPconv, dout[1] = seconds spent in the EM loop
 * (sweeps + M-steps + theta updates, what bench.py's CPU baseline reports). */
int sbm_oracle_em(int64_t N, int64_t K, const int64_t* links, int B, double maxCrs, const double* gr0,
                  const double* Crs0, const double* psicav0, int degreecorrection, int itrmax, int priorlearning,
                  double learningrate, double bpconvthreshold, int float32_logbiprob, uint64_t seed,
                  const int64_t* perms, double* gr, double* Crs, double* PSI, double* PSIcav, double* theta,
                  double* scalars, int* iout, double* dout) {
  if (B < 1 || B > MAXB) return -10;
  Ctx c;
  memset(&c, 0, sizeof c);
  c.N = N; c.K = K; c.B = B; c.links = links;
  c.PSIcav = PSIcav; c.PSI = PSI; c.theta = theta; c.gr = gr; c.Crs = Crs;
  double h[MAXB];
  c.h = h;
  int rc = build(&c);
  if (rc) return rc;
  int fail = 0, cnv = 0, itrnum = 0, itrs = 0;
  double conv = NAN, tloop = 0.0;
  for (int64_t i = 0; i < N; ++i) theta[i] = 1.0;                      /* sbm.jl:277 */
  for (int r = 0; r < B; ++r) gr[r] = priorlearning ? gr0[r] : 1.0 / B; /* sbm.jl:304-306 */
  for (int k = 0; k < B * B; ++k) {                                      /* sbm.jl:308-317 */
    double v = Crs0[k];
    if (isnan(v) || isinf(v)) fail = 1;
    else if (v < pow(10.0, -6.0)) v = pow(10.0, -6.0);
    Crs[k] = v;
  }
  memcpy(PSIcav, psicav0, (size_t)K * B * sizeof(double));
  for (int t = 0; t < B; ++t) h[t] = 0.0;
  memset(PSI, 0, (size_t)N * B * sizeof(double));
  for (int k = 0; k < 9; ++k) scalars[k] = 0.0;
  if (!fail) {
    double v[MAXB];
    for (int64_t i = 0; i < N; ++i) {                                    /* updatePSI, sbm.jl:340 */
      logunnormalizedMessage(&c, i, v);
      normalize_logprob(v, B);
      memcpy(PSI + i * B, v, B * sizeof(double));
    }
    for (int t = 0; t < B; ++t) {                                        /* gr = update_gr(PSI), sbm.jl:341 */
      double s = 0.0;
      for (int64_t i = 0; i < N; ++i) s += PSI[i * B + t];
      gr[t] = s / (double)N;
    }
    int64_t* perm = malloc((N + 1) * sizeof(int64_t));
    rng_state = seed;
    double t0 = now_s();
    for (int itr = 1; itr <= itrmax; ++itr) {
      if (perms) memcpy(perm, perms + (int64_t)(itr - 1) * N, N * sizeof(int64_t));
      else {
        for (int64_t i = 0; i < N; ++i) perm[i] = i;
        for (int64_t i = N - 1; i > 0; --i) {
          int64_t j = (int64_t)(rng_next() % (uint64_t)(i + 1));
          int64_t tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
        }
      }
      conv = BP(&c, perm, priorlearning);
      update_Crs(&c, maxCrs, learningrate);
      if (degreecorrection) update_theta(&c);
      itrs = itr;
      if (conv < bpconvthreshold) { cnv = 1; itrnum = itr; break; }
    }
    tloop = now_s() - t0;
    free(perm);
    update_h(&c);                                                        /* sbm.jl:364 */
    for (int t = 0; t < B; ++t) {                                        /* sbm.jl:365 */
      double s = 0.0;
      for (int64_t i = 0; i < N; ++i) s += PSI[i * B + t];
      gr[t] = s / (double)N;
    }
    freeenergy(&c, float32_logbiprob, scalars);
  }
  iout[0] = fail; iout[1] = cnv; iout[2] = itrnum; iout[3] = itrs;
  dout[0] = conv; dout[1] = tloop;
  free(c.ptr); free(c.in_link); free(c.in_src); free(c.rev); free(c.degrees);
  return 0;
}
