/* A plain C program against include/graphbix_cuda.h: what a non-Python host (the ccall stub of INTEGRATION.md) does.
 * Reads a problem from a binary file written by tests/test_gpu_cabi_client.py, runs EM() through the C ABI and writes
 * the results back; the test compares them with the ctypes path.  Test infrastructure only.
 *   gcc -O2 -I include tests/cabi/cabi_client.c -o cabi_client -L graphbix_b200 -lgraphbix_cuda -Wl,-rpath,...
 * file layout (little endian): int64 n, K, q; int64 links[2*K] (column-major K x 2); double gr[q]; double Crs[q*q];
 * double psi[K*q] */
#include <stdio.h>
#include <stdlib.h>

#include "graphbix_cuda.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != GBX_OK) {                                                              \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, gbx_last_error());                \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int64_t hdr[3];
  if (fread(hdr, sizeof(int64_t), 3, f) != 3) return 1;
  const int64_t n = hdr[0], K = hdr[1];
  const int q = (int)hdr[2];
  int64_t* links = malloc(sizeof(int64_t) * 2 * K);
  double* gr = malloc(sizeof(double) * q);
  double* Crs = malloc(sizeof(double) * q * q);
  double* psi = malloc(sizeof(double) * K * q);
  double* PSI = malloc(sizeof(double) * n * q);
  int32_t* block = malloc(sizeof(int32_t) * n);
  if (fread(links, sizeof(int64_t), 2 * K, f) != (size_t)(2 * K) || fread(gr, sizeof(double), q, f) != (size_t)q ||
      fread(Crs, sizeof(double), q * q, f) != (size_t)(q * q) || fread(psi, sizeof(double), K * q, f) != (size_t)(K * q))
    return 1;
  fclose(f);

  gbx_sbm* h = NULL;
  CHECK(gbx_sbm_create(&h, 0, n, K, links, q));
  gbx_sbm_options opt;
  gbx_sbm_default_options(&opt);
  opt.itrmax = 300;
  CHECK(gbx_sbm_set_options(h, &opt));
  CHECK(gbx_sbm_init(h, gr, Crs, psi, 0));
  gbx_sbm_result res;
  CHECK(gbx_sbm_em(h, &res));
  CHECK(gbx_sbm_get_state(h, gr, Crs, PSI, NULL, NULL));
  CHECK(gbx_sbm_get_assignment(h, block));
  const long long launches = (long long)gbx_sbm_launch_count(h);
  CHECK(gbx_sbm_destroy(h));

  /* an error path: a self loop must be refused with GBX_ERR_INVALID and a message */
  links[K] = links[0];
  const int rc_bad = gbx_sbm_create(&h, 0, n, K, links, q);

  FILE* o = fopen(argv[2], "wb");
  if (!o) return 1;
  double scal[12] = {res.FE, res.CVBayes, res.CVGP, res.CVGT, res.CVMAP, (double)res.cnv, (double)res.itrnum,
                     (double)res.fail, (double)res.nonfinite, res.BPconv, (double)launches, (double)rc_bad};
  fwrite(scal, sizeof(double), 12, o);
  fwrite(gr, sizeof(double), q, o);
  fwrite(Crs, sizeof(double), q * q, o);
  fwrite(PSI, sizeof(double), n * q, o);
  fwrite(block, sizeof(int32_t), n, o);
  fclose(o);
  printf("cabi_client: n=%lld K=%lld q=%d itrnum=%d cnv=%d FE=%.12f launches=%lld bad_links_rc=%d (%s)\n", (long long)n,
         (long long)K, q, res.itrnum, res.cnv, res.FE, launches, rc_bad, rc_bad == GBX_OK ? "" : gbx_last_error());
  return 0;
}
