/* Plain-C client of the drop-in boundary (include/b200lops.h): no CUDA headers, no Python, no torch.
 * Applies MPIFirstDerivative (centered, order 3, edge) to a HOST array through b2_first_derivative_host and
 * checks it against the stencil written out below (FirstDerivative.py:201-219).  Built and run by
 * tests/test_gpu_kernels.py::test_c_abi_client; exit code 0 = match. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "b200lops.h"

int main(void) {
  const size_t nr = 1037, nc = 259;
  const double h = 0.5;
  b2_ctx* ctx = NULL;
  int rc = b2_ctx_create(0, &ctx);
  if (rc != B2_OK) { fprintf(stderr, "b2_ctx_create: %s\n", b2_strerror(rc)); return 2; }
  double* x = (double*)malloc(nr * nc * sizeof(double));
  double* y = (double*)malloc(nr * nc * sizeof(double));
  if (!x || !y) return 3;
  unsigned long long s = 88172645463325252ULL;
  for (size_t i = 0; i < nr * nc; ++i) {  /* xorshift64 */
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    x[i] = (double)(s >> 11) / 9007199254740992.0 - 0.5;
  }
  rc = b2_first_derivative_host(ctx, x, y, nr, nc, 0, nr, B2_FD_CENTERED, 3, 1, h, 0, B2_F64);
  if (rc != B2_OK) { fprintf(stderr, "b2_first_derivative_host: %s\n", b2_strerror(rc)); return 4; }
  double worst = 0.0;
  for (size_t i = 0; i < nr; ++i)
    for (size_t j = 0; j < nc; ++j) {
      double ref;
      if (i == 0) ref = (x[nc + j] - x[j]) / h;
      else if (i == nr - 1) ref = (x[i * nc + j] - x[(i - 1) * nc + j]) / h;
      else ref = 0.5 * (x[(i + 1) * nc + j] - x[(i - 1) * nc + j]) / h;
      double e = fabs(ref - y[i * nc + j]);
      if (e > worst) worst = e;
    }
  int sms = 0;
  b2_ctx_sm_count(ctx, &sms);
  printf("abi_smoke: version %d, sm_count %d, max |err| = %.3e\n", b2_version(), sms, worst);
  b2_ctx_destroy(ctx);
  free(x);
  free(y);
  return worst < 1e-12 ? 0 : 1;
}
