/* examples/c_abi_cold_start.c -- libl2b used from plain C through include/l2b.h: no Python, no torch.
 *
 * Cold start (every link = identity) on a 4 x 4 x 4 x 6 lattice, 3 chains:
 *   - l2b_su3_plaq_sums must return sum Re tr P = 18 V and sum Im tr P = 0 per chain
 *     (plaq = 1, S = -6 beta V; SURVEY appendix A.2 known answers),
 *   - one HMC trajectory with zero momenta must leave the links untouched (the force of a
 *     cold configuration vanishes) and return energies (KE0, S0, KE1, S1) with S0 == S1.
 *
 *   gcc -std=c99 -I. examples/c_abi_cold_start.c -o /tmp/cold -L l2hmc_b200 -ll2b \
 *       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/l2hmc_b200 -lm
 * Exit code 0 = checks passed, 77 = no CUDA device (nothing computed: there is no CPU fallback).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "include/l2b.h"

/* the four CUDA runtime calls this program needs, declared by hand so that it stays C99 */
int cudaMalloc(void** p, size_t n);
int cudaFree(void* p);
int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
int cudaDeviceSynchronize(void);
int cudaGetDeviceCount(int* n);
enum { H2D = 1, D2H = 2 };

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != 0) {                                                          \
      fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, l2b_last_error()); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  const int dims[4] = {4, 4, 4, 6}, nb = 3, nlf = 4;
  const size_t V = 4 * 4 * 4 * 6, nlinks = (size_t)nb * 4 * V, fbytes = nlinks * 9 * 2 * sizeof(double);
  const double beta = 6.0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != 0 || ndev == 0) {
    fprintf(stderr, "no CUDA device: libl2b has no CPU fallback (ABI version %d)\n", l2b_version());
    return 77;
  }
  double* hx = (double*)calloc(nlinks * 18, sizeof(double));
  for (size_t l = 0; l < nlinks; ++l) hx[l * 18 + 0] = hx[l * 18 + 8] = hx[l * 18 + 16] = 1.0;   /* identity */
  void *x, *v, *xo, *vo, *ws;
  double *sums, *en;
  const size_t wsb = l2b_su3_ws_bytes(nb, dims, L2B_F64);
  CHECK(cudaMalloc(&x, fbytes));
  CHECK(cudaMalloc(&v, fbytes));
  CHECK(cudaMalloc(&xo, fbytes));
  CHECK(cudaMalloc(&vo, fbytes));
  CHECK(cudaMalloc(&ws, wsb));
  CHECK(cudaMalloc((void**)&sums, nb * 2 * sizeof(double)));
  CHECK(cudaMalloc((void**)&en, nb * 4 * sizeof(double)));
  CHECK(cudaMemcpy(x, hx, fbytes, H2D));
  memset(hx, 0, fbytes);
  CHECK(cudaMemcpy(v, hx, fbytes, H2D));                                  /* zero momenta */

  CHECK(l2b_su3_plaq_sums(x, sums, nb, dims, L2B_F64, ws, wsb, NULL));
  CHECK(l2b_su3_hmc_trajectory(x, v, beta, 0.1, nlf, xo, vo, en, nb, dims, L2B_F64, ws, wsb, NULL));
  CHECK(cudaDeviceSynchronize());
  double hs[6], he[12];
  CHECK(cudaMemcpy(hs, sums, sizeof hs, D2H));
  CHECK(cudaMemcpy(he, en, sizeof he, D2H));
  CHECK(cudaMemcpy(hx, xo, fbytes, D2H));
  int bad = 0;
  for (int b = 0; b < nb; ++b) {
    if (fabs(hs[2 * b] - 18.0 * V) > 1e-9 || fabs(hs[2 * b + 1]) > 1e-9) ++bad;
    const double s0 = -6.0 * beta * V;       /* S = -(beta/3) * 18 V */
    if (fabs(he[4 * b + 1] - s0) > 1e-8 || fabs(he[4 * b + 3] - s0) > 1e-8) ++bad;
    if (fabs(he[4 * b + 0] - he[4 * b + 2]) > 1e-9) ++bad;   /* no force: kinetic energy unchanged */
  }
  for (size_t l = 0; l < nlinks && !bad; ++l)
    for (int e = 0; e < 18; ++e) {
      const double want = (e == 0 || e == 8 || e == 16) ? 1.0 : 0.0;
      if (fabs(hx[l * 18 + e] - want) > 1e-12) { ++bad; break; }
    }
  printf("cold start: sum Re tr P = %.1f (want %.1f), S = %.3f, launches = %llu -> %s\n", hs[0], 18.0 * V, he[1],
         (unsigned long long)l2b_launch_count(), bad ? "FAILED" : "ok");
  cudaFree(x); cudaFree(v); cudaFree(xo); cudaFree(vo); cudaFree(ws); cudaFree(sums); cudaFree(en);
  free(hx);
  return bad ? 1 : 0;
}
