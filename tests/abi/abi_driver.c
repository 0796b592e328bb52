/* tests/abi/abi_driver.c -- drives libastr_gpu.so from plain C through include/astr_gpu.h only,
 * the way the Fortran bind(C) shim does (fortran/astr_gpu_mod.F90): fills astr_cfg, hands over
 * Fortran-layout host arrays, runs RK stages, reads the state back.  No Python, no torch.
 *
 *   abi_driver <n> <nsteps> <out.bin>
 * Periodic Taylor-Green block n^3 (src/initialisation.F90:621-702, gridcube
 * src/gridgeneration.F90:233-265), nsteps RK3 steps; writes q(-hm:n+hm,..,5) to out.bin.
 * tests/test_gpu_abi_c.py compares that file with the oracle. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "astr_gpu.h"

#define HM ASTR_GPU_HM
#define CHECK(call)                                                            \
  do {                                                                         \
    if ((call) != 0) {                                                         \
      fprintf(stderr, "abi_driver: %s failed: %s\n", #call, astr_gpu_last_error()); \
      return 2;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: abi_driver n nsteps out.bin\n"); return 1; }
  const int n = atoi(argv[1]), nsteps = atoi(argv[2]);
  const double gamma = 1.4, mach = 0.1, reynolds = 1600.0, prandtl = 0.72, ref_tem = 273.15;
  const double pi = 4.0 * atan(1.0);

  astr_cfg c;
  memset(&c, 0, sizeof c);
  if (astr_gpu_sizeof_cfg() != (int)sizeof c) { fprintf(stderr, "astr_cfg layout mismatch\n"); return 1; }
  c.abi_version = ASTR_GPU_ABI_VERSION; c.device = 0;
  c.im = c.jm = c.km = n; c.ia = c.ja = c.ka = n;
  c.hm = HM; c.numq = ASTR_GPU_NUMQ; c.ndims = 3;
  for (int d = 0; d < 3; ++d) { c.npdc[d] = 3; c.lhomo[d] = 1; c.rank[d] = 0; c.size[d] = 1; }
  c.is = c.js = c.ks = 0; c.ie = c.je = c.ke = n;
  for (int s = 0; s < 6; ++s) { c.nbr[s] = -1; c.bctype[s] = 1; }
  c.conschm = c.difschm = 643; c.scheme_compact = 1; c.rkscheme = 3;
  c.lfilter = 1; c.diffterm = 1; c.nondimen = 1; c.flowtype = 0;
  c.alfa_filter = 0.49;
  c.uinf = 1.0; c.vinf = 0.0; c.winf = 0.0; c.roinf = 1.0;
  c.reynolds = reynolds; c.mach = mach; c.prandtl = prandtl; c.gamma = gamma; c.ref_tem = ref_tem;
  /* refcal, src/solver.F90:104-126 */
  c.const1 = 1.0 / (gamma * (gamma - 1.0) * mach * mach);
  c.const2 = gamma * mach * mach;
  c.const3 = (gamma - 1.0) / 3.0 * prandtl * mach * mach;
  c.const4 = (gamma - 1.0) * mach * mach * reynolds * prandtl;
  c.const5 = (gamma - 1.0) * mach * mach;
  c.const6 = 1.0 / (gamma - 1.0);
  c.const7 = (gamma - 1.0) * mach * mach * reynolds * prandtl;
  c.tempconst = 110.3 / ref_tem; c.tempconst1 = 1.0 + c.tempconst;
  c.deltat = 1e-3;
  c.pinf = 1.0 / c.const2; c.bfacmpld = 0.3; c.shkcrt = 0.01;
  CHECK(astr_gpu_init(&c));

  const size_t m = (size_t)n + 1 + 2 * HM, ne = m * m * m;
#define IX(i, j, k) ((size_t)((i) + HM) + m * ((size_t)((j) + HM) + m * (size_t)((k) + HM)))
  double* x = calloc(3 * ne, sizeof(double));
  double* q = calloc(5 * ne, sizeof(double));
  double* rho = calloc(ne, sizeof(double));
  double* vel = calloc(3 * ne, sizeof(double));
  double* prs = calloc(ne, sizeof(double));
  double* tmp = calloc(ne, sizeof(double));
  if (!x || !q || !rho || !vel || !prs || !tmp) return 1;
  const double pinf = 1.0 / c.const2;
  for (int k = 0; k <= n; ++k)
    for (int j = 0; j <= n; ++j)
      for (int i = 0; i <= n; ++i) {
        const size_t p = IX(i, j, k);
        const double X = 2.0 * pi / (double)n * (double)i, Y = 2.0 * pi / (double)n * (double)j,
                     Z = 2.0 * pi / (double)n * (double)k;
        x[p] = X; x[ne + p] = Y; x[2 * ne + p] = Z;
        const double r = 1.0, u = sin(X) * cos(Y) * cos(Z), v = -cos(X) * sin(Y) * cos(Z), w = 0.0;
        const double pp = pinf + r / 16.0 * (cos(2.0 * X) + cos(2.0 * Y)) * (cos(2.0 * Z) + 2.0);
        const double t = pp / r * c.const2;
        rho[p] = r; vel[p] = u; vel[ne + p] = v; vel[2 * ne + p] = w; prs[p] = pp; tmp[p] = t;
        q[p] = r; q[ne + p] = r * u; q[2 * ne + p] = r * v; q[3 * ne + p] = r * w;
        q[4 * ne + p] = r * (t * c.const1 + 0.5 * (u * u + v * v + w * w));
      }
  CHECK(astr_gpu_gridgeom(x));                       /* geomcal */
  CHECK(astr_gpu_upload_state(q, rho, vel, prs, tmp)); /* end of flowinit */
  for (int s = 0; s < nsteps; ++s)
    for (int rk = 1; rk <= 3; ++rk) {                /* time_integration_rk, src/mainloop.F90:396-482 */
      CHECK(astr_gpu_filterq());
      CHECK(astr_gpu_boucon());
      CHECK(astr_gpu_qswap());
      CHECK(astr_gpu_gradcal());
      CHECK(astr_gpu_rhscal());
      CHECK(astr_gpu_rk_update(rk, c.deltat));
      CHECK(astr_gpu_updatefvar());
    }
  CHECK(astr_gpu_download_state(q, NULL, NULL, NULL, NULL));
  long long launches = 0;
  CHECK(astr_gpu_kernel_launches(&launches));
  CHECK(astr_gpu_finalize());
  FILE* f = fopen(argv[3], "wb");
  if (!f || fwrite(q, sizeof(double), 5 * ne, f) != 5 * ne) return 3;
  fclose(f);
  printf("abi_driver ok: n=%d steps=%d kernel launches=%lld\n", n, nsteps, launches);
  free(x); free(q); free(rho); free(vel); free(prs); free(tmp);
  return 0;
}
