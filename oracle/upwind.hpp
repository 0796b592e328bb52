// oracle/upwind.hpp -- TEST INFRASTRUCTURE ONLY (see astr_oracle.hpp).  Included by solver.cpp
// inside namespace astr_oracle, after Block / Case / at() / for_each_pencil are defined.
//
// Restatement of the reference's upwind-biased compact convection path (conschm = '543c'):
//   compact_flux_initiate / flux_compact / compact_flux_rhs   src/flux.F90:32-266
//   mplimiter, MP5                                             src/flux.F90:363-379, :434-501
//   minmod2, minmod4                                           src/commfunc.F90:699-738
//   flux_steger_warming                                        src/riemann.F90:23-161
//   chardecomp                                                 src/solver.F90:1958-2162
//   iwind6                                                     src/solver.F90:1236-1257
//   ducrossensor                                               src/commcal.F90:196-357
//   convrsdcmp                                                 src/solver.F90:1271-1937
// Non-dimensional, no species (numq = 5), crinod all false (lcracon off: the first-order
// fallback of solver.F90:1466-1470 is unreachable).
//
// Parity status: UNPINNED -- the reference ships no stored number for this path.

constexpr double num1d18 = 1.0 / 18.0;
constexpr double num19d18 = 19.0 / 18.0;
constexpr double num5d9 = 5.0 / 9.0;
constexpr double num9d36 = 9.0 / 36.0;
constexpr double num37d60 = 37.0 / 60.0;
constexpr double num2d15 = 2.0 / 15.0;

// src/flux.F90:32-118
static void compact_flux_initiate(CompactScheme& s, int scheme, int ntype, int dim, char wind, double bfacmpld) {
  int i_0 = 0, i_m = 0;
  switch (ntype) {
    case 1: i_0 = -1; i_m = dim + 1; break;
    case 2: i_0 = -2; i_m = dim; break;
    case 3: i_0 = -2; i_m = dim + 1; break;
    default: i_0 = -1; i_m = dim; break;
  }
  s.first_node = i_0; s.last_node = i_m; s.dimension = dim; s.nbctype = ntype;
  const int n = s.size();
  if (scheme != 543) { std::fprintf(stderr, "oracle: only scheme 543 is restated\n"); std::abort(); }
  const double a = (wind == '+') ? 0.5 - num1d6 * bfacmpld : num1d6 + num1d6 * bfacmpld;
  const double c = (wind == '+') ? num1d6 + num1d6 * bfacmpld : 0.5 - num1d6 * bfacmpld;
  s.a.assign(n, a); s.c.assign(n, c);
  s.a[0] = 0.0; s.c[0] = 0.0;
  s.a[n - 1] = 0.0; s.c[n - 1] = 0.0;
  if (ntype == 1 || ntype == 4) { s.a[0] = 2.0; s.c[0] = 2.0; s.a[1] = 0.25; s.c[1] = 0.25; }
  if (ntype == 2 || ntype == 4) { s.a[n - 1] = 2.0; s.c[n - 1] = 2.0; s.a[n - 2] = 0.25; s.c[n - 2] = 0.25; }
  thomas_preprocess(s);
}

// src/flux.F90:163-266 ; f points at node 0, d[row] with row = node - first_node
static void compact_flux_rhs(const CompactScheme& s, char wind, double bfacmpld, const double* f, double* d) {
  const int i_0 = s.first_node, i_m = s.last_node, ntype = s.nbctype;
  auto D = [&](int j) -> double& { return d[j - i_0]; };
  int i_s = i_0 + 1, i_e = i_m - 1;
  if (ntype == 1 || ntype == 4) {
    int j = i_0;
    D(j) = 2.5 * f[j + 1] + 0.5 * f[j + 2];
    j = i_0 + 1;
    D(j) = 0.75 * f[j] + 0.75 * f[j + 1];
    i_s = i_0 + 2;
  } else {
    const int j = i_0;
    const double var1 = f[j] + f[j + 1], var2 = f[j - 1] + f[j + 2], var3 = f[j - 2] + f[j + 3];
    D(j) = num37d60 * var1 - num2d15 * var2 + num1d60 * var3;
  }
  if (ntype == 2 || ntype == 4) {
    int j = i_m - 1;
    D(j) = 0.75 * f[j + 1] + 0.75 * f[j];
    j = i_m;
    D(j) = 2.5 * f[j] + 0.5 * f[j - 1];
    i_e = i_m - 2;
  } else {
    const int j = i_m;
    const double var1 = f[j] + f[j + 1], var2 = f[j - 1] + f[j + 2], var3 = f[j - 2] + f[j + 3];
    D(j) = num37d60 * var1 - num2d15 * var2 + num1d60 * var3;
  }
  if (wind == '+') {
    for (int j = i_s; j <= i_e; ++j)
      D(j) = (num1d18 - num1d36 * bfacmpld) * f[j - 1] + (num19d18 - num9d36 * bfacmpld) * f[j] +
             (num5d9 + num9d36 * bfacmpld) * f[j + 1] + num1d36 * bfacmpld * f[j + 2];
  } else {
    for (int j = i_s; j <= i_e; ++j)
      D(j) = (num1d18 - num1d36 * bfacmpld) * f[j + 2] + (num19d18 - num9d36 * bfacmpld) * f[j + 1] +
             (num5d9 + num9d36 * bfacmpld) * f[j] + num1d36 * bfacmpld * f[j - 1];
  }
}

// src/flux.F90:125-151 ; fh points at interface 0: fh[-1..dim] are written
static void flux_compact(const CompactScheme& s, char wind, double bfacmpld, const double* f, double* fh,
                         double* work /*2*size*/) {
  double* d = work;
  double* xx = work + s.size();
  compact_flux_rhs(s, wind, bfacmpld, f, d);
  thomas_solve(s, d, xx);
  for (int l = -1; l <= s.dimension; ++l) fh[l] = xx[l - s.first_node];
}

// src/commfunc.F90:699-738
static inline double minmod2(double v1, double v2) {
  if (v1 > 0.0 && v2 > 0.0) return std::min(std::fabs(v1), std::fabs(v2));
  if (v1 < 0.0 && v2 < 0.0) return -1.0 * std::min(std::fabs(v1), std::fabs(v2));
  return 0.0;
}
static inline double minmod4(double v1, double v2, double v3, double v4) {
  if (v1 > 0.0 && v2 > 0.0 && v3 > 0.0 && v4 > 0.0)
    return std::min(std::min(std::fabs(v1), std::fabs(v2)), std::min(std::fabs(v3), std::fabs(v4)));
  if (v1 < 0.0 && v2 < 0.0 && v3 < 0.0 && v4 < 0.0)
    return -1.0 * std::min(std::min(std::fabs(v1), std::fabs(v2)), std::min(std::fabs(v3), std::fabs(v4)));
  return 0.0;
}

// src/flux.F90:434-501  MP5(u(1:5), ul, discont) ; u[0..4] = u(1..5)
static double mp5(const double* u, double ul, bool discont) {
  const double ulinear = ul;
  double var1 = u[3] - u[2];
  double var2 = 4.0 * (u[2] - u[1]);
  const double uMP = u[2] + minmod2(var1, var2);
  var1 = (ulinear - u[2]) * (ulinear - uMP);
  if (discont && var1 >= 1.e-10) {
    const double dm1 = u[0] - 2.0 * u[1] + u[2];
    const double d0 = u[1] - 2.0 * u[2] + u[3];
    const double d1 = u[2] - 2.0 * u[3] + u[4];
    const double dhm1 = minmod4(4.0 * dm1 - d0, 4.0 * d0 - dm1, dm1, d0);
    const double dh0 = minmod4(4.0 * d0 - d1, 4.0 * d1 - d0, d0, d1);
    const double uUL = u[2] + 4.0 * (u[2] - u[1]);
    const double uAV = 0.5 * (u[2] + u[3]);
    const double uMD = uAV - 0.5 * dh0;
    const double uLC = u[2] + 0.5 * (u[2] - u[1]) + 1.333333333333333 * dhm1;
    var1 = std::min(std::min(u[2], u[3]), uMD);
    var2 = std::min(std::min(u[2], uUL), uLC);
    const double uMIN = std::max(var1, var2);
    var1 = std::max(std::max(u[2], u[3]), uMD);
    var2 = std::max(std::max(u[2], uUL), uLC);
    const double uMAX = std::min(var1, var2);
    var1 = uMIN - ulinear;
    var2 = uMAX - ulinear;
    return ulinear + minmod2(var1, var2);
  }
  return ulinear;
}

// src/flux.F90:363-379
static inline double mplimiter(const double* f, double fl, bool shock, int inode, int dim, int ntype) {
  if ((ntype == 1 && inode == 0) || (ntype == 2 && inode == dim - 1) || (ntype == 1 && inode == 1) ||
      (ntype == 2 && inode == dim - 2))
    return fl;
  return mp5(f, fl, shock);
}

// src/solver.F90:1236-1257
static inline int iwind6(int i, int n /*1..5*/, int imin, int imax, char dir) {
  int w = (dir == '+') ? i + n - 3 : i + 4 - n;
  if (w < imin) w = imin;
  if (w > imax) w = imax;
  return w;
}

// src/riemann.F90:23-161 at one node (nondimen: sos = sqrt(T)/Mach, fludyna.F90:851)
static void steger_warming_node(const Thermo& th, double rho, const double* vel, double prs, double tmp,
                                const double* q, const double* dxi, double jacob, double* fplus, double* fmius) {
  const double eps = 0.04;
  const double gamma = th.gamma;
  const double uu = dxi[0] * vel[0] + dxi[1] * vel[1] + dxi[2] * vel[2];
  const double var0 = 1.0 / std::sqrt(dxi[0] * dxi[0] + dxi[1] * dxi[1] + dxi[2] * dxi[2]);
  const double gpd[3] = {dxi[0] * var0, dxi[1] * var0, dxi[2] * var0};
  const double gm2 = 0.5 / gamma;
  const double css = th.sos(tmp);
  const double csa = css / var0;
  const double lmach = uu / csa;
  const double lmda[5] = {uu, uu, uu, uu + csa, uu - csa};
  double lmdap[5], lmdam[5];
  for (int n = 0; n < 5; ++n) {
    lmdap[n] = 0.5 * (lmda[n] + std::sqrt(lmda[n] * lmda[n] + eps * eps));
    lmdam[n] = lmda[n] - lmdap[n];
  }
  if (lmach >= 1.0) {
    fplus[0] = jacob * q[0] * uu;
    fplus[1] = jacob * (q[1] * uu + dxi[0] * prs);
    fplus[2] = jacob * (q[2] * uu + dxi[1] * prs);
    fplus[3] = jacob * (q[3] * uu + dxi[2] * prs);
    fplus[4] = jacob * (q[4] + prs) * uu;
    for (int n = 0; n < 5; ++n) fmius[n] = 0.0;
  } else if (lmach <= -1.0) {
    for (int n = 0; n < 5; ++n) fplus[n] = 0.0;
    fmius[0] = jacob * q[0] * uu;
    fmius[1] = jacob * (q[1] * uu + dxi[0] * prs);
    fmius[2] = jacob * (q[2] * uu + dxi[1] * prs);
    fmius[3] = jacob * (q[3] * uu + dxi[2] * prs);
    fmius[4] = jacob * (q[4] + prs) * uu;
  } else {
    const double fhi = 0.5 * (gamma - 1.0) * (vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
    const double jro = jacob * rho;
    double var1 = lmdap[0];
    double var2 = lmdap[3] - lmdap[4];
    double var3 = 2.0 * lmdap[0] - lmdap[3] - lmdap[4];
    double var4 = var1 - var3 * gm2;
    fplus[0] = jro * var4;
    fplus[1] = jro * (var4 * vel[0] + var2 * css * gpd[0] * gm2);
    fplus[2] = jro * (var4 * vel[1] + var2 * css * gpd[1] * gm2);
    fplus[3] = jro * (var4 * vel[2] + var2 * css * gpd[2] * gm2);
    fplus[4] = jacob * (var1 * q[4] + rho * (var2 * uu * var0 * css * gm2 - var3 * (fhi + css * css) * gm2 / (gamma - 1.0)));
    var1 = lmdam[0];
    var2 = lmdam[3] - lmdam[4];
    var3 = 2.0 * lmdam[0] - lmdam[3] - lmdam[4];
    var4 = var1 - var3 * gm2;
    fmius[0] = jro * var4;
    fmius[1] = jro * (var4 * vel[0] + var2 * css * gpd[0] * gm2);
    fmius[2] = jro * (var4 * vel[1] + var2 * css * gpd[1] * gm2);
    fmius[3] = jro * (var4 * vel[2] + var2 * css * gpd[2] * gm2);
    fmius[4] = jacob * (var1 * q[4] + rho * (var2 * uu * var0 * css * gm2 - var3 * (fhi + css * css) * gm2 / (gamma - 1.0)));
  }
}

// src/solver.F90:1958-2162 (no COMB).  REV/LEV are [row][col], 0-based.  Returns false where
// the reference stops (' !! ERROR 1 @ chardecomp': degenerate metric normal).
static bool chardecomp(double gamma, double ro_l, double p_l, double E_l, const double* vel_l, const double* ddi_l,
                       double ro_r, double p_r, double E_r, const double* vel_r, const double* ddi_r,
                       double REV[5][5], double LEV[5][5]) {
  const double rero = 1.e-12;
  const double WRoe = std::sqrt(ro_l) / (std::sqrt(ro_l) + std::sqrt(ro_r));
  const double WRoe1 = 1.0 - WRoe;
  const double u1Roe = WRoe * vel_l[0] + WRoe1 * vel_r[0];
  const double u2Roe = WRoe * vel_l[1] + WRoe1 * vel_r[1];
  const double u3Roe = WRoe * vel_l[2] + WRoe1 * vel_r[2];
  const double KRoe = 0.5 * (u1Roe * u1Roe + u2Roe * u2Roe + u3Roe * u3Roe);
  const double HL = (E_l + p_l) / ro_l;
  const double HR = (E_r + p_r) / ro_r;
  const double HRoe = WRoe * HL + WRoe1 * HR;
  const double CssRoe = std::sqrt((gamma - 1.0) * (HRoe - KRoe));
  const double rcs = 1.0 / CssRoe;
  const double var1 = 0.5 * (ddi_l[0] + ddi_r[0]);
  const double var2 = 0.5 * (ddi_l[1] + ddi_r[1]);
  const double var3 = 0.5 * (ddi_l[2] + ddi_r[2]);
  const double var4 = 1.0 / std::sqrt(var1 * var1 + var2 * var2 + var3 * var3);
  const double gpd[3] = {var1 * var4, var2 * var4, var3 * var4};
  const double ugp = u1Roe * gpd[0] + u2Roe * gpd[1] + u3Roe * gpd[2];
  const double b1 = (gamma - 1.0) / (CssRoe * CssRoe);
  const double b2 = 1.0 + 2.0 * b1 * KRoe - b1 * HRoe;
  // (row, col) 1-based accessors in the reference's notation
#define LE(m, n) LEV[(m) - 1][(n) - 1]
#define RE(m, n) REV[(m) - 1][(n) - 1]
  LE(1, 1) = 0.5 * (b2 + ugp * rcs);
  LE(1, 2) = -0.5 * (b1 * u1Roe + gpd[0] * rcs);
  LE(1, 3) = -0.5 * (b1 * u2Roe + gpd[1] * rcs);
  LE(1, 4) = -0.5 * (b1 * u3Roe + gpd[2] * rcs);
  LE(1, 5) = 0.5 * b1;
  RE(1, 1) = 1.0;
  RE(2, 1) = u1Roe - CssRoe * gpd[0];
  RE(3, 1) = u2Roe - CssRoe * gpd[1];
  RE(4, 1) = u3Roe - CssRoe * gpd[2];
  RE(5, 1) = HRoe - ugp * CssRoe;
  LE(2, 1) = 1.0 - b2;
  LE(2, 2) = b1 * u1Roe;
  LE(2, 3) = b1 * u2Roe;
  LE(2, 4) = b1 * u3Roe;
  LE(2, 5) = -b1;
  RE(1, 2) = 1.0;
  RE(2, 2) = u1Roe;
  RE(3, 2) = u2Roe;
  RE(4, 2) = u3Roe;
  RE(5, 2) = HRoe - 1.0 / b1;
  if (std::fabs(var1) > rero) {
    const double rgp = 1.0 / gpd[0];
    LE(3, 1) = (ugp * gpd[1] - u2Roe) * rgp;
    LE(3, 2) = -gpd[1];
    LE(3, 3) = (1.0 - gpd[1] * gpd[1]) * rgp;
    LE(3, 4) = -gpd[1] * gpd[2] * rgp;
    LE(3, 5) = 0.0;
    LE(4, 1) = (ugp * gpd[2] - u3Roe) * rgp;
    LE(4, 2) = -gpd[2];
    LE(4, 3) = -gpd[1] * gpd[2] * rgp;
    LE(4, 4) = (1.0 - gpd[2] * gpd[2]) * rgp;
    LE(4, 5) = 0.0;
    RE(1, 3) = 0.0; RE(2, 3) = -gpd[1]; RE(3, 3) = gpd[0]; RE(4, 3) = 0.0;
    RE(5, 3) = u2Roe * gpd[0] - u1Roe * gpd[1];
    RE(1, 4) = 0.0; RE(2, 4) = -gpd[2]; RE(3, 4) = 0.0; RE(4, 4) = gpd[0];
    RE(5, 4) = u3Roe * gpd[0] - u1Roe * gpd[2];
  } else if (std::fabs(var2) > rero) {
    const double rgp = 1.0 / gpd[1];
    LE(3, 1) = (ugp * gpd[0] - u1Roe) * rgp;
    LE(3, 2) = (1.0 - gpd[0] * gpd[0]) * rgp;
    LE(3, 3) = -gpd[0];
    LE(3, 4) = -gpd[0] * gpd[2] * rgp;
    LE(3, 5) = 0.0;
    LE(4, 1) = (ugp * gpd[2] - u3Roe) * rgp;
    LE(4, 2) = -gpd[0] * gpd[2] * rgp;
    LE(4, 3) = -gpd[2];
    LE(4, 4) = (1.0 - gpd[2] * gpd[2]) * rgp;
    LE(4, 5) = 0.0;
    RE(1, 3) = 0.0; RE(2, 3) = gpd[1]; RE(3, 3) = -gpd[0]; RE(4, 3) = 0.0;
    RE(5, 3) = u1Roe * gpd[1] - u2Roe * gpd[0];
    RE(1, 4) = 0.0; RE(2, 4) = 0.0; RE(3, 4) = -gpd[2]; RE(4, 4) = gpd[1];
    RE(5, 4) = u3Roe * gpd[1] - u2Roe * gpd[2];
  } else if (std::fabs(var3) > rero) {
    const double rgp = 1.0 / gpd[2];
    LE(3, 1) = (ugp * gpd[0] - u1Roe) * rgp;
    LE(3, 2) = (1.0 - gpd[0] * gpd[0]) * rgp;
    LE(3, 3) = -gpd[0] * gpd[1] * rgp;
    LE(3, 4) = -gpd[0];
    LE(3, 5) = 0.0;
    LE(4, 1) = (ugp * gpd[1] - u2Roe) * rgp;
    LE(4, 2) = -gpd[0] * gpd[1] * rgp;
    LE(4, 3) = (1.0 - gpd[1] * gpd[1]) * rgp;
    LE(4, 4) = -gpd[1];
    LE(4, 5) = 0.0;
    RE(1, 3) = 0.0; RE(2, 3) = gpd[2]; RE(3, 3) = 0.0; RE(4, 3) = -gpd[0];
    RE(5, 3) = u1Roe * gpd[2] - u3Roe * gpd[0];
    RE(1, 4) = 0.0; RE(2, 4) = 0.0; RE(3, 4) = gpd[2]; RE(4, 4) = -gpd[1];
    RE(5, 4) = u2Roe * gpd[2] - u3Roe * gpd[1];
  } else {
    return false;
  }
  LE(5, 1) = 0.5 * (b2 - ugp * rcs);
  LE(5, 2) = -0.5 * (b1 * u1Roe - gpd[0] * rcs);
  LE(5, 3) = -0.5 * (b1 * u2Roe - gpd[1] * rcs);
  LE(5, 4) = -0.5 * (b1 * u3Roe - gpd[2] * rcs);
  LE(5, 5) = 0.5 * b1;
  RE(1, 5) = 1.0;
  RE(2, 5) = u1Roe + CssRoe * gpd[0];
  RE(3, 5) = u2Roe + CssRoe * gpd[1];
  RE(4, 5) = u3Roe + CssRoe * gpd[2];
  RE(5, 5) = HRoe + ugp * CssRoe;
#undef LE
#undef RE
  return true;
}

// hdiss of the interface i+1/2 (src/solver.F90:1456-1462, :660-666): a critical node (crash control) on either side
static inline bool crinod_interface(const Block& b, int d, int i, int dm, int p1, int p2) {
  if (b.crinod.v.empty()) return false;
  if (i < 0) return at(b.crinod, d, i + 1, p1, p2) != 0.0;
  if (i + 1 > dm) return at(b.crinod, d, i, p1, p2) != 0.0;
  return at(b.crinod, d, i, p1, p2) != 0.0 || at(b.crinod, d, i + 1, p1, p2) != 0.0;
}

// src/commcal.F90:196-357 ducrossensor -> b.ssf, b.lshock (0/1 as doubles)
static void ducrossensor(Case& c) {
  for (Block& b : c.blk) {
    if (b.ssf.v.empty()) { b.ssf.alloc(b.im, b.jm, b.km); b.lshock.alloc(b.im, b.jm, b.km); }
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) {
          auto dv = [&](int m, int n) { return b.dvel[m - 1][n - 1](i, j, k); };
          const double s = dv(1, 1) + dv(2, 2) + dv(3, 3);
          const double div2 = s * s;
          const double vortx = dv(3, 2) - dv(2, 3), vorty = dv(1, 3) - dv(3, 1), vortz = dv(2, 1) - dv(1, 2);
          const double vort = vortx * vortx + vorty * vorty + vortz * vortz;
          int ip1 = i + 1, im1 = i - 1, jp1 = j + 1, jm1 = j - 1, kp1 = k + 1, km1 = k - 1;
          if (b.npdc[0] == 1 && im1 < 0) im1 = 0;
          if (b.npdc[1] == 1 && jm1 < 0) jm1 = 0;
          if (b.npdc[2] == 1 && km1 < 0) km1 = 0;
          if (b.npdc[0] == 2 && ip1 > b.im) ip1 = b.im;
          if (b.npdc[1] == 2 && jp1 > b.jm) jp1 = b.jm;
          if (b.npdc[2] == 2 && kp1 > b.km) kp1 = b.km;
          const double p0 = b.prs(i, j, k);
          const double dpdi = std::fabs(b.prs(ip1, j, k) - 2.0 * p0 + b.prs(im1, j, k)) /
                              (b.prs(ip1, j, k) + 2.0 * p0 + b.prs(im1, j, k));
          const double dpdj = std::fabs(b.prs(i, jp1, k) - 2.0 * p0 + b.prs(i, jm1, k)) /
                              (b.prs(i, jp1, k) + 2.0 * p0 + b.prs(i, jm1, k));
          const double dpdk = std::fabs(b.prs(i, j, kp1) - 2.0 * p0 + b.prs(i, j, km1)) /
                              (b.prs(i, j, kp1) + 2.0 * p0 + b.prs(i, j, km1));
          b.ssf(i, j, k) = div2 / (div2 + vort + 1.e-30) * std::max(std::max(dpdi, dpdj), dpdk);
        }
  }
  Getter gs = [](Block& b) { return FieldList{&b.ssf}; };
  dataswap(c, gs);
  for (Block& b : c.blk) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) {
          double m = 0.0;
          for (int i1 = -hm + 1; i1 <= hm; ++i1) {
            int ii = i + i1;
            if (b.npdc[0] == 1 && ii < 0) ii = 0;
            if (b.npdc[0] == 2 && ii > b.im) ii = b.im;
            m = std::max(m, b.ssf(ii, j, k));
          }
          for (int j1 = -hm + 1; j1 <= hm; ++j1) {
            int jj = j + j1;
            if (b.npdc[1] == 1 && jj < 0) jj = 0;
            if (b.npdc[1] == 2 && jj > b.jm) jj = b.jm;
            m = std::max(m, b.ssf(i, jj, k));
          }
          for (int k1 = -hm + 1; k1 <= hm; ++k1) {
            int kk = k + k1;
            if (b.npdc[2] == 1 && kk < 0) kk = 0;
            if (b.npdc[2] == 2 && kk > b.km) kk = b.km;
            m = std::max(m, b.ssf(i, j, kk));
          }
          b.lshock(i, j, k) = (m > c.shkcrt) ? 1.0 : 0.0;
        }
  }
}

// src/solver.F90:1271-1937 convrsdcmp, the three directions written once over `d`
static int convrsdcmp(Case& c) {
  const int md = std::max(c.ia, std::max(c.ja, c.ka));
  const double bf = c.bfacmpld;
  int bad = 0;
  for (Block& b : c.blk) {
    const bool sson = !b.lshock.v.empty();   // allocated(lshock)
    for (int d = 0; d < c.ndims(); ++d) {   // solver.F90:1719 `if(ndims==2) return` after j
      const int dm = b.dim(d), nt = b.npdc[d];
      CompactScheme uw, dw;
      compact_flux_initiate(uw, 543, nt, dm, '+', bf);   // comsolver.F90:103-108
      compact_flux_initiate(dw, 543, nt, dm, '-', bf);
      int lss, lee;                                       // :1313-1327
      if (nt == 1) { lss = 0; lee = dm + hm; }
      else if (nt == 2) { lss = -hm; lee = dm; }
      else if (nt == 3) { lss = -hm; lee = dm + hm; }
      else { lss = 0; lee = dm; }
      int o1, o2;
      if (d == 0) { o1 = 1; o2 = 2; } else if (d == 1) { o1 = 0; o2 = 2; } else { o1 = 0; o2 = 1; }
      const int np = md + 1 + 2 * hm;
      for_each_pencil(b, d, [&](int p1, int p2) {
        if (p1 < b.s[o1] || p1 > b.e[o1] || p2 < b.s[o2] || p2 > b.e[o2]) return;
        std::vector<double> fsw(10 * np, 0.0), fhc(10 * (md + 2), 0.0), Fh(5 * (md + 2), 0.0), work(2 * (md + 8));
        double* fswp[5]; double* fswm[5]; double* fhcp[5]; double* fhcm[5]; double* fh[5];
        for (int n = 0; n < 5; ++n) {
          fswp[n] = fsw.data() + n * np + hm;
          fswm[n] = fsw.data() + (5 + n) * np + hm;
          fhcp[n] = fhc.data() + n * (md + 2) + 1;
          fhcm[n] = fhc.data() + (5 + n) * (md + 2) + 1;
          fh[n] = Fh.data() + n * (md + 2) + 1;
        }
        for (int l = lss; l <= lee; ++l) {
          const double vel[3] = {at(b.vel[0], d, l, p1, p2), at(b.vel[1], d, l, p1, p2), at(b.vel[2], d, l, p1, p2)};
          const double q[5] = {at(b.q[0], d, l, p1, p2), at(b.q[1], d, l, p1, p2), at(b.q[2], d, l, p1, p2),
                               at(b.q[3], d, l, p1, p2), at(b.q[4], d, l, p1, p2)};
          const double dxi[3] = {at(b.dxi[d][0], d, l, p1, p2), at(b.dxi[d][1], d, l, p1, p2), at(b.dxi[d][2], d, l, p1, p2)};
          double fp[5], fm[5];
          steger_warming_node(c.th, at(b.rho, d, l, p1, p2), vel, at(b.prs, d, l, p1, p2), at(b.tmp, d, l, p1, p2), q, dxi,
                              at(b.jacob, d, l, p1, p2), fp, fm);
          for (int n = 0; n < 5; ++n) { fswp[n][l] = fp[n]; fswm[n][l] = fm[n]; }
        }
        for (int n = 0; n < 5; ++n) {
          flux_compact(uw, '+', bf, fswp[n], fhcp[n], work.data());
          flux_compact(dw, '-', bf, fswm[n], fhcm[n], work.data());
        }
        for (int i = b.s[d] - 1; i <= b.e[d]; ++i) {
          double flcp[5][5], flcm[5][5], fhcpc[5], fhcmc[5], Fhc[5];
          double REV[5][5], LEV[5][5];
          if (c.lchardecomp) {
            const double vl[3] = {at(b.vel[0], d, i, p1, p2), at(b.vel[1], d, i, p1, p2), at(b.vel[2], d, i, p1, p2)};
            const double vr[3] = {at(b.vel[0], d, i + 1, p1, p2), at(b.vel[1], d, i + 1, p1, p2), at(b.vel[2], d, i + 1, p1, p2)};
            const double dl[3] = {at(b.dxi[d][0], d, i, p1, p2), at(b.dxi[d][1], d, i, p1, p2), at(b.dxi[d][2], d, i, p1, p2)};
            const double dr[3] = {at(b.dxi[d][0], d, i + 1, p1, p2), at(b.dxi[d][1], d, i + 1, p1, p2),
                                  at(b.dxi[d][2], d, i + 1, p1, p2)};
            if (!chardecomp(c.th.gamma, at(b.rho, d, i, p1, p2), at(b.prs, d, i, p1, p2), at(b.q[4], d, i, p1, p2), vl, dl,
                            at(b.rho, d, i + 1, p1, p2), at(b.prs, d, i + 1, p1, p2), at(b.q[4], d, i + 1, p1, p2), vr, dr,
                            REV, LEV)) {
#pragma omp atomic
              bad += 1;
              continue;
            }
            for (int m = 0; m < 5; ++m) {
              for (int n = 1; n <= 5; ++n) {
                int nwd = iwind6(i, n, lss, lee, '+');
                flcp[m][n - 1] = LEV[m][0] * fswp[0][nwd] + LEV[m][1] * fswp[1][nwd] + LEV[m][2] * fswp[2][nwd] +
                                 LEV[m][3] * fswp[3][nwd] + LEV[m][4] * fswp[4][nwd];
                nwd = iwind6(i, n, lss, lee, '-');
                flcm[m][n - 1] = LEV[m][0] * fswm[0][nwd] + LEV[m][1] * fswm[1][nwd] + LEV[m][2] * fswm[2][nwd] +
                                 LEV[m][3] * fswm[3][nwd] + LEV[m][4] * fswm[4][nwd];
              }
              fhcpc[m] = LEV[m][0] * fhcp[0][i] + LEV[m][1] * fhcp[1][i] + LEV[m][2] * fhcp[2][i] + LEV[m][3] * fhcp[3][i] +
                         LEV[m][4] * fhcp[4][i];
              fhcmc[m] = LEV[m][0] * fhcm[0][i] + LEV[m][1] * fhcm[1][i] + LEV[m][2] * fhcm[2][i] + LEV[m][3] * fhcm[3][i] +
                         LEV[m][4] * fhcm[4][i];
            }
          } else {
            for (int n = 1; n <= 5; ++n) {
              int nwd = iwind6(i, n, lss, lee, '+');
              for (int m = 0; m < 5; ++m) flcp[m][n - 1] = fswp[m][nwd];
              nwd = iwind6(i, n, lss, lee, '-');
              for (int m = 0; m < 5; ++m) flcm[m][n - 1] = fswm[m][nwd];
            }
            for (int m = 0; m < 5; ++m) { fhcpc[m] = fhcp[m][i]; fhcmc[m] = fhcm[m][i]; }
          }
          bool lsh = true;                                  // :1437-1447
          if (sson) {
            if (i < 0) lsh = at(b.lshock, d, i + 1, p1, p2) != 0.0;
            else if (i + 1 > dm) lsh = at(b.lshock, d, i, p1, p2) != 0.0;
            else lsh = at(b.lshock, d, i, p1, p2) != 0.0 || at(b.lshock, d, i + 1, p1, p2) != 0.0;
          }
          const bool hdiss = crinod_interface(b, d, i, dm, p1, p2);       // :1456-1462
          for (int m = 0; m < 5; ++m) {
            // :1466-1481: at a critical node the interface value is the (projected) split flux of one node
            const double var1 = hdiss ? flcp[m][3] : mplimiter(flcp[m], fhcpc[m], lsh, i, dm, nt);
            const double var2 = hdiss ? flcm[m][3] : mplimiter(flcm[m], fhcmc[m], lsh, i, dm, nt);
            Fhc[m] = var1 + var2;
          }
          if (c.lchardecomp) {
            for (int m = 0; m < 5; ++m)
              fh[m][i] = REV[m][0] * Fhc[0] + REV[m][1] * Fhc[1] + REV[m][2] * Fhc[2] + REV[m][3] * Fhc[3] + REV[m][4] * Fhc[4];
          } else {
            for (int m = 0; m < 5; ++m) fh[m][i] = Fhc[m];
          }
        }
        for (int i = b.s[d]; i <= b.e[d]; ++i)
          for (int m = 0; m < 5; ++m) {
            double& r = at(b.qrhs[m], d, i, p1, p2);
            r = r + fh[m][i] - fh[m][i - 1];
          }
      });
    }
  }
  return bad;
}
