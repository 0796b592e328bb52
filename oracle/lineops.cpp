// oracle/lineops.cpp -- TEST INFRASTRUCTURE ONLY (see astr_oracle.hpp).
// 1-D line operators of the reference: Thomas solver, 6th-order compact first
// derivative, explicit 6th-order derivative, 10th-order compact filter.
#include "astr_oracle.hpp"
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace astr_oracle {

double Thermo::std_sqrt(double v) { return std::sqrt(v); }

// src/commfunc.F90:752-774  tridiagonal_thomas_proprocess
// ac(1,1)=c(1); ac(2,1), ac(3,1) are never set nor read by the solver.
void thomas_preprocess(CompactScheme& s) {
  const int n = s.size();
  s.ac1.assign(n, 0.0);
  s.ac2.assign(n, 0.0);
  s.ac3.assign(n, 0.0);
  s.ac1[0] = s.c[0];
  for (int i = 1; i < n; ++i) {
    s.ac1[i] = s.c[i] / (1.0 - s.a[i] * s.ac1[i - 1]);
    s.ac2[i] = 1.0 / (1.0 - s.a[i] * s.ac1[i - 1]);
    s.ac3[i] = s.a[i] / (1.0 - s.a[i] * s.ac1[i - 1]);
  }
}

// src/commfunc.F90:790-813  tridiagonal_thomas_solver
void thomas_solve(const CompactScheme& s, double* d, double* x) {
  const int n = s.size();
  for (int i = 1; i < n; ++i) d[i] = d[i] * s.ac2[i] - d[i - 1] * s.ac3[i];
  x[n - 1] = d[n - 1];
  for (int i = n - 2; i >= 0; --i) x[i] = d[i] - s.ac1[i] * x[i + 1];
}

// src/derivative.F90:63-158  fd_scheme_initiate
void fd_scheme_initiate(CompactScheme& s, int nscheme, char kind, int ntype, int dim) {
  int i_0, i_m;
  switch (ntype) {  // :74-93
    case 1: i_0 = 0;  i_m = dim + 1; break;
    case 2: i_0 = -1; i_m = dim;     break;
    case 3: i_0 = -1; i_m = dim + 1; break;
    case 4: i_0 = 0;  i_m = dim;     break;
    default: std::fprintf(stderr, "oracle: bad ntype %d\n", ntype); std::abort();
  }
  s.first_node = i_0; s.last_node = i_m; s.dimension = dim; s.nbctype = ntype;
  const int n = s.size();
  s.a.assign(n, 0.0); s.c.assign(n, 0.0);
  if (kind != 'c') { s.ac1.clear(); s.ac2.clear(); s.ac3.clear(); return; }  // :104
  if (nscheme / 100 != 6 || nscheme != 643) {
    std::fprintf(stderr, "oracle: scheme %d%c not defined (reference stops too)\n", nscheme, kind);
    std::abort();
  }
  auto A = [&](int node) -> double& { return s.a[node - i_0]; };
  auto C = [&](int node) -> double& { return s.c[node - i_0]; };
  for (int i = 0; i < n; ++i) { s.a[i] = num1d3; s.c[i] = num1d3; }  // :108-109
  A(i_0) = 0.0; C(i_0) = 0.0;                                         // :112-113
  A(i_m) = 0.0; C(i_m) = 0.0;
  if (ntype == 1 || ntype == 4) {  // :131-134
    A(i_0) = 2.0;      C(i_0) = 2.0;
    A(i_0 + 1) = 0.25; C(i_0 + 1) = 0.25;
  }
  if (ntype == 2 || ntype == 4) {  // :136-139
    A(i_m) = 2.0;      C(i_m) = 2.0;
    A(i_m - 1) = 0.25; C(i_m - 1) = 0.25;
  }
  thomas_preprocess(s);  // :154
}

// src/derivative.F90:210-306  compact_fd_rhs.  d is indexed [node-first_node].
void compact_fd_rhs(const CompactScheme& s, const double* f, double* dd) {
  const int i_0 = s.first_node, i_m = s.last_node, ntype = s.nbctype;
  auto d = [&](int node) -> double& { return dd[node - i_0]; };
  int i_s, i_e, j;
  double var1, var2, var3;
  if (ntype == 1 || ntype == 4) {  // physical boundary :230-248
    i_s = i_0 + 2;
    j = i_0 + 1;
    var1 = f[j + 1] - f[j - 1];
    d(j) = 0.75 * var1;
    j = i_0;
    d(j) = -2.5 * f[j] + 2.0 * f[j + 1] + 0.5 * f[j + 2];
  } else {  // interface :250-260
    i_s = i_0 + 1;
    j = i_0;
    var1 = f[j + 1] - f[j - 1];
    var2 = f[j + 2] - f[j - 2];
    var3 = f[j + 3] - f[j - 3];
    d(j) = 0.75 * var1 - 0.15 * var2 + num1d60 * var3;
  }
  if (ntype == 2 || ntype == 4) {  // :264-281
    i_e = i_m - 2;
    j = i_m - 1;
    var1 = f[j + 1] - f[j - 1];
    d(j) = 0.75 * var1;
    j = i_m;
    d(j) = 2.5 * f[j] - 2.0 * f[j - 1] - 0.5 * f[j - 2];
  } else {  // :283-292
    i_e = i_m - 1;
    j = i_m;
    var1 = f[j + 1] - f[j - 1];
    var2 = f[j + 2] - f[j - 2];
    var3 = f[j + 3] - f[j - 3];
    d(j) = 0.75 * var1 - 0.15 * var2 + num1d60 * var3;
  }
  for (j = i_s; j <= i_e; ++j) {  // :296-304
    var1 = f[j + 1] - f[j - 1];
    var2 = f[j + 2] - f[j - 2];
    d(j) = num7d9 * var1 + num1d36 * var2;
  }
}

// src/derivative.F90:171-198  df_compact
void df_compact(const CompactScheme& s, const double* f, double* df, double* work) {
  const int n = s.size();
  double* d = work;
  double* xx = work + n;
  compact_fd_rhs(s, f, d);
  thomas_solve(s, d, xx);
  for (int i = 0; i <= s.dimension; ++i) df[i] = xx[i - s.first_node];
}

// src/derivative.F90:350-413  diff6ec
void diff6ec(const double* vin, int dim, int ntype, double* vout) {
  auto c6 = [&](int i) {
    return 0.75 * (vin[i + 1] - vin[i - 1]) - 0.15 * (vin[i + 2] - vin[i - 2]) +
           num1d60 * (vin[i + 3] - vin[i - 3]);
  };
  int lo = 0, hi = dim;
  if (ntype == 1 || ntype == 4) {
    vout[0] = -0.5 * vin[2] + 2.0 * vin[1] - 1.5 * vin[0];
    vout[1] = 0.5 * (vin[2] - vin[0]);
    vout[2] = num2d3 * (vin[3] - vin[1]) - num1d12 * (vin[4] - vin[0]);
    lo = 3;
  }
  if (ntype == 2 || ntype == 4) {
    vout[dim - 2] = num2d3 * (vin[dim - 1] - vin[dim - 3]) - num1d12 * (vin[dim] - vin[dim - 4]);
    vout[dim - 1] = 0.5 * (vin[dim] - vin[dim - 2]);
    vout[dim] = 0.5 * vin[dim - 2] - 2.0 * vin[dim - 1] + 1.5 * vin[dim];
    hi = dim - 3;
  }
  for (int i = lo; i <= hi; ++i) vout[i] = c6(i);
}

// src/filter.F90:299-432  filter_coefficient_cal (only the tables the hot path reads)
void filter_coefficient_cal(FilterCoef& fc, double alfa, double beter_halo, double beter_bouond) {
  fc.coef2i[0] = (1.0 + 2.0 * alfa) / 4.0;
  fc.coef2i[1] = (1.0 + 2.0 * alfa) / 4.0;
  fc.coef4i[0] = (5.0 + 6.0 * alfa) / 16.0;
  fc.coef4i[1] = (1.0 + 2.0 * alfa) / 4.0;
  fc.coef4i[2] = (-1.0 + 2.0 * alfa) / 16.0;
  fc.coef6i[0] = (11.0 + 10.0 * alfa) / 32.0;
  fc.coef6i[1] = (15.0 + 34.0 * alfa) / 64.0;
  fc.coef6i[2] = (-3.0 + 6.0 * alfa) / 32.0;
  fc.coef6i[3] = (1.0 - 2.0 * alfa) / 64.0;
  fc.coef8i[0] = (93.0 + 70.0 * alfa) / 256.0;
  fc.coef8i[1] = (7.0 + 18.0 * alfa) / 32.0;
  fc.coef8i[2] = (-7.0 + 14.0 * alfa) / 64.0;
  fc.coef8i[3] = (1.0 - 2.0 * alfa) / 32.0;
  fc.coef8i[4] = (-1.0 + 2.0 * alfa) / 256.0;
  fc.coef10i[0] = (193.0 + 126.0 * alfa) / 512.0;
  fc.coef10i[1] = (105.0 + 302.0 * alfa) / 512.0;
  fc.coef10i[2] = (-15.0 + 30.0 * alfa) / 128.0;
  fc.coef10i[3] = (45.0 - 90.0 * alfa) / 1024.0;
  fc.coef10i[4] = (-5.0 + 10.0 * alfa) / 512.0;
  fc.coef10i[5] = (1.0 - 2.0 * alfa) / 1024.0;
  for (auto& row : fc.coefb) for (double& v : row) v = 0.0;
  // coefb(4,:) is allocated but never assigned in the reference (:304) and never read.
  fc.coefb[3][0] = (1.0 - 2.0 * alfa) / 256.0;
  fc.coefb[3][1] = (-1.0 + 2.0 * alfa) / 32.0;
  fc.coefb[3][2] = (7.0 + 50.0 * alfa) / 64.0;
  fc.coefb[3][3] = (25.0 + 14.0 * alfa) / 32.0;
  fc.coefb[3][4] = (35.0 + 58.0 * alfa) / 128.0;
  fc.coefb[3][5] = (-7.0 + 14.0 * alfa) / 32.0;
  fc.coefb[3][6] = (7.0 - 14.0 * alfa) / 64.0;
  fc.coefb[3][7] = (-1.0 + 2.0 * alfa) / 32.0;
  fc.coefb[3][8] = (1.0 - 2.0 * alfa) / 256.0;
  fc.coefb[2][0] = (-1.0 + 2.0 * alfa) / 64.0;
  fc.coefb[2][1] = (3.0 + 26.0 * alfa) / 32.0;
  fc.coefb[2][2] = (49.0 + 30.0 * alfa) / 64.0;
  fc.coefb[2][3] = (5.0 + 6.0 * alfa) / 16.0;
  fc.coefb[2][4] = (-15.0 + 30.0 * alfa) / 64.0;
  fc.coefb[2][5] = (3.0 - 6.0 * alfa) / 32.0;
  fc.coefb[2][6] = (-1.0 + 2.0 * alfa) / 64.0;
  fc.coefb[1][0] = (1.0 + 62.0 * alfa) / 64.0;
  fc.coefb[1][1] = (29.0 + 6.0 * alfa) / 32.0;
  fc.coefb[1][2] = (15.0 + 34.0 * alfa) / 64.0;
  fc.coefb[1][3] = (-5.0 + 10.0 * alfa) / 16.0;
  fc.coefb[1][4] = (15.0 - 30.0 * alfa) / 64.0;
  fc.coefb[1][5] = (-3.0 + 6.0 * alfa) / 32.0;
  fc.coefb[1][6] = (1.0 - 2.0 * alfa) / 64.0;
  fc.coefb[0][0] = (63.0 + 1.0 * beter_bouond) / 64.0;
  fc.coefb[0][1] = (3.0 + 29.0 * beter_bouond) / 32.0;
  fc.coefb[0][2] = (-15.0 + 15.0 * beter_bouond) / 64.0;
  fc.coefb[0][3] = (5.0 - 5.0 * beter_bouond) / 16.0;
  fc.coefb[0][4] = (-15.0 + 15.0 * beter_bouond) / 64.0;
  fc.coefb[0][5] = (3.0 - 3.0 * beter_bouond) / 32.0;
  fc.coefb[0][6] = (-1.0 + 1.0 * beter_bouond) / 64.0;
  for (auto& row : fc.coefh) for (double& v : row) v = 0.0;
  // coefh(3:4,:) allocated, never assigned, never read.
  fc.coefh[0][0] = (-1.0 + 1.0 * beter_halo) / 1024.0;
  fc.coefh[0][1] = (5.0 - 5.0 * beter_halo) / 512.0;
  fc.coefh[0][2] = (979.0 + 45.0 * beter_halo) / 1024.0;
  fc.coefh[0][3] = (15.0 + 113.0 * beter_halo) / 128.0;
  fc.coefh[0][4] = (-105.0 + 105.0 * beter_halo) / 512.0;
  fc.coefh[0][5] = (63.0 - 63.0 * beter_halo) / 256.0;
  fc.coefh[0][6] = (-105.0 + 105.0 * beter_halo) / 512.0;
  fc.coefh[0][7] = (15.0 - 15.0 * beter_halo) / 128.0;
  fc.coefh[0][8] = (-45.0 + 45.0 * beter_halo) / 1024.0;
  fc.coefh[0][9] = (5.0 - 5.0 * beter_halo) / 512.0;
  fc.coefh[0][10] = (-1.0 + 1.0 * beter_halo) / 1024.0;
  // The reference writes 1024. / 512. / 256. (default-real literals) in a few of these
  // (:409,411-413); they are exact powers of two, so promotion to real(8) is exact.
  fc.coefh[1][0] = (1.0 - 2.0 * alfa) / 1024.0;
  fc.coefh[1][1] = (-5.0 + 10.0 * alfa) / 512.0;
  fc.coefh[1][2] = (45.0 + 934.0 * alfa) / 1024.0;
  fc.coefh[1][3] = (113.0 + 30.0 * alfa) / 128.0;
  fc.coefh[1][4] = (105.0 + 302.0 * alfa) / 512.0;
  fc.coefh[1][5] = (-63.0 + 126.0 * alfa) / 256.0;
  fc.coefh[1][6] = (105.0 - 210.0 * alfa) / 512.0;
  fc.coefh[1][7] = (-15.0 + 30.0 * alfa) / 128.0;
  fc.coefh[1][8] = (45.0 - 90.0 * alfa) / 1024.0;
  fc.coefh[1][9] = (-5.0 + 10.0 * alfa) / 512.0;
  fc.coefh[1][10] = (1.0 - 2.0 * alfa) / 1024.0;
  fc.coefh[2][0] = (-1.0 + 2.0 * alfa) / 1024.0;
  fc.coefh[2][1] = (5.0 - 10.0 * alfa) / 512.0;
  fc.coefh[2][2] = (-45.0 + 90.0 * alfa) / 1024.0;
  fc.coefh[2][3] = (15.0 + 98.0 * alfa) / 128.0;
  fc.coefh[2][4] = (407.0 + 210.0 * alfa) / 512.0;
  fc.coefh[2][5] = (63.0 + 130.0 * alfa) / 256.0;
  fc.coefh[2][6] = (-105.0 + 210.0 * alfa) / 512.0;
  fc.coefh[2][7] = (15.0 - 30.0 * alfa) / 128.0;
  fc.coefh[2][8] = (-45.0 + 90.0 * alfa) / 1024.0;
  fc.coefh[2][9] = (5.0 - 10.0 * alfa) / 512.0;
  fc.coefh[2][10] = (-1.0 + 2.0 * alfa) / 1024.0;
}

// src/filter.F90:31-100  compact_filter_initiate (note= absent on the hot path)
void compact_filter_initiate(CompactScheme& s, int ntype, int dim, double alfa) {
  int i_0, i_m;
  double beter_0, beter_m;
  switch (ntype) {  // :44-71
    case 1: i_0 = 0;  i_m = dim + 3; beter_0 = 0.98; beter_m = 1.11; break;
    case 2: i_0 = -3; i_m = dim;     beter_0 = 1.11; beter_m = 0.98; break;
    case 3: i_0 = -3; i_m = dim + 3; beter_0 = 1.11; beter_m = 1.11; break;
    case 4: i_0 = 0;  i_m = dim;     beter_0 = 0.98; beter_m = 0.98; break;
    default: std::fprintf(stderr, "oracle: bad ntype %d\n", ntype); std::abort();
  }
  s.first_node = i_0; s.last_node = i_m; s.dimension = dim; s.nbctype = ntype;
  const int n = s.size();
  s.a.assign(n, alfa); s.c.assign(n, alfa);  // :80-81
  s.a[0] = beter_0;     s.c[0] = beter_0;    // :91-95
  s.a[n - 1] = beter_m; s.c[n - 1] = beter_m;
  thomas_preprocess(s);  // :98
}

// src/filter.F90:156-285  compact_filter_rhs (note= absent)
void compact_filter_rhs(const CompactScheme& s, const FilterCoef& fc, const double* f, double* dd) {
  const int i_0 = s.first_node, i_m = s.last_node, ntype = s.nbctype;
  auto d = [&](int node) -> double& { return dd[node - i_0]; };
  int i_s, i_e, j, k;
  double var0, var1, var2, var3, var4, var5;
  if (ntype == 1 || ntype == 4) {  // :176-204
    i_s = i_0 + 5;
    for (k = i_0; k <= i_0 + 2; ++k) {
      var0 = 0.0;
      for (j = 0; j <= 6; ++j) var0 = var0 + fc.coefb[k - i_0][j] * f[i_0 + j];
      d(k) = var0;
    }
    j = i_0 + 3;
    var0 = f[j] + f[j];
    var1 = f[j + 1] + f[j - 1];
    var2 = f[j + 2] + f[j - 2];
    var3 = f[j + 3] + f[j - 3];
    d(j) = fc.coef6i[0] * var0 + fc.coef6i[1] * var1 + fc.coef6i[2] * var2 + fc.coef6i[3] * var3;
    j = i_0 + 4;
    var0 = f[j] + f[j];
    var1 = f[j + 1] + f[j - 1];
    var2 = f[j + 2] + f[j - 2];
    var3 = f[j + 3] + f[j - 3];
    var4 = f[j + 4] + f[j - 4];
    d(j) = fc.coef8i[0] * var0 + fc.coef8i[1] * var1 + fc.coef8i[2] * var2 +
           fc.coef8i[3] * var3 + fc.coef8i[4] * var4;
  } else {  // interface :206-218
    i_s = i_0 + 3;
    for (k = i_0; k <= i_0 + 2; ++k) {
      var0 = 0.0;
      for (j = 0; j <= 10; ++j) var0 = var0 + fc.coefh[k - i_0][j] * f[i_0 - 2 + j];
      d(k) = var0;
    }
  }
  if (ntype == 2 || ntype == 4) {  // :222-249
    i_e = i_m - 5;
    for (k = i_m - 2; k <= i_m; ++k) {
      var0 = 0.0;
      for (j = 0; j <= 6; ++j) var0 = var0 + fc.coefb[i_m - k][j] * f[i_m - j];
      d(k) = var0;
    }
    j = i_m - 3;
    var0 = f[j] + f[j];
    var1 = f[j + 1] + f[j - 1];
    var2 = f[j + 2] + f[j - 2];
    var3 = f[j + 3] + f[j - 3];
    d(j) = fc.coef6i[0] * var0 + fc.coef6i[1] * var1 + fc.coef6i[2] * var2 + fc.coef6i[3] * var3;
    j = i_m - 4;
    var0 = f[j] + f[j];
    var1 = f[j + 1] + f[j - 1];
    var2 = f[j + 2] + f[j - 2];
    var3 = f[j + 3] + f[j - 3];
    var4 = f[j + 4] + f[j - 4];
    d(j) = fc.coef8i[0] * var0 + fc.coef8i[1] * var1 + fc.coef8i[2] * var2 +
           fc.coef8i[3] * var3 + fc.coef8i[4] * var4;
  } else {  // :251-261
    i_e = i_m - 3;
    for (k = i_m - 2; k <= i_m; ++k) {
      var0 = 0.0;
      for (j = 0; j <= 10; ++j) var0 = var0 + fc.coefh[i_m - k][j] * f[i_m + 2 - j];
      d(k) = var0;
    }
  }
  for (j = i_s; j <= i_e; ++j) {  // :271-283
    var0 = f[j] + f[j];
    var1 = f[j + 1] + f[j - 1];
    var2 = f[j + 2] + f[j - 2];
    var3 = f[j + 3] + f[j - 3];
    var4 = f[j + 4] + f[j - 4];
    var5 = f[j + 5] + f[j - 5];
    d(j) = fc.coef10i[0] * var0 + fc.coef10i[1] * var1 + fc.coef10i[2] * var2 +
           fc.coef10i[3] * var3 + fc.coef10i[4] * var4 + fc.coef10i[5] * var5;
  }
}

// src/filter.F90:112-144  compact_filter
void compact_filter(const CompactScheme& s, const FilterCoef& fc, const double* f, double* ff,
                    double* work) {
  const int n = s.size();
  double* d = work;
  double* xx = work + n;
  compact_filter_rhs(s, fc, f, d);
  thomas_solve(s, d, xx);
  for (int i = 0; i <= s.dimension; ++i) ff[i] = xx[i - s.first_node];
  if (s.nbctype == 1 || s.nbctype == 4) ff[0] = f[0];                          // :141
  if (s.nbctype == 2 || s.nbctype == 4) ff[s.dimension] = f[s.dimension];      // :142
}

// src/solver.F90:104-126 (nondimen branch) and miniapps/tgv_solver_3d/tgvsolver.F90:183-195
void Thermo::refcal(double sutherland_s) {
  const1 = 1.0 / (gamma * (gamma - 1.0) * (mach * mach));
  const2 = gamma * (mach * mach);
  const3 = (gamma - 1.0) / 3.0 * prandtl * (mach * mach);
  const4 = (gamma - 1.0) * (mach * mach) * reynolds * prandtl;
  const5 = (gamma - 1.0) * (mach * mach);
  const6 = 1.0 / (gamma - 1.0);
  const7 = (gamma - 1.0) * (mach * mach) * reynolds * prandtl;
  tempconst = sutherland_s / ref_tem;
  tempconst1 = 1.0 + tempconst;
  nondimen = true;
  roinf = 1.0; tinf = 1.0;
  pinf = roinf * tinf / const2;
}

// src/solver.F90:124-148: ref_tem, ref_vel, ref_len, ref_den are inputs; Mach and Reynolds follow
void Thermo::refcal_dimensional() {
  nondimen = false;
  rgas = 287.1;
  cp = gamma / (gamma - 1.0) * rgas;
  cv = rgas / (gamma - 1.0);
  tinf = ref_tem; roinf = ref_den;
  pinf = thermal_p(roinf, tinf);
  const double ref_miu = miucal(ref_tem);
  mach = ref_vel / sos(ref_tem);
  reynolds = ref_den * ref_vel * ref_len / ref_miu;
  const1 = 1.0 / (gamma * (gamma - 1.0) * (mach * mach));
  const2 = gamma * (mach * mach);
  const3 = (gamma - 1.0) / 3.0 * prandtl * (mach * mach);
  const4 = (gamma - 1.0) * (mach * mach) * reynolds * prandtl;
  const5 = (gamma - 1.0) * (mach * mach);
  const6 = 1.0 / (gamma - 1.0);
  const7 = (gamma - 1.0) * (mach * mach) * reynolds * prandtl;
}

}  // namespace astr_oracle
