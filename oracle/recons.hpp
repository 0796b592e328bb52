// oracle/recons.hpp -- TEST INFRASTRUCTURE ONLY (see astr_oracle.hpp).  Included by solver.cpp after
// upwind.hpp (uses minmod2 / minmod4 / chardecomp / steger_warming_node from there).
//
// Restatement of the reference's EXPLICIT upwind-biased convection path (conschm = '..e' with an odd
// first digit):
//   recons_exp                          src/flux.F90:269-350
//   suw3, suw5, suw7                    src/flux.F90:388-422
//   MP5 (ul absent), MP7                src/flux.F90:434-551
//   MP5LD, MP7LD                        src/flux.F90:556-708
//   round                               src/flux.F90:717-752
//   WENO5, WENO7, WENO5Z, WENO7Z        src/flux.F90:761-1055
//   iwind8                              src/solver.F90:1213-1234
//   convrsduwd                          src/solver.F90:548-1201
// Arrays are 0-based here: u[k] is the reference's u(k+1) (u(k) for the 0-based MP7 / MP7LD).
//
// Parity status: UNPINNED -- the reference ships no stored number for this path.

static inline double suw3(const double* u) { return -num1d6 * u[0] + (5.0 / 6.0) * u[1] + num1d3 * u[2]; }
static inline double suw5(const double* u) {
  return (2.0 * u[0] - 13.0 * u[1] + 47.0 * u[2] + 27.0 * u[3] - 3.0 * u[4]) / 60.0;
}
static inline double suw7(const double* u) {
  return (-3.0 * u[0] + 25.0 * u[1] - 101.0 * u[2] + 319.0 * u[3] + 214.0 * u[4] - 38.0 * u[5] + 4.0 * u[6]) / 420.0;
}

// the monotonicity-preserving bounds shared by MP5 / MP7 / MP5LD / MP7LD: w[0..4] = the five values
// centred on the upwind cell (the reference's u(1:5) of MP5, u(1:5) of the 0-based MP7)
static inline double mp_limit(const double* w, double ulinear) {
  const double dm1 = w[0] - 2.0 * w[1] + w[2];
  const double d0 = w[1] - 2.0 * w[2] + w[3];
  const double d1 = w[2] - 2.0 * w[3] + w[4];
  const double dhm1 = minmod4(4.0 * dm1 - d0, 4.0 * d0 - dm1, dm1, d0);
  const double dh0 = minmod4(4.0 * d0 - d1, 4.0 * d1 - d0, d0, d1);
  const double uUL = w[2] + 4.0 * (w[2] - w[1]);
  const double uAV = 0.5 * (w[2] + w[3]);
  const double uMD = uAV - 0.5 * dh0;
  const double uLC = w[2] + 0.5 * (w[2] - w[1]) + 1.333333333333333 * dhm1;
  double var1 = std::min(std::min(w[2], w[3]), uMD);
  double var2 = std::min(std::min(w[2], uUL), uLC);
  const double uMIN = std::max(var1, var2);
  var1 = std::max(std::max(w[2], w[3]), uMD);
  var2 = std::max(std::max(w[2], uUL), uLC);
  const double uMAX = std::min(var1, var2);
  var1 = uMIN - ulinear;
  var2 = uMAX - ulinear;
  return ulinear + minmod2(var1, var2);
}
static inline double mp_switch(const double* w, double ulinear) {   // (ulinear-u3)*(ulinear-uMP)
  const double uMP = w[2] + minmod2(w[3] - w[2], 4.0 * (w[2] - w[1]));
  return (ulinear - w[2]) * (ulinear - uMP);
}

// MP5(u) with ul and discont absent (src/flux.F90:434-496): u[0..4]
static inline double mp5_std(const double* u) {
  const double ulinear = suw5(u);
  return mp_switch(u, ulinear) >= 1.e-10 ? mp_limit(u, ulinear) : ulinear;
}
// MP7(u(0:6)) (src/flux.F90:498-551): the limiter window is u(1:5)
static inline double mp7_std(const double* u) {
  const double ulinear = suw7(u);
  return mp_switch(u + 1, ulinear) >= 1.e-10 ? mp_limit(u + 1, ulinear) : ulinear;
}
// MP5LD(u(1:6), weightBW, lskt, lsod) (src/flux.F90:556-634)
static inline double mp5ld(const double* u, double weightBW) {
  const double vadp = 1.666666666666667e-2 * weightBW;
  const double b1 = -vadp + 3.333333333333333e-2;
  const double b2 = 5.0 * vadp - 2.166666666666667e-1;
  const double b3 = -10.0 * vadp + 7.833333333333333e-1;
  const double b4 = 10.0 * vadp + 0.45;
  const double b5 = -5.0 * vadp - 5.e-2;
  const double b6 = vadp;
  const double ulinear = b1 * u[0] + b2 * u[1] + b3 * u[2] + b4 * u[3] + b5 * u[4] + b6 * u[5];
  return mp_switch(u, ulinear) >= 1.e-10 ? mp_limit(u, ulinear) : ulinear;
}
// MP7LD(u(0:7), weightBW, lskt, lsod) (src/flux.F90:636-708): limited wherever lskt, window u(1:5)
static inline double mp7ld(const double* u, double weightBW, bool lskt) {
  const double vadp = 3.571428571428571e-3 * weightBW;
  const double b0 = 1.0 * vadp - 7.142857142857143e-3;
  const double b1 = -7.0 * vadp + 5.952380952380952e-2;
  const double b2 = 21.0 * vadp - 0.240476190476190;
  const double b3 = -35.0 * vadp + 0.759523809523809;
  const double b4 = 35.0 * vadp + 0.509523809523809;
  const double b5 = -21.0 * vadp - 9.047619047619047e-2;
  const double b6 = 7.0 * vadp + 9.523809523809525e-3;
  const double b7 = -1.0 * vadp;
  const double ulinear = b0 * u[0] + b1 * u[1] + b2 * u[2] + b3 * u[3] + b4 * u[4] + b5 * u[5] + b6 * u[6] + b7 * u[7];
  return lskt ? mp_limit(u + 1, ulinear) : ulinear;
}

// round(u(1:3)) (src/flux.F90:717-752)
static inline double round3(const double* u) {
  const double eps = 1.e-16;
  const double z0c = (u[1] - u[0] + eps) / (u[2] - u[0] + eps);
  const double a1c = 1.0 + 12.0 * z0c * z0c;
  const double a2c = 1.0 + 5.0 * (z0c - 1.0) * (z0c - 1.0);
  const double plc = 1100.0 * (z0c - 0.05) * (z0c - 0.05) * (z0c - 0.05) * (0.47 - z0c) * (0.47 - z0c) * (0.47 - z0c);
  const double prc = 18000.0 * (z0c - 0.55) * (z0c - 0.55) * (z0c - 0.55) * (0.97 - z0c) * (0.97 - z0c) * (0.97 - z0c) *
                     (0.97 - z0c) * (0.97 - z0c);
  const double p1c = 0.833333333333333 * z0c + 0.333333333333333 + std::max(plc, 0.0) + std::max(prc, 0.0);
  const double p2c = 1.5 * z0c;
  const double p3c = 0.5 * z0c + 0.5;
  const double wc1c = 1.0 / a1c / a1c / a1c / a1c;
  const double wc2c = 1.0 / a2c / a2c / a2c / a2c / a2c / a2c / a2c / a2c;
  const double gc = (p1c * (1.0 - wc1c) + p2c * wc1c) * (1.0 - wc2c) + p3c * wc2c;
  return gc * (u[2] - u[0]) + u[0];
}

static inline double sq(double v) { return v * v; }

// WENO5 / WENO5Z (src/flux.F90:761-815, :911-968): u[0..4]
static inline double weno5_any(const double* u, bool z) {
  const double eps = 1.e-6;
  const double uh1 = 0.333333333333333 * u[0] - 1.16666666666667 * u[1] + 1.83333333333333 * u[2];
  const double uh2 = -0.166666666666667 * u[1] + 0.833333333333333 * u[2] + 0.333333333333333 * u[3];
  const double uh3 = 0.333333333333333 * u[2] + 0.833333333333333 * u[3] - 0.166666666666667 * u[4];
  const double beter1 = 1.08333333333333 * sq(u[0] - 2.0 * u[1] + u[2]) + 0.25 * sq(u[0] - 4.0 * u[1] + 3.0 * u[2]);
  const double beter2 = 1.08333333333333 * sq(u[1] - 2.0 * u[2] + u[3]) + 0.25 * sq(u[1] - u[3]);
  const double beter3 = 1.08333333333333 * sq(u[2] - 2.0 * u[3] + u[4]) + 0.25 * sq(3.0 * u[2] - 4.0 * u[3] + u[4]);
  const double C1 = 0.1, C2 = 0.6, C3 = 0.3;
  double alfa1, alfa2, alfa3;
  if (z) {
    const double tau5 = std::fabs(beter3 - beter1);
    alfa1 = C1 + C1 * (tau5 / sq(beter1 + eps));
    alfa2 = C2 + C2 * (tau5 / sq(beter2 + eps));
    alfa3 = C3 + C3 * (tau5 / sq(beter3 + eps));
  } else {
    alfa1 = C1 / sq(beter1 + eps);
    alfa2 = C2 / sq(beter2 + eps);
    alfa3 = C3 / sq(beter3 + eps);
  }
  const double alfaS = alfa1 + alfa2 + alfa3;
  const double WT1 = alfa1 / alfaS, WT2 = alfa2 / alfaS, WT3 = alfa3 / alfaS;
  return WT1 * uh1 + WT2 * uh2 + WT3 * uh3;
}

// WENO7 / WENO7Z (src/flux.F90:823-899, :977-1055): u[0..6]
static inline double weno7_any(const double* u, bool z) {
  const double eps = 1.e-6;
  const double uh1 = -0.25 * u[0] + 1.083333333333333 * u[1] - 1.916666666666667 * u[2] + 2.083333333333333 * u[3];
  const double uh2 = 8.333333333333333e-2 * u[1] - 4.166666666666667e-1 * u[2] + 1.083333333333333 * u[3] + 0.25 * u[4];
  const double uh3 = -8.333333333333333e-2 * u[2] + 5.833333333333333e-1 * u[3] + 5.833333333333333e-1 * u[4] -
                     8.333333333333333e-2 * u[5];
  const double uh4 = 0.25 * u[3] + 1.083333333333333 * u[4] - 4.166666666666667e-1 * u[5] + 8.333333333333333e-2 * u[6];
  const double k1 = 2.777777777777778e-2, k2 = 1.083333333333333, k3 = 1.084722222222222;
  double df1 = k1 * sq(-2.0 * u[0] + 9.0 * u[1] - 18.0 * u[2] + 11.0 * u[3]);
  double df2 = k2 * sq(-1.0 * u[0] + 4.0 * u[1] - 5.0 * u[2] + 2.0 * u[3]);
  double df3 = k3 * sq(-1.0 * u[0] + 3.0 * u[1] - 3.0 * u[2] + 1.0 * u[3]);
  const double beter1 = df1 + df2 + df3;
  df1 = k1 * sq(u[1] - 6.0 * u[2] + 3.0 * u[3] + 2.0 * u[4]);
  df2 = k2 * sq(u[2] - 2.0 * u[3] + u[4]);
  df3 = k3 * sq(-1.0 * u[1] + 3.0 * u[2] - 3.0 * u[3] + 1.0 * u[4]);
  const double beter2 = df1 + df2 + df3;
  df1 = k1 * sq(-2.0 * u[2] - 3.0 * u[3] + 6.0 * u[4] - 1.0 * u[5]);
  df2 = k2 * sq(u[2] - 2.0 * u[3] + u[4]);
  df3 = k3 * sq(-1.0 * u[2] + 3.0 * u[3] - 3.0 * u[4] + 1.0 * u[5]);
  const double beter3 = df1 + df2 + df3;
  df1 = k1 * sq(-11.0 * u[3] + 18.0 * u[4] - 9.0 * u[5] + 2.0 * u[6]);
  df2 = k2 * sq(2.0 * u[3] - 5.0 * u[4] + 4.0 * u[5] - u[6]);
  df3 = k3 * sq(-1.0 * u[3] + 3.0 * u[4] - 3.0 * u[5] + 1.0 * u[6]);
  const double beter4 = df1 + df2 + df3;
  const double C1 = 2.857142857142857e-2, C2 = 3.428571428571429e-1, C3 = 5.142857142857143e-1, C4 = 1.142857142857143e-1;
  double alfa1, alfa2, alfa3, alfa4;
  if (z) {
    const double tau7 = std::fabs(beter4 - beter1);
    alfa1 = C1 + C1 * (tau7 / sq(beter1 + eps));
    alfa2 = C2 + C2 * (tau7 / sq(beter2 + eps));
    alfa3 = C3 + C3 * (tau7 / sq(beter3 + eps));
    alfa4 = C4 + C4 * (tau7 / sq(beter4 + eps));
  } else {
    alfa1 = C1 / sq(beter1 + eps);
    alfa2 = C2 / sq(beter2 + eps);
    alfa3 = C3 / sq(beter3 + eps);
    alfa4 = C4 / sq(beter4 + eps);
  }
  const double alfaS = alfa1 + alfa2 + alfa3 + alfa4;
  const double WT1 = alfa1 / alfaS, WT2 = alfa2 / alfaS, WT3 = alfa3 / alfaS, WT4 = alfa4 / alfaS;
  return WT1 * uh1 + WT2 * uh2 + WT3 * uh3 + WT4 * uh4;
}

// recons_exp(f(1:8), inode, dim, ntype, reschem, shock, solid) (src/flux.F90:269-350); returns NaN for a
// reschem the reference stops on
static double recons_exp(const double* f, int inode, int dim, int ntype, int reschem, bool shock, double bfacmpld) {
  const double bad = std::nan("");
  if ((ntype == 1 && inode == 0) || (ntype == 2 && inode == dim - 1)) return reschem == -1 ? f[3] : 0.5 * (f[3] + f[4]);
  if ((ntype == 1 && inode == 1) || (ntype == 2 && inode == dim - 2)) return reschem == -1 ? f[3] : suw3(f + 2);
  if ((ntype == 1 && inode == 2) || (ntype == 2 && inode == dim - 3)) {
    switch (reschem) {
      case -1: return f[3];
      case 0: return suw5(f + 1);
      case 1: return weno5_any(f + 1, false);
      case 2: return weno5_any(f + 1, true);
      case 3: return mp5_std(f + 1);
      case 5: return mp5ld(f + 1, bfacmpld);
      case 6: return round3(f + 2);
      default: return bad;
    }
  }
  switch (reschem) {
    case -1: return f[3];
    case 0: return suw7(f);
    case 1: return weno7_any(f, false);
    case 2: return weno7_any(f, true);
    case 3: return mp7_std(f);
    case 5: return mp7ld(f, bfacmpld, shock);
    case 6: return round3(f + 2);
    default: return bad;
  }
}

// src/solver.F90:1213-1234
static inline int iwind8(int i, int n /*1..8*/, int imin, int imax, char dir) {
  int w = (dir == '+') ? i + n - 4 : i + 5 - n;
  if (w < imin) w = imin;
  if (w > imax) w = imax;
  return w;
}

// src/solver.F90:548-1201 convrsduwd, the three directions written once over `d`.  Returns -1 where the
// reference stops (npdcj / npdck == 4, :810-821, :1002-1013), the number of degenerate eigen-decompositions
// otherwise.
static int convrsduwd(Case& c) {
  const int md = std::max(c.ia, std::max(c.ja, c.ka));
  int bad = 0;
  for (Block& b : c.blk) {
    if (b.npdc[1] == 4 || (c.ndims() == 3 && b.npdc[2] == 4)) return -1;
    const bool sson = !b.lshock.v.empty();
    for (int d = 0; d < c.ndims(); ++d) {
      const int dm = b.dim(d), nt = b.npdc[d];
      int lss, lee;
      if (nt == 1) { lss = 0; lee = dm + hm; }
      else if (nt == 2) { lss = -hm; lee = dm; }
      else if (nt == 3) { lss = -hm; lee = dm + hm; }
      else { lss = 0; lee = dm; }
      int o1, o2;
      if (d == 0) { o1 = 1; o2 = 2; } else if (d == 1) { o1 = 0; o2 = 2; } else { o1 = 0; o2 = 1; }
      const int np = md + 1 + 2 * hm;
      for_each_pencil(b, d, [&](int p1, int p2) {
        if (p1 < b.s[o1] || p1 > b.e[o1] || p2 < b.s[o2] || p2 > b.e[o2]) return;
        std::vector<double> fsw(10 * np, 0.0), Fh(5 * (md + 2), 0.0);
        double* fswp[5]; double* fswm[5]; double* fh[5];
        for (int n = 0; n < 5; ++n) {
          fswp[n] = fsw.data() + n * np + hm;
          fswm[n] = fsw.data() + (5 + n) * np + hm;
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
        for (int i = b.s[d] - 1; i <= b.e[d]; ++i) {
          bool lsh = true;                                  // :640-650
          if (sson) {
            if (i < 0) lsh = at(b.lshock, d, i + 1, p1, p2) != 0.0;
            else if (i + 1 > dm) lsh = at(b.lshock, d, i, p1, p2) != 0.0;
            else lsh = at(b.lshock, d, i, p1, p2) != 0.0 || at(b.lshock, d, i + 1, p1, p2) != 0.0;
          }
          const bool chr = c.lchardecomp && lsh;            // :661
          double flcp[5][8], flcm[5][8], Fhc[5], REV[5][5], LEV[5][5];
          if (chr) {
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
            for (int m = 0; m < 5; ++m)
              for (int n = 1; n <= 8; ++n) {
                int nwd = iwind8(i, n, lss, lee, '+');
                flcp[m][n - 1] = LEV[m][0] * fswp[0][nwd] + LEV[m][1] * fswp[1][nwd] + LEV[m][2] * fswp[2][nwd] +
                                 LEV[m][3] * fswp[3][nwd] + LEV[m][4] * fswp[4][nwd];
                nwd = iwind8(i, n, lss, lee, '-');
                flcm[m][n - 1] = LEV[m][0] * fswm[0][nwd] + LEV[m][1] * fswm[1][nwd] + LEV[m][2] * fswm[2][nwd] +
                                 LEV[m][3] * fswm[3][nwd] + LEV[m][4] * fswm[4][nwd];
              }
          } else {
            for (int n = 1; n <= 8; ++n) {
              int nwd = iwind8(i, n, lss, lee, '+');
              for (int m = 0; m < 5; ++m) flcp[m][n - 1] = fswp[m][nwd];
              nwd = iwind8(i, n, lss, lee, '-');
              for (int m = 0; m < 5; ++m) flcm[m][n - 1] = fswm[m][nwd];
            }
          }
          const bool hdiss = crinod_interface(b, d, i, dm, p1, p2);       // :660-666
          for (int m = 0; m < 5; ++m) {
            // :755-769: hdiss (critical node on either side) takes the split flux of one node
            const double var1 = hdiss ? flcp[m][3] : recons_exp(flcp[m], i, dm, nt, c.recon_schem, lsh, c.bfacmpld);
            const double var2 = hdiss ? flcm[m][3] : recons_exp(flcm[m], i, dm, nt, c.recon_schem, lsh, c.bfacmpld);
            Fhc[m] = var1 + var2;
          }
          if (chr) {
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
