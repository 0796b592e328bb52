// =====================================================================================
// oracle/ -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (C++17, IEEE fp64, compiled with -ffp-contract=off) of the ASTR
// right-hand-side / Runge-Kutta-stage arithmetic.  It is the CHECKER for the CUDA
// path in astr_b200/csrc: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call anything in this
// directory.  Nothing here is shipped or linked into libastr_gpu.so.
//
// Parity pin: the mini-app mode (miniapp.cpp) reproduces the reference's shipped
// golden history miniapps/tgv_solver_3d/state.ref_128 (committed as
// tests/golden/state.ref_128); see tests/test_oracle_golden.py.  The main-solver
// mode (solver.cpp: stage order of src/mainloop.F90, curvilinear metrics, qswap
// averaging, multi-block layouts, ntype 1/2/4 closures) has no stored number in the
// reference: for those parts parity is UNPINNED and rests on the line-by-line
// restatement below (every function cites the reference file:line it follows).
//
// The reference (Fortran 90) cannot be built in this image (no Fortran compiler,
// no MPI, no HDF5), so there is no oracle/_ref binary.
// =====================================================================================
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <string>

namespace astr_oracle {

constexpr int hm = 5;  // src/commvar.F90:184  parameter(hm=5)

// src/constdef.F90:12-27 -- rational constants are QUOTIENTS, not decimal literals.
constexpr double num1d3 = 1.0 / 3.0;
constexpr double num2d3 = 2.0 / 3.0;
constexpr double num1d6 = 1.0 / 6.0;
constexpr double num1d12 = 1.0 / 12.0;
constexpr double num7d9 = 7.0 / 9.0;
constexpr double num1d36 = 1.0 / 36.0;
constexpr double num1d60 = 1.0 / 60.0;

// src/commtype.F90:13-23  type compact_scheme.  Arrays are indexed [node-first_node].
struct CompactScheme {
  int first_node = 0, last_node = 0, dimension = 0, nbctype = 0;
  std::vector<double> a, c, ac1, ac2, ac3;
  int size() const { return last_node - first_node + 1; }
};

// src/commfunc.F90:752-774
void thomas_preprocess(CompactScheme& s);
// src/commfunc.F90:790-813  (d is overwritten by the forward sweep, as in the reference)
void thomas_solve(const CompactScheme& s, double* d, double* x);

// src/derivative.F90:63-158   scheme = e.g. 643, kind = 'c' (compact) or 'e' (explicit)
void fd_scheme_initiate(CompactScheme& s, int nscheme, char kind, int ntype, int dim);
// src/derivative.F90:210-306   f points at node 0; f[-hm .. dim+hm] must be readable
void compact_fd_rhs(const CompactScheme& s, const double* f, double* d);
// src/derivative.F90:171-198   df[0..dim]; work needs 2*s.size() doubles
void df_compact(const CompactScheme& s, const double* f, double* df, double* work);
// src/derivative.F90:350-413
void diff6ec(const double* f, int dim, int ntype, double* out);

// src/filter.F90:299-432
struct FilterCoef {
  double coef2i[2], coef4i[3], coef6i[4], coef8i[5], coef10i[6];
  double coefb[5][9], coefh[5][11];
};
void filter_coefficient_cal(FilterCoef& fc, double alfa, double beter_halo, double beter_bound);
// src/filter.F90:31-100
void compact_filter_initiate(CompactScheme& s, int ntype, int dim, double alfa);
// src/filter.F90:156-285
void compact_filter_rhs(const CompactScheme& s, const FilterCoef& fc, const double* f, double* d);
// src/filter.F90:112-144   ff[0..dim]; work needs 2*s.size() doubles
void compact_filter(const CompactScheme& s, const FilterCoef& fc, const double* f, double* ff,
                    double* work);

// ---------------------------------------------------------------------------------
// Thermodynamics (non-dimensional branch; src/fludyna.F90)
// ---------------------------------------------------------------------------------
struct Thermo {
  double gamma = 1.4, mach = 0.1, reynolds = 1600.0, prandtl = 0.72, ref_tem = 273.15;
  double const1, const2, const3, const4, const5, const6, const7;
  double tempconst, tempconst1;  // Sutherland
  // nondimen=f (src/solver.F90:124-148): SI units, rgas=287.1, reference state ref_tem/ref_vel/ref_len/ref_den
  bool nondimen = true;
  double rgas = 287.1, cp = 0.0, cv = 0.0, ref_vel = 1.0, ref_len = 1.0, ref_den = 1.0;
  double roinf = 1.0, tinf = 1.0, pinf = 0.0;
  // sutherland_s: 110.3 in src/solver.F90:122, 110.4 in miniapps/tgv_solver_3d/tgvsolver.F90:133
  void refcal(double sutherland_s);  // src/solver.F90:104-126
  void refcal_dimensional();         // src/solver.F90:124-148
  double miucal(double t) const {    // src/fludyna.F90:791-817
    if (nondimen) return t * std_sqrt(t) * tempconst1 / (t + tempconst);
    const double tnondim = t / 273.15;
    return 1.716e-5 * tnondim * std_sqrt(tnondim) * (273.15 + 110.4) / (t + 110.4);
  }
  // viscosity as diffrsdcal6 uses it (src/solver.F90:2456-2460) and the conductivity factor (:2519-2523)
  double miu_eff(double t) const { return nondimen ? miucal(t) / reynolds : miucal(t); }
  double hcc(double miu) const { return nondimen ? (miu / prandtl) / const5 : cp * miu / prandtl; }
  // thermal_scar (src/fludyna.F90:45-88)
  double thermal_T(double p, double rho) const { return nondimen ? p / rho * const2 : p / rho / rgas; }
  double thermal_rho(double p, double t) const { return nondimen ? p / t * const2 : p / t / rgas; }
  double thermal_p(double rho, double t) const { return nondimen ? rho * t / const2 : rho * t * rgas; }
  double cotem() const { return nondimen ? const1 : cv; }            // fvar2q, src/fludyna.F90:344-348
  double sos(double t) const { return nondimen ? std_sqrt(t) / mach : std_sqrt(gamma * rgas * t); }  // :850-854
  static double std_sqrt(double v);
};

// A 3-D array with hm halo cells on every side, Fortran order (i fastest).
struct Field {
  int im = 0, jm = 0, km = 0;
  long ni = 0, nj = 0, nk = 0;
  std::vector<double> v;
  void alloc(int im_, int jm_, int km_) {
    im = im_; jm = jm_; km = km_;
    ni = im + 1 + 2 * hm; nj = jm + 1 + 2 * hm; nk = km + 1 + 2 * hm;
    v.assign((size_t)ni * nj * nk, 0.0);
  }
  inline size_t idx(int i, int j, int k) const {
    return (size_t)(i + hm) + (size_t)ni * ((size_t)(j + hm) + (size_t)nj * (size_t)(k + hm));
  }
  inline double& operator()(int i, int j, int k) { return v[idx(i, j, k)]; }
  inline double operator()(int i, int j, int k) const { return v[idx(i, j, k)]; }
};

}  // namespace astr_oracle
