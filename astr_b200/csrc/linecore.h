// astr_b200/csrc/linecore.h -- the register-resident partitioned line solve (sweep2.cu).
//
// Same linear systems as the reference's compact operators -- `fds%central`
// (src/derivative.F90:171-306), `compact_filter` (src/filter.F90:112-285), unit-diagonal
// tridiagonal rows a(r) x(r-1) + x(r) + c(r) x(r+1) = d(r) (src/commfunc.F90:752-813) --
// but factorised for a GPU: the rows are cut into ELEMENTS
//
//     [head block | regular chunk 0 | regular chunk 1 | ... | regular chunk NW-1 | tail block]
//
// * head / tail block: the 1-2 closure rows whose LEFT-hand side differs from the interior
//   (explicit interface rows a=c=0, wall rows a=c=2 and 1/4, filter end rows 1.11 / 0.98);
// * regular chunks: rows with the interior coefficients (alpha, 1, alpha).  Every chunk is
//   factorised from a FRESH start, so all chunks share one set of coefficient tables
//   (m, g, ac1, ev below), which the kernels read as immediate constant-bank operands.
//
// A thread owns one chunk in registers:
//   pass 1  e(s) = d(s) m(s) - e(s-1) g(s)                       local forward elimination
//   pass 2  y(s) = e(s) - ac1(s) y(s+1)                          local solution, zero carries;
//           only its head yh = y(0) and tail yt = e(len-1) leave the thread
//   scan    the heads/tails of all elements of a pencil obey a block-tridiagonal system
//           whose block LU collapses to two scalar recurrences (P forward, h backward)
//   pass 3  x(s) = (e(s) - ev(s) t_prev) - ac1(s) x(s+1),  x(len) = h_next
// with t_prev the true solution at the last row of the previous element and h_next at the
// first row of the next one.  Exact algebra (no truncation of the coupling); differences
// from the reference's single Thomas sweep are rounding only.
//
// Header shared by nvcc (device + host) and g++ (tests/emul_sweep2.cpp drives the same
// functions on the CPU against the oracle).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define ASTR_HD __host__ __device__ __forceinline__
#else
#define ASTR_HD inline
#endif

#define ASTR_LMAX 33   // rows of a full regular chunk
#define ASTR_EMAX 18   // elements per line: 16 regular chunks + head + tail block
#define ASTR_SMAX 2    // rows of a head / tail block

struct RegTab {        // fresh-start factorisation of rows (alpha, 1, alpha)
  double m[ASTR_LMAX], g[ASTR_LMAX], ac1[ASTR_LMAX], ev[ASTR_LMAX];
};
struct SpecTab {       // head / tail block
  int len, pad;
  double m[ASTR_SMAX], g[ASTR_SMAX], ac1[ASTR_SMAX], ev[ASTR_SMAX];
};
struct ElemTab {       // reduced-system coefficients of one element (reduced_scan below)
  double gamma, K, Q, gammap, Kp, Qp, D, pad;
};
struct LinePlan {
  int optype, ntype, n, first_node, nrows;
  int sh, st;          // rows of the head / tail block
  int nsf, nsl;        // rows with a closure RIGHT-hand side at each end (>= sh, st)
  int NW, Lr, len0;    // regular chunks: chunk 0 has len0 rows, the others Lr
  int E;               // NW + 2
  int ok;              // 0: this line cannot be handled by sweep2 (fall back)
  RegTab reg;
  SpecTab head, tail;
  ElemTab el[ASTR_EMAX];
};

// first row of regular chunk w
ASTR_HD int plan_chunk_row(const LinePlan& p, int w) {
  return p.sh + (w == 0 ? 0 : p.len0 + (w - 1) * p.Lr);
}

ASTR_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// ---- reduced system -------------------------------------------------------------------
// Heads h(e) / tails t(e) of the elements obey
//     h(e) = yh(e) - vh(e) t(e-1) - wh(e) h(e+1),   t(e) = yt(e) - vt(e) t(e-1) - wt(e) h(e+1).
// Block elimination from the first end gives t(e) = P(e) - Q(e) h(e+1) with the scalar recurrence
//     P(e) = S(e) - K(e) P(e-1),    S(e)  = yt(e) + gamma(e)  yh(e),
// and from the last end h(e) = P'(e) - Q'(e) t(e-1) with
//     P'(e) = S'(e) - K'(e) P'(e+1), S'(e) = yh(e) + gamma'(e) yt(e);
// at the interface in front of element e the two meet:
//     t(e-1) = D(e) (P(e-1) - Q(e-1) P'(e)),   h(e) = D(e) (P'(e) - Q'(e) P(e-1)).
// A thread of element `me` runs the forward recurrence up to me and the backward one down to me.
struct ScanOut { double Pm1, Pm, Pb0, Pb1; };   // P(me-1), P(me), P'(me), P'(me+1)

template <class GETS, class GETSP>
ASTR_HD ScanOut reduced_scan(const LinePlan& pl, GETS S, GETSP SP, int me) {
  ScanOut o;
  const int E = pl.E;
  double pf = 0.0, pf1 = 0.0;
#pragma unroll
  for (int e = 0; e < ASTR_EMAX; ++e) {
    if (e > me) break;
    pf1 = pf;
    pf = fma_(-pl.el[e].K, pf, S(e));
  }
  o.Pm = pf; o.Pm1 = pf1;
  double pb = 0.0, pb1 = 0.0;
#pragma unroll
  for (int e = ASTR_EMAX - 1; e >= 0; --e) {
    if (e < me) break;
    if (e < E) {
      pb1 = pb;
      pb = fma_(-pl.el[e].Kp, pb, SP(e));
    }
  }
  o.Pb0 = pb; o.Pb1 = pb1;
  return o;
}
// true solution at the last row of element me-1 / the first row of element me+1
ASTR_HD double scan_t_prev(const LinePlan& pl, int me, double Pm1, double Pb0) {
  return (me > 0) ? pl.el[me].D * fma_(-pl.el[me > 0 ? me - 1 : 0].Q, Pb0, Pm1) : 0.0;
}
ASTR_HD double scan_h_next(const LinePlan& pl, int me, double Pm, double Pb1) {
  return (me + 1 < pl.E) ? pl.el[me + 1].D * fma_(-pl.el[me + 1].Qp, Pm, Pb1) : 0.0;
}

// ---- head / tail block (<= ASTR_SMAX rows) ----------------------------------------------
ASTR_HD void spec_forward(const SpecTab& t, const double* d, double* e, double& yh, double& yt) {
  e[0] = d[0];
#pragma unroll
  for (int s = 1; s < ASTR_SMAX; ++s) e[s] = (s < t.len) ? fma_(-e[s - 1], t.g[s], d[s] * t.m[s]) : 0.0;
  double y = 0.0;
  yt = 0.0;
#pragma unroll
  for (int s = ASTR_SMAX - 1; s >= 0; --s)
    if (s < t.len) {
      if (s == t.len - 1) yt = e[s];
      y = fma_(-t.ac1[s], y, e[s]);
    }
  yh = y;
}
ASTR_HD void spec_back(const SpecTab& t, const double* e, double t_prev, double h_next, double* x) {
  double xn = h_next;
#pragma unroll
  for (int s = ASTR_SMAX - 1; s >= 0; --s) {
    if (s < t.len) { xn = fma_(-t.ac1[s], xn, fma_(-t.ev[s], t_prev, e[s])); x[s] = xn; } else x[s] = 0.0;
  }
}

// ---- right-hand sides -----------------------------------------------------------------
// coefficient tables of src/filter.F90:299-432 (only the rows the hot path reads)
struct FilterCoef {
  double coef6i[4], coef8i[5], coef10i[6];
  double coefb[4][9];
  double coefh[3][11];
};

template <int OP> struct OpT;
template <> struct OpT<0> { static constexpr int H = 2; };   // OP_DERIV
template <> struct OpT<1> { static constexpr int H = 5; };   // OP_FILTER

// interior right-hand side of the row whose node sits at window slot s + H
template <int OP, int WN>
ASTR_HD double reg_rhs(const double (&wv)[WN], int s, const FilterCoef& fc) {
  constexpr int H = OpT<OP>::H;
#define WS(k) wv[s + H + (k)]
  if (OP == 0) {
    // src/derivative.F90:296-304
    const double var1 = WS(1) - WS(-1);
    const double var2 = WS(2) - WS(-2);
    return (7.0 / 9.0) * var1 + (1.0 / 36.0) * var2;
  } else {
    // src/filter.F90:271-283
    const double var0 = WS(0) + WS(0);
    const double var1 = WS(1) + WS(-1);
    const double var2 = WS(2) + WS(-2);
    const double var3 = WS(3) + WS(-3);
    const double var4 = WS(4) + WS(-4);
    const double var5 = WS(5) + WS(-5);
    return fc.coef10i[0] * var0 + fc.coef10i[1] * var1 + fc.coef10i[2] * var2 + fc.coef10i[3] * var3 +
           fc.coef10i[4] * var4 + fc.coef10i[5] * var5;
  }
#undef WS
}

// Closure rows at the first end.  hw[k] = f(k - 5), k = 0..13 (nodes -5..8).  sf[k] is row k.
template <int OP>
ASTR_HD void closure_head(const double (&hw)[14], bool phys, const FilterCoef& fc, double (&sf)[5]) {
#define F(node) hw[(node) + 5]
#pragma unroll
  for (int k = 0; k < 5; ++k) sf[k] = 0.0;
  if (OP == 0) {
    if (phys) {  // src/derivative.F90:230-248
      sf[0] = -2.5 * F(0) + 2.0 * F(1) + 0.5 * F(2);
      sf[1] = 0.75 * (F(2) - F(0));
    } else {     // :250-260, row of ghost node -1
      sf[0] = 0.75 * (F(0) - F(-2)) - 0.15 * (F(1) - F(-3)) + (1.0 / 60.0) * (F(2) - F(-4));
    }
  } else {
    if (phys) {  // src/filter.F90:176-204
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 6; ++j) v = v + fc.coefb[k][j] * F(j);
        sf[k] = v;
      }
      {
        const double v0 = F(3) + F(3), v1 = F(4) + F(2), v2 = F(5) + F(1), v3 = F(6) + F(0);
        sf[3] = fc.coef6i[0] * v0 + fc.coef6i[1] * v1 + fc.coef6i[2] * v2 + fc.coef6i[3] * v3;
      }
      {
        const double v0 = F(4) + F(4), v1 = F(5) + F(3), v2 = F(6) + F(2), v3 = F(7) + F(1), v4 = F(8) + F(0);
        sf[4] = fc.coef8i[0] * v0 + fc.coef8i[1] * v1 + fc.coef8i[2] * v2 + fc.coef8i[3] * v3 + fc.coef8i[4] * v4;
      }
    } else {     // :206-218, ghost rows -3..-1 against the fixed window f(-5..5)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 10; ++j) v = v + fc.coefh[k][j] * F(-5 + j);
        sf[k] = v;
      }
    }
  }
#undef F
}

// Closure rows at the last end.  tw[k] = f(n - 8 + k), k = 0..13 (nodes n-8..n+5).
// sl[k] is row nrows - nsl + k.
template <int OP>
ASTR_HD void closure_tail(const double (&tw)[14], bool phys, const FilterCoef& fc, double (&sl)[5]) {
#define F(off) tw[(off) + 8]     // F(off) = f(n + off)
#pragma unroll
  for (int k = 0; k < 5; ++k) sl[k] = 0.0;
  if (OP == 0) {
    if (phys) {  // src/derivative.F90:264-281
      sl[0] = 0.75 * (F(0) - F(-2));
      sl[1] = 2.5 * F(0) - 2.0 * F(-1) - 0.5 * F(-2);
    } else {     // :283-292, row of ghost node n+1
      sl[0] = 0.75 * (F(2) - F(0)) - 0.15 * (F(3) - F(-1)) + (1.0 / 60.0) * (F(4) - F(-2));
    }
  } else {
    if (phys) {  // src/filter.F90:222-249 ; rows n-4, n-3, n-2, n-1, n
      {
        const double v0 = F(-4) + F(-4), v1 = F(-3) + F(-5), v2 = F(-2) + F(-6), v3 = F(-1) + F(-7),
                     v4 = F(0) + F(-8);
        sl[0] = fc.coef8i[0] * v0 + fc.coef8i[1] * v1 + fc.coef8i[2] * v2 + fc.coef8i[3] * v3 + fc.coef8i[4] * v4;
      }
      {
        const double v0 = F(-3) + F(-3), v1 = F(-2) + F(-4), v2 = F(-1) + F(-5), v3 = F(0) + F(-6);
        sl[1] = fc.coef6i[0] * v0 + fc.coef6i[1] * v1 + fc.coef6i[2] * v2 + fc.coef6i[3] * v3;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // node n-k uses coefb[k]
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 6; ++j) v = v + fc.coefb[k][j] * F(-j);
        sl[4 - k] = v;
      }
    } else {     // :251-261 ; ghost rows n+1..n+3 ; node n+3-k uses coefh[k]
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j <= 10; ++j) v = v + fc.coefh[k][j] * F(5 - j);
        sl[2 - k] = v;
      }
    }
  }
#undef F
}

// ---- regular chunk, passes 1 and 2 ------------------------------------------------------
// ROLE_MID : full chunk (ASTR_LMAX rows), interior right-hand sides only.
// ROLE_TAIL: full chunk whose last rows are closure rows with interior left-hand side (filter:
//            2 rows at an interface end, 4 at a wall) -> ov[0..]; static slots.
// ROLE_HEAD: chunk 0, `len` <= ASTR_LMAX rows, first rows closure rows (same counts) -> ov[0..].
enum { ROLE_MID = 0, ROLE_HEAD = 1, ROLE_TAIL = 2 };

// closure rows inside the regular chunk at one end
template <int OP> ASTR_HD int n_override(bool phys) { return OP == 0 ? 0 : (phys ? 4 : 2); }

template <int OP, int ROLE, int WN>
ASTR_HD void chunk_forward(const RegTab& t, const FilterCoef& fc, const double (&wv)[WN], int len, bool phys,
                           const double (&ov)[4], double (&e)[ASTR_LMAX], double& yh, double& yt) {
  constexpr int L = ASTR_LMAX;
  if (ROLE != ROLE_HEAD) {
#pragma unroll
    for (int s = 0; s < L; ++s) {
      double d = reg_rhs<OP>(wv, s, fc);
      if (ROLE == ROLE_TAIL && OP == 1) {
        if (s >= L - 4 && phys) d = ov[s >= L - 4 ? s - (L - 4) : 0];
        if (s >= L - 2 && !phys) d = ov[s >= L - 2 ? s - (L - 2) : 0];
      }
      e[s] = (s == 0) ? d : fma_(-e[s > 0 ? s - 1 : 0], t.g[s], d * t.m[s]);
    }
    double y = e[L - 1];
#pragma unroll
    for (int s = L - 2; s >= 0; --s) y = fma_(-t.ac1[s], y, e[s]);
    yh = y; yt = e[L - 1];
  } else {
    // chunk 0: `len` rows (uniform across the warp: the guards are branches, not selects)
    yt = 0.0;
#pragma unroll
    for (int s = 0; s < L; ++s) {
      if (s >= len) break;
      double d = reg_rhs<OP>(wv, s, fc);
      if (OP == 1 && s < 4 && (phys || s < 2)) d = ov[s < 4 ? s : 0];
      e[s] = (s == 0) ? d : fma_(-e[s > 0 ? s - 1 : 0], t.g[s], d * t.m[s]);
      yt = e[s];
    }
    double y = 0.0;
#pragma unroll
    for (int s = L - 1; s >= 0; --s)
      if (s < len) y = fma_(-t.ac1[s], y, e[s]);
    yh = y;
  }
}

// ---- regular chunk, pass 3: ST(s, x) receives the solution of row s (descending) ---------
template <int ROLE, class ST>
ASTR_HD void chunk_back(const RegTab& t, const double (&e)[ASTR_LMAX], int len, double t_prev, double h_next, ST st) {
  double x = h_next;
#pragma unroll
  for (int s = ASTR_LMAX - 1; s >= 0; --s) {
    if (ROLE != ROLE_HEAD || s < len) {
      x = fma_(-t.ac1[s], x, fma_(-t.ev[s], t_prev, e[s]));
      st(s, x);
    }
  }
}

// =======================================================================================
// host: plan construction
// =======================================================================================
#include <vector>

// Local (fresh-start) factorisation of rows [r0, r0+len) of the system (a, c):
// m, g, ac1, ev as used by pass 1-3, plus the four spike values of the element.
struct LocalFac { std::vector<double> m, g, ac1, ev; double vh, vt, wh, wt; };

inline LocalFac local_factor(const std::vector<double>& a, const std::vector<double>& c, int r0, int len,
                             bool first_elem, bool last_elem) {
  LocalFac f;
  f.m.assign(len, 1.0); f.g.assign(len, 0.0); f.ac1.assign(len, 0.0); f.ev.assign(len, 0.0);
  const double a_first = first_elem ? 0.0 : a[r0];
  f.ac1[0] = c[r0]; f.ev[0] = a_first;
  for (int i = 1; i < len; ++i) {
    const double den = 1.0 - a[r0 + i] * f.ac1[i - 1];
    f.m[i] = 1.0 / den; f.g[i] = a[r0 + i] / den; f.ac1[i] = c[r0 + i] / den;
    f.ev[i] = -f.ev[i - 1] * f.g[i];
  }
  if (last_elem) f.ac1[len - 1] = 0.0;   // no coupling past the last row of the system
  double vb = f.ev[len - 1];
  for (int i = len - 2; i >= 0; --i) vb = f.ev[i] - f.ac1[i] * vb;
  f.vh = vb; f.vt = f.ev[len - 1];
  double w = f.ac1[len - 1];
  for (int i = len - 2; i >= 0; --i) w = -f.ac1[i] * w;
  f.wh = w; f.wt = f.ac1[len - 1];
  return f;
}

// LHS rows of the reference's operators: fd_scheme_initiate (src/derivative.F90:63-158; scheme 643:
// a=c=1/3, explicit interface rows a=c=0, wall rows 2 and 1/4) and compact_filter_initiate
// (src/filter.F90:31-100; alfa with beter_halo=1.11 / beter_bound=0.98 end rows, comsolver.F90:121).
inline void build_lhs(int optype, int ntype, int n, double alfa, std::vector<double>& a, std::vector<double>& c,
                      int& first_node, int& nsf, int& nsl) {
  const bool p0 = (ntype == 1 || ntype == 4), pm = (ntype == 2 || ntype == 4);
  if (optype == 2 || optype == 3) {
    // compact_flux_initiate (src/flux.F90:32-118), scheme 543; alfa carries bfacmpld.
    // Unknowns are the interface values fh(first_node..last); '+' (optype 2) / '-' (optype 3).
    first_node = p0 ? -1 : -2;
    const int last = pm ? n : n + 1;
    const int N = last - first_node + 1;
    const double up = 0.5 - (1.0 / 6.0) * alfa, dn = (1.0 / 6.0) + (1.0 / 6.0) * alfa;
    a.assign(N, optype == 2 ? up : dn); c.assign(N, optype == 2 ? dn : up);
    const int e = N - 1;
    a[0] = c[0] = 0.0; a[e] = c[e] = 0.0;
    if (p0) { a[0] = c[0] = 2.0; a[1] = c[1] = 0.25; }
    if (pm) { a[e] = c[e] = 2.0; a[e - 1] = c[e - 1] = 0.25; }
    nsf = p0 ? 2 : 1; nsl = pm ? 2 : 1;
    return;
  }
  if (optype == 0) {
    first_node = p0 ? 0 : -1;
    const int last = pm ? n : n + 1;
    const int N = last - first_node + 1;
    a.assign(N, 1.0 / 3.0); c.assign(N, 1.0 / 3.0);
    const int e = N - 1;
    a[0] = c[0] = 0.0; a[e] = c[e] = 0.0;
    if (p0) { a[0] = c[0] = 2.0; a[1] = c[1] = 0.25; }
    if (pm) { a[e] = c[e] = 2.0; a[e - 1] = c[e - 1] = 0.25; }
    nsf = p0 ? 2 : 1; nsl = pm ? 2 : 1;
  } else {
    first_node = p0 ? 0 : -3;
    const int last = pm ? n : n + 3;
    const int N = last - first_node + 1;
    a.assign(N, alfa); c.assign(N, alfa);
    const int e = N - 1;
    a[0] = c[0] = p0 ? 0.98 : 1.11;
    a[e] = c[e] = pm ? 0.98 : 1.11;
    nsf = p0 ? 5 : 3; nsl = pm ? 5 : 3;
  }
}

// src/filter.F90:299-432
inline void build_filter_coef(FilterCoef& fc, double alfa, double bh, double bb) {
  const double c6[4] = {(11.0 + 10.0 * alfa) / 32.0, (15.0 + 34.0 * alfa) / 64.0, (-3.0 + 6.0 * alfa) / 32.0,
                        (1.0 - 2.0 * alfa) / 64.0};
  const double c8[5] = {(93.0 + 70.0 * alfa) / 256.0, (7.0 + 18.0 * alfa) / 32.0, (-7.0 + 14.0 * alfa) / 64.0,
                        (1.0 - 2.0 * alfa) / 32.0, (-1.0 + 2.0 * alfa) / 256.0};
  const double c10[6] = {(193.0 + 126.0 * alfa) / 512.0, (105.0 + 302.0 * alfa) / 512.0,
                         (-15.0 + 30.0 * alfa) / 128.0, (45.0 - 90.0 * alfa) / 1024.0,
                         (-5.0 + 10.0 * alfa) / 512.0,  (1.0 - 2.0 * alfa) / 1024.0};
  for (int i = 0; i < 4; ++i) fc.coef6i[i] = c6[i];
  for (int i = 0; i < 5; ++i) fc.coef8i[i] = c8[i];
  for (int i = 0; i < 6; ++i) fc.coef10i[i] = c10[i];
  for (int k = 0; k < 4; ++k) for (int j = 0; j < 9; ++j) fc.coefb[k][j] = 0.0;
  for (int k = 0; k < 3; ++k) for (int j = 0; j < 11; ++j) fc.coefh[k][j] = 0.0;
  const double b3[9] = {(1.0 - 2.0 * alfa) / 256.0, (-1.0 + 2.0 * alfa) / 32.0, (7.0 + 50.0 * alfa) / 64.0,
                        (25.0 + 14.0 * alfa) / 32.0, (35.0 + 58.0 * alfa) / 128.0, (-7.0 + 14.0 * alfa) / 32.0,
                        (7.0 - 14.0 * alfa) / 64.0, (-1.0 + 2.0 * alfa) / 32.0, (1.0 - 2.0 * alfa) / 256.0};
  const double b2[7] = {(-1.0 + 2.0 * alfa) / 64.0, (3.0 + 26.0 * alfa) / 32.0, (49.0 + 30.0 * alfa) / 64.0,
                        (5.0 + 6.0 * alfa) / 16.0, (-15.0 + 30.0 * alfa) / 64.0, (3.0 - 6.0 * alfa) / 32.0,
                        (-1.0 + 2.0 * alfa) / 64.0};
  const double b1[7] = {(1.0 + 62.0 * alfa) / 64.0, (29.0 + 6.0 * alfa) / 32.0, (15.0 + 34.0 * alfa) / 64.0,
                        (-5.0 + 10.0 * alfa) / 16.0, (15.0 - 30.0 * alfa) / 64.0, (-3.0 + 6.0 * alfa) / 32.0,
                        (1.0 - 2.0 * alfa) / 64.0};
  const double b0[7] = {(63.0 + 1.0 * bb) / 64.0, (3.0 + 29.0 * bb) / 32.0, (-15.0 + 15.0 * bb) / 64.0,
                        (5.0 - 5.0 * bb) / 16.0, (-15.0 + 15.0 * bb) / 64.0, (3.0 - 3.0 * bb) / 32.0,
                        (-1.0 + 1.0 * bb) / 64.0};
  for (int j = 0; j < 9; ++j) fc.coefb[3][j] = b3[j];
  for (int j = 0; j < 7; ++j) { fc.coefb[2][j] = b2[j]; fc.coefb[1][j] = b1[j]; fc.coefb[0][j] = b0[j]; }
  const double h0[11] = {(-1.0 + 1.0 * bh) / 1024.0, (5.0 - 5.0 * bh) / 512.0, (979.0 + 45.0 * bh) / 1024.0,
                         (15.0 + 113.0 * bh) / 128.0, (-105.0 + 105.0 * bh) / 512.0, (63.0 - 63.0 * bh) / 256.0,
                         (-105.0 + 105.0 * bh) / 512.0, (15.0 - 15.0 * bh) / 128.0, (-45.0 + 45.0 * bh) / 1024.0,
                         (5.0 - 5.0 * bh) / 512.0, (-1.0 + 1.0 * bh) / 1024.0};
  const double h1[11] = {(1.0 - 2.0 * alfa) / 1024.0, (-5.0 + 10.0 * alfa) / 512.0, (45.0 + 934.0 * alfa) / 1024.0,
                         (113.0 + 30.0 * alfa) / 128.0, (105.0 + 302.0 * alfa) / 512.0,
                         (-63.0 + 126.0 * alfa) / 256.0, (105.0 - 210.0 * alfa) / 512.0,
                         (-15.0 + 30.0 * alfa) / 128.0, (45.0 - 90.0 * alfa) / 1024.0,
                         (-5.0 + 10.0 * alfa) / 512.0, (1.0 - 2.0 * alfa) / 1024.0};
  const double h2[11] = {(-1.0 + 2.0 * alfa) / 1024.0, (5.0 - 10.0 * alfa) / 512.0, (-45.0 + 90.0 * alfa) / 1024.0,
                         (15.0 + 98.0 * alfa) / 128.0, (407.0 + 210.0 * alfa) / 512.0,
                         (63.0 + 130.0 * alfa) / 256.0, (-105.0 + 210.0 * alfa) / 512.0,
                         (15.0 - 30.0 * alfa) / 128.0, (-45.0 + 90.0 * alfa) / 1024.0,
                         (5.0 - 10.0 * alfa) / 512.0, (-1.0 + 2.0 * alfa) / 1024.0};
  for (int j = 0; j < 11; ++j) { fc.coefh[0][j] = h0[j]; fc.coefh[1][j] = h1[j]; fc.coefh[2][j] = h2[j]; }
}

// a, c: the LHS rows exactly as the reference sets them (fd_scheme_initiate
// src/derivative.F90:63-158, compact_filter_initiate src/filter.F90:31-100).
// maxw: most regular chunks a line may be cut into (warps per CTA).
inline void build_line_plan(LinePlan& p, int optype, int ntype, int n, int first_node, int nsf, int nsl,
                            const std::vector<double>& a, const std::vector<double>& c, int maxw) {
  p = LinePlan();
  p.optype = optype; p.ntype = ntype; p.n = n; p.first_node = first_node;
  p.nrows = (int)a.size(); p.nsf = nsf; p.nsl = nsl;
  const int N = p.nrows;
  const bool p0 = (ntype == 1 || ntype == 4), pm = (ntype == 2 || ntype == 4);
  // head / tail block = leading / trailing rows whose coefficients differ from the interior
  const bool deriv = (optype == 0);
  p.sh = (deriv && p0) ? 2 : 1;
  p.st = (deriv && pm) ? 2 : 1;
  const int nreg = N - p.sh - p.st;
  const double alpha = (nreg > 0) ? a[p.sh] : 0.0;
  p.ok = 0;
  if (nreg < 12 || n < 9) return;
  for (int r = p.sh; r < N - p.st; ++r)
    if (a[r] != alpha || c[r] != alpha) return;     // not a Toeplitz interior: not ours
  int NW = (nreg + ASTR_LMAX - 1) / ASTR_LMAX;
  if (NW < 2 || NW > maxw || NW + 2 > ASTR_EMAX) return;
  // every chunk is full except chunk 0, which takes the remainder and must hold the closure
  // right-hand-side rows of the first end
  const int Lr = ASTR_LMAX;
  const int len0 = nreg - (NW - 1) * Lr;
  if (len0 < (nsf - p.sh > 1 ? nsf - p.sh : 1)) return;
  p.NW = NW; p.Lr = Lr; p.len0 = len0; p.E = NW + 2;
  // tables
  {
    std::vector<double> ra(ASTR_LMAX, alpha), rc(ASTR_LMAX, alpha);
    LocalFac f = local_factor(ra, rc, 0, ASTR_LMAX, false, false);
    for (int i = 0; i < ASTR_LMAX; ++i) { p.reg.m[i] = f.m[i]; p.reg.g[i] = f.g[i]; p.reg.ac1[i] = f.ac1[i]; p.reg.ev[i] = f.ev[i]; }
  }
  std::vector<LocalFac> fac(p.E);
  fac[0] = local_factor(a, c, 0, p.sh, true, false);
  for (int w = 0; w < NW; ++w) fac[w + 1] = local_factor(a, c, plan_chunk_row(p, w), w == 0 ? len0 : Lr, false, false);
  fac[p.E - 1] = local_factor(a, c, N - p.st, p.st, false, true);
  auto fill = [](SpecTab& t, const LocalFac& f, int len) {
    t.len = len; t.pad = 0;
    for (int i = 0; i < ASTR_SMAX; ++i) {
      t.m[i] = i < len ? f.m[i] : 1.0; t.g[i] = i < len ? f.g[i] : 0.0;
      t.ac1[i] = i < len ? f.ac1[i] : 0.0; t.ev[i] = i < len ? f.ev[i] : 0.0;
    }
  };
  fill(p.head, fac[0], p.sh);
  fill(p.tail, fac[p.E - 1], p.st);
  for (int e = 0; e < ASTR_EMAX; ++e) p.el[e] = ElemTab{0, 0, 0, 0, 0, 0, 1, 0};
  double Qprev = 0.0;
  for (int e = 0; e < p.E; ++e) {          // elimination from the first end
    const LocalFac& f = fac[e];
    ElemTab& t = p.el[e];
    const double mu = 1.0 / (1.0 - f.vh * Qprev);
    const double B = mu * f.wh;
    t.gamma = f.vt * Qprev * mu;
    t.K = f.vt * mu;
    t.Q = f.wt + f.vt * Qprev * B;
    Qprev = t.Q;
  }
  double Qnext = 0.0;
  for (int e = p.E - 1; e >= 0; --e) {     // elimination from the last end (roles of v and w swapped)
    const LocalFac& f = fac[e];
    ElemTab& t = p.el[e];
    const double mu = 1.0 / (1.0 - f.wt * Qnext);
    const double B = mu * f.vt;
    t.gammap = f.wh * Qnext * mu;
    t.Kp = f.wh * mu;
    t.Qp = f.vh + f.wh * Qnext * B;
    Qnext = t.Qp;
  }
  for (int e = 1; e < p.E; ++e) p.el[e].D = 1.0 / (1.0 - p.el[e - 1].Q * p.el[e].Qp);
  p.ok = 1;
}
