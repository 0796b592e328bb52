// astr_b200/csrc/linecore.h -- the register-resident partitioned line solve (sweep2.cu).
//
// Same linear systems as the reference's compact operators -- `fds%central`
// (src/derivative.F90:171-306), `compact_filter` (src/filter.F90:112-285), `flux_compact`
// (src/flux.F90:125-266): unit-diagonal tridiagonal rows a(r) x(r-1) + x(r) + c(r) x(r+1) = d(r)
// (src/commfunc.F90:752-813) -- but factorised for a GPU.  The rows are cut into ELEMENTS
//
//     [head block | regular chunk 0 | regular chunk 1 | ... | regular chunk NW-1 | tail block]
//
// * regular chunks: ASTR_LMAX rows with the interior left- AND right-hand side.  Every chunk
//   is factorised from a FRESH start, so all chunks of all pencils share ONE set of coefficient
//   tables (m, g, ac1, ev), which the kernels read as immediate constant-bank operands, and all
//   chunk threads run the same straight-line code.
// * head block: the closure rows of the first end (explicit interface rows, wall rows, filter
//   end rows) plus the rows that do not fill a whole chunk, when the block then has at most 8 rows;
//   tail block: the closure rows of the last end (+ at most one interior row, see `align_even`).
//   Generic tables, one dedicated warp.
// * SHORT first chunk: when head block + left-over rows would exceed 8 rows, the left-over rows
//   become regular chunk 0 with ls0 < ASTR_LMAX rows instead.  A fresh-start factorisation does not
//   depend on where the chunk ends, so the short chunk runs the same tables truncated at ls0 (the
//   slots behind pass values through); only its element coefficients differ.  The head block then
//   stays on the 8-slot code path for every line length (round 1-2a: up to 40 slots, one warp, the
//   straggler of every bundle on e.g. 513-row wall-bounded lines).
//
// A thread owns one element of one pencil in registers:
//   pass 1  e(s) = d(s) m(s) - e(s-1) g(s)                       local forward elimination
//   pass 2  y(s) = e(s) - ac1(s) y(s+1)                          local solution, zero carries;
//           only its head yh = y(0) and tail yt = e(len-1) leave the thread
//   scan    the heads/tails of all elements of a pencil obey a block-tridiagonal system whose
//           block LU collapses to two scalar recurrences (P forward, P' backward)
//   pass 3  x(s) = (e(s) - ev(s) t_prev) - ac1(s) x(s+1),  x(len) = h_next
// with t_prev the true solution at the last row of the previous element and h_next at the first
// row of the next one.  The recurrences P(e) = S(e) - K(e) P(e-1) contract by |K| (the product of
// the elimination factors over a whole chunk: 1e-14 per chunk for the derivative, 1e-3 for the
// filter at alfa = 0.49), so each thread evaluates them over the W nearest elements only, with W
// chosen by the host so that the dropped terms are below 1e-20 relative -- four decades under the
// rounding of a double.  Everything else is exact algebra; differences from the reference's
// single Thomas sweep are rounding only (tests/test_host_logic.py drives these functions on the
// CPU against the oracle).
//
// Header shared by nvcc (device + host) and g++ (tests/emul_sweep2.cpp).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define ASTR_HD __host__ __device__ __forceinline__
#else
#define ASTR_HD inline
#endif

#define ASTR_LMAX 34     // rows of a regular chunk (even: all chunk windows of an i line share one 16-byte alignment)
#define ASTR_NWMAX 15    // regular chunks per line (warps 0..NW-1; warp NW owns the head and tail blocks)
#define ASTR_HS 40       // slots of the head block: <= 5 closure rows + ASTR_LMAX remainder rows (+1 alignment row)
#define ASTR_TS 6        // slots of the tail block: <= 5 closure rows + 1 interior row
#define ASTR_EMAX (ASTR_NWMAX + 2)
#define ASTR_WPAD 8      // most elements the truncated reduced scan may look at on each side

struct RegTab {          // fresh-start factorisation of ASTR_LMAX interior rows (a, 1, c)
  double m[ASTR_LMAX], g[ASTR_LMAX], ac1[ASTR_LMAX], ev[ASTR_LMAX];
};
// SHORT first chunk (build_line_plan): the first ls0 slots are the rows (the fresh-start tables truncated), the rest
// NEUTRAL slots, as in the head / tail blocks below: the code stays straight-line with immediate operands
struct ShortTab {
  double m[ASTR_LMAX], g[ASTR_LMAX], q[ASTR_LMAX], ac1[ASTR_LMAX], ev[ASTR_LMAX];
};
// Head / tail block: rows first, then NEUTRAL slots (m = 0, g = -1, q = 0, ac1 = -1, ev = 0) that
// pass e forward and x backward unchanged, so the unrolled code needs no length guards.
template <int S> struct SpecTab {
  int len, nclos;        // rows of the block; rows among them with a closure right-hand side
  double m[S], g[S], q[S], ac1[S], ev[S];
};
struct ElemTab {         // reduced-system coefficients of one element (reduced_scan below)
  double gamma, K, Q, gammap, Kp, Qp, D, pad;
};
struct LinePlan {
  int optype, ntype, n, first_node, nrows;
  int nsf, nsl;          // rows with a closure RIGHT-hand side at each end
  int sh, st;            // rows of the head / tail block (sh >= nsf, st = nsl or nsl + 1)
  int NW;                // regular chunks
  int E;                 // NW + 2
  int W;                 // reduced scan: elements looked at on each side (<= ASTR_WPAD)
  int ok;                // 0: this line cannot be handled by sweep2 (fall back to sweep.cu)
  int ls0;               // rows of regular chunk 0: ASTR_LMAX, or fewer when the rows that do not fill whole chunks
                         // are too many for the short head block (they then form a SHORT first chunk, see build_line_plan)
  RegTab reg;
  ShortTab sreg;         // chunk 0 when ls0 < ASTR_LMAX
  SpecTab<ASTR_HS> head;
  SpecTab<ASTR_TS> tail;
  ElemTab el[ASTR_EMAX + 2 * ASTR_WPAD];   // element e at index e + ASTR_WPAD; neutral entries around
};

// first row of regular chunk w
ASTR_HD int plan_chunk_row(const LinePlan& p, int w) { return w == 0 ? p.sh : p.sh + p.ls0 + (w - 1) * ASTR_LMAX; }

// ASTR_NO_FMA (the -fmad=false debug library, `make nofma`): every product is rounded before it is added, as
// in the oracle (which is compiled with -ffp-contract=off, like the reference's golden run).  Whatever then
// still differs from the oracle is re-association, not contraction (tests/test_gpu_nofma.py).
ASTR_HD double fma_(double a, double b, double c) {
#if defined(ASTR_NO_FMA) && defined(__CUDA_ARCH__)
  return __dadd_rn(__dmul_rn(a, b), c);
#elif defined(__CUDA_ARCH__)
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
// S(idx), SP(idx): published sums at PADDED element index idx = e + ASTR_WPAD (zero outside 0..E-1).
struct ScanOut { double t_prev, h_next; };

template <class GETS, class GETSP>
ASTR_HD ScanOut reduced_scan(const LinePlan& pl, GETS S, GETSP SP, int me) {
  const int W = pl.W;
  const int c = me + ASTR_WPAD;
  double pf = 0.0, pf1 = 0.0;
  for (int t = W; t >= 0; --t) {
    pf1 = pf;
    pf = fma_(-pl.el[c - t].K, pf, S(c - t));
  }
  double pb = 0.0, pb1 = 0.0;
  for (int t = W; t >= 0; --t) {
    pb1 = pb;
    pb = fma_(-pl.el[c + t].Kp, pb, SP(c + t));
  }
  ScanOut o;
  o.t_prev = pl.el[c].D * fma_(-pl.el[c - 1].Q, pb, pf1);        // element 0: P(-1) = 0, Q(-1) = 0
  o.h_next = pl.el[c + 1].D * fma_(-pl.el[c + 1].Qp, pf, pb1);   // element E-1: P'(E) = 0, Q'(E) = 0
  return o;
}

// ---- right-hand sides -----------------------------------------------------------------
// coefficient tables of src/filter.F90:299-432 (only the rows the hot path reads) and the interior
// coefficients of compact_flux_rhs (src/flux.F90:247-262)
struct FilterCoef {
  double coef6i[4], coef8i[5], coef10i[6];
  double coefb[4][9];
  double coefh[3][11];
  double flx[4];
};

template <int OP> struct OpT;
// H: half width of the interior stencil; HB: reach of the head block's window in front of the first row
// CR: window slots the closure rows of the first end may read
template <> struct OpT<0> { static constexpr int H = 2, HB = 3, CR = 7; };    // OP_DERIV (interface row: 7-point explicit)
template <> struct OpT<1> { static constexpr int H = 5, HB = 5, CR = 14; };   // OP_FILTER
template <> struct OpT<2> { static constexpr int H = 2, HB = 2, CR = 6; };    // OP_FLUXP
template <> struct OpT<3> { static constexpr int H = 2, HB = 2, CR = 6; };    // OP_FLUXM

// first row of the system as a node index: fd_scheme_initiate (src/derivative.F90:74-93),
// compact_filter_initiate (src/filter.F90:44-71), compact_flux_initiate (src/flux.F90:44-70)
template <int OP> ASTR_HD constexpr int op_first_node(bool p0) {
  return OP == 0 ? (p0 ? 0 : -1) : OP == 1 ? (p0 ? 0 : -3) : (p0 ? -1 : -2);
}
// node of the row in front of the closure rows of the last end, relative to n
template <int OP> ASTR_HD constexpr int op_tail_interior_off(bool pm) {
  return OP == 0 ? (pm ? -2 : 0) : OP == 1 ? (pm ? -5 : 0) : (pm ? -2 : 0);
}

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
  } else if (OP == 2) {
    // src/flux.F90:247-253
    return fc.flx[0] * WS(-1) + fc.flx[1] * WS(0) + fc.flx[2] * WS(1) + fc.flx[3] * WS(2);
  } else if (OP == 3) {
    // src/flux.F90:255-261
    return fc.flx[0] * WS(2) + fc.flx[1] * WS(1) + fc.flx[2] * WS(0) + fc.flx[3] * WS(-1);
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

// Closure rows at the first end.  hw[k] = f(first_node - HB + k): the head block's window (HB: how far the
// closure and interior stencils reach in front of the first row).  sf[k] is row k.
template <int OP, bool P0, int HWN>
ASTR_HD void closure_head(const double (&hw)[HWN], const FilterCoef& fc, double (&sf)[5]) {
  constexpr int HB = OpT<OP>::HB;
  constexpr int FN = op_first_node<OP>(P0);
#define F(node) hw[(node) - FN + HB]
#pragma unroll
  for (int k = 0; k < 5; ++k) sf[k] = 0.0;
  if (OP == 0) {
    if (P0) {    // src/derivative.F90:230-248
      sf[0] = -2.5 * F(0) + 2.0 * F(1) + 0.5 * F(2);
      sf[1] = 0.75 * (F(2) - F(0));
    } else {     // :250-260, row of ghost node -1
      sf[0] = 0.75 * (F(0) - F(-2)) - 0.15 * (F(1) - F(-3)) + (1.0 / 60.0) * (F(2) - F(-4));
    }
  } else if (OP == 2 || OP == 3) {
    if (P0) {    // src/flux.F90:189-199, rows -1 and 0
      sf[0] = 2.5 * F(0) + 0.5 * F(1);
      sf[1] = 0.75 * F(0) + 0.75 * F(1);
    } else {     // :205-209, row -2: 6th-order explicit
      const double var1 = F(-2) + F(-1), var2 = F(-3) + F(0), var3 = F(-4) + F(1);
      sf[0] = (37.0 / 60.0) * var1 - (2.0 / 15.0) * var2 + (1.0 / 60.0) * var3;
    }
  } else {
    if (P0) {    // src/filter.F90:176-204
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

// Closure rows at the last end.  tw[k] = f(n - 10 + k), k = 0..15 (nodes n-10..n+5).
// sl[k] is row nrows - nsl + k.
template <int OP, bool PM>
ASTR_HD void closure_tail(const double (&tw)[16], const FilterCoef& fc, double (&sl)[5]) {
#define F(off) tw[(off) + 10]    // F(off) = f(n + off)
#pragma unroll
  for (int k = 0; k < 5; ++k) sl[k] = 0.0;
  if (OP == 0) {
    if (PM) {    // src/derivative.F90:264-281
      sl[0] = 0.75 * (F(0) - F(-2));
      sl[1] = 2.5 * F(0) - 2.0 * F(-1) - 0.5 * F(-2);
    } else {     // :283-292, row of ghost node n+1
      sl[0] = 0.75 * (F(2) - F(0)) - 0.15 * (F(3) - F(-1)) + (1.0 / 60.0) * (F(4) - F(-2));
    }
  } else if (OP == 2 || OP == 3) {
    if (PM) {    // src/flux.F90:217-225, rows n-1 and n
      sl[0] = 0.75 * F(0) + 0.75 * F(-1);
      sl[1] = 2.5 * F(0) + 0.5 * F(-1);
    } else {     // :231-236, row n+1
      const double var1 = F(1) + F(2), var2 = F(0) + F(3), var3 = F(-1) + F(4);
      sl[0] = (37.0 / 60.0) * var1 - (2.0 / 15.0) * var2 + (1.0 / 60.0) * var3;
    }
  } else {
    if (PM) {    // src/filter.F90:222-249 ; rows n-4, n-3, n-2, n-1, n
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

// ---- regular chunk ------------------------------------------------------------------------
// passes 1 and 2: wv[s] = f(node0 - H + s), s = 0 .. ASTR_LMAX + 2H - 1
template <int OP, int WN>
ASTR_HD void chunk_forward(const RegTab& t, const FilterCoef& fc, const double (&wv)[WN], double (&e)[ASTR_LMAX],
                           double& yh, double& yt) {
  constexpr int L = ASTR_LMAX;
#ifdef ASTR_SKELETON   // experiment build (`make skel`): data movement only, no arithmetic
#pragma unroll
  for (int s = 0; s < L; ++s) e[s] = wv[s + OpT<OP>::H];
  yh = 0.0; yt = 0.0;
  return;
#endif
#pragma unroll
  for (int s = 0; s < L; ++s) {
    const double d = reg_rhs<OP>(wv, s, fc);
    e[s] = (s == 0) ? d : fma_(-e[s > 0 ? s - 1 : 0], t.g[s], d * t.m[s]);
  }
  double y = e[L - 1];
#pragma unroll
  for (int s = L - 2; s >= 0; --s) y = fma_(-t.ac1[s], y, e[s]);
  yh = y; yt = e[L - 1];
}
// the SHORT chunk: same passes over the ShortTab (neutral slots behind the rows carry e forward and leave y alone;
// their right-hand sides are finite values of the following chunk's rows, multiplied by zero)
template <int OP, int WN>
ASTR_HD void chunk_forward_short(const ShortTab& t, const FilterCoef& fc, const double (&wv)[WN], double (&e)[ASTR_LMAX],
                                 double& yh, double& yt) {
  constexpr int L = ASTR_LMAX;
#pragma unroll
  for (int s = 0; s < L; ++s) {
    const double d = reg_rhs<OP>(wv, s, fc);
    e[s] = (s == 0) ? d : fma_(-e[s > 0 ? s - 1 : 0], t.g[s], d * t.m[s]);
  }
  double y = 0.0;
#pragma unroll
  for (int s = L - 1; s >= 0; --s) y = fma_(-t.ac1[s], y, e[s] * t.q[s]);
  yh = y; yt = e[L - 1];
}
// pass 3: ST(s, x) receives the solution of row s (descending)
template <class ST>
ASTR_HD void chunk_back(const RegTab& t, const double (&e)[ASTR_LMAX], double t_prev, double h_next, ST st) {
#ifdef ASTR_SKELETON
#pragma unroll
  for (int s = ASTR_LMAX - 1; s >= 0; --s) st(s, e[s]);
  return;
#endif
  double x = h_next;
#pragma unroll
  for (int s = ASTR_LMAX - 1; s >= 0; --s) {
    x = fma_(-t.ac1[s], x, fma_(-t.ev[s], t_prev, e[s]));
    st(s, x);
  }
}
// SHORT chunk: ST is called for every slot; the caller keeps the slots s < ls0
template <class ST>
ASTR_HD void chunk_back_short(const ShortTab& t, const double (&e)[ASTR_LMAX], double t_prev, double h_next, ST st) {
  double x = h_next;
#pragma unroll
  for (int s = ASTR_LMAX - 1; s >= 0; --s) {
    x = fma_(-t.ac1[s], x, fma_(-t.ev[s], t_prev, e[s] * t.q[s]));
    st(s, x);
  }
}

// ---- head / tail block ----------------------------------------------------------------------
// The slots are processed in groups of ASTR_SG; groups that lie entirely behind the rows of the block
// (warp-uniform test on `len`) are skipped, neutral slots inside the last group pass values through.
#define ASTR_SG 8
template <int S> struct SpecGroups { static constexpr int G = (S % ASTR_SG == 0) ? ASTR_SG : S, N = S / G; };

// d[s]: right-hand side of slot s (any finite value on neutral slots of a processed group)
// S: slots processed (the block has at most S rows); TS: slots of the table
template <int S, int TS>
ASTR_HD void spec_forward(const SpecTab<TS>& t, const double (&d)[S], double (&e)[S], double& yh, double& yt) {
  static_assert(S <= TS, "slots");
  constexpr int G = SpecGroups<S>::G, NG = SpecGroups<S>::N;
  const int len = t.len;
  yt = 0.0;
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g * G < len) {
#pragma unroll
      for (int s = g * G; s < g * G + G; ++s)
        e[s] = (s == 0) ? d[0] * t.m[0] : fma_(-e[s > 0 ? s - 1 : 0], t.g[s], d[s] * t.m[s]);
      yt = e[g * G + G - 1];
    }
  }
  double y = 0.0;
#pragma unroll
  for (int g = NG - 1; g >= 0; --g) {
    if (g * G < len) {
#pragma unroll
      for (int s = g * G + G - 1; s >= g * G; --s) y = fma_(-t.ac1[s], y, e[s] * t.q[s]);
    }
  }
  yh = y;
}
template <int S, int TS, class ST>
ASTR_HD void spec_back(const SpecTab<TS>& t, const double (&e)[S], double t_prev, double h_next, ST st) {
  constexpr int G = SpecGroups<S>::G, NG = SpecGroups<S>::N;
  const int len = t.len;
  double x = h_next;
#pragma unroll
  for (int g = NG - 1; g >= 0; --g) {
    if (g * G < len) {
#pragma unroll
      for (int s = g * G + G - 1; s >= g * G; --s) {
        x = fma_(-t.ac1[s], x, fma_(-t.ev[s], t_prev, e[s] * t.q[s]));
        st(s, x);
      }
    }
  }
}
// right-hand sides of the head block: closure rows first, interior rows behind them (groups behind `len` skipped)
template <int OP, bool P0, int HWN, int S>
ASTR_HD void head_rhs(const double (&hw)[HWN], const FilterCoef& fc, int nsf, int len, double (&d)[S]) {
  constexpr int G = SpecGroups<S>::G, NG = SpecGroups<S>::N;
  double sf[5];
  closure_head<OP, P0>(hw, fc, sf);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g * G < len) {
#pragma unroll
      for (int s = g * G; s < g * G + G; ++s) {
        double v = reg_rhs<OP>(hw, s + (OpT<OP>::HB - OpT<OP>::H), fc);
        if (s < 5) { if (s < nsf) v = sf[s]; }
        d[s] = v;
      }
    }
  }
}
// tail block: (one interior row when st == nsl + 1,) then the closure rows
template <int OP, bool PM>
ASTR_HD void tail_rhs(const double (&tw)[16], const FilterCoef& fc, int extra, double (&d)[ASTR_TS]) {
  constexpr int H = OpT<OP>::H;
  constexpr int TO = op_tail_interior_off<OP>(PM);
  double sl[5];
  closure_tail<OP, PM>(tw, fc, sl);
  // the interior stencil of node n + TO reads tw[TO + 10 - H .. TO + 10 + H]
  double ww[2 * H + 1];
#pragma unroll
  for (int k = 0; k <= 2 * H; ++k) ww[k] = tw[TO + 10 - H + k];
  const double di = reg_rhs<OP>(ww, 0, fc);
  d[0] = extra ? di : sl[0];
#pragma unroll
  for (int s = 1; s < ASTR_TS; ++s) d[s] = extra ? sl[s - 1] : (s < 5 ? sl[s < 5 ? s : 4] : 0.0);
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

// src/filter.F90:299-432 ; flux coefficients src/flux.F90:247-262 for bfacmpld = b
inline void build_filter_coef(FilterCoef& fc, double alfa, double bh, double bb, double b = 0.0) {
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
  // compact_flux_rhs interior coefficients (src/flux.F90:247-262)
  fc.flx[0] = 1.0 / 18.0 - (1.0 / 36.0) * b; fc.flx[1] = 19.0 / 18.0 - (9.0 / 36.0) * b;
  fc.flx[2] = 5.0 / 9.0 + (9.0 / 36.0) * b;  fc.flx[3] = (1.0 / 36.0) * b;
}

template <int S>
inline void fill_spec(SpecTab<S>& t, const LocalFac& f, int len, int nclos) {
  t.len = len; t.nclos = nclos;
  for (int i = 0; i < S; ++i) {
    const bool r = i < len;
    t.m[i] = r ? f.m[i] : 0.0; t.g[i] = r ? f.g[i] : -1.0; t.q[i] = r ? 1.0 : 0.0;
    t.ac1[i] = r ? f.ac1[i] : -1.0; t.ev[i] = r ? f.ev[i] : 0.0;
  }
}

// a, c: the LHS rows exactly as the reference sets them (fd_scheme_initiate
// src/derivative.F90:63-158, compact_filter_initiate src/filter.F90:31-100, compact_flux_initiate
// src/flux.F90:32-118).  maxw: most regular chunks a line may be cut into.  align_even (i lines):
// node0 - H + 6 of every chunk must be even, so that the chunk windows start on a 16-byte boundary of
// the shared-memory line (node -6 sits at an aligned position); one interior row may move to the tail block.
inline void build_line_plan(LinePlan& p, int optype, int ntype, int n, int first_node, int nsf, int nsl,
                            const std::vector<double>& a, const std::vector<double>& c, int maxw, bool align_even) {
  p = LinePlan();
  p.optype = optype; p.ntype = ntype; p.n = n; p.first_node = first_node;
  p.nrows = (int)a.size(); p.nsf = nsf; p.nsl = nsl;
  const int N = p.nrows, L = ASTR_LMAX;
  const int H = (optype == 1) ? 5 : 2;
  p.ok = 0;
  const int nreg = N - nsf - nsl;
  if (nreg < L || n < 12) return;
  const double aa = a[nsf], cc = c[nsf];
  for (int r = nsf; r < N - nsl; ++r)
    if (a[r] != aa || c[r] != cc) return;           // not a Toeplitz interior: not ours
  if (maxw > ASTR_NWMAX) maxw = ASTR_NWMAX;
  int NW = nreg / L;
  if (NW > maxw) return;
  int rem = nreg - NW * L, rem_t = 0, rem_h = 0, ls0 = L;
  // (A) the left-over rows join the head block
  int nwA = NW, remA = rem, remtA = 0;
  bool okA = true;
  if (align_even && (((first_node + nsf + remA - H + 6) & 1) != 0)) {
    if (remA == 0) { nwA -= 1; remA = L; }
    if (nwA < 1) okA = false;
    remA -= 1; remtA = 1;
  }
  // (B) SHORT first chunk: head block = closure rows (+1 row for the window alignment of i lines), chunk 0 = the
  // left-over rows (an even number for i lines: one more row may move to the tail block)
  int nwB = NW, R = rem, remhB = 0, remtB = 0;
  if (align_even && (((first_node + nsf - H + 6) & 1) != 0)) remhB = 1;
  if (R - remhB < 2 && nwB > 1) { nwB -= 1; R += L; }          // too few left-over rows: borrow a chunk
  int lsB = R - remhB;
  if (align_even && (lsB & 1)) { lsB -= 1; remtB = 1; }
  const bool okB = lsB >= 2 && lsB < L && nwB + 1 <= maxw && nwB >= 1;
  if (okA && nsf + remA <= 8) { NW = nwA; rem_h = remA; rem_t = remtA; }
  else if (okB) { NW = nwB + 1; rem_h = remhB; rem_t = remtB; ls0 = lsB; }
  else if (okA) { NW = nwA; rem_h = remA; rem_t = remtA; }      // long head block (one warp, up to ASTR_HS slots)
  else return;
  p.sh = nsf + rem_h; p.st = nsl + rem_t; p.NW = NW; p.E = NW + 2; p.ls0 = ls0;
  if (p.sh > ASTR_HS || p.st > ASTR_TS) return;
  // tables
  {
    std::vector<double> ra(L, aa), rc(L, cc);
    const LocalFac f = local_factor(ra, rc, 0, L, false, false);
    for (int i = 0; i < L; ++i) { p.reg.m[i] = f.m[i]; p.reg.g[i] = f.g[i]; p.reg.ac1[i] = f.ac1[i]; p.reg.ev[i] = f.ev[i]; }
  }
  std::vector<LocalFac> fac(p.E);
  fac[0] = local_factor(a, c, 0, p.sh, true, false);
  for (int w = 0; w < NW; ++w) fac[w + 1] = local_factor(a, c, plan_chunk_row(p, w), w == 0 ? p.ls0 : L, false, false);
  fac[p.E - 1] = local_factor(a, c, N - p.st, p.st, false, true);
  for (int i = 0; i < L; ++i) {            // SHORT first chunk: rows, then neutral slots (unused when ls0 == L)
    const bool r = i < p.ls0;
    const LocalFac& f = fac[1];
    p.sreg.m[i] = r ? f.m[i] : 0.0; p.sreg.g[i] = r ? f.g[i] : -1.0; p.sreg.q[i] = r ? 1.0 : 0.0;
    p.sreg.ac1[i] = r ? f.ac1[i] : -1.0; p.sreg.ev[i] = r ? f.ev[i] : 0.0;
  }
  fill_spec(p.head, fac[0], p.sh, nsf);
  fill_spec(p.tail, fac[p.E - 1], p.st, nsl);
  const int NE = ASTR_EMAX + 2 * ASTR_WPAD;
  for (int e = 0; e < NE; ++e) p.el[e] = ElemTab{0, 0, 0, 0, 0, 0, 1, 0};
  double Qprev = 0.0;
  for (int e = 0; e < p.E; ++e) {          // elimination from the first end
    const LocalFac& f = fac[e];
    ElemTab& t = p.el[e + ASTR_WPAD];
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
    ElemTab& t = p.el[e + ASTR_WPAD];
    const double mu = 1.0 / (1.0 - f.wt * Qnext);
    const double B = mu * f.vt;
    t.gammap = f.wh * Qnext * mu;
    t.Kp = f.wh * mu;
    t.Qp = f.vh + f.wh * Qnext * B;
    Qnext = t.Qp;
  }
  // P(E-1) and P'(0) are never used (they would only feed the interfaces outside the system)
  p.el[p.E - 1 + ASTR_WPAD].K = 0.0;
  p.el[ASTR_WPAD].Kp = 0.0;
  for (int e = 1; e < p.E; ++e) p.el[e + ASTR_WPAD].D = 1.0 / (1.0 - p.el[e - 1 + ASTR_WPAD].Q * p.el[e + ASTR_WPAD].Qp);
  // truncation window of the reduced scan: P(e) = S(e) - K(e) P(e-1).  A thread of element e needs P(e) and
  // P(e-1); evaluated over the elements e-W..e, the first dropped term of P(e-1) carries the factor
  // K(e-1) K(e-2) ... K(e-W).  Smallest W that keeps every product of W consecutive factors under 1e-20.
  int W = 1;
  for (; W <= p.E; ++W) {
    double worst = 0.0;
    for (int e = 0; e < p.E; ++e) {
      double pk = 1.0, pkp = 1.0;
      for (int k = 0; k < W; ++k) {
        pk *= (e - k >= 0) ? std::fabs(p.el[e - k + ASTR_WPAD].K) : 0.0;
        pkp *= (e + k < p.E) ? std::fabs(p.el[e + k + ASTR_WPAD].Kp) : 0.0;
      }
      worst = std::fmax(worst, std::fmax(pk, pkp));
    }
    if (worst < 1e-20) break;
  }
  if (W > ASTR_WPAD) return;               // slowly decaying coupling (alfa very close to 1/2): fall back
  p.W = W;
  p.ok = 1;
}
