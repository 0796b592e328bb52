// tests/emul_sweep2.cpp -- TEST INFRASTRUCTURE.  Drives the math of astr_b200/csrc/linecore.h
// (the functions the sm_100a kernels in sweep2.cu are built from) on the CPU, one pencil at a
// time, thread by thread in the kernels' phase order.  tests/test_host_logic.py compares the
// result with the oracle's single Thomas sweep.
#include "../astr_b200/csrc/linecore.h"
#include <vector>
#include <algorithm>

namespace {

template <int OP>
int run_line(const LinePlan& pl, const FilterCoef& fc, const double* f5 /* f5[k] = f(k-5) */, double* out, int use_fast) {
  constexpr int H = OpT<OP>::H;
  constexpr int WN = ASTR_LMAX + 2 * H;
  const int n = pl.n, NW = pl.NW, E = pl.E;
  const bool p0 = (pl.ntype == 1 || pl.ntype == 4), pm = (pl.ntype == 2 || pl.ntype == 4);
  auto F = [&](int node) { return f5[std::min(std::max(node, -5), n + 5) + 5]; };
  std::vector<double> S(E, 0.0), SP(E, 0.0);
  std::vector<std::vector<double>> e(NW, std::vector<double>(ASTR_LMAX, 0.0));
  double he[ASTR_SMAX] = {0, 0}, te[ASTR_SMAX] = {0, 0};
  // ---- phase A: every thread (here: chunk) on its own
  for (int w = 0; w < NW; ++w) {
    const bool head = (w == 0), tail = (w == NW - 1);
    const int len = head ? pl.len0 : pl.Lr;
    const int node0 = pl.first_node + plan_chunk_row(pl, w);
    double wv[WN];
    for (int s = 0; s < WN; ++s) wv[s] = (s < len + 2 * H) ? F(node0 - H + s) : 0.0;
    double ov[4] = {0, 0, 0, 0};
    if (head) {
      double hw[14], sf[5];
      for (int k = 0; k < 14; ++k) hw[k] = F(k - 5);
      closure_head<OP>(hw, p0, fc, sf);
      for (int k = 0; k < pl.nsf - pl.sh; ++k) ov[k] = sf[pl.sh + k];
      double yh, yt;
      spec_forward(pl.head, sf, he, yh, yt);
      S[0] = fma_(pl.el[0].gamma, yh, yt); SP[0] = fma_(pl.el[0].gammap, yt, yh);
    }
    if (tail) {
      double tw[14], sl[5];
      for (int k = 0; k < 14; ++k) tw[k] = F(n - 8 + k);
      closure_tail<OP>(tw, pm, fc, sl);
      const int nov_t = pl.nsl - pl.st;
      for (int k = 0; k < nov_t; ++k) ov[k] = sl[k];
      double yh, yt;
      spec_forward(pl.tail, sl + nov_t, te, yh, yt);
      S[E - 1] = fma_(pl.el[E - 1].gamma, yh, yt); SP[E - 1] = fma_(pl.el[E - 1].gammap, yt, yh);
    }
    double ee[ASTR_LMAX], yh, yt;
    if (head) chunk_forward<OP, ROLE_HEAD>(pl.reg, fc, wv, len, p0, ov, ee, yh, yt);
    else if (tail) chunk_forward<OP, ROLE_TAIL>(pl.reg, fc, wv, len, pm, ov, ee, yh, yt);
    else chunk_forward<OP, ROLE_MID>(pl.reg, fc, wv, len, false, ov, ee, yh, yt);
    for (int s = 0; s < ASTR_LMAX; ++s) e[w][s] = ee[s];
    S[w + 1] = fma_(pl.el[w + 1].gamma, yh, yt); SP[w + 1] = fma_(pl.el[w + 1].gammap, yt, yh);
  }
  // ---- barrier; phase B
  auto GS = [&](int el) { return S[el]; };
  auto GP = [&](int el) { return SP[el]; };
  auto put = [&](int node, double x) { if (node >= 0 && node <= n) out[node] = x; };
  for (int w = 0; w < NW; ++w) {
    const bool head = (w == 0), tail = (w == NW - 1);
    const int len = head ? pl.len0 : pl.Lr;
    const int node0 = pl.first_node + plan_chunk_row(pl, w);
    const int me = w + 1;
    const ScanOut so = reduced_scan(pl, GS, GP, me);
    const double tp = scan_t_prev(pl, me, so.Pm1, so.Pb0), hn = scan_h_next(pl, me, so.Pm, so.Pb1);
    // head block = element 0: P(0) = Pm1 and P'(1) = Pb0 of element 1; tail block likewise
    const double hn2 = scan_h_next(pl, 0, so.Pm1, so.Pb0);
    const double tp2 = scan_t_prev(pl, E - 1, so.Pm, so.Pb1);
    double ee[ASTR_LMAX];
    for (int s = 0; s < ASTR_LMAX; ++s) ee[s] = e[w][s];
    auto st = [&](int s, double x) { put(node0 + s, x); };
    if (head) chunk_back<ROLE_HEAD>(pl.reg, ee, len, tp, hn, st);
    else chunk_back<ROLE_MID>(pl.reg, ee, len, tp, hn, st);
    if (head) {
      double x[ASTR_SMAX];
      spec_back(pl.head, he, 0.0, hn2, x);
      for (int k = 0; k < pl.sh; ++k) put(pl.first_node + k, x[k]);
    }
    if (tail && NW > 1) {
      double x[ASTR_SMAX];
      spec_back(pl.tail, te, tp2, 0.0, x);
      for (int k = 0; k < pl.st; ++k) put(pl.first_node + pl.nrows - pl.st + k, x[k]);
    }
  }
  return 0;
}

}  // namespace

extern "C" int emul_line(int optype, int ntype, int n, double alfa, int maxw, int use_fast, const double* f5, double* out,
                         int* info /* NW, Lr, len0 */) {
  std::vector<double> a, c;
  int first_node, nsf, nsl;
  build_lhs(optype, ntype, n, alfa, a, c, first_node, nsf, nsl);
  LinePlan pl;
  build_line_plan(pl, optype, ntype, n, first_node, nsf, nsl, a, c, maxw);
  if (!pl.ok) return 1;
  info[0] = pl.NW; info[1] = pl.Lr; info[2] = pl.len0;
  FilterCoef fc;
  build_filter_coef(fc, alfa, 1.11, 0.98);
  if (optype == 0) return run_line<0>(pl, fc, f5, out, use_fast);
  return run_line<1>(pl, fc, f5, out, use_fast);
}
