// tests/emul_sweep2.cpp -- TEST INFRASTRUCTURE.  Drives the math of astr_b200/csrc/linecore.h
// (the functions the sm_100a kernels in sweep2.cu are built from) on the CPU, one pencil at a
// time, "thread" by "thread" in the kernels' phase order: every regular chunk and the head / tail
// blocks run passes 1-2 and publish S / S', then (barrier) every element runs the truncated
// reduced scan and pass 3.  tests/test_host_logic.py compares the result with the oracle's single
// Thomas sweep (df_compact / compact_filter / flux_compact).
#include "../astr_b200/csrc/linecore.h"
#include <vector>
#include <algorithm>

namespace {

template <int OP, bool P0, bool PM>
int run_line(const LinePlan& pl, const FilterCoef& fc, const double* f5 /* f5[k] = f(k-5) */, double* out /* out[r] = row r */) {
  constexpr int H = OpT<OP>::H;
  constexpr int L = ASTR_LMAX;
  constexpr int WN = L + 2 * H;
  constexpr int HB = OpT<OP>::HB;
  constexpr int HWN = ASTR_HS + HB + H;
  const int n = pl.n, NW = pl.NW, E = pl.E;
  auto F = [&](int node) { return (node >= -5 && node <= n + 5) ? f5[node + 5] : 0.0; };
  std::vector<double> S(ASTR_EMAX + 2 * ASTR_WPAD, 0.0), SP(ASTR_EMAX + 2 * ASTR_WPAD, 0.0);
  std::vector<std::vector<double>> e(NW, std::vector<double>(L, 0.0));
  double he[ASTR_HS] = {0}, te[ASTR_TS] = {0};
  auto publish = [&](int el, double yh, double yt) {
    S[el + ASTR_WPAD] = fma_(pl.el[el + ASTR_WPAD].gamma, yh, yt);
    SP[el + ASTR_WPAD] = fma_(pl.el[el + ASTR_WPAD].gammap, yt, yh);
  };
  // ---- phase A/B: every element on its own
  {
    double hw[HWN];
    for (int k = 0; k < HWN; ++k) hw[k] = (k < std::max(OpT<OP>::CR, pl.sh + HB + H)) ? F(pl.first_node - HB + k) : 0.0;   // CR: reach of the closure rows
    double d[ASTR_HS] = {0}, yh, yt;
    head_rhs<OP, P0>(hw, fc, pl.nsf, pl.sh, d);
    spec_forward(pl.head, d, he, yh, yt);
    publish(0, yh, yt);
  }
  {
    double tw[16];
    for (int k = 0; k < 16; ++k) tw[k] = F(n - 10 + k);
    double d[ASTR_TS], yh, yt;
    tail_rhs<OP, PM>(tw, fc, pl.st - pl.nsl, d);
    spec_forward(pl.tail, d, te, yh, yt);
    publish(E - 1, yh, yt);
  }
  for (int w = 0; w < NW; ++w) {
    const int node0 = pl.first_node + plan_chunk_row(pl, w);
    double wv[WN];
    for (int s = 0; s < WN; ++s) wv[s] = F(node0 - H + s);
    double ee[L], yh, yt;
    // chunk 0 may be SHORT (pl.ls0 < L rows): the kernels run it through the SHORT instantiation
    if (w == 0 && pl.ls0 != L) chunk_forward_short<OP>(pl.sreg, fc, wv, ee, yh, yt);
    else chunk_forward<OP>(pl.reg, fc, wv, ee, yh, yt);
    for (int s = 0; s < L; ++s) e[w][s] = ee[s];
    publish(w + 1, yh, yt);
  }
  // ---- barrier; phase C
  auto GS = [&](int idx) { return S[idx]; };
  auto GP = [&](int idx) { return SP[idx]; };
  {
    const ScanOut so = reduced_scan(pl, GS, GP, 0);
    spec_back(pl.head, he, so.t_prev, so.h_next, [&](int s, double x) { if (s < pl.sh) out[s] = x; });
  }
  {
    const ScanOut so = reduced_scan(pl, GS, GP, E - 1);
    spec_back(pl.tail, te, so.t_prev, so.h_next, [&](int s, double x) { if (s < pl.st) out[pl.nrows - pl.st + s] = x; });
  }
  for (int w = 0; w < NW; ++w) {
    const ScanOut so = reduced_scan(pl, GS, GP, w + 1);
    double ee[L];
    for (int s = 0; s < L; ++s) ee[s] = e[w][s];
    const int r0 = plan_chunk_row(pl, w);
    if (w == 0 && pl.ls0 != L) chunk_back_short(pl.sreg, ee, so.t_prev, so.h_next, [&](int s, double x) { if (s < pl.ls0) out[r0 + s] = x; });
    else chunk_back(pl.reg, ee, so.t_prev, so.h_next, [&](int s, double x) { out[r0 + s] = x; });
  }
  return 0;
}

template <int OP>
int run_op(const LinePlan& pl, const FilterCoef& fc, const double* f5, double* out) {
  const bool p0 = (pl.ntype == 1 || pl.ntype == 4), pm = (pl.ntype == 2 || pl.ntype == 4);
  if (p0 && pm) return run_line<OP, true, true>(pl, fc, f5, out);
  if (p0) return run_line<OP, true, false>(pl, fc, f5, out);
  if (pm) return run_line<OP, false, true>(pl, fc, f5, out);
  return run_line<OP, false, false>(pl, fc, f5, out);
}

}  // namespace

// out[r], r = 0..nrows-1: the solution of every row of the system (row r sits at node first_node + r).
// info: first_node, nrows, NW, sh, st, W.  Returns 1 when the plan does not cover the line.
extern "C" int emul_line(int optype, int ntype, int n, double alfa, int align_even, const double* f5, double* out, int* info) {
  std::vector<double> a, c;
  int first_node, nsf, nsl;
  build_lhs(optype, ntype, n, alfa, a, c, first_node, nsf, nsl);
  LinePlan pl;
  build_line_plan(pl, optype, ntype, n, first_node, nsf, nsl, a, c, ASTR_NWMAX, align_even != 0);
  info[0] = first_node; info[1] = (int)a.size();
  if (!pl.ok) return 1;
  info[2] = pl.NW; info[3] = pl.sh; info[4] = pl.st; info[5] = pl.W; info[6] = pl.ls0;
  FilterCoef fc;
  build_filter_coef(fc, optype == 1 ? alfa : 0.49, 1.11, 0.98, (optype >= 2) ? alfa : 0.0);
  switch (optype) {
    case 0: return run_op<0>(pl, fc, f5, out);
    case 1: return run_op<1>(pl, fc, f5, out);
    case 2: return run_op<2>(pl, fc, f5, out);
    default: return run_op<3>(pl, fc, f5, out);
  }
}
