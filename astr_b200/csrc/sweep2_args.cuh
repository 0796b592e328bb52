// astr_b200/csrc/sweep2_args.cuh -- launch arguments shared by the host glue (sweep2.cu) and the kernels (sweep2_impl.cuh)
#pragma once
#include "common.cuh"

struct FastDiv {           // n / d for 0 <= n < 2^31 without a hardware divide
  unsigned d, m, s1, s2;
};
inline FastDiv make_fastdiv(unsigned d) {
  FastDiv f;
  f.d = d;
  unsigned l = 0;
  while ((1ull << l) < d) ++l;
  f.m = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.s1 = l < 1 ? l : 1;
  f.s2 = l < 1 ? 0 : l - 1;
  return f;
}
__device__ __forceinline__ unsigned fdiv(unsigned n, const FastDiv& f) {
  const unsigned t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.s1)) >> f.s2;
}

struct Sweep2Args {
  Layout L;
  int rb, nbox;              // j/k: line nodes per TMA box, boxes per bundle
  int sp;                    // i: row pitch of the shared-memory line tile (doubles, sp/2 odd)
  int slot[ASTR_MAXF];       // j/k: 4th tensor coordinate of each input field
  const double* in[ASTR_MAXF];
  double* out[ASTR_MAXF];
  int nf, epi, o_lo, o_hi;
  FastDiv dx, dxy;           // bundle index -> (bx, by, bz)
  int nby;
  FastDiv dpair;             // i: copy-out item -> (line, node pair)
  // j/k: PACKED bundles (indices >= nfull).  When the pencils of a row do not fill whole bundles (513 = 16 x 32 + 1)
  // the rag_r left-over pencils of rag_G different rows share one bundle: lane = g * rag_r + x.  Their windows are
  // read straight from global memory (no tile), so they cost their share of the traffic instead of a whole bundle.
  int nfull, rag_r, rag_G, rag_i0;
  FastDiv dpk, drr;          // packed index -> (bundle of the field, field); lane -> g
  // i: lines are numbered j + (jm + 1) k across the planes of a field; a bundle is LINES consecutive lines
  FastDiv dline;             // line -> (j, k)
  int nlines;                // lines of a field
  int nbundles;
};

