// oracle/solver.cpp -- TEST INFRASTRUCTURE ONLY (see astr_oracle.hpp).
// Restatement of the MAIN solver's Runge-Kutta stage (src/mainloop.F90:396-482) on a
// virtual isize x jsize x ksize block grid held in one process.  Each virtual block
// is what one MPI rank (reference) / one GPU (astr_b200) owns; halo exchanges between
// blocks are plain copies that follow src/parallel.F90 statement by statement
// (pack everything, then unpack -- the blocking mpi_sendrecv semantics).
//
// Parity status: UNPINNED (no stored number in the reference covers this mode);
// cross-checked against the pinned mini-app mode at the 1e-13 level in
// tests/test_oracle_solver.py.
#include "astr_oracle.hpp"
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <functional>

namespace astr_oracle {

struct Block {
  int im, jm, km;
  int rk[3];           // irk,jrk,krk
  int g0[3];           // ig0,jg0,kg0
  int npdc[3];         // npdci,npdcj,npdck (src/parallel.F90:1042-1228)
  int s[3], e[3];      // is..ie, js..je, ks..ke
  int nb[3][2];        // neighbour block index per direction (lo,hi), -1 = MPI_PROC_NULL
  CompactScheme fds[3], fil[3];
  Field x[3], q[5], rho, vel[3], prs, tmp, jacob, dxi[3][3];
  Field qrhs[5], qsave[5], dvel[3][3], dtmp[3], sigma[6], qflux[3], vor[3];
  Field rhsav[5];      // rk4 only (src/mainloop.F90:394), allocated by oracle_case_set_rkscheme
  // spongefilter_layer (src/sponge_layer.F90:67-319): per face (i0, im, j0, jm, k0, km) the node range of the
  // layer along the face direction on this block (beg<0: none here) and sponge_damp_coef over the box
  // [beg:end] x [s:e] x [s:e] in Fortran order
  struct SpongeLayer { int beg = -1, end = -2; std::vector<double> coef; };
  SpongeLayer spg[6];
  // spg_def='circl' (src/sponge_layer.F90:369-440): sponge_damp_coef over the whole [s:e]^3 box of the block
  std::vector<double> spg_circ;
  bool spg_circ_loc = false;
  Field ssf, lshock;   // ducrossensor (allocated on first use, commcal.F90:218-219); lshock as 0/1
  Field crinod;        // critical nodes of the crash control as 0/1 (src/commarray.F90:105; mainloop.F90:81 .false.)
  int dim(int d) const { return d == 0 ? im : (d == 1 ? jm : km); }
};

struct Case {
  int ia, ja, ka, size[3];
  int ndims() const { return ka == 0 ? 2 : 3; }   // src/parallel.F90:208-214
  bool homo[3];
  bool lfilter = true, diffterm = true;
  double alfa_filter = 0.49, deltat = 1e-3, time = 0.0;
  int nstep = 0;
  int nthreads = 0;
  int rkscheme = 3;          // 3: 'rk3' (TVD), 4: 'rk4' (src/mainloop.F90:348-388)
  Thermo th;
  FilterCoef fc;
  std::vector<Block> blk;
  double force[3] = {0, 0, 0};
  int flowtype = 0;          // 0 tgv / generic (no source term), 1 channel (src_chan)
  char scheme_kind = 'c';    // difschm(4:4): 'c' compact, 'e' explicit (comsolver.F90:76-84)
  // conschm: 643 central (convrsdcal6) or 543 upwind compact (convrsdcmp); input-file line
  // `recon_schem, lchardecomp, bfacmpld, shkcrt` (src/readwrite.F90)
  int conschm = 643;
  bool conschm_explicit = false;   // conschm(4:4)=='e' with an odd first digit: convrsduwd (solver.F90:220-223)
  int recon_schem = 3;             // reconstruction scheme of recons_exp (flux.F90:269-350)
  bool lchardecomp = false;
  double bfacmpld = 0.3, shkcrt = 0.01;
  int bctype[6] = {1, 1, 1, 1, 1, 1};      // src/readwrite.F90 bctype(1:6): imin,imax,jmin,jmax,kmin,kmax
  double twall[6] = {0, 0, 0, 0, 0, 0};
  double pinf = 0.0;                       // src/solver.F90:120 pinf=roinf*tinf/const2 (set in create)
  // free stream of the far-field faces (commvar uinf, vinf, winf, roinf; src/solver.F90:113-120): nondimensional 1, 0, 0, 1
  double uinf = 1.0, vinf = 0.0, winf = 0.0, roinf = 1.0;
  // inflow(1) data of the blocks with irk==0 (bc.F90 alloinflow :69-83): vel_in(0:jm,0:km,3),
  // tmp_in(0:jm,0:km), tmp_prof(0:jm); indexed [block]
  std::vector<std::vector<double>> vel_in, tmp_in, tmp_prof;
  bool lspg[6] = {false, false, false, false, false, false};   // lspg_i0 .. lspg_km (global flags)
  // databakup (src/mainloop.F90:826-972): two alternating in-memory copies of q with their step counters
  struct DataBack { int nstep = 0; double time = 0.0; int recover_counter = 0; std::vector<std::vector<double>> q; };
  DataBack dat_a, dat_b;
  char datpnt = 'o';
  int crinod_ntimes = 0;
  bool spg_def_circl = false;   // spg_def: 'layer' (default) or 'circl'
  bool lspg_circ = false;    // spg_def='circl': lsponge (global flag, por of lsponge_loc)
  std::vector<double> hist;  // nstep,time,kenergy,enstrophy (statistic.F90:871-990)
  double xmax = 0.0;
};

// ---------------------------------------------------------------------------------
// generic accessors along a direction: l is the index along dir, (p1,p2) the others
// ---------------------------------------------------------------------------------
static inline double& at(Field& a, int d, int l, int p1, int p2) {
  return d == 0 ? a(l, p1, p2) : (d == 1 ? a(p1, l, p2) : a(p1, p2, l));
}
static inline double at(const Field& a, int d, int l, int p1, int p2) {
  return d == 0 ? a(l, p1, p2) : (d == 1 ? a(p1, l, p2) : a(p1, p2, l));
}
// extents of the two "other" indices for direction d
static inline void others(const Block& b, int d, int& n1, int& n2) {
  if (d == 0) { n1 = b.jm; n2 = b.km; }
  else if (d == 1) { n1 = b.im; n2 = b.km; }
  else { n1 = b.im; n2 = b.jm; }
}

template <class Fn>
static void for_each_pencil(const Block& b, int d, Fn fn) {
  int n1, n2;
  others(b, d, n1, n2);
#pragma omp parallel for collapse(2) schedule(static)
  for (int p2 = 0; p2 <= n2; ++p2)
    for (int p1 = 0; p1 <= n1; ++p1) fn(p1, p2);
}

// ---------------------------------------------------------------------------------
// src/parallel.F90:356-484 (parapp) + :919-1248 (parallelini)
// ---------------------------------------------------------------------------------
static void decompose(Case& c) {
  const int n[3] = {c.ia, c.ja, c.ka};
  std::vector<int> mp[3], off[3];
  for (int d = 0; d < 3; ++d) {
    const int sz = c.size[d];
    mp[d].assign(sz, n[d] / sz);
    const int n2 = n[d] % sz;
    for (int r = sz - 1; r >= sz - n2; --r) mp[d][r] += 1;  // :399-405
    off[d].assign(sz, 0);
    for (int r = 1; r < sz; ++r) off[d][r] = off[d][r - 1] + mp[d][r - 1];
  }
  const int isize = c.size[0], jsize = c.size[1], ksize = c.size[2];
  c.blk.resize((size_t)isize * jsize * ksize);
  for (int krk = 0; krk < ksize; ++krk)
    for (int jrk = 0; jrk < jsize; ++jrk)
      for (int irk = 0; irk < isize; ++irk) {
        Block& b = c.blk[(size_t)krk * (isize * jsize) + jrk * isize + irk];
        b.rk[0] = irk; b.rk[1] = jrk; b.rk[2] = krk;
        b.im = mp[0][irk]; b.jm = mp[1][jrk]; b.km = mp[2][krk];
        b.g0[0] = off[0][irk]; b.g0[1] = off[1][jrk]; b.g0[2] = off[2][krk];
        for (int d = 0; d < 3; ++d) {
          const int sz = c.size[d], r = b.rk[d], rm = sz - 1, dm = b.dim(d);
          auto rank_of = [&](int rr) {
            int cc[3] = {irk, jrk, krk};
            cc[d] = rr;
            return cc[2] * (isize * jsize) + cc[1] * isize + cc[0];
          };
          if (c.homo[d]) {  // :1042-1066
            b.s[d] = 0; b.e[d] = dm; b.npdc[d] = 3;
            if (sz == 1) { b.nb[d][0] = b.nb[d][1] = -1; }
            else {
              b.nb[d][0] = rank_of(r == 0 ? rm : r - 1);
              b.nb[d][1] = rank_of(r == rm ? 0 : r + 1);
            }
          } else if (sz == 1) {  // :1071-1077 -- is..ie are NOT assigned by the reference
            b.nb[d][0] = b.nb[d][1] = -1; b.npdc[d] = 4;
            b.s[d] = 1; b.e[d] = dm - 1;  // extension, SURVEY.md Q4
          } else if (r == 0) {
            b.s[d] = 1; b.e[d] = dm; b.npdc[d] = 1;
            b.nb[d][0] = -1; b.nb[d][1] = rank_of(r + 1);
          } else if (r == rm) {
            b.s[d] = 0; b.e[d] = dm - 1; b.npdc[d] = 2;
            b.nb[d][0] = rank_of(r - 1); b.nb[d][1] = -1;
          } else {
            b.s[d] = 0; b.e[d] = dm; b.npdc[d] = 3;
            b.nb[d][0] = rank_of(r - 1); b.nb[d][1] = rank_of(r + 1);
          }
        }
      }
}

static void alloc_block(Block& b) {
  auto A = [&](Field& f) { f.alloc(b.im, b.jm, b.km); };
  for (auto& f : b.x) A(f);
  for (auto& f : b.q) A(f);
  A(b.rho); for (auto& f : b.vel) A(f); A(b.prs); A(b.tmp); A(b.jacob);
  for (auto& r : b.dxi) for (auto& f : r) A(f);
  for (auto& f : b.qrhs) A(f);
  for (auto& f : b.qsave) A(f);
  for (auto& r : b.dvel) for (auto& f : r) A(f);
  for (auto& f : b.dtmp) A(f);
  for (auto& f : b.sigma) A(f);
  for (auto& f : b.qflux) A(f);
  for (auto& f : b.vor) A(f);
}

// ---------------------------------------------------------------------------------
// Halo exchanges.  `get(b)` returns the list of fields of block b to exchange.
// ---------------------------------------------------------------------------------
using FieldList = std::vector<Field*>;
using Getter = std::function<FieldList(Block&)>;

enum class Xmode { SWAP, QSWAP, SYNC };

// One direction of array{3,4,5}d_sendrecv (src/parallel.F90:4132-4370; SWAP),
// qswap (:4848-5318; QSWAP: 0:hm planes + shared-node average) or array3d_sync
// (:3725-3934; SYNC: shared-node average only).
static void exchange_dir(Case& c, int d, const Getter& get, Xmode mode) {
  const size_t nblk = c.blk.size();
  if (c.size[d] == 1) {
    if (!c.homo[d]) return;
    for (Block& b : c.blk) {
      const int dm = b.dim(d);
      int n1, n2;
      others(b, d, n1, n2);
      for (Field* f : get(b)) {
        if (dm == 0) {  // 2-D replicate (:4296-4299)
          if (mode == Xmode::SYNC) continue;
          for (int p2 = 0; p2 <= n2; ++p2)
            for (int p1 = 0; p1 <= n1; ++p1)
              for (int l = -hm; l <= hm; ++l) at(*f, d, l, p1, p2) = at(*f, d, 0, p1, p2);
          continue;
        }
#pragma omp parallel for collapse(2) schedule(static)
        for (int p2 = 0; p2 <= n2; ++p2)
          for (int p1 = 0; p1 <= n1; ++p1) {
            if (mode != Xmode::SYNC) {
              for (int l = -hm; l <= -1; ++l) at(*f, d, l, p1, p2) = at(*f, d, dm + l, p1, p2);
              for (int l = 1; l <= hm; ++l) at(*f, d, dm + l, p1, p2) = at(*f, d, l, p1, p2);
            }
            if (mode != Xmode::SWAP) {
              const double v = 0.5 * (at(*f, d, 0, p1, p2) + at(*f, d, dm, p1, p2));
              at(*f, d, 0, p1, p2) = v;
              at(*f, d, dm, p1, p2) = v;
            }
          }
      }
    }
    return;
  }
  // multi-block: pack all, then unpack all
  struct Buf { std::vector<double> lo, hi; };  // lo: sent to the low neighbour (planes 0|1..hm)
  std::vector<std::vector<Buf>> bufs(nblk);
  const int l0 = (mode == Xmode::SWAP) ? 1 : 0;
  const int l1 = (mode == Xmode::SYNC) ? 0 : hm;
  const int np = l1 - l0 + 1;
  for (size_t ib = 0; ib < nblk; ++ib) {
    Block& b = c.blk[ib];
    const int dm = b.dim(d);
    int n1, n2;
    others(b, d, n1, n2);
    FieldList fl = get(b);
    bufs[ib].resize(fl.size());
    for (size_t n = 0; n < fl.size(); ++n) {
      Buf& B = bufs[ib][n];
      const size_t cnt = (size_t)np * (n1 + 1) * (n2 + 1);
      B.lo.resize(cnt); B.hi.resize(cnt);
      size_t t = 0;
      for (int p2 = 0; p2 <= n2; ++p2)
        for (int p1 = 0; p1 <= n1; ++p1)
          for (int l = l0; l <= l1; ++l, ++t) {
            B.lo[t] = at(*fl[n], d, l, p1, p2);             // sbuf1 = var(l0:hm)
            B.hi[t] = at(*fl[n], d, dm - l, p1, p2);        // sbuf2 (stored mirrored: dm-l)
          }
    }
  }
  for (size_t ib = 0; ib < nblk; ++ib) {
    Block& b = c.blk[ib];
    const int dm = b.dim(d);
    int n1, n2;
    others(b, d, n1, n2);
    FieldList fl = get(b);
    for (size_t n = 0; n < fl.size(); ++n) {
      if (b.nb[d][1] >= 0) {  // from the high neighbour: its `lo` buffer
        const Buf& R = bufs[b.nb[d][1]][n];
        size_t t = 0;
        for (int p2 = 0; p2 <= n2; ++p2)
          for (int p1 = 0; p1 <= n1; ++p1)
            for (int l = l0; l <= l1; ++l, ++t) {
              if (l == 0) at(*fl[n], d, dm, p1, p2) = 0.5 * (at(*fl[n], d, dm, p1, p2) + R.lo[t]);
              else at(*fl[n], d, dm + l, p1, p2) = R.lo[t];
            }
      }
      if (b.nb[d][0] >= 0) {  // from the low neighbour: its `hi` buffer
        const Block& nbk = c.blk[b.nb[d][0]];
        (void)nbk;
        const Buf& R = bufs[b.nb[d][0]][n];
        size_t t = 0;
        for (int p2 = 0; p2 <= n2; ++p2)
          for (int p1 = 0; p1 <= n1; ++p1)
            for (int l = l0; l <= l1; ++l, ++t) {
              if (l == 0) at(*fl[n], d, 0, p1, p2) = 0.5 * (at(*fl[n], d, 0, p1, p2) + R.hi[t]);
              else at(*fl[n], d, -l, p1, p2) = R.hi[t];
            }
      }
    }
  }
}

static void dataswap(Case& c, const Getter& get, int direction = -1) {
  for (int d = 0; d < 3; ++d)
    if (direction < 0 || direction == d) exchange_dir(c, d, get, Xmode::SWAP);
}
static void datasync(Case& c, const Getter& get) {
  for (int d = 0; d < 3; ++d) exchange_dir(c, d, get, Xmode::SYNC);
}

// src/fludyna.F90:545-634 q2fvar_3da (non-COMB, nondimen) at one node
static inline void q2fvar_node(const Thermo& th, Block& b, int i, int j, int k) {
  const double r = b.q[0](i, j, k);
  b.rho(i, j, k) = r;
  const double u = b.q[1](i, j, k) / r, v = b.q[2](i, j, k) / r, w = b.q[3](i, j, k) / r;
  b.vel[0](i, j, k) = u; b.vel[1](i, j, k) = v; b.vel[2](i, j, k) = w;
  const double p = (b.q[4](i, j, k) - 0.5 * r * (u * u + v * v + w * w)) / th.const6;
  b.prs(i, j, k) = p;
  b.tmp(i, j, k) = th.thermal_T(p, r);  // thermal_3d(pressure,density) fludyna.F90:136-179
}

// src/parallel.F90:4848-5318 qswap: per direction (i then j then k) halo copy,
// shared-node average, then q2fvar on the slabs -hm:0 and dim:dim+hm of sides that
// have a neighbour (or both sides for a single periodic block).
static void qswap(Case& c) {
  Getter gq = [](Block& b) { return FieldList{&b.q[0], &b.q[1], &b.q[2], &b.q[3], &b.q[4]}; };
  for (int d = 0; d < 3; ++d) {
    exchange_dir(c, d, gq, Xmode::QSWAP);
    for (Block& b : c.blk) {
      const int dm = b.dim(d);
      int n1, n2;
      others(b, d, n1, n2);
      bool lo, hi;
      if (c.size[d] == 1) { lo = hi = c.homo[d]; }
      else { lo = b.nb[d][0] >= 0; hi = b.nb[d][1] >= 0; }
      auto slab = [&](int la, int lb) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int p2 = 0; p2 <= n2; ++p2)
          for (int p1 = 0; p1 <= n1; ++p1)
            for (int l = la; l <= lb; ++l) {
              if (d == 0) q2fvar_node(c.th, b, l, p1, p2);
              else if (d == 1) q2fvar_node(c.th, b, p1, l, p2);
              else q2fvar_node(c.th, b, p1, p2, l);
            }
      };
      if (hi) slab(dm, dm + hm);
      if (lo) slab(-hm, 0);
    }
  }
}

// ---------------------------------------------------------------------------------
// Grid + metrics: src/gridgeneration.F90:233-265 (gridcube), src/parallel.F90:2780-3035
// (gridsendrecv), src/geom.F90:99-700 (gridgeom, ndims==3 branch, no lfftk).
// ---------------------------------------------------------------------------------
static void gridsendrecv(Case& c) {
  // Pack/unpack relative offsets; physical boundaries get an even reflection.
  for (int d = 0; d < 3; ++d) {
    const size_t nblk = c.blk.size();
    if (c.size[d] == 1) {
      for (Block& b : c.blk) {
        const int dm = b.dim(d);
        int n1, n2;
        others(b, d, n1, n2);
        if (dm == 0) {  // ka==0 (:3015-3021): planes -hm..hm copy x,y of plane 0, z = z(0)+k
          for (int p2 = 0; p2 <= n2; ++p2)
            for (int p1 = 0; p1 <= n1; ++p1)
              for (int l = -hm; l <= hm; ++l) {
                if (l == 0) continue;
                at(b.x[0], d, l, p1, p2) = at(b.x[0], d, 0, p1, p2);
                at(b.x[1], d, l, p1, p2) = at(b.x[1], d, 0, p1, p2);
                at(b.x[2], d, l, p1, p2) = at(b.x[2], d, 0, p1, p2) + double(l);
              }
          continue;
        }
        for (int m = 0; m < 3; ++m)
          for (int p2 = 0; p2 <= n2; ++p2)
            for (int p1 = 0; p1 <= n1; ++p1)
              for (int n = 1; n <= hm; ++n) {  // :2861-2866
                Field& x = b.x[m];
                at(x, d, dm + n, p1, p2) = at(x, d, dm, p1, p2) + (at(x, d, n, p1, p2) - at(x, d, 0, p1, p2));
                at(x, d, -n, p1, p2) = at(x, d, 0, p1, p2) - (at(x, d, dm, p1, p2) - at(x, d, dm - n, p1, p2));
              }
      }
      continue;
    }
    std::vector<std::vector<double>> lo(nblk), hi(nblk);
    for (size_t ib = 0; ib < nblk; ++ib) {
      Block& b = c.blk[ib];
      const int dm = b.dim(d);
      int n1, n2;
      others(b, d, n1, n2);
      for (int m = 0; m < 3; ++m)
        for (int p2 = 0; p2 <= n2; ++p2)
          for (int p1 = 0; p1 <= n1; ++p1)
            for (int n = 1; n <= hm; ++n) {  // :2812-2815
              lo[ib].push_back(at(b.x[m], d, n, p1, p2) - at(b.x[m], d, 0, p1, p2));
              hi[ib].push_back(at(b.x[m], d, dm - n, p1, p2) - at(b.x[m], d, dm, p1, p2));
            }
    }
    for (size_t ib = 0; ib < nblk; ++ib) {
      Block& b = c.blk[ib];
      const int dm = b.dim(d);
      int n1, n2;
      others(b, d, n1, n2);
      size_t t = 0;
      for (int m = 0; m < 3; ++m)
        for (int p2 = 0; p2 <= n2; ++p2)
          for (int p1 = 0; p1 <= n1; ++p1)
            for (int n = 1; n <= hm; ++n, ++t) {
              Field& x = b.x[m];
              if (b.nb[d][0] < 0) at(x, d, -n, p1, p2) = 2.0 * at(x, d, 0, p1, p2) - at(x, d, n, p1, p2);
              else at(x, d, -n, p1, p2) = hi[b.nb[d][0]][t] + at(x, d, 0, p1, p2);
              if (b.nb[d][1] < 0) at(x, d, dm + n, p1, p2) = 2.0 * at(x, d, dm, p1, p2) - at(x, d, dm - n, p1, p2);
              else at(x, d, dm + n, p1, p2) = lo[b.nb[d][1]][t] + at(x, d, dm, p1, p2);
            }
    }
  }
}

// derivative of one pencil of `src` along d at (p1,p2) -> out[0..dim]
struct PencilWork {
  std::vector<double> buf, df, work;
  explicit PencilWork(int md) : buf(md + 1 + 2 * hm), df(md + 1), work(2 * (md + 8)) {}
  double* f() { return buf.data() + hm; }
};
static inline void gather(const Field& a, int d, int p1, int p2, int dim, double* f) {
  for (int l = -hm; l <= dim + hm; ++l) f[l] = at(a, d, l, p1, p2);
}

// fds%central (src/derivative.F90:16-34): compact_central -> df_compact (:171-198),
// explicit_central -> df_explicit -> diff6ec (:319-413)
static inline void central(const Case& c, const CompactScheme& s, const double* f, double* df, double* work) {
  if (c.scheme_kind == 'e') diff6ec(f, s.dimension, s.nbctype, df);
  else df_compact(s, f, df, work);
}

static void gridgeom(Case& c) {
  gridsendrecv(c);
  const int md = std::max(c.ia, std::max(c.ja, c.ka));
  // dx(m,n) = d x_m / d xi_n   (geom.F90:130-164)
  std::vector<std::vector<Field>> DX(c.blk.size());
  for (size_t ib = 0; ib < c.blk.size(); ++ib) {
    Block& b = c.blk[ib];
    DX[ib].resize(9);
    for (auto& f : DX[ib]) f.alloc(b.im, b.jm, b.km);
    for (int d = 0; d < c.ndims(); ++d)
      for_each_pencil(b, d, [&](int p1, int p2) {
        PencilWork w(md);
        for (int m = 0; m < 3; ++m) {
          gather(b.x[m], d, p1, p2, b.dim(d), w.f());
          central(c, b.fds[d], w.f(), w.df.data(), w.work.data());
          for (int l = 0; l <= b.dim(d); ++l) at(DX[ib][m * 3 + d], d, l, p1, p2) = w.df[l];
        }
      });
  }
  size_t cur = 0;
  std::vector<std::vector<Field>>* cur_list = &DX;
  Getter gdx = [&](Block& b) {
    size_t ib = &b - c.blk.data();
    FieldList fl;
    for (auto& f : (*cur_list)[ib]) fl.push_back(&f);
    return fl;
  };
  (void)cur;
  dataswap(c, gdx);  // geom.F90:333
  // jacobian :340-357
  for (size_t ib = 0; ib < c.blk.size(); ++ib) {
    Block& b = c.blk[ib];
    auto dx = [&](int m, int n) -> Field& { return DX[ib][(m - 1) * 3 + (n - 1)]; };
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i)
          if (c.ndims() == 2)   // :371-375
            b.jacob(i, j, k) = dx(1, 1)(i, j, k) * dx(2, 2)(i, j, k) - dx(1, 2)(i, j, k) * dx(2, 1)(i, j, k);
          else
          b.jacob(i, j, k) = dx(1, 1)(i, j, k) * dx(2, 2)(i, j, k) * dx(3, 3)(i, j, k) +
                             dx(1, 2)(i, j, k) * dx(2, 3)(i, j, k) * dx(3, 1)(i, j, k) +
                             dx(1, 3)(i, j, k) * dx(2, 1)(i, j, k) * dx(3, 2)(i, j, k) -
                             dx(1, 3)(i, j, k) * dx(2, 2)(i, j, k) * dx(3, 1)(i, j, k) -
                             dx(1, 2)(i, j, k) * dx(2, 1)(i, j, k) * dx(3, 3)(i, j, k) -
                             dx(1, 1)(i, j, k) * dx(2, 3)(i, j, k) * dx(3, 2)(i, j, k);
  }
  Getter gj = [](Block& b) { return FieldList{&b.jacob}; };
  dataswap(c, gj);   // :383
  datasync(c, gj);   // :385
  // conservative-form d xi/d x :399-519.  dxi(a,b): a = xi index, b = x index.
  for (size_t ib = 0; ib < c.blk.size(); ++ib) {
    Block& b = c.blk[ib];
    auto dx = [&](int m, int n) -> const Field& { return DX[ib][(m - 1) * 3 + (n - 1)]; };
    // term list: {sweep dir, target dxi(a,b), dx(m1,n1)*x(c1) - dx(m2,n2)*x(c2)}
    struct Term { int d, a, bb, m1, n1, c1, m2, n2, c2; };
    static const Term terms[] = {
        // i-sweeps (geom.F90:402-427)
        {0, 2, 1, 2, 3, 3, 3, 3, 2}, {0, 2, 2, 3, 3, 1, 1, 3, 3}, {0, 2, 3, 1, 3, 2, 2, 3, 1},
        {0, 3, 1, 3, 2, 2, 2, 2, 3}, {0, 3, 2, 1, 2, 3, 3, 2, 1}, {0, 3, 3, 2, 2, 1, 1, 2, 2},
        // j-sweeps (:433-457)
        {1, 1, 1, 3, 3, 2, 2, 3, 3}, {1, 1, 2, 1, 3, 3, 3, 3, 1}, {1, 1, 3, 2, 3, 1, 1, 3, 2},
        {1, 3, 1, 2, 1, 3, 3, 1, 2}, {1, 3, 2, 3, 1, 1, 1, 1, 3}, {1, 3, 3, 1, 1, 2, 2, 1, 1},
        // k-sweeps (:464-510)
        {2, 1, 1, 2, 2, 3, 3, 2, 2}, {2, 1, 2, 3, 2, 1, 1, 2, 3}, {2, 1, 3, 1, 2, 2, 2, 2, 1},
        {2, 2, 1, 3, 1, 2, 2, 1, 3}, {2, 2, 2, 1, 1, 3, 3, 1, 1}, {2, 2, 3, 2, 1, 1, 1, 1, 2}};
    for (auto& r : b.dxi) for (auto& f : r) std::fill(f.v.begin(), f.v.end(), 0.0);
    if (c.ndims() == 2) {   // :520-527
      for (int k = 0; k <= b.km; ++k)
        for (int j = 0; j <= b.jm; ++j)
          for (int i = 0; i <= b.im; ++i) {
            b.dxi[0][0](i, j, k) = dx(2, 2)(i, j, k);
            b.dxi[0][1](i, j, k) = -dx(1, 2)(i, j, k);
            b.dxi[1][0](i, j, k) = -dx(2, 1)(i, j, k);
            b.dxi[1][1](i, j, k) = dx(1, 1)(i, j, k);
          }
      continue;
    }
    for (const Term& t : terms) {
      for_each_pencil(b, t.d, [&](int p1, int p2) {
        PencilWork w(md);
        double* phi = w.f();
        const int dm = b.dim(t.d);
        for (int l = -hm; l <= dm + hm; ++l)
          phi[l] = 0.5 * (at(dx(t.m1, t.n1), t.d, l, p1, p2) * at(b.x[t.c1 - 1], t.d, l, p1, p2) -
                          at(dx(t.m2, t.n2), t.d, l, p1, p2) * at(b.x[t.c2 - 1], t.d, l, p1, p2));
        central(c, b.fds[t.d], phi, w.df.data(), w.work.data());
        Field& tgt = b.dxi[t.a - 1][t.bb - 1];
        for (int l = 0; l <= dm; ++l) at(tgt, t.d, l, p1, p2) = at(tgt, t.d, l, p1, p2) + w.df[l];
      });
    }
  }
  Getter gdxi = [](Block& b) {
    FieldList fl;
    for (auto& r : b.dxi) for (auto& f : r) fl.push_back(&f);
    return fl;
  };
  dataswap(c, gdxi);  // :541  (the metric-identity check :546-646 has no side effects)
  for (Block& b : c.blk)  // :662-666
    for (int n = 0; n < 3; ++n)
      for (int m = 0; m < 3; ++m)
        for (int k = 0; k <= b.km; ++k)
          for (int j = 0; j <= b.jm; ++j)
            for (int i = 0; i <= b.im; ++i) b.dxi[m][n](i, j, k) = b.dxi[m][n](i, j, k) / b.jacob(i, j, k);
  dataswap(c, gdxi);  // :674
  datasync(c, gdxi);  // :676-680
  // geombc (:682) only acts for wall bctype 41: not restated (TGV/periodic scope).
}

// ---------------------------------------------------------------------------------
// Stage operators
// ---------------------------------------------------------------------------------
// src/comsolver.F90:514-632 filterq
static void filterq(Case& c) {
  const int md = std::max(c.ia, std::max(c.ja, c.ka));
  Getter gq = [](Block& b) { return FieldList{&b.q[0], &b.q[1], &b.q[2], &b.q[3], &b.q[4]}; };
  for (int d = 0; d < c.ndims(); ++d) {   // comsolver.F90:588 `if(ndims==3)` for k
    dataswap(c, gq, d);
    for (Block& b : c.blk) {
      const int dm = b.dim(d);
      for_each_pencil(b, d, [&](int p1, int p2) {
        PencilWork w(md);
        for (int n = 0; n < 5; ++n) {
          gather(b.q[n], d, p1, p2, dm, w.f());
          compact_filter(b.fil[d], c.fc, w.f(), w.df.data(), w.work.data());
          for (int l = 0; l <= dm; ++l) at(b.q[n], d, l, p1, p2) = w.df[l];
        }
      });
    }
  }
}

// src/comsolver.F90:244-497 gradcal
static void gradcal(Case& c) {
  const int md = std::max(c.ia, std::max(c.ja, c.ka));
  for (Block& b : c.blk) {
    for (auto& r : b.dvel) for (auto& f : r) std::fill(f.v.begin(), f.v.end(), 0.0);
    for (auto& f : b.dtmp) std::fill(f.v.begin(), f.v.end(), 0.0);
    for (int d = 0; d < c.ndims(); ++d) {   // comsolver.F90:418 `if(ndims==3)` for k
      const int dm = b.dim(d);
      for_each_pencil(b, d, [&](int p1, int p2) {
        PencilWork w(md);
        for (int n = 0; n < 4; ++n) {
          const Field& src = n < 3 ? b.vel[n] : b.tmp;
          gather(src, d, p1, p2, dm, w.f());
          central(c, b.fds[d], w.f(), w.df.data(), w.work.data());
          for (int m = 0; m < 3; ++m) {
            Field& dst = n < 3 ? b.dvel[n][m] : b.dtmp[m];
            for (int l = 0; l <= dm; ++l)
              at(dst, d, l, p1, p2) = at(dst, d, l, p1, p2) + w.df[l] * at(b.dxi[d][m], d, l, p1, p2);
          }
        }
      });
    }
  }
}

// src/solver.F90:2173-2341 convrsdcal6
static void convrsdcal6(Case& c) {
  const int md = std::max(c.ia, std::max(c.ja, c.ka));
  for (Block& b : c.blk) {
    for (int d = 0; d < c.ndims(); ++d) {   // solver.F90:2284 `if(ndims==3)` for k
      const int dm = b.dim(d);
      const int np = md + 1 + 2 * hm;
      int o1, o2;  // the two other directions
      if (d == 0) { o1 = 1; o2 = 2; } else if (d == 1) { o1 = 0; o2 = 2; } else { o1 = 0; o2 = 1; }
      for_each_pencil(b, d, [&](int p1, int p2) {
        // loop bounds are js:je / ks:ke etc. for the non-swept indices (:2197-2198)
        if (p1 < b.s[o1] || p1 > b.e[o1] || p2 < b.s[o2] || p2 > b.e[o2]) return;
        std::vector<double> fb(5 * np), df(md + 1), work(2 * (md + 8));
        double* fcs[5];
        for (int n = 0; n < 5; ++n) fcs[n] = fb.data() + n * np + hm;
        for (int l = -hm; l <= dm + hm; ++l) {
          const double d1 = at(b.dxi[d][0], d, l, p1, p2), d2 = at(b.dxi[d][1], d, l, p1, p2),
                       d3 = at(b.dxi[d][2], d, l, p1, p2);
          const double uu = d1 * at(b.vel[0], d, l, p1, p2) + d2 * at(b.vel[1], d, l, p1, p2) +
                            d3 * at(b.vel[2], d, l, p1, p2);
          const double jac = at(b.jacob, d, l, p1, p2), p = at(b.prs, d, l, p1, p2);
          fcs[0][l] = jac * at(b.q[0], d, l, p1, p2) * uu;
          fcs[1][l] = jac * (at(b.q[1], d, l, p1, p2) * uu + d1 * p);
          fcs[2][l] = jac * (at(b.q[2], d, l, p1, p2) * uu + d2 * p);
          fcs[3][l] = jac * (at(b.q[3], d, l, p1, p2) * uu + d3 * p);
          fcs[4][l] = jac * (at(b.q[4], d, l, p1, p2) + p) * uu;
        }
        for (int n = 0; n < 5; ++n) {
          central(c, b.fds[d], fcs[n], df.data(), work.data());
          for (int l = b.s[d]; l <= b.e[d]; ++l) {
            double& r = at(b.qrhs[n], d, l, p1, p2);
            r = r + df[l];
          }
        }
      });
    }
  }
}

// src/solver.F90:2354-2873 diffrsdcal6 (turbmode none, nondimen, no species)
static void diffrsdcal6(Case& c) {
  const int md = std::max(c.ia, std::max(c.ja, c.ka));
  const Thermo& th = c.th;
  for (Block& b : c.blk) {
    for (auto& f : b.sigma) std::fill(f.v.begin(), f.v.end(), 0.0);  // :2399-2400
    for (auto& f : b.qflux) std::fill(f.v.begin(), f.v.end(), 0.0);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) {
          const double miu = th.miu_eff(b.tmp(i, j, k));   // solver.F90:2456-2460
          auto dv = [&](int m, int n) { return b.dvel[m - 1][n - 1](i, j, k); };
          const double s11 = dv(1, 1);
          const double s12 = 0.5 * (dv(1, 2) + dv(2, 1));
          const double s13 = 0.5 * (dv(1, 3) + dv(3, 1));
          const double s22 = dv(2, 2);
          const double s23 = 0.5 * (dv(2, 3) + dv(3, 2));
          const double s33 = dv(3, 3);
          const double skk = num1d3 * (s11 + s22 + s33);
          b.vor[0](i, j, k) = dv(3, 2) - dv(2, 3);
          b.vor[1](i, j, k) = dv(1, 3) - dv(3, 1);
          b.vor[2](i, j, k) = dv(2, 1) - dv(1, 2);
          const double miu2 = 2.0 * miu;
          const double hcc = th.hcc(miu);                  // solver.F90:2519-2523
          const double detk = 0.0, tau = 0.0;
          const double sg1 = miu2 * (s11 - skk) - detk + tau;
          const double sg2 = miu2 * s12 + tau;
          const double sg3 = miu2 * s13 + tau;
          const double sg4 = miu2 * (s22 - skk) - detk + tau;
          const double sg5 = miu2 * s23 + tau;
          const double sg6 = miu2 * (s33 - skk) - detk + tau;
          b.sigma[0](i, j, k) = sg1; b.sigma[1](i, j, k) = sg2; b.sigma[2](i, j, k) = sg3;
          b.sigma[3](i, j, k) = sg4; b.sigma[4](i, j, k) = sg5; b.sigma[5](i, j, k) = sg6;
          const double u = b.vel[0](i, j, k), v = b.vel[1](i, j, k), w = b.vel[2](i, j, k);
          b.qflux[0](i, j, k) = hcc * b.dtmp[0](i, j, k) + sg1 * u + sg2 * v + sg3 * w;
          b.qflux[1](i, j, k) = hcc * b.dtmp[1](i, j, k) + sg2 * u + sg4 * v + sg5 * w;
          b.qflux[2](i, j, k) = hcc * b.dtmp[2](i, j, k) + sg3 * u + sg5 * v + sg6 * w;
        }
  }
  Getter gs = [](Block& b) {
    return FieldList{&b.sigma[0], &b.sigma[1], &b.sigma[2], &b.sigma[3], &b.sigma[4], &b.sigma[5]};
  };
  Getter gqf = [](Block& b) { return FieldList{&b.qflux[0], &b.qflux[1], &b.qflux[2]}; };
  dataswap(c, gs);   // :2604
  dataswap(c, gqf);  // :2606
  static const int sel[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};  // rows of the symmetric sigma
  for (Block& b : c.blk) {
    for (int d = 0; d < c.ndims(); ++d) {   // solver.F90:2776 `if(ndims==3)` for k
      const int dm = b.dim(d);
      const int np = md + 1 + 2 * hm;
      for_each_pencil(b, d, [&](int p1, int p2) {  // all pencils 0:jm,0:km (:2623-2624)
        std::vector<double> fb(4 * np), df(md + 1), work(2 * (md + 8));
        double* ff[4];
        for (int n = 0; n < 4; ++n) ff[n] = fb.data() + n * np + hm;
        for (int l = -hm; l <= dm + hm; ++l) {
          const double d1 = at(b.dxi[d][0], d, l, p1, p2), d2 = at(b.dxi[d][1], d, l, p1, p2),
                       d3 = at(b.dxi[d][2], d, l, p1, p2), jac = at(b.jacob, d, l, p1, p2);
          for (int n = 0; n < 3; ++n)
            ff[n][l] = (at(b.sigma[sel[n][0]], d, l, p1, p2) * d1 + at(b.sigma[sel[n][1]], d, l, p1, p2) * d2 +
                        at(b.sigma[sel[n][2]], d, l, p1, p2) * d3) * jac;
          ff[3][l] = (at(b.qflux[0], d, l, p1, p2) * d1 + at(b.qflux[1], d, l, p1, p2) * d2 +
                      at(b.qflux[2], d, l, p1, p2) * d3) * jac;
        }
        for (int n = 0; n < 4; ++n) {
          central(c, b.fds[d], ff[n], df.data(), work.data());
          for (int l = b.s[d]; l <= b.e[d]; ++l) {
            double& r = at(b.qrhs[n + 1], d, l, p1, p2);
            r = r + df[l];
          }
        }
      });
    }
  }
}

#include "upwind.hpp"
#include "recons.hpp"

// src/solver.F90:295-353 src_chan: bulk velocities by trapezoidal integration in y over
// nodes 1..im,1..jm,1..km of every rank (psum = sum over blocks), then the body force
static void src_chan(Case& c) {
  double robulk = 0.0, u1bulk = 0.0, u2bulk = 0.0, u3bulk = 0.0;
  for (Block& b : c.blk) {
    double ro = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
    const int k1 = c.ndims() == 2 ? 0 : 1, k2 = c.ndims() == 2 ? 0 : b.km;   // :306-315
    for (int k = k1; k <= k2; ++k)
      for (int j = 1; j <= b.jm; ++j)
        for (int i = 1; i <= b.im; ++i) {
          const double dy = b.x[1](i, j, k) - b.x[1](i, j - 1, k);
          ro = ro + 0.5 * (b.q[0](i, j - 1, k) + b.q[0](i, j, k)) * dy;
          u1 = u1 + 0.5 * (b.q[1](i, j - 1, k) + b.q[1](i, j, k)) * dy;
          u2 = u2 + 0.5 * (b.q[2](i, j - 1, k) + b.q[2](i, j, k)) * dy;
          u3 = u3 + 0.5 * (b.q[3](i, j - 1, k) + b.q[3](i, j, k)) * dy;
        }
    robulk += ro; u1bulk += u1; u2bulk += u2; u3bulk += u3;  // psum
  }
  u1bulk = u1bulk / robulk; u2bulk = u2bulk / robulk; u3bulk = u3bulk / robulk;
  const double fe = c.force[0] * u1bulk + c.force[1] * u2bulk + c.force[2] * u3bulk;
  for (Block& b : c.blk)
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) {
          const double jac = b.jacob(i, j, k);
          b.qrhs[1](i, j, k) = b.qrhs[1](i, j, k) + c.force[0] * jac;
          b.qrhs[2](i, j, k) = b.qrhs[2](i, j, k) + c.force[1] * jac;
          b.qrhs[3](i, j, k) = b.qrhs[3](i, j, k) + c.force[2] * jac;
          b.qrhs[4](i, j, k) = b.qrhs[4](i, j, k) + fe * jac;
        }
}

// src/solver.F90:185-282 rhscal (conschm even -> central; flowtype channel adds src_chan :262)
static void rhscal(Case& c) {
  if ((c.conschm / 100) % 2 == 0) convrsdcal6(c);      // :216-218
  else if (c.conschm_explicit) {                       // conschm(4:4)=='e', :220-223
    if (c.recon_schem == 5 || c.lchardecomp) ducrossensor(c);
    convrsduwd(c);
  } else {                                             // conschm(4:4)=='c', :224-226
    if (c.lchardecomp) ducrossensor(c);
    convrsdcmp(c);
  }
  for (Block& b : c.blk)
    for (auto& f : b.qrhs)
      for (double& v : f.v) v = -v;  // :242
  if (c.diffterm) diffrsdcal6(c);
  if (c.flowtype == 1) src_chan(c);
}

// src/bc.F90:6306-6723 noslip(ndir,tw) (nondimen, no species, turbmode none, no wall blowing):
// face node: u=0, T=tw, p extrapolated from the two interior neighbours, rho=thermal(p,T),
// q=fvar2q(rho,vel,p) (fludyna.F90:312-376, pressure branch).  ndir: 1 imin .. 6 kmax.
static void noslip(Case& c, int ndir, double tw) {
  const int d = (ndir - 1) / 2, side = (ndir - 1) % 2;
  for (Block& b : c.blk) {
    if (side == 0 ? b.rk[d] != 0 : b.rk[d] != c.size[d] - 1) continue;
    const int dm = b.dim(d), l = side ? dm : 0, sg = side ? -1 : 1;
    int n1, n2;
    others(b, d, n1, n2);
    for (int p2 = 0; p2 <= n2; ++p2)
      for (int p1 = 0; p1 <= n1; ++p1) {
        const double pe = num1d3 * (4.0 * at(b.prs, d, l + sg, p1, p2) - at(b.prs, d, l + 2 * sg, p1, p2));
        for (int m = 0; m < 3; ++m) at(b.vel[m], d, l, p1, p2) = 0.0;
        at(b.prs, d, l, p1, p2) = pe;
        at(b.tmp, d, l, p1, p2) = tw;
        const double rho = c.th.thermal_rho(pe, tw);  // thermal(pressure,temperature) fludyna.F90:45-88
        at(b.rho, d, l, p1, p2) = rho;
        at(b.q[0], d, l, p1, p2) = rho;
        at(b.q[1], d, l, p1, p2) = rho * 0.0;
        at(b.q[2], d, l, p1, p2) = rho * 0.0;
        at(b.q[3], d, l, p1, p2) = rho * 0.0;
        const double var1 = 0.5 * (0.0 * 0.0 + 0.0 * 0.0 + 0.0 * 0.0);
        // fvar2q with pressure (nondimen) or with temperature (dimensional), bc.F90:6331-6348
        at(b.q[4], d, l, p1, p2) = c.th.nondimen ? pe * c.th.const6 + rho * var1 : rho * (tw * c.th.cotem() + var1);
      }
  }
}

// src/commfunc.F90:277-285 extrapolate_2o with dv=0
static inline double extrapolate2(double v1, double v2) { return num1d3 * (4.0 * v1 - v2 - 2.0 * 0.0); }

// fvar2q_sca (src/fludyna.F90:312-376), nondimen: energy from temperature or from pressure
static inline void fvar2q_T(const Thermo& th, Block& b, int i, int j, int k) {
  const double r = b.rho(i, j, k), u = b.vel[0](i, j, k), v = b.vel[1](i, j, k), w = b.vel[2](i, j, k);
  b.q[0](i, j, k) = r; b.q[1](i, j, k) = r * u; b.q[2](i, j, k) = r * v; b.q[3](i, j, k) = r * w;
  const double var1 = 0.5 * (u * u + v * v + w * w);
  b.q[4](i, j, k) = r * (b.tmp(i, j, k) * th.cotem() + var1);
}
static inline void fvar2q_P(const Thermo& th, Block& b, int i, int j, int k) {
  const double r = b.rho(i, j, k), u = b.vel[0](i, j, k), v = b.vel[1](i, j, k), w = b.vel[2](i, j, k);
  b.q[0](i, j, k) = r; b.q[1](i, j, k) = r * u; b.q[2](i, j, k) = r * v; b.q[3](i, j, k) = r * w;
  const double var1 = 0.5 * (u * u + v * v + w * w);
  b.q[4](i, j, k) = b.prs(i, j, k) * th.const6 + r * var1;
}

// src/bc.F90:1366-1562 inflow(ndir): only ndir==1 exists in the reference (face i=0 of irk==0).
// vel_in / tmp_in / tmp_prof are what profileinflow / freestreaminflow (+ inflowintp at rkstep 1) leave.
static int inflow(Case& c, int ndir) {
  if (ndir != 1) return 0;
  for (size_t ib = 0; ib < c.blk.size(); ++ib) {
    Block& b = c.blk[ib];
    if (b.rk[0] != 0) continue;
    if (c.vel_in.size() <= ib || c.vel_in[ib].empty()) return -1;
    const int i = 0, nj = b.jm + 1, nk = b.km + 1;
    const std::vector<double>& vin = c.vel_in[ib];
    auto VIN = [&](int j, int k, int m) { return vin[(size_t)j + (size_t)nj * ((size_t)k + (size_t)nk * m)]; };
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j) {
        const double rho_ref = b.rho(i + 1, j, k);
        const double css = c.th.sos(c.tmp_prof[ib][j]);                    // sos(tmp_prof(j)) fludyna.F90:832-859
        b.vel[1](i, j, k) = VIN(j, k, 1);
        b.vel[2](i, j, k) = VIN(j, k, 2);
        b.tmp(i, j, k) = c.tmp_in[ib][(size_t)j + (size_t)nj * k];
        const double pe = extrapolate2(b.prs(i + 1, j, k), b.prs(i + 2, j, k));
        const double ue = extrapolate2(b.vel[0](i + 1, j, k), b.vel[0](i + 2, j, k));
        const double pwave_in = c.pinf;
        const double malo = VIN(j, k, 0) / css;
        const double blend = 0.5 * (std::tanh((malo - 1.0) * 6.0) + 1.0);
        b.prs(i, j, k) = (0.5 * (pwave_in + pe) + 0.5 * rho_ref * css * (VIN(j, k, 0) - ue)) * (1.0 - blend) + pwave_in * blend;
        b.vel[0](i, j, k) = VIN(j, k, 0) + (c.pinf - b.prs(i, j, k)) / rho_ref / css;
        b.rho(i, j, k) = c.th.thermal_rho(b.prs(i, j, k), b.tmp(i, j, k)); // thermal(pressure,temperature)
        fvar2q_T(c.th, b, i, j, k);
      }
  }
  return 0;
}

// src/bc.F90:3404-3617 outflow(ndir): the reference has ndir==2 (first-order copy) and ndir==4
// (extrapolation, subsonic branch relaxes the pressure towards pinf)
static int outflow(Case& c, int ndir) {
  for (Block& b : c.blk) {
    if (ndir == 2 && b.rk[0] == c.size[0] - 1) {
      const int i = b.im;
      for (int k = 0; k <= b.km; ++k)
        for (int j = 0; j <= b.jm; ++j) {
          for (int m = 0; m < 3; ++m) b.vel[m](i, j, k) = b.vel[m](i - 1, j, k);
          b.prs(i, j, k) = b.prs(i - 1, j, k);
          b.tmp(i, j, k) = b.tmp(i - 1, j, k);
          b.rho(i, j, k) = c.th.thermal_rho(b.prs(i, j, k), b.tmp(i, j, k));
          fvar2q_T(c.th, b, i, j, k);
        }
    } else if (ndir == 4 && b.rk[1] == c.size[1] - 1) {
      const int j = b.jm;
      const double alpha = 0.25;
      for (int k = 0; k <= b.km; ++k)
        for (int i = 0; i <= b.im; ++i) {
          const double css = c.th.sos(b.tmp(i, j, k));
          const double ub = b.vel[1](i, j, k);
          const double ue = extrapolate2(b.vel[0](i, j - 1, k), b.vel[0](i, j - 2, k));
          const double ve = extrapolate2(b.vel[1](i, j - 1, k), b.vel[1](i, j - 2, k));
          const double we = extrapolate2(b.vel[2](i, j - 1, k), b.vel[2](i, j - 2, k));
          const double pe = extrapolate2(b.prs(i, j - 1, k), b.prs(i, j - 2, k));
          const double te = extrapolate2(b.tmp(i, j - 1, k), b.tmp(i, j - 2, k));
          const double roe = extrapolate2(b.rho(i, j - 1, k), b.rho(i, j - 2, k));
          if (ub >= css) {
            b.prs(i, j, k) = pe;
            b.rho(i, j, k) = roe;
          } else {
            const double pwave_in = (b.prs(i, j, k) + alpha * c.deltat * c.pinf + b.rho(i, j, k) * css * (ve - b.vel[1](i, j, k))) /
                                    (1.0 + alpha * c.deltat);
            b.prs(i, j, k) = pwave_in;
            b.tmp(i, j, k) = te;
            b.rho(i, j, k) = c.th.thermal_rho(b.prs(i, j, k), b.tmp(i, j, k));
          }
          b.vel[0](i, j, k) = ue; b.vel[1](i, j, k) = ve; b.vel[2](i, j, k) = we;
          if (c.th.nondimen) {        // :3596-3600
            b.tmp(i, j, k) = c.th.thermal_T(b.prs(i, j, k), b.rho(i, j, k));
            fvar2q_P(c.th, b, i, j, k);
          } else {                    // :3601-3605
            b.rho(i, j, k) = c.th.thermal_rho(b.prs(i, j, k), b.tmp(i, j, k));
            fvar2q_T(c.th, b, i, j, k);
          }
        }
    } else if (ndir != 2 && ndir != 4) {
      return -1;
    }
  }
  return 0;
}

// src/bc.F90:3008-3392 farfield(ndir).  ndir==4 (jmax, the face the HBL / SWLBI inputs use): plain second-order
// extrapolation of the primitives (:3106-3232).  ndir==3 (jmin, :3024-3104), 5 (kmin, :3235-3312), 6 (kmax,
// :3314-3390): subsonic characteristic inflow / outflow against the free stream (uinf, vinf, winf, roinf, pinf).
static int farfield(Case& c, int ndir) {
  if (ndir < 3 || ndir > 6) return -1;
  const int d = (ndir - 1) / 2, side = (ndir - 1) % 2;      // direction (1: j, 2: k), 0 = low face
  const double vinf3[3] = {c.uinf, c.vinf, c.winf};
  for (Block& b : c.blk) {
    if (side ? (b.rk[d] != c.size[d] - 1) : (b.rk[d] != 0)) continue;
    const int dm = b.dim(d), l = side ? dm : 0, sg = side ? -1 : 1;
    const int n1 = b.im, n2 = (d == 1) ? b.km : b.jm;
    for (int p2 = 0; p2 <= n2; ++p2)
      for (int i = 0; i <= n1; ++i) {
        auto at = [&](Field& f, int off) -> double& { return d == 1 ? f(i, l + sg * off, p2) : f(i, p2, l + sg * off); };
        const double ue = extrapolate2(at(b.vel[0], 1), at(b.vel[0], 2));
        const double ve = extrapolate2(at(b.vel[1], 1), at(b.vel[1], 2));
        const double we = extrapolate2(at(b.vel[2], 1), at(b.vel[2], 2));
        const double pe = extrapolate2(at(b.prs, 1), at(b.prs, 2));
        const double roe = extrapolate2(at(b.rho, 1), at(b.rho, 2));
        const int ii = i, jj = d == 1 ? l : p2, kk = d == 1 ? p2 : l;
        if (ndir == 4) {
          b.prs(ii, jj, kk) = pe; b.rho(ii, jj, kk) = roe;
          b.vel[0](ii, jj, kk) = ue; b.vel[1](ii, jj, kk) = ve; b.vel[2](ii, jj, kk) = we;
          b.tmp(ii, jj, kk) = c.th.thermal_T(pe, roe);
          fvar2q_T(c.th, b, ii, jj, kk);
          continue;
        }
        const double css = c.th.sos(b.tmp(ii, jj, kk));
        const double csse = extrapolate2(c.th.sos(at(b.tmp, 1)), c.th.sos(at(b.tmp, 2)));
        const double ext[3] = {ue, ve, we};
        const double vn = b.vel[d](ii, jj, kk), vne = ext[d], vninf = vinf3[d];
        const bool inflow = side ? (vn <= 0.0) : (vn >= 0.0);
        if (inflow) {
          const double rho0 = b.rho(ii, jj, kk);
          double vnew, pnew;
          if (!side) {
            vnew = 0.5 * (c.pinf - pe) / (rho0 * css) + 0.5 * (vninf + vne);
            pnew = 0.5 * (c.pinf + pe) + 0.5 * rho0 * css * (vninf - vne);
          } else {
            vnew = -0.5 * (c.pinf - pe) / (rho0 * css) + 0.5 * (vninf + vne);
            pnew = 0.5 * (c.pinf + pe) - 0.5 * rho0 * css * (vninf - vne);
          }
          for (int m = 0; m < 3; ++m) b.vel[m](ii, jj, kk) = vinf3[m];
          b.vel[d](ii, jj, kk) = vnew;
          b.prs(ii, jj, kk) = pnew;
          b.rho(ii, jj, kk) = c.roinf * std::pow(pnew / c.pinf, 1.0 / c.th.gamma);
        } else {
          const double pnew = c.pinf;
          b.prs(ii, jj, kk) = pnew;
          b.rho(ii, jj, kk) = roe + (pnew - pe) / csse / csse;
          for (int m = 0; m < 3; ++m) b.vel[m](ii, jj, kk) = ext[m];
          b.vel[d](ii, jj, kk) = side ? vne + (pe - pnew) / roe / csse : vne - (pe - pnew) / roe / csse;
        }
        b.tmp(ii, jj, kk) = c.th.thermal_T(b.prs(ii, jj, kk), b.rho(ii, jj, kk));
        fvar2q_P(c.th, b, ii, jj, kk);
      }
  }
  return 0;
}

// src/bc.F90:7231-7430 slipadibwall(ndir): adiabatic (T extrapolated), p extrapolated, rho=thermal(p,T),
// q=fvar2q(pressure).  ndir==3 (jmin, the SWLBI input, :7255-7373): u extrapolated, v=w=0; ndir==4 (jmax,
// :7375-7424): u and v extrapolated, w=0 -- as the reference writes them.
static int slipadibwall(Case& c, int ndir) {
  if (ndir != 3 && ndir != 4) return -1;
  for (Block& b : c.blk) {
    if (ndir == 3 ? (b.rk[1] != 0) : (b.rk[1] != c.size[1] - 1)) continue;
    const int j = ndir == 3 ? 0 : b.jm, sg = ndir == 3 ? 1 : -1;
    for (int k = 0; k <= b.km; ++k)
      for (int i = 0; i <= b.im; ++i) {
        const double pe = num1d3 * (4.0 * b.prs(i, j + sg, k) - b.prs(i, j + 2 * sg, k));
        const double te = num1d3 * (4.0 * b.tmp(i, j + sg, k) - b.tmp(i, j + 2 * sg, k));
        const double ue = num1d3 * (4.0 * b.vel[0](i, j + sg, k) - b.vel[0](i, j + 2 * sg, k));
        const double ve = num1d3 * (4.0 * b.vel[1](i, j + sg, k) - b.vel[1](i, j + 2 * sg, k));
        b.vel[0](i, j, k) = ue; b.vel[1](i, j, k) = ndir == 3 ? 0.0 : ve; b.vel[2](i, j, k) = 0.0;
        b.tmp(i, j, k) = te; b.prs(i, j, k) = pe;
        b.rho(i, j, k) = c.th.thermal_rho(pe, te);
        fvar2q_P(c.th, b, i, j, k);
      }
  }
  return 0;
}

// src/bc.F90:327-407 boucon: faces in the order n=1..6; only the bctypes restated so far
static int boucon(Case& c) {
  for (int n = 1; n <= 6; ++n) {
    const int bt = c.bctype[n - 1];
    int rc = 0;
    if (bt == 41) noslip(c, n, c.twall[n - 1]);
    else if (bt == 51) rc = farfield(c, n);
    else if (bt == 11) rc = inflow(c, n);
    else if (bt == 21) rc = outflow(c, n);
    else if (bt == 421) rc = slipadibwall(c, n);
    else if (bt != 1) rc = -1;
    if (rc) return rc;
  }
  return 0;
}

// src/sponge_layer.F90:67-319 spongefilter_layer: faces i0, im, jm, k0, km (the reference has no j0 block);
// per face a one-direction dataswap of q, then a damped 7-point average over the layer (Jacobi: qtemp)
// src/sponge_layer.F90:321-364 spongefilter_global (spg_def='circl'): dataswap(q) in every direction, then the
// damped 7-point average over the whole [s:e]^3 box of every block that has a damped node (lsponge_loc)
static void spongefilter_global(Case& c) {
  Getter gq = [](Block& b) { return FieldList{&b.q[0], &b.q[1], &b.q[2], &b.q[3], &b.q[4]}; };
  if (!c.lspg_circ) return;
  for (int d = 0; d < 3; ++d) dataswap(c, gq, d);
  for (Block& b : c.blk) {
    if (b.spg_circ.empty()) continue;
    const int ni = b.e[0] - b.s[0] + 1, nj = b.e[1] - b.s[1] + 1, nk = b.e[2] - b.s[2] + 1;
    std::vector<double> qtemp((size_t)5 * ni * nj * nk);
    for (int n = 0; n < 5; ++n)
      for (int k = b.s[2]; k <= b.e[2]; ++k)
        for (int j = b.s[1]; j <= b.e[1]; ++j)
          for (int i = b.s[0]; i <= b.e[0]; ++i) {
            const size_t t = (size_t)(i - b.s[0]) + (size_t)ni * ((size_t)(j - b.s[1]) + (size_t)nj * (k - b.s[2]));
            const double var1 = b.spg_circ[t];
            const Field& q = b.q[n];
            qtemp[t + (size_t)n * ni * nj * nk] =
                (1.0 - var1) * q(i, j, k) + num1d6 * var1 * (q(i + 1, j, k) + q(i - 1, j, k) + q(i, j + 1, k) +
                                                             q(i, j - 1, k) + q(i, j, k + 1) + q(i, j, k - 1));
          }
    for (int n = 0; n < 5; ++n)
      for (int k = b.s[2]; k <= b.e[2]; ++k)
        for (int j = b.s[1]; j <= b.e[1]; ++j)
          for (int i = b.s[0]; i <= b.e[0]; ++i) {
            const size_t t = (size_t)(i - b.s[0]) + (size_t)ni * ((size_t)(j - b.s[1]) + (size_t)nj * (k - b.s[2]));
            b.q[n](i, j, k) = qtemp[t + (size_t)n * ni * nj * nk];
          }
  }
}

// src/sponge_layer.F90:369-440 spongelayer_define_circle: damping grows with the square of the distance beyond
// `range_spange` from the centre (xc, yc, zc), normalised by its global maximum (pmax) and scaled by dampfac.
// The reference hard-codes centre 0, range 0.06 and dampfac 0.05; they are arguments here so that a test box of any
// size has an undamped core.
static void spongelayer_define_circle(Case& c, double xc, double yc, double zc, double range_spange, double dampfac) {
  double max_dis = 0.0;
  c.lspg_circ = false;
  for (Block& b : c.blk) {
    const int ni = b.e[0] - b.s[0] + 1, nj = b.e[1] - b.s[1] + 1, nk = b.e[2] - b.s[2] + 1;
    b.spg_circ.assign((size_t)ni * nj * nk, 0.0);
    bool loc = false;
    for (int k = b.s[2]; k <= b.e[2]; ++k)
      for (int j = b.s[1]; j <= b.e[1]; ++j)
        for (int i = b.s[0]; i <= b.e[0]; ++i) {
          const double dx = b.x[0](i, j, k) - xc, dy = b.x[1](i, j, k) - yc, dz = b.x[2](i, j, k) - zc;
          const double var1 = std::sqrt(dx * dx + dy * dy + dz * dz);
          double var2 = 0.0;
          if (var1 >= range_spange) { var2 = (var1 - range_spange) * (var1 - range_spange); loc = true; }
          b.spg_circ[(size_t)(i - b.s[0]) + (size_t)ni * ((size_t)(j - b.s[1]) + (size_t)nj * (k - b.s[2]))] = var2;
          max_dis = std::max(max_dis, var2);
        }
    if (loc) c.lspg_circ = true;
    b.spg_circ_loc = loc;
  }
  for (Block& b : c.blk) {
    for (double& v : b.spg_circ) v = v / max_dis * dampfac;
    if (!b.spg_circ_loc) b.spg_circ.clear();     // lsponge_loc false: the block skips the update
  }
}

static void spongefilter(Case& c) {
  Getter gq = [](Block& b) { return FieldList{&b.q[0], &b.q[1], &b.q[2], &b.q[3], &b.q[4]}; };
  if (c.spg_def_circl) { spongefilter_global(c); return; }     // src/sponge_layer.F90:59-63
  static const int faces[5] = {0, 1, 3, 4, 5};
  for (int f : faces) {
    if (!c.lspg[f]) continue;
    const int d = f / 2;
    dataswap(c, gq, d);
    for (Block& b : c.blk) {
      const Block::SpongeLayer& sp = b.spg[f];
      if (sp.beg < 0) continue;
      int lo[3] = {b.s[0], b.s[1], b.s[2]}, hi[3] = {b.e[0], b.e[1], b.e[2]};
      lo[d] = sp.beg; hi[d] = sp.end;
      const int ni = hi[0] - lo[0] + 1, nj = hi[1] - lo[1] + 1, nk = hi[2] - lo[2] + 1;
      std::vector<double> qtemp((size_t)5 * ni * nj * nk);
      for (int n = 0; n < 5; ++n)
        for (int k = lo[2]; k <= hi[2]; ++k)
          for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) {
              const size_t t = (size_t)(i - lo[0]) + (size_t)ni * ((size_t)(j - lo[1]) + (size_t)nj * (k - lo[2]));
              const double var1 = sp.coef[t];
              const Field& q = b.q[n];
              qtemp[t + (size_t)n * ni * nj * nk] =
                  (1.0 - var1) * q(i, j, k) + num1d6 * var1 * (q(i + 1, j, k) + q(i - 1, j, k) + q(i, j + 1, k) +
                                                               q(i, j - 1, k) + q(i, j, k + 1) + q(i, j, k - 1));
            }
      for (int n = 0; n < 5; ++n)
        for (int k = lo[2]; k <= hi[2]; ++k)
          for (int j = lo[1]; j <= hi[1]; ++j)
            for (int i = lo[0]; i <= hi[0]; ++i) {
              const size_t t = (size_t)(i - lo[0]) + (size_t)ni * ((size_t)(j - lo[1]) + (size_t)nj * (k - lo[2]));
              b.q[n](i, j, k) = qtemp[t + (size_t)n * ni * nj * nk];
            }
    }
  }
}

// src/fludyna.F90:191-242 updatefvar
static void updatefvar(Case& c) {
  for (Block& b : c.blk) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) q2fvar_node(c.th, b, i, j, k);
  }
}

// ---------------------------------------------------------------------------------
// crash control (lcracon), src/mainloop.F90:709-1198.  nodestat <= 0 everywhere (no immersed body).
// ---------------------------------------------------------------------------------
static void ensure_crinod(Case& c) {
  for (Block& b : c.blk)
    if (b.crinod.v.empty()) b.crinod.alloc(b.im, b.jm, b.km);    // crinod=.false. (mainloop.F90:81)
}
// crinod_expansion (:985-1033): 3x3x3 dilation of the flags over -1..dim+1, then dataswap(crinod)
static long long crinod_expansion(Case& c) {
  ensure_crinod(c);
  long long counter = 0;
  for (Block& b : c.blk) {
    Field tmp; tmp.alloc(b.im, b.jm, b.km);
    for (int k = -1; k <= b.km + 1; ++k)
      for (int j = -1; j <= b.jm + 1; ++j)
        for (int i = -1; i <= b.im + 1; ++i)
          if (b.crinod(i, j, k) != 0.0)
            for (int k1 = k - 1; k1 <= k + 1; ++k1)
              for (int j1 = j - 1; j1 <= j + 1; ++j1)
                for (int i1 = i - 1; i1 <= i + 1; ++i1) { tmp(i1, j1, k1) = 1.0; counter += 1; }
    for (int k = -2; k <= b.km + 2; ++k)
      for (int j = -2; j <= b.jm + 2; ++j)
        for (int i = -2; i <= b.im + 2; ++i) b.crinod(i, j, k) = tmp(i, j, k);
  }
  Getter gc = [](Block& b) { return FieldList{&b.crinod}; };
  dataswap(c, gc);
  c.crinod_ntimes += 1;
  return counter;
}
// databakup (:826-972).  mode 0 'backup', 1 'recovery'.  Returns 0, or 1 when a recovery finds no backup.
static int databakup(Case& c, int mode) {
  auto save = [&](Case::DataBack& d) {
    d.nstep = c.nstep; d.time = c.time; d.recover_counter = 0;
    d.q.resize(c.blk.size() * 5);
    for (size_t ib = 0; ib < c.blk.size(); ++ib)
      for (int m = 0; m < 5; ++m) d.q[ib * 5 + m] = c.blk[ib].q[m].v;
  };
  auto load = [&](Case::DataBack& d) {
    c.nstep = d.nstep; c.time = d.time;
    for (size_t ib = 0; ib < c.blk.size(); ++ib) {
      Block& b = c.blk[ib];
      for (int m = 0; m < 5; ++m) {
        Field old; old.alloc(b.im, b.jm, b.km); old.v = d.q[ib * 5 + m];
        for (int k = 0; k <= b.km; ++k)               // q(0:im,0:jm,0:km,:) only: the halos keep their values
          for (int j = 0; j <= b.jm; ++j)
            for (int i = 0; i <= b.im; ++i) b.q[m](i, j, k) = old(i, j, k);
      }
    }
    d.recover_counter += 1;
    return d.recover_counter;
  };
  if (mode == 0) {
    if (c.datpnt == 'o') c.datpnt = 'a';
    if (c.datpnt == 'a') { save(c.dat_a); c.datpnt = 'b'; }
    else { save(c.dat_b); c.datpnt = 'a'; }
    return 0;
  }
  if (c.datpnt == 'o') return 1;          // ' !! not backup data avaliable !!'
  // after a single backup datpnt points at the copy that was never written: the reference then reads an unallocated
  // array (:941); reported as "no backup" here
  if ((c.datpnt == 'a' ? c.dat_a : c.dat_b).q.empty()) return 1;
  int counter;
  if (c.datpnt == 'a') { counter = load(c.dat_a); c.datpnt = 'b'; }
  else { counter = load(c.dat_b); c.datpnt = 'a'; }
  updatefvar(c);
  if (counter > 1) crinod_expansion(c);
  return 0;
}
// crashcheck (:709-818), the detection part: fluid nodes whose density is not >= 0 become critical nodes;
// returns how many (the caller recovers or stops, :776-812)
static long long crashcheck(Case& c) {
  ensure_crinod(c);
  long long n = 0;
  for (Block& b : c.blk)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i)
          if (!(b.q[0](i, j, k) >= 0.0)) { b.crinod(i, j, k) = 1.0; n += 1; }
  return n;
}
// crashfix (:1045-1198): nodes whose density, pressure or temperature fell under eps are flagged and replaced by
// the mean of their admissible neighbours, one after the other in storage order (later nodes see earlier repairs)
static long long crashfix(Case& c) {
  ensure_crinod(c);
  const double eps_rho = 1.0e-5, eps_tmp = 1.0e-5;
  const double eps_prs = c.th.thermal_p(eps_rho, eps_tmp);       // thermal(density,temperature), fludyna.F90:45-88
  long long counter = 0;
  for (Block& b : c.blk)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) {
          if (b.rho(i, j, k) >= eps_rho && b.prs(i, j, k) >= eps_prs && b.tmp(i, j, k) >= eps_tmp) continue;
          b.crinod(i, j, k) = 1.0;
          double qavg[5] = {0, 0, 0, 0, 0};
          int norm = 0;
          for (int kk = -1; kk <= 1; ++kk)
            for (int jj = -1; jj <= 1; ++jj)
              for (int ii = -1; ii <= 1; ++ii) {
                if (ii == 0 && jj == 0 && kk == 0) continue;
                if (b.g0[0] + i + ii < 0 || b.g0[0] + i + ii > c.ia || b.g0[1] + j + jj < 0 || b.g0[1] + j + jj > c.ja) continue;
                if (b.rho(i + ii, j + jj, k + kk) >= 0.0 && b.prs(i + ii, j + jj, k + kk) >= 0.0 &&
                    b.tmp(i + ii, j + jj, k + kk) >= 0.0) {
                  for (int m = 0; m < 5; ++m) qavg[m] = qavg[m] + b.q[m](i + ii, j + jj, k + kk);
                  norm += 1;
                }
              }
          if (norm >= 1) {
            for (int m = 0; m < 5; ++m) b.q[m](i, j, k) = qavg[m] / double(norm);
            q2fvar_node(c.th, b, i, j, k);
            counter += 1;
          }
        }
  return counter;
}

// src/statistic.F90:871-990 kenergycal / enstophycal (ndims==3)
static void statcal(Case& c) {
  double ke = 0.0, en = 0.0;
  for (Block& b : c.blk) {
    double ke_b = 0.0, en_b = 0.0;
    for (int k = 1; k <= b.km; ++k)
      for (int j = 1; j <= b.jm; ++j)
        for (int i = 1; i <= b.im; ++i) {
          const double u = b.vel[0](i, j, k), v = b.vel[1](i, j, k), w = b.vel[2](i, j, k);
          const double var1 = u * u + v * v + w * w;
          ke_b = ke_b + b.rho(i, j, k) * var1;
          const double o1 = b.dvel[2][1](i, j, k) - b.dvel[1][2](i, j, k);
          const double o2 = b.dvel[0][2](i, j, k) - b.dvel[2][0](i, j, k);
          const double o3 = b.dvel[1][0](i, j, k) - b.dvel[0][1](i, j, k);
          const double omegam = o1 * o1 + o2 * o2 + o3 * o3;
          en_b = en_b + b.rho(i, j, k) * omegam;
        }
    ke += ke_b; en += en_b;  // psum
  }
  const double cnt = double(c.ia * c.ja * c.ka);
  const double pi = 4.0 * std::atan(1.0);
  const double roinf = 1.0, uinf = 1.0;
  ke = 0.5 * ke / cnt;
  ke = ke / (roinf * uinf * uinf);
  en = 0.5 * en / cnt;
  const double l_0 = c.xmax / (2.0 * pi);
  en = en / (roinf * ((uinf / l_0) * (uinf / l_0)));
  c.hist.push_back(double(c.nstep));
  c.hist.push_back(c.time);
  c.hist.push_back(ke);
  c.hist.push_back(en);
}

// RK update, src/mainloop.F90:441-476
static void rk_update(Case& c, int rkstep /*1-based*/) {
  if (c.rkscheme == 4) {
    // src/mainloop.F90:368-376 (coefficients), :452-476 (update); rhsav is zeroed at rkstep 1 (:436)
    static const double rk4coe[2][4] = {{0.5, 0.5, 1.0, num1d6}, {1.0, 2.0, 2.0, 1.0}};
    const double c1 = rk4coe[0][rkstep - 1], c2 = rk4coe[1][rkstep - 1];
    for (Block& b : c.blk)
      for (int m = 0; m < 5; ++m) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int k = 0; k <= b.km; ++k)
          for (int j = 0; j <= b.jm; ++j)
            for (int i = 0; i <= b.im; ++i) {
              if (rkstep == 1) b.rhsav[m](i, j, k) = 0.0;
              if (rkstep <= 3) {
                double v = b.qsave[m](i, j, k) + c1 * c.deltat * b.qrhs[m](i, j, k);
                b.q[m](i, j, k) = v / b.jacob(i, j, k);
                b.rhsav[m](i, j, k) = b.rhsav[m](i, j, k) + c2 * b.qrhs[m](i, j, k);
              } else {
                double v = b.qsave[m](i, j, k) + c1 * c.deltat * (b.qrhs[m](i, j, k) + b.rhsav[m](i, j, k));
                b.q[m](i, j, k) = v / b.jacob(i, j, k);
              }
            }
      }
    return;
  }
  static const double rkcoe[3][3] = {{1.0, 0.0, 1.0}, {0.75, 0.25, 0.25}, {num1d3, num2d3, num2d3}};
  const double c1 = rkcoe[rkstep - 1][0], c2 = rkcoe[rkstep - 1][1], c3 = rkcoe[rkstep - 1][2];
  for (Block& b : c.blk)
    for (int m = 0; m < 5; ++m) {
#pragma omp parallel for collapse(2) schedule(static)
      for (int k = 0; k <= b.km; ++k)
        for (int j = 0; j <= b.jm; ++j)
          for (int i = 0; i <= b.im; ++i) {
            double v = c1 * b.qsave[m](i, j, k) + c2 * b.q[m](i, j, k) * b.jacob(i, j, k) +
                       c3 * b.qrhs[m](i, j, k) * c.deltat;
            b.q[m](i, j, k) = v / b.jacob(i, j, k);
          }
    }
}

static void save_q(Case& c) {  // mainloop.F90:429-433
  for (Block& b : c.blk)
    for (int m = 0; m < 5; ++m)
      for (int k = 0; k <= b.km; ++k)
        for (int j = 0; j <= b.jm; ++j)
          for (int i = 0; i <= b.im; ++i) b.qsave[m](i, j, k) = b.q[m](i, j, k) * b.jacob(i, j, k);
}

static void zero_qrhs(Case& c) {
  for (Block& b : c.blk)
    for (auto& f : b.qrhs) std::fill(f.v.begin(), f.v.end(), 0.0);
}

// One RK stage, src/mainloop.F90:396-482 (spongefilter is a no-op without sponge layers)
static void rk_stage(Case& c, int rkstep) {
  if (c.lfilter) filterq(c);
  zero_qrhs(c);
  boucon(c);
  qswap(c);
  gradcal(c);
  if (rkstep == 1) {
    save_q(c);
    if (c.ndims() == 3) statcal(c);  // rkfirst -> statcal (the 2-D statistics are not restated)
  }
  rhscal(c);
  rk_update(c, rkstep);
  spongefilter(c);   // mainloop.F90:478
  updatefvar(c);
}

// src/initialisation.F90:621-702 tgvini + fludyna.F90:254-300 updateq
static void tgvini(Case& c) {
  const double roinf = 1.0, uinf = 1.0, l_0 = 1.0;
  const double pinf = roinf * 1.0 / c.th.const2;  // solver.F90:120 pinf=roinf*tinf/const2
  for (Block& b : c.blk)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) {
          const double X = b.x[0](i, j, k), Y = b.x[1](i, j, k), Z = b.x[2](i, j, k);
          b.rho(i, j, k) = roinf;
          b.vel[0](i, j, k) = uinf * std::sin(X / l_0) * std::cos(Y / l_0) * std::cos(Z / l_0);
          b.vel[1](i, j, k) = -uinf * std::cos(X / l_0) * std::sin(Y / l_0) * std::cos(Z / l_0);
          b.vel[2](i, j, k) = 0.0;
          b.prs(i, j, k) = pinf + b.rho(i, j, k) / 16.0 * (uinf * uinf) *
                                      (std::cos(2.0 * X / l_0) + std::cos(2.0 * Y / l_0)) *
                                      (std::cos(2.0 * Z / l_0) + 2.0);
          b.tmp(i, j, k) = b.prs(i, j, k) / b.rho(i, j, k) * c.th.const2;
          // fvar2q_3da with temperature (fludyna.F90:501-505)
          const double r = b.rho(i, j, k), u = b.vel[0](i, j, k), v = b.vel[1](i, j, k), w = b.vel[2](i, j, k);
          b.q[0](i, j, k) = r; b.q[1](i, j, k) = r * u; b.q[2](i, j, k) = r * v; b.q[3](i, j, k) = r * w;
          b.q[4](i, j, k) = r * (b.tmp(i, j, k) * c.th.const1 + 0.5 * (u * u + v * v + w * w));
        }
}

}  // namespace astr_oracle

// ---------------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------------
using namespace astr_oracle;
extern "C" {

// Create a case: global grid ia x ja x ka intervals on an isize x jsize x ksize block
// grid, periodic flags, cube [0,lx]x[0,ly]x[0,lz].  sutherland_s: 110.3 (src/) or 110.4.
void* oracle_case_create(int ia, int ja, int ka, int isize, int jsize, int ksize, int lihomo,
                         int ljhomo, int lkhomo, double lx, double ly, double lz, double reynolds,
                         double mach, double alfa_filter, double deltat, double sutherland_s) {
  Case* c = new Case();
  c->ia = ia; c->ja = ja; c->ka = ka;
  c->size[0] = isize; c->size[1] = jsize; c->size[2] = ksize;
  c->homo[0] = lihomo; c->homo[1] = ljhomo; c->homo[2] = lkhomo;
  c->alfa_filter = alfa_filter; c->deltat = deltat;
  c->th.reynolds = reynolds; c->th.mach = mach; c->th.ref_tem = 273.15;
  c->th.refcal(sutherland_s);
  c->pinf = 1.0 * 1.0 / c->th.const2;   // roinf*tinf/const2, roinf=tinf=1 (src/solver.F90:113-120)
  decompose(*c);
  filter_coefficient_cal(c->fc, alfa_filter, 1.11, 0.98);  // comsolver.F90:121
  for (Block& b : c->blk) {
    alloc_block(b);
    for (int d = 0; d < c->ndims(); ++d) {
      fd_scheme_initiate(b.fds[d], 643, 'c', b.npdc[d], b.dim(d));
      compact_filter_initiate(b.fil[d], b.npdc[d], b.dim(d), alfa_filter);
    }
    // gridcube (gridgeneration.F90:246-262)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) {
          b.x[0](i, j, k) = lx / double(ia) * double(i + b.g0[0]);
          b.x[1](i, j, k) = ly / double(ja) * double(j + b.g0[1]);
          b.x[2](i, j, k) = ka == 0 ? 0.0 : lz / double(ka) * double(k + b.g0[2]);
        }
  }
  c->xmax = lx / double(ia) * double(ia);
  return c;
}
void oracle_case_destroy(void* h) { delete static_cast<Case*>(h); }
int oracle_case_nblocks(void* h) { return int(static_cast<Case*>(h)->blk.size()); }
// info[0..2]=im,jm,km  [3..5]=npdc  [6..11]=is,ie,js,je,ks,ke  [12..14]=ig0..  [15..20]=neighbours
void oracle_case_block_info(void* h, int ib, int* info) {
  const Block& b = static_cast<Case*>(h)->blk[ib];
  info[0] = b.im; info[1] = b.jm; info[2] = b.km;
  for (int d = 0; d < 3; ++d) {
    info[3 + d] = b.npdc[d];
    info[6 + 2 * d] = b.s[d]; info[7 + 2 * d] = b.e[d];
    info[12 + d] = b.g0[d];
    info[15 + 2 * d] = b.nb[d][0]; info[16 + 2 * d] = b.nb[d][1];
  }
}
// Overwrite the node coordinates of a block (0:im,0:jm,0:km, Fortran order, 3 comps)
void oracle_case_set_x(void* h, int ib, const double* x) {
  Block& b = static_cast<Case*>(h)->blk[ib];
  size_t t = 0;
  for (int m = 0; m < 3; ++m)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) b.x[m](i, j, k) = x[t++];
}
void oracle_case_gridgeom(void* h) { gridgeom(*static_cast<Case*>(h)); }
void oracle_case_tgvini(void* h) { tgvini(*static_cast<Case*>(h)); }

static Field* field_by_id(Block& b, int id) {
  // 0-4 q | 5 rho | 6-8 vel | 9 prs | 10 tmp | 11-15 qrhs | 16 jacob | 17-25 dxi(a,b) a-major
  // 26-34 dvel(m,n) m-major | 35-37 dtmp | 38-43 sigma | 44-46 qflux | 47-49 x | 50-54 qsave
  // 55-57 vor
  if (id < 5) return &b.q[id];
  if (id == 5) return &b.rho;
  if (id < 9) return &b.vel[id - 6];
  if (id == 9) return &b.prs;
  if (id == 10) return &b.tmp;
  if (id < 16) return &b.qrhs[id - 11];
  if (id == 16) return &b.jacob;
  if (id < 26) return &b.dxi[(id - 17) / 3][(id - 17) % 3];
  if (id < 35) return &b.dvel[(id - 26) / 3][(id - 26) % 3];
  if (id < 38) return &b.dtmp[id - 35];
  if (id < 44) return &b.sigma[id - 38];
  if (id < 47) return &b.qflux[id - 44];
  if (id < 50) return &b.x[id - 47];
  if (id < 55) return &b.qsave[id - 50];
  if (id < 58) return &b.vor[id - 55];
  if (id == 58) return &b.ssf;
  if (id == 59) return &b.lshock;
  if (id == 60) { if (b.crinod.v.empty()) b.crinod.alloc(b.im, b.jm, b.km); return &b.crinod; }
  return nullptr;
}
// Copy a field (always in the halo'd shape (im+11)(jm+11)(km+11)) out of / into a block.
void oracle_case_get(void* h, int ib, int id, double* out) {
  Field* f = field_by_id(static_cast<Case*>(h)->blk[ib], id);
  std::memcpy(out, f->v.data(), f->v.size() * sizeof(double));
}
void oracle_case_set(void* h, int ib, int id, const double* in) {
  Field* f = field_by_id(static_cast<Case*>(h)->blk[ib], id);
  std::memcpy(f->v.data(), in, f->v.size() * sizeof(double));
}
// Stage operators, individually callable so each can be compared with its CUDA twin.
void oracle_case_filterq(void* h) { filterq(*static_cast<Case*>(h)); }
void oracle_case_qswap(void* h) { qswap(*static_cast<Case*>(h)); }
void oracle_case_gradcal(void* h) { gradcal(*static_cast<Case*>(h)); }
void oracle_case_zero_qrhs(void* h) { zero_qrhs(*static_cast<Case*>(h)); }
void oracle_case_convrsdcal6(void* h) { convrsdcal6(*static_cast<Case*>(h)); }
void oracle_case_rhscal(void* h) { rhscal(*static_cast<Case*>(h)); }
void oracle_case_save_q(void* h) { save_q(*static_cast<Case*>(h)); }
void oracle_case_rk_update(void* h, int rkstep) { rk_update(*static_cast<Case*>(h), rkstep); }
void oracle_case_updatefvar(void* h) { updatefvar(*static_cast<Case*>(h)); }
// updateq (src/fludyna.F90:254-300): q(0:im,0:jm,0:km,:) from density, velocity and temperature
void oracle_case_updateq(void* h) {
  Case* c = static_cast<Case*>(h);
  for (Block& b : c->blk)
    for (int k = 0; k <= b.km; ++k)
      for (int j = 0; j <= b.jm; ++j)
        for (int i = 0; i <= b.im; ++i) fvar2q_T(c->th, b, i, j, k);
}
void oracle_case_rk_stage(void* h, int rkstep) { rk_stage(*static_cast<Case*>(h), rkstep); }
int oracle_case_boucon(void* h) { return boucon(*static_cast<Case*>(h)); }
// bctype(1:6), twall(1:6) of the input file; flowtype 0 generic / 1 channel; force(1:3) of
// src_chan; scheme_kind 'c' (643c) or 'e' (642e: diff6ec everywhere fds%central is called)
void oracle_case_set_bc(void* h, const int* bctype, const double* twall) {
  Case* c = static_cast<Case*>(h);
  for (int n = 0; n < 6; ++n) { c->bctype[n] = bctype[n]; c->twall[n] = twall[n]; }
}
// inflow data of block ib: vel_in(0:jm,0:km,3), tmp_in(0:jm,0:km), tmp_prof(0:jm) (Fortran order)
void oracle_case_set_inflow(void* h, int ib, const double* vel_in, const double* tmp_in, const double* tmp_prof) {
  Case* c = static_cast<Case*>(h);
  const Block& b = c->blk[ib];
  const size_t nf = (size_t)(b.jm + 1) * (b.km + 1);
  c->vel_in.resize(c->blk.size()); c->tmp_in.resize(c->blk.size()); c->tmp_prof.resize(c->blk.size());
  c->vel_in[ib].assign(vel_in, vel_in + 3 * nf);
  c->tmp_in[ib].assign(tmp_in, tmp_in + nf);
  c->tmp_prof[ib].assign(tmp_prof, tmp_prof + b.jm + 1);
}
// sponge layer of face f (0 i0, 1 im, 3 jm, 4 k0, 5 km) on block ib: node range beg..end along the face
// direction (beg<0: the layer exists but not on this block) and the damping coefficients over the box
void oracle_case_set_sponge(void* h, int ib, int face, int beg, int end, const double* coef) {
  Case* c = static_cast<Case*>(h);
  Block& b = c->blk[ib];
  c->lspg[face] = true;
  b.spg[face].beg = beg; b.spg[face].end = end;
  if (beg >= 0) {
    const int d = face / 2;
    size_t cnt = (size_t)(end - beg + 1);
    for (int o = 0; o < 3; ++o) if (o != d) cnt *= (size_t)(b.e[o] - b.s[o] + 1);
    b.spg[face].coef.assign(coef, coef + cnt);
  }
}
void oracle_case_spongefilter(void* h) { spongefilter(*static_cast<Case*>(h)); }
// crash control (src/mainloop.F90:709-1198)
long long oracle_case_crashcheck(void* h) { return crashcheck(*static_cast<Case*>(h)); }
long long oracle_case_crashfix(void* h) { return crashfix(*static_cast<Case*>(h)); }
long long oracle_case_crinod_expansion(void* h) { return crinod_expansion(*static_cast<Case*>(h)); }
int oracle_case_databakup(void* h, int mode) { return databakup(*static_cast<Case*>(h), mode); }
int oracle_case_nstep(void* h) { return static_cast<Case*>(h)->nstep; }
// spg_def='circl': defines the coefficients (spongelayer_define_circle) and switches spongefilter to the global form
void oracle_case_set_sponge_circle(void* h, double xc, double yc, double zc, double range_spange, double dampfac) {
  Case* c = static_cast<Case*>(h);
  c->spg_def_circl = true;
  spongelayer_define_circle(*c, xc, yc, zc, range_spange, dampfac);
}
// sponge_damp_coef(is:ie,js:je,ks:ke) of block ib (Fortran order); returns 0 when the block has no damped node
int oracle_case_sponge_circle_coef(void* h, int ib, double* out) {
  const Block& b = static_cast<Case*>(h)->blk[ib];
  if (b.spg_circ.empty()) return 0;
  std::memcpy(out, b.spg_circ.data(), b.spg_circ.size() * sizeof(double));
  return 1;
}
double oracle_case_pinf(void* h) { return static_cast<Case*>(h)->pinf; }
// nondimen=f (src/solver.F90:124-148): SI reference state; Mach, Reynolds, const1..7, pinf follow
void oracle_case_set_dimensional(void* h, double ref_tem, double ref_vel, double ref_len, double ref_den) {
  Case* c = static_cast<Case*>(h);
  c->th.ref_tem = ref_tem; c->th.ref_vel = ref_vel; c->th.ref_len = ref_len; c->th.ref_den = ref_den;
  c->th.refcal_dimensional();
  c->pinf = c->th.pinf;
  // src/solver.F90:131-139: uinf=ref_vel, vinf=winf=0, roinf=ref_den (tinf=ref_tem, pinf=thermal(tinf,roinf))
  c->uinf = ref_vel; c->vinf = 0.0; c->winf = 0.0; c->roinf = ref_den;
}
// out: reynolds, mach, const1..const7, rgas, cp, cv, pinf, nondimen
void oracle_case_thermo(void* h, double* out) {
  const Thermo& t = static_cast<Case*>(h)->th;
  const double v[14] = {t.reynolds, t.mach, t.const1, t.const2, t.const3, t.const4, t.const5, t.const6, t.const7,
                        t.rgas, t.cp, t.cv, static_cast<Case*>(h)->pinf, t.nondimen ? 1.0 : 0.0};
  std::memcpy(out, v, sizeof v);
}
void oracle_case_set_flow(void* h, int flowtype, const double* force) {
  Case* c = static_cast<Case*>(h);
  c->flowtype = flowtype;
  for (int n = 0; n < 3; ++n) c->force[n] = force[n];
}
// conschm digits (643 / 543) and the upwind parameters lchardecomp, bfacmpld, shkcrt
void oracle_case_set_upwind(void* h, int conschm, int lchardecomp, double bfacmpld, double shkcrt) {
  Case* c = static_cast<Case*>(h);
  c->conschm = conschm; c->conschm_explicit = false;
  c->lchardecomp = lchardecomp != 0; c->bfacmpld = bfacmpld; c->shkcrt = shkcrt;
}
// explicit upwind family: conschm='<odd>..e' with recon_schem (-1, 0, 1, 2, 3, 5, 6)
void oracle_case_set_upwind_explicit(void* h, int recon_schem, int lchardecomp, double bfacmpld, double shkcrt) {
  Case* c = static_cast<Case*>(h);
  c->conschm = 753; c->conschm_explicit = true; c->recon_schem = recon_schem;
  c->lchardecomp = lchardecomp != 0; c->bfacmpld = bfacmpld; c->shkcrt = shkcrt;
}
int oracle_case_convrsduwd(void* h) { return convrsduwd(*static_cast<Case*>(h)); }
double oracle_recons_exp(const double* f8, int inode, int dim, int ntype, int reschem, int shock, double bfacmpld) {
  return recons_exp(f8, inode, dim, ntype, reschem, shock != 0, bfacmpld);
}
void oracle_case_ducrossensor(void* h) { ducrossensor(*static_cast<Case*>(h)); }
int oracle_case_convrsdcmp(void* h) { return convrsdcmp(*static_cast<Case*>(h)); }
// single-pencil entry points of the upwind building blocks (tests)
void oracle_flux_compact(int ntype, int dim, int plus, double bfacmpld, const double* f /*-hm..dim+hm*/, double* fh /*-1..dim*/) {
  CompactScheme s;
  compact_flux_initiate(s, 543, ntype, dim, plus ? '+' : '-', bfacmpld);
  std::vector<double> work(2 * s.size());
  flux_compact(s, plus ? '+' : '-', bfacmpld, f + hm, fh + 1, work.data());
}
double oracle_mp5(const double* u5, double ul, int discont) { return mp5(u5, ul, discont != 0); }
void oracle_steger_warming(double gamma, double mach, const double* in /*rho,vel3,prs,tmp,q5,dxi3,jacob*/, double* fp, double* fm) {
  Thermo th; th.gamma = gamma; th.mach = mach;
  steger_warming_node(th, in[0], in + 1, in[4], in[5], in + 6, in + 11, in[14], fp, fm);
}
int oracle_chardecomp(double gamma, const double* l /*ro,p,E,vel3,ddi3*/, const double* r, double* rev25, double* lev25) {
  double REV[5][5], LEV[5][5];
  const bool ok = chardecomp(gamma, l[0], l[1], l[2], l + 3, l + 6, r[0], r[1], r[2], r + 3, r + 6, REV, LEV);
  std::memcpy(rev25, REV, sizeof REV); std::memcpy(lev25, LEV, sizeof LEV);
  return ok ? 0 : 1;
}
void oracle_case_set_scheme(void* h, int explicit_scheme) {
  static_cast<Case*>(h)->scheme_kind = explicit_scheme ? 'e' : 'c';
}
// rkscheme 'rk3' / 'rk4' (src/mainloop.F90:348-388)
int oracle_case_set_rkscheme(void* h, int scheme) {
  Case* c = static_cast<Case*>(h);
  if (scheme != 3 && scheme != 4) return 1;
  c->rkscheme = scheme;
  if (scheme == 4)
    for (Block& b : c->blk)
      for (int m = 0; m < 5; ++m) { b.rhsav[m] = b.qsave[m]; std::fill(b.rhsav[m].v.begin(), b.rhsav[m].v.end(), 0.0); }
  return 0;
}
void oracle_case_set_flags(void* h, int lfilter, int diffterm) {
  Case* c = static_cast<Case*>(h);
  c->lfilter = lfilter; c->diffterm = diffterm;
}
// nsteps full RK3 steps (mainloop.F90:103-205: rk stages, then nstep++, time+=dt)
// ---- per-step diagnostics of rkfirst / steploop, as raw partial sums / maxima summed (psum) or
// maximised (pmax) over the blocks, BEFORE the reference's normalisation by ia*ja*ka etc. ----
// what = 0: kenergycal, enstophycal, diss_rate_cal  (src/statistic.F90:938, :871, :994)    -> out[0..2]
//        1: cflcal deltai, deltaj, deltak           (src/commcal.F90:27-74)                -> out[0..2]
//        2: massfluxchan, fbcxchan                  (src/statistic.F90:1437-1476, :1303)   -> out[0..1]
// Needs gradcal of the current state (dvel) for 0 and 2.
void oracle_case_reduce(void* h, int what, double* out) {
  Case& c = *static_cast<Case*>(h);
  const Thermo& th = c.th;
  out[0] = out[1] = out[2] = 0.0;
  const int jrkm = c.size[1] - 1;
  for (Block& b : c.blk) {
    if (what == 0) {
      double ke = 0.0, en = 0.0, di = 0.0;
      for (int k = 1; k <= b.km; ++k)
        for (int j = 1; j <= b.jm; ++j)
          for (int i = 1; i <= b.im; ++i) {
            const double u = b.vel[0](i, j, k), v = b.vel[1](i, j, k), w = b.vel[2](i, j, k);
            const double var1 = u * u + v * v + w * w;
            ke = ke + b.rho(i, j, k) * var1;
            const double o1 = b.dvel[2][1](i, j, k) - b.dvel[1][2](i, j, k);
            const double o2 = b.dvel[0][2](i, j, k) - b.dvel[2][0](i, j, k);
            const double o3 = b.dvel[1][0](i, j, k) - b.dvel[0][1](i, j, k);
            en = en + b.rho(i, j, k) * (o1 * o1 + o2 * o2 + o3 * o3);
            // diss_rate_cal, src/statistic.F90:1011-1036
            const double miu = th.miu_eff(b.tmp(i, j, k));
            const double du11 = b.dvel[0][0](i, j, k), du12 = b.dvel[0][1](i, j, k), du13 = b.dvel[0][2](i, j, k);
            const double du21 = b.dvel[1][0](i, j, k), du22 = b.dvel[1][1](i, j, k), du23 = b.dvel[1][2](i, j, k);
            const double du31 = b.dvel[2][0](i, j, k), du32 = b.dvel[2][1](i, j, k), du33 = b.dvel[2][2](i, j, k);
            const double s11 = du11, s12 = 0.5 * (du12 + du21), s13 = 0.5 * (du13 + du31);
            const double s22 = du22, s23 = 0.5 * (du23 + du32), s33 = du33;
            const double div = s11 + s22 + s33;
            const double var2 = 2.0 * miu * (s11 * s11 + s22 * s22 + s33 * s33 + 2.0 * (s12 * s12 + s13 * s13 + s23 * s23) -
                                             num1d3 * (div * div));
            di = di + var2;
          }
      out[0] += ke; out[1] += en; out[2] += di;
    } else if (what == 1) {
      double di = 0.0, dj = 0.0, dk = 0.0;
      for (int k = 0; k <= b.km; ++k)
        for (int j = 0; j <= b.jm; ++j)
          for (int i = 0; i <= b.im; ++i) {
            const double u = b.vel[0](i, j, k), v = b.vel[1](i, j, k), w = b.vel[2](i, j, k);
            const double ubar = b.dxi[0][0](i, j, k) * u + b.dxi[0][1](i, j, k) * v + b.dxi[0][2](i, j, k) * w;
            const double vbar = b.dxi[1][0](i, j, k) * u + b.dxi[1][1](i, j, k) * v + b.dxi[1][2](i, j, k) * w;
            const double wbar = b.dxi[2][0](i, j, k) * u + b.dxi[2][1](i, j, k) * v + b.dxi[2][2](i, j, k) * w;
            const double css = th.sos(b.tmp(i, j, k));
            auto nrm = [&](int a) {
              const double d1 = b.dxi[a][0](i, j, k), d2 = b.dxi[a][1](i, j, k), d3 = b.dxi[a][2](i, j, k);
              return Thermo::std_sqrt(d1 * d1 + d2 * d2 + d3 * d3);
            };
            const double csi = css * nrm(0), csj = css * nrm(1), csk = css * nrm(2);
            di = std::max(std::max(di, ubar), std::max(ubar - csi, ubar + csi));
            dj = std::max(std::max(dj, vbar), std::max(vbar - csj, vbar + csj));
            dk = std::max(std::max(dk, wbar), std::max(wbar - csk, wbar + csk));
          }
      out[0] = std::max(out[0], di); out[1] = std::max(out[1], dj); out[2] = std::max(out[2], dk);
    } else {
      const int k1 = c.ndims() == 2 ? 0 : 1, k2 = c.ndims() == 2 ? 0 : b.km;
      double mf = 0.0, fb = 0.0;
      for (int k = k1; k <= k2; ++k)
        for (int j = 1; j <= b.jm; ++j)
          for (int i = 1; i <= b.im; ++i) {
            const double dy = b.x[1](i, j, k) - b.x[1](i, j - 1, k);
            const double var1 = 0.5 * (b.q[1](i, j, k) + b.q[1](i, j - 1, k));
            mf = mf + var1 * dy;
          }
      if (b.rk[1] == 0)
        for (int k = k1; k <= k2; ++k)
          for (int i = 1; i <= b.im; ++i) fb = fb + th.miu_eff(b.tmp(i, 0, k)) * b.dvel[0][1](i, 0, k);
      if (b.rk[1] == jrkm)
        for (int k = k1; k <= k2; ++k)
          for (int i = 1; i <= b.im; ++i) fb = fb - th.miu_eff(b.tmp(i, b.jm, k)) * b.dvel[0][1](i, b.jm, k);
      out[0] += mf; out[1] += fb;
    }
  }
}

int oracle_case_run(void* h, int nsteps) {
  Case* c = static_cast<Case*>(h);
  for (int s = 0; s < nsteps; ++s) {
    for (int rk = 1; rk <= c->rkscheme; ++rk) rk_stage(*c, rk);
    c->nstep += 1;
    c->time = c->time + c->deltat;
  }
  return int(c->hist.size() / 4);
}
void oracle_case_history(void* h, double* out) {
  Case* c = static_cast<Case*>(h);
  std::memcpy(out, c->hist.data(), c->hist.size() * sizeof(double));
}
}
