// oracle/miniapp.cpp -- TEST INFRASTRUCTURE ONLY (see astr_oracle.hpp).
// Restatement of the reference's self-contained single-block TGV solver,
// miniapps/tgv_solver_3d/tgvsolver.F90, which `use`s the real src/derivative.F90,
// src/filter.F90 and src/commfunc.F90.  This is the mode that is pinned against the
// golden history miniapps/tgv_solver_3d/state.ref_128.
#include "astr_oracle.hpp"
#include <cmath>
#include <cstring>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace astr_oracle {

struct MiniApp {
  int im, jm, km, numq = 5;
  double alfa_filter = 0.49, deltat = 1.0e-3, time = 0.0;
  int nstep = 0;
  double dx, dy, dz;
  Thermo th;
  CompactScheme fds_i, fds_j, fds_k, fil_i, fil_j, fil_k;
  FilterCoef fc;
  Field q[5], rho, prs, tmp, vel[3];
  Field qrhs[5], dvel[3][3], dtmp[3], qsave[5], sigma[6], qflux[3];
  std::vector<double> hist;  // rows of (nstep,time,tke,enst), written by stacal

  void program_init(int n);
  void solver_init();
  void flowfield_init();
  void bchomo();
  void bchomovec(Field* var, int nvar, int dir);
  void gradcal();
  void convection();
  void diffusion();
  void filterq();
  void stacal();
  void rhscal();
  void rk3();
};

static inline int maxdim(const MiniApp& m) { return std::max(m.im, std::max(m.jm, m.km)); }

// tgvsolver.F90:152-170 (im=jm=km=128 in the shipped program; size is a parameter here)
void MiniApp::program_init(int n) {
  im = jm = km = n;
  alfa_filter = 0.49;
}

// tgvsolver.F90:236-261
void MiniApp::solver_init() {
  fd_scheme_initiate(fds_i, 643, 'c', 3, im);
  fd_scheme_initiate(fds_j, 643, 'c', 3, jm);
  fd_scheme_initiate(fds_k, 643, 'c', 3, km);
  filter_coefficient_cal(fc, alfa_filter, 1.11, 1.09);  // :253 (beter_bouond=1.09 here)
  compact_filter_initiate(fil_i, 3, im, alfa_filter);
  compact_filter_initiate(fil_j, 3, jm, alfa_filter);
  compact_filter_initiate(fil_k, 3, km, alfa_filter);
}

// tgvsolver.F90:13-38 + 172-234
void MiniApp::flowfield_init() {
  for (auto& f : q) f.alloc(im, jm, km);
  rho.alloc(im, jm, km); prs.alloc(im, jm, km); tmp.alloc(im, jm, km);
  for (auto& f : vel) f.alloc(im, jm, km);
  for (auto& f : qrhs) f.alloc(im, jm, km);
  for (auto& r : dvel) for (auto& f : r) f.alloc(im, jm, km);
  for (auto& f : dtmp) f.alloc(im, jm, km);
  for (auto& f : qsave) f.alloc(im, jm, km);
  for (auto& f : sigma) f.alloc(im, jm, km);
  for (auto& f : qflux) f.alloc(im, jm, km);

  th.ref_tem = 273.15; th.gamma = 1.4; th.mach = 0.1; th.reynolds = 1600.0; th.prandtl = 0.72;
  th.refcal(110.4);  // tgvsolver.F90:133
  const double pi = 4.0 * std::atan(1.0);
  const double pinf = 1.0 / th.const2;
  dx = 2.0 * pi / double(im);
  dy = 2.0 * pi / double(jm);
  dz = 2.0 * pi / double(km);
  for (int k = 0; k <= km; ++k)
    for (int j = 0; j <= jm; ++j)
      for (int i = 0; i <= im; ++i) {
        const double x1 = 2.0 * pi / double(im) * double(i);
        const double x2 = 2.0 * pi / double(jm) * double(j);
        const double x3 = 2.0 * pi / double(km) * double(k);
        rho(i, j, k) = 1.0;
        vel[0](i, j, k) = std::sin(x1) * std::cos(x2) * std::cos(x3);
        vel[1](i, j, k) = -std::cos(x1) * std::sin(x2) * std::cos(x3);
        vel[2](i, j, k) = 0.0;
        prs(i, j, k) = pinf + 1.0 / 16.0 * (std::cos(2.0 * x1) + std::cos(2.0 * x2)) *
                                  (std::cos(2.0 * x3) + 2.0);
        tmp(i, j, k) = prs(i, j, k) / rho(i, j, k) * th.const2;  // thermal_scar :86
        // fvar2q with pressure :59-68
        const double r = rho(i, j, k), u = vel[0](i, j, k), v = vel[1](i, j, k), w = vel[2](i, j, k);
        q[0](i, j, k) = r;
        q[1](i, j, k) = r * u;
        q[2](i, j, k) = r * v;
        q[3](i, j, k) = r * w;
        const double var1 = 0.5 * ((u * u + v * v) + w * w);
        q[4](i, j, k) = prs(i, j, k) * th.const6 + r * var1;
      }
  nstep = 0; time = 0.0; deltat = 1.0e-3;
}

// tgvsolver.F90:271-369  periodic halos of primitives, q rebuilt with fvar2q(pressure)
void MiniApp::bchomo() {
  auto fill = [&](int i, int j, int k, int is, int js, int ks) {
    rho(i, j, k) = rho(is, js, ks);
    for (int m = 0; m < 3; ++m) vel[m](i, j, k) = vel[m](is, js, ks);
    tmp(i, j, k) = tmp(is, js, ks);
    prs(i, j, k) = prs(is, js, ks);
    const double r = rho(i, j, k), u = vel[0](i, j, k), v = vel[1](i, j, k), w = vel[2](i, j, k);
    q[0](i, j, k) = r;
    q[1](i, j, k) = r * u;
    q[2](i, j, k) = r * v;
    q[3](i, j, k) = r * w;
    const double var1 = 0.5 * ((u * u + v * v) + w * w);
    q[4](i, j, k) = prs(i, j, k) * th.const6 + r * var1;
  };
  for (int k = 0; k <= km; ++k)
    for (int j = 0; j <= jm; ++j) {
      for (int i = -hm; i <= -1; ++i) fill(i, j, k, im + i, j, k);
      for (int i = im + 1; i <= im + hm; ++i) fill(i, j, k, i - im, j, k);
    }
  for (int k = 0; k <= km; ++k)
    for (int i = 0; i <= im; ++i) {
      for (int j = -hm; j <= -1; ++j) fill(i, j, k, i, jm + j, k);
      for (int j = jm + 1; j <= jm + hm; ++j) fill(i, j, k, i, j - jm, k);
    }
  for (int j = 0; j <= jm; ++j)
    for (int i = 0; i <= im; ++i) {
      for (int k = -hm; k <= -1; ++k) fill(i, j, k, i, j, km + k);
      for (int k = km + 1; k <= km + hm; ++k) fill(i, j, k, i, j, k - km);
    }
}

// tgvsolver.F90:371-440
void MiniApp::bchomovec(Field* var, int nvar, int dir) {
  for (int n = 0; n < nvar; ++n) {
    Field& a = var[n];
    if (dir == 1 || dir == 0)
      for (int k = 0; k <= km; ++k)
        for (int j = 0; j <= jm; ++j) {
          for (int i = -hm; i <= -1; ++i) a(i, j, k) = a(im + i, j, k);
          for (int i = im + 1; i <= im + hm; ++i) a(i, j, k) = a(i - im, j, k);
        }
    if (dir == 2 || dir == 0)
      for (int k = 0; k <= km; ++k)
        for (int i = 0; i <= im; ++i) {
          for (int j = -hm; j <= -1; ++j) a(i, j, k) = a(i, jm + j, k);
          for (int j = jm + 1; j <= jm + hm; ++j) a(i, j, k) = a(i, j - jm, k);
        }
    if (dir == 3 || dir == 0)
      for (int j = 0; j <= jm; ++j)
        for (int i = 0; i <= im; ++i) {
          for (int k = -hm; k <= -1; ++k) a(i, j, k) = a(i, j, km + k);
          for (int k = km + 1; k <= km + hm; ++k) a(i, j, k) = a(i, j, k - km);
        }
  }
}

// Helper: apply `op` along direction dir (1,2,3) to every pencil of `nin` input
// fields; the callback receives the gathered pencils (pointing at node 0) and the
// pencil's fixed indices, and scatters the result itself.
template <class Fn>
static void for_each_pencil(int im, int jm, int km, int dir, Fn fn) {
  if (dir == 1) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= km; ++k)
      for (int j = 0; j <= jm; ++j) fn(j, k);
  } else if (dir == 2) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= km; ++k)
      for (int i = 0; i <= im; ++i) fn(i, k);
  } else {
#pragma omp parallel for collapse(2) schedule(static)
    for (int j = 0; j <= jm; ++j)
      for (int i = 0; i <= im; ++i) fn(i, j);
  }
}

static inline void gather(const Field& a, int dir, int p1, int p2, int dim, double* buf) {
  // buf points at node 0; fills buf[-hm..dim+hm]
  if (dir == 1) for (int i = -hm; i <= dim + hm; ++i) buf[i] = a(i, p1, p2);
  else if (dir == 2) for (int j = -hm; j <= dim + hm; ++j) buf[j] = a(p1, j, p2);
  else for (int k = -hm; k <= dim + hm; ++k) buf[k] = a(p1, p2, k);
}
static inline double& at(Field& a, int dir, int p1, int p2, int l) {
  return dir == 1 ? a(l, p1, p2) : (dir == 2 ? a(p1, l, p2) : a(p1, p2, l));
}
static inline double at(const Field& a, int dir, int p1, int p2, int l) {
  return dir == 1 ? a(l, p1, p2) : (dir == 2 ? a(p1, l, p2) : a(p1, p2, l));
}

// tgvsolver.F90:450-537
void MiniApp::gradcal() {
  const int md = maxdim(*this);
  for (int dir = 1; dir <= 3; ++dir) {
    const int dim = dir == 1 ? im : (dir == 2 ? jm : km);
    const double h = dir == 1 ? dx : (dir == 2 ? dy : dz);
    const CompactScheme& s = dir == 1 ? fds_i : (dir == 2 ? fds_j : fds_k);
    for_each_pencil(im, jm, km, dir, [&](int p1, int p2) {
      std::vector<double> ffb(md + 1 + 2 * hm), df(md + 1), work(2 * (md + 3));
      double* ff = ffb.data() + hm;
      for (int n = 0; n < 4; ++n) {
        const Field& src = n < 3 ? vel[n] : tmp;
        gather(src, dir, p1, p2, dim, ff);
        df_compact(s, ff, df.data(), work.data());
        Field& dst = n < 3 ? dvel[n][dir - 1] : dtmp[dir - 1];
        for (int l = 0; l <= dim; ++l) at(dst, dir, p1, p2, l) = df[l] / h;
      }
    });
  }
}

// tgvsolver.F90:539-633
void MiniApp::convection() {
  const int md = maxdim(*this);
  for (int dir = 1; dir <= 3; ++dir) {
    const int dim = dir == 1 ? im : (dir == 2 ? jm : km);
    const double h = dir == 1 ? dx : (dir == 2 ? dy : dz);
    const CompactScheme& s = dir == 1 ? fds_i : (dir == 2 ? fds_j : fds_k);
    for_each_pencil(im, jm, km, dir, [&](int p1, int p2) {
      const int np = md + 1 + 2 * hm;
      std::vector<double> fb(5 * np), df(md + 1), work(2 * (md + 3));
      double* fcs[5];
      for (int n = 0; n < 5; ++n) fcs[n] = fb.data() + n * np + hm;
      for (int l = -hm; l <= dim + hm; ++l) {
        const double uu = at(vel[dir - 1], dir, p1, p2, l);
        const double p = at(prs, dir, p1, p2, l);
        for (int n = 0; n < 4; ++n) {
          double v = at(q[n], dir, p1, p2, l) * uu;
          if (n == dir) v = v + p;  // momentum component aligned with the sweep
          fcs[n][l] = v;
        }
        fcs[4][l] = (at(q[4], dir, p1, p2, l) + p) * uu;
      }
      for (int n = 0; n < 5; ++n) {
        df_compact(s, fcs[n], df.data(), work.data());
        for (int l = 0; l <= dim; ++l) {
          double& r = at(qrhs[n], dir, p1, p2, l);
          r = r + df[l] / h;
        }
      }
    });
  }
}

// tgvsolver.F90:635-789
void MiniApp::diffusion() {
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k <= km; ++k)
    for (int j = 0; j <= jm; ++j)
      for (int i = 0; i <= im; ++i) {
        const double s11 = dvel[0][0](i, j, k);
        const double s12 = 0.5 * (dvel[0][1](i, j, k) + dvel[1][0](i, j, k));
        const double s13 = 0.5 * (dvel[0][2](i, j, k) + dvel[2][0](i, j, k));
        const double s22 = dvel[1][1](i, j, k);
        const double s23 = 0.5 * (dvel[1][2](i, j, k) + dvel[2][1](i, j, k));
        const double s33 = dvel[2][2](i, j, k);
        const double skk = num1d3 * (s11 + s22 + s33);
        const double miu = th.miucal(tmp(i, j, k)) / th.reynolds;
        const double miu2 = 2.0 * miu;
        const double hcc = (miu / th.prandtl) / th.const5;
        const double sg1 = miu2 * (s11 - skk), sg2 = miu2 * s12, sg3 = miu2 * s13;
        const double sg4 = miu2 * (s22 - skk), sg5 = miu2 * s23, sg6 = miu2 * (s33 - skk);
        sigma[0](i, j, k) = sg1; sigma[1](i, j, k) = sg2; sigma[2](i, j, k) = sg3;
        sigma[3](i, j, k) = sg4; sigma[4](i, j, k) = sg5; sigma[5](i, j, k) = sg6;
        const double u = vel[0](i, j, k), v = vel[1](i, j, k), w = vel[2](i, j, k);
        qflux[0](i, j, k) = hcc * dtmp[0](i, j, k) + sg1 * u + sg2 * v + sg3 * w;
        qflux[1](i, j, k) = hcc * dtmp[1](i, j, k) + sg2 * u + sg4 * v + sg5 * w;
        qflux[2](i, j, k) = hcc * dtmp[2](i, j, k) + sg3 * u + sg5 * v + sg6 * w;
      }
  bchomovec(sigma, 6, 0);
  bchomovec(qflux, 3, 0);
  const int md = maxdim(*this);
  // which sigma components feed the momentum equations per direction (:707-710,738-741,769-772)
  static const int sel[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
  for (int dir = 1; dir <= 3; ++dir) {
    const int dim = dir == 1 ? im : (dir == 2 ? jm : km);
    const double h = dir == 1 ? dx : (dir == 2 ? dy : dz);
    const CompactScheme& s = dir == 1 ? fds_i : (dir == 2 ? fds_j : fds_k);
    for_each_pencil(im, jm, km, dir, [&](int p1, int p2) {
      std::vector<double> ffb(md + 1 + 2 * hm), df(md + 1), work(2 * (md + 3));
      double* ff = ffb.data() + hm;
      for (int n = 0; n < 4; ++n) {
        const Field& src = n < 3 ? sigma[sel[dir - 1][n]] : qflux[dir - 1];
        gather(src, dir, p1, p2, dim, ff);
        df_compact(s, ff, df.data(), work.data());
        for (int l = 0; l <= dim; ++l) {
          double& r = at(qrhs[n + 1], dir, p1, p2, l);
          r = r + df[l] / h;
        }
      }
    });
  }
}

// tgvsolver.F90:791-883
void MiniApp::filterq() {
  const int md = maxdim(*this);
  for (int dir = 1; dir <= 3; ++dir) {
    bchomovec(q, 5, dir);
    const int dim = dir == 1 ? im : (dir == 2 ? jm : km);
    const CompactScheme& s = dir == 1 ? fil_i : (dir == 2 ? fil_j : fil_k);
    for_each_pencil(im, jm, km, dir, [&](int p1, int p2) {
      std::vector<double> ffb(md + 1 + 2 * hm), fph(md + 1), work(2 * (md + 7));
      double* phi = ffb.data() + hm;
      for (int n = 0; n < 5; ++n) {
        gather(q[n], dir, p1, p2, dim, phi);
        compact_filter(s, fc, phi, fph.data(), work.data());
        for (int l = 0; l <= dim; ++l) at(q[n], dir, p1, p2, l) = fph[l];
      }
    });
  }
}

// tgvsolver.F90:893-942 -- strictly sequential accumulation, k outer / j / i inner
void MiniApp::stacal() {
  double rhom = 0.0, tke = 0.0, enst = 0.0;
  for (int k = 1; k <= km; ++k)
    for (int j = 1; j <= jm; ++j)
      for (int i = 1; i <= im; ++i) {
        rhom = rhom + rho(i, j, k);
        const double u = vel[0](i, j, k), v = vel[1](i, j, k), w = vel[2](i, j, k);
        const double var1 = u * u + v * v + w * w;
        tke = tke + rho(i, j, k) * var1;
        const double o1 = dvel[2][1](i, j, k) - dvel[1][2](i, j, k);
        const double o2 = dvel[0][2](i, j, k) - dvel[2][0](i, j, k);
        const double o3 = dvel[1][0](i, j, k) - dvel[0][1](i, j, k);
        enst = enst + rho(i, j, k) * (o1 * o1 + o2 * o2 + o3 * o3);
      }
  const double cnt = double(im * jm * km);  // im*jm*km is integer arithmetic in the reference
  tke = 0.5 * tke / cnt;
  enst = 0.5 * enst / cnt;
  hist.push_back(double(nstep));
  hist.push_back(time);
  hist.push_back(tke);
  hist.push_back(enst);
}

// tgvsolver.F90:962-992
void MiniApp::rhscal() {
  for (auto& f : qrhs) std::fill(f.v.begin(), f.v.end(), 0.0);
  gradcal();
  convection();
  for (auto& f : qrhs)
    for (double& v : f.v) v = -v;
  diffusion();
}

// tgvsolver.F90:994-1082
void MiniApp::rk3() {
  const double rkcoe[3][3] = {{1.0, 0.0, 1.0}, {0.75, 0.25, 0.25}, {num1d3, num2d3, num2d3}};
  for (int rkstep = 0; rkstep < 3; ++rkstep) {
    bchomo();
    rhscal();
    if (rkstep == 0) {
      stacal();
      for (int m = 0; m < 5; ++m)
        for (int k = 0; k <= km; ++k)
          for (int j = 0; j <= jm; ++j)
            for (int i = 0; i <= im; ++i) qsave[m](i, j, k) = q[m](i, j, k);
    }
    const double c1 = rkcoe[rkstep][0], c2 = rkcoe[rkstep][1], c3 = rkcoe[rkstep][2];
    for (int m = 0; m < 5; ++m) {
#pragma omp parallel for collapse(2) schedule(static)
      for (int k = 0; k <= km; ++k)
        for (int j = 0; j <= jm; ++j)
          for (int i = 0; i <= im; ++i)
            q[m](i, j, k) = c1 * qsave[m](i, j, k) + c2 * q[m](i, j, k) +
                            c3 * qrhs[m](i, j, k) * deltat;
    }
    filterq();
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 0; k <= km; ++k)
      for (int j = 0; j <= jm; ++j)
        for (int i = 0; i <= im; ++i) {  // q2fvar :92-121
          const double r = q[0](i, j, k);
          rho(i, j, k) = r;
          const double u = q[1](i, j, k) / r, v = q[2](i, j, k) / r, w = q[3](i, j, k) / r;
          vel[0](i, j, k) = u; vel[1](i, j, k) = v; vel[2](i, j, k) = w;
          const double p = (q[4](i, j, k) - 0.5 * r * (u * u + v * v + w * w)) / th.const6;
          prs(i, j, k) = p;
          tmp(i, j, k) = p / r * th.const2;
        }
  }
}

}  // namespace astr_oracle

// ---------------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------------
using astr_oracle::MiniApp;
extern "C" {

void* oracle_miniapp_create(int n) {
  auto* m = new MiniApp();
  m->program_init(n);
  m->solver_init();
  m->flowfield_init();
  return m;
}
void oracle_miniapp_destroy(void* h) { delete static_cast<MiniApp*>(h); }

// Advance nsteps RK3 steps (tgvsolver.F90:944-960).  Returns number of history rows.
int oracle_miniapp_run(void* h, int nsteps) {
  auto* m = static_cast<MiniApp*>(h);
  for (int s = 0; s < nsteps; ++s) {
    m->rk3();
    m->nstep += 1;
    m->time = m->time + m->deltat;
  }
  return int(m->hist.size() / 4);
}
// out must hold 4*rows doubles: nstep,time,tke,enst
void oracle_miniapp_history(void* h, double* out) {
  auto* m = static_cast<MiniApp*>(h);
  std::memcpy(out, m->hist.data(), m->hist.size() * sizeof(double));
}
// name: 0..4 q, 5 rho, 6..8 vel, 9 prs, 10 tmp, 11..15 qrhs.  out has the halo'd Fortran shape.
void oracle_miniapp_get(void* h, int id, double* out) {
  auto* m = static_cast<MiniApp*>(h);
  const astr_oracle::Field* f = nullptr;
  if (id < 5) f = &m->q[id];
  else if (id == 5) f = &m->rho;
  else if (id < 9) f = &m->vel[id - 6];
  else if (id == 9) f = &m->prs;
  else if (id == 10) f = &m->tmp;
  else f = &m->qrhs[id - 11];
  std::memcpy(out, f->v.data(), f->v.size() * sizeof(double));
}
// One derivative / filter call on a single pencil, for unit tests.
// f has dim+1+2*hm entries (node -hm first).
void oracle_df_compact(int ntype, int dim, const double* f, double* df) {
  astr_oracle::CompactScheme s;
  astr_oracle::fd_scheme_initiate(s, 643, 'c', ntype, dim);
  std::vector<double> work(2 * s.size());
  astr_oracle::df_compact(s, f + astr_oracle::hm, df, work.data());
}
void oracle_diff6ec(int ntype, int dim, const double* f, double* df) {
  astr_oracle::diff6ec(f + astr_oracle::hm, dim, ntype, df);
}
void oracle_compact_filter(int ntype, int dim, double alfa, double beter_bound, const double* f,
                           double* ff) {
  astr_oracle::CompactScheme s;
  astr_oracle::FilterCoef fc;
  astr_oracle::filter_coefficient_cal(fc, alfa, 1.11, beter_bound);
  astr_oracle::compact_filter_initiate(s, ntype, dim, alfa);
  std::vector<double> work(2 * s.size());
  astr_oracle::compact_filter(s, fc, f + astr_oracle::hm, ff, work.data());
}
// Raw LHS / factorisation tables for tests: out_a,out_c,out_ac1..3 each size n; returns n.
int oracle_scheme_tables(int is_filter, int ntype, int dim, double alfa, int* first_node,
                         double* out_a, double* out_c, double* ac1, double* ac2, double* ac3) {
  astr_oracle::CompactScheme s;
  if (is_filter) astr_oracle::compact_filter_initiate(s, ntype, dim, alfa);
  else astr_oracle::fd_scheme_initiate(s, 643, 'c', ntype, dim);
  const int n = s.size();
  *first_node = s.first_node;
  if (out_a) {
    for (int i = 0; i < n; ++i) {
      out_a[i] = s.a[i]; out_c[i] = s.c[i];
      ac1[i] = s.ac1[i]; ac2[i] = s.ac2[i]; ac3[i] = s.ac3[i];
    }
  }
  return n;
}
int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// bench.py: launchers such as torch.distributed.run export OMP_NUM_THREADS=1; the CPU baseline sets its
// thread count explicitly and reports what it got
int oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
}
