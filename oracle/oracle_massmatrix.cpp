/*
 * oracle_massmatrix.cpp -- TEST INFRASTRUCTURE ONLY (see picnic_oracle.h).
 *
 * CPU restatement of the reference's per-particle mass-matrix deposit:
 *   PicChargedSpecies::accumulateMassMatrices      src/species/pic/charged/PicChargedSpecies.cpp:3671-3761
 *   MeshInterp::depositMassMatrices                src/particle_tools/MeshInterpI.H:231-475
 *   cc1_1d_deposit_mass_matrix                     src/particle_tools/MeshInterpMassMatrixF.ChF:835-1220
 *   cc1_2d_deposit_mass_matrix                     src/particle_tools/MeshInterpMassMatrixF.ChF:1228-1862
 *   compute_mm_kernals (planar push, inert_type 0) src/particle_tools/MeshInterpMassMatrixF.ChF:1869-2076
 * of the component counts of the sigma containers
 *   PicSpeciesInterface::initializeMassMatrices    src/species/pic/PicSpeciesInterface.cpp:256-350
 * and of the grid contraction J = J0 + sigma (E - E0)
 *   compute_J{x,y,z}_from_mass_matrix              src/fields/FieldsF.ChF:3-415
 *   PicSpeciesInterface::computeJfromMassMatrices  src/species/pic/PicSpeciesInterface.cpp:567-753
 *
 * PARITY STATUS: "parity unpinned" (no golden vectors in the reference, Fortran not buildable here).  Pinned by
 * the reference-derived identity tested in tests/test_oracle_massmatrix.py: with the particle orbits frozen,
 * the CC1 current deposit of Boris(u_old, E_p(E), B_p) equals J0 + sigma (E - E0) to round-off for any E,
 * which exercises every weight, every component index Nc and the contraction offsets at once.
 *
 * Fortran literals without a d0 suffix are taken as double (Chombo builds with -fdefault-real-8); the only
 * place where that matters is the 1.01 threshold of the relativistic correction.
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "picnic_oracle.h"

namespace {

struct View {
  double *p;
  int lo0, lo1, n0, n1;
  explicit View(const orc_fab &f)
      : p(f.p), lo0(f.lo[0]), lo1(f.lo[1]), n0(f.hi[0] - f.lo[0] + 1), n1(f.hi[1] - f.lo[1] + 1) {}
  inline bool in(int i, int j) const {
    return i >= lo0 && i < lo0 + n0 && j >= lo1 && j < lo1 + n1;
  }
  inline double &operator()(int i, int j) const { return p[(i - lo0) + (long)(j - lo1) * n0]; }
};

/* CHF_FRA: components are the slowest index */
struct MView {
  double *p;
  int lo0, lo1, n0, n1, ncomp;
  explicit MView(const orc_mfab &f)
      : p(f.p), lo0(f.lo[0]), lo1(f.lo[1]), n0(f.hi[0] - f.lo[0] + 1), n1(f.hi[1] - f.lo[1] + 1), ncomp(f.ncomp) {}
  inline bool in(int i, int j, int c) const {
    return i >= lo0 && i < lo0 + n0 && j >= lo1 && j < lo1 + n1 && c >= 0 && c < ncomp;
  }
  inline double &operator()(int i, int j, int c) const {
    return p[(i - lo0) + (long)(j - lo1) * n0 + (long)c * n0 * n1];
  }
};

inline int ifloor(double a) { return (int)std::floor(a); }

struct Kern {
  double fp[3];      /* fpx, fpy, fpz */
  double f[3][3];    /* fpxx ... fpzz */
};

/* compute_mm_kernals, planar branch (MeshInterpMassMatrixF.ChF:1941-2074).  Bp comes in as gathered and is
 * scaled in place like the Fortran does. */
void mm_kernels(Kern &k, double *Bp, double qp, double alphas, double volume, const double *upold,
                const double *upbar, int anticyclic, int relativistic) {
  double gammap_bar = 1.0, gammap_new = 1.0, gammap_tilde = 1.0;
  double upnew[3] = {0.0, 0.0, 0.0};
  double rhop;
  if (relativistic) {
    for (int c = 0; c < 3; ++c) upnew[c] = 2.0 * upbar[c] - upold[c];
    gammap_bar = std::sqrt(1.0 + upbar[0] * upbar[0] + upbar[1] * upbar[1] + upbar[2] * upbar[2]);
    const double gammap_old = std::sqrt(1.0 + upold[0] * upold[0] + upold[1] * upold[1] + upold[2] * upold[2]);
    gammap_new = std::sqrt(1.0 + upnew[0] * upnew[0] + upnew[1] * upnew[1] + upnew[2] * upnew[2]);
    gammap_tilde = 0.5 * (gammap_old + gammap_new);
    rhop = qp / volume / gammap_tilde;
    for (int c = 0; c < 3; ++c) Bp[c] = alphas * Bp[c] / gammap_bar;
  } else {
    rhop = qp / volume;
    for (int c = 0; c < 3; ++c) Bp[c] = alphas * Bp[c];
  }
  const double Bpx = Bp[0], Bpy = Bp[1], Bpz = Bp[2];
  const double ac = (double)anticyclic;
  const double Bpsq = Bpx * Bpx + Bpy * Bpy + Bpz * Bpz;
  const double arogp = alphas * rhop / (1.0 + Bpsq);
  k.f[0][0] = arogp * (Bpx * Bpx + 1.0);
  k.f[0][1] = arogp * (Bpx * Bpy + ac * Bpz);
  k.f[0][2] = arogp * (Bpx * Bpz - ac * Bpy);
  k.f[1][0] = arogp * (Bpy * Bpx - ac * Bpz);
  k.f[1][1] = arogp * (Bpy * Bpy + 1.0);
  k.f[1][2] = arogp * (Bpy * Bpz + ac * Bpx);
  k.f[2][0] = arogp * (Bpz * Bpx + ac * Bpy);
  k.f[2][1] = arogp * (Bpz * Bpy - ac * Bpx);
  k.f[2][2] = arogp * (Bpz * Bpz + 1.0);
  if (relativistic && gammap_bar > 1.01) {
    /* relativistic effects in the Lorentz force (:2016-2038) */
    const double upBp = upbar[0] * Bpx + upbar[1] * Bpy + upbar[2] * Bpz;
    double gp_denom = gammap_bar * gammap_bar + Bpsq + upBp * upBp;
    double gp[3];
    gp[0] = (Bpsq * upbar[0] - upBp * Bpx - ac * (upbar[1] * Bpz - upbar[2] * Bpy)) / gp_denom;
    gp[1] = (Bpsq * upbar[1] - upBp * Bpy - ac * (upbar[2] * Bpx - upbar[0] * Bpz)) / gp_denom;
    gp[2] = (Bpsq * upbar[2] - upBp * Bpz - ac * (upbar[0] * Bpy - upbar[1] * Bpx)) / gp_denom;
    double upf[3];
    for (int e = 0; e < 3; ++e) upf[e] = upbar[0] * k.f[0][e] + upbar[1] * k.f[1][e] + upbar[2] * k.f[2][e];
    for (int j = 0; j < 3; ++j)
      for (int e = 0; e < 3; ++e) k.f[j][e] = k.f[j][e] + gp[j] * upf[e];
    /* relativistic effects in vp = up/gammap_tilde (:2040-2060) */
    gp_denom = gammap_tilde * gammap_new;
    for (int j = 0; j < 3; ++j) gp[j] = -upbar[j] / gp_denom;
    for (int e = 0; e < 3; ++e) upf[e] = upnew[0] * k.f[0][e] + upnew[1] * k.f[1][e] + upnew[2] * k.f[2][e];
    for (int j = 0; j < 3; ++j)
      for (int e = 0; e < 3; ++e) k.f[j][e] = k.f[j][e] + gp[j] * upf[e];
  }
  /* J0 kernels (:2071-2073) */
  for (int c = 0; c < 3; ++c) k.fp[c] = rhop * upbar[c];
}

struct Target {
  View J[3];
  MView s[9];   /* xx xy xz yx yy yz zx zy zz */
  int err;
  inline void addJ(int c, int i, int j, double v) {
    if (!J[c].in(i, j)) { err = 1; return; }
    J[c](i, j) = J[c](i, j) + v;
  }
  inline void addS(int k, int i, int j, int nc, double v) {
    if (!s[k].in(i, j, nc)) { err = 1; return; }
    s[k](i, j, nc) = s[k](i, j, nc) + v;
  }
};
enum { XX = 0, XY, XZ, YX, YY, YZ, ZX, ZY, ZZ };

/* ---- cc1_1d_deposit_mass_matrix (:835-1220) ------------------------------------------------ */
int mm_cc1_1d(const orc_geom &g, Target &T, const View *B, const double *upold, const double *upbar,
              double alphas, int anticyclic, double qp, double xpold, double xpbar, int relativistic) {
  const double dx = g.dx[0], le = g.le[0], re = g.re[0];
  const int index = ifloor((xpbar - le - 0.5 * dx) / dx);
  const int index_stag = ifloor((xpbar - le) / dx);

  /* magnetic field at the particle (:899-943) */
  double Bp[3] = {0.0, 0.0, 0.0};
  for (int ii = index; ii <= index + 1; ++ii) {
    const double l0 = ii * dx + 0.5 * dx - xpbar + le;
    const int ii_stag = ii - index + index_stag;
    const double l0_stag = ii_stag * dx - xpbar + le;
    const double w0 = 1.0 - std::fabs(l0 / dx);
    const double w0_stag = 1.0 - std::fabs(l0_stag / dx);
    if (!B[0].in(ii_stag, 0) || !B[1].in(ii, 0) || !B[2].in(ii, 0)) { T.err = 1; return 0; }
    Bp[0] = Bp[0] + w0_stag * B[0](ii_stag, 0);
    Bp[1] = Bp[1] + w0 * B[1](ii, 0);
    Bp[2] = Bp[2] + w0 * B[2](ii, 0);
  }
  Kern k;
  mm_kernels(k, Bp, qp, alphas, dx, upold, upbar, anticyclic, relativistic);

  /* y and z rows against Ey, Ez (:965-1009) */
  for (int ii_stag = index_stag; ii_stag <= index_stag + 1; ++ii_stag) {
    const double l0_stag = ii_stag * dx - xpbar + le;
    const double w0_stag = 1.0 - std::fabs(l0_stag / dx);
    const int off_diag_comp = (ii_stag == index_stag) ? 2 : 0;
    const double weight = w0_stag;
    T.addS(YY, ii_stag, 0, 1, k.f[1][1] * weight * weight);
    T.addS(YY, ii_stag, 0, off_diag_comp, k.f[1][1] * weight * (1.0 - weight));
    T.addS(YZ, ii_stag, 0, 1, k.f[1][2] * weight * weight);
    T.addS(YZ, ii_stag, 0, off_diag_comp, k.f[1][2] * weight * (1.0 - weight));
    T.addJ(1, ii_stag, 0, k.fp[1] * weight);
    T.addS(ZY, ii_stag, 0, 1, k.f[2][1] * weight * weight);
    T.addS(ZY, ii_stag, 0, off_diag_comp, k.f[2][1] * weight * (1.0 - weight));
    T.addS(ZZ, ii_stag, 0, 1, k.f[2][2] * weight * weight);
    T.addS(ZZ, ii_stag, 0, off_diag_comp, k.f[2][2] * weight * (1.0 - weight));
    T.addJ(2, ii_stag, 0, k.fp[2] * weight);
  }

  /* cell crossings (:1017-1090) */
  double xpnew = 2.0 * xpbar - xpold;
  const double dXp = std::fabs(xpnew - xpold);
  double bc_seg_factor = 1.0;
  double xpold0 = xpold;
  const int shift = (index == index_stag) ? 0 : 1;
  if (g.bc_lo[0] == 1) {
    if (xpold0 < le) {
      xpold0 = le;
      bc_seg_factor = std::fabs(xpnew - xpold0) / dXp;
    }
    if (xpnew < le) {
      xpnew = le;
      bc_seg_factor = std::fabs(xpnew - xpold0) / dXp;
    }
  }
  if (g.bc_hi[0] == 1) {
    if (xpold0 > re) {
      xpold0 = re;
      bc_seg_factor = std::fabs(xpnew - xpold0) / dXp;
    }
    if (xpnew > re) {
      xpnew = re;
      bc_seg_factor = std::fabs(xpnew - xpold0) / dXp;
    }
  }
  const double l0_stag = xpbar - index_stag * dx - le;
  const double wx_up_stag = l0_stag / dx;
  const double wx_dn_stag = 1.0 - wx_up_stag;
  const double l0 = xpbar - (index + 0.5) * dx - le;
  const double wx_up = l0 / dx;
  const double wx_dn = 1.0 - wx_up;
  const int index_old = ifloor((xpold0 - le - 0.5 * dx) / dx);
  const int index_new = ifloor((xpnew - le - 0.5 * dx) / dx);
  const int num_segments = 1 + std::abs(index_new - index_old);
  const int index_min = index_old < index_new ? index_old : index_new;
  const int maxXings = (T.s[XY].ncomp - 2) / 2;
  if (num_segments > maxXings + 1) return -1;
  /* the Fortran's weight vectors hold three segments (:884-889) */
  if (num_segments > 3) return -1;

  int SegNumX[3] = {1, 0, 0};
  double dn[3] = {0.0, wx_dn * bc_seg_factor, 0.0};
  double up[3] = {0.0, wx_up * bc_seg_factor, 0.0};
  const double xmin = xpold0 < xpnew ? xpold0 : xpnew;
  const double xmax = xpold0 > xpnew ? xpold0 : xpnew;
  if (num_segments == 2) {
    if (index_min < index) {
      const double Xcell = le + (index + 0.5) * dx;
      double dXp_sub = Xcell - xmin;
      dn[0] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      up[0] = dXp_sub / dXp - dn[0];
      dXp_sub = xmax - Xcell;
      up[1] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      dn[1] = dXp_sub / dXp - up[1];
      SegNumX[0] = 0;
      SegNumX[1] = 1;
    } else {
      const double Xcell = le + (index + 1.5) * dx;
      double dXp_sub = Xcell - xmin;
      dn[1] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      up[1] = dXp_sub / dXp - dn[1];
      dXp_sub = xmax - Xcell;
      up[2] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      dn[2] = dXp_sub / dXp - up[2];
      SegNumX[0] = 1;
      SegNumX[1] = 2;
    }
  }
  if (num_segments == 3) {
    double Xcell = le + (index + 0.5) * dx;
    double dXp_sub = Xcell - xmin;
    dn[0] = dXp_sub * dXp_sub / 2.0 / dXp / dx;
    up[0] = dXp_sub / dXp - dn[0];
    up[1] = dx / dXp / 2.0;
    dn[1] = up[1];
    Xcell = Xcell + dx;
    dXp_sub = xmax - Xcell;
    up[2] = dXp_sub * dXp_sub / 2.0 / dXp / dx;
    dn[2] = dXp_sub / dXp - up[2];
    SegNumX[0] = 0;
    SegNumX[1] = 1;
    SegNumX[2] = 2;
  }

  /* x row (:1153-1195) */
  for (int nJ = 0; nJ < num_segments; ++nJ) {
    const int llJ = SegNumX[nJ];
    const int ii = index - 1 + llJ;
    const double w0_cic[2] = {dn[llJ], up[llJ]};
    for (int iiJ = 0; iiJ < 2; ++iiJ) {
      const double weight_J = w0_cic[iiJ];
      T.addJ(0, ii + iiJ, 0, k.fp[0] * weight_J);
      for (int nE = 0; nE < num_segments; ++nE) {
        const int llE = SegNumX[nE];
        const double w0E_cic[2] = {dn[llE], up[llE]};
        for (int iiE = 0; iiE < 2; ++iiE) {
          const double weight_E = w0E_cic[iiE];
          const int Nc = 1 + maxXings + iiE - iiJ + llE - llJ;
          T.addS(XX, ii + iiJ, 0, Nc, k.f[0][0] * weight_J * weight_E);
        }
      }
      for (int iiE = 0; iiE < 2; ++iiE) {
        const double weight_E = (1 - iiE) * wx_dn_stag + iiE * wx_up_stag;
        const int Nc = 1 + maxXings + shift + iiE - iiJ - llJ;
        T.addS(XY, ii + iiJ, 0, Nc, k.f[0][1] * weight_J * weight_E);
        T.addS(XZ, ii + iiJ, 0, Nc, k.f[0][2] * weight_J * weight_E);
      }
    }
  }
  /* y and z rows against Ex (:1200-1220) */
  for (int iiJ = 0; iiJ < 2; ++iiJ) {
    const double weight_J = (1 - iiJ) * wx_dn_stag + iiJ * wx_up_stag;
    for (int nE = 0; nE < num_segments; ++nE) {
      const int llE = SegNumX[nE];
      const double w0E_cic[2] = {dn[llE], up[llE]};
      for (int iiE = 0; iiE < 2; ++iiE) {
        const double weight_E = w0E_cic[iiE];
        const int Nc = maxXings - shift + iiE - iiJ + llE;
        T.addS(YX, index_stag + iiJ, 0, Nc, k.f[1][0] * weight_J * weight_E);
        T.addS(ZX, index_stag + iiJ, 0, Nc, k.f[2][0] * weight_J * weight_E);
      }
    }
  }
  return 0;
}

/* truncate_boundaries (MeshInterpChargeConservingF.ChF:2021-2084), restated in oracle_interp.cpp */
}  // namespace
extern "C" void orc_truncate_boundaries_2d(const orc_geom *g, double *xpold, double *xpnew, double slope,
                                           double slope_inv);
namespace {

/* ---- cc1_2d_deposit_mass_matrix (:1228-1862) ------------------------------------------------ */
enum { MAXSEG = 8 };
int mm_cc1_2d(const orc_geom &g, Target &T, const View *B, const double *upold, const double *upbar,
              double alphas, int anticyclic, double qp, const double *xpold_in, const double *xpbar,
              int relativistic) {
  const int i0 = 0, i1 = 1;
  const double *dx = g.dx, *le = g.le;
  int index[2], index_stag[2];
  double wv[2][2], wsv[2][2];   /* w{d}_vec, w{d}_stag_vec */
  for (int d = 0; d < 2; ++d) {
    index[d] = ifloor((xpbar[d] - le[d] - 0.5 * dx[d]) / dx[d]);
    index_stag[d] = ifloor((xpbar[d] - le[d]) / dx[d]);
    const double l = xpbar[d] - ((index[d] + 0.5) * dx[d] + le[d]);
    wv[d][1] = l / dx[d];
    wv[d][0] = 1.0 - wv[d][1];
    const double ls = xpbar[d] - (index_stag[d] * dx[d] + le[d]);
    wsv[d][1] = ls / dx[d];
    wsv[d][0] = 1.0 - wsv[d][1];
  }
  /* magnetic field at the particle (:1326-1360) */
  double Bp[3] = {0.0, 0.0, 0.0};
  for (int iiJ = 0; iiJ < 2; ++iiJ) {
    const int ii = index[0] + iiJ, ii_stag = index_stag[0] + iiJ;
    const double w0 = wv[0][iiJ], w0_stag = wsv[0][iiJ];
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const int jj = index[1] + jjJ, jj_stag = index_stag[1] + jjJ;
      const double w1 = wv[1][jjJ], w1_stag = wsv[1][jjJ];
      if (!B[0].in(ii_stag, jj) || !B[1].in(ii, jj_stag) || !B[2].in(ii, jj)) { T.err = 1; return 0; }
      double weight = w0_stag * w1;
      Bp[0] = Bp[0] + weight * B[0](ii_stag, jj);
      weight = w0 * w1_stag;
      Bp[1] = Bp[1] + weight * B[1](ii, jj_stag);
      weight = w0 * w1;
      Bp[2] = Bp[2] + weight * B[2](ii, jj);
    }
  }
  Kern k;
  mm_kernels(k, Bp, qp, alphas, dx[0] * dx[1], upold, upbar, anticyclic, relativistic);

  /* Jz and sigma_zz (:1375-1411) */
  for (int iiJ = 0; iiJ < 2; ++iiJ) {
    const int ii_stag = index_stag[0] + iiJ;
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const int jj_stag = index_stag[1] + jjJ;
      const double weight_J = wsv[0][iiJ] * wsv[1][jjJ];
      T.addJ(2, ii_stag, jj_stag, k.fp[2] * weight_J);
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE) {
          const double weight_E = wsv[0][iiE] * wsv[1][jjE];
          const int Nc = 1 + iiE - iiJ + 3 * (1 + jjE - jjJ);
          T.addS(ZZ, ii_stag, jj_stag, Nc, k.f[2][2] * weight_J * weight_E);
        }
    }
  }

  /* cell crossings (:1417-1470) */
  const int maxXings = (int)((std::sqrt(1.0 * T.s[XY].ncomp) - 4) / 2);
  double xpold[2] = {xpold_in[0], xpold_in[1]};
  double xpnew[2], dXp[2];
  int num_segments = 1;
  for (int d = 0; d < 2; ++d) {
    xpnew[d] = 2.0 * xpbar[d] - xpold[d];
    dXp[d] = xpnew[d] - xpold[d];
  }
  const double slope = dXp[i1] / dXp[i0];
  const double slope_inv = 1.0 / slope;
  orc_truncate_boundaries_2d(&g, xpold, xpnew, slope, slope_inv);
  int index_old[2], index_new[2], sign[2], cell_crossings[2], shift[2];
  for (int d = 0; d < 2; ++d) {
    index_old[d] = ifloor((xpold[d] - le[d] - 0.5 * dx[d]) / dx[d]);
    index_new[d] = ifloor((xpnew[d] - le[d] - 0.5 * dx[d]) / dx[d]);
    sign[d] = (index_new[d] < index_old[d]) ? -1 : 1;
    cell_crossings[d] = std::abs(index_new[d] - index_old[d]);
    num_segments = num_segments + cell_crossings[d];
    if (cell_crossings[d] > maxXings) return -1;
  }
  if (num_segments > MAXSEG) return -1;   /* the Fortran's arrays hold 5 (:1287-1296) */
  for (int d = 0; d < 2; ++d) shift[d] = (index[d] == index_stag[d]) ? 0 : 1;
  double Xcell[2];
  for (int d = 0; d < 2; ++d) Xcell[d] = le[d] + (index_old[d] + 0.5 * (1 - sign[d]) + 0.5) * dx[d];

  /* pre-define all interpolation weights (:1476-1597) */
  int SegNumX[MAXSEG], SegNumY[MAXSEG];
  double cicX[MAXSEG][2], cicY[MAXSEG][2], tscX[MAXSEG][3], tscY[MAXSEG][3];
  double xpold0[2] = {xpold[0], xpold[1]}, xpnew0[2] = {0.0, 0.0}, dXp_sub[2] = {0.0, 0.0};
  int ii_next = index_old[0], jj_next = index_old[1];
  for (int nn = 0; nn < num_segments; ++nn) {
    const int ii = ii_next, jj = jj_next;
    if (nn == num_segments - 1) {
      xpnew0[0] = xpnew[0];
      xpnew0[1] = xpnew[1];
      dXp_sub[0] = xpnew0[0] - xpold0[0];
      dXp_sub[1] = xpnew0[1] - xpold0[1];
    } else if (cell_crossings[i0] == 0) {
      jj_next = jj + sign[i1];
      Xcell[i1] = Xcell[i1] + sign[i1] * dx[i1];
      xpnew0[i1] = Xcell[i1];
      dXp_sub[i1] = xpnew0[i1] - xpold0[i1];
      dXp_sub[i0] = slope_inv * dXp_sub[i1];
      xpnew0[i0] = xpold0[i0] + dXp_sub[i0];
    } else if (cell_crossings[i1] == 0) {
      ii_next = ii + sign[i0];
      Xcell[i0] = Xcell[i0] + sign[i0] * dx[i0];
      xpnew0[i0] = Xcell[i0];
      dXp_sub[i0] = xpnew0[i0] - xpold0[i0];
      dXp_sub[i1] = slope * dXp_sub[i0];
      xpnew0[i1] = xpold0[i1] + dXp_sub[i1];
    } else {
      xpnew0[i0] = Xcell[i0] + sign[i0] * dx[i0];
      xpnew0[i1] = Xcell[i1] + sign[i1] * dx[i1];
      dXp_sub[i0] = xpnew0[i0] - xpold0[i0];
      dXp_sub[i1] = xpnew0[i1] - xpold0[i1];
      const double dXp_sub02 = slope_inv * dXp_sub[i1];
      if (std::fabs(dXp_sub[i0]) < std::fabs(dXp_sub02)) {
        dXp_sub[i1] = slope * dXp_sub[i0];
        xpnew0[i1] = xpold0[i1] + dXp_sub[i1];
        Xcell[i0] = xpnew0[i0];
        ii_next = ii + sign[i0];
        cell_crossings[i0] = cell_crossings[i0] - 1;
      } else {
        dXp_sub[i0] = slope_inv * dXp_sub[i1];
        xpnew0[i0] = xpold0[i0] + dXp_sub[i0];
        Xcell[i1] = xpnew0[i1];
        jj_next = jj + sign[i1];
        cell_crossings[i1] = cell_crossings[i1] - 1;
      }
    }
    double seg_factor[2];
    for (int d = 0; d < 2; ++d) {
      if (dXp[d] != 0.0) seg_factor[d] = dXp_sub[d] / dXp[d];
      else seg_factor[d] = 1.0;
    }
    double xpbar0[2];
    int index_start[2];
    for (int d = 0; d < 2; ++d) {
      xpbar0[d] = 0.5 * (xpold0[d] + xpnew0[d]);
      index_start[d] = ifloor((xpbar0[d] - le[d] - 0.5 * dx[d]) / dx[d]);
    }
    SegNumX[nn] = 1 + index_start[i0] - index[i0];
    SegNumY[nn] = 1 + index_start[i1] - index[i1];
    const double delta0 = (xpbar0[i0] - (le[i0] + (ii + 0.5) * dx[i0])) / dx[i0];
    const double delta1 = (xpbar0[i1] - (le[i1] + (jj + 0.5) * dx[i1])) / dx[i1];
    cicX[nn][0] = (1.0 - delta0) * seg_factor[i0];
    cicX[nn][1] = delta0 * seg_factor[i0];
    cicY[nn][0] = (1.0 - delta1) * seg_factor[i1];
    cicY[nn][1] = delta1 * seg_factor[i1];
    /* TSC weights averaged over the end points of the segment (:1546-1592) */
    for (int d = 0; d < 2; ++d) {
      double(*tsc)[3] = (d == 0) ? tscX : tscY;
      for (int b = 0; b < 3; ++b) {
        double l = (index_start[d] + b) * dx[d] - xpold0[d] + le[d];
        double delta = std::fabs(l / dx[d]);
        double t = 1.5 - delta;
        const double w_old = (b == 1) ? 0.75 - delta * delta : 0.5 * (t * t);
        l = (index_start[d] + b) * dx[d] - xpnew0[d] + le[d];
        delta = std::fabs(l / dx[d]);
        t = 1.5 - delta;
        const double w_new = (b == 1) ? 0.75 - delta * delta : 0.5 * (t * t);
        tsc[nn][b] = 0.5 * (w_old + w_new);
      }
    }
    xpold0[0] = xpnew0[0];
    xpold0[1] = xpnew0[1];
  }

  const int mX = maxXings;
  /* loop over segments and deposit (:1603-1793) */
  for (int nJ = 0; nJ < num_segments; ++nJ) {
    const int llJ = SegNumX[nJ], mmJ = SegNumY[nJ];
    const int ii = index[i0] - 1 + llJ, jj = index[i1] - 1 + mmJ;
    const int ii_stag = ii, jj_stag = jj;
    /* Jx rows */
    for (int iiJ = 0; iiJ < 2; ++iiJ)
      for (int jjJ = 0; jjJ < 3; ++jjJ) {
        const double weight_J = cicX[nJ][iiJ] * tscY[nJ][jjJ];
        T.addJ(0, ii + iiJ, jj_stag + jjJ, k.fp[0] * weight_J);
        for (int nE = 0; nE < num_segments; ++nE) {
          const int llE = SegNumX[nE], mmE = SegNumY[nE];
          for (int iiE = 0; iiE < 2; ++iiE)
            for (int jjE = 0; jjE < 3; ++jjE) {
              const int Nc = 1 + mX + llE - llJ + iiE - iiJ + (3 + 2 * mX) * (2 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = cicX[nE][iiE] * tscY[nE][jjE];
              T.addS(XX, ii + iiJ, jj_stag + jjJ, Nc, k.f[0][0] * weight_J * weight_E);
            }
          for (int iiE = 0; iiE < 3; ++iiE)
            for (int jjE = 0; jjE < 2; ++jjE) {
              const int Nc = 1 + mX + llE - llJ + iiE - iiJ + (4 + 2 * mX) * (2 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = tscX[nE][iiE] * cicY[nE][jjE];
              T.addS(XY, ii + iiJ, jj_stag + jjJ, Nc, k.f[0][1] * weight_J * weight_E);
            }
        }
        for (int iiE = 0; iiE < 2; ++iiE)
          for (int jjE = 0; jjE < 2; ++jjE) {
            const int Nc = 1 + mX + shift[i0] - llJ + iiE - iiJ + (2 + 2 * mX) * (2 + mX + shift[i1] - mmJ + jjE - jjJ);
            const double weight_E = wsv[0][iiE] * wsv[1][jjE];
            T.addS(XZ, ii + iiJ, jj_stag + jjJ, Nc, k.f[0][2] * weight_J * weight_E);
          }
      }
    /* Jy rows */
    for (int iiJ = 0; iiJ < 3; ++iiJ)
      for (int jjJ = 0; jjJ < 2; ++jjJ) {
        const double weight_J = tscX[nJ][iiJ] * cicY[nJ][jjJ];
        T.addJ(1, ii_stag + iiJ, jj + jjJ, k.fp[1] * weight_J);
        for (int nE = 0; nE < num_segments; ++nE) {
          const int llE = SegNumX[nE], mmE = SegNumY[nE];
          for (int iiE = 0; iiE < 2; ++iiE)
            for (int jjE = 0; jjE < 3; ++jjE) {
              const int Nc = 2 + mX + llE - llJ + iiE - iiJ + (4 + 2 * mX) * (1 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = cicX[nE][iiE] * tscY[nE][jjE];
              T.addS(YX, ii_stag + iiJ, jj + jjJ, Nc, k.f[1][0] * weight_J * weight_E);
            }
          for (int iiE = 0; iiE < 3; ++iiE)
            for (int jjE = 0; jjE < 2; ++jjE) {
              const int Nc = 2 + mX + llE - llJ + iiE - iiJ + (5 + 2 * mX) * (1 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = tscX[nE][iiE] * cicY[nE][jjE];
              T.addS(YY, ii_stag + iiJ, jj + jjJ, Nc, k.f[1][1] * weight_J * weight_E);
            }
        }
        for (int iiE = 0; iiE < 2; ++iiE)
          for (int jjE = 0; jjE < 2; ++jjE) {
            const int Nc = 2 + mX + shift[i0] - llJ + iiE - iiJ + (3 + 2 * mX) * (1 + mX + shift[i1] - mmJ + jjE - jjJ);
            const double weight_E = wsv[0][iiE] * wsv[1][jjE];
            T.addS(YZ, ii_stag + iiJ, jj + jjJ, Nc, k.f[1][2] * weight_J * weight_E);
          }
      }
  }
  /* Jz rows against Ex, Ey (:1798-1858) */
  for (int iiJ = 0; iiJ < 2; ++iiJ)
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const double weight_J = wsv[0][iiJ] * wsv[1][jjJ];
      for (int nE = 0; nE < num_segments; ++nE) {
        const int llE = SegNumX[nE], mmE = SegNumY[nE];
        for (int iiE = 0; iiE < 2; ++iiE)
          for (int jjE = 0; jjE < 3; ++jjE) {
            const int Nc = mX - shift[i0] + llE + iiE - iiJ + (2 + 2 * mX) * (mX - shift[i1] + mmE + jjE - jjJ);
            const double weight_E = cicX[nE][iiE] * tscY[nE][jjE];
            T.addS(ZX, index_stag[i0] + iiJ, index_stag[i1] + jjJ, Nc, k.f[2][0] * weight_J * weight_E);
          }
        for (int iiE = 0; iiE < 3; ++iiE)
          for (int jjE = 0; jjE < 2; ++jjE) {
            const int Nc = mX - shift[i0] + llE + iiE - iiJ + (3 + 2 * mX) * (mX - shift[i1] + mmE + jjE - jjJ);
            const double weight_E = tscX[nE][iiE] * cicY[nE][jjE];
            T.addS(ZY, index_stag[i0] + iiJ, index_stag[i1] + jjJ, Nc, k.f[2][1] * weight_J * weight_E);
          }
      }
    }
  return 0;
}

}  // namespace

/* PicSpeciesInterface::initializeMassMatrices (PicSpeciesInterface.cpp:256-350): components per direction of
 * the nine sigma containers, order xx xy xz yx yy yz zx zy zz.  Returns -1 for a combination the reference
 * asserts against. */
extern "C" int orc_mm_ncomp(int D, int interp, int ghosts, int *ncomp /* [9][2] */) {
  const int tsc = (interp == ORC_TSC) ? 2 : 0;
  const int d0[9] = {3, 4, 4, 4, 3, 3, 4, 3, 3};
  const int d1[9] = {3, 4, 3, 4, 3, 4, 3, 4, 3};
  for (int k = 0; k < 9; ++k) {
    ncomp[2 * k] = d0[k] + tsc;
    ncomp[2 * k + 1] = (D >= 2) ? d1[k] + tsc : 1;
  }
  if (interp == ORC_TSC && ghosts < 3) return -1;
  if (interp == ORC_CIC && ghosts < 2) return -1;
  if (interp == ORC_CC0) {
    if (D != 1 || ghosts < 2) return -1;
    ncomp[0] = 5;
  }
  if (interp == ORC_CC1 && D == 1) {
    if (ghosts < 2) return -1;
    const int m = ghosts - 1;
    ncomp[2 * 0] = 3 + 2 * m;
    ncomp[2 * 1] = 2 + 2 * m;
    ncomp[2 * 2] = 2 + 2 * m;
    ncomp[2 * 3] = 2 + 2 * m;
    ncomp[2 * 6] = 2 + 2 * m;
  }
  if (interp == ORC_CC1 && D == 2) {
    if (ghosts < 3) return -1;
    const int m = ghosts - 2;
    const int c0[9] = {3, 4, 2, 4, 5, 3, 2, 3, 3};
    const int c1[9] = {5, 4, 3, 4, 3, 2, 3, 2, 3};
    for (int k = 0; k < 9; ++k) {
      ncomp[2 * k] = (k == 8) ? 3 : c0[k] + 2 * m;
      ncomp[2 * k + 1] = (k == 8) ? 3 : c1[k] + 2 * m;
    }
  }
  return 0;
}

/* test hook: compute_mm_kernals for one particle; out = fpx,fpy,fpz, fpxx..fpzz (12 doubles) */
extern "C" void orc_mm_kernels(const double *Bp_in, double qp, double alphas, double volume, const double *upold,
                               const double *upbar, int anticyclic, int relativistic, double *out) {
  Kern k;
  double Bp[3] = {Bp_in[0], Bp_in[1], Bp_in[2]};
  mm_kernels(k, Bp, qp, alphas, volume, upold, upbar, anticyclic, relativistic);
  for (int c = 0; c < 3; ++c) out[c] = k.fp[c];
  for (int j = 0; j < 3; ++j)
    for (int e = 0; e < 3; ++e) out[3 + 3 * j + e] = k.f[j][e];
}

/* accumulateMassMatrices for one species on one box (CC1, planar push).  Accumulates into J0[3] and sigma[9]
 * (the caller zeroes them, PicSpeciesInterface::setMassMatrices :1044-1058).  qovs = charge/volume_scale,
 * alphas = fnorm*cnormDt/2.  Returns 0, -1 (too many crossings: Fortran STOP), -2 (unsupported), -3 (a
 * stencil left the arrays: undefined behaviour in the reference). */
extern "C" int orc_deposit_mass_matrices(const orc_geom *gp, int interp, long n, const double *x, const double *xold,
                                         const double *v, const double *vold, const double *w, double qovs,
                                         double alphas, double cnormDt, int anticyclic, int relativistic,
                                         const orc_fab *B, orc_fab *J0, orc_mfab *sigma) {
  (void)cnormDt;   /* only used by the NEW_EXACT_CHARGE_CONSERVATION build */
  const orc_geom &g = *gp;
  if (interp != ORC_CC1) return -2;
  const View Bv[3] = {View(B[0]), View(B[1]), View(B[2])};
  Target T{{View(J0[0]), View(J0[1]), View(J0[2])},
           {MView(sigma[0]), MView(sigma[1]), MView(sigma[2]), MView(sigma[3]), MView(sigma[4]), MView(sigma[5]),
            MView(sigma[6]), MView(sigma[7]), MView(sigma[8])},
           0};
  const int ac = anticyclic ? -1 : 1;
  int rc = 0;
  for (long p = 0; p < n; ++p) {
    const double upold[3] = {vold[p], vold[n + p], vold[2 * n + p]};
    const double upbar[3] = {v[p], v[n + p], v[2 * n + p]};
    const double wp = w[p] * qovs;
    int r;
    if (g.D == 1) {
      r = mm_cc1_1d(g, T, Bv, upold, upbar, alphas, ac, wp, xold[p], x[p], relativistic);
    } else {
      const double xo[2] = {xold[p], xold[n + p]};
      const double xb[2] = {x[p], x[n + p]};
      r = mm_cc1_2d(g, T, Bv, upold, upbar, alphas, ac, wp, xo, xb, relativistic);
    }
    if (r) rc = r;
  }
  if (rc == 0 && T.err) rc = -3;
  return rc;
}

/* compute_J{x,y,z}_from_mass_matrix (FieldsF.ChF:3-415) over the whole box of J (ghosts included), with the
 * index clipping of the Fortran against the bounds of the E arrays.  E0/E/J0/J = x, y, z components in the order
 * computeJfromMassMatrices passes them (1D: Ex, Ev comp 0, Ev comp 1). */
extern "C" void orc_compute_J_from_mass_matrices(int D, const int *ncomp /* [9][2] */, const orc_mfab *sigma,
                                                 const orc_fab *E0, const orc_fab *E, const orc_fab *J0,
                                                 orc_fab *J) {
  for (int row = 0; row < 3; ++row) {
    const View Jv(J[row]), J0v(J0[row]);
    /* offsets (:30-40, :168-178, :306-316): (Nc-1)/2, but Nc/2 in the directions listed per row */
    int off[3][2];
    for (int e = 0; e < 3; ++e)
      for (int d = 0; d < 2; ++d) {
        const int Nc = ncomp[2 * (3 * row + e) + d];
        off[e][d] = (Nc - 1) / 2;
      }
    if (row == 0) {
      off[1][1] = ncomp[2 * (3 * 0 + 1) + 1] / 2;          /* xy, dir 1 */
    } else if (row == 1) {
      off[0][0] = ncomp[2 * (3 * 1 + 0) + 0] / 2;          /* yx, dir 0 */
    } else {
      off[0][0] = ncomp[2 * (3 * 2 + 0) + 0] / 2;          /* zx, dir 0 */
      off[1][1] = ncomp[2 * (3 * 2 + 1) + 1] / 2;          /* zy, dir 1 */
    }
    for (int j = Jv.lo1; j < Jv.lo1 + Jv.n1; ++j)
      for (int i = Jv.lo0; i < Jv.lo0 + Jv.n0; ++i) {
        double sigdE_e[3];
        for (int e = 0; e < 3; ++e) {
          const MView S(sigma[3 * row + e]);
          const View Ev(E[e]), E0v(E0[e]);
          const int N0 = ncomp[2 * (3 * row + e)], N1 = (D >= 2) ? ncomp[2 * (3 * row + e) + 1] : 1;
          const int o0 = off[e][0], o1 = (D >= 2) ? off[e][1] : 0;
          const int ehi0 = Ev.lo0 + Ev.n0 - 1, ehi1 = Ev.lo1 + Ev.n1 - 1;
          const int ii_min = std::max(0, o0 + Ev.lo0 - i), ii_max = std::min(N0 - 1, o0 + ehi0 - i);
          int jj_min = 0, jj_max = 0;
          if (D >= 2) {
            jj_min = std::max(0, o1 + Ev.lo1 - j);
            jj_max = std::min(N1 - 1, o1 + ehi1 - j);
          }
          double acc = 0.0;
          for (int ii = ii_min; ii <= ii_max; ++ii)
            for (int jj = jj_min; jj <= jj_max; ++jj) {
              const double dE = Ev(i + ii - o0, j + jj - o1) - E0v(i + ii - o0, j + jj - o1);
              const int Nc = ii + N0 * jj;
              acc = acc + S(i, j, Nc) * dE;
            }
          sigdE_e[e] = acc;
        }
        const double sigdE = sigdE_e[0] + sigdE_e[1] + sigdE_e[2];
        Jv(i, j) = J0v(i, j) + sigdE;
      }
  }
}
