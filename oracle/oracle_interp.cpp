/*
 * oracle_interp.cpp -- TEST INFRASTRUCTURE ONLY (see picnic_oracle.h).
 *
 * CPU restatement of the reference's per-particle gather / deposit kernels:
 *   src/particle_tools/MeshInterpF.ChF                 (CIC, TSC)
 *   src/particle_tools/MeshInterpChargeConservingF.ChF (CC0, CC1, truncate_boundaries)
 * and of the dispatch in src/particle_tools/MeshInterpI.H:19-228,537-690.
 *
 * The Fortran is dimension-generic through CHF_DTERM macros; here every routine
 * is written out for D==1 and D==2 with the operation order of the expanded
 * Fortran (left-to-right evaluation, true divides, integer->real conversions at
 * the same places).  Build with -O2 -ffp-contract=off.
 */
#include <cmath>
#include <vector>
#include <cstdlib>

#include "picnic_oracle.h"

namespace {

struct View {
  double *p;
  int lo0, lo1, n0;
  explicit View(const orc_fab &f)
      : p(f.p), lo0(f.lo[0]), lo1(f.lo[1]), n0(f.hi[0] - f.lo[0] + 1) {}
  inline double &operator()(int i, int j) const {
    return p[(i - lo0) + (long)(j - lo1) * n0];
  }
  inline double &operator()(int i) const { return p[i - lo0]; }
};

inline int ifloor(double a) { return (int)std::floor(a); }

/* TSC shape value for |r| (MeshInterpF.ChF:431-435 and every TSC site) */
inline double tsc_w(double r_abs) {
  if (r_abs < 0.5) return 0.75 - r_abs * r_abs;
  const double t = 1.5 - r_abs;
  return 0.5 * (t * t);
}

/* --------------------------------------------------------------------------
 * CIC: cic_interpolate_fields (MeshInterpF.ChF:497-569)
 *      cic_deposit_current    (MeshInterpF.ChF:323-390)
 * ------------------------------------------------------------------------ */
void cic_gather(const orc_geom &g, const double *xp, const View *E,
                const View *B, double *Ep, double *Bp) {
  const int D = g.D;
  int index[2] = {0, 0}, index_stag[2] = {0, 0};
  for (int d = 0; d < D; ++d) {
    index[d] = ifloor((xp[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
    index_stag[d] = ifloor((xp[d] - g.le[d]) / g.dx[d]);
  }
  for (int ii = index[0]; ii <= index[0] + 1; ++ii) {
    const double l0 = ii * g.dx[0] + 0.5 * g.dx[0] - xp[0] + g.le[0];
    const int ii_stag = ii - index[0] + index_stag[0];
    const double l0_stag = ii_stag * g.dx[0] - xp[0] + g.le[0];
    const double w0 = 1.0 - std::fabs(l0 / g.dx[0]);
    const double w0_stag = 1.0 - std::fabs(l0_stag / g.dx[0]);
    if (D == 1) {
      Ep[0] = Ep[0] + w0 * E[0](ii);
      Ep[1] = Ep[1] + w0_stag * E[1](ii_stag);
      Ep[2] = Ep[2] + w0_stag * E[2](ii_stag);
      Bp[0] = Bp[0] + w0_stag * B[0](ii_stag);
      Bp[1] = Bp[1] + w0 * B[1](ii);
      Bp[2] = Bp[2] + w0 * B[2](ii);
      continue;
    }
    for (int jj = index[1]; jj <= index[1] + 1; ++jj) {
      const double l1 = jj * g.dx[1] + 0.5 * g.dx[1] - xp[1] + g.le[1];
      const int jj_stag = jj - index[1] + index_stag[1];
      const double l1_stag = jj_stag * g.dx[1] - xp[1] + g.le[1];
      const double w1 = 1.0 - std::fabs(l1 / g.dx[1]);
      const double w1_stag = 1.0 - std::fabs(l1_stag / g.dx[1]);
      double weight;
      weight = w0 * w1_stag;
      Ep[0] = Ep[0] + weight * E[0](ii, jj_stag);
      weight = w0_stag * w1;
      Ep[1] = Ep[1] + weight * E[1](ii_stag, jj);
      weight = w0_stag * w1_stag;
      Ep[2] = Ep[2] + weight * E[2](ii_stag, jj_stag);
      weight = w0_stag * w1;
      Bp[0] = Bp[0] + weight * B[0](ii_stag, jj);
      weight = w0 * w1_stag;
      Bp[1] = Bp[1] + weight * B[1](ii, jj_stag);
      weight = w0 * w1;
      Bp[2] = Bp[2] + weight * B[2](ii, jj);
    }
  }
}

void cic_deposit_current(const orc_geom &g, const double *xp, double vpx,
                         double vpy, double vpz, double qp, const View *J) {
  const int D = g.D;
  const double volume = (D == 1) ? g.dx[0] : g.dx[0] * g.dx[1];
  const double rhop = qp / volume;
  int index[2] = {0, 0}, index_stag[2] = {0, 0};
  for (int d = 0; d < D; ++d) {
    index[d] = ifloor((xp[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
    index_stag[d] = ifloor((xp[d] - g.le[d]) / g.dx[d]);
  }
  for (int ii = index[0]; ii <= index[0] + 1; ++ii) {
    const double l0 = ii * g.dx[0] + 0.5 * g.dx[0] - xp[0] + g.le[0];
    const int ii_stag = ii - index[0] + index_stag[0];
    const double l0_stag = ii_stag * g.dx[0] - xp[0] + g.le[0];
    const double w0 = 1.0 - std::fabs(l0 / g.dx[0]);
    const double w0_stag = 1.0 - std::fabs(l0_stag / g.dx[0]);
    if (D == 1) {
      J[0](ii) = J[0](ii) + vpx * rhop * w0;
      J[1](ii_stag) = J[1](ii_stag) + vpy * rhop * w0_stag;
      J[2](ii_stag) = J[2](ii_stag) + vpz * rhop * w0_stag;
      continue;
    }
    for (int jj = index[1]; jj <= index[1] + 1; ++jj) {
      const double l1 = jj * g.dx[1] + 0.5 * g.dx[1] - xp[1] + g.le[1];
      const int jj_stag = jj - index[1] + index_stag[1];
      const double l1_stag = jj_stag * g.dx[1] - xp[1] + g.le[1];
      const double w1 = 1.0 - std::fabs(l1 / g.dx[1]);
      const double w1_stag = 1.0 - std::fabs(l1_stag / g.dx[1]);
      double weight;
      weight = w0 * w1_stag;
      J[0](ii, jj_stag) = J[0](ii, jj_stag) + vpx * rhop * weight;
      weight = w0_stag * w1;
      J[1](ii_stag, jj) = J[1](ii_stag, jj) + vpy * rhop * weight;
      weight = w0_stag * w1_stag;
      J[2](ii_stag, jj_stag) = J[2](ii_stag, jj_stag) + vpz * rhop * weight;
    }
  }
}

/* --------------------------------------------------------------------------
 * TSC: tsc_interpolate_fields (MeshInterpF.ChF:577-673)
 *      tsc_deposit_current    (MeshInterpF.ChF:398-489)
 * Note the reference evaluates (l/dx)**2 on the signed ratio in the inner
 * branch and |l/dx| in the outer one; both are reproduced through tsc_w on
 * |l/dx| because (l/dx)^2 == |l/dx|^2 exactly.
 * ------------------------------------------------------------------------ */
inline void tsc_pair(const orc_geom &g, int d, int i, int i_stag, double xp,
                     double &w, double &w_stag) {
  const double l = i * g.dx[d] + 0.5 * g.dx[d] - xp + g.le[d];
  w = tsc_w(std::fabs(l / g.dx[d]));
  const double l_stag = i_stag * g.dx[d] - xp + g.le[d];
  w_stag = tsc_w(std::fabs(l_stag / g.dx[d]));
}

void tsc_gather(const orc_geom &g, const double *xp, const View *E,
                const View *B, double *Ep, double *Bp) {
  const int D = g.D;
  int index[2] = {0, 0}, index_stag[2] = {0, 0};
  for (int d = 0; d < D; ++d) {
    index[d] = ifloor((xp[d] - g.le[d] - g.dx[d]) / g.dx[d]);
    index_stag[d] = ifloor((xp[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
  }
  for (int ii = index[0]; ii <= index[0] + 2; ++ii) {
    const int ii_stag = ii - index[0] + index_stag[0];
    double w0, w0_stag;
    tsc_pair(g, 0, ii, ii_stag, xp[0], w0, w0_stag);
    if (D == 1) {
      Ep[0] = Ep[0] + w0 * E[0](ii);
      Ep[1] = Ep[1] + w0_stag * E[1](ii_stag);
      Ep[2] = Ep[2] + w0_stag * E[2](ii_stag);
      Bp[0] = Bp[0] + w0_stag * B[0](ii_stag);
      Bp[1] = Bp[1] + w0 * B[1](ii);
      Bp[2] = Bp[2] + w0 * B[2](ii);
      continue;
    }
    for (int jj = index[1]; jj <= index[1] + 2; ++jj) {
      const int jj_stag = jj - index[1] + index_stag[1];
      double w1, w1_stag;
      tsc_pair(g, 1, jj, jj_stag, xp[1], w1, w1_stag);
      double weight;
      weight = w0 * w1_stag;
      Ep[0] = Ep[0] + weight * E[0](ii, jj_stag);
      weight = w0_stag * w1;
      Ep[1] = Ep[1] + weight * E[1](ii_stag, jj);
      weight = w0_stag * w1_stag;
      Ep[2] = Ep[2] + weight * E[2](ii_stag, jj_stag);
      weight = w0_stag * w1;
      Bp[0] = Bp[0] + weight * B[0](ii_stag, jj);
      weight = w0 * w1_stag;
      Bp[1] = Bp[1] + weight * B[1](ii, jj_stag);
      weight = w0 * w1;
      Bp[2] = Bp[2] + weight * B[2](ii, jj);
    }
  }
}

void tsc_deposit_current(const orc_geom &g, const double *xp, double vpx,
                         double vpy, double vpz, double qp, const View *J) {
  const int D = g.D;
  const double volume = (D == 1) ? g.dx[0] : g.dx[0] * g.dx[1];
  const double rhop = qp / volume;
  int index[2] = {0, 0}, index_stag[2] = {0, 0};
  for (int d = 0; d < D; ++d) {
    index[d] = ifloor((xp[d] - g.le[d] - g.dx[d]) / g.dx[d]);
    index_stag[d] = ifloor((xp[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
  }
  for (int ii = index[0]; ii <= index[0] + 2; ++ii) {
    const int ii_stag = ii - index[0] + index_stag[0];
    double w0, w0_stag;
    tsc_pair(g, 0, ii, ii_stag, xp[0], w0, w0_stag);
    if (D == 1) {
      J[0](ii) = J[0](ii) + vpx * rhop * w0;
      J[1](ii_stag) = J[1](ii_stag) + vpy * rhop * w0_stag;
      J[2](ii_stag) = J[2](ii_stag) + vpz * rhop * w0_stag;
      continue;
    }
    for (int jj = index[1]; jj <= index[1] + 2; ++jj) {
      const int jj_stag = jj - index[1] + index_stag[1];
      double w1, w1_stag;
      tsc_pair(g, 1, jj, jj_stag, xp[1], w1, w1_stag);
      double weight;
      weight = w0 * w1_stag;
      J[0](ii, jj_stag) = J[0](ii, jj_stag) + vpx * rhop * weight;
      weight = w0_stag * w1;
      J[1](ii_stag, jj) = J[1](ii_stag, jj) + vpy * rhop * weight;
      weight = w0_stag * w1_stag;
      J[2](ii_stag, jj_stag) = J[2](ii_stag, jj_stag) + vpz * rhop * weight;
    }
  }
}

/* --------------------------------------------------------------------------
 * truncate_boundaries (MeshInterpChargeConservingF.ChF:2021-2084); the 1D
 * routines inline the i0-only version (e.g. :989-1004).
 * ------------------------------------------------------------------------ */
void truncate_boundaries_2d(const orc_geom &g, double *xpold, double *xpnew,
                            double slope, double slope_inv) {
  const double xpold_save[2] = {xpold[0], xpold[1]};
  const int i0 = 0, i1 = 1;
  if (g.bc_lo[i0] == 1) {
    if (xpold[i0] < g.le[i0]) {
      xpold[i0] = g.le[i0];
      xpold[i1] = xpold_save[i1] + slope * (xpold[i0] - xpold_save[i0]);
    }
    if (xpnew[i0] < g.le[i0]) {
      xpnew[i0] = g.le[i0];
      xpnew[i1] = xpold_save[i1] + slope * (xpnew[i0] - xpold_save[i0]);
    }
  }
  if (g.bc_hi[i0] == 1) {
    if (xpold[i0] > g.re[i0]) {
      xpold[i0] = g.re[i0];
      xpold[i1] = xpold_save[i1] + slope * (xpold[i0] - xpold_save[i0]);
    }
    if (xpnew[i0] > g.re[i0]) {
      xpnew[i0] = g.re[i0];
      xpnew[i1] = xpold_save[i1] + slope * (xpnew[i0] - xpold_save[i0]);
    }
  }
  if (g.bc_lo[i1] == 1) {
    if (xpold[i1] < g.le[i1]) {
      xpold[i1] = g.le[i1];
      xpold[i0] = xpold_save[i0] + slope_inv * (xpold[i1] - xpold_save[i1]);
    }
    if (xpnew[i1] < g.le[i1]) {
      xpnew[i1] = g.le[i1];
      xpnew[i0] = xpold_save[i0] + slope_inv * (xpnew[i1] - xpold_save[i1]);
    }
  }
  if (g.bc_hi[i1] == 1) {
    if (xpold[i1] > g.re[i1]) {
      xpold[i1] = g.re[i1];
      xpold[i0] = xpold_save[i0] + slope_inv * (xpold[i1] - xpold_save[i1]);
    }
    if (xpnew[i1] > g.re[i1]) {
      xpnew[i1] = g.re[i1];
      xpnew[i0] = xpold_save[i0] + slope_inv * (xpnew[i1] - xpold_save[i1]);
    }
  }
}

inline void truncate_boundaries_1d(const orc_geom &g, double &xpold,
                                   double &xpnew) {
  if (g.bc_lo[0] == 1) {
    if (xpold < g.le[0]) xpold = g.le[0];
    if (xpnew < g.le[0]) xpnew = g.le[0];
  }
  if (g.bc_hi[0] == 1) {
    if (xpold > g.re[0]) xpold = g.re[0];
    if (xpnew > g.re[0]) xpnew = g.re[0];
  }
}

/* The CIC tail shared by the CC0/CC1 gathers: out-of-plane E and all of B at
 * xpbar (e.g. MeshInterpChargeConservingF.ChF:1971-2011 for CC1 2D, :1254-1297
 * for CC1 1D, :725-765 / :500-543 for CC0).  first_E is the first E component
 * gathered here: 1 in 1D (Ey,Ez), 2 in 2D (Ez). */
void cc_cic_tail(const orc_geom &g, const double *xpbar, const View *E,
                 const View *B, double *Ep, double *Bp) {
  const int D = g.D;
  int index[2] = {0, 0}, index_stag[2] = {0, 0};
  for (int d = 0; d < D; ++d) {
    index[d] = ifloor((xpbar[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
    index_stag[d] = ifloor((xpbar[d] - g.le[d]) / g.dx[d]);
  }
  for (int ii = index[0]; ii <= index[0] + 1; ++ii) {
    const double l0 = ii * g.dx[0] + 0.5 * g.dx[0] - xpbar[0] + g.le[0];
    const int ii_stag = ii - index[0] + index_stag[0];
    const double l0_stag = ii_stag * g.dx[0] - xpbar[0] + g.le[0];
    const double w0 = 1.0 - std::fabs(l0 / g.dx[0]);
    const double w0_stag = 1.0 - std::fabs(l0_stag / g.dx[0]);
    if (D == 1) {
      Ep[1] = Ep[1] + w0_stag * E[1](ii_stag);
      Ep[2] = Ep[2] + w0_stag * E[2](ii_stag);
      Bp[0] = Bp[0] + w0_stag * B[0](ii_stag);
      Bp[1] = Bp[1] + w0 * B[1](ii);
      Bp[2] = Bp[2] + w0 * B[2](ii);
      continue;
    }
    for (int jj = index[1]; jj <= index[1] + 1; ++jj) {
      const double l1 = jj * g.dx[1] + 0.5 * g.dx[1] - xpbar[1] + g.le[1];
      const int jj_stag = jj - index[1] + index_stag[1];
      const double l1_stag = jj_stag * g.dx[1] - xpbar[1] + g.le[1];
      const double w1 = 1.0 - std::fabs(l1 / g.dx[1]);
      const double w1_stag = 1.0 - std::fabs(l1_stag / g.dx[1]);
      double weight;
      weight = w0_stag * w1_stag;
      Ep[2] = Ep[2] + weight * E[2](ii_stag, jj_stag);
      weight = w0_stag * w1;
      Bp[0] = Bp[0] + weight * B[0](ii_stag, jj);
      weight = w0 * w1_stag;
      Bp[1] = Bp[1] + weight * B[1](ii, jj_stag);
      weight = w0 * w1;
      Bp[2] = Bp[2] + weight * B[2](ii, jj);
    }
  }
}

/* --------------------------------------------------------------------------
 * CC0 1D: cc0_1d_deposit_current (:9-158), cc0_1d_interpolate_fields (:377-545)
 * `mode` 0 = gather Ex, 1 = deposit Jx.
 * ------------------------------------------------------------------------ */
void cc0_1d_inplane(const orc_geom &g, double xpold_save, double xpbar, int mode,
                    const View &F, double &Epx, double vpx_rhop) {
  double xpold = xpold_save;
  double xpnew = 2.0 * xpbar - xpold;
  const double dXp = xpnew - xpold;
  int sign = 1;
  truncate_boundaries_1d(g, xpold, xpnew);
  const int index_old = ifloor((xpold - g.le[0]) / g.dx[0]);
  const int index_new = ifloor((xpnew - g.le[0]) / g.dx[0]);
  if (index_new < index_old) sign = -1;
  const int crossings = std::abs(index_new - index_old);
  const int num_segments = 1 + crossings;
  double Xcell = g.le[0] + (index_old + 0.5 * (1 - sign)) * g.dx[0];
  int ii_next = index_old;
  double xpold0 = xpold, xpnew0 = 0.0, dXp_sub = 0.0;
  for (int nn = 0; nn < num_segments; ++nn) {
    const int ii = ii_next;
    if (nn == num_segments - 1) {
      xpnew0 = xpnew;
      dXp_sub = xpnew0 - xpold0;
    } else {
      ii_next = ii + sign;
      Xcell = Xcell + sign * g.dx[0];
      xpnew0 = Xcell;
      dXp_sub = xpnew0 - xpold0;
    }
    double seg_factor;
    if (dXp != 0.0) seg_factor = dXp_sub / dXp;
    else seg_factor = 1.0;
    const double weight = seg_factor;
    if (mode == 0) Epx = Epx + weight * F(ii);
    else F(ii) = F(ii) + vpx_rhop * weight;
    xpold0 = xpnew0;
  }
}

/* --------------------------------------------------------------------------
 * CC1 1D: cc1_1d_deposit_current (:941-1109), cc1_1d_interpolate_fields
 * (:1118-1299).  Returns -1 when num_segments > ghosts+1 (:1018-1021).
 * The last-segment guard differs: deposit tests |dXp_sub|>0 (:1042), gather
 * tests dXp!=0 (:1222); seg_factor persists across segments (initial 1.0).
 * ------------------------------------------------------------------------ */
int cc1_1d_inplane(const orc_geom &g, double xpold_save, double xpbar, int mode,
                   const View &F, double &Epx, double vpx_rhop) {
  double xpold = xpold_save;
  double xpnew = 2.0 * xpbar - xpold;
  const double dXp = xpnew - xpold;
  int sign = 1;
  double seg_factor = 1.0;
  truncate_boundaries_1d(g, xpold, xpnew);
  const int index_old = ifloor((xpold - g.le[0] - 0.5 * g.dx[0]) / g.dx[0]);
  const int index_new = ifloor((xpnew - g.le[0] - 0.5 * g.dx[0]) / g.dx[0]);
  if (index_new < index_old) sign = -1;
  const int crossings = std::abs(index_new - index_old);
  const int num_segments = 1 + crossings;
  if (num_segments > g.ghosts + 1) return -1;
  double Xcell = g.le[0] + (index_old + 0.5 * (1 - sign) + 0.5) * g.dx[0];
  int ii_next = index_old;
  double xpold0 = xpold, xpnew0 = 0.0, dXp_sub = 0.0;
  for (int nn = 0; nn < num_segments; ++nn) {
    const int ii = ii_next;
    if (nn == num_segments - 1) {
      xpnew0 = xpnew;
      dXp_sub = xpnew0 - xpold0;
      if (mode == 1) {
        if (std::fabs(dXp_sub) > 0.0) seg_factor = dXp_sub / dXp;
      } else {
        if (dXp != 0.0) seg_factor = dXp_sub / dXp;
      }
    } else {
      ii_next = ii + sign;
      Xcell = Xcell + sign * g.dx[0];
      xpnew0 = Xcell;
      dXp_sub = xpnew0 - xpold0;
      seg_factor = dXp_sub / dXp;
    }
    const double xpbar0 = 0.5 * (xpnew0 + xpold0);
    const double l0 = g.le[0] + ii * g.dx[0] + 0.5 * g.dx[0] - xpbar0;
    const double w0 = 1.0 - std::fabs(l0 / g.dx[0]);
    double weight = w0 * seg_factor;
    if (mode == 0) Epx = Epx + weight * F(ii);
    else F(ii) = F(ii) + vpx_rhop * weight;
    weight = (1.0 - w0) * seg_factor;
    if (mode == 0) Epx = Epx + weight * F(ii + 1);
    else F(ii + 1) = F(ii + 1) + vpx_rhop * weight;
    xpold0 = xpnew0;
  }
  return 0;
}

/* Nodal CIC deposit of the out-of-plane components at xpbar in 1D
 * (cc0: :131-156, cc1: :1082-1107): Jy,Jz share w0. */
void cc_1d_virtual_deposit(const orc_geom &g, double xpbar, double vpy_rhop,
                           double vpz_rhop, const View &Jy, const View &Jz) {
  const int index_stag = ifloor((xpbar - g.le[0]) / g.dx[0]);
  for (int ii = index_stag; ii <= index_stag + 1; ++ii) {
    const double l0 = ii * g.dx[0] - xpbar + g.le[0];
    const double w0 = 1.0 - std::fabs(l0 / g.dx[0]);
    Jy(ii) = Jy(ii) + vpy_rhop * w0;
    Jz(ii) = Jz(ii) + vpz_rhop * w0;
  }
}

/* --------------------------------------------------------------------------
 * 2D orbit segmentation shared by cc0_2d_* (:204-289, :593-678) and cc1_2d_*
 * (:1524-1614, :1793-1883).  shift = 0 for CC0, 0.5 for CC1.
 * A segment callback receives (ii, jj, xpold0[2], xpnew0[2], dXp_sub[2]).
 * ------------------------------------------------------------------------ */
template <class SegFn>
int walk_2d(const orc_geom &g, const double *xpold_save, const double *xpbar,
            double shift, bool limit_segments, double *dXp_out, SegFn &&seg) {
  const int i0 = 0, i1 = 1;
  double xpold[2] = {xpold_save[0], xpold_save[1]};
  double xpnew[2], dXp[2];
  int sign[2] = {1, 1};
  for (int d = 0; d < 2; ++d) {
    xpnew[d] = 2.0 * xpbar[d] - xpold[d];
    dXp[d] = xpnew[d] - xpold[d];
  }
  const double slope = dXp[i1] / dXp[i0];
  const double slope_inv = 1 / slope;
  truncate_boundaries_2d(g, xpold, xpnew, slope, slope_inv);

  int index_old[2], index_new[2], cell_crossings[2];
  int num_segments = 1;
  for (int d = 0; d < 2; ++d) {
    index_old[d] = ifloor((xpold[d] - g.le[d] - shift * g.dx[d]) / g.dx[d]);
    index_new[d] = ifloor((xpnew[d] - g.le[d] - shift * g.dx[d]) / g.dx[d]);
    if (index_new[d] < index_old[d]) sign[d] = -1;
    cell_crossings[d] = std::abs(index_new[d] - index_old[d]);
    num_segments = num_segments + cell_crossings[d];
  }
  if (limit_segments && num_segments > g.ghosts + 1) return -1;

  double Xcell[2];
  for (int d = 0; d < 2; ++d) {
    /* CC1: (index_old + half*(1-sign) + 0.5)*dx ; CC0: (index_old + half*(1-sign))*dx */
    if (shift != 0.0)
      Xcell[d] = g.le[d] + (index_old[d] + 0.5 * (1 - sign[d]) + 0.5) * g.dx[d];
    else
      Xcell[d] = g.le[d] + (index_old[d] + 0.5 * (1 - sign[d])) * g.dx[d];
  }
  double xpold0[2] = {xpold[0], xpold[1]};
  double xpnew0[2] = {0.0, 0.0}, dXp_sub[2] = {0.0, 0.0};
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
      Xcell[i1] = Xcell[i1] + sign[i1] * g.dx[i1];
      xpnew0[i1] = Xcell[i1];
      dXp_sub[i1] = xpnew0[i1] - xpold0[i1];
      dXp_sub[i0] = slope_inv * dXp_sub[i1];
      xpnew0[i0] = xpold0[i0] + dXp_sub[i0];
    } else if (cell_crossings[i1] == 0) {
      ii_next = ii + sign[i0];
      Xcell[i0] = Xcell[i0] + sign[i0] * g.dx[i0];
      xpnew0[i0] = Xcell[i0];
      dXp_sub[i0] = xpnew0[i0] - xpold0[i0];
      dXp_sub[i1] = slope * dXp_sub[i0];
      xpnew0[i1] = xpold0[i1] + dXp_sub[i1];
    } else {
      xpnew0[i0] = Xcell[i0] + sign[i0] * g.dx[i0];
      xpnew0[i1] = Xcell[i1] + sign[i1] * g.dx[i1];
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
    seg(ii, jj, xpold0, xpnew0, dXp_sub);
    xpold0[0] = xpnew0[0];
    xpold0[1] = xpnew0[1];
  }
  dXp_out[0] = dXp[0];
  dXp_out[1] = dXp[1];
  return 0;
}

/* CC0 2D in-plane part (deposit :291-333, gather :680-714). */
int cc0_2d_inplane(const orc_geom &g, const double *xpold_save,
                   const double *xpbar, int mode, const View &Fx, const View &Fy,
                   double &Epx, double &Epy, double vpx_rhop, double vpy_rhop) {
  /* dXp is needed inside the callback: recompute it the way the reference does */
  double dXp[2];
  for (int d = 0; d < 2; ++d) {
    const double xpnew = 2.0 * xpbar[d] - xpold_save[d];
    dXp[d] = xpnew - xpold_save[d];
  }
  double dXp_unused[2];
  return walk_2d(
      g, xpold_save, xpbar, 0.0, false, dXp_unused,
      [&](int ii, int jj, const double *xpold0, const double *xpnew0,
          const double *dXp_sub) {
        const double l0 = 0.5 * (xpold0[0] + xpnew0[0]) - (g.le[0] + ii * g.dx[0]);
        const double l1 = 0.5 * (xpold0[1] + xpnew0[1]) - (g.le[1] + jj * g.dx[1]);
        double w0, w1, weight;
        /* x component */
        if (dXp[0] == 0.0) w0 = 1.0;
        else w0 = dXp_sub[0] / dXp[0];
        w1 = 1.0 - l1 / g.dx[1];
        weight = w0 * w1;
        if (mode == 0) Epx = Epx + weight * Fx(ii, jj);
        else Fx(ii, jj) = Fx(ii, jj) + vpx_rhop * weight;
        weight = w0 * (1.0 - w1);
        if (mode == 0) Epx = Epx + weight * Fx(ii, jj + 1);
        else Fx(ii, jj + 1) = Fx(ii, jj + 1) + vpx_rhop * weight;
        /* y component */
        w0 = 1.0 - l0 / g.dx[0];
        if (dXp[1] == 0.0) w1 = 1.0;
        else w1 = dXp_sub[1] / dXp[1];
        weight = w0 * w1;
        if (mode == 0) Epy = Epy + weight * Fy(ii, jj);
        else Fy(ii, jj) = Fy(ii, jj) + vpy_rhop * weight;
        weight = (1.0 - w0) * w1;
        if (mode == 0) Epy = Epy + weight * Fy(ii + 1, jj);
        else Fy(ii + 1, jj) = Fy(ii + 1, jj) + vpy_rhop * weight;
      });
}

/* CC1 2D in-plane part (deposit :1616-1709, gather :1885-1960). */
int cc1_2d_inplane(const orc_geom &g, const double *xpold_save,
                   const double *xpbar, int mode, const View &Fx, const View &Fy,
                   double &Epx, double &Epy, double vpx_rhop, double vpy_rhop) {
  double dXp[2];
  for (int d = 0; d < 2; ++d) {
    const double xpnew = 2.0 * xpbar[d] - xpold_save[d];
    dXp[d] = xpnew - xpold_save[d];
  }
  double dXp_unused[2];
  return walk_2d(
      g, xpold_save, xpbar, 0.5, true, dXp_unused,
      [&](int ii, int jj, const double *xpold0, const double *xpnew0,
          const double *dXp_sub) {
        double seg_factor[2];
        for (int d = 0; d < 2; ++d) {
          if (dXp[d] != 0.0) seg_factor[d] = dXp_sub[d] / dXp[d];
          else seg_factor[d] = 1.0;
        }
        double xpbar0[2];
        xpbar0[0] = 0.5 * (xpold0[0] + xpnew0[0]);
        xpbar0[1] = 0.5 * (xpold0[1] + xpnew0[1]);
        const double delta0 = (xpbar0[0] - (g.le[0] + (ii + 0.5) * g.dx[0])) / g.dx[0];
        const double delta1 = (xpbar0[1] - (g.le[1] + (jj + 0.5) * g.dx[1])) / g.dx[1];
        int index_start[2];
        for (int d = 0; d < 2; ++d)
          index_start[d] = ifloor((xpbar0[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);

        for (int jj_stag = index_start[1]; jj_stag <= index_start[1] + 2; ++jj_stag) {
          double l1_stag = jj_stag * g.dx[1] - xpold0[1] + g.le[1];
          double delta1_stag = std::fabs(l1_stag / g.dx[1]);
          double w1_stag = tsc_w(delta1_stag);
          l1_stag = jj_stag * g.dx[1] - xpnew0[1] + g.le[1];
          delta1_stag = std::fabs(l1_stag / g.dx[1]);
          /* w1_stag + 0.75 - d**2  evaluates as (w1_stag + 0.75) - d*d */
          if (delta1_stag < 0.5) w1_stag = w1_stag + 0.75 - delta1_stag * delta1_stag;
          else {
            const double t = 1.5 - delta1_stag;
            w1_stag = w1_stag + 0.5 * (t * t);
          }
          w1_stag = 0.5 * w1_stag;
          double weight = (1.0 - delta0) * w1_stag * seg_factor[0];
          if (mode == 0) Epx = Epx + weight * Fx(ii, jj_stag);
          else Fx(ii, jj_stag) = Fx(ii, jj_stag) + vpx_rhop * weight;
          weight = delta0 * w1_stag * seg_factor[0];
          if (mode == 0) Epx = Epx + weight * Fx(ii + 1, jj_stag);
          else Fx(ii + 1, jj_stag) = Fx(ii + 1, jj_stag) + vpx_rhop * weight;
        }
        for (int ii_stag = index_start[0]; ii_stag <= index_start[0] + 2; ++ii_stag) {
          double l0_stag = ii_stag * g.dx[0] - xpold0[0] + g.le[0];
          double delta0_stag = std::fabs(l0_stag / g.dx[0]);
          double w0_stag = tsc_w(delta0_stag);
          l0_stag = ii_stag * g.dx[0] - xpnew0[0] + g.le[0];
          delta0_stag = std::fabs(l0_stag / g.dx[0]);
          if (delta0_stag < 0.5) w0_stag = w0_stag + 0.75 - delta0_stag * delta0_stag;
          else {
            const double t = 1.5 - delta0_stag;
            w0_stag = w0_stag + 0.5 * (t * t);
          }
          w0_stag = 0.5 * w0_stag;
          double weight = w0_stag * (1.0 - delta1) * seg_factor[1];
          if (mode == 0) Epy = Epy + weight * Fy(ii_stag, jj);
          else Fy(ii_stag, jj) = Fy(ii_stag, jj) + vpy_rhop * weight;
          weight = w0_stag * delta1 * seg_factor[1];
          if (mode == 0) Epy = Epy + weight * Fy(ii_stag, jj + 1);
          else Fy(ii_stag, jj + 1) = Fy(ii_stag, jj + 1) + vpy_rhop * weight;
        }
      });
}

/* Nodal CIC deposit of Jz at xpbar in 2D.  CC0 (:344-366) forms
 * w = 1-|l/dx| with l = i*dx - xpbar + le; CC1 (:1720-1740) forms
 * delta = (i*dx - xpbar + le)/dx, w = 1-|delta|: the same arithmetic. */
void cc_2d_virtual_deposit(const orc_geom &g, const double *xpbar,
                           double vpz_rhop, const View &Jz) {
  int index_stag[2];
  for (int d = 0; d < 2; ++d) index_stag[d] = ifloor((xpbar[d] - g.le[d]) / g.dx[d]);
  for (int ii = index_stag[0]; ii <= index_stag[0] + 1; ++ii) {
    const double l0 = ii * g.dx[0] - xpbar[0] + g.le[0];
    const double w0 = 1.0 - std::fabs(l0 / g.dx[0]);
    for (int jj = index_stag[1]; jj <= index_stag[1] + 1; ++jj) {
      const double l1 = jj * g.dx[1] - xpbar[1] + g.le[1];
      const double w1 = 1.0 - std::fabs(l1 / g.dx[1]);
      const double weight = w0 * w1;
      Jz(ii, jj) = Jz(ii, jj) + vpz_rhop * weight;
    }
  }
}

}  // namespace

/* ==========================================================================
 * Dispatch: MeshInterp::interpolateEMfieldsToPart (MeshInterpI.H:537-690).
 * Ep/Bp start from zero for every particle (:552-553).
 * ======================================================================== */
/* truncate_boundaries for the mass-matrix restatement (oracle_massmatrix.cpp) */
extern "C" void orc_truncate_boundaries_2d(const orc_geom *g, double *xpold, double *xpnew, double slope,
                                           double slope_inv) {
  truncate_boundaries_2d(*g, xpold, xpnew, slope, slope_inv);
}

extern "C" int orc_gather(const orc_geom *gp, int interp, long n, const double *x,
                          const double *xold, const orc_fab *Ef,
                          const orc_fab *Bf, double *Ep_out, double *Bp_out) {
  const orc_geom &g = *gp;
  const View E[3] = {View(Ef[0]), View(Ef[1]), View(Ef[2])};
  const View B[3] = {View(Bf[0]), View(Bf[1]), View(Bf[2])};
  int rc = 0;
  for (long p = 0; p < n; ++p) {
    double Ep[3] = {0.0, 0.0, 0.0}, Bp[3] = {0.0, 0.0, 0.0};
    double xp[2] = {x[p], g.D == 2 ? x[n + p] : 0.0};
    double xpo[2] = {xold[p], g.D == 2 ? xold[n + p] : 0.0};
    switch (interp) {
      case ORC_CIC: cic_gather(g, xp, E, B, Ep, Bp); break;
      case ORC_TSC: tsc_gather(g, xp, E, B, Ep, Bp); break;
      case ORC_CC0:
        if (g.D == 1) cc0_1d_inplane(g, xpo[0], xp[0], 0, E[0], Ep[0], 0.0);
        else cc0_2d_inplane(g, xpo, xp, 0, E[0], E[1], Ep[0], Ep[1], 0.0, 0.0);
        cc_cic_tail(g, xp, E, B, Ep, Bp);
        break;
      case ORC_CC1:
        if (g.D == 1) {
          if (cc1_1d_inplane(g, xpo[0], xp[0], 0, E[0], Ep[0], 0.0)) rc = -1;
        } else {
          if (cc1_2d_inplane(g, xpo, xp, 0, E[0], E[1], Ep[0], Ep[1], 0.0, 0.0)) rc = -1;
        }
        cc_cic_tail(g, xp, E, B, Ep, Bp);
        break;
      default: return -2;
    }
    for (int c = 0; c < 3; ++c) {
      Ep_out[c * n + p] = Ep[c];
      Bp_out[c * n + p] = Bp[c];
    }
  }
  return rc;
}

/* ==========================================================================
 * Dispatch: MeshInterp::depositCurrent (MeshInterpI.H:48-228), non-relativistic
 * planar build: wpog = wp, up0 = up.
 * ======================================================================== */
/* MeshInterp::depositCurrent of the RELATIVISTIC_PARTICLES build (MeshInterpI.H:72-91): the particle weight
 * is divided by gamma -- of the stored velocity (explicit solvers) or the time-centred one of getImplicitGamma. */
extern "C" double orc_implicit_gamma(const double *upold, const double *upbar);
extern "C" int orc_deposit_current(const orc_geom *gp, int interp, long n, const double *x, const double *xold,
                                   const double *v, const double *w, double cnormDt, orc_fab *Jf);
extern "C" int orc_deposit_current_rel(const orc_geom *gp, int interp, long n, const double *x, const double *xold,
                                       const double *v, const double *vold, const double *w, double cnormDt,
                                       int from_explicit_solver, orc_fab *Jf) {
  std::vector<double> wpog(n);
  for (long p = 0; p < n; ++p) {
    double gammap = 1.0;
    if (from_explicit_solver) {
      gammap += v[p] * v[p] + v[n + p] * v[n + p] + v[2 * n + p] * v[2 * n + p];
      gammap = std::sqrt(gammap);
    } else {
      const double uo[3] = {vold[p], vold[n + p], vold[2 * n + p]}, ub[3] = {v[p], v[n + p], v[2 * n + p]};
      gammap = orc_implicit_gamma(uo, ub);
    }
    wpog[p] = w[p] / gammap;
  }
  return orc_deposit_current(gp, interp, n, x, xold, v, wpog.data(), cnormDt, Jf);
}

extern "C" int orc_deposit_current(const orc_geom *gp, int interp, long n,
                                   const double *x, const double *xold,
                                   const double *v, const double *w,
                                   double /*cnormDt*/, orc_fab *Jf) {
  const orc_geom &g = *gp;
  const View J[3] = {View(Jf[0]), View(Jf[1]), View(Jf[2])};
  const double volume = (g.D == 1) ? g.dx[0] : g.dx[0] * g.dx[1];
  int rc = 0;
  for (long p = 0; p < n; ++p) {
    double xp[2] = {x[p], g.D == 2 ? x[n + p] : 0.0};
    double xpo[2] = {xold[p], g.D == 2 ? xold[n + p] : 0.0};
    const double vpx = v[p], vpy = v[n + p], vpz = v[2 * n + p];
    const double qp = w[p];
    const double rhop = qp / volume;
    double dummy = 0.0, dummy2 = 0.0;
    switch (interp) {
      case ORC_CIC: cic_deposit_current(g, xp, vpx, vpy, vpz, qp, J); break;
      case ORC_TSC: tsc_deposit_current(g, xp, vpx, vpy, vpz, qp, J); break;
      case ORC_CC0:
        if (g.D == 1) {
          cc0_1d_inplane(g, xpo[0], xp[0], 1, J[0], dummy, vpx * rhop);
          cc_1d_virtual_deposit(g, xp[0], vpy * rhop, vpz * rhop, J[1], J[2]);
        } else {
          cc0_2d_inplane(g, xpo, xp, 1, J[0], J[1], dummy, dummy2, vpx * rhop, vpy * rhop);
          cc_2d_virtual_deposit(g, xp, vpz * rhop, J[2]);
        }
        break;
      case ORC_CC1:
        if (g.D == 1) {
          if (cc1_1d_inplane(g, xpo[0], xp[0], 1, J[0], dummy, vpx * rhop)) rc = -1;
          cc_1d_virtual_deposit(g, xp[0], vpy * rhop, vpz * rhop, J[1], J[2]);
        } else {
          if (cc1_2d_inplane(g, xpo, xp, 1, J[0], J[1], dummy, dummy2, vpx * rhop, vpy * rhop)) rc = -1;
          cc_2d_virtual_deposit(g, xp, vpz * rhop, J[2]);
        }
        break;
      default: return -2;
    }
  }
  return rc;
}

/* ==========================================================================
 * MeshInterp::deposit (MeshInterpI.H:19-46) -> cic_deposit (MeshInterpF.ChF:
 * 206-255) / tsc_deposit (:739-800); kernal == 1.
 * ======================================================================== */
extern "C" void orc_deposit_rho(const orc_geom *gp, int interp, long n,
                                const double *x, const double *w, const int *stag,
                                orc_fab *rhof) {
  const orc_geom &g = *gp;
  const View rho(*rhof);
  const int D = g.D;
  const double volume = (D == 1) ? g.dx[0] : g.dx[0] * g.dx[1];
  const double kernal = 1.0;
  for (long p = 0; p < n; ++p) {
    const double xp[2] = {x[p], D == 2 ? x[n + p] : 0.0};
    const double particle_rho = w[p] / volume;
    int index[2] = {0, 0};
    if (interp == ORC_CIC) {
      for (int d = 0; d < D; ++d)
        index[d] = ifloor((xp[d] - g.le[d] - 0.5 * g.dx[d] * (1.0 - stag[d])) / g.dx[d]);
      for (int ii = index[0]; ii <= index[0] + 1; ++ii) {
        const double l0 = ii * g.dx[0] + 0.5 * g.dx[0] * (1.0 - stag[0]) - xp[0] + g.le[0];
        const double w0 = 1.0 - std::fabs(l0 / g.dx[0]);
        if (D == 1) {
          rho(ii) = rho(ii) + kernal * particle_rho * w0;
          continue;
        }
        for (int jj = index[1]; jj <= index[1] + 1; ++jj) {
          const double l1 = jj * g.dx[1] + 0.5 * g.dx[1] * (1.0 - stag[1]) - xp[1] + g.le[1];
          const double w1 = 1.0 - std::fabs(l1 / g.dx[1]);
          const double weight = w0 * w1;
          rho(ii, jj) = rho(ii, jj) + kernal * particle_rho * weight;
        }
      }
    } else {
      for (int d = 0; d < D; ++d)
        index[d] = ifloor((xp[d] - g.le[d] - 0.5 * g.dx[d] - 0.5 * g.dx[d] * (1.0 - stag[d])) / g.dx[d]);
      for (int ii = index[0]; ii <= index[0] + 2; ++ii) {
        const double l0 = ii * g.dx[0] + 0.5 * g.dx[0] * (1.0 - stag[0]) - xp[0] + g.le[0];
        const double w0 = tsc_w(std::fabs(l0 / g.dx[0]));
        if (D == 1) {
          rho(ii) = rho(ii) + kernal * particle_rho * w0;
          continue;
        }
        for (int jj = index[1]; jj <= index[1] + 2; ++jj) {
          const double l1 = jj * g.dx[1] + 0.5 * g.dx[1] * (1.0 - stag[1]) - xp[1] + g.le[1];
          const double w1 = tsc_w(std::fabs(l1 / g.dx[1]));
          const double weight = w0 * w1;
          rho(ii, jj) = rho(ii, jj) + kernal * particle_rho * weight;
        }
      }
    }
  }
}

/* FArrayBox::mult(scalar) as used in PicChargedSpecies::setCurrentDensity
 * (PicChargedSpecies.cpp:3240-3248) and setChargeDensity* (:3072,:3110,:3155). */
extern "C" void orc_scale_fab(orc_fab *f, int D, double s) {
  long n = (long)(f->hi[0] - f->lo[0] + 1);
  if (D == 2) n *= (long)(f->hi[1] - f->lo[1] + 1);
  for (long i = 0; i < n; ++i) f->p[i] *= s;
}

/* Periodic ghost add-exchange of a deposited field held in ONE box (what
 * LevelData::exchange(reverseCopier, LDadd*Op) + exchange*() achieve for a
 * single periodic box, PicSpeciesInterface.cpp:766-772, PicChargedSpecies.cpp:
 * 3163-3166).  Chombo is not available here, so the exact copier semantics are
 * restated from the physics: every entry outside the owned index range is an
 * image of an owned entry and is added onto it; afterwards images are refreshed. */
extern "C" void orc_fold_periodic(orc_fab *f, int D, const int *stag,
                                  const int *valid_lo, const int *valid_hi,
                                  const int *periodic) {
  const View a(*f);
  const int lo0 = f->lo[0], hi0 = f->hi[0];
  const int lo1 = (D == 2) ? f->lo[1] : 0, hi1 = (D == 2) ? f->hi[1] : 0;
  /* direction 0 */
  if (periodic[0]) {
    const int N = valid_hi[0] - valid_lo[0] + 1;
    const int own_lo = valid_lo[0], own_hi = valid_hi[0]; /* nodal: node hi+1 is image of lo */
    for (int j = lo1; j <= hi1; ++j) {
      for (int i = lo0; i <= hi0; ++i) {
        if (i >= own_lo && i <= own_hi) continue;
        int im = i;
        while (im < own_lo) im += N;
        while (im > own_hi) im -= N;
        if (D == 2) a(im, j) += a(i, j); else a(im) += a(i);
      }
      for (int i = lo0; i <= hi0; ++i) {
        if (i >= own_lo && i <= own_hi) continue;
        int im = i;
        while (im < own_lo) im += N;
        while (im > own_hi) im -= N;
        if (D == 2) a(i, j) = a(im, j); else a(i) = a(im);
      }
    }
  }
  if (D == 2 && periodic[1]) {
    const int N = valid_hi[1] - valid_lo[1] + 1;
    const int own_lo = valid_lo[1], own_hi = valid_hi[1];
    for (int i = lo0; i <= hi0; ++i) {
      for (int j = lo1; j <= hi1; ++j) {
        if (j >= own_lo && j <= own_hi) continue;
        int jm = j;
        while (jm < own_lo) jm += N;
        while (jm > own_hi) jm -= N;
        a(i, jm) += a(i, j);
      }
      for (int j = lo1; j <= hi1; ++j) {
        if (j >= own_lo && j <= own_hi) continue;
        int jm = j;
        while (jm < own_lo) jm += N;
        while (jm > own_hi) jm -= N;
        a(i, j) = a(i, jm);
      }
    }
  }
  (void)stag;
}
