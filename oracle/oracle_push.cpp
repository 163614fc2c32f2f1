/*
 * oracle_push.cpp -- TEST INFRASTRUCTURE ONLY (see picnic_oracle.h).
 *
 * CPU restatement of the particle advance loops of
 *   src/species/pic/PicSpeciesUtils.cpp:8-101            (Boris, PLANAR)
 *   src/species/pic/charged/PicChargedSpecies.cpp        (position updates,
 *       stepNormTransfer, advanceParticles[Iteratively], binning, moments, BCs)
 *   src/species/pic/PicSpeciesInterface.cpp:1627-1721    (Debye length)
 * Planar push.  Default = the non-relativistic build; orc_set_relativistic(1, higuera_cary) switches every
 * routine here (and orc_deposit_current_rel) to the code the reference compiles with -DRELATIVISTIC_PARTICLES.
 */
#include <algorithm>
#include <cmath>
#include <vector>

#include "picnic_oracle.h"

namespace {
/* src/core/PicnicConstants.H:13-51 */
const double kPI = M_PI;
const double kTWOPI = 2.0 * M_PI;
const double kFOURPI = 4.0 * M_PI;
const double kCVAC = 2.99792458e+08;
const double kMU0 = kFOURPI * 1.0e-7;
const double kEP0 = 1.0 / kCVAC / kCVAC / kMU0;
const double kME = 9.10938370e-31;
const double kQE = 1.60217663e-19;
const double kH = 6.62607015e-34;
const double kHBAR = kH / kTWOPI;
const double kEV_PER_JOULE = 1.0 / kQE;
}  // namespace

static int g_rel = 0, g_hc = 0;
extern "C" void orc_set_relativistic(int relativistic, int higuera_cary) {
  g_rel = relativistic;
  g_hc = higuera_cary;
}
extern "C" int orc_get_relativistic(void) { return g_rel; }

/* PicSpeciesUtils::getImplicitGamma (PicSpeciesUtils.H:43-52) */
extern "C" double orc_implicit_gamma(const double *upold, const double *upbar) {
  double upnew[3];
  for (int n = 0; n < 3; ++n) upnew[n] = 2.0 * upbar[n] - upold[n];
  const double gbsq_old = upold[0] * upold[0] + upold[1] * upold[1] + upold[2] * upold[2];
  const double gbsq_new = upnew[0] * upnew[0] + upnew[1] * upnew[1] + upnew[2] * upnew[2];
  return 0.5 * (std::sqrt(1.0 + gbsq_old) + std::sqrt(1.0 + gbsq_new));
}
static double implicit_gamma_soa(long n, long p, const double *vold, const double *v) {
  const double uo[3] = {vold[p], vold[n + p], vold[2 * n + p]}, ub[3] = {v[p], v[n + p], v[2 * n + p]};
  return orc_implicit_gamma(uo, ub);
}

/* The curvilinear velocity pushes of PicChargedSpecies::applyForces (PicChargedSpecies.cpp:341-355):
 *   type 1 CYL_CYL  PicSpeciesUtils::applyForces_CYL_CYL (PicSpeciesUtils.cpp:103-206): Boris with the inertia term as an
 *                   extra B_z = sin(dtheta); dtheta (pos_virt[0]) is predicted from u_old and r_old when it is zero and
 *                   then corrected with the new velocity and stored (the caller iterates);
 *   type 2 SPH_SPH  applyForces_SPH_SPH (:208-309): the same with dtheta, dphi (pos_virt[0], [1]), sines of the angles;
 *   type 3 CYL_HYB  applyForces_CYL_HYB (:311-383): Cartesian Boris on u_old rotated by the time-centred angle pos_virt[0];
 *   type 4 SPH_HYB  applyForces_SPH_HYB (:385-473): the same with two angles.
 * r_old = x_old[0] (types 1, 2).  virt[k*n+p], k < 2, is read and (types 1, 2) written.  anticyclic: particle arrays are
 * stored {X, Z, Y} (cyl_RZ).  HYB types always return the time-centred velocity (the reference has no byHalfDt there).
 * The relativistic build divides the scaled B by gamma = sqrt(1 + |vm|^2) (no Higuera-Cary variant in these pushes). */
extern "C" int orc_boris_curvilinear(int type, long n, double *v, const double *vold, const double *Ep, const double *Bp,
                                     const double *r_old, double *virt, double fnorm, double cnormDt, int byHalfDt,
                                     int anticyclic) {
  if (type < 1 || type > 4) return -1;
  const double alpha = fnorm * cnormDt / 2.0;
  int dirp[3] = {0, 1, 2};
  if (anticyclic && (type == 1 || type == 3)) {
    dirp[1] = 2;
    dirp[2] = 1;
  }
  for (long p = 0; p < n; ++p) {
    const double uo[3] = {vold[dirp[0] * n + p], vold[dirp[1] * n + p], vold[dirp[2] * n + p]};
    const double E[3] = {Ep[dirp[0] * n + p], Ep[dirp[1] * n + p], Ep[dirp[2] * n + p]};
    const double B[3] = {Bp[dirp[0] * n + p], Bp[dirp[1] * n + p], Bp[dirp[2] * n + p]};
    double vm0 = uo[0] + alpha * E[0];
    double vm1 = uo[1] + alpha * E[1];
    double vm2 = uo[2] + alpha * E[2];
    double bp0, bp1, bp2, gammap = 1.0, up[3];
    if (type == 1) {
      bp0 = alpha * B[0];
      bp1 = alpha * B[1];
      if (g_rel) {
        gammap = std::sqrt(1.0 + vm0 * vm0 + vm1 * vm1 + vm2 * vm2);
        bp0 /= gammap;
        bp1 /= gammap;
      }
      double dtheta = virt[p];
      bool set_dtheta = false;
      if (dtheta == 0.0) {
        dtheta = cnormDt / 2.0 * uo[1] / r_old[p] / gammap;
        set_dtheta = true;
      }
      bp2 = alpha * B[2] / gammap + std::sin(dtheta);
      const double denom = 1.0 + bp0 * bp0 + bp1 * bp1 + bp2 * bp2;
      const double vpr0 = vm0 + vm1 * bp2 - vm2 * bp1;
      const double vpr1 = vm1 + vm2 * bp0 - vm0 * bp2;
      const double vpr2 = vm2 + vm0 * bp1 - vm1 * bp0;
      up[0] = vm0 + (vpr1 * bp2 - vpr2 * bp1) / denom;
      up[1] = vm1 + (vpr2 * bp0 - vpr0 * bp2) / denom;
      up[2] = vm2 + (vpr0 * bp1 - vpr1 * bp0) / denom;
      if (set_dtheta) {
        const double rpbar = r_old[p] + cnormDt / 2.0 * up[0];
        dtheta = cnormDt / 2.0 * up[1] / rpbar / gammap;
        virt[p] = dtheta;
      }
      if (!byHalfDt)
        for (int k = 0; k < 3; ++k) up[k] = 2.0 * up[k] - uo[k];
    } else if (type == 2) {
      if (g_rel) gammap = std::sqrt(1.0 + vm0 * vm0 + vm1 * vm1 + vm2 * vm2);
      bp0 = alpha * B[0] / gammap;
      bp1 = alpha * B[1] / gammap;
      bp2 = alpha * B[2] / gammap;
      double dtheta = virt[p], dphi = virt[n + p];
      bool set_dtheta = false;
      if (dtheta == 0.0) {
        dtheta = cnormDt / 2.0 * uo[1] / r_old[p] / gammap;
        dphi = cnormDt / 2.0 * uo[2] / r_old[p] / gammap;
        dtheta = std::sin(dtheta);
        dphi = std::sin(dphi);
        set_dtheta = true;
      }
      bp0 += std::sin(dphi) * dtheta;
      bp1 -= dphi;
      bp2 += std::cos(dphi) * dtheta;
      const double denom = 1.0 + bp0 * bp0 + bp1 * bp1 + bp2 * bp2;
      const double vpr0 = vm0 + vm1 * bp2 - vm2 * bp1;
      const double vpr1 = vm1 + vm2 * bp0 - vm0 * bp2;
      const double vpr2 = vm2 + vm0 * bp1 - vm1 * bp0;
      up[0] = vm0 + (vpr1 * bp2 - vpr2 * bp1) / denom;
      up[1] = vm1 + (vpr2 * bp0 - vpr0 * bp2) / denom;
      up[2] = vm2 + (vpr0 * bp1 - vpr1 * bp0) / denom;
      if (set_dtheta) {
        const double rpbar = r_old[p] + cnormDt / 2.0 * up[0];
        dphi = cnormDt / 2.0 * up[2] / rpbar / gammap;
        dphi = std::sin(dphi);
        dtheta = cnormDt / 2.0 * up[1] / rpbar / gammap / std::cos(dphi);
        dtheta = std::sin(dtheta);
        virt[p] = dtheta;
        virt[n + p] = dphi;
      }
      if (!byHalfDt)
        for (int k = 0; k < 3; ++k) up[k] = 2.0 * up[k] - uo[k];
    } else {
      bp0 = alpha * B[0];
      bp1 = alpha * B[1];
      bp2 = alpha * B[2];
      if (g_rel) {
        gammap = std::sqrt(1.0 + vm0 * vm0 + vm1 * vm1 + vm2 * vm2);
        bp0 /= gammap;
        bp1 /= gammap;
        bp2 /= gammap;
      }
      const double denom = 1.0 + bp0 * bp0 + bp1 * bp1 + bp2 * bp2;
      if (type == 3) {
        const double dthp = virt[p];
        const double costhp = std::cos(dthp), sinthp = std::sin(dthp);
        const double upoldr_2 = costhp * uo[0] + sinthp * uo[1];
        const double upoldth_2 = -sinthp * uo[0] + costhp * uo[1];
        vm0 = upoldr_2 + alpha * E[0];
        vm1 = upoldth_2 + alpha * E[1];
      } else {
        const double thp = virt[p], php = virt[n + p];
        const double costhp = std::cos(thp), sinthp = std::sin(thp), cosphp = std::cos(php), sinphp = std::sin(php);
        const double upoldr_2 = cosphp * (costhp * uo[0] + sinthp * uo[1]) + sinphp * uo[2];
        const double upoldth_2 = -sinthp * uo[0] + costhp * uo[1];
        const double upoldph_2 = -sinphp * (costhp * uo[0] + sinthp * uo[1]) + cosphp * uo[2];
        vm0 = upoldr_2 + alpha * E[0];
        vm1 = upoldth_2 + alpha * E[1];
        vm2 = upoldph_2 + alpha * E[2];
      }
      const double vpr0 = vm0 + vm1 * bp2 - vm2 * bp1;
      const double vpr1 = vm1 + vm2 * bp0 - vm0 * bp2;
      const double vpr2 = vm2 + vm0 * bp1 - vm1 * bp0;
      up[0] = vm0 + (vpr1 * bp2 - vpr2 * bp1) / denom;
      up[1] = vm1 + (vpr2 * bp0 - vpr0 * bp2) / denom;
      up[2] = vm2 + (vpr0 * bp1 - vpr1 * bp0) / denom;
    }
    v[dirp[0] * n + p] = up[0];
    v[dirp[1] * n + p] = up[1];
    v[dirp[2] * n + p] = up[2];
  }
  return 0;
}

/* PicSpeciesUtils::applyForces (PicSpeciesUtils.cpp:8-101), dirp = {0,1,2}. */
extern "C" void orc_boris(long n, double *v, const double *vold, const double *Ep,
                          const double *Bp, double fnorm, double cnormDt,
                          int byHalfDt) {
  const double alpha = fnorm * cnormDt / 2.0;
  for (long p = 0; p < n; ++p) {
    const double vm0 = vold[p] + alpha * Ep[p];
    const double vm1 = vold[n + p] + alpha * Ep[n + p];
    const double vm2 = vold[2 * n + p] + alpha * Ep[2 * n + p];
    double bp0 = alpha * Bp[p];
    double bp1 = alpha * Bp[n + p];
    double bp2 = alpha * Bp[2 * n + p];
    if (g_rel) { /* time-centred relativistic factor (:55-78) */
      double root;
      if (g_hc) {
        const double vmsq = vm0 * vm0 + vm1 * vm1 + vm2 * vm2;
        const double vmdbp = vm0 * bp0 + vm1 * bp1 + vm2 * bp2;
        const double bpsq = bp0 * bp0 + bp1 * bp1 + bp2 * bp2;
        const double c1 = 1.0 + vmsq - bpsq;
        const double c2 = bpsq + vmdbp * vmdbp;
        root = 0.5 * (c1 + std::sqrt(c1 * c1 + 4.0 * c2));
      } else {
        root = 1.0 + vm0 * vm0 + vm1 * vm1 + vm2 * vm2;
      }
      const double gammap = std::sqrt(root);
      bp0 /= gammap;
      bp1 /= gammap;
      bp2 /= gammap;
    }
    const double denom = 1.0 + bp0 * bp0 + bp1 * bp1 + bp2 * bp2;
    const double vpr0 = vm0 + vm1 * bp2 - vm2 * bp1;
    const double vpr1 = vm1 + vm2 * bp0 - vm0 * bp2;
    const double vpr2 = vm2 + vm0 * bp1 - vm1 * bp0;
    double up0 = vm0 + (vpr1 * bp2 - vpr2 * bp1) / denom;
    double up1 = vm1 + (vpr2 * bp0 - vpr0 * bp2) / denom;
    double up2 = vm2 + (vpr0 * bp1 - vpr1 * bp0) / denom;
    if (!byHalfDt) {
      up0 = 2.0 * up0 - vold[p];
      up1 = 2.0 * up1 - vold[n + p];
      up2 = 2.0 * up2 - vold[2 * n + p];
    }
    v[p] = up0;
    v[n + p] = up1;
    v[2 * n + p] = up2;
  }
}

/* PicChargedSpecies::advancePositionsExplicit (PicChargedSpecies.cpp:483-504) */
extern "C" void orc_advance_positions_explicit(int D, long n, double *x,
                                               const double *xold,
                                               const double *v, double cnormDt) {
  for (long p = 0; p < n; ++p) {
    if (g_rel) { /* :496-498; v holds the 3 components of gamma*beta */
      const double gammap = std::sqrt(1.0 + v[p] * v[p] + v[n + p] * v[n + p] + v[2 * n + p] * v[2 * n + p]);
      for (int d = 0; d < D; ++d) x[d * n + p] = xold[d * n + p] + v[d * n + p] / gammap * cnormDt;
      continue;
    }
    for (int d = 0; d < D; ++d) x[d * n + p] = xold[d * n + p] + v[d * n + p] * cnormDt;
  }
}

/* PicChargedSpecies::advancePositionsImplicit (PicChargedSpecies.cpp:523-561) */
extern "C" void orc_advance_positions_implicit(int D, long n, double *x,
                                               const double *xold,
                                               const double *v, double cnormDt) {
  const double cnormHalfDt = cnormDt * 0.5;
  for (long p = 0; p < n; ++p)
    for (int d = 0; d < D; ++d) x[d * n + p] = xold[d * n + p] + v[d * n + p] * cnormHalfDt;
}
/* the RELATIVISTIC_PARTICLES build of the same (:548-553) needs the old velocity as well */
extern "C" void orc_advance_positions_implicit_rel(int D, long n, double *x, const double *xold, const double *v,
                                                   const double *vold, double cnormDt) {
  const double cnormHalfDt = cnormDt * 0.5;
  for (long p = 0; p < n; ++p) {
    const double gammap = implicit_gamma_soa(n, p, vold, v);
    for (int d = 0; d < D; ++d) x[d * n + p] = xold[d * n + p] + v[d * n + p] / gammap * cnormHalfDt;
  }
}

/* PicChargedSpecies::advancePositions_2ndHalf (PicChargedSpecies.cpp:1015-1025) */
extern "C" void orc_advance_positions_2nd_half(int D, long n, double *x,
                                               const double *xold) {
  for (long i = 0; i < (long)D * n; ++i) x[i] = 2.0 * x[i] - xold[i];
}

/* PicChargedSpecies::advanceVelocities_2ndHalf (PicChargedSpecies.cpp:1228-1244) */
extern "C" void orc_advance_velocities_2nd_half(long n, double *v,
                                                const double *vold) {
  for (long i = 0; i < 3 * n; ++i) v[i] = 2.0 * v[i] - vold[i];
}

/* PicChargedSpecies::averageVelocities (PicChargedSpecies.cpp:1202-1226) */
extern "C" void orc_average_velocities(long n, double *v, const double *vold) {
  for (long i = 0; i < 3 * n; ++i) v[i] = (v[i] + vold[i]) / 2.0;
}

/* ---- external fields ------------------------------------------------------------------------------------------- */
static orc_ext_fields g_ext = {};
extern "C" void orc_set_external_fields(const orc_ext_fields *ext) {
  if (ext) g_ext = *ext;
  else g_ext.on = 0;
}
extern "C" double orc_ext_value(const orc_ext_fn *f, int D, const double *x) {
  const double PI = M_PI, TWOPI = 2.0 * M_PI;   /* PicnicConstants.H:15-17 */
  switch (f->type) {
    case 1: return f->value;                     /* Constant.H:30-32 */
    case 2: {                                    /* Cosine.H:31-45 */
      double value = f->value;
      for (int dir = 0; dir < D; ++dir) {
        const double X0 = x[dir], L0 = f->L[dir];
        double arg = TWOPI * f->mode[dir] * X0 / L0 + f->phase[dir] * PI;
        arg = std::fmod(arg, TWOPI);
        value = value * std::cos(arg);
      }
      value = value + f->constant;
      return value;
    }
    case 3: {                                    /* Heavyside.H:40-54 */
      double prod = 1.0;
      for (int dir = 0; dir < D; ++dir) {
        const double arg = x[dir] - f->X0[dir];
        double H = (arg < 0.0) ? 0.0 : 1.0;
        if (std::fabs(arg) < f->eps[dir]) H = 0.5;
        const double v = f->C[dir] + f->A[dir] * H;
        prod = (dir == 0) ? v : prod * v;        /* RealVect::product() */
      }
      return prod;
    }
    default: return 0.0;
  }
}
/* PicChargedSpecies::addExternalFieldsToParticles (PicChargedSpecies.cpp:3967-3996) */
extern "C" void orc_add_external_fields(int D, long n, const double *x, double *Ep, double *Bp) {
  if (!g_ext.on) return;
  for (long p = 0; p < n; ++p) {
    double xp[2] = {0.0, 0.0};
    for (int d = 0; d < D; ++d) xp[d] = x[d * n + p];
    for (int c = 0; c < 3; ++c) {
      Ep[c * n + p] += orc_ext_value(&g_ext.f[c], D, xp);
      Bp[c * n + p] += orc_ext_value(&g_ext.f[3 + c], D, xp);
    }
  }
}

/* PicChargedSpecies::advanceParticles (PicChargedSpecies.cpp:1594-1612) */
extern "C" int orc_advance_particles(const orc_geom *g, int interpE, long n,
                                     double *x, const double *xold, double *v,
                                     const double *vold, const orc_fab *E,
                                     const orc_fab *B, double fnorm,
                                     double cnormDt, int order_swap) {
  std::vector<double> Ep(3 * n), Bp(3 * n);
  auto move = [&]() {
    if (g_rel) orc_advance_positions_implicit_rel(g->D, n, x, xold, v, vold, cnormDt);
    else orc_advance_positions_implicit(g->D, n, x, xold, v, cnormDt);
  };
  if (order_swap) move();
  const int rc = orc_gather(g, interpE, n, x, xold, E, B, Ep.data(), Bp.data());
  orc_add_external_fields(g->D, n, x, Ep.data(), Bp.data());   /* :1606 */
  orc_boris(n, v, vold, Ep.data(), Bp.data(), fnorm, cnormDt, 1);
  if (!order_swap) move();
  return rc;
}

/* stepNormTransfer (PicChargedSpecies.cpp:658-733) for one particle.
 * Returns true if the particle is converged; updates xbar as the reference does. */
static bool step_norm(const orc_geom *g, long n, long p, double *x,
                      const double *xold, const double *v, const double *vold, double cnormDt,
                      double rtol, bool reverse) {
  const double cnormHalfDt = 0.5 * cnormDt;
  double dxp[2] = {0.0, 0.0};
  double rel_diff_max = 0.0;
  const double gammap = g_rel ? implicit_gamma_soa(n, p, vold, v) : 1.0;   /* :693-698 */
  for (int d = 0; d < g->D; ++d) {
    const double dxp0 = x[d * n + p] - xold[d * n + p];
    dxp[d] = g_rel ? v[d * n + p] / gammap * cnormHalfDt : v[d * n + p] * cnormHalfDt;
    const double rel_diff_dir = std::fabs(dxp0 - dxp[d]) / g->dx[d];
    rel_diff_max = std::max(rel_diff_max, rel_diff_dir);
  }
  if (reverse) {
    if (rel_diff_max < rtol) return true;
    for (int d = 0; d < g->D; ++d) x[d * n + p] = xold[d * n + p] + dxp[d];
    return false;
  }
  for (int d = 0; d < g->D; ++d) x[d * n + p] = xold[d * n + p] + dxp[d];
  return !(rel_diff_max >= rtol);
}

/* PicChargedSpecies::advanceParticlesIteratively (PicChargedSpecies.cpp:1614-1716).
 * The linked-list transfers are restated as an "active" index list; particles
 * are independent, so list order has no effect on per-particle results. */
extern "C" int orc_advance_particles_iteratively(
    const orc_geom *g, int interpE, long n, double *x, const double *xold,
    double *v, const double *vold, const orc_fab *E, const orc_fab *B,
    double fnorm, double cnormDt, double rtol, int iter_max, long *num_apply_its,
    long *num_unconverged, int *its_out) {
  int rc = 0;
  long apply_its = 0;
  std::vector<double> Ep(3 * n), Bp(3 * n);
  if (orc_gather(g, interpE, n, x, xold, E, B, Ep.data(), Bp.data())) rc = -1;
  orc_add_external_fields(g->D, n, x, Ep.data(), Bp.data());   /* :1652 */
  orc_boris(n, v, vold, Ep.data(), Bp.data(), fnorm, cnormDt, 1);
  apply_its += n;
  std::vector<long> temp;
  for (long p = 0; p < n; ++p) {
    if (its_out) its_out[p] = 1;
    if (!step_norm(g, n, p, x, xold, v, vold, cnormDt, rtol, false)) temp.push_back(p);
  }
  int iter = 1;
  double xp[2], xpo[2], vo[3], vn[3], ep[3], bp[3];
  while (!temp.empty()) {
    std::vector<long> still;
    for (long p : temp) {
      /* single-particle gather + Boris through the same array routines */
      for (int d = 0; d < g->D; ++d) {
        xp[d] = x[d * n + p];
        xpo[d] = xold[d * n + p];
      }
      for (int c = 0; c < 3; ++c) vo[c] = vold[c * n + p];
      if (orc_gather(g, interpE, 1, xp, xpo, E, B, ep, bp)) rc = -1;
      orc_add_external_fields(g->D, 1, xp, ep, bp);              /* :1669 */
      orc_boris(1, vn, vo, ep, bp, fnorm, cnormDt, 1);
      for (int c = 0; c < 3; ++c) v[c * n + p] = vn[c];
      if (its_out) its_out[p] += 1;
    }
    apply_its += (long)temp.size();
    for (long p : temp)
      if (!step_norm(g, n, p, x, xold, v, vold, cnormDt, rtol, true)) still.push_back(p);
    temp.swap(still);
    if (temp.empty()) break;
    if (iter >= iter_max) break;
    iter += 1;
  }
  if (num_apply_its) *num_apply_its = apply_its;
  if (num_unconverged) *num_unconverged = (long)temp.size();
  return rc;
}

/* PicChargedSpecies::advanceSubOrbitParticlesAndSetJ (PicChargedSpecies.cpp:3376-3669) for the bulk sub-orbit container
 * (is_inflow_list = false), PLANAR push, no interp_bc_check.  Each particle -- one that the particle Picard loop left
 * unconverged (:1699-1706) or a "fast" one (transferFastParticles, :894-956) -- is advanced through nsub[p] equal
 * sub-steps of the time step, each an implicit push of its own (gather at the sub-orbit's x_bar, Boris half step over
 * cnormDt/nsub, stepNormTransfer in reverse mode until converged), deposits each sub-orbit's current and ends at the
 * NEW-time position and velocity with x_old, u_old restored to the start of the step.  A sub-orbit that does not
 * converge in iter_max passes restarts the particle with one more sub-orbit (unless from_emjacobian: then iter_max is
 * doubled and the unconverged state is deposited as it is).  J (three arrays, zeroed by the caller) receives the sum
 * over particles of (sum over sub-orbits of the deposit)/nsub, un-scaled (the caller multiplies by charge/volume_scale).
 * Returns 0, -1 on a gather/deposit failure, -2 if a particle needed more than max_suborbits. */
static int suborbit_core(const orc_geom *g, int interpE, int interpJ, long n, double *x, double *xold, double *v,
                         double *vold, const double *w, int *nsub, const orc_fab *E, const orc_fab *B, double fnorm,
                         double cnormDt, double rtol, int iter_max_in, int from_emjacobian, int max_suborbits, orc_fab *J,
                         int bdry_dir, int bdry_side) {
  const int D = g->D;
  const bool is_inflow_list = (bdry_dir >= 0 && bdry_side >= 0);   /* :3401-3402 */
  int rc = 0;
  int iter_max = iter_max_in;
  if (from_emjacobian) iter_max += iter_max;   /* :3400 */
  /* per-particle current, as the reference's this_Jp / this_Jpv: same boxes as J */
  std::vector<std::vector<double>> Jp(3);
  orc_fab Jpf[3];
  for (int c = 0; c < 3; ++c) {
    long sz = 1;
    for (int d = 0; d < D; ++d) sz *= (J[c].hi[d] - J[c].lo[d] + 1);
    Jp[c].assign(sz, 0.0);
    Jpf[c] = J[c];
    Jpf[c].p = Jp[c].data();
  }
  for (long p = 0; p < n; ++p) {
    int num_suborbits = nsub[p];
    double cnormDt_sub = cnormDt / num_suborbits;
    for (int c = 0; c < 3; ++c) std::fill(Jp[c].begin(), Jp[c].end(), 0.0);
    double xpold0_save[2], xpold0[2] = {0.0, 0.0}, vpold0[3];
    for (int d = 0; d < D; ++d) xpold0[d] = xpold0_save[d] = xold[d * n + p];
    for (int c = 0; c < 3; ++c) vpold0[c] = vold[c * n + p];
    if (is_inflow_list) {
      /* advanceInflowPartToBdry (:958-995): free streaming from where createInflowParticles put the particle (outside
       * the domain) to the boundary plane; the rest of the step is what the sub-orbits share (:3437-3442) */
      const double X0 = bdry_side == 0 ? g->le[bdry_dir] : g->re[bdry_dir];
      const double cnormDt0 = (X0 - xpold0[bdry_dir]) / (vpold0[bdry_dir] / 1.0);
      for (int d = 0; d < D; ++d) {
        if (d == bdry_dir) xpold0[d] = X0;
        else xpold0[d] = xpold0[d] + vpold0[d] / 1.0 * cnormDt0;
      }
      cnormDt_sub = cnormDt - cnormDt0;
      cnormDt_sub /= num_suborbits;
    }
    double xp[2] = {0.0, 0.0}, xo[2] = {0.0, 0.0}, vp[3], vo[3];
    for (int d = 0; d < D; ++d) xp[d] = xo[d] = xpold0[d];   /* x_bar guess = x_old (:3443) */
    for (int c = 0; c < 3; ++c) vp[c] = vo[c] = vpold0[c];
    bool failed = false, reflected = false;
    for (int nv = 0; nv < num_suborbits; nv++) {
      int iter = 0;
      bool restart = false;
      while (true) {
        double ep[3], bp[3];
        if (orc_gather(g, interpE, 1, xp, xo, E, B, ep, bp)) rc = -1;
        orc_add_external_fields(D, 1, xp, ep, bp);
        orc_boris(1, vp, vo, ep, bp, fnorm, cnormDt_sub, 1);
        /* stepNormTransfer(single, temp, cnormDt_sub, reverse = true) (:658-733; iter_min = 0) */
        const double cnormHalfDt = 0.5 * cnormDt_sub;
        double dxp[2] = {0.0, 0.0}, rel_diff_max = 0.0;
        for (int d = 0; d < D; ++d) {
          const double dxp0 = xp[d] - xo[d];
          dxp[d] = vp[d] * cnormHalfDt;
          rel_diff_max = std::max(rel_diff_max, std::fabs(dxp0 - dxp[d]) / g->dx[d]);
        }
        if (rel_diff_max < rtol) break;
        for (int d = 0; d < D; ++d) xp[d] = xo[d] + dxp[d];
        if (is_inflow_list) {
          /* a particle the fields turn around before it is inside (:3486-3509): it ends the call at rest normal to the
           * boundary, where it started, as a one-sub-orbit particle that deposited nothing */
          const double xpnew0 = xo[bdry_dir] + vp[bdry_dir] * cnormDt_sub;
          if ((bdry_side == 0 && xpnew0 < g->le[bdry_dir]) || (bdry_side == 1 && xpnew0 > g->re[bdry_dir])) {
            reflected = true;
            break;
          }
        }
        iter += 1;
        if (iter >= iter_max) {
          if (!from_emjacobian) {   /* :3521-3541: one more sub-orbit, start again */
            num_suborbits++;
            if (num_suborbits > max_suborbits) {
              failed = true;
              break;
            }
            for (int d = 0; d < D; ++d) xp[d] = xo[d] = xpold0[d];
            for (int c = 0; c < 3; ++c) vp[c] = vo[c] = vpold0[c];
            /* :3532-3534: the inflow branch's own value (the remaining time split once more) is overwritten by the
             * bulk formula on the next line of the reference; restated as it stands */
            cnormDt_sub = cnormDt / num_suborbits;
            for (int c = 0; c < 3; ++c) std::fill(Jp[c].begin(), Jp[c].end(), 0.0);
            restart = true;
          }
          break;   /* from_emjacobian: deposit the unconverged state (:3543-3548) */
        }
      }
      if (failed || reflected) break;
      if (restart) {
        nv = -1;
        continue;
      }
      /* 3) deposit this sub-orbit's current (:3557-3574) */
      const double wp = w[p];
      if (orc_deposit_current(g, interpJ, 1, xp, xo, vp, &wp, cnormDt_sub, Jpf)) rc = -1;
      /* 4) time-centred -> new (:3591-3594) */
      for (int c = 0; c < 3; ++c) vp[c] = 2.0 * vp[c] - vo[c];
      for (int d = 0; d < D; ++d) xp[d] = 2.0 * xp[d] - xo[d];
      /* 5) next sub-orbit starts from here (:3634-3640); the last one keeps the new values (:3603-3618) */
      if (nv < num_suborbits - 1) {
        for (int d = 0; d < D; ++d) xo[d] = xp[d];
        for (int c = 0; c < 3; ++c) vo[c] = vp[c];
      }
    }
    if (failed) {
      rc = -2;
      continue;
    }
    if (reflected) {
      nsub[p] = 1;
      for (int d = 0; d < D; ++d) x[d * n + p] = xold[d * n + p] = xpold0_save[d];
      for (int c = 0; c < 3; ++c) v[c * n + p] = vold[c * n + p] = vpold0[c];
      v[bdry_dir * n + p] = 0.0;
      continue;   /* this_Jp zeroed: nothing is added */
    }
    nsub[p] = num_suborbits;
    for (int d = 0; d < D; ++d) {
      /* inflow particles are handed back time-centred against their ORIGINAL old state (:3608-3617) */
      x[d * n + p] = is_inflow_list ? (xp[d] + xpold0_save[d]) / 2.0 : xp[d];
      xold[d * n + p] = xpold0_save[d];
    }
    for (int c = 0; c < 3; ++c) {
      v[c * n + p] = is_inflow_list ? (vp[c] + vpold0[c]) / 2.0 : vp[c];
      vold[c * n + p] = vpold0[c];
    }
    /* divide the particle's J by its number of sub-orbits (inflow: scale by the sub-step over the step) and add it to
     * the total (:3647-3654) */
    for (int c = 0; c < 3; ++c)
      for (size_t k = 0; k < Jp[c].size(); ++k) {
        if (is_inflow_list) J[c].p[k] += Jp[c][k] * (cnormDt_sub / cnormDt);
        else J[c].p[k] += Jp[c][k] / num_suborbits;
      }
  }
  return rc;
}

extern "C" int orc_advance_suborbit_particles_and_set_J(const orc_geom *g, int interpE, int interpJ, long n, double *x,
                                                        double *xold, double *v, double *vold, const double *w, int *nsub,
                                                        const orc_fab *E, const orc_fab *B, double fnorm, double cnormDt,
                                                        double rtol, int iter_max_in, int from_emjacobian,
                                                        int max_suborbits, orc_fab *J) {
  return suborbit_core(g, interpE, interpJ, n, x, xold, v, vold, w, nsub, E, B, fnorm, cnormDt, rtol, iter_max_in,
                       from_emjacobian, max_suborbits, J, -1, -1);
}

/* PicChargedSpecies::advanceInflowParticlesAndSetJ (PicChargedSpecies.cpp:3255-3322; pic_species.N.suborbit_inflow_J): the
 * particles createInflowParticles left in the inflow list of boundary (bdry_dir, bdry_side) -- x_old outside the domain,
 * moving in -- stream freely to the boundary plane and take the REST of the step as nsub[p] sub-orbits of the same kind
 * as above (advanceSubOrbitParticlesAndSetJ with is_inflow_list, :3376-3669).  On return x, v are time-centred against
 * the original x_old, u_old (which are kept), so that PicChargedSpeciesBC::inflow_Lo/Hi (2 x - x_old, :965-972) finds the
 * new-time state; a particle turned around before it is inside comes back with x = x_old, zero normal velocity and no
 * current.  J receives sum over sub-orbits of the deposit times cnormDt_sub / cnormDt, un-scaled. */
extern "C" int orc_advance_inflow_particles_and_set_J(const orc_geom *g, int interpE, int interpJ, long n, double *x,
                                                      double *xold, double *v, double *vold, const double *w, int *nsub,
                                                      const orc_fab *E, const orc_fab *B, double fnorm, double cnormDt,
                                                      double rtol, int iter_max_in, int from_emjacobian, int max_suborbits,
                                                      orc_fab *J, int bdry_dir, int bdry_side) {
  if (bdry_dir < 0 || bdry_dir >= g->D || bdry_side < 0 || bdry_side > 1) return -3;
  return suborbit_core(g, interpE, interpJ, n, x, xold, v, vold, w, nsub, E, B, fnorm, cnormDt, rtol, iter_max_in,
                       from_emjacobian, max_suborbits, J, bdry_dir, bdry_side);
}

/* PicChargedSpecies::transferFastParticles (PicChargedSpecies.cpp:894-956): flag[p] = 1 if the orbit x_old -> 2 x_bar -
 * x_old crosses more than ghosts - D faces of the half-shifted grid in any direction (CC1 only) */
extern "C" void orc_fast_particles(const orc_geom *g, long n, const double *x, const double *xold, int *flag) {
  const int D = g->D, max_crossings = g->ghosts - D;
  for (long p = 0; p < n; ++p) {
    flag[p] = 0;
    for (int dir = 0; dir < D; ++dir) {
      const double xpnew = 2.0 * x[dir * n + p] - xold[dir * n + p];
      /* the reference leaves out the domain's left edge here (:930-931): kept, it matters only for Xmin != 0 */
      const int index_old = (int)std::floor((xold[dir * n + p] - 0.5 * g->dx[dir]) / g->dx[dir]);
      const int index_new = (int)std::floor((xpnew - 0.5 * g->dx[dir]) / g->dx[dir]);
      if (std::abs(index_new - index_old) > max_crossings) flag[p] = 1;
    }
  }
}

/* BinFab::locateBin (BinFabImplem.H:582-594): (int)floor((x-origin)/dx). */
extern "C" void orc_bin(const orc_geom *g, long n, const double *x, int *cell) {
  for (long p = 0; p < n; ++p)
    for (int d = 0; d < g->D; ++d) {
      double t = x[d * n + p];
      t -= g->le[d];
      t /= g->dx[d];
      cell[d * n + p] = (int)std::floor(t);
    }
}

/* set{Number,Momentum,Energy}DensityFromBinFab (PicChargedSpecies.cpp:2881-3047),
 * cartesian (Jacobian == 1).  Sums run over the particles of a cell in the
 * order they appear in the input arrays. */
extern "C" void orc_cell_moments(const orc_geom *g, long n, const double *x,
                                 const double *v, const double *w, double mass,
                                 double volume_scale, const int *lo, const int *hi,
                                 double *dens, double *mom, double *ene) {
  const int D = g->D;
  const int n0 = hi[0] - lo[0] + 1;
  const int n1 = (D == 2) ? hi[1] - lo[1] + 1 : 1;
  const long ncell = (long)n0 * n1;
  const double dV_mapped = (D == 1) ? g->dx[0] : g->dx[0] * g->dx[1];
  const double dV_phys = dV_mapped * volume_scale;
  std::fill(dens, dens + ncell, 0.0);
  std::fill(mom, mom + 3 * ncell, 0.0);
  std::fill(ene, ene + 3 * ncell, 0.0);
  std::vector<int> cell((size_t)D * n);
  orc_bin(g, n, x, cell.data());
  for (long p = 0; p < n; ++p) {
    const int i = cell[p] - lo[0];
    const int j = (D == 2) ? cell[n + p] - lo[1] : 0;
    if (i < 0 || i >= n0 || j < 0 || j >= n1) continue;
    const long c = i + (long)j * n0;
    const double wp = w[p];
    dens[c] += wp;
    for (int k = 0; k < 3; ++k) {
      const double up = v[k * n + p];
      mom[k * ncell + c] += wp * up;
      ene[k * ncell + c] += wp * up * up;
    }
  }
  const double kn = 1.0 / dV_phys, km = mass / dV_phys, ke = 0.5 * mass / dV_phys;
  for (long c = 0; c < ncell; ++c) dens[c] *= kn;
  for (long c = 0; c < 3 * ncell; ++c) {
    mom[c] *= km;
    ene[c] *= ke;
  }
}

/* PicSpeciesInterface::setDebyeLength (PicSpeciesInterface.cpp:1627-1721) */
extern "C" void orc_debye_accumulate(long ncell, const double *dens,
                                     const double *mom, const double *ene,
                                     double mass, double charge, double *sum_inv) {
  const double mcSq_eV = kME * kCVAC * kCVAC * kEV_PER_JOULE;
  const double Aconst = kEP0 / kQE / (charge * charge);
  for (long c = 0; c < ncell; ++c) {
    const double N = dens[c];
    if (N == 0.0) continue;
    const double R = 1.0 / std::cbrt(4.0 / 3.0 * kPI * N);
    const double rho = N * mass;
    const double rhoUx = mom[c], rhoUy = mom[ncell + c], rhoUz = mom[2 * ncell + c];
    const double meanE = (rhoUx * rhoUx + rhoUy * rhoUy + rhoUz * rhoUz) / rho / 2.0;
    const double EF_eV = kHBAR * kHBAR / (2.0 * kME * mass) *
                         std::pow(3.0 * kPI * kPI * N, 2.0 / 3.0) * kEV_PER_JOULE;
    double rhoE = 0.0;
    for (int dir = 0; dir < 3; ++dir) rhoE += ene[dir * ncell + c];
    double T_eV = 2.0 / 3.0 * (rhoE - meanE) / N * mcSq_eV;
    T_eV = std::max(T_eV, 0.01);
    const double LDe_sq = std::max(Aconst * (T_eV + 2.0 / 3.0 * EF_eV) / N, R * R);
    sum_inv[c] += 1.0 / LDe_sq;
  }
}

extern "C" void orc_debye_finish(long ncell, double *a) {
  for (long c = 0; c < ncell; ++c) a[c] = 1.0 / std::sqrt(a[c]);
}

/* PicChargedSpeciesBC::enforcePeriodic (PicChargedSpeciesBC.cpp:738-765) */
extern "C" void orc_bc_periodic(long n, double *x, double *xold, double left,
                                double right) {
  const double Lbox = right - left;
  for (long p = 0; p < n; ++p) {
    if (x[p] < left) {
      x[p] = x[p] + Lbox;
      xold[p] = xold[p] + Lbox;
    }
    if (x[p] >= right) {
      x[p] = x[p] - Lbox;
      xold[p] = xold[p] - Lbox;
    }
  }
}

/* PicChargedSpeciesBC::symmetry_Lo / symmetry_Hi (PicChargedSpeciesBC.cpp:808-870) */
extern "C" void orc_bc_symmetry(long n, double *x, double *xold, double *v,
                                double *vold, double left, double right,
                                int do_lo, int do_hi) {
  for (long p = 0; p < n; ++p) {
    if (do_lo && x[p] <= left) {
      x[p] = 2. * left - x[p];
      v[p] = -v[p];
      xold[p] = 2. * left - xold[p];
      vold[p] = -vold[p];
    }
    if (do_hi && x[p] >= right) {
      x[p] = 2. * right - x[p];
      v[p] = -v[p];
      xold[p] = 2. * right - xold[p];
      vold[p] = -vold[p];
      if (x[p] == right) x[p] = 0.999999999 * right;
    }
  }
}
