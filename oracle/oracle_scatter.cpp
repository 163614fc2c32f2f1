/*
 * oracle_scatter.cpp -- TEST INFRASTRUCTURE ONLY (see picnic_oracle.h).
 *
 * CPU restatement of the reference's intra-cell binary Coulomb collisions:
 *   src/scattering/ScatteringUtils.H:78-105       (computeDeltaU rotation)
 *   src/scattering/TakizukaAbe.cpp:12-53,263-578  (TA77 pairing + scattering)
 * RNG: one global std::mt19937 with std::shuffle, std::uniform_real_distribution
 * and std::normal_distribution, as exec/picnic.cpp:51 + src/utils/MathUtils.cpp:
 * 97-131.  The GPU path uses Philox and cannot share this stream; parity for
 * collisions is per-pair (explicit random numbers) and statistical.
 *
 * Masses/reduced mass are `long double` in the reference (TakizukaAbe.H:137-139);
 * kept here because x86-64 g++ provides the same 80-bit type.
 */
#include <algorithm>
#include <cmath>
#include <limits>
#include <random>
#include <vector>

#include "picnic_oracle.h"

namespace {
const double kPI = M_PI;
const double kTWOPI = 2.0 * M_PI;
const double kFOURPI = 4.0 * M_PI;
const double kCVAC = 2.99792458e+08;
const double kMU0 = kFOURPI * 1.0e-7;
const double kEP0 = 1.0 / kCVAC / kCVAC / kMU0;
const double kME = 9.10938370e-31;
const double kQE = 1.60217663e-19;

std::mt19937 global_rand_gen;

double mu_rand() {
  static std::uniform_real_distribution<> dis(0, 1);
  return dis(global_rand_gen);
}
double mu_randn() {
  static std::normal_distribution<double> disNorm(0.0, 1.0);
  return disNorm(global_rand_gen);
}
}  // namespace

extern "C" void orc_rng_seed(uint64_t seed) { global_rand_gen.seed((unsigned)seed); }

/* ScatteringUtils::collapseThreeToTwo (ScatteringUtils.H:20-47): particles 2 (weight wp2 > wp2p), its scattered
 * fraction 2p (weight wp2p) and 3 become two equally weighted particles with the same momentum and the same
 * energy per direction.  PINNED bit for bit on the reference (tests/golden/ref_pins_collapse.npz). */
extern "C" void orc_collapse_three_to_two(double *vp2, double *wp2, double *vp3, double *wp3, const double *vp2p,
                                          double wp2p) {
  const double wp23 = 0.5 * (*wp2 + *wp3);
  for (int dir = 0; dir < 3; dir++) {
    const double c23 = (wp2p * vp2p[dir] + (*wp2 - wp2p) * vp2[dir] + *wp3 * vp3[dir]) / wp23;
    const double d23 =
        (wp2p * vp2p[dir] * vp2p[dir] + (*wp2 - wp2p) * vp2[dir] * vp2[dir] + *wp3 * vp3[dir] * vp3[dir]) / wp23;
    const double arg23 = 2.0 * d23 - c23 * c23;
    vp2[dir] = 0.5 * (c23 + sqrt(arg23));
    vp3[dir] = 0.5 * (c23 - sqrt(arg23));
  }
  *wp2 = wp23;
  *wp3 = wp23;
}

/* ScatteringUtils::computeDeltaU (ScatteringUtils.H:78-105) */
extern "C" void orc_scatter_delta_u(double ux, double uy, double uz, double costh,
                                    double sinth, double cosphi, double sinphi,
                                    double *dU) {
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  const double uperp = sqrt(ux * ux + uy * uy);
  if (uperp == 0.0) {
    dU[0] = u * sinth * cosphi;
    dU[1] = u * sinth * sinphi;
    dU[2] = u * costh - u;
  } else {
    dU[0] = ux * uz / uperp * sinth * cosphi - uy * u / uperp * sinth * sinphi - ux * (1. - costh);
    dU[1] = uy * uz / uperp * sinth * cosphi + ux * u / uperp * sinth * sinphi - uy * (1. - costh);
    dU[2] = -uperp * sinth * cosphi - uz * (1. - costh);
  }
}

/* m_b90_fact (TakizukaAbe.cpp:44-49, TakizukaAbe.H:30), non-relativistic */
extern "C" double orc_ta_b90_fact(double charge1, double charge2, double mass1,
                                  double mass2) {
  const long double m1 = mass1, m2 = mass2;
  const long double mu = m1 * m2 / (m1 + m2);
  const double b90_codeToPhys = kQE * kQE / (4.0 * kPI * kEP0 * kME);
  const double cvacSq = kCVAC * kCVAC;
  /* m_charge1/2 are signed int in the reference */
  const int q1 = (int)charge1, q2 = (int)charge2;
  return (double)(abs(q1 * q2) / (mu * cvacSq) * b90_codeToPhys);
}

/* TakizukaAbe::computeDeltaU (TakizukaAbe.cpp:538-578) with the three random
 * draws made explicit. */
extern "C" void orc_ta_delta_u(const double *vp1, double den1, const double *vp2,
                               double den2, double b90_fact, double Clog,
                               double dt_sec, double gauss, double u_theta,
                               double u_phi, double *dU) {
  const double ux = vp1[0] - vp2[0];
  const double uy = vp1[1] - vp2[1];
  const double uz = vp1[2] - vp2[2];
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  const double den = std::min(den1, den2);
  const double b90 = b90_fact / (u * u);
  const double deltasq_var = kTWOPI * b90 * b90 * den * Clog * u * kCVAC * dt_sec;
  double sinth, costh;
  if (deltasq_var < 1.0) {
    const double delta = sqrt(deltasq_var) * gauss;
    const double deltasq = delta * delta;
    sinth = 2.0 * delta / (1.0 + deltasq);
    costh = 1.0 - 2.0 * deltasq / (1.0 + deltasq);
  } else {
    const double theta = kPI * u_theta;
    costh = cos(theta);
    sinth = sin(theta);
  }
  const double phi = kTWOPI * u_phi;
  orc_scatter_delta_u(ux, uy, uz, costh, sinth, cos(phi), sin(phi), dU);
}

/* ScatteringUtils::rotateVelocity (ScatteringUtils.H:49-75) */
extern "C" void orc_rotate_velocity(double *a_u, double costh, double sinth, double cosphi, double sinphi) {
  const double ux = a_u[0], uy = a_u[1], uz = a_u[2];
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  const double uperp = sqrt(ux * ux + uy * uy);
  if (uperp == 0.0) {
    a_u[0] = u * sinth * cosphi;
    a_u[1] = u * sinth * sinphi;
    a_u[2] = u * costh;
  } else {
    a_u[0] = ux * uz / uperp * sinth * cosphi - uy * u / uperp * sinth * sinphi + ux * costh;
    a_u[1] = uy * uz / uperp * sinth * cosphi + ux * u / uperp * sinth * sinphi + uy * costh;
    a_u[2] = -uperp * sinth * cosphi + uz * costh;
  }
}

/* m_b90_fact of the RELATIVISTIC_PARTICLES build (TakizukaAbe.cpp:45-46, TakizukaAbe.H:27-28): no 1/mu, and
 * twice the conversion factor */
extern "C" double orc_ta_b90_fact_rel(double charge1, double charge2) {
  const double b90_codeToPhys = kQE * kQE / (2.0 * kPI * kEP0 * kME);   /* TakizukaAbe.H:27-28: 2 pi, not 4 pi */
  const double cvacSq = kCVAC * kCVAC;
  const int q1 = (int)charge1, q2 = (int)charge2;
  return abs(q1 * q2) / cvacSq * b90_codeToPhys;
}

/* TakizukaAbe::LorentzScatter (TakizukaAbe.cpp:580-659) with the three random draws made explicit
 * (gauss is used if s12 < 2, u_theta otherwise).  long double intermediates as in the reference.
 * Returns 1 if the small-angle (gaussian) branch was taken. */
extern "C" int orc_ta_lorentz_scatter(double *a_up1, double *a_up2, double mass1, double mass2, double a_den2,
                                      double a_dt_sec, double b90_fact, double Clog, double gauss, double u_theta,
                                      double u_phi) {
  const long double a_mass1 = mass1, a_mass2 = mass2;
  long double gamma1, gamma2, Etot, gammacm, gamma1st, gamma2st;
  long double vcmdotup, vrelst, s12, muRst, upst_fact, upstsq, denom;
  double vcm[3], upst[3];
  gamma1 = sqrt(1.0 + a_up1[0] * a_up1[0] + a_up1[1] * a_up1[1] + a_up1[2] * a_up1[2]);
  gamma2 = sqrt(1.0 + a_up2[0] * a_up2[0] + a_up2[1] * a_up2[1] + a_up2[2] * a_up2[2]);
  Etot = gamma1 * a_mass1 + gamma2 * a_mass2;
  for (int n = 0; n < 3; n++) vcm[n] = (a_mass1 * a_up1[n] + a_mass2 * a_up2[n]) / Etot;
  gammacm = 1.0 / sqrt(1.0 - vcm[0] * vcm[0] - vcm[1] * vcm[1] - vcm[2] * vcm[2]);
  vcmdotup = vcm[0] * a_up2[0] + vcm[1] * a_up2[1] + vcm[2] * a_up2[2];
  gamma2st = gammacm * (gamma2 - vcmdotup);
  vcmdotup = vcm[0] * a_up1[0] + vcm[1] * a_up1[1] + vcm[2] * a_up1[2];
  gamma1st = gammacm * (gamma1 - vcmdotup);
  upst_fact = (gammacm / (1.0 + gammacm) * vcmdotup - gamma1) * gammacm;
  for (int n = 0; n < 3; n++) upst[n] = a_up1[n] + upst_fact * vcm[n];
  muRst = gamma1st * a_mass1 * gamma2st * a_mass2 / (a_mass2 * gamma2st + a_mass1 * gamma1st);
  upstsq = upst[0] * upst[0] + upst[1] * upst[1] + upst[2] * upst[2];
  denom = 1.0 + upstsq * a_mass1 / a_mass2 / gamma1st / gamma2st;
  vrelst = sqrt(upstsq) * a_mass1 / muRst / denom;
  s12 = kPI * b90_fact * b90_fact * a_den2 * Clog * vrelst * kCVAC * a_dt_sec;
  s12 *= gamma1st * gamma2st / gamma1 / gamma2;
  s12 /= pow(muRst * vrelst * vrelst, 2);
  double costh, sinth;
  int small = 0;
  if (s12 < 2.0) {
    const double delta = sqrt(s12 / 2.0) * gauss;
    const double deltasq = delta * delta;
    sinth = 2.0 * delta / (1.0 + deltasq);
    costh = 1.0 - 2.0 * deltasq / (1.0 + deltasq);
    small = 1;
  } else {
    const double theta = kPI * u_theta;
    costh = cos(theta);
    sinth = sin(theta);
  }
  const double phi = kTWOPI * u_phi;
  orc_rotate_velocity(upst, costh, sinth, cos(phi), sin(phi));
  vcmdotup = vcm[0] * upst[0] + vcm[1] * upst[1] + vcm[2] * upst[2];
  upst_fact = (gammacm / (1.0 + gammacm) * vcmdotup + gamma1st) * gammacm;
  for (int n = 0; n < 3; n++) a_up1[n] = upst[n] + upst_fact * vcm[n];
  for (int n = 0; n < 3; n++) upst[n] *= -a_mass1 / a_mass2;
  vcmdotup *= -a_mass1 / a_mass2;
  upst_fact = (gammacm / (1.0 + gammacm) * vcmdotup + gamma2st) * gammacm;
  for (int n = 0; n < 3; n++) a_up2[n] = upst[n] + upst_fact * vcm[n];
  return small;
}

extern "C" int orc_get_relativistic(void);

namespace {
/* draws exactly what TakizukaAbe::computeDeltaU draws, in its order: randn only
 * in the small-angle branch, rand for theta only in the other, then rand for phi */
void ta_pair(double *a, double *b, double den1, double den2, double b90_fact,
             double Clog, double dt_sec, long double mu, long double m1,
             long double m2, bool inter) {
  if (orc_get_relativistic()) {
    /* TakizukaAbe.cpp:336-337, 371-372, 503-505: LorentzScatter(up1, up2, m1, m2, den2); between species the one
     * with the lower density goes second.  Draws in its order: randn in the small-angle branch or rand for
     * theta, then rand for phi; the branch is found with a dry run (the draws do not enter s12). */
    double *p1 = a, *p2 = b;
    double ma = (double)m1, mb = (double)m2, den = den2;
    if (inter && den1 <= den2) {
      p1 = b, p2 = a;
      ma = (double)m2, mb = (double)m1;
      den = den1;
    }
    double t1[3] = {p1[0], p1[1], p1[2]}, t2[3] = {p2[0], p2[1], p2[2]};
    const int small = orc_ta_lorentz_scatter(t1, t2, ma, mb, den, dt_sec, b90_fact, Clog, 0.0, 0.5, 0.5);
    double gauss = 0.0, uth = 0.0;
    if (small) gauss = mu_randn();
    else uth = mu_rand();
    const double uphi = mu_rand();
    orc_ta_lorentz_scatter(p1, p2, ma, mb, den, dt_sec, b90_fact, Clog, gauss, uth, uphi);
    return;
  }
  const double ux = a[0] - b[0], uy = a[1] - b[1], uz = a[2] - b[2];
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  const double den = std::min(den1, den2);
  const double b90 = b90_fact / (u * u);
  const double deltasq_var = kTWOPI * b90 * b90 * den * Clog * u * kCVAC * dt_sec;
  double gauss = 0.0, uth = 0.0;
  if (deltasq_var < 1.0) gauss = mu_randn();
  else uth = mu_rand();
  const double uphi = mu_rand();
  double dU[3];
  orc_ta_delta_u(a, den1, b, den2, b90_fact, Clog, dt_sec, gauss, uth, uphi, dU);
  for (int dir = 0; dir < 3; ++dir) {
    a[dir] = a[dir] + mu / m1 * dU[dir];
    b[dir] = b[dir] - mu / m2 * dU[dir];
  }
}
}  // namespace

/* TakizukaAbe::applySelfScattering (TakizukaAbe.cpp:263-402).  v is [3][n]
 * component-major over cell-sorted particles. */
extern "C" void orc_ta_self(long ncell, const long *cell_start, double *v, long n,
                            const double *dens, double mass, double charge,
                            double Clog, double dt_sec, long *npairs_out) {
  const long double m1 = mass, m2 = mass;
  const long double mu = m1 * m2 / (m1 + m2);
  const double b90_fact = orc_get_relativistic() ? orc_ta_b90_fact_rel(charge, charge)
                                                 : orc_ta_b90_fact(charge, charge, mass, mass);
  long npairs = 0;
  std::vector<long> idx;
  for (long c = 0; c < ncell; ++c) {
    const double numDen = dens[c];
    if (numDen == 0.0) continue;
    const long numCell = cell_start[c + 1] - cell_start[c];
    if (numCell < 2) continue;
    int pstart = 3;
    if (numCell % 2 == 0) pstart = 0;
    idx.resize(numCell);
    for (long k = 0; k < numCell; ++k) idx[k] = cell_start[c] + k;
    std::shuffle(idx.begin(), idx.end(), global_rand_gen);
    auto scatter = [&](long p1, long p2, double den) {
      double a[3] = {v[p1], v[n + p1], v[2 * n + p1]};
      double b[3] = {v[p2], v[n + p2], v[2 * n + p2]};
      ta_pair(a, b, den, den, b90_fact, Clog, dt_sec, mu, m1, m2, false);
      for (int k = 0; k < 3; ++k) {
        v[k * n + p1] = a[k];
        v[k * n + p2] = b[k];
      }
      ++npairs;
    };
    for (long p = pstart; p < numCell; p++) {
      const long p1 = idx[p];
      p++;
      const long p2 = idx[p];
      scatter(p1, p2, numDen);
    }
    if (pstart == 3) {
      for (int p = 0; p < pstart; p++) {
        const int q1 = p % 2;
        int q2 = 2;
        if (p == 0) q2 = 1;
        /* the relativistic build passes the full density here (TakizukaAbe.cpp:372 vs :377) */
        scatter(idx[q1], idx[q2], orc_get_relativistic() ? numDen : numDen / 2.0);
      }
    }
  }
  if (npairs_out) *npairs_out = npairs;
}

/* TakizukaAbe::applyInterScattering (TakizukaAbe.cpp:404-536) */
extern "C" void orc_ta_inter(long ncell, const long *cell_start1, double *v1,
                             long n1, const double *dens1, double mass1,
                             double charge1, const long *cell_start2, double *v2,
                             long n2, const double *dens2, double mass2,
                             double charge2, double Clog, double dt_sec,
                             long *npairs_out) {
  const long double m1 = mass1, m2 = mass2;
  const long double mu = m1 * m2 / (m1 + m2);
  const double b90_fact = orc_get_relativistic() ? orc_ta_b90_fact_rel(charge1, charge2)
                                                 : orc_ta_b90_fact(charge1, charge2, mass1, mass2);
  long npairs = 0;
  std::vector<long> idx1, idx2;
  for (long c = 0; c < ncell; ++c) {
    const double numDen1 = dens1[c], numDen2 = dens2[c];
    if (numDen1 * numDen2 == 0.0) continue;
    const long numCell1 = cell_start1[c + 1] - cell_start1[c];
    const long numCell2 = cell_start2[c + 1] - cell_start2[c];
    if (numCell1 * numCell2 < 2) continue;
    const long pMin = std::min(numCell1, numCell2);
    const long pMax = std::max(numCell1, numCell2);
    idx1.resize(numCell1);
    for (long k = 0; k < numCell1; ++k) idx1[k] = cell_start1[c] + k;
    std::shuffle(idx1.begin(), idx1.end(), global_rand_gen);
    idx2.resize(numCell2);
    for (long k = 0; k < numCell2; ++k) idx2[k] = cell_start2[c] + k;
    std::shuffle(idx2.begin(), idx2.end(), global_rand_gen);
    for (long p = 0; p < pMax; p++) {
      long p1, p2;
      if (pMin == numCell1) {
        p1 = p % numCell1;
        p2 = p;
      } else {
        p1 = p;
        p2 = p % numCell2;
      }
      const long i1 = idx1[p1], i2 = idx2[p2];
      double a[3] = {v1[i1], v1[n1 + i1], v1[2 * n1 + i1]};
      double b[3] = {v2[i2], v2[n2 + i2], v2[2 * n2 + i2]};
      ta_pair(a, b, numDen1, numDen2, b90_fact, Clog, dt_sec, mu, m1, m2, true);
      for (int k = 0; k < 3; ++k) {
        v1[k * n1 + i1] = a[k];
        v2[k * n2 + i2] = b[k];
      }
      ++npairs;
    }
  }
  if (npairs_out) *npairs_out = npairs;
}

/* =====================================================================================
 * Coulomb, PROBABILISTIC weight method (src/scattering/Coulomb.cpp:14-77, 400-592,
 * 919-1180, 1642-1692, 1795-1903; Coulomb.H:339-363) and Elastic::electronImpact
 * (src/scattering/Elastic.cpp:225-476).  Non-relativistic build (Galilean scatter).
 * ===================================================================================== */
namespace {
const double kHBAR = 6.62607015e-34 / (2.0 * M_PI);   /* PicnicConstants.H:31,34: H/TWOPI */

struct CoulombConsts {
  double mass1, mass2, mu, b90_fact, bqm_fact, EF_fact;
};
CoulombConsts coulomb_consts(double charge1, double charge2, double mass1, double mass2) {
  CoulombConsts k;
  k.mass1 = mass1;
  k.mass2 = mass2;
  k.mu = mass1 * mass2 / (mass1 + mass2);
  const double qocSq = kQE * kQE / (kCVAC * kCVAC);
  const double b90_codeToPhys = qocSq / (kTWOPI * kEP0 * kME);
  k.b90_fact = std::abs(charge1 * charge2) * b90_codeToPhys;
  k.bqm_fact = kHBAR / (2.0 * kME * kCVAC);
  k.EF_fact = 0.0;
  if (mass1 == 1.0 || mass2 == 1.0)
    k.EF_fact = kHBAR * kHBAR / (2.0 * kME * k.mu) * std::pow(3.0 * kPI * kPI, 2.0 / 3.0) / (kME * kCVAC * kCVAC);
  return k;
}
}  // namespace

/* Coulomb::setNANBUcosthsinth (Coulomb.H:339-363) with the uniform draw made explicit */
extern "C" void orc_nanbu_costh_sinth(double s12, double U, double *costh, double *sinth) {
  double A12, c;
  if (s12 < 0.1466) {
    A12 = 1.0 / (s12 * (1.0 - s12 / 2.0 + s12 * s12 / 6.0));
    c = 1.0 + 1.0 / A12 * std::log(1.0 - U * (1.0 - std::exp(-2.0 * A12)));
  } else if (s12 < 3.0) {
    const double s12sq = s12 * s12, s12cu = s12 * s12sq;
    A12 = 1.0 / (0.0056958 + 0.9560202 * s12 - 0.508139 * s12sq + 0.47913906 * s12cu - 0.12788975 * s12sq * s12sq +
                 0.02389567 * s12cu * s12sq);
    c = 1.0 + 1.0 / A12 * std::log(1.0 - U * (1.0 - std::exp(-2.0 * A12)));
  } else if (s12 < 6.0) {
    A12 = 3.0 * std::exp(-s12);
    c = 1.0 + 1.0 / A12 * std::log(1.0 - U * (1.0 - std::exp(-2.0 * A12)));
  } else {
    c = 2.0 * U - 1.0;
  }
  *costh = c;
  *sinth = std::sqrt(1.0 - c * c);
}

/* scattering.coulomb.include_large_angle_scattering: the first half of Coulomb::SetPolarScattering (Coulomb.cpp:1801-1863).
 * A pair makes one Rutherford scattering event with an impact parameter below the cutoff b_c with probability SL, and the
 * variance of the cumulative small-angle part is reduced so that the total stays s12.  RL is the reference's one uniform
 * draw; returns true when the small-angle part is skipped (costh/sinth are then final).  The test value of RL used by the
 * explicit-draw entry points (orc_coulomb_delta_u, orc_coulomb_lorentz_scatter) comes from orc_coulomb_set_large_angle. */
static int g_large_angle = 0;
static double g_large_angle_draw = 0.5;
extern "C" void orc_coulomb_set_large_angle(int on, double test_draw) {
  g_large_angle = on ? 1 : 0;
  g_large_angle_draw = test_draw;
}
static bool large_angle_part(double &s12, double Clog, double b0, double bmin_qm, double sigma_eff, double RL,
                             double &costh, double &sinth) {
  double N12, N12_min, N12_tr;
  double bperp_sq, bmin_sq, bmax_sq, bc_sq, SL, ClogM;
  bperp_sq = b0 * b0 / 4.0;
  bmin_sq = bmin_qm * bmin_qm;
  bmax_sq = std::exp(2.0 * Clog) * (bperp_sq + bmin_sq) - bperp_sq;
  N12 = s12 / sigma_eff * kPI * (bmax_sq - bmin_sq);
  bc_sq = bperp_sq + bmin_sq;
  N12_min = 0.1;
  N12_tr = 80.0;
  const double N12_tr0 = N12_min / 2.0 * (bmax_sq - bmin_sq) / (bc_sq - bmin_sq);
  if (N12_tr0 < N12_tr) N12_tr = N12_tr0;
  if (N12 <= N12_min) {
    SL = N12;
  } else if (N12 <= N12_tr) {
    const double SL_tr = N12_tr * (bc_sq - bmin_sq) / (bmax_sq - bmin_sq);
    const double SL_min = N12_min;
    SL = (N12 - N12_min) / (N12_tr - N12_min) * SL_tr + (N12_tr - N12) / (N12_tr - N12_min) * SL_min;
  } else {
    SL = std::min(0.1, N12 * (bc_sq - bmin_sq) / (bmax_sq - bmin_sq));
  }
  bc_sq = bmin_sq + SL / N12 * (bmax_sq - bmin_sq);
  ClogM = 0.5 * std::log((bperp_sq + bmax_sq) / (bperp_sq + bc_sq));
  s12 *= ClogM / Clog / (1.0 - SL);
  costh = 1.0;
  sinth = 0.0;
  if (RL < SL) {
    const double bsq = bc_sq - RL / SL * (bc_sq - bmin_sq);
    costh = (bsq - bperp_sq) / (bsq + bperp_sq);
    sinth = std::sqrt(1.0 - costh * costh);
    return true;
  }
  if (N12 <= N12_min) return true;
  return false;
}

/* Coulomb angular_scattering = NANBU_FAS (3) and NANBU_FAS_v2 (4): Coulomb::setNANBUFAScosthsinth (Coulomb.H:365-428),
 * setNANBUFAS_v2_costhsinth (:430-571), setFAScoefficients (:573-618), setFAS_v2_coefficients (:620-678) and
 * getTransitionX_NANBU (:680-718).  Full-angle scattering after Higginson, JCP 2017: Nanbu's cumulative small-angle
 * distribution joined to single Rutherford events, the coefficients from a fixed-point solve per pair.
 * The reference draws its uniforms one after the other, and only those a branch needs; FasDraws hands them out in
 * that order -- from the explicit test values (u_polar, then the two of orc_coulomb_set_fas_draws) or, inside the pair
 * driver, live from the generator (the azimuth then follows the polar draws, as in GalileanScatter).
 * Where the reference prints a message and exits on a failed v2 solve, this falls back to setNANBUcosthsinth. */
static double g_fas_draw2 = 0.5, g_fas_draw3 = 0.5;
static int g_fas_live = 0;
extern "C" void orc_coulomb_set_fas_draws(double second, double third) {
  g_fas_draw2 = second;
  g_fas_draw3 = third;
}
namespace {
struct FasDraws {
  double u[3];
  int pos;
  double next() {
    if (g_fas_live) return mu_rand();
    const double v = u[pos < 3 ? pos : 2];
    ++pos;
    return v;
  }
};
double transition_x_nanbu(double Clog, double s12, double alpha_g, double sA) {
  double xc = 2.0;
  const double C0 = s12 / (8.0 * Clog * alpha_g * sA);
  if (C0 > std::exp(-2.0)) return 1.0;
  double error = 1.0;
  int iter = 0;
  while (error > 1.0e-4) {
    const double xold = xc;
    const double y0 = xc * xc * std::exp(-2.0 * xc) - C0;
    const double dy0dx = 2.0 * xc * (1.0 - xc) * std::exp(-2.0 * xc);
    xc = xc - y0 / dy0dx;
    error = std::abs(1.0 - xold / xc);
    iter += 1;
    if (iter > 20) break;
  }
  if (sA * xc > 1.0) xc = 1.0 / sA;
  return xc;
}
int fas_coefficients(double &alpha_g, double &sA, double &muc, double Clog, double s12, double mu_max) {
  alpha_g = 1.0;
  sA = s12 / 2.0;
  double Xc = transition_x_nanbu(Clog, s12, alpha_g, sA);
  muc = sA * Xc;
  int iter = 0, success = 1;
  double error = 1.0;
  while (error > 1.0e-4) {
    const double sAold = sA;
    const double f1 = 1.0 - std::exp(-s12) + s12 / Clog * 0.5 * std::log(Xc * sAold / mu_max);
    const double f2 = (1.0 - std::exp(-2.0 * Xc)) / (1.0 - (1.0 + Xc) * std::exp(-2.0 * Xc));
    sA = (4.0 * mu_max * Clog / (4.0 * mu_max * Clog + s12)) * (s12 / (4.0 * Xc * Clog) + f1 * f2);
    Xc = transition_x_nanbu(Clog, s12, alpha_g, sA);
    muc = sA * Xc;
    alpha_g = (1.0 - s12 / (4.0 * Clog) * (mu_max - muc) / mu_max / muc) / (1.0 - std::exp(-2.0 * muc / sA));
    error = std::abs(1.0 - sAold / sA);
    iter += 1;
    if (iter > 20) {
      success = -1;
      break;
    }
  }
  return success;
}
int fas_v2_coefficients(double &alpha_g, double &sA, double Clog, double s12, double mu_max, double mu_tr) {
  const double S_L = s12 / (4.0 * Clog) * mu_max / mu_tr / (mu_max + mu_tr);
  const double mu_L = s12 / (2.0 * Clog) * std::log((mu_max + mu_tr) / mu_tr);
  const double mu_N97 = 1.0 - std::exp(-s12);
  if (mu_L > mu_N97 || S_L > 1.0) {
    alpha_g = 0.0;
    sA = mu_tr;
    return -1;
  }
  sA = s12 / 2.0;
  alpha_g = std::max(0.0, (1.0 - S_L) / (1.0 - std::exp(-2.0 * mu_max / sA)));
  int iter = 0, success = 1;
  double error = 1.0;
  while (error > 1.0e-4) {
    const double sAold = sA;
    const double f1 = mu_N97 - mu_L + alpha_g * mu_max * std::exp(-2.0 * mu_max / sAold);
    const double f2 = 1.0 - S_L;
    sA = f1 / f2;
    alpha_g = std::max(0.0, (1.0 - S_L) / (1.0 - std::exp(-2.0 * mu_max / sA)));
    error = std::abs(1.0 - sAold / sA);
    iter += 1;
    if (iter > 20) {
      success = -1;
      break;
    }
  }
  return success;
}
void nanbu_fas(double s12, double Clog, double b0, double bmin_qm, double sigma_eff, FasDraws &D, double &costh,
               double &sinth) {
  const double bperp_sq = b0 * b0 / 4.0, bmin_sq = bmin_qm * bmin_qm;
  const double bmax_sq = std::exp(2.0 * Clog) * (bperp_sq + bmin_sq) - bperp_sq;
  const double s12_min = 1.33 * 4.0 * Clog / (std::exp(2.0 * Clog) - 1.0);
  costh = 1.0;
  sinth = 0.0;
  if (s12 < s12_min) {
    const double N12 = s12 / sigma_eff * kPI * (bmax_sq - bmin_sq);
    const double PL = 1.0 - std::exp(-N12);
    if (D.next() < PL) {
      const double RL = D.next();
      const double bsq = bmax_sq - RL * (bmax_sq - bmin_sq);
      costh = (bsq - bperp_sq) / (bsq + bperp_sq);
      sinth = std::sqrt(1.0 - costh * costh);
    }
  } else if (s12 < 0.5) {
    double alpha_g, sA, muc;
    const double costhmax = (bmin_sq - bperp_sq) / (bmin_sq + bperp_sq);
    const double mu_max = (1.0 - costhmax) / 2.0;
    const int success = fas_coefficients(alpha_g, sA, muc, Clog, s12, mu_max);
    if (success < 0 || muc != muc) {
      orc_nanbu_costh_sinth(s12, D.next(), &costh, &sinth);
    } else {
      const double Uc = 1.0 - s12 / (4.0 * Clog) * (mu_max - muc) / (mu_max * muc);
      const double R = D.next();
      if (R < Uc) {
        costh = 1.0 + sA * std::log(1.0 - R / Uc * (1.0 - std::exp(-2.0 * muc / sA)));
        sinth = std::sqrt(1.0 - costh * costh);
      } else {
        const double costhc = 1.0 - 2.0 * muc;
        const double R2 = D.next();
        const double numer = R2 * (costhc - costhmax) - costhc * (1.0 - costhmax);
        const double denom = R2 * (costhc - costhmax) - (1.0 - costhmax);
        costh = numer / denom;
        sinth = std::sqrt(1.0 - costh * costh);
      }
    }
  } else {
    orc_nanbu_costh_sinth(s12, D.next(), &costh, &sinth);
  }
}
void nanbu_fas_v2(double s12, double Clog, double b0, double bmin_qm, double /*sigma_eff*/, FasDraws &D, double &costh,
                  double &sinth) {
  const double bperp_sq = b0 * b0 / 4.0, bmin_sq = bmin_qm * bmin_qm;
  const double bmax_sq = std::exp(2.0 * Clog) * (bperp_sq + bmin_sq) - bperp_sq;
  const double costhmax = (bmin_sq - bperp_sq) / (bmin_sq + bperp_sq);
  const double mu_max = (1.0 - costhmax) / 2.0;
  if (s12 > 0.6) {
    orc_nanbu_costh_sinth(s12, D.next(), &costh, &sinth);
    return;
  }
  double Nmax = 2.0 * Clog;
  const double mutr_factor = Clog;
  const double sig_ratio = bperp_sq / (bmax_sq - bmin_sq);
  const double s12_Nmin = 4.0 * Clog * sig_ratio;
  double s12_Nmax = Nmax * s12_Nmin;
  const double Nmax_min = 2.0 * mutr_factor * mu_max / (std::exp(2.0 * Clog) - 1.0) / s12_Nmin;
  if (Nmax < Nmax_min) {
    Nmax = Nmax_min;
    s12_Nmax = Nmax * s12_Nmin;
  }
  if (s12 < s12_Nmax) {
    const double Ntot = s12 / s12_Nmin;
    const double Pscatter = 1.0 - std::exp(-Ntot);
    if (D.next() > Pscatter) {
      costh = 1.0;
      sinth = 0.0;
      return;
    }
    double alpha_g, sA;
    double mu_tr = s12_Nmax / mutr_factor;
    const int success = fas_v2_coefficients(alpha_g, sA, Clog, s12_Nmax, mu_max, mu_tr);
    if (success < 0 || sA != sA) {
      orc_nanbu_costh_sinth(s12, D.next(), &costh, &sinth);
      return;
    }
    alpha_g = std::max(0.0, alpha_g * (s12 - s12_Nmin) / (s12_Nmax - s12_Nmin));
    if (alpha_g == 0.0) sA = mu_max;
    else sA *= s12 / s12_Nmax;
    double mu0;
    const double coefc = 4.0 * sig_ratio / mu_max;
    if (coefc < 1.0e-10) mu0 = sig_ratio;
    else mu0 = mu_max * (-1.0 + std::sqrt(1.0 + coefc)) / 2.0;
    const double S_N97 = alpha_g * (1.0 - std::exp(-2.0 * mu_max / sA));
    if (s12 <= s12_Nmin) {
      mu_tr = mu0;
    } else {
      const double C0 = (1.0 - S_N97) * 4.0 * Clog / s12;
      const double coef = 4.0 / mu_max / C0;
      if (coef < 1.0e-10) mu_tr = 1.0 / C0;
      else mu_tr = mu_max * (-1.0 + std::sqrt(1.0 + coef)) / 2.0;
      mu_tr = std::max(mu_tr, mu0);
    }
    const double R = D.next();
    if (R < S_N97) {
      costh = 1.0 + sA * std::log(1.0 - R / S_N97 * (1.0 - std::exp(-2.0 * mu_max / sA)));
    } else {
      const double RL = D.next();
      costh = 1.0 - 2.0 * RL * mu_tr * mu_max / (mu_max * (1.0 - RL) + mu_tr);
    }
    sinth = std::sqrt(1.0 - costh * costh);
  } else {
    double alpha_g, sA;
    const double mu_tr = s12 / mutr_factor;
    const int success = fas_v2_coefficients(alpha_g, sA, Clog, s12, mu_max, mu_tr);
    if (success < 0 || sA != sA) {
      orc_nanbu_costh_sinth(s12, D.next(), &costh, &sinth);
      return;
    }
    const double S_L = s12 / (4.0 * Clog) * mu_max / mu_tr / (mu_max + mu_tr);
    const double S_N97 = 1.0 - S_L;
    const double R = D.next();
    if (R < S_N97) {
      costh = 1.0 + sA * std::log(1.0 - R / S_N97 * (1.0 - std::exp(-2.0 * mu_max / sA)));
    } else {
      const double RL = D.next();
      costh = 1.0 - 2.0 * RL * mu_tr * mu_max / (mu_max * (1.0 - RL) + mu_tr);
    }
    sinth = std::sqrt(1.0 - costh * costh);
  }
}
}  // namespace

/* the polar part alone, for the parity tests of the device kernels (variant 3 NANBU_FAS, 4 NANBU_FAS_v2) */
extern "C" void orc_nanbu_fas_costh_sinth(int variant, double s12, double Clog, double b0, double bmin_qm, double sigma_eff,
                                          double u1, double u2, double u3, double *costh, double *sinth) {
  FasDraws D = {{u1, u2, u3}, 0};
  const int live = g_fas_live;
  g_fas_live = 0;
  if (variant == 3) nanbu_fas(s12, Clog, b0, bmin_qm, sigma_eff, D, *costh, *sinth);
  else nanbu_fas_v2(s12, Clog, b0, bmin_qm, sigma_eff, D, *costh, *sinth);
  g_fas_live = live;
}

/* Coulomb::GalileanScatter + SetPolarScattering (TAKIZUKA=0, NANBU=1, BOBYLEV=2, ISOTROPIC=5)
 * and NANBU_FAS=3, NANBU_FAS_v2=4 (their second and third uniforms: orc_coulomb_set_fas_draws)
 * with explicit draws: r_polar = |randn| for TAKIZUKA's small-angle branch, else a uniform;
 * u_phi uniform.  Returns 0 and leaves dU = 0 when the reference returns early (Appendix B:
 * the reference then uses an uninitialised deltaU; treated as zero). */
extern "C" int orc_coulomb_delta_u(const double *vp1, const double *vp2, double charge1, double charge2,
                                   double mass1, double mass2, double EF_norm, double Clog_in, int angular,
                                   double den12, double bmax, double sigma_max, double dt_sec, double gauss,
                                   double u_polar, double u_phi, double *dU, double *s12_out) {
  const CoulombConsts k = coulomb_consts(charge1, charge2, mass1, mass2);
  dU[0] = dU[1] = dU[2] = 0.0;
  const double ux = vp1[0] - vp2[0], uy = vp1[1] - vp2[1], uz = vp1[2] - vp2[2];
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  if (u <= std::numeric_limits<double>::min()) return 0;
  const double vsum = std::sqrt(vp1[0] * vp1[0] + vp1[1] * vp1[1] + vp1[2] * vp1[2]) +
                      std::sqrt(vp2[0] * vp2[0] + vp2[1] * vp2[1] + vp2[2] * vp2[2]);
  if (u <= 1.0e-14 * vsum) return 0;
  double b0 = k.b90_fact / (k.mu * u * u + 2.0 * EF_norm);
  const double bmin_qm = k.bqm_fact / (k.mu * u + std::sqrt(2.0 * EF_norm * k.mu));
  double Clog = Clog_in;
  if (Clog == 0.0 && u > 0.0) {
    Clog = 0.5 * std::log((b0 * b0 / 4.0 + bmax * bmax) / (b0 * b0 / 4.0 + bmin_qm * bmin_qm));
    Clog = std::max(2.0, Clog);
  }
  b0 = k.b90_fact / (k.mu * u * u);
  double sigma_eff = kPI * b0 * b0 * Clog;
  sigma_eff = std::min(sigma_eff, sigma_max);
  double s12 = sigma_eff * den12 * u * kCVAC * dt_sec;
  double costh = 1.0, sinth = 0.0;
  bool skip_small = false;
  if (g_large_angle) skip_small = large_angle_part(s12, Clog, b0, bmin_qm, sigma_eff, g_large_angle_draw, costh, sinth);
  if (s12_out) *s12_out = skip_small ? -1.0 : s12;   /* -1: no polar draw follows */
  if (!skip_small) switch (angular) {
    case 0:
      if (s12 < 2.0) {
        const double delta = sqrt(s12 / 2.0) * std::abs(gauss);
        const double deltasq = delta * delta;
        sinth = 2.0 * delta / (1.0 + deltasq);
        costh = 1.0 - 2.0 * deltasq / (1.0 + deltasq);
      } else {
        const double theta = kPI * u_polar;
        costh = std::cos(theta);
        sinth = std::sin(theta);
      }
      break;
    case 1:
      orc_nanbu_costh_sinth(s12, u_polar, &costh, &sinth);
      break;
    case 2:
      costh = 1.0 - std::min(s12, 2.0);
      sinth = std::sin(std::acos(costh));
      break;
    case 3:
    case 4: {
      FasDraws D = {{u_polar, g_fas_draw2, g_fas_draw3}, 0};
      if (angular == 3) nanbu_fas(s12, Clog, b0, bmin_qm, sigma_eff, D, costh, sinth);
      else nanbu_fas_v2(s12, Clog, b0, bmin_qm, sigma_eff, D, costh, sinth);
      break;
    }
    default: {
      const double theta = kPI * u_polar;
      costh = std::cos(theta);
      sinth = std::sin(theta);
    }
  }
  const double phi = kTWOPI * (g_fas_live ? mu_rand() : u_phi);   /* live: the azimuth follows the polar draws */
  orc_scatter_delta_u(ux, uy, uz, costh, sinth, std::cos(phi), std::sin(phi), dU);
  return 1;
}

/* Coulomb::LorentzScatter (Coulomb.cpp:1694-1793) + SetPolarScattering (:1795-1903, small-angle part) with the
 * draws made explicit, long double scalars as in the reference.  Particle 1 always scatters, particle 2 only if
 * scatter2 (its proper velocity then follows from momentum conservation).  Returns 0 when the reference returns
 * early (vanishing relative velocity); s12_out receives m_s12. */
extern "C" int orc_coulomb_lorentz_scatter(double *a_up1, double *a_up2, int a_scatter2, double charge1,
                                           double charge2, double mass1, double mass2, double EF_norm, double Clog_in,
                                           int angular, double a_den12, double a_bmax, double a_sigma_max,
                                           double a_dt_sec, double gauss, double u_polar, double u_phi,
                                           double *s12_out) {
  const CoulombConsts k = coulomb_consts(charge1, charge2, mass1, mass2);
  const long double a_mass1 = mass1, a_mass2 = mass2;
  long double g1, g2, vcmsq, Etot, gcm, g1st, g2st;
  long double ucmdotup, vrelst, vrelst_invar, muRst, upst_fact, upstsq, denom;
  double ptot[3], vcm[3], upst[3];
  if (s12_out) *s12_out = 0.0;
  const double gb1sq = a_up1[0] * a_up1[0] + a_up1[1] * a_up1[1] + a_up1[2] * a_up1[2];
  const double gb2sq = a_up2[0] * a_up2[0] + a_up2[1] * a_up2[1] + a_up2[2] * a_up2[2];
  g1 = sqrt(1.0 + gb1sq);
  g2 = sqrt(1.0 + gb2sq);
  Etot = g1 * a_mass1 + g2 * a_mass2;
  for (int n = 0; n < 3; n++) ptot[n] = a_mass1 * a_up1[n] + a_mass2 * a_up2[n];
  for (int n = 0; n < 3; n++) vcm[n] = ptot[n] / Etot;
  vcmsq = vcm[0] * vcm[0] + vcm[1] * vcm[1] + vcm[2] * vcm[2];
  gcm = 1.0 / std::sqrt(1.0 - vcmsq);
  ucmdotup = gcm * (vcm[0] * a_up2[0] + vcm[1] * a_up2[1] + vcm[2] * a_up2[2]);
  g2st = gcm * g2 - ucmdotup;
  ucmdotup = gcm * (vcm[0] * a_up1[0] + vcm[1] * a_up1[1] + vcm[2] * a_up1[2]);
  g1st = gcm * g1 - ucmdotup;
  upst_fact = gcm * (ucmdotup / (1.0 + gcm) - g1);
  for (int n = 0; n < 3; n++) upst[n] = a_up1[n] + upst_fact * vcm[n];
  muRst = g1st * a_mass1 * g2st * a_mass2 / (g1st * a_mass1 + g2st * a_mass2);
  upstsq = upst[0] * upst[0] + upst[1] * upst[1] + upst[2] * upst[2];
  vrelst = std::sqrt(upstsq) * a_mass1 / muRst;
  if (vrelst <= std::numeric_limits<double>::min()) return 0;
  const double vsum = std::sqrt(gb1sq) / g1 + std::sqrt(gb2sq) / g2;
  if (vrelst <= 1.0e-14 * vsum) return 0;
  denom = 1.0 + upstsq * a_mass1 / a_mass2 / g1st / g2st;
  vrelst_invar = vrelst / denom;
  double b0 = k.b90_fact / (muRst * vrelst * vrelst_invar + 2.0 * EF_norm);
  const double bmin_qm = k.bqm_fact / (muRst * vrelst + std::sqrt(2.0 * EF_norm * muRst));
  double Clog = Clog_in;
  if (Clog == 0.0 && upstsq > 0.0) {
    Clog = 0.5 * std::log((b0 * b0 / 4.0 + a_bmax * a_bmax) / (b0 * b0 / 4.0 + bmin_qm * bmin_qm));
    Clog = std::max(2.0, Clog);
  }
  b0 = k.b90_fact / (muRst * vrelst * vrelst_invar);
  double sigma_eff = kPI * b0 * b0 * Clog;
  sigma_eff = std::min(sigma_eff, a_sigma_max);
  double s12 = sigma_eff * a_den12 * vrelst * kCVAC * a_dt_sec;
  s12 *= g1st * g2st / g1 / g2;
  double costh = 1.0, sinth = 0.0;
  bool skip_small = false;
  if (g_large_angle) skip_small = large_angle_part(s12, Clog, b0, bmin_qm, sigma_eff, g_large_angle_draw, costh, sinth);
  if (s12_out) *s12_out = skip_small ? -1.0 : s12;
  if (!skip_small) switch (angular) {
    case 0:
      if (s12 < 2.0) {
        const double delta = sqrt(s12 / 2.0) * std::abs(gauss);
        const double deltasq = delta * delta;
        sinth = 2.0 * delta / (1.0 + deltasq);
        costh = 1.0 - 2.0 * deltasq / (1.0 + deltasq);
      } else {
        const double theta = kPI * u_polar;
        costh = std::cos(theta);
        sinth = std::sin(theta);
      }
      break;
    case 1:
      orc_nanbu_costh_sinth(s12, u_polar, &costh, &sinth);
      break;
    case 2:
      costh = 1.0 - std::min(s12, 2.0);
      sinth = std::sin(std::acos(costh));
      break;
    case 3:
    case 4: {
      FasDraws D = {{u_polar, g_fas_draw2, g_fas_draw3}, 0};
      if (angular == 3) nanbu_fas(s12, Clog, b0, bmin_qm, sigma_eff, D, costh, sinth);
      else nanbu_fas_v2(s12, Clog, b0, bmin_qm, sigma_eff, D, costh, sinth);
      break;
    }
    default: {
      const double theta = kPI * u_polar;
      costh = std::cos(theta);
      sinth = std::sin(theta);
    }
  }
  const double phi = kTWOPI * (g_fas_live ? mu_rand() : u_phi);   /* live: the azimuth follows the polar draws */
  orc_rotate_velocity(upst, costh, sinth, std::cos(phi), std::sin(phi));
  ucmdotup = gcm * (vcm[0] * upst[0] + vcm[1] * upst[1] + vcm[2] * upst[2]);
  upst_fact = gcm * (ucmdotup / (1.0 + gcm) + g1st);
  for (int n = 0; n < 3; n++) a_up1[n] = upst[n] + upst_fact * vcm[n];
  if (a_scatter2)
    for (int n = 0; n < 3; n++) a_up2[n] = (ptot[n] - a_mass1 * a_up1[n]) / a_mass2;
  return 1;
}

/* Coulomb weight_method: 0 PROBABILISTIC, 1 CONSERVATIVE (the Sentoku-Kemp update of applyIntra/InterScattering_SK08;
 * O(N) pairs only) */
static int g_sk08 = 0;
extern "C" void orc_coulomb_set_weight_method(int conservative) { g_sk08 = conservative ? 1 : 0; }
namespace {
/* one pair with the reference's draw order: polar draw(s) inside SetPolarScattering, then phi,
 * then the weight-rejection uniform (only for unequal weights) */
struct PairCtx {
  double charge1, charge2, mass1, mass2, EF_norm, Clog;
  int angular;
  double bmax, sigma_max, dt_sec;
};
void coulomb_pair(const PairCtx &c, double *b1, double w1, double *b2, double w2, double den12, double f1,
                  double f2) {
  if (orc_get_relativistic()) {
    /* Coulomb.cpp:548-559 / 1139-1150: the weight-rejection draw comes first, then LorentzScatter with the
     * lighter-weight particle in the first slot; draws inside it as in GalileanScatter (polar, then phi) */
    double *p1 = b1, *p2 = b2;
    double q1 = c.charge1, q2 = c.charge2, m1 = c.mass1, m2 = c.mass2;
    bool scatter2 = true;
    if ((float)w2 < (float)w1) {
      if (mu_rand() > w2 / w1) scatter2 = false;
      std::swap(p1, p2);
      std::swap(q1, q2);
      std::swap(m1, m2);
    } else if ((float)w1 < (float)w2) {
      if (mu_rand() > w1 / w2) scatter2 = false;
    }
    double t1[3] = {p1[0], p1[1], p1[2]}, t2[3] = {p2[0], p2[1], p2[2]}, s12 = 0.0;
    const int live = orc_coulomb_lorentz_scatter(t1, t2, 0, q1, q2, m1, m2, c.EF_norm, c.Clog, 2, den12, c.bmax,
                                                 c.sigma_max, c.dt_sec, 0.0, 0.0, 0.0, &s12);
    if (!live) return;
    const double saved_draw = g_large_angle_draw;
    if (g_large_angle) {   /* SetPolarScattering draws RL first; it decides whether a polar draw follows */
      g_large_angle_draw = mu_rand();
      double u1[3] = {p1[0], p1[1], p1[2]}, u2[3] = {p2[0], p2[1], p2[2]};
      orc_coulomb_lorentz_scatter(u1, u2, 0, q1, q2, m1, m2, c.EF_norm, c.Clog, 2, den12, c.bmax, c.sigma_max, c.dt_sec,
                                  0.0, 0.0, 0.0, &s12);
    }
    double gauss = 0.0, upol = 0.0;
    if (s12 < 0.0) {
      /* large-angle event (or too few collisions): no polar draw */
    } else if (c.angular == 0) {
      if (s12 < 2.0) gauss = mu_randn();
      else upol = mu_rand();
    } else if (c.angular == 3 || c.angular == 4) {
      g_fas_live = 1;          /* the polar uniforms and then the azimuth are drawn inside, in the reference's order */
    } else if (c.angular != 2) {
      upol = mu_rand();
    }
    const double uphi = g_fas_live ? 0.0 : mu_rand();
    orc_coulomb_lorentz_scatter(p1, p2, scatter2 ? 1 : 0, q1, q2, m1, m2, c.EF_norm, c.Clog, c.angular, den12,
                                c.bmax, c.sigma_max, c.dt_sec, gauss, upol, uphi, nullptr);
    g_fas_live = 0;
    g_large_angle_draw = saved_draw;
    return;
  }
  /* replicate which draws GalileanScatter makes: decide the branch from s12 first */
  double dU[3], s12 = 0.0;
  double probe[3];
  const int live = orc_coulomb_delta_u(b1, b2, c.charge1, c.charge2, c.mass1, c.mass2, c.EF_norm, c.Clog, 2,
                                       den12, c.bmax, c.sigma_max, c.dt_sec, 0.0, 0.0, 0.0, probe, &s12);
  if (live) {
    const double saved_draw = g_large_angle_draw;
    if (g_large_angle) {
      g_large_angle_draw = mu_rand();
      orc_coulomb_delta_u(b1, b2, c.charge1, c.charge2, c.mass1, c.mass2, c.EF_norm, c.Clog, 2, den12, c.bmax,
                          c.sigma_max, c.dt_sec, 0.0, 0.0, 0.0, probe, &s12);
    }
    double gauss = 0.0, upol = 0.0;
    if (s12 < 0.0) {
      /* large-angle event (or too few collisions): no polar draw */
    } else if (c.angular == 0) {
      if (s12 < 2.0) gauss = mu_randn();
      else upol = mu_rand();
    } else if (c.angular == 3 || c.angular == 4) {
      g_fas_live = 1;          /* the polar uniforms and then the azimuth are drawn inside, in the reference's order */
    } else if (c.angular != 2) {
      upol = mu_rand();
    }
    const double uphi = g_fas_live ? 0.0 : mu_rand();
    orc_coulomb_delta_u(b1, b2, c.charge1, c.charge2, c.mass1, c.mass2, c.EF_norm, c.Clog, c.angular, den12,
                        c.bmax, c.sigma_max, c.dt_sec, gauss, upol, uphi, dU, nullptr);
    g_fas_live = 0;
    g_large_angle_draw = saved_draw;
  } else {
    dU[0] = dU[1] = dU[2] = 0.0;
  }
  if (g_sk08 && (float)w1 != (float)w2) {
    /* weight_method = CONSERVATIVE: Coulomb::applyIntra/InterScattering_SK08 (Coulomb.cpp:849-897, 1575-1621) with
     * Coulomb::enforceEnergyConservation (Coulomb.H:796-823) */
    const bool first_light = (float)w1 < (float)w2;
    double *vl = first_light ? b1 : b2, *vh = first_light ? b2 : b1;
    const double fl = first_light ? f1 : -f2, fh = first_light ? -f2 : f1;
    const double mh = first_light ? c.mass2 : c.mass1;
    const long double ratio = first_light ? w1 / w2 : w2 / w1;
    double before[3] = {vh[0], vh[1], vh[2]}, vhp[3];
    const double Ebefore = mh * (vh[0] * vh[0] + vh[1] * vh[1] + vh[2] * vh[2]) / 2.0;
    for (int n = 0; n < 3; n++) vl[n] += fl * dU[n];
    for (int n = 0; n < 3; n++) vhp[n] = vh[n] + fh * dU[n];
    const double Escatter = mh * (vhp[0] * vhp[0] + vhp[1] * vhp[1] + vhp[2] * vhp[2]) / 2.0;
    const double Eafter = Ebefore + ratio * (Escatter - Ebefore);
    for (int n = 0; n < 3; n++) vh[n] = before[n] + ratio * (vhp[n] - before[n]);
    double betap_r = vh[0] * vh[0] + vh[1] * vh[1];
    double betap_mag = betap_r + vh[2] * vh[2];
    const double Eafter2 = mh / 2.0 * betap_mag;
    betap_mag = std::sqrt(betap_mag);
    betap_r = std::sqrt(betap_r);
    if (Eafter < Eafter2) return;
    const double dmag = std::sqrt(2.0 / mh * (Eafter - Eafter2));
    const double phi = kTWOPI * mu_rand();
    const double cosphi = cos(phi), sinphi = sin(phi);
    double d[3];
    d[0] = (vh[2] * vh[0] * cosphi - betap_mag * vh[1] * sinphi) / betap_r * dmag / betap_mag;
    d[1] = (vh[2] * vh[1] * cosphi + betap_mag * vh[0] * sinphi) / betap_r * dmag / betap_mag;
    d[2] = -betap_r * cosphi * dmag / betap_mag;
    for (int n = 0; n < 3; n++) vh[n] += d[n];
    return;
  }
  if ((float)w1 == (float)w2) {
    for (int n = 0; n < 3; n++) b1[n] += f1 * dU[n];
    for (int n = 0; n < 3; n++) b2[n] -= f2 * dU[n];
  } else if ((float)w1 < (float)w2) {
    for (int n = 0; n < 3; n++) b1[n] += f1 * dU[n];
    if (mu_rand() < w1 / w2)
      for (int n = 0; n < 3; n++) b2[n] -= f2 * dU[n];
  } else {
    if (mu_rand() < w2 / w1)
      for (int n = 0; n < 3; n++) b1[n] += f1 * dU[n];
    for (int n = 0; n < 3; n++) b2[n] -= f2 * dU[n];
  }
}
}  // namespace

/* ScatteringUtils::modEnergyPairwise (ScatteringUtils.H:113-205): zero-angle inelastic "collision" of a pair that removes
 * (a_deltaE > 0) or adds (< 0) up to Erel_frac of the pair's relative energy without touching its momentum.  The scalars
 * are long double as in the reference.  rel selects the RELATIVISTIC_PARTICLES branch. */
static void mod_energy_pairwise(double *b1, double *b2, double wpmp1, double wpmp2, double Erel_frac, double &Erel_cumm,
                                long double &a_deltaE, int rel) {
  int sign = 1;
  if (a_deltaE < 0.0) sign = -1;
  const double ux = b1[0] - b2[0], uy = b1[1] - b2[1], uz = b1[2] - b2[2];
  const long double usq = ux * ux + uy * uy + uz * uz;
  long double Erel, muR = 0.0, E1 = 0.0, E2 = 0.0, Etot = 0.0, pxtot = 0.0, pytot = 0.0, pztot = 0.0;
  if (rel) {
    long double gbsq1 = 0.0, gbsq2 = 0.0;
    for (int n = 0; n < 3; n++) gbsq1 += b1[n] * b1[n];
    for (int n = 0; n < 3; n++) gbsq2 += b2[n] * b2[n];
    const long double gamma1 = std::sqrt(1.0 + gbsq1), gamma2 = std::sqrt(1.0 + gbsq2);
    E1 = wpmp1 * gamma1;
    E2 = wpmp2 * gamma2;
    Etot = E1 + E2;
    pxtot = wpmp1 * b1[0] + wpmp2 * b2[0];
    pytot = wpmp1 * b1[1] + wpmp2 * b2[1];
    pztot = wpmp1 * b1[2] + wpmp2 * b2[2];
    const long double Ecm = std::sqrt(Etot * Etot - pxtot * pxtot - pytot * pytot - pztot * pztot);
    Erel = Ecm - wpmp1 - wpmp2;
  } else {
    muR = wpmp1 * wpmp2 / (wpmp1 + wpmp2);
    Erel = muR / 2.0 * usq;
  }
  if (Erel <= 0.0) return;
  long double deltaE = sign * Erel_frac * Erel;
  if (std::abs(deltaE) > std::abs(a_deltaE)) {
    deltaE = a_deltaE;
    a_deltaE = 0.0;
  } else {
    a_deltaE -= deltaE;
  }
  Erel_cumm += Erel - deltaE;
  if (rel) {
    const long double A = Etot - deltaE;
    const long double D = A * A + E2 * E2 - E1 * E1;
    const long double p2dotu = wpmp2 * (b2[0] * ux + b2[1] * uy + b2[2] * uz);
    const long double ptdotu = pxtot * ux + pytot * uy + pztot * uz;
    const long double a = A * A * usq - ptdotu * ptdotu;
    const long double b = D * ptdotu - 2 * A * A * p2dotu;
    const long double c = A * A * E2 * E2 - D * D / 4.0;
    const long double root = b * b - 4.0 * a * c;
    if (root < 0.0 || a == 0.0) return;
    const long double alpha = (-b + std::sqrt(root)) / (2.0 * a);
    const long double ratio1 = alpha / wpmp1, ratio2 = alpha / wpmp2;
    b1[0] += ratio1 * ux; b1[1] += ratio1 * uy; b1[2] += ratio1 * uz;
    b2[0] -= ratio2 * ux; b2[1] -= ratio2 * uy; b2[2] -= ratio2 * uz;
  } else {
    const long double uprime_over_u = std::sqrt(1.0 - deltaE / Erel);
    double deltaU[3];
    deltaU[0] = uprime_over_u * ux - ux;
    deltaU[1] = uprime_over_u * uy - uy;
    deltaU[2] = uprime_over_u * uz - uz;
    for (int n = 0; n < 3; n++) {
      b1[n] += muR / wpmp1 * deltaU[n];
      b2[n] -= muR / wpmp2 * deltaU[n];
    }
  }
}
extern "C" void orc_mod_energy_pairwise(double *b1, double *b2, double wpmp1, double wpmp2, double Erel_frac,
                                        double *Erel_cumm, double *deltaE, int rel) {
  long double dE = *deltaE;
  mod_energy_pairwise(b1, b2, wpmp1, wpmp2, Erel_frac, *Erel_cumm, dE, rel);
  *deltaE = (double)dE;
}

/* scattering.coulomb.enforce_conservations and its companions (Coulomb.H:286-293, defaults :32-33, :325-333) */
struct EnforcePrm {
  int on;
  double energy_fraction, energy_fraction_max;
  int beta_weight_exponent, sort_weighted, nmin_save;
};
static EnforcePrm g_enf = {0, 0.05, 0.5, 1, 0, 100000};
extern "C" void orc_coulomb_set_enforce(int on, double energy_fraction, double energy_fraction_max, int beta_weight_exponent,
                                        int sort_weighted, int nmin_save) {
  g_enf = {on, energy_fraction, energy_fraction_max, beta_weight_exponent, sort_weighted, nmin_save};
}
namespace {
struct CellSums {
  double W;
  double p[3];
  long double E;
};
/* Wtot0 / ptot0 / Etot0 of a cell list (Coulomb.cpp:486-512; the Galilean Efact = 1/2) */
CellSums cell_sums(const std::vector<long> &idx, const double *v, const double *w, long n) {
  CellSums s = {0.0, {0.0, 0.0, 0.0}, 0.0};
  for (long i : idx) {
    s.W += std::pow(w[i], g_enf.beta_weight_exponent);
    double gbsq = 0.0;
    for (int q = 0; q < 3; ++q) {
      s.p[q] += w[i] * v[q * n + i];
      gbsq += v[q * n + i] * v[q * n + i];
    }
    s.E += w[i] * 0.5 * gbsq;
  }
  return s;
}
/* the pair sweep that absorbs deltaE inside one list (Coulomb.cpp:643-706 and :1257-1300): returns false if the
 * correction failed (energy_frac_eff > energy_fraction_max or more than ten sweeps) */
bool absorb_energy(std::vector<long> order, double *v, const double *w, long n, double mass, long double &deltaE,
                   long *count) {
  if (g_enf.sort_weighted)
    std::sort(order.begin(), order.end(), [&](long a, long b) { return w[a] > w[b]; });
  const int N = (int)order.size();
  int loop_count = 0;
  double Erel_cumm = 0.0, fmult_fact = 1.0;
  for (int p = 0; p < N; p++) {
    if (deltaE == 0.0) break;
    const int p1 = p;
    p++;
    if (p == N) {
      loop_count++;
      p = 0;
    }
    const int p2 = p;
    const long i1 = order[p1], i2 = order[p2];
    double a[3] = {v[i1], v[n + i1], v[2 * n + i1]}, b[3] = {v[i2], v[n + i2], v[2 * n + i2]};
    mod_energy_pairwise(a, b, mass * w[i1], mass * w[i2], g_enf.energy_fraction * fmult_fact, Erel_cumm, deltaE, 0);
    for (int q = 0; q < 3; ++q) {
      v[q * n + i1] = a[q];
      v[q * n + i2] = b[q];
    }
    if (count) ++*count;
    if (deltaE == 0.0) break;
    if (p == N - 1) {
      loop_count++;
      const double energy_frac_eff = (double)std::abs(deltaE) / Erel_cumm;
      if (energy_frac_eff > g_enf.energy_fraction_max || loop_count > 10) return false;
      else if (energy_frac_eff > g_enf.energy_fraction) fmult_fact = energy_frac_eff / g_enf.energy_fraction;
      Erel_cumm = 0.0;
      p = -1;
    }
  }
  return true;
}
}  // namespace

/* Coulomb::applyIntraScattering_PROB (Coulomb.cpp:400-728); enforce_conservations per orc_coulomb_set_enforce */
extern "C" void orc_coulomb_intra(long ncell, const long *cell_start, double *v, const double *w, long n,
                                  const double *dens, const double *LDe, double cellV_SI, double mass, double charge,
                                  double Clog, int angular, int NxN_in, int NxN_Nthresh, double dt_sec,
                                  long *npairs_out) {
  const CoulombConsts k = coulomb_consts(charge, charge, mass, mass);
  long npairs = 0;
  std::vector<long> idx;
  for (long c = 0; c < ncell; ++c) {
    const double numDen = dens[c];
    if (numDen == 0.0) continue;
    const double atomic_spacing = 1.0 / std::cbrt(4.0 / 3.0 * kPI * numDen);
    PairCtx ctx = {charge, charge, mass, mass, k.EF_fact * std::pow(numDen, 2.0 / 3.0), Clog, angular,
                   LDe[c], 1.0 / (numDen * atomic_spacing), dt_sec};
    const long numCell = cell_start[c + 1] - cell_start[c];
    if (numCell < 2) continue;
    bool NxN = NxN_in != 0;
    if (numCell < NxN_Nthresh) NxN = true;
    if (g_sk08) NxN = false;   /* _SK08 pairs in O(N) only (its first three pairs of an odd cell are the same three) */
    bool odd_NxN = false;
    if (!NxN && numCell % 2 == 1) odd_NxN = true;
    const long Naa = numCell - 1;
    idx.resize(numCell);
    for (long q = 0; q < numCell; ++q) idx[q] = cell_start[c] + q;
    std::shuffle(idx.begin(), idx.end(), global_rand_gen);
    CellSums s0 = {0.0, {0.0, 0.0, 0.0}, 0.0};
    std::vector<double> vsave;
    if (g_enf.on) {   /* :490-512 */
      s0 = cell_sums(idx, v, w, n);
      if (numCell <= g_enf.nmin_save)
        for (long i : idx)
          for (int q = 0; q < 3; ++q) vsave.push_back(v[q * n + i]);
    }
    const long p1_max = numCell - 2;
    for (long p1 = 0; p1 <= p1_max; p1++) {
      long p2_max = p1 + 1;
      if (NxN) p2_max = numCell - 1;
      else if (odd_NxN) p2_max = 2;
      for (long p2 = p1 + 1; p2 <= p2_max; p2++) {
        const long i1 = idx[p1], i2 = idx[p2];
        const double wpMax = std::max(w[i1], w[i2]);
        double den12;
        if (NxN) den12 = wpMax / cellV_SI;
        else if (odd_NxN) den12 = wpMax * Naa / cellV_SI / 2.0;
        else den12 = wpMax * Naa / cellV_SI;
        double a[3] = {v[i1], v[n + i1], v[2 * n + i1]}, b[3] = {v[i2], v[n + i2], v[2 * n + i2]};
        coulomb_pair(ctx, a, w[i1], b, w[i2], den12, 0.5, 0.5);
        for (int q = 0; q < 3; ++q) {
          v[q * n + i1] = a[q];
          v[q * n + i2] = b[q];
        }
        ++npairs;
      }
      if (odd_NxN && p1 == 1) odd_NxN = false;
      if (!odd_NxN && !NxN) ++p1;
    }
    if (g_enf.on) {   /* :611-711.  dBetaAvg = the cell's weighted momentum change, taken from the sums as the
                         RELATIVISTIC_PARTICLES build does (:596-609); the Galilean build accumulates the same number pair by pair */
      const CellSums s1 = cell_sums(idx, v, w, n);
      double dBetaAvg[3], dBetaSq = 0.0;
      for (int q = 0; q < 3; ++q) {
        dBetaAvg[q] = s1.p[q] - s0.p[q];
        dBetaSq += dBetaAvg[q] * dBetaAvg[q];
      }
      if (dBetaSq > 0.0) {
        for (int q = 0; q < 3; ++q) dBetaAvg[q] /= s0.W;
        long double Etot1 = 0.0;
        for (long i : idx) {
          double gbsq = 0.0;
          for (int q = 0; q < 3; ++q) {
            v[q * n + i] -= std::pow(w[i], g_enf.beta_weight_exponent - 1) * dBetaAvg[q];
            gbsq += v[q * n + i] * v[q * n + i];
          }
          Etot1 += w[i] * 0.5 * gbsq;
        }
        long double deltaE = mass * (Etot1 - s0.E);
        if (!absorb_energy(idx, v, w, n, mass, deltaE, nullptr) && numCell <= g_enf.nmin_save) {
          long k = 0;
          for (long i : idx)
            for (int q = 0; q < 3; ++q) v[q * n + i] = vsave[k++];
        }
      }
    }
  }
  if (npairs_out) *npairs_out = npairs;
}

/* Coulomb::applyInterScattering_PROB (Coulomb.cpp:919-1438); enforce_conservations per orc_coulomb_set_enforce */
extern "C" void orc_coulomb_inter(long ncell, const long *cs1, double *v1, const double *w1, long n1,
                                  const double *dens1, double mass1, double charge1, const long *cs2, double *v2,
                                  const double *w2, long n2, const double *dens2, double mass2, double charge2,
                                  const double *LDe, double cellV_SI, double Clog, int angular, int NxN_in,
                                  int NxN_Nthresh, double dt_sec, long *npairs_out) {
  const CoulombConsts k = coulomb_consts(charge1, charge2, mass1, mass2);
  long npairs = 0;
  std::vector<long> idx1, idx2;
  for (long c = 0; c < ncell; ++c) {
    const double numDen1 = dens1[c], numDen2 = dens2[c];
    if (numDen1 * numDen2 == 0.0) continue;
    const double minn = std::min(numDen1, numDen2), maxn = std::max(numDen1, numDen2);
    const double atomic_spacing = 1.0 / std::cbrt(4.0 / 3.0 * kPI * minn);
    PairCtx ctx = {charge1, charge2, mass1, mass2, k.EF_fact * std::pow(maxn, 2.0 / 3.0), Clog, angular,
                   LDe[c], 1.0 / (minn * atomic_spacing), dt_sec};
    const long numCell1 = cs1[c + 1] - cs1[c], numCell2 = cs2[c + 1] - cs2[c];
    if (numCell1 * numCell2 < 2) continue;
    const long Nmin = std::min(numCell1, numCell2), Nmax = std::max(numCell1, numCell2);
    bool NxN = NxN_in != 0;
    if (Nmin < NxN_Nthresh) NxN = true;
    if (g_sk08) NxN = false;
    idx1.resize(numCell1);
    for (long q = 0; q < numCell1; ++q) idx1[q] = cs1[c] + q;
    std::shuffle(idx1.begin(), idx1.end(), global_rand_gen);
    idx2.resize(numCell2);
    for (long q = 0; q < numCell2; ++q) idx2[q] = cs2[c] + q;
    std::shuffle(idx2.begin(), idx2.end(), global_rand_gen);
    CellSums s01 = {0.0, {0.0, 0.0, 0.0}, 0.0}, s02 = s01;
    double wp1_mean = 0.0, wp2_mean = 0.0;
    std::vector<double> vsave1, vsave2;
    if (g_enf.on) {   /* :1024-1083 */
      s01 = cell_sums(idx1, v1, w1, n1);
      s02 = cell_sums(idx2, v2, w2, n2);
      for (long i : idx1) wp1_mean += w1[i];
      for (long i : idx2) wp2_mean += w2[i];
      wp1_mean /= numCell1;
      wp2_mean /= numCell2;
      if (numCell1 <= g_enf.nmin_save || numCell2 <= g_enf.nmin_save) {
        for (long i : idx1)
          for (int q = 0; q < 3; ++q) vsave1.push_back(v1[q * n1 + i]);
        for (long i : idx2)
          for (int q = 0; q < 3; ++q) vsave2.push_back(v2[q * n2 + i]);
      }
    }
    for (long p = 0; p < Nmax; p++) {
      long p1, p2, pmin_start;
      if (Nmin == numCell1) {
        p1 = p % numCell1;
        p2 = p;
        pmin_start = p1;
      } else {
        p1 = p;
        p2 = p % numCell2;
        pmin_start = p2;
      }
      long pmin_end = pmin_start;
      if (NxN) {
        pmin_start = 0;
        pmin_end = Nmin - 1;
      }
      for (long pmin = pmin_start; pmin <= pmin_end; pmin++) {
        if (NxN) {
          if (Nmin == numCell1) p1 = pmin;
          else p2 = pmin;
        }
        const long i1 = idx1[p1], i2 = idx2[p2];
        const double wpMax = std::max(w1[i1], w2[i2]);
        const double den12 = NxN ? wpMax / cellV_SI : wpMax * Nmin / cellV_SI;
        double a[3] = {v1[i1], v1[n1 + i1], v1[2 * n1 + i1]}, b[3] = {v2[i2], v2[n2 + i2], v2[2 * n2 + i2]};
        coulomb_pair(ctx, a, w1[i1], b, w2[i2], den12, k.mu / mass1, k.mu / mass2);
        for (int q = 0; q < 3; ++q) {
          v1[q * n1 + i1] = a[q];
          v2[q * n2 + i2] = b[q];
        }
        ++npairs;
      }
    }
    if (g_enf.on) {   /* :1182-1430 */
      const double Wtot0 = mass1 * s01.W + mass2 * s02.W;
      const long double Etot0 = mass1 * s01.E + mass2 * s02.E;
      const CellSums s11 = cell_sums(idx1, v1, w1, n1), s12 = cell_sums(idx2, v2, w2, n2);
      double dBetaAvg[3], dBetaSq = 0.0;
      for (int q = 0; q < 3; ++q) {
        dBetaAvg[q] = (mass1 * s11.p[q] + mass2 * s12.p[q]) - (mass1 * s01.p[q] + mass2 * s02.p[q]);
        dBetaSq += dBetaAvg[q] * dBetaAvg[q];
      }
      if (dBetaSq > 0.0) {
        for (int q = 0; q < 3; ++q) dBetaAvg[q] /= Wtot0;
        long double Etot11 = 0.0, Etot12 = 0.0;
        for (long i : idx1) {
          double gbsq = 0.0;
          for (int q = 0; q < 3; ++q) {
            v1[q * n1 + i] -= std::pow(w1[i], g_enf.beta_weight_exponent - 1) * dBetaAvg[q];
            gbsq += v1[q * n1 + i] * v1[q * n1 + i];
          }
          Etot11 += w1[i] * 0.5 * gbsq;
        }
        Etot11 *= mass1;
        for (long i : idx2) {
          double gbsq = 0.0;
          for (int q = 0; q < 3; ++q) {
            v2[q * n2 + i] -= std::pow(w2[i], g_enf.beta_weight_exponent - 1) * dBetaAvg[q];
            gbsq += v2[q * n2 + i] * v2[q * n2 + i];
          }
          Etot12 += w2[i] * 0.5 * gbsq;
        }
        Etot12 *= mass2;
        const long double deltaE = (Etot11 + Etot12) - Etot0;
        const long double Etotdenom = wp1_mean * Etot11 + wp2_mean * Etot12;
        long double deltaEp1, deltaEp2;
        if (numCell1 == 1) {
          deltaEp1 = 0.0;
          deltaEp2 = deltaE;
        } else if (numCell2 == 1) {
          deltaEp1 = deltaE;
          deltaEp2 = 0.0;
        } else {
          deltaEp1 = wp1_mean * Etot11 / Etotdenom * deltaE;
          deltaEp2 = wp2_mean * Etot12 / Etotdenom * deltaE;
        }
        bool ok = absorb_energy(idx1, v1, w1, n1, mass1, deltaEp1, nullptr);
        if (ok) ok = absorb_energy(idx2, v2, w2, n2, mass2, deltaEp2, nullptr);
        if (!ok && (numCell1 <= g_enf.nmin_save || numCell2 <= g_enf.nmin_save)) {
          long k = 0;
          for (long i : idx1)
            for (int q = 0; q < 3; ++q) v1[q * n1 + i] = vsave1[k++];
          k = 0;
          for (long i : idx2)
            for (int q = 0; q < 3; ++q) v2[q * n2 + i] = vsave2[k++];
        }
      }
    }
  }
  if (npairs_out) *npairs_out = npairs;
}

/* Elastic::getSigma / getTextSigma (Elastic.cpp:390-476): const sigma (ntab == 0) or a table
 * (E [eV] ascending, Q, xi) with the reference's interpolation; angular: 0 ISOTROPIC (Q = Qelm),
 * 1 OKHRIMOVSKYY (Q = Qela).  The interpolation formulas, including which end they weight, are
 * the reference's (MathUtils::linearInterp, ScatteringUtils::semilogInterp/loglogInterp). */
extern "C" double orc_elastic_sigma(double g12, double mu, double const_sigma, int ntab, const double *E,
                                    const double *Q, const double *XI, int angular, int loglog, double *xi_out) {
  *xi_out = 0.0;
  if (ntab == 0) return const_sigma;
  const double mcSq = kME * kCVAC * kCVAC / kQE;
  const double KE = mu * mcSq * g12 * g12 / 2.0;
  double sigma = 0.0, xi = 0.0;
  auto lin = [&](const double *Y, int i) {   /* MathUtils::linearInterp (MathUtils.cpp:134-148) */
    return (Y[i + 1] * (KE - E[i]) + Y[i] * (E[i + 1] - KE)) / (E[i + 1] - E[i]);
  };
  auto loglogI = [&](const double *Y, int i) {
    const double l0 = log10(KE), lu = log10(E[i]), ld = log10(E[i + 1]);
    return pow(10.0, (log10(Y[i + 1]) * (l0 - ld) + log10(Y[i]) * (lu - l0)) / (lu - ld));
  };
  auto semilogI = [&](const double *Y, int i) {
    const double l0 = log10(KE), lu = log10(E[i]), ld = log10(E[i + 1]);
    return (Y[i + 1] * (l0 - ld) + Y[i] * (lu - l0)) / (lu - ld);
  };
  if (KE >= E[ntab - 1]) {
    if (angular == 0) sigma = Q[ntab - 1] * log(KE) / log(E[ntab - 1]) * E[ntab - 1] / KE;
    else {
      sigma = Q[ntab - 1] * E[ntab - 1] / KE;
      xi = XI[ntab - 1];
    }
  } else {
    int index = ntab / 2;
    while (KE < E[index]) index--;
    while (KE > E[index + 1]) index++;
    if (loglog && Q[index] * E[index] > 0.0) sigma = loglogI(Q, index);
    else sigma = lin(Q, index);
    if (angular == 1) {
      if (E[index] * KE > 0.0) xi = semilogI(XI, index);
      else xi = lin(XI, index);
    }
  }
  *xi_out = xi;
  return sigma;
}

/* Elastic::electronImpact (Elastic.cpp:225-388), PROBABILISTIC weight method */
extern "C" void orc_elastic_wm(long ncell, const long *cs1, double *v1, const double *w1, long n1, double mass1,
                               const long *cs2, double *v2, double *w2, long n2, const double *dens2, double mass2,
                               double const_sigma, int ntab, const double *E, const double *Q, const double *XI,
                               int angular, int loglog, int conservative, double dt_sec, long *ncoll_out);
extern "C" void orc_elastic(long ncell, const long *cs1, double *v1, const double *w1, long n1, double mass1,
                            const long *cs2, double *v2, const double *w2, long n2, const double *dens2,
                            double mass2, double const_sigma, int ntab, const double *E, const double *Q,
                            const double *XI, int angular, int loglog, double dt_sec, long *ncoll_out) {
  orc_elastic_wm(ncell, cs1, v1, w1, n1, mass1, cs2, v2, const_cast<double *>(w2), n2, dens2, mass2, const_sigma, ntab, E,
                 Q, XI, angular, loglog, 0, dt_sec, ncoll_out);
}
/* conservative != 0: weight_method = CONSERVATIVE (Elastic.cpp:333-358): when the projectile is the lighter one
 * (wp1 < wp2) it scatters, and the target, its scattered fraction and a second target of the cell are merged into two
 * equally weighted particles (collapseThreeToTwo); w2 changes.  Oracle only: the device has no such branch yet. */
extern "C" void orc_elastic_wm(long ncell, const long *cs1, double *v1, const double *w1, long n1, double mass1,
                               const long *cs2, double *v2, double *w2, long n2, const double *dens2, double mass2,
                               double const_sigma, int ntab, const double *E, const double *Q, const double *XI,
                               int angular, int loglog, int conservative, double dt_sec, long *ncoll_out) {
  const double mu = mass1 * mass2 / (mass1 + mass2);
  long ncoll = 0;
  for (long c = 0; c < ncell; ++c) {
    const long numCell1 = cs1[c + 1] - cs1[c], numCell2 = cs2[c + 1] - cs2[c];
    if (numCell1 < 1 || numCell2 < 1) continue;
    for (long q = 0; q < numCell1; ++q) {
      const long i1 = cs1[c] + q;
      std::uniform_int_distribution<> pick(0, (int)numCell2 - 1);   /* MathUtils::randInt */
      const long i2 = cs2[c] + pick(global_rand_gen);
      double a[3] = {v1[i1], v1[n1 + i1], v1[2 * n1 + i1]}, b[3] = {v2[i2], v2[n2 + i2], v2[2 * n2 + i2]};
      double g12 = 0.0;
      for (int d = 0; d < 3; ++d) g12 += pow(a[d] - b[d], 2);
      g12 = sqrt(g12);
      double xi;
      const double sigma = orc_elastic_sigma(g12, mu, const_sigma, ntab, E, Q, XI, angular, loglog, &xi);
      if (sigma == 0.0) continue;
      const double arg = g12 * kCVAC * sigma * dens2[c] * dt_sec;
      const double q12 = 1.0 - exp(-arg);
      if (mu_rand() <= q12) {
        ++ncoll;
        const double phi = kTWOPI * mu_rand();
        const double R = mu_rand();
        const double costh = 1.0 - 2.0 * R * (1.0 - xi) / (1.0 + xi * (1.0 - 2.0 * R));
        const double sinth = sqrt(1.0 - costh * costh);
        double dU[3];
        orc_scatter_delta_u(a[0] - b[0], a[1] - b[1], a[2] - b[2], costh, sinth, cos(phi), sin(phi), dU);
        if (conservative && w1[i1] < w2[i2]) {
          if (numCell2 < 2) continue;
          double b2p[3];
          for (int d = 0; d < 3; ++d) {
            a[d] += mu / mass1 * dU[d];
            b2p[d] = b[d] - mu / mass2 * dU[d];
          }
          long i3 = cs2[c] + pick(global_rand_gen);
          while (i3 == i2) i3 = cs2[c] + pick(global_rand_gen);
          double b3[3] = {v2[i3], v2[n2 + i3], v2[2 * n2 + i3]};
          double wq2 = w2[i2], wq3 = w2[i3];
          orc_collapse_three_to_two(b, &wq2, b3, &wq3, b2p, w1[i1]);
          w2[i2] = wq2;
          w2[i3] = wq3;
          for (int d = 0; d < 3; ++d) {
            v1[d * n1 + i1] = a[d];
            v2[d * n2 + i2] = b[d];
            v2[d * n2 + i3] = b3[d];
          }
          continue;
        }
        const double r2 = mu_rand();
        if (r2 <= w2[i2] / w1[i1])
          for (int d = 0; d < 3; ++d) v1[d * n1 + i1] = a[d] + mu / mass1 * dU[d];
        if (r2 <= w1[i1] / w2[i2])
          for (int d = 0; d < 3; ++d) v2[d * n2 + i2] = b[d] - mu / mass2 * dU[d];
      }
    }
  }
  if (ncoll_out) *ncoll_out = ncoll;
}

/* =====================================================================================
 * Scattering::setMeanFreeTime -- the per-cell collision frequency whose box maximum sets
 * m_scatter_dt = 1/nu_max (the MPI MAX over ranks is commented out in the reference, the
 * caller min-reduces the resulting dt: ScatteringInterface.cpp:324-345).  Inputs are the
 * cell moments of set{Number,Momentum,Energy}DensityFromBinFab: dens[c], mom[k*ncell+c],
 * ene[k*ncell+c], and the Debye length LDe[c].  Returned: nu_max [Hz] (0 if no cell counts).
 * ===================================================================================== */
namespace {
/* MathUtils::gammainc (MathUtils.cpp:65-95): Taylor series for a = 3/2, x < 10 */
double mu_gammainc_3half(double x) {
  double soln = std::tgamma(1.5);
  if (x < 10.0) {
    const int pmax = 41;
    double coef, sign = -1.0, factorial = 1;
    soln = 0.0;
    for (int p = 1; p < pmax; p++) {
      coef = 0.5 + p;
      if (p > 1) factorial = factorial * (p - 1);
      sign = -sign;
      soln = soln + sign * pow(x, coef) / coef / factorial;
    }
  }
  return soln;
}
const double kEV_PER_JOULE = 1.0 / kQE;
const double kM3_PER_CM3 = 1.0 / 1.0e+06;
}  // namespace

extern "C" double orc_gammainc_3half(double x) { return mu_gammainc_3half(x); }

/* TakizukaAbe::setIntraMFT / setInterMFT (TakizukaAbe.cpp:80-238) */
extern "C" double orc_ta_nu_max(long ncell, const double *dens1, const double *ene1, const double *dens2,
                                const double *ene2, double charge1, double charge2, double mass1, double mass2,
                                double Clog, int intra) {
  const double cvacSq = kCVAC * kCVAC;
  double box_nuMax = 0.0;
  if (intra) {
    for (long c = 0; c < ncell; ++c) {
      const double numberDensity = dens1[c];
      if (numberDensity == 0.0) continue;
      double energyDensity = 0.0;
      for (int dir = 0; dir < 3; dir++) energyDensity = energyDensity + ene1[dir * ncell + c];
      double Teff_eV = kME * 2.0 / 3.0 * energyDensity / numberDensity * cvacSq;
      Teff_eV = kEV_PER_JOULE * Teff_eV;
      double tau = 3.44e5 * pow(Teff_eV, 1.5) / (numberDensity * kM3_PER_CM3) / Clog;
      tau = tau * sqrt(mass1 / 2.0) / pow(charge1 * charge2, 2);
      box_nuMax = std::max(box_nuMax, 1.0 / tau);
    }
    return box_nuMax;
  }
  const double factor = pow(kQE * charge1 * kQE * charge2 / kEP0, 2) / kFOURPI;
  for (long c = 0; c < ncell; ++c) {
    const double numberDensity1 = dens1[c], numberDensity2 = dens2[c];
    if (numberDensity1 * numberDensity2 == 0.0) continue;
    double energyDensity1 = 0.0, energyDensity2 = 0.0;
    for (int dir = 0; dir < 3; dir++) {
      energyDensity1 = energyDensity1 + ene1[dir * ncell + c];
      energyDensity2 = energyDensity2 + ene2[dir * ncell + c];
    }
    const double energy1 = kME * energyDensity1 / numberDensity1 * cvacSq;
    const double energy2 = kME * energyDensity2 / numberDensity2 * cvacSq;
    const double Teff1_eV = kEV_PER_JOULE * 2.0 / 3.0 * energy1;
    const double Teff2_eV = kEV_PER_JOULE * 2.0 / 3.0 * energy2;
    const double VT1 = sqrt(kQE * Teff1_eV / (kME * mass1));
    const double VT2 = sqrt(kQE * Teff2_eV / (kME * mass2));
    const double x12 = (Teff1_eV / mass1) / (Teff2_eV / mass2);
    const double x21 = 1. / x12;
    const double psi12 = 2.0 / sqrt(kPI) * mu_gammainc_3half(x12);
    const double psi21 = 2.0 / sqrt(kPI) * mu_gammainc_3half(x21);
    const double nu012 = factor * Clog * numberDensity2 / (pow(energy1, 2)) * VT1;
    const double nu021 = factor * Clog * numberDensity1 / (pow(energy2, 2)) * VT2;
    const double nu12 = (1.0 + mass1 / mass2) * psi12 * nu012;
    const double nu21 = (1.0 + mass2 / mass1) * psi21 * nu021;
    box_nuMax = std::max(box_nuMax, nu12);
    box_nuMax = std::max(box_nuMax, nu21);
  }
  return box_nuMax;
}

/* ==========================================================================================
 * HardSphere, PROBABILISTIC weight method (src/scattering/HardSphere.cpp:30-52, 223-418, 419-665):
 * no-time-counter pair selection.  Per cell: gmax = 5 thermal speeds from the cell's energy density,
 * Nmax candidate pairs (fractional part by a draw), each accepted with probability g12/gmax, isotropic
 * scattering, velocity update with probability w_other/w_self.  ene = [3][ncell] energy densities as
 * set{Energy}DensityFromBinFab leaves them.
 * ======================================================================================== */
namespace {
int mu_randint(int a, int b) {   /* MathUtils::randInt */
  std::uniform_int_distribution<int> dist(a, b);
  return dist(global_rand_gen);
}
}  // namespace

extern "C" double orc_hs_sigmaT(double r1, double r2) { return kPI * (r1 + r2) * (r1 + r2); }   /* HardSphere.cpp:52 */

extern "C" void orc_hs_self_wm(long ncell, const long *cs, double *v, double *w, long n, const double *dens,
                               const double *ene, double mass, double sigmaT, int conservative, double dt_sec,
                               long *ncand_out, long *ncoll_out);
extern "C" void orc_hs_self(long ncell, const long *cs, double *v, const double *w, long n, const double *dens,
                            const double *ene, double mass, double sigmaT, double dt_sec, long *ncand_out,
                            long *ncoll_out) {
  orc_hs_self_wm(ncell, cs, v, const_cast<double *>(w), n, dens, ene, mass, sigmaT, 0, dt_sec, ncand_out, ncoll_out);
}
/* conservative != 0: the CONSERVATIVE weight method (HardSphere.cpp:357-392): for unequal weights the lighter
 * particle scatters and the heavier one, its scattered fraction and a third particle of the cell are merged into two
 * equally weighted particles (collapseThreeToTwo); the weights w change. */
extern "C" void orc_hs_self_wm(long ncell, const long *cs, double *v, double *w, long n, const double *dens,
                               const double *ene, double mass, double sigmaT, int conservative, double dt_sec,
                               long *ncand_out, long *ncoll_out) {
  const double cvacSq = kCVAC * kCVAC;
  long ncand = 0, ncoll = 0;
  for (long c = 0; c < ncell; ++c) {
    const double local_numberDensity = dens[c];
    if (local_numberDensity == 0.0) continue;
    double local_energyDensity = 0.0;
    for (int dir = 0; dir < 3; dir++) local_energyDensity = local_energyDensity + ene[dir * ncell + c];
    const double local_Teff = 2.0 / 3.0 * local_energyDensity / local_numberDensity * cvacSq;
    const double local_gmax = 5.0 * sqrt(local_Teff / mass);
    const double local_nuMaxDt = local_numberDensity * sigmaT * local_gmax * dt_sec;
    const int local_numCell = (int)(cs[c + 1] - cs[c]);
    if (local_numCell < 2) continue;
    const double local_Nmax = 0.5 * (local_numCell - 1) * std::min(local_nuMaxDt, 1.0);
    double local_Nmax_whole;
    const double local_Nmax_remainder = modf(local_Nmax, &local_Nmax_whole);
    const double rand_num = mu_rand();
    int local_Nmax_integer = static_cast<int>(local_Nmax_whole);
    if (rand_num <= local_Nmax_remainder) local_Nmax_integer = local_Nmax_integer + 1;
    ncand += local_Nmax_integer;
    for (int k = 0; k < local_Nmax_integer; k++) {
      const int random_index1 = mu_randint(0, local_numCell - 1);
      int random_index2 = mu_randint(0, local_numCell - 1);
      while (random_index2 == random_index1) random_index2 = mu_randint(0, local_numCell - 1);
      const long i1 = cs[c] + random_index1, i2 = cs[c] + random_index2;
      double b1[3] = {v[i1], v[n + i1], v[2 * n + i1]}, b2[3] = {v[i2], v[n + i2], v[2 * n + i2]};
      const double wp1 = w[i1], wp2 = w[i2];
      double g12 = 0.0;
      for (int dir = 0; dir < 3; dir++) g12 += pow(b1[dir] - b2[dir], 2);
      g12 = sqrt(g12) * kCVAC;
      const double q12 = g12 * sigmaT / (local_gmax * sigmaT);
      if (mu_rand() > q12) continue;
      ncoll += 1;
      const double R = mu_rand();
      const double costh = 1.0 - 2.0 * R;
      const double sinth = sqrt(1.0 - costh * costh);
      const double phi = kTWOPI * mu_rand();
      double dU[3];
      orc_scatter_delta_u(b1[0] - b2[0], b1[1] - b2[1], b1[2] - b2[2], costh, sinth, cos(phi), sin(phi), dU);
      if (conservative && wp1 != wp2) {
        if (local_numCell < 3) continue;
        int random_index3 = mu_randint(0, local_numCell - 1);
        while (random_index3 == random_index2 || random_index3 == random_index1)
          random_index3 = mu_randint(0, local_numCell - 1);
        const long i3 = cs[c] + random_index3;
        double b3[3] = {v[i3], v[n + i3], v[2 * n + i3]};
        double wq3 = w[i3];
        if (wp1 < wp2) {
          double b2p[3], wq2 = wp2;
          for (int dir = 0; dir < 3; dir++) {
            b1[dir] += 0.5 * dU[dir];
            b2p[dir] = b2[dir] - 0.5 * dU[dir];
          }
          orc_collapse_three_to_two(b2, &wq2, b3, &wq3, b2p, wp1);
          w[i2] = wq2;
        } else {
          double b1p[3], wq1 = wp1;
          for (int dir = 0; dir < 3; dir++) {
            b1p[dir] = b1[dir] + 0.5 * dU[dir];
            b2[dir] -= 0.5 * dU[dir];
          }
          orc_collapse_three_to_two(b1, &wq1, b3, &wq3, b1p, wp2);
          w[i1] = wq1;
        }
        w[i3] = wq3;
        for (int dir = 0; dir < 3; dir++) {
          v[dir * n + i1] = b1[dir];
          v[dir * n + i2] = b2[dir];
          v[dir * n + i3] = b3[dir];
        }
        continue;
      }
      const double rand_num3 = mu_rand();
      if (rand_num3 <= wp2 / wp1)
        for (int dir = 0; dir < 3; dir++) v[dir * n + i1] = b1[dir] + 0.5 * dU[dir];
      if (rand_num3 <= wp1 / wp2)
        for (int dir = 0; dir < 3; dir++) v[dir * n + i2] = b2[dir] - 0.5 * dU[dir];
    }
  }
  if (ncand_out) *ncand_out = ncand;
  if (ncoll_out) *ncoll_out = ncoll;
}

/* VariableHardSphere (src/scattering/VariableHardSphere.cpp:28-47 constants, 217-412 applySelfScattering; the
 * reference implements self-scattering only): sigmaT(g) = 4 pi A g^(-4/alpha), alpha = 4/(2 eta - 1). */
extern "C" void orc_vhs_consts(double mass, double eta, double T0, double mu0, double *fourPiA, double *fourOverAlpha) {
  const double KB = 1.380649e-23;
  const double alpha = 4. / (2. * eta - 1.);
  const double Mass_kg = mass * kME;
  const double VT0 = sqrt(KB * T0 / Mass_kg);
  const double Gamma0 = tgamma(4. - 2. / alpha);
  const double Aconst = 15. / 32. / Gamma0 / mu0 * Mass_kg / sqrt(kPI) * VT0 * pow(4. * VT0 * VT0, 2. / alpha);
  *fourPiA = 4. * kPI * Aconst;
  *fourOverAlpha = 4.0 / alpha;
}

extern "C" void orc_vhs_self(long ncell, const long *cs, double *v, long n, const double *dens, const double *ene,
                             double mass, double fourPiA, double fourOverAlpha, double dt_sec, long *ncand_out,
                             long *ncoll_out) {
  const double cvacSq = kCVAC * kCVAC;
  long ncand = 0, ncoll = 0;
  for (long c = 0; c < ncell; ++c) {
    const double local_numberDensity = dens[c];
    if (local_numberDensity == 0.0) continue;
    double local_energyDensity = 0.0;
    for (int dir = 0; dir < 3; dir++) local_energyDensity = local_energyDensity + ene[dir * ncell + c];
    const double local_Teff = 2.0 / 3.0 * local_energyDensity / local_numberDensity * cvacSq;
    const double local_gmax = 5.0 * sqrt(local_Teff / mass);
    const double local_sigmaTmax = fourPiA * pow(local_gmax, -fourOverAlpha);
    const double local_nuMaxDt = local_numberDensity * local_sigmaTmax * local_gmax * dt_sec;
    const int local_numCell = (int)(cs[c + 1] - cs[c]);
    if (local_numCell < 2) continue;
    const double local_Nmax = 0.5 * (local_numCell - 1) * local_nuMaxDt;
    double whole;
    const double rem = modf(local_Nmax, &whole);
    double rand_num = mu_rand();
    int Nint = static_cast<int>(whole);
    if (rand_num < rem) Nint = Nint + 1;
    ncand += Nint;
    for (int k = 0; k < Nint; k++) {
      const int r1 = mu_randint(0, local_numCell - 1);
      int r2 = mu_randint(0, local_numCell - 1);
      while (r2 == r1) r2 = mu_randint(0, local_numCell - 1);
      const long i1 = cs[c] + r1, i2 = cs[c] + r2;
      double b1[3] = {v[i1], v[n + i1], v[2 * n + i1]}, b2[3] = {v[i2], v[n + i2], v[2 * n + i2]};
      double g12 = 0.0;
      for (int dir = 0; dir < 3; dir++) g12 = g12 + pow(b1[dir] - b2[dir], 2);
      g12 = sqrt(g12) * kCVAC;
      const double local_sigmaT = fourPiA * pow(g12, -fourOverAlpha);
      const double q12 = g12 * local_sigmaT / (local_gmax * local_sigmaTmax);
      rand_num = mu_rand();
      if (rand_num <= q12) {
        ncoll += 1;
        const double R = mu_rand();
        const double costh = 1.0 - 2.0 * R;
        const double sinth = sqrt(1.0 - costh * costh);
        const double phi = kTWOPI * mu_rand();
        double dU[3];
        orc_scatter_delta_u(b1[0] - b2[0], b1[1] - b2[1], b1[2] - b2[2], costh, sinth, cos(phi), sin(phi), dU);
        for (int dir = 0; dir < 3; dir++) {
          v[dir * n + i1] = b1[dir] + 0.5 * dU[dir];
          v[dir * n + i2] = b2[dir] - 0.5 * dU[dir];
        }
      }
    }
  }
  if (ncand_out) *ncand_out = ncand;
  if (ncoll_out) *ncoll_out = ncoll;
}

extern "C" void orc_hs_inter_wm(long ncell, const long *cs1, double *v1, double *w1, long n1, const double *dens1,
                                const double *ene1, double mass1, const long *cs2, double *v2, double *w2, long n2,
                                const double *dens2, const double *ene2, double mass2, double Vc, double sigmaT,
                                int conservative, double dt_sec, long *ncand_out, long *ncoll_out);
extern "C" void orc_hs_inter(long ncell, const long *cs1, double *v1, const double *w1, long n1, const double *dens1,
                             const double *ene1, double mass1, const long *cs2, double *v2, const double *w2, long n2,
                             const double *dens2, const double *ene2, double mass2, double Vc, double sigmaT,
                             double dt_sec, long *ncand_out, long *ncoll_out) {
  orc_hs_inter_wm(ncell, cs1, v1, const_cast<double *>(w1), n1, dens1, ene1, mass1, cs2, v2, const_cast<double *>(w2), n2,
                  dens2, ene2, mass2, Vc, sigmaT, 0, dt_sec, ncand_out, ncoll_out);
}
/* conservative != 0: the CONSERVATIVE weight method between species (HardSphere.cpp:594-636).  For unequal weights the
 * lighter particle moves by 0.5 deltaU and the heavier one's scattered copy by -0.5 deltaU -- NOT mu/m deltaU, as the
 * PROBABILISTIC branch below and the self-scattering code do: the reference conserves the pair's momentum here only
 * for equal masses (SURVEY App. B style quirk, restated as it stands) -- then the heavier particle, its scattered
 * fraction and a third particle of the heavier particle's species are merged (collapseThreeToTwo); weights change. */
extern "C" void orc_hs_inter_wm(long ncell, const long *cs1, double *v1, double *w1, long n1, const double *dens1,
                                const double *ene1, double mass1, const long *cs2, double *v2, double *w2, long n2,
                                const double *dens2, const double *ene2, double mass2, double Vc, double sigmaT,
                                int conservative, double dt_sec, long *ncand_out, long *ncoll_out) {
  const double cvacSq = kCVAC * kCVAC;
  const double mu = mass1 * mass2 / (mass1 + mass2);
  long ncand = 0, ncoll = 0;
  for (long c = 0; c < ncell; ++c) {
    const double nd1 = dens1[c], nd2 = dens2[c];
    if (nd1 * nd2 == 0.0) continue;
    double e1 = 0.0, e2 = 0.0;
    for (int dir = 0; dir < 3; dir++) {
      e1 = e1 + ene1[dir * ncell + c];
      e2 = e2 + ene2[dir * ncell + c];
    }
    const double Teff1 = 2.0 / 3.0 * e1 / nd1 * cvacSq, Teff2 = 2.0 / 3.0 * e2 / nd2 * cvacSq;
    const double local_gmax = 2.5 * sqrt(2.0 * std::max(Teff1, Teff2) / mu);
    const int numCell1 = (int)(cs1[c + 1] - cs1[c]), numCell2 = (int)(cs2[c + 1] - cs2[c]);
    if (numCell1 < 2 && numCell2 < 2) continue;
    if (numCell1 < 1 || numCell2 < 1) continue;   /* randInt(0,-1) in the reference: no pair exists */
    const double W1 = nd1 / numCell1 * Vc, W2 = nd2 / numCell2 * Vc;
    const double Wmax = std::max(W1, W2);
    const double local_Nmax = Wmax * numCell1 * numCell2 / Vc * sigmaT * local_gmax * dt_sec;
    double whole;
    const double rem = modf(local_Nmax, &whole);
    const double rand_num = mu_rand();
    int Nint = static_cast<int>(whole);
    if (rand_num < rem) Nint = Nint + 1;
    ncand += Nint;
    for (int k = 0; k < Nint; k++) {
      const long i1 = cs1[c] + mu_randint(0, numCell1 - 1), i2 = cs2[c] + mu_randint(0, numCell2 - 1);
      double b1[3] = {v1[i1], v1[n1 + i1], v1[2 * n1 + i1]}, b2[3] = {v2[i2], v2[n2 + i2], v2[2 * n2 + i2]};
      const double wp1 = w1[i1], wp2 = w2[i2];
      double g12 = 0.0;
      for (int dir = 0; dir < 3; dir++) g12 = g12 + pow(b1[dir] - b2[dir], 2);
      g12 = sqrt(g12) * kCVAC;
      const double q12 = g12 * sigmaT / (local_gmax * sigmaT);
      if (mu_rand() > q12) continue;
      ncoll += 1;
      const double R = mu_rand();
      const double costh = 1.0 - 2.0 * R;
      const double sinth = sqrt(1.0 - costh * costh);
      const double phi = kTWOPI * mu_rand();
      double dU[3];
      orc_scatter_delta_u(b1[0] - b2[0], b1[1] - b2[1], b1[2] - b2[2], costh, sinth, cos(phi), sin(phi), dU);
      if (conservative && wp1 != wp2) {
        if (wp1 < wp2) {
          if (numCell2 < 2) continue;
          double b2p[3], wq2 = wp2;
          for (int dir = 0; dir < 3; dir++) {
            b1[dir] += 0.5 * dU[dir];
            b2p[dir] = b2[dir] - 0.5 * dU[dir];
          }
          const int q2 = (int)(i2 - cs2[c]);
          int q3 = mu_randint(0, numCell2 - 1);
          while (q3 == q2) q3 = mu_randint(0, numCell2 - 1);
          const long i3 = cs2[c] + q3;
          double b3[3] = {v2[i3], v2[n2 + i3], v2[2 * n2 + i3]}, wq3 = w2[i3];
          orc_collapse_three_to_two(b2, &wq2, b3, &wq3, b2p, wp1);
          w2[i2] = wq2;
          w2[i3] = wq3;
          for (int dir = 0; dir < 3; dir++) {
            v1[dir * n1 + i1] = b1[dir];
            v2[dir * n2 + i2] = b2[dir];
            v2[dir * n2 + i3] = b3[dir];
          }
        } else {
          if (numCell1 < 2) continue;
          double b1p[3], wq1 = wp1;
          for (int dir = 0; dir < 3; dir++) {
            b1p[dir] = b1[dir] + 0.5 * dU[dir];
            b2[dir] -= 0.5 * dU[dir];
          }
          const int q1 = (int)(i1 - cs1[c]);
          int q3 = mu_randint(0, numCell1 - 1);
          while (q3 == q1) q3 = mu_randint(0, numCell1 - 1);
          const long i3 = cs1[c] + q3;
          double b3[3] = {v1[i3], v1[n1 + i3], v1[2 * n1 + i3]}, wq3 = w1[i3];
          orc_collapse_three_to_two(b1, &wq1, b3, &wq3, b1p, wp2);
          w1[i1] = wq1;
          w1[i3] = wq3;
          for (int dir = 0; dir < 3; dir++) {
            v1[dir * n1 + i1] = b1[dir];
            v1[dir * n1 + i3] = b3[dir];
            v2[dir * n2 + i2] = b2[dir];
          }
        }
        continue;
      }
      const double rand_num3 = mu_rand();
      if (rand_num3 <= wp2 / wp1)
        for (int dir = 0; dir < 3; dir++) v1[dir * n1 + i1] = b1[dir] + mu / mass1 * dU[dir];
      if (rand_num3 <= wp1 / wp2)
        for (int dir = 0; dir < 3; dir++) v2[dir * n2 + i2] = b2[dir] - mu / mass2 * dU[dir];
    }
  }
  if (ncand_out) *ncand_out = ncand;
  if (ncoll_out) *ncoll_out = ncoll;
}

/* Coulomb::setIntraMFT / setInterMFT (Coulomb.cpp:108-356) */
extern "C" double orc_coulomb_nu_max(long ncell, const double *LDe, const double *dens1, const double *mom1,
                                     const double *ene1, const double *dens2, const double *mom2,
                                     const double *ene2, double charge1, double charge2, double mass1,
                                     double mass2, double Clog_in, int intra) {
  const CoulombConsts k = coulomb_consts(charge1, charge2, mass1, mass2);
  const double cvacSq = kCVAC * kCVAC;
  const double mcSq_eV = kME * kEV_PER_JOULE * cvacSq;
  double box_nuMax = 0.0;
  for (long c = 0; c < ncell; ++c) {
    double g12sq, numFreq, sigma_max, EF_norm;
    const double bmax = LDe[c];
    if (intra) {
      const double numDen = dens1[c];
      if (numDen == 0.0) continue;
      const double rho = mass1 * numDen;
      const double atomic_spacing = 1.0 / std::cbrt(4.0 / 3.0 * kPI * numDen);
      sigma_max = 1.0 / (numDen * atomic_spacing);
      EF_norm = k.EF_fact * std::pow(numDen, 2.0 / 3.0);
      const double rhoUx = mom1[c], rhoUy = mom1[ncell + c], rhoUz = mom1[2 * ncell + c];
      const double meanE = (rhoUx * rhoUx + rhoUy * rhoUy + rhoUz * rhoUz) / rho / 2.0;
      double eneDen = 0.0;
      for (int dir = 0; dir < 3; dir++) eneDen += ene1[dir * ncell + c];
      double T_eV = 2.0 / 3.0 * (eneDen - meanE) / numDen * mcSq_eV;
      T_eV = std::max(T_eV, 0.01);
      g12sq = 6.0 * kQE / kME * T_eV / mass1;
      numFreq = numDen;
    } else {
      const double numDen1 = dens1[c], numDen2 = dens2[c];
      if (numDen1 * numDen2 == 0.0) continue;
      const double rho1 = mass1 * numDen1, rho2 = mass2 * numDen2;
      const double minn = std::min(numDen1, numDen2);
      const double atomic_spacing = 1.0 / std::cbrt(4.0 / 3.0 * kPI * minn);
      sigma_max = 1.0 / (minn * atomic_spacing);
      const double maxn = std::max(numDen1, numDen2);
      EF_norm = k.EF_fact * std::pow(maxn, 2.0 / 3.0);
      const double rhoUx1 = mom1[c], rhoUy1 = mom1[ncell + c], rhoUz1 = mom1[2 * ncell + c];
      const double meanE1 = (rhoUx1 * rhoUx1 + rhoUy1 * rhoUy1 + rhoUz1 * rhoUz1) / rho1 / 2.0;
      const double rhoUx2 = mom2[c], rhoUy2 = mom2[ncell + c], rhoUz2 = mom2[2 * ncell + c];
      const double meanE2 = (rhoUx2 * rhoUx2 + rhoUy2 * rhoUy2 + rhoUz2 * rhoUz2) / rho2 / 2.0;
      double eneDen1 = 0.0, eneDen2 = 0.0;
      for (int dir = 0; dir < 3; dir++) eneDen1 += ene1[dir * ncell + c];
      for (int dir = 0; dir < 3; dir++) eneDen2 += ene2[dir * ncell + c];
      double T1_eV = 2.0 / 3.0 * (eneDen1 - meanE1) / numDen1 * mcSq_eV;
      double T2_eV = 2.0 / 3.0 * (eneDen2 - meanE2) / numDen2 * mcSq_eV;
      T1_eV = std::max(T1_eV, 0.01);
      T2_eV = std::max(T2_eV, 0.01);
      const double VT1 = std::sqrt(kQE * T1_eV / (kME * mass1));
      const double VT2 = std::sqrt(kQE * T2_eV / (kME * mass2));
      g12sq = (3.0 * VT1 * VT1 + 3.0 * VT2 * VT2);
      g12sq += std::pow((rhoUx1 / rho1 - rhoUx2 / rho2), 2) * cvacSq;
      g12sq += std::pow((rhoUy1 / rho1 - rhoUy2 / rho2), 2) * cvacSq;
      g12sq += std::pow((rhoUz1 / rho1 - rhoUz2 / rho2), 2) * cvacSq;
      numFreq = maxn;
    }
    const double g12sq_norm = g12sq / cvacSq;
    const double b90 = k.b90_fact / (k.mu * g12sq_norm + 2.0 * EF_norm);
    double Clog = Clog_in;
    if (Clog == 0.0 && g12sq > 0.0) { /* Lee and More 1984 Eqs 20-22 */
      const double bmin_qm = k.bqm_fact / (k.mu * std::sqrt(g12sq_norm));
      const double bmin = std::max(b90 / 2.0, bmin_qm);
      Clog = 0.5 * std::log(1.0 + bmax * bmax / bmin / bmin);
      Clog = std::max(2.0, Clog);
    }
    double sigma90 = 8.0 / kPI * b90 * b90 * Clog;
    sigma90 = std::min(sigma90, sigma_max);
    const double nu90 = sqrt(g12sq) * numFreq * sigma90;
    box_nuMax = std::max(box_nuMax, nu90);
  }
  return box_nuMax;
}

/* Elastic::setInterMFT (Elastic.cpp:146-202).  Reference quirk kept by the caller: setMeanFreeTime
 * hands species 1's moments in for BOTH species (Elastic.cpp:130-131) while m_mass2 stays species 2's. */
extern "C" double orc_elastic_nu_max(long ncell, const double *dens1, const double *ene1, const double *dens2,
                                     const double *ene2, double mass1, double mass2, double const_sigma, int ntab,
                                     const double *E, const double *Q, const double *XI, int angular, int loglog) {
  const double mu = mass1 * mass2 / (mass1 + mass2);
  double box_nuMax = 0.0;
  for (long c = 0; c < ncell; ++c) {
    const double n1 = dens1[c], n2 = dens2[c];
    if (n1 * n2 == 0.0) continue;
    double b1 = 0.0, b2 = 0.0;
    for (int dir = 0; dir < 3; dir++) {
      b1 += 2.0 * ene1[dir * ncell + c];
      b2 += 2.0 * ene2[dir * ncell + c];
    }
    b1 /= n1 * mass1;
    b2 /= n2 * mass2;
    const double g12 = sqrt(b1 + b2);
    double xi;
    const double sigma = orc_elastic_sigma(g12, mu, const_sigma, ntab, E, Q, XI, angular, loglog, &xi);
    const double local_nuMax = n2 * sigma * g12 * kCVAC;
    box_nuMax = std::max(box_nuMax, local_nuMax);
  }
  return box_nuMax;
}
