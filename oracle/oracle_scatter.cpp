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

namespace {
/* draws exactly what TakizukaAbe::computeDeltaU draws, in its order: randn only
 * in the small-angle branch, rand for theta only in the other, then rand for phi */
void ta_pair(double *a, double *b, double den1, double den2, double b90_fact,
             double Clog, double dt_sec, long double mu, long double m1,
             long double m2) {
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
  const double b90_fact = orc_ta_b90_fact(charge, charge, mass, mass);
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
      ta_pair(a, b, den, den, b90_fact, Clog, dt_sec, mu, m1, m2);
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
        scatter(idx[q1], idx[q2], numDen / 2.0);
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
  const double b90_fact = orc_ta_b90_fact(charge1, charge2, mass1, mass2);
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
      ta_pair(a, b, numDen1, numDen2, b90_fact, Clog, dt_sec, mu, m1, m2);
      for (int k = 0; k < 3; ++k) {
        v1[k * n1 + i1] = a[k];
        v2[k * n2 + i2] = b[k];
      }
      ++npairs;
    }
  }
  if (npairs_out) *npairs_out = npairs;
}
