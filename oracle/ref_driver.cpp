// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" entry points that drive the REFERENCE's own code (compiled from /root/reference by
// oracle/ref_build.sh against oracle/chombo_mock/) on plain arrays, so that the oracle restatement
// can be pinned against the real thing:
//   ref_boris           -> PicSpeciesUtils::applyForces        (src/species/pic/PicSpeciesUtils.cpp:8-101)
//   ref_boris_curvilinear -> PicSpeciesUtils::applyForces_CYL_CYL/_SPH_SPH/_CYL_HYB/_SPH_HYB (PicSpeciesUtils.cpp:103-473)
//   ref_delta_u         -> ScatteringUtils::computeDeltaU      (src/scattering/ScatteringUtils.H:84-111)
//   ref_rotate_velocity -> ScatteringUtils::rotateVelocity     (src/scattering/ScatteringUtils.H:51-82)
//   ref_scattering_cos  -> ScatteringUtils::getScatteringCos   (src/scattering/ScatteringUtils.H:12-18)
//   ref_collapse_three_to_two -> ScatteringUtils::collapseThreeToTwo (src/scattering/ScatteringUtils.H:20-47)
//   ref_mod_energy_pair -> ScatteringUtils::modEnergyPairwise  (src/scattering/ScatteringUtils.H:113-205)
//   ref_particle_wire   -> JustinsParticle::linearOut          (src/particle_tools/JustinsParticle.cpp:339-383)
//   ref_implicit_gamma  -> PicSpeciesUtils::getImplicitGamma   (src/species/pic/PicSpeciesUtils.H:43-52; -DRELATIVISTIC_PARTICLES)
// This file contains no reference source; it only calls it.
#include <array>
#include <cstring>

#include "PicSpeciesUtils.H"
#include "ScatteringUtils.H"

extern "C" {

int ref_spacedim(void) { return SpaceDim; }

// the RELATIVISTIC_PARTICLES build: pusher variant of ref_boris, and PicSpeciesUtils::getImplicitGamma
static int g_higuera_cary = 0;
void ref_set_higuera_cary(int on) { g_higuera_cary = on; }
#ifdef RELATIVISTIC_PARTICLES
int ref_is_relativistic(void) { return 1; }
double ref_implicit_gamma(const double *upold, const double *upbar) {
  std::array<Real, 3> uo = {upold[0], upold[1], upold[2]}, un;
  for (int n = 0; n < 3; ++n) un[n] = 2.0 * upbar[n] - upold[n];
  return PicSpeciesUtils::getImplicitGamma(uo, un);
}
#else
int ref_is_relativistic(void) { return 0; }
#endif

// v[c*n+p] (out), vold/Ep/Bp[c*n+p] (in): SoA, component-major like the oracle
void ref_boris(long n, double *v, const double *vold, const double *Ep, const double *Bp, double fnorm,
               double cnormDt, int byHalfDt) {
  List<JustinsParticle> lst;
  for (long p = 0; p < n; ++p) {
    JustinsParticle q;
    q.setOldVelocity({vold[p], vold[n + p], vold[2 * n + p]});
    q.setVelocity({v[p], v[n + p], v[2 * n + p]});
    q.setElectricField({Ep[p], Ep[n + p], Ep[2 * n + p]});
    q.setMagneticField({Bp[p], Bp[n + p], Bp[2 * n + p]});
    lst.add(q);
  }
#ifdef RELATIVISTIC_PARTICLES
  PicSpeciesUtils::applyForces(lst, fnorm, cnormDt, g_higuera_cary != 0, byHalfDt != 0, false);
#else
  PicSpeciesUtils::applyForces(lst, fnorm, cnormDt, byHalfDt != 0, false);
#endif
  long p = 0;
  for (ListIterator<JustinsParticle> lit(lst); lit.ok(); ++lit, ++p) {
    const std::array<Real, 3> &u = lit().velocity();
    v[p] = u[0];
    v[n + p] = u[1];
    v[2 * n + p] = u[2];
  }
}

// PicSpeciesUtils::applyForces_CYL_CYL / _SPH_SPH / _CYL_HYB / _SPH_HYB (src/species/pic/PicSpeciesUtils.cpp:103-473),
// type 1..4; r_old -> position_old()[0]; virt[k*n+p] <-> position_virt()[k]
void ref_boris_curvilinear(int type, long n, double *v, const double *vold, const double *Ep, const double *Bp,
                           const double *r_old, double *virt, double fnorm, double cnormDt, int byHalfDt,
                           int anticyclic) {
  List<JustinsParticle> lst;
  for (long p = 0; p < n; ++p) {
    JustinsParticle q;
    q.setOldVelocity({vold[p], vold[n + p], vold[2 * n + p]});
    q.setVelocity({0.0, 0.0, 0.0});
    q.setElectricField({Ep[p], Ep[n + p], Ep[2 * n + p]});
    q.setMagneticField({Bp[p], Bp[n + p], Bp[2 * n + p]});
    RealVect xo;
    for (int d = 0; d < SpaceDim; ++d) xo[d] = 0.0;
    xo[0] = r_old[p];
    q.setOldPosition(xo);
    std::array<Real, 4 - CH_SPACEDIM> &pv = q.position_virt();
    for (int k = 0; k < 4 - CH_SPACEDIM && k < 2; ++k) pv[k] = virt[k * n + p];
    lst.add(q);
  }
  if (type == 1) PicSpeciesUtils::applyForces_CYL_CYL(lst, fnorm, cnormDt, byHalfDt != 0, anticyclic != 0);
  else if (type == 2) PicSpeciesUtils::applyForces_SPH_SPH(lst, fnorm, cnormDt, byHalfDt != 0);
  else if (type == 3) PicSpeciesUtils::applyForces_CYL_HYB(lst, fnorm, cnormDt, anticyclic != 0);
  else PicSpeciesUtils::applyForces_SPH_HYB(lst, fnorm, cnormDt);
  long p = 0;
  for (ListIterator<JustinsParticle> lit(lst); lit.ok(); ++lit, ++p) {
    const std::array<Real, 3> &u = lit().velocity();
    v[p] = u[0];
    v[n + p] = u[1];
    v[2 * n + p] = u[2];
    const std::array<Real, 4 - CH_SPACEDIM> &pv = lit().position_virt();
    for (int k = 0; k < 4 - CH_SPACEDIM && k < 2; ++k) virt[k * n + p] = pv[k];
  }
}

void ref_delta_u(double ux, double uy, double uz, double costh, double sinth, double cosphi, double sinphi,
                 double *dU) {
  std::array<Real, 3> d;
  ScatteringUtils::computeDeltaU(d, ux, uy, uz, costh, sinth, cosphi, sinphi);
  dU[0] = d[0];
  dU[1] = d[1];
  dU[2] = d[2];
}

void ref_rotate_velocity(double *u, double costh, double sinth, double cosphi, double sinphi) {
  std::array<Real, 3> a = {u[0], u[1], u[2]};
  ScatteringUtils::rotateVelocity(a, costh, sinth, cosphi, sinphi);
  u[0] = a[0];
  u[1] = a[1];
  u[2] = a[2];
}

double ref_scattering_cos(double R, double xi) { return ScatteringUtils::getScatteringCos(R, xi); }

void ref_mod_energy_pair(double *b1, double *b2, double wpmp1, double wpmp2, double Erel_frac, double *Erel_cumm,
                         double *deltaE) {
  std::array<Real, 3> a = {b1[0], b1[1], b1[2]}, b = {b2[0], b2[1], b2[2]};
  long double dE = *deltaE;
  ScatteringUtils::modEnergyPairwise(a, b, wpmp1, wpmp2, Erel_frac, *Erel_cumm, dE);
  for (int i = 0; i < 3; ++i) {
    b1[i] = a[i];
    b2[i] = b[i];
  }
  *deltaE = (double)dE;
}

// ScatteringUtils::collapseThreeToTwo (src/scattering/ScatteringUtils.H:20-47): vp2/wp2 and vp3/wp3 are updated
void ref_collapse_three_to_two(double *vp2, double *wp2, double *vp3, double *wp3, const double *vp2p, double wp2p) {
  std::array<Real, 3> a = {vp2[0], vp2[1], vp2[2]}, b = {vp3[0], vp3[1], vp3[2]};
  const std::array<Real, 3> c = {vp2p[0], vp2p[1], vp2p[2]};
  ScatteringUtils::collapseThreeToTwo(a, *wp2, b, *wp3, c, wp2p);
  for (int i = 0; i < 3; ++i) {
    vp2[i] = a[i];
    vp3[i] = b[i];
  }
}

// wire format of one particle (what MPI migration and checkpoints carry); returns bytes
int ref_particle_wire(double w, const double *x, const double *xold, const double *v, const double *vold,
                      unsigned long long id, double *buf) {
  RealVect X, Xo;
  for (int d = 0; d < SpaceDim; ++d) {
    X[d] = x[d];
    Xo[d] = xold[d];
  }
  JustinsParticle q(w, X, {v[0], v[1], v[2]});
  q.setOldPosition(Xo);
  q.setOldVelocity({vold[0], vold[1], vold[2]});
  q.setID(id);
  q.linearOut(buf);
  return q.size();
}

}  // extern "C"
