// ref_meshinterp.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the REFERENCE's own C++ gathers -- MeshInterp::interpolateEMfieldsToPart_testing and the private
// interpolateEMfieldsToPart_CIC / _TSC, interpolateBfieldsToPart_CIC, interpolateEToPart_CC0 (1D and 2D) and
// interpolateEToPart_CC1 (1D) it dispatches to (src/particle_tools/MeshInterpI.H:1013-1850; the NEW_EM_INTERP_METHOD
// path of PicChargedSpecies::interpolateFieldsToParticles, PicChargedSpecies.cpp:3895-3898) -- on plain arrays, compiled
// from where they lie under /root/reference by oracle/ref_build.sh against oracle/chombo_mock/.  They are the author's
// second, C++ statement of the gathers the production build runs from MeshInterpF.ChF / MeshInterpChargeConservingF.ChF,
// with the same stencils, indices and weights in a different operation order; tests/test_ref_pin_gather.py pins the
// oracle's restatement of the Fortran on them (indices exactly, values to a few ulp of the stencil sum).
// This file contains no reference source; it only includes and calls it.
#include <array>
#include <cstdlib>

// the reference's generated Fortran prototypes (a stale Chombo artefact) and the Fortran kernels themselves are not
// needed by the C++ gathers: skip the header, and turn the calls in the (never instantiated) templates into no-ops
#define _MESHINTERPF_F_H_
#define FORT_CIC_DEPOSIT_CURRENT(...) ((void)0)
#define FORT_TSC_DEPOSIT_CURRENT(...) ((void)0)
#define FORT_CC0_1D_DEPOSIT_CURRENT(...) ((void)0)
#define FORT_CC0_2D_DEPOSIT_CURRENT(...) ((void)0)
#define FORT_CC1_1D_DEPOSIT_CURRENT(...) ((void)0)
#define FORT_CC1_2D_DEPOSIT_CURRENT(...) ((void)0)
#define FORT_CIC_DEPOSIT_MASS_MATRIX(...) ((void)0)
#define FORT_TSC_DEPOSIT_MASS_MATRIX(...) ((void)0)
#define FORT_CC0_1D_DEPOSIT_MASS_MATRIX(...) ((void)0)
#define FORT_CC1_1D_DEPOSIT_MASS_MATRIX(...) ((void)0)
#define FORT_CC1_2D_DEPOSIT_MASS_MATRIX(...) ((void)0)
#define FORT_CIC_INTERPOLATE_FIELDS(...) ((void)0)
#define FORT_TSC_INTERPOLATE_FIELDS(...) ((void)0)
#define FORT_CC0_1D_INTERPOLATE_FIELDS(...) ((void)0)
#define FORT_CC0_2D_INTERPOLATE_FIELDS(...) ((void)0)
#define FORT_CC1_1D_INTERPOLATE_FIELDS(...) ((void)0)
#define FORT_CC1_2D_INTERPOLATE_FIELDS(...) ((void)0)

#include "ListBox.H"
#include "JustinsParticle.H"
#define private public   // the gathers are private members; the object is filled in directly (MeshInterp.cpp is not linked)
#include "MeshInterp.H"
#undef private

const IntVect IntVect::Zero = [] { IntVect v; return v; }();
const IntVect IntVect::Unit = [] { IntVect v; for (int d = 0; d < SpaceDim; ++d) v[d] = 1; return v; }();
#ifdef REFMI_DEFINE_REALVECT_ZERO
const RealVect RealVect::Zero;
#endif

namespace {
// the constructors live in MeshInterp.cpp (which needs the Fortran kernels to link); all members are plain data
MeshInterp* make_interp(const double* le, const double* re, const double* dx, int ghosts) {
  MeshInterp* m = static_cast<MeshInterp*>(std::calloc(1, sizeof(MeshInterp)));
  for (int d = 0; d < SpaceDim; ++d) {
    m->m_dx[d] = dx[d];
    m->m_domainLeftEdge[d] = le[d];
    m->m_domainRightEdge[d] = re[d];
  }
  m->m_ghosts = ghosts;
  return m;
}
FArrayBox make_fab(double* data, const int* lo, const int* hi, const int* typ, int ncomp) {
  IntVect l, h, t;
  for (int d = 0; d < SpaceDim; ++d) {
    l[d] = lo[d];
    h[d] = hi[d];
    t[d] = typ[d];
  }
  return FArrayBox(Box(l, h, t), ncomp, data);
}
}  // namespace

extern "C" {

int refmi_spacedim(void) { return SpaceDim; }

// interp: the reference's InterpType (MeshInterp.H:27): CIC = 2, TSC = 3, CC0 = 5, CC1 = 6.
// x = xbar, xold: [D][n] component major.  Six field arrays F[c] (Ex Ey Ez Bx By Bz as MeshInterp receives them) with
// inclusive bounds lo/hi [6][D], centring typ [6][D] and component counts ncomp[6] (Chombo layout, component slowest).
// Ep, Bp: [3][n] out.  Returns 0, or 1 if interp is not one of the four.
int refmi_gather(int interp, long n, const double* x, const double* xold, const double* le, const double* re,
                 const double* dx, int ghosts, double* const* F, const int* lo, const int* hi, const int* typ,
                 const int* ncomp, double* Ep, double* Bp) {
  if (interp != CIC && interp != TSC && interp != CC0 && interp != CC1) return 1;
  MeshInterp* mi = make_interp(le, re, dx, ghosts);
  List<JustinsParticle> lst;
  for (long p = 0; p < n; ++p) {
    JustinsParticle q;
    RealVect xb, xo;
    for (int d = 0; d < SpaceDim; ++d) {
      xb[d] = x[d * n + p];
      xo[d] = xold[d * n + p];
    }
    q.setPosition(xb);
    q.setOldPosition(xo);
    lst.add(q);
  }
  FArrayBox f0 = make_fab(F[0], lo + 0 * SpaceDim, hi + 0 * SpaceDim, typ + 0 * SpaceDim, ncomp[0]);
  FArrayBox f1 = make_fab(F[1], lo + 1 * SpaceDim, hi + 1 * SpaceDim, typ + 1 * SpaceDim, ncomp[1]);
  FArrayBox f2 = make_fab(F[2], lo + 2 * SpaceDim, hi + 2 * SpaceDim, typ + 2 * SpaceDim, ncomp[2]);
  FArrayBox f3 = make_fab(F[3], lo + 3 * SpaceDim, hi + 3 * SpaceDim, typ + 3 * SpaceDim, ncomp[3]);
  FArrayBox f4 = make_fab(F[4], lo + 4 * SpaceDim, hi + 4 * SpaceDim, typ + 4 * SpaceDim, ncomp[4]);
  FArrayBox f5 = make_fab(F[5], lo + 5 * SpaceDim, hi + 5 * SpaceDim, typ + 5 * SpaceDim, ncomp[5]);
  mi->interpolateEMfieldsToPart_testing(lst, f0, f1, f2, f3, f4, f5, static_cast<InterpType>(interp));
  long p = 0;
  for (ListIterator<JustinsParticle> lit(lst); lit.ok(); ++lit, ++p) {
    const std::array<Real, 3>& e = lit().electric_field();
    const std::array<Real, 3>& b = lit().magnetic_field();
    for (int c = 0; c < 3; ++c) {
      Ep[c * n + p] = e[c];
      Bp[c * n + p] = b[c];
    }
  }
  std::free(mi);
  return 0;
}

}  // extern "C"
