// example_shock_1d.cpp -- the particle side of the implicit shock decks (regression_tests/1d/plasma_shock_implicit:
// inflow_outflow at the low end with pic_species.N.suborbit_inflow_J = true, outflow at the high end) written against the
// host classes the way PICTimeIntegrator_EM_ThetaImplicit and PicSpeciesInterface::preRHSOp drive PicChargedSpecies
// (src/time/PICTimeIntegrator_EM_ThetaImplicit.cpp:193-364, src/species/pic/PicSpeciesInterface.cpp:899-994), field free and
// self-checking: exact particle bookkeeping through the inflow list, the species and the outflow list, packed host I/O,
// the List<JustinsParticle> record round trip, and the fused explicit step.  Exit code 0 = all checks passed.
//   usage: example_shock_1d [steps]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "PicGpuHost.H"

using namespace picnic_gpu;

static unsigned long long s_rng = 88172645463325252ull;
static double urand() {   // xorshift64*: the host's createInflowParticles stand-in needs some generator
  s_rng ^= s_rng >> 12;
  s_rng ^= s_rng << 25;
  s_rng ^= s_rng >> 27;
  return (double)((s_rng * 2685821657736338717ull) >> 11) * (1.0 / 9007199254740992.0);
}
static int fail(const char *what) {
  std::fprintf(stderr, "example_shock_1d: FAILED: %s\n", what);
  return 1;
}

int main(int argc, char **argv) {
  const int steps = argc > 1 ? std::atoi(argv[1]) : 120;
  const int D = 1, ncell[2] = {32, 1}, nghost = 4, periodic[2] = {0, 0}, lo[2] = {0, 0}, hi[2] = {31, 0};
  const double xmin[2] = {0.0, 0.0}, dx[2] = {0.25, 1.0}, dt = 0.5, cvac = 0.9986, cdt = dt * cvac;
  initialize(0);
  int rc = 0;
  {
    Mesh mesh(D, ncell, xmin, dx, nghost, periodic, lo, hi, 2.0);
    // zero fields through the packed path: one buffer, six components back to back
    std::vector<double> fields((size_t)mesh.packedFieldSize(), 0.0), Jbuf((size_t)mesh.packedCurrentSize(), 0.0);
    PicChargedSpecies sp(mesh, "electron", 1.0, -1.0, -0.7, cvac, TSC, CC1, CC1);
    sp.setParticleSolverParams(false, false, 25, 1.0e-12, 0, 0);
    sp.setParticles(0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    const int bc_lo[2] = {PGPU_BC_INFLOW_OUTFLOW, PGPU_BC_NONE}, bc_hi[2] = {PGPU_BC_OUTFLOW, PGPU_BC_NONE};
    const int nin = 30;
    const double ud = 0.4 * dx[0] / cdt;
    double injected = 0.0, left = 0.0;
    uint64_t next_id = 1;
    for (int step = 0; step < steps && !rc; ++step) {
      sp.removeOutflowParticles();
      // createInflowParticles (host): where the particles would be one step before they cross the boundary
      std::vector<double> x(nin), v(3 * nin), w(nin, 1.0);
      std::vector<uint64_t> id(nin);
      for (int k = 0; k < nin; ++k) {
        v[k] = ud * (0.8 + 0.4 * urand());
        v[nin + k] = v[2 * nin + k] = 0.0;
        x[k] = 0.0 - v[k] * cdt * urand();
        id[k] = next_id++;
      }
      sp.addToInflowList(nin, x.data(), v.data(), w.data(), id.data(), 0, 0);
      sp.updateOldParticlePositions();
      sp.updateOldParticleVelocities();
      for (int it = 0; it < 2; ++it) {                   // two nonlinear evaluations: preRHSOp
        mesh.setEMfieldsPacked(fields.data());
        mesh.zeroCurrentDensity();
        sp.advanceParticlesIteratively(dt, true);
        mesh.addSpeciesCurrentDensity(sp);
        sp.advanceInflowParticlesAndSetJ(dt, false);       // PicSpeciesInterface::addInflowJ
        mesh.addInflowJ(sp);
        mesh.finalizeSettingJ();
        mesh.getCurrentDensityPacked(Jbuf.data());
        if (pgpu_synchronize()) fatal("example_shock_1d: synchronize");   // the host reads J here
      }
      sp.advanceVelocities_2ndHalf();
      sp.advancePositions_2ndHalf();
      sp.applyBCs(bc_lo, bc_hi);                           // inflow_Lo: the list joins; outflow_Hi: leavers to their list
      if (sp.numInflowParticles() != 0) rc = fail("inflow list not emptied by applyBCs");
      double fin[20], fout[20];
      sp.inflowProbes(fin);
      sp.outflowProbes(fout);
      injected += fin[0];
      left += fout[5];
    }
    if (!rc && injected != (double)steps * nin) rc = fail("not every inflow particle joined");
    if (!rc && injected != (double)sp.numParticles() + left) rc = fail("bookkeeping: injected != inside + left");
    // the current of the last evaluation: electrons moving to +x everywhere inside
    if (!rc) {
      int jl[2], jh[2];
      mesh.fieldBounds(0, jl, jh);
      for (int k = 1; k < 31 && !rc; ++k)
        if (!(Jbuf[(size_t)(k - jl[0])] < 0.0)) rc = fail("J_x must be negative on every interior edge");
    }
    // List<JustinsParticle> records: [w | x | xold | virt virt | v[3] | vold[3] | ID], round trip
    if (!rc) {
      const long n = sp.numParticles(), nw = sp.linearSize() / (long)sizeof(double);
      if (nw != 2 * D + 10) rc = fail("linear record size");
      std::vector<double> rec((size_t)(n * nw)), rec2((size_t)(n * nw));
      sp.getParticlesLinear(rec.data());
      for (long i = 0; i < n && !rc; ++i)
        if (rec[i * nw] != 1.0 || rec[i * nw + 1] < 0.0 || rec[i * nw + 1] >= 8.0) rc = fail("linear record content");
      sp.setParticlesLinear(n, rec.data());
      sp.getParticlesLinear(rec2.data());
      if (!rc && std::memcmp(rec.data(), rec2.data(), rec.size() * sizeof(double)) != 0) rc = fail("linear round trip");
    }
    // the explicit leap-frog step in one call moves the (field-free) particles by u dt
    if (!rc) {
      const long n = sp.numParticles();
      std::vector<double> x0(n), x1(n), v(3 * n), tmp(3 * n), w(n);
      std::vector<uint64_t> id(n);
      sp.getParticles(x0.data(), tmp.data(), v.data(), tmp.data(), w.data(), id.data());
      sp.updateOldParticlePositions();
      sp.updateOldParticleVelocities();
      const int pbc[2] = {PGPU_BC_NONE, PGPU_BC_NONE};
      sp.explicitStep(dt, pbc, pbc, true);
      sp.getParticles(x1.data(), tmp.data(), tmp.data(), tmp.data(), w.data(), id.data());
      for (long i = 0; i < n && !rc; ++i)
        if (std::fabs(x1[i] - (x0[i] + v[i] * cdt)) > 1.0e-13) rc = fail("explicit step of a free particle");
    }
    std::printf("example_shock_1d: steps=%d injected=%.0f inside=%d left=%.0f %s\n", steps, injected, sp.numParticles(), left,
                rc ? "FAILED" : "ok");
  }
  finalize();
  return rc;
}
