// example_driver.cpp -- the particle side of one theta-implicit step written against the host
// classes exactly as PICTimeIntegrator_EM_ThetaImplicit + PicSpeciesInterface::preRHSOp drive
// PicChargedSpecies (src/time/PICTimeIntegrator_EM_ThetaImplicit.cpp:193-364,
// src/species/pic/PicSpeciesInterface.cpp:899-994).  Reads a problem written by
// tests/test_host_shim.py, runs it on the GPU and writes J and the particles back.
//   usage: example_driver <problem.bin> <result.bin>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "PicGpuHost.H"

using namespace picnic_gpu;

static void rd(FILE *f, void *p, size_t n) {
  if (fread(p, 1, n, f) != n) fatal("example_driver: short read");
}

int main(int argc, char **argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s problem.bin result.bin\n", argv[0]);
    return 2;
  }
  FILE *f = std::fopen(argv[1], "rb");
  if (!f) fatal("example_driver: cannot open problem file");
  int hdr[8];   // D, ncell0, ncell1, nghost, n, n_outer, iter_max, pad
  double par[10];  // xmin0 xmin1 dx0 dx1 dt rtol fnorm cvac_norm volume_scale pad
  rd(f, hdr, sizeof(hdr));
  rd(f, par, sizeof(par));
  const int D = hdr[0], ncell[2] = {hdr[1], hdr[2]}, nghost = hdr[3], n_outer = hdr[5];
  const long n = hdr[4];
  const double xmin[2] = {par[0], par[1]}, dx[2] = {par[2], par[3]}, dt = par[4];
  const int periodic[2] = {1, 1}, lo[2] = {0, 0}, hi[2] = {ncell[0] - 1, ncell[1] - 1};

  initialize(0);
  {
    Mesh mesh(D, ncell, xmin, dx, nghost, periodic, lo, hi, par[8]);
    // fields: six components with the bounds the mesh reports
    std::vector<std::vector<double>> F(6);
    FabRef R[6];
    for (int c = 0; c < 6; ++c) {
      mesh.fieldBounds(c, R[c].lo, R[c].hi);
      R[c].lo[1] = D == 2 ? R[c].lo[1] : 0;
      R[c].hi[1] = D == 2 ? R[c].hi[1] : 0;
      size_t sz = 1;
      for (int d = 0; d < D; ++d) sz *= (size_t)(R[c].hi[d] - R[c].lo[d] + 1);
      F[c].resize(sz);
      rd(f, F[c].data(), sz * sizeof(double));
      R[c].data = F[c].data();
    }
    std::vector<double> x(D * n), v(3 * n), w(n);
    std::vector<uint64_t> id(n);
    rd(f, x.data(), x.size() * 8);
    rd(f, v.data(), v.size() * 8);
    rd(f, w.data(), w.size() * 8);
    for (long i = 0; i < n; ++i) id[i] = (uint64_t)i;
    std::fclose(f);

    PicChargedSpecies sp(mesh, "electron", 1.0, -1.0, par[6], par[7], TSC, CC1, CC1);
    sp.setParticleSolverParams(false, false, hdr[6], par[5], 0, 0);
    sp.setParticles(n, x.data(), x.data(), v.data(), v.data(), w.data(), id.data());
    sp.binTheParticles();

    // ---- one time step -------------------------------------------------------------------
    sp.updateOldParticlePositions();
    sp.updateOldParticleVelocities();
    for (int it = 0; it < n_outer; ++it) {             // nonlinear function evaluations
      mesh.setEMfields(R[0], R[1], R[2], R[3], R[4], R[5]);     // preRHSOp: E,B of this iterate
      mesh.zeroCurrentDensity();
      sp.advanceParticlesIteratively(dt, true);         // + setCurrentDensity, fused
      mesh.addSpeciesCurrentDensity(sp);
      mesh.finalizeSettingJ();
    }
    std::vector<std::vector<double>> J(3);
    for (int c = 0; c < 3; ++c) {
      J[c].resize(F[c].size());
      FabRef out = R[c];
      out.data = J[c].data();
      mesh.getCurrentDensity(c, out);
    }
    // ---- mass matrices (use_mass_matrices decks): PicSpeciesInterface::setMassMatrices at the converged orbits,
    // then computeJfromMassMatrices with E == E0 must give back the current just deposited (J0 is that current)
    if (D == 2 && nghost >= 3) {
      std::vector<PicChargedSpecies *> one(1, &sp);
      int ncomp[18];
      mesh.initializeMassMatrices(PGPU_CC1, ncomp);
      mesh.setMassMatrices(one, dt);
      mesh.computeJfromMassMatrices();
      mesh.finalizeSettingJ();
      double worst = 0.0;
      for (int c = 0; c < 3; ++c) {
        std::vector<double> Jm(F[c].size());
        FabRef out = R[c];
        out.data = Jm.data();
        mesh.getCurrentDensity(c, out);
        double scale = 0.0, diff = 0.0;
        for (size_t k = 0; k < Jm.size(); ++k) {
          scale = std::fmax(scale, std::fabs(J[c][k]));
          diff = std::fmax(diff, std::fabs(Jm[k] - J[c][k]));
        }
        worst = std::fmax(worst, diff / scale);
      }
      std::printf("example_driver: mass matrices ncomp_xx=%dx%d, J(E0) vs deposited J: %.3e\n", ncomp[0], ncomp[1], worst);
    }
    const int bc[2] = {PGPU_BC_PERIODIC, PGPU_BC_PERIODIC};
    sp.advanceVelocities_2ndHalf();
    sp.advancePositions_2ndHalf();
    sp.applyBCs(bc, bc);
    uint64_t parts_its, apply_its;
    sp.picardParams(parts_its, apply_its);

    std::vector<double> xo(D * n), vo(3 * n);
    sp.getParticles(x.data(), xo.data(), v.data(), vo.data(), w.data(), id.data());
    FILE *g = std::fopen(argv[2], "wb");
    if (!g) fatal("example_driver: cannot open result file");
    for (int c = 0; c < 3; ++c) std::fwrite(J[c].data(), 8, J[c].size(), g);
    std::fwrite(x.data(), 8, x.size(), g);
    std::fwrite(v.data(), 8, v.size(), g);
    std::fwrite(id.data(), 8, id.size(), g);
    std::fclose(g);
    std::printf("example_driver: n=%ld apply_its=%llu unconverged=%llu\n", n, (unsigned long long)apply_its,
                (unsigned long long)sp.numUnconverged());

    // ---- collisions: ScatteringInterface::applyScattering for one self-scattering model ------
    // prepForScatter (bin + cell moments), Scattering::setMeanFreeTime -> scatterDt, applyScattering;
    // done after the result file is written so that the step above stays checkable on its own
    sp.binTheParticles();
    sp.setNumberDensityFromBinFab();
    std::vector<PicChargedSpecies *> all(1, &sp);
    TakizukaAbe ta(0, 0, 3.0);
    ta.setMeanFreeTime(all);
    Scattering::setRandomState(1983, 0);
    ta.applyScattering(all, dt * 1.77e-17);
    std::printf("example_driver: TA scatterDt=%.17g pairs=%ld\n", ta.scatterDt(), ta.lastPairCount());
  }
  finalize();
  return 0;
}
