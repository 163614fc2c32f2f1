// example_two_boxes.cpp -- the multi-box part of the host shim on ONE device: a periodic 32 x 16 domain as one
// box and as 2 x 2 boxes that live in this process.  The boxes sum their ghost currents through
// picnic_gpu::GhostExchange (peer-memory kernels; here plain device pointers instead of CUDA IPC) and hand over
// their leavers through picnic_gpu::ParticleMigration.  Checks: J of every box (ghosts included) equals the
// single-box J at the periodic image; after migration every particle sits in the box that owns it and none is lost.
// A PICNIC build does the same with one box per MPI rank and GhostExchange::connect(MPI_Allgather wrapper).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "PicGpuHost.H"

using namespace picnic_gpu;

static unsigned long long g_state = 88172645463325252ULL;
static double urand() {   // xorshift64*
  g_state ^= g_state >> 12;
  g_state ^= g_state << 25;
  g_state ^= g_state >> 27;
  return (double)((g_state * 2685821657736338717ULL) >> 11) / 9007199254740992.0;
}

struct Fab {
  std::vector<double> a;
  int lo[2], hi[2];
  FabRef ref() { FabRef r = {a.data(), {lo[0], lo[1]}, {hi[0], hi[1]}}; return r; }
  double at(int i, int j) const { return a[(size_t)(i - lo[0]) + (size_t)(j - lo[1]) * (hi[0] - lo[0] + 1)]; }
};
static Fab fetchJ(Mesh &m, int comp) {
  Fab f;
  m.fieldBounds(comp, f.lo, f.hi);
  f.a.assign((size_t)(f.hi[0] - f.lo[0] + 1) * (f.hi[1] - f.lo[1] + 1), 0.0);
  m.getCurrentDensity(comp, f.ref());
  return f;
}

int main() {
  initialize(0);
  const int NC[2] = {32, 16}, NB[2] = {16, 8}, NG = 2, PER[2] = {1, 1};
  const double XMIN[2] = {0.0, -1.0}, DX[2] = {0.25, 0.5};
  const long n = 20000;
  std::vector<double> x(2 * n), xo(2 * n), v(3 * n), w(n);
  std::vector<uint64_t> id(n);
  for (long p = 0; p < n; ++p) {
    for (int d = 0; d < 2; ++d) {
      xo[d * n + p] = XMIN[d] + urand() * NC[d] * DX[d];
      x[d * n + p] = xo[d * n + p] + (urand() - 0.5) * 0.8 * DX[d];
    }
    for (int c = 0; c < 3; ++c) v[c * n + p] = 0.1 * (urand() - 0.5);
    w[p] = 0.5 + urand();
    id[p] = 1000 + p;
  }
  // one box spanning the domain
  std::vector<Fab> Jg;
  Fab rho_g;
  {
    const int lo[2] = {0, 0}, hi[2] = {NC[0] - 1, NC[1] - 1};
    Mesh mesh(2, NC, XMIN, DX, NG, PER, lo, hi, 1.0);
    PicChargedSpecies sp(mesh, "electron", 1.0, -1.0, 1.0, 1.0, TSC, CC1, CC1);
    sp.setParticles(n, x.data(), xo.data(), v.data(), v.data(), w.data(), id.data());
    sp.setCurrentDensity(1.0);
    mesh.zeroCurrentDensity();
    mesh.addSpeciesCurrentDensity(sp);
    mesh.finalizeSettingJ();
    for (int c = 0; c < 3; ++c) Jg.push_back(fetchJ(mesh, c));
    // charge density on the nodes of the one box (setChargeDensityOnNodes)
    rho_g.lo[0] = rho_g.lo[1] = -NG;
    rho_g.hi[0] = NC[0] - 1 + NG + 1;
    rho_g.hi[1] = NC[1] - 1 + NG + 1;
    rho_g.a.assign((size_t)(rho_g.hi[0] - rho_g.lo[0] + 1) * (rho_g.hi[1] - rho_g.lo[1] + 1), 0.0);
    sp.setChargeDensityOnNodes(rho_g.ref());
  }
  // 2 x 2 boxes in this process
  BoxLayout lay(2, NC, NB, NG, PER);
  const int world = lay.numBoxes();
  std::vector<Mesh *> mesh(world);
  std::vector<PicChargedSpecies *> sp(world);
  std::vector<GhostExchange *> gx(world), gxr(world);
  const int NODES[2] = {1, 1};
  std::vector<ParticleMigration *> mg(world);
  for (int r = 0; r < world; ++r) {
    int lo[2], hi[2];
    lay.box(r, lo, hi);
    mesh[r] = new Mesh(2, NC, XMIN, DX, NG, PER, lo, hi, 1.0);
    sp[r] = new PicChargedSpecies(*mesh[r], "electron", 1.0, -1.0, 1.0, 1.0, TSC, CC1, CC1);
    std::vector<double> bx[2], bxo[2], bv[3], bw;
    std::vector<uint64_t> bid;
    for (long p = 0; p < n; ++p) {
      const int b0 = (int)std::floor((xo[p] - XMIN[0]) / (DX[0] * NB[0])), b1 = (int)std::floor((xo[n + p] - XMIN[1]) / (DX[1] * NB[1]));
      if (b0 + b1 * lay.nb[0] != r) continue;
      for (int d = 0; d < 2; ++d) { bx[d].push_back(x[d * n + p]); bxo[d].push_back(xo[d * n + p]); }
      for (int c = 0; c < 3; ++c) bv[c].push_back(v[c * n + p]);
      bw.push_back(w[p]);
      bid.push_back(id[p]);
    }
    const long m = (long)bw.size();
    std::vector<double> X, XO, V;
    for (int d = 0; d < 2; ++d) { X.insert(X.end(), bx[d].begin(), bx[d].end()); XO.insert(XO.end(), bxo[d].begin(), bxo[d].end()); }
    for (int c = 0; c < 3; ++c) V.insert(V.end(), bv[c].begin(), bv[c].end());
    sp[r]->setParticles(m, X.data(), XO.data(), V.data(), V.data(), bw.data(), bid.data());
    sp[r]->setCurrentDensity(1.0);
    mesh[r]->zeroCurrentDensity();
    mesh[r]->addSpeciesCurrentDensity(*sp[r]);
    gx[r] = new GhostExchange(*mesh[r], lay, r);
    gxr[r] = new GhostExchange(*mesh[r], lay, r, NODES);
    mg[r] = new ParticleMigration(*sp[r], lay, r, 4096);
  }
  GhostExchange::connectLocal(gx);
  GhostExchange::connectLocal(gxr);
  ParticleMigration::connectLocal(mg);
  // several boxes per process: every box sends before any box receives
  for (int r = 0; r < world; ++r) gx[r]->begin();
  for (int ph = 0; ph < gx[0]->numPhases(); ++ph) {
    for (int r = 0; r < world; ++r) gx[r]->send(ph);
    for (int r = 0; r < world; ++r) gx[r]->recvAdd(ph);
  }
  double worst = 0.0, scale = 0.0;
  for (int c = 0; c < 3; ++c)
    for (size_t k = 0; k < Jg[c].a.size(); ++k) scale = std::fmax(scale, std::fabs(Jg[c].a[k]));
  for (int r = 0; r < world; ++r) {
    mesh[r]->finalizeSettingJ();
    for (int c = 0; c < 3; ++c) {
      Fab f = fetchJ(*mesh[r], c);
      for (int j = f.lo[1]; j <= f.hi[1]; ++j)
        for (int i = f.lo[0]; i <= f.hi[0]; ++i) {
          const int gi = ((i % NC[0]) + NC[0]) % NC[0], gj = ((j % NC[1]) + NC[1]) % NC[1];
          worst = std::fmax(worst, std::fabs(f.at(i, j) - Jg[c].at(gi, gj)) / scale);
        }
    }
  }
  // the same for the nodal charge density: deposit per box, add-exchange of the ghost layers, read
  double worst_rho = 0.0, scale_rho = 0.0;
  for (size_t k = 0; k < rho_g.a.size(); ++k) scale_rho = std::fmax(scale_rho, std::fabs(rho_g.a[k]));
  for (int r = 0; r < world; ++r) sp[r]->depositChargeDensity(NODES);
  for (int r = 0; r < world; ++r) gxr[r]->begin();
  for (int ph = 0; ph < gxr[0]->numPhases(); ++ph) {
    for (int r = 0; r < world; ++r) gxr[r]->send(ph);
    for (int r = 0; r < world; ++r) gxr[r]->recvAdd(ph);
  }
  for (int r = 0; r < world; ++r) {
    int lo[2], hi[2];
    lay.box(r, lo, hi);
    Fab f;
    for (int d = 0; d < 2; ++d) { f.lo[d] = lo[d] - NG; f.hi[d] = hi[d] + NG + 1; }
    f.a.assign((size_t)(f.hi[0] - f.lo[0] + 1) * (f.hi[1] - f.lo[1] + 1), 0.0);
    sp[r]->getChargeDensity(NODES, f.ref());
    for (int j = f.lo[1]; j <= f.hi[1]; ++j)
      for (int i = f.lo[0]; i <= f.hi[0]; ++i) {
        const int gi = ((i % NC[0]) + NC[0]) % NC[0], gj = ((j % NC[1]) + NC[1]) % NC[1];
        worst_rho = std::fmax(worst_rho, std::fabs(f.at(i, j) - rho_g.at(gi, gj)) / scale_rho);
      }
  }
  // migration: periodic wrap of the new positions, then every leaver to the box that owns it
  const int bc[2] = {PGPU_BC_PERIODIC, PGPU_BC_PERIODIC};
  for (int r = 0; r < world; ++r) sp[r]->applyBCs(bc, bc);
  for (int r = 0; r < world; ++r) mg[r]->send();
  for (int r = 0; r < world; ++r) mg[r]->recv();
  long moved = 0, total = 0, misplaced = 0;
  unsigned long long idsum = 0, idsum0 = 0;
  for (long p = 0; p < n; ++p) idsum0 += id[p];
  for (int r = 0; r < world; ++r) moved += mg[r]->finish();
  for (int r = 0; r < world; ++r) {
    const long m = sp[r]->numParticles();
    std::vector<double> X(2 * m), XO(2 * m), V(3 * m), VO(3 * m), W(m);
    std::vector<uint64_t> ID(m);
    sp[r]->getParticles(X.data(), XO.data(), V.data(), VO.data(), W.data(), ID.data());
    int lo[2], hi[2];
    lay.box(r, lo, hi);
    for (long p = 0; p < m; ++p) {
      idsum += ID[p];
      for (int d = 0; d < 2; ++d) {
        const int cell = (int)std::floor((X[d * m + p] - XMIN[d]) / DX[d]);
        if (cell < lo[d] || cell > hi[d]) ++misplaced;
      }
    }
    total += m;
  }
  std::printf("max_rel_err_J %.3e max_rel_err_rho %.3e moved %ld total %ld misplaced %ld ids_ok %d\n", worst, worst_rho, moved,
              total, misplaced, (int)(idsum == idsum0));
  for (int r = 0; r < world; ++r) {
    delete mg[r];
    delete gx[r];
    delete gxr[r];
    delete sp[r];
    delete mesh[r];
  }
  finalize();
  return (worst < 1e-13 && worst_rho < 1e-13 && total == n && misplaced == 0 && moved > 100 && idsum == idsum0) ? 0 : 1;
}
