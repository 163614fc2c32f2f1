// PicGpuHost.cpp -- see PicGpuHost.H.  Every method is a forward to the C ABI; nothing here
// computes on the CPU.
#include "PicGpuHost.H"

#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace picnic_gpu {

static void default_fatal(const char *msg) {
  std::fprintf(stderr, "picnic_gpu: %s\n", msg);   // MayDay::Error prints and aborts the run
  std::exit(EXIT_FAILURE);
}
static FatalHandler g_fatal = default_fatal;
void setFatalHandler(FatalHandler h) { g_fatal = h ? h : default_fatal; }
void fatal(const std::string &msg) { g_fatal(msg.c_str()); }
void check(int rc, const char *where) {
  if (rc != 0) fatal(std::string(where) + ": " + pgpu_last_error());
}

void initialize(int device) { check(pgpu_init(device), "picnic_gpu::initialize"); }
void finalize() { pgpu_finalize(); }

// ---- Mesh -----------------------------------------------------------------------------------
Mesh::Mesh(int D, const int *num_cells, const Real *Xmin, const Real *dX, int num_ghosts, const int *is_periodic,
           const int *box_lo, const int *box_hi, Real volume_scale)
    : m_h(nullptr), m_D(D) {
  pgpu_grid_desc d;
  std::memset(&d, 0, sizeof(d));
  d.D = D;
  for (int k = 0; k < D; ++k) {
    d.ncell[k] = num_cells[k];
    d.xmin[k] = Xmin[k];
    d.dx[k] = dX[k];
    d.periodic[k] = is_periodic[k];
    d.box_lo[k] = box_lo[k];
    d.box_hi[k] = box_hi[k];
  }
  d.nghost = num_ghosts;
  d.volume_scale = volume_scale;
  check(pgpu_grid_create(&d, &m_h), "Mesh::Mesh");
}
Mesh::~Mesh() { pgpu_grid_destroy(m_h); }

void Mesh::setEMfields(const FabRef &Ex, const FabRef &Ey, const FabRef &Ez, const FabRef &Bx, const FabRef &By,
                       const FabRef &Bz) {
  const FabRef *f[6] = {&Ex, &Ey, &Ez, &Bx, &By, &Bz};
  for (int c = 0; c < 6; ++c) check(pgpu_fields_set(m_h, c, f[c]->data, f[c]->lo, f[c]->hi), "Mesh::setEMfields");
  check(pgpu_synchronize(), "Mesh::setEMfields");   // the host arrays may change after this returns
}
void Mesh::fieldBounds(int comp, int *lo, int *hi) const {
  int l[2], h[2];
  check(pgpu_field_bounds(m_h, comp, l, h), "Mesh::fieldBounds");
  for (int d = 0; d < m_D; ++d) {
    lo[d] = l[d];
    hi[d] = h[d];
  }
}
void Mesh::zeroCurrentDensity() { check(pgpu_current_zero(m_h), "Mesh::zeroCurrentDensity"); }
void Mesh::addSpeciesCurrentDensity(const PicChargedSpecies &sp) {
  check(pgpu_current_add_species(m_h, sp.handle()), "Mesh::addSpeciesCurrentDensity");
}
void Mesh::finalizeSettingJ() { check(pgpu_current_finalize(m_h), "Mesh::finalizeSettingJ"); }
void Mesh::getCurrentDensity(int comp, const FabRef &out) const {
  check(pgpu_current_get(m_h, comp, out.data, out.lo, out.hi), "Mesh::getCurrentDensity");
}
void Mesh::setDebyeLength(const std::vector<PicChargedSpecies *> &species, Real *LDe) {
  std::vector<pgpu_species_t> h;
  for (size_t i = 0; i < species.size(); ++i) h.push_back(species[i]->handle());
  check(pgpu_debye_length(m_h, h.data(), (int)h.size(), LDe), "Mesh::setDebyeLength");
}

// ---- PicChargedSpecies ------------------------------------------------------------------------
PicChargedSpecies::PicChargedSpecies(Mesh &a_mesh, const std::string &a_name, Real a_mass, Real a_charge,
                                     Real a_fnorm_const, Real a_cvac_norm, InterpType a_interpRhoToGrid,
                                     InterpType a_interpJToGrid, InterpType a_interpEToParts)
    : m_mesh(a_mesh), m_name(a_name), m_h(nullptr), m_stable_dt(DBL_MAX), m_num_parts_its(0), m_num_apply_its(0),
      m_num_unconverged(0) {
  std::memset(&m_desc, 0, sizeof(m_desc));
  m_desc.mass = a_mass;
  m_desc.charge = a_charge;
  m_desc.fnorm_const = a_fnorm_const;
  m_desc.cvac_norm = a_cvac_norm;
  m_desc.interp_N = a_interpRhoToGrid;
  m_desc.interp_J = a_interpJToGrid;
  m_desc.interp_E = a_interpEToParts;
  m_desc.rtol = 1.0e-12;     // PicChargedSpecies.cpp defaults (pic_species.rtol_particles, iter_max_particles)
  m_desc.iter_max = 0;
  m_desc.order_swap = 0;
  m_desc.motion = 1;
  m_desc.forces = 1;
  check(pgpu_species_create(m_mesh.handle(), &m_desc, &m_h), "PicChargedSpecies::PicChargedSpecies");
}
PicChargedSpecies::~PicChargedSpecies() { pgpu_species_destroy(m_h); }

void PicChargedSpecies::setParticleSolverParams(bool, bool a_iter_order_swap, int a_iter_max, Real a_rtol, int, int) {
  m_desc.order_swap = a_iter_order_swap ? 1 : 0;
  m_desc.iter_max = a_iter_max;
  m_desc.rtol = a_rtol;
  check(pgpu_species_set_solver_params(m_h, m_desc.order_swap, a_iter_max, a_rtol),
        "PicChargedSpecies::setParticleSolverParams");
}
int PicChargedSpecies::numParticles() const { return (int)pgpu_species_count(m_h); }
void PicChargedSpecies::setParticles(long n, const Real *x, const Real *xold, const Real *v, const Real *vold,
                                     const Real *w, const uint64_t *id) {
  check(pgpu_species_upload(m_h, n, x, xold, v, vold, w, id), "PicChargedSpecies::setParticles");
}
void PicChargedSpecies::getParticles(Real *x, Real *xold, Real *v, Real *vold, Real *w, uint64_t *id) const {
  check(pgpu_species_download(m_h, x, xold, v, vold, w, id), "PicChargedSpecies::getParticles");
}

#define FWD0(method, call) \
  void PicChargedSpecies::method() { check(call(m_h), "PicChargedSpecies::" #method); }
FWD0(advancePositions_2ndHalf, pgpu_advance_positions_2nd_half)
FWD0(advanceVelocities_2ndHalf, pgpu_advance_velocities_2nd_half)
FWD0(averageVelocities, pgpu_average_velocities)
FWD0(updateOldParticlePositions, pgpu_update_old_particle_positions)
FWD0(updateOldParticleVelocities, pgpu_update_old_particle_velocities)
FWD0(resetParticles, pgpu_reset_particles)
FWD0(interpolateFieldsToParticles, pgpu_interpolate_fields_to_particles)
FWD0(binTheParticles, pgpu_bin_particles)
FWD0(setNumberDensityFromBinFab, pgpu_set_moments_from_bins)
#undef FWD0

void PicChargedSpecies::advancePositionsExplicit(Real a_full_dt, bool a_half_step) {
  check(pgpu_advance_positions_explicit(m_h, a_full_dt, a_half_step ? 1 : 0), "PicChargedSpecies::advancePositionsExplicit");
}
void PicChargedSpecies::advancePositionsImplicit(Real a_full_dt) {
  check(pgpu_advance_positions_implicit(m_h, a_full_dt), "PicChargedSpecies::advancePositionsImplicit");
}
void PicChargedSpecies::advanceVelocities(Real a_full_dt, bool a_half_step) {
  check(pgpu_advance_velocities(m_h, a_full_dt, a_half_step ? 1 : 0), "PicChargedSpecies::advanceVelocities");
}
void PicChargedSpecies::advanceParticles(Real a_dt) {
  check(pgpu_advance_particles(m_h, a_dt), "PicChargedSpecies::advanceParticles");
}
void PicChargedSpecies::advanceParticlesIteratively(Real a_dt, bool a_deposit_current) {
  pgpu_picard_stats st = {0, 0, 0};
  const int rc = pgpu_advance_particles_iteratively(m_h, a_dt, a_deposit_current ? 1 : 0, &st);
  if (rc == PGPU_ERR_SEGMENTS)   // the reference's Fortran STOP (MeshInterpChargeConservingF.ChF:1018-1021)
    fatal("cc1 deposit: particle crossing more cells than num_ghosts allows (" + m_name + ")");
  check(rc, "PicChargedSpecies::advanceParticlesIteratively");
  m_num_parts_its += (uint64_t)st.num_parts_its;
  m_num_apply_its += (uint64_t)st.num_apply_its;
  m_num_unconverged = (uint64_t)st.num_unconverged;
}
void PicChargedSpecies::setCurrentDensity(Real a_dt, bool a_from_explicit_solver) {
  check(pgpu_set_current_density(m_h, a_dt, a_from_explicit_solver ? 1 : 0), "PicChargedSpecies::setCurrentDensity");
}
void PicChargedSpecies::getCurrentDensity(int comp, const FabRef &out) const {
  check(pgpu_species_current_get(m_h, comp, out.data, out.lo, out.hi), "PicChargedSpecies::getCurrentDensity");
}
void PicChargedSpecies::setChargeDensity(const FabRef &out) {
  const int stag[2] = {0, 0};
  check(pgpu_set_charge_density(m_h, stag, out.data, out.lo, out.hi), "PicChargedSpecies::setChargeDensity");
}
void PicChargedSpecies::setChargeDensityOnFaces(int dir, const FabRef &out) {
  const int stag[2] = {dir == 0 ? 1 : 0, dir == 1 ? 1 : 0};
  check(pgpu_set_charge_density(m_h, stag, out.data, out.lo, out.hi), "PicChargedSpecies::setChargeDensityOnFaces");
}
void PicChargedSpecies::setChargeDensityOnNodes(const FabRef &out) {
  const int stag[2] = {1, 1};
  check(pgpu_set_charge_density(m_h, stag, out.data, out.lo, out.hi), "PicChargedSpecies::setChargeDensityOnNodes");
}
void PicChargedSpecies::getMomentsFromBinFab(Real *dens, Real *mom, Real *ene) const {
  check(pgpu_species_moments_get(m_h, dens, mom, ene), "PicChargedSpecies::getMomentsFromBinFab");
}
void PicChargedSpecies::applyBCs(const int *bc_lo, const int *bc_hi) {
  check(pgpu_apply_bcs(m_h, bc_lo, bc_hi), "PicChargedSpecies::applyBCs");
}
void PicChargedSpecies::finishImplicitStep(const int *bc_lo, const int *bc_hi) {
  check(pgpu_finish_implicit_step(m_h, bc_lo, bc_hi), "PicChargedSpecies::finishImplicitStep");
}
void PicChargedSpecies::setStableDt() { check(pgpu_stable_dt(m_h, &m_stable_dt), "PicChargedSpecies::setStableDt"); }
void PicChargedSpecies::globalMoments(std::array<Real, 7> &a_moments) const {
  check(pgpu_global_moments(m_h, a_moments.data()), "PicChargedSpecies::globalMoments");
}
void PicChargedSpecies::picardParams(uint64_t &a_num_parts_its, uint64_t &a_num_apply_its) const {
  a_num_parts_its = m_num_parts_its;
  a_num_apply_its = m_num_apply_its;
}

// ---- scattering --------------------------------------------------------------------------------
uint64_t Scattering::s_seed = 1983, Scattering::s_step = 0;
void Scattering::setRandomState(uint64_t a_seed, uint64_t a_step) {
  s_seed = a_seed;
  s_step = a_step;
}

TakizukaAbe::TakizukaAbe(int a_sp1, int a_sp2, Real a_Clog)
    : m_sp1(a_sp1), m_sp2(a_sp2), m_Clog(a_Clog), m_scatter_dt(DBL_MAX), m_npairs(0) {
  // TakizukaAbe.H:50 asserts Clog >= 3 (assert off in OPT builds); report instead of ignoring
  if (a_Clog < 3.0) std::cerr << "picnic_gpu::TakizukaAbe: coulomb_logarithm " << a_Clog << " < 3" << std::endl;
}
void TakizukaAbe::setMeanFreeTime(const std::vector<PicChargedSpecies *> &a_species) const {
  // m_scatter_dt = 1/nu_max over the cells of this box (TakizukaAbe.cpp:55-238); needs the cell
  // moments of prepForScatter.  nu_max == 0 (no cell holds both species) leaves the dt unlimited.
  double nu = 0.0;
  check(pgpu_scatter_nu_max_ta(a_species[m_sp1]->handle(), a_species[m_sp2]->handle(), m_Clog, &nu),
        "TakizukaAbe::setMeanFreeTime");
  m_scatter_dt = nu > 0.0 ? 1.0 / nu : DBL_MAX;
}
void TakizukaAbe::applyScattering(std::vector<PicChargedSpecies *> &a_species, Real a_dt_sec) const {
  PicChargedSpecies *a = a_species[m_sp1], *b = a_species[m_sp2];
  if (a->numParticles() == 0 || b->numParticles() == 0) return;
  long np = 0;
  // prepForScatter (PicSpeciesInterface.cpp:1593-1625) binned the particles and set the moments
  check(pgpu_collide_ta(a->handle(), b->handle(), m_Clog, a_dt_sec, s_seed, s_step, &np), "TakizukaAbe::applyScattering");
  m_npairs = np;
}
void TakizukaAbe::printParameters() const {
  std::cout << " TakizukaAbe scattering parameters:" << std::endl;
  std::cout << "  species A = " << m_sp1 << ", species B = " << m_sp2 << std::endl;
  std::cout << "  Coulomb Logarithm = " << m_Clog << std::endl;
}

Coulomb::Coulomb(int a_sp1, int a_sp2, Real a_Clog, AngularScattering a_angular, bool a_NxN, int a_NxN_Nthresh,
                 int a_num_subcycles)
    : m_sp1(a_sp1), m_sp2(a_sp2), m_scatter_dt(DBL_MAX), m_npairs(0) {
  m_prm.Clog = a_Clog;
  m_prm.angular_scattering = a_angular;
  m_prm.NxN = a_NxN ? 1 : 0;
  m_prm.NxN_Nthresh = a_NxN_Nthresh;
  m_prm.num_subcycles = a_num_subcycles;
  if (a_Clog != 0.0 && a_Clog < 2.0) fatal("Coulomb: coulomb_logarithm must be 0 (computed) or >= 2");   // Coulomb.H:224
}
void Coulomb::setMeanFreeTime(const std::vector<PicChargedSpecies *> &a_species) const {
  // Coulomb.cpp:79-356; needs Mesh::setDebyeLength and the cell moments
  double nu = 0.0;
  check(pgpu_scatter_nu_max_coulomb(a_species[m_sp1]->handle(), a_species[m_sp2]->handle(), &m_prm, &nu),
        "Coulomb::setMeanFreeTime");
  m_scatter_dt = nu > 0.0 ? 1.0 / nu : DBL_MAX;
}
void Coulomb::applyScattering(std::vector<PicChargedSpecies *> &a_species, Real a_dt_sec) const {
  PicChargedSpecies *a = a_species[m_sp1], *b = a_species[m_sp2];
  if (a->numParticles() == 0 || b->numParticles() == 0) return;
  long np = 0;
  check(pgpu_collide_coulomb(a->handle(), b->handle(), &m_prm, a_dt_sec, s_seed, s_step, &np), "Coulomb::applyScattering");
  m_npairs = np;
}
void Coulomb::printParameters() const {
  std::cout << " Coulomb scattering parameters:" << std::endl;
  std::cout << "  species A = " << m_sp1 << ", species B = " << m_sp2 << std::endl;
  std::cout << "  Coulomb Logarithm = " << m_prm.Clog << (m_prm.Clog == 0.0 ? " (computed per pair)" : "") << std::endl;
  std::cout << "  angular scattering = " << m_prm.angular_scattering << ", weight method = PROBABILISTIC" << std::endl;
  std::cout << "  NxN pairings = " << (m_prm.NxN ? "true" : "false") << ", NxN Nthresh = " << m_prm.NxN_Nthresh << std::endl;
}

Elastic::Elastic(int a_sp1, int a_sp2, Real a_const_sigma)
    : m_sp1(a_sp1), m_sp2(a_sp2), m_const_sigma(a_const_sigma), m_okhrimovskyy(false), m_loglog(false), m_scatter_dt(DBL_MAX), m_ncoll(0) {}
Elastic::Elastic(int a_sp1, int a_sp2, const std::vector<Real> &a_E_eV, const std::vector<Real> &a_Q,
                 const std::vector<Real> &a_xi, bool a_okhrimovskyy, bool a_use_loglog_interp)
    : m_sp1(a_sp1), m_sp2(a_sp2), m_const_sigma(0.0), m_E(a_E_eV), m_Q(a_Q), m_xi(a_xi),
      m_okhrimovskyy(a_okhrimovskyy), m_loglog(a_use_loglog_interp), m_scatter_dt(DBL_MAX), m_ncoll(0) {
  if (m_E.size() < 2 || m_Q.size() != m_E.size() || (a_okhrimovskyy && m_xi.size() != m_E.size()))
    fatal("Elastic: cross-section table columns differ in length");
}
pgpu_elastic_params Elastic::params() const {
  pgpu_elastic_params prm;
  prm.const_sigma = m_const_sigma;
  prm.ntab = (int)m_E.size();
  prm.E = m_E.empty() ? nullptr : m_E.data();
  prm.Q = m_Q.empty() ? nullptr : m_Q.data();
  prm.xi = m_xi.empty() ? nullptr : m_xi.data();
  prm.angular_scattering = m_okhrimovskyy ? 1 : 0;
  prm.use_loglog_interp = m_loglog ? 1 : 0;
  return prm;
}
void Elastic::setMeanFreeTime(const std::vector<PicChargedSpecies *> &a_species) const {
  // Elastic.cpp:122-202
  const pgpu_elastic_params prm = params();
  double nu = 0.0;
  check(pgpu_scatter_nu_max_elastic(a_species[m_sp1]->handle(), a_species[m_sp2]->handle(), &prm, &nu),
        "Elastic::setMeanFreeTime");
  m_scatter_dt = nu > 0.0 ? 1.0 / nu : DBL_MAX;
}
void Elastic::applyScattering(std::vector<PicChargedSpecies *> &a_species, Real a_dt_sec) const {
  PicChargedSpecies *a = a_species[m_sp1], *b = a_species[m_sp2];
  if (a->numParticles() == 0 || b->numParticles() == 0) return;
  const pgpu_elastic_params prm = params();
  long nc = 0;
  check(pgpu_collide_elastic(a->handle(), b->handle(), &prm, a_dt_sec, s_seed, s_step, &nc), "Elastic::applyScattering");
  m_ncoll = nc;
}
void Elastic::printParameters() const {
  std::cout << " Elastic scattering parameters:" << std::endl;
  std::cout << "  species A = " << m_sp1 << ", species B = " << m_sp2 << std::endl;
  if (m_E.empty()) std::cout << "  constant cross section = " << m_const_sigma << " m^2" << std::endl;
  else std::cout << "  tabulated cross section, " << m_E.size() << " rows, "
                 << (m_okhrimovskyy ? "OKHRIMOVSKYY" : "ISOTROPIC") << std::endl;
}

}  // namespace picnic_gpu
