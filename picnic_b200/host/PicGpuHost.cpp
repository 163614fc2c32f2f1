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
    : m_h(nullptr), m_D(D), m_gx(nullptr) {
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
long Mesh::packedFieldSize() const {
  long n = 0;
  check(pgpu_fields_packed_size(m_h, &n), "Mesh::packedFieldSize");
  return n;
}
long Mesh::packedCurrentSize() const {
  long n = 0;
  check(pgpu_current_packed_size(m_h, &n), "Mesh::packedCurrentSize");
  return n;
}
void Mesh::setEMfieldsPacked(const Real *six_components) {
  check(pgpu_fields_set_packed(m_h, six_components), "Mesh::setEMfieldsPacked");
}
void Mesh::getCurrentDensityPacked(Real *three_components) const {
  check(pgpu_current_get_packed_async(m_h, three_components), "Mesh::getCurrentDensityPacked");
}
void Mesh::setExternalFields(const pgpu_ext_fn *six_or_null) {
  check(pgpu_grid_set_external_fields(m_h, six_or_null), "Mesh::setExternalFields");
}
void Mesh::addSubOrbitJ(const PicChargedSpecies &sp) {
  check(pgpu_current_add_suborbit(m_h, sp.handle()), "Mesh::addSubOrbitJ");
}
void Mesh::addInflowJ(const PicChargedSpecies &sp) {
  check(pgpu_current_add_inflow(m_h, sp.handle()), "Mesh::addInflowJ");
}
void Mesh::filterJ(bool filterE_inPlane, bool filterE_virtual) {
  check(pgpu_current_filter(m_h, filterE_inPlane ? 1 : 0, filterE_virtual ? 1 : 0), "PicSpeciesInterface::filterJ");
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
void Mesh::finalizeSettingJ() {
  if (m_gx) m_gx->addExchange();
  check(pgpu_current_finalize(m_h), "Mesh::finalizeSettingJ");
}
void Mesh::getCurrentDensity(int comp, const FabRef &out) const {
  check(pgpu_current_get(m_h, comp, out.data, out.lo, out.hi), "Mesh::getCurrentDensity");
}
void Mesh::initializeMassMatrices(int a_interp_type, int *a_ncomp) {
  check(pgpu_mass_matrices_init(m_h, a_interp_type, a_ncomp), "Mesh::initializeMassMatrices");
}
void Mesh::setMassMatrices(const std::vector<PicChargedSpecies *> &species, Real a_dt) {
  check(pgpu_mass_matrices_zero(m_h), "Mesh::setMassMatrices");
  for (size_t i = 0; i < species.size(); ++i) species[i]->accumulateMassMatrices(a_dt);
  check(pgpu_mass_matrices_save_E0(m_h), "Mesh::setMassMatrices");
}
void Mesh::computeJfromMassMatrices() {
  check(pgpu_compute_J_from_mass_matrices(m_h), "Mesh::computeJfromMassMatrices");
}
void Mesh::getMassMatrix(int a_which, const FabRef &out, int a_ncomp) const {
  check(pgpu_mass_matrix_get(m_h, a_which, out.data, out.lo, out.hi, a_ncomp), "Mesh::getMassMatrix");
}
void Mesh::getJ0(int comp, const FabRef &out) const {
  check(pgpu_mass_matrix_J0_get(m_h, comp, out.data, out.lo, out.hi), "Mesh::getJ0");
}
void Mesh::setDebyeLength(const std::vector<PicChargedSpecies *> &species, Real *LDe) {
  std::vector<pgpu_species_t> h;
  for (size_t i = 0; i < species.size(); ++i) h.push_back(species[i]->handle());
  check(pgpu_debye_length(m_h, h.data(), (int)h.size(), LDe), "Mesh::setDebyeLength");
}

// ---- BoxLayout / GhostExchange / ParticleMigration ---------------------------------------------
BoxLayout::BoxLayout(int a_D, const int *a_num_cells, const int *a_box_cells, int a_num_ghosts, const int *a_is_periodic)
    : D(a_D), num_ghosts(a_num_ghosts) {
  for (int d = 0; d < 2; ++d) {
    num_cells[d] = d < D ? a_num_cells[d] : 1;
    box_cells[d] = d < D ? a_box_cells[d] : 1;
    is_periodic[d] = d < D ? a_is_periodic[d] : 0;
    if (num_cells[d] % box_cells[d]) fatal("BoxLayout: boxes must tile the domain");
    nb[d] = num_cells[d] / box_cells[d];
    if (d < D && nb[d] > 1 && box_cells[d] < 2 * num_ghosts + 1) fatal("BoxLayout: box narrower than its ghost overlap");
  }
}
void BoxLayout::box(int rank, int *lo, int *hi) const {
  const int c[2] = {rank % nb[0], rank / nb[0]};
  for (int d = 0; d < D; ++d) {
    lo[d] = c[d] * box_cells[d];
    hi[d] = lo[d] + box_cells[d] - 1;
  }
}
int BoxLayout::neighborCode(int rank, int code) const {
  int c[2] = {rank % nb[0], rank / nb[0]};
  const int off[2] = {code % 3 - 1, code / 3 - 1};
  for (int d = 0; d < D; ++d) {
    c[d] += off[d];
    if (c[d] < 0 || c[d] >= nb[d]) {
      if (!is_periodic[d]) return -1;
      c[d] = (c[d] + nb[d]) % nb[d];
    }
  }
  return c[0] + (D == 2 ? c[1] * nb[0] : 0);
}
int BoxLayout::neighbor(int rank, int dir, int side) const {
  return neighborCode(rank, 4 + side * (dir == 0 ? 1 : 3));
}
void BoxLayout::overlap(int rank, const int *stag, int dir, int side, int *lo, int *hi) const {
  int blo[2] = {0, 0}, bhi[2] = {0, 0};
  box(rank, blo, bhi);
  const int g = num_ghosts;
  for (int d = 0; d < D; ++d) {
    lo[d] = blo[d] - g;
    hi[d] = bhi[d] + g + stag[d];
  }
  if (side > 0) {
    lo[dir] = bhi[dir] + 1 - g;
    hi[dir] = bhi[dir] + g + stag[dir];
  } else {
    lo[dir] = blo[dir] - g;
    hi[dir] = blo[dir] - 1 + g + stag[dir];
  }
}

GhostExchange::GhostExchange(Mesh &mesh, const BoxLayout &layout, int rank, const int *a_rho_stag)
    : m_mesh(mesh), m_layout(layout), m_rank(rank), m_h(nullptr), m_for_rho(a_rho_stag != nullptr) {
  static const int STAG_J[2][3][2] = {{{0, 0}, {1, 0}, {1, 0}}, {{0, 1}, {1, 0}, {1, 1}}};   // Jx, Jy, Jz centring
  const int D = layout.D;
  std::vector<pgpu_halo_msg> recs;
  int phase = 0;
  for (int d = 0; d < D; ++d) {
    if (layout.nb[d] == 1) continue;      // spans the domain: folded locally by pgpu_current_finalize
    for (int side = -1; side <= 1; side += 2) {
      const int peer = layout.neighbor(rank, d, side);
      if (peer < 0) continue;
      pgpu_halo_msg m;
      std::memset(&m, 0, sizeof(m));
      m.phase = phase;
      m.recv_area = (int)recs.size();
      if (a_rho_stag) layout.overlap(rank, a_rho_stag, d, side, m.lo[0], m.hi[0]);
      else for (int c = 0; c < 3; ++c) layout.overlap(rank, STAG_J[D - 1][c], d, side, m.lo[c], m.hi[c]);
      recs.push_back(m);
      Msg q = {phase, side, peer, 0};
      m_msgs.push_back(q);
    }
    ++phase;
  }
  if (a_rho_stag) {
    const int st[2] = {a_rho_stag[0], D == 2 ? a_rho_stag[1] : 0};
    check(pgpu_halo_create_rho(mesh.handle(), st, (int)recs.size(), recs.data(), &m_h), "GhostExchange::GhostExchange");
  } else {
    check(pgpu_halo_create(mesh.handle(), (int)recs.size(), recs.data(), &m_h), "GhostExchange::GhostExchange");
  }
  for (size_t i = 0; i < m_msgs.size(); ++i)
    check(pgpu_halo_area_offset(m_h, (int)i, &m_msgs[i].offset, nullptr), "GhostExchange::GhostExchange");
}
GhostExchange::~GhostExchange() {
  if (!m_for_rho) m_mesh.setGhostExchange(nullptr);
  pgpu_halo_destroy(m_h);
}
int GhostExchange::numPhases() const { return pgpu_halo_phases(m_h); }
void GhostExchange::begin() { check(pgpu_halo_begin(m_h), "GhostExchange::begin"); }
void GhostExchange::send(int phase) { check(pgpu_halo_send(m_h, phase), "GhostExchange::send"); }
void GhostExchange::recvAdd(int phase) { check(pgpu_halo_recv_add(m_h, phase), "GhostExchange::recvAdd"); }
void GhostExchange::addExchange() {
  begin();
  for (int ph = 0; ph < numPhases(); ++ph) {
    send(ph);
    recvAdd(ph);
  }
}
std::vector<long> GhostExchange::areaTable() const {
  std::vector<long> t;
  for (size_t i = 0; i < m_msgs.size(); ++i) {
    t.push_back(m_msgs[i].phase);
    t.push_back(m_msgs[i].side);
    t.push_back((long)i);
    t.push_back(m_msgs[i].offset);
  }
  return t;
}
void GhostExchange::connectWith(const std::vector<void *> &inbox_of_rank,
                                const std::vector<std::vector<long> > &table_of_rank) {
  for (size_t i = 0; i < m_msgs.size(); ++i) {
    const Msg &q = m_msgs[i];
    const std::vector<long> &t = table_of_rank[q.peer];
    bool found = false;
    for (size_t k = 0; k + 3 < t.size(); k += 4)
      if (t[k] == q.phase && t[k + 1] == -q.side) {      // my +side message is the peer's -side arrival
        check(pgpu_halo_connect(m_h, (int)i, inbox_of_rank[q.peer], (int)t[k + 2], t[k + 3]), "GhostExchange::connect");
        found = true;
      }
    if (!found) fatal("GhostExchange::connect: the neighbour has no matching message");
  }
}
void GhostExchange::connectLocal(const std::vector<GhostExchange *> &all) {
  std::vector<void *> inbox(all.size(), nullptr);
  std::vector<std::vector<long> > table(all.size());
  for (size_t k = 0; k < all.size(); ++k) {
    const int r = all[k]->m_rank;
    check(pgpu_halo_inbox(all[k]->m_h, &inbox[r], nullptr), "GhostExchange::connectLocal");
    table[r] = all[k]->areaTable();
  }
  for (size_t k = 0; k < all.size(); ++k) all[k]->connectWith(inbox, table);
}
void GhostExchange::connect(AllGatherFn allgather, void *user) {
  // record per rank: 64-byte IPC handle + message count + (phase, side, area, offset) of up to 16 messages
  const int world = m_layout.numBoxes();
  const size_t NREC = 8 + 1 + 4 * 16;
  std::vector<long> mine(NREC, 0), all(NREC * world, 0);
  check(pgpu_halo_ipc_handle(m_h, mine.data()), "GhostExchange::connect");
  const std::vector<long> t = areaTable();
  mine[8] = (long)m_msgs.size();
  for (size_t k = 0; k < t.size(); ++k) mine[9 + k] = t[k];
  allgather(mine.data(), all.data(), NREC * sizeof(long), user);
  std::vector<void *> inbox(world, nullptr);
  std::vector<std::vector<long> > table(world);
  for (size_t i = 0; i < m_msgs.size(); ++i) {
    const int r = m_msgs[i].peer;
    if (inbox[r]) continue;
    if (r == m_rank) check(pgpu_halo_inbox(m_h, &inbox[r], nullptr), "GhostExchange::connect");
    else check(pgpu_halo_ipc_open(m_h, &all[NREC * r], &inbox[r]), "GhostExchange::connect");
    table[r].assign(all.begin() + NREC * r + 9, all.begin() + NREC * r + 9 + 4 * all[NREC * r + 8]);
  }
  connectWith(inbox, table);
}

ParticleMigration::ParticleMigration(PicChargedSpecies &species, const BoxLayout &layout, int rank,
                                     long capacity_records)
    : m_layout(layout), m_rank(rank), m_h(nullptr), m_lost(0) {
  check(pgpu_migrator_create(species.handle(), capacity_records, &m_h), "ParticleMigration::ParticleMigration");
}
ParticleMigration::~ParticleMigration() { pgpu_migrator_destroy(m_h); }
void ParticleMigration::connectLocal(const std::vector<ParticleMigration *> &all) {
  std::vector<void *> inbox(all.size(), nullptr);
  for (size_t k = 0; k < all.size(); ++k)
    check(pgpu_migrator_inbox(all[k]->m_h, &inbox[all[k]->m_rank], nullptr), "ParticleMigration::connectLocal");
  for (size_t k = 0; k < all.size(); ++k)
    for (int code = 0; code < 9; ++code) {
      const int peer = code == 4 ? -1 : all[k]->m_layout.neighborCode(all[k]->m_rank, code);
      if (peer >= 0 && peer != all[k]->m_rank)
        check(pgpu_migrator_connect(all[k]->m_h, code, inbox[peer]), "ParticleMigration::connectLocal");
    }
}
void ParticleMigration::connect(AllGatherFn allgather, void *user) {
  const int world = m_layout.numBoxes();
  std::vector<long> mine(8, 0), all(8 * world, 0);
  check(pgpu_migrator_ipc_handle(m_h, mine.data()), "ParticleMigration::connect");
  allgather(mine.data(), all.data(), 8 * sizeof(long), user);
  std::vector<void *> inbox(world, nullptr);
  for (int code = 0; code < 9; ++code) {
    const int peer = code == 4 ? -1 : m_layout.neighborCode(m_rank, code);
    if (peer < 0 || peer == m_rank) continue;
    if (!inbox[peer]) check(pgpu_migrator_ipc_open(m_h, &all[8 * peer], &inbox[peer]), "ParticleMigration::connect");
    check(pgpu_migrator_connect(m_h, code, inbox[peer]), "ParticleMigration::connect");
  }
}
void ParticleMigration::send() { check(pgpu_migrate_send(m_h), "ParticleMigration::send"); }
void ParticleMigration::recv() { check(pgpu_migrate_recv(m_h), "ParticleMigration::recv"); }
long ParticleMigration::finish() {
  long a = 0, l = 0;
  check(pgpu_migrate_finish(m_h, &a, &l, &m_lost), "ParticleMigration::finish");
  return a;
}

// ---- PicChargedSpecies ------------------------------------------------------------------------
PicChargedSpecies::PicChargedSpecies(Mesh &a_mesh, const std::string &a_name, Real a_mass, Real a_charge,
                                     Real a_fnorm_const, Real a_cvac_norm, InterpType a_interpRhoToGrid,
                                     InterpType a_interpJToGrid, InterpType a_interpEToParts, bool a_relativistic,
                                     bool a_higuera_cary)
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
  m_desc.relativistic = a_relativistic ? 1 : 0;
  m_desc.higuera_cary = a_higuera_cary ? 1 : 0;
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
void PicChargedSpecies::accumulateMassMatrices(Real a_dt) {
  check(pgpu_accumulate_mass_matrices(m_h, a_dt), "PicChargedSpecies::accumulateMassMatrices");
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
void PicChargedSpecies::depositChargeDensity(const int *stag) {
  check(pgpu_charge_density_deposit(m_h, stag), "PicChargedSpecies::depositChargeDensity");
}
void PicChargedSpecies::getChargeDensity(const int *stag, const FabRef &out) const {
  check(pgpu_charge_density_get(m_mesh.handle(), stag, out.data, out.lo, out.hi), "PicChargedSpecies::getChargeDensity");
}
void PicChargedSpecies::getMomentsFromBinFab(Real *dens, Real *mom, Real *ene) const {
  check(pgpu_species_moments_get(m_h, dens, mom, ene), "PicChargedSpecies::getMomentsFromBinFab");
}
void PicChargedSpecies::explicitStep(Real a_dt, const int *bc_lo, const int *bc_hi, bool a_second_half) {
  check(pgpu_explicit_step(m_h, a_dt, bc_lo, bc_hi, a_second_half ? 1 : 0), "PicChargedSpecies::explicitStep");
}
void PicChargedSpecies::addExternalFieldsToParticles() {
  check(pgpu_add_external_fields_to_particles(m_h), "PicChargedSpecies::addExternalFieldsToParticles");
}
void PicChargedSpecies::applyForcesCurvilinear(int a_push_type, Real a_full_dt, bool a_byHalfDt, bool a_anticyclic) {
  check(pgpu_apply_forces_curvilinear(m_h, a_push_type, a_full_dt, a_byHalfDt ? 1 : 0, a_anticyclic ? 1 : 0),
        "PicChargedSpecies::applyForcesCurvilinear");
}
void PicChargedSpecies::setVirtualPositions(const Real *a_virt) {
  check(pgpu_species_virtual_positions_set(m_h, a_virt), "PicChargedSpecies::setVirtualPositions");
}
void PicChargedSpecies::getVirtualPositions(Real *a_virt) const {
  check(pgpu_species_virtual_positions_get(m_h, a_virt), "PicChargedSpecies::getVirtualPositions");
}
void PicChargedSpecies::setSubOrbitModel(bool a_use_suborbit_model, bool a_suborbit_fast_particles) {
  check(pgpu_species_set_suborbit_model(m_h, a_use_suborbit_model ? 1 : 0, a_suborbit_fast_particles ? 1 : 0),
        "PicChargedSpecies::setSubOrbitModel");
}
void PicChargedSpecies::transferFastParticles() {
  check(pgpu_transfer_fast_particles(m_h), "PicChargedSpecies::transferFastParticles");
}
void PicChargedSpecies::advanceSubOrbitParticlesAndSetJ(Real a_dt, bool a_from_emjacobian) {
  check(pgpu_advance_suborbit_particles_and_set_J(m_h, a_dt, a_from_emjacobian ? 1 : 0),
        "PicChargedSpecies::advanceSubOrbitParticlesAndSetJ");
}
void PicChargedSpecies::mergeSubOrbitParticles() {
  check(pgpu_merge_suborbit_particles(m_h), "PicChargedSpecies::mergeSubOrbitParticles");
}
long PicChargedSpecies::numSubOrbitParticles() const { return pgpu_species_suborbit_count(m_h); }
void PicChargedSpecies::removeOutflowParticles() {
  check(pgpu_remove_outflow_particles(m_h), "PicChargedSpecies::removeOutflowParticles");
}
void PicChargedSpecies::outflowProbes(Real *a_flux20) const {
  check(pgpu_species_outflow_fluxes(m_h, a_flux20), "PicChargedSpecies::outflowProbes");
}
void PicChargedSpecies::injectInflowParticles(long n, const Real *x, const Real *v, const Real *w, const uint64_t *id) {
  check(pgpu_species_append(m_h, n, x, nullptr, v, nullptr, w, id), "PicChargedSpecies::injectInflowParticles");
}
void PicChargedSpecies::addToInflowList(long n, const Real *x, const Real *v, const Real *w, const uint64_t *id,
                                        int a_bdry_dir, int a_bdry_side) {
  check(pgpu_species_inflow_append(m_h, n, x, v, w, id, a_bdry_dir, a_bdry_side), "PicChargedSpecies::addToInflowList");
}
void PicChargedSpecies::advanceInflowParticlesAndSetJ(Real a_dt, bool a_from_emjacobian) {
  check(pgpu_advance_inflow_particles_and_set_J(m_h, a_dt, a_from_emjacobian ? 1 : 0),
        "PicChargedSpecies::advanceInflowParticlesAndSetJ");
}
void PicChargedSpecies::inflowProbes(Real *a_flux20) {
  check(pgpu_species_inflow_fluxes(m_h, a_flux20), "PicChargedSpecies::inflowProbes");
}
long PicChargedSpecies::numInflowParticles() const { return pgpu_species_inflow_count(m_h); }
long PicChargedSpecies::linearSize() const { return pgpu_particle_linear_size(m_h); }
void PicChargedSpecies::getParticlesLinear(void *a_records) const {
  check(pgpu_species_download_linear(m_h, a_records), "PicChargedSpecies::getParticlesLinear");
}
void PicChargedSpecies::setParticlesLinear(long n, const void *a_records) {
  check(pgpu_species_upload_linear(m_h, n, a_records), "PicChargedSpecies::setParticlesLinear");
}
void PicChargedSpecies::sortForLocality() { check(pgpu_sort_for_locality(m_h), "PicChargedSpecies::sortForLocality"); }
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
uint64_t Scattering::s_seed = 1983, Scattering::s_step = 0, Scattering::s_next_stream = 0;
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
  check(pgpu_collide_ta(a->handle(), b->handle(), m_Clog, a_dt_sec, streamSeed(), nextStep(), &np), "TakizukaAbe::applyScattering");
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
  m_prm.enforce_conservations = 0;        // Coulomb.H:325-333 defaults; setEnforceConservations switches it on
  m_prm.energy_fraction = 0.05;
  m_prm.energy_fraction_max = 0.5;
  m_prm.beta_weight_exponent = 1;
  m_prm.sort_weighted_particles = 0;
  m_prm.conservation_Nmin_save = 100000;
  m_prm.weight_method = 0;
  m_prm.include_large_angle_scattering = 0;   // Coulomb.H:329; setLargeAngleScattering switches it on
  m_prm.test_large_angle_draw = 0.5;
  m_prm.test_fas_draw2 = m_prm.test_fas_draw3 = 0.5;
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
  check(pgpu_collide_coulomb(a->handle(), b->handle(), &m_prm, a_dt_sec, streamSeed(), nextStep(), &np), "Coulomb::applyScattering");
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
    : m_sp1(a_sp1), m_sp2(a_sp2), m_const_sigma(a_const_sigma), m_okhrimovskyy(false), m_loglog(false), m_conservative(false), m_scatter_dt(DBL_MAX), m_ncoll(0) {}
Elastic::Elastic(int a_sp1, int a_sp2, const std::vector<Real> &a_E_eV, const std::vector<Real> &a_Q,
                 const std::vector<Real> &a_xi, bool a_okhrimovskyy, bool a_use_loglog_interp)
    : m_sp1(a_sp1), m_sp2(a_sp2), m_const_sigma(0.0), m_E(a_E_eV), m_Q(a_Q), m_xi(a_xi),
      m_okhrimovskyy(a_okhrimovskyy), m_loglog(a_use_loglog_interp), m_conservative(false), m_scatter_dt(DBL_MAX), m_ncoll(0) {
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
  prm.weight_method = m_conservative ? 1 : 0;
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
  check(pgpu_collide_elastic(a->handle(), b->handle(), &prm, a_dt_sec, streamSeed(), nextStep(), &nc), "Elastic::applyScattering");
  m_ncoll = nc;
}
void Elastic::printParameters() const {
  std::cout << " Elastic scattering parameters:" << std::endl;
  std::cout << "  species A = " << m_sp1 << ", species B = " << m_sp2 << std::endl;
  if (m_E.empty()) std::cout << "  constant cross section = " << m_const_sigma << " m^2" << std::endl;
  else std::cout << "  tabulated cross section, " << m_E.size() << " rows, "
                 << (m_okhrimovskyy ? "OKHRIMOVSKYY" : "ISOTROPIC") << std::endl;
}

// ---- HardSphere (HardSphere.cpp:30-52, 65-194, 196-665) ------------------------------------------
HardSphere::HardSphere(int a_sp1, int a_sp2, Real a_r1, Real a_r2, bool a_conservative)
    : m_sp1(a_sp1), m_sp2(a_sp2), m_r1(a_r1), m_r2(a_r2), m_sigmaT(3.14159265358979323846 * (a_r1 + a_r2) * (a_r1 + a_r2)),
      m_conservative(a_conservative), m_scatter_dt(DBL_MAX), m_ncoll(0) {
  if (a_conservative && a_sp1 != a_sp2) fatal("HardSphere: weight_method = CONSERVATIVE is implemented for self-scattering only");
}
void HardSphere::setMeanFreeTime(const std::vector<PicChargedSpecies *> &a_species) const {
  double nu = 0.0;
  check(pgpu_scatter_nu_max_hard_sphere(a_species[m_sp1]->handle(), a_species[m_sp2]->handle(), m_sigmaT, &nu),
        "HardSphere::setMeanFreeTime");
  m_scatter_dt = nu > 0.0 ? 1.0 / nu : DBL_MAX;
}
void HardSphere::applyScattering(std::vector<PicChargedSpecies *> &a_species, Real a_dt_sec) const {
  PicChargedSpecies *a = a_species[m_sp1], *b = a_species[m_sp2];
  if (a->numParticles() == 0 || b->numParticles() == 0) return;
  long nc = 0;
  check(pgpu_collide_hard_sphere_wm(a->handle(), b->handle(), m_sigmaT, m_conservative ? 1 : 0, a_dt_sec, streamSeed(), nextStep(),
                                    &nc),
        "HardSphere::applyScattering");
  m_ncoll = nc;
}
void HardSphere::printParameters() const {
  std::cout << " HardSphere scattering parameters:" << std::endl;
  std::cout << "  species A = " << m_sp1 << ", species B = " << m_sp2 << std::endl;
  std::cout << "  r1 = " << m_r1 << " m, r2 = " << m_r2 << " m, sigmaT = " << m_sigmaT << " m^2" << std::endl;
}

// ---- VariableHardSphere (VariableHardSphere.cpp:28-47, 217-412) ------------------------------------
VariableHardSphere::VariableHardSphere(int a_sp, Real a_eta, Real a_T0, Real a_mu0)
    : m_sp(a_sp), m_eta(a_eta), m_T0(a_T0), m_mu0(a_mu0), m_scatter_dt(DBL_MAX), m_ncoll(0) {}
void VariableHardSphere::setMeanFreeTime(const std::vector<PicChargedSpecies *> &a_species) const {
  double nu = 0.0;
  check(pgpu_scatter_nu_max_vhs(a_species[m_sp]->handle(), m_eta, m_T0, m_mu0, &nu), "VariableHardSphere::setMeanFreeTime");
  m_scatter_dt = nu > 0.0 ? 1.0 / nu : DBL_MAX;
}
void VariableHardSphere::applyScattering(std::vector<PicChargedSpecies *> &a_species, Real a_dt_sec) const {
  PicChargedSpecies *a = a_species[m_sp];
  if (a->numParticles() == 0) return;
  long nc = 0;
  check(pgpu_collide_vhs(a->handle(), m_eta, m_T0, m_mu0, a_dt_sec, streamSeed(), nextStep(), &nc),
        "VariableHardSphere::applyScattering");
  m_ncoll = nc;
}
void VariableHardSphere::printParameters() const {
  std::cout << " VariableHardSphere scattering parameters:" << std::endl;
  std::cout << "  species = " << m_sp << ", eta = " << m_eta << ", T0 = " << m_T0 << " K, mu0 = " << m_mu0 << " Pa s"
            << std::endl;
}

}  // namespace picnic_gpu
