/*
 * picnic_gpu.h -- C ABI of the B200 particle engine behind PICNIC's species and
 * scattering interfaces.
 *
 * The reference has no FFI for this path: its seams are the C++
 * classes PicChargedSpecies / PicSpeciesInterface / Scattering, plus the Fortran
 * array ABI of the per-particle kernels (CHF_FRA1 = pointer + inclusive lo/hi
 * bounds, column-major, ghosts included; src/particle_tools/MeshInterpF_F.H:8-44).
 * Each entry point below names the reference method it replaces; the C++ shim in
 * picnic_b200/host/ forwards the same-named methods to these calls.
 *
 * Conventions
 *   - plain C, opaque handles, no exceptions, no exit(): every call returns 0 on
 *     success and <0 on error (text via pgpu_last_error()).  PGPU_ERR_SEGMENTS is
 *     the reference's Fortran STOP "particle crossing more cell than allowed"
 *     (MeshInterpChargeConservingF.ChF:1018-1021,1555-1558).
 *   - all pointers are HOST pointers borrowed for the duration of the call;
 *     the library owns every device allocation.
 *   - grid arrays: one component, column-major, inclusive lo/hi per direction in
 *     GLOBAL cell/node indices, ghosts included (Chombo FArrayBox layout).
 *   - particle arrays: SoA, component-major (x[d*n+p], v[c*n+p]).
 *   - one caller thread per device; calls that return data synchronise.
 *   - there is NO CPU fallback: if no CUDA device is usable pgpu_init fails.
 */
#ifndef PICNIC_GPU_H
#define PICNIC_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pgpu_grid_s *pgpu_grid_t;
typedef struct pgpu_species_s *pgpu_species_t;

enum { PGPU_CIC = 0, PGPU_TSC = 1, PGPU_CC0 = 2, PGPU_CC1 = 3 }; /* InterpType */
enum { PGPU_EX = 0, PGPU_EY, PGPU_EZ, PGPU_BX, PGPU_BY, PGPU_BZ, PGPU_NFIELD };
enum { PGPU_JX = 0, PGPU_JY, PGPU_JZ };
enum { PGPU_BC_NONE = 0, PGPU_BC_PERIODIC = 1, PGPU_BC_SYMMETRY = 2, PGPU_BC_OUTFLOW = 3, PGPU_BC_INFLOW_OUTFLOW = 4 };
enum {
  PGPU_OK = 0,
  PGPU_ERR_ARG = -1,
  PGPU_ERR_CUDA = -2,
  PGPU_ERR_SEGMENTS = -3, /* CC1: num_segments > ghosts+1 */
  PGPU_ERR_BOUNDS = -4,   /* a particle stencil left the ghosted arrays */
  PGPU_ERR_STATE = -5,
  PGPU_ERR_COMM = -6
};

/* ---- lifecycle ---------------------------------------------------------- */
int pgpu_init(int device);
int pgpu_finalize(void);
const char *pgpu_last_error(void);
int pgpu_abi_version(void);
/* Run all kernels on this cudaStream_t (0 = the library's own stream). */
int pgpu_set_stream(void *cuda_stream);
int pgpu_synchronize(void);
/* 0: fast arithmetic (FMA contraction, reciprocal multiplies, guarded floors);
 * 1: reference operation order (no contraction, true divides).  Cell/node indices
 * are bit-exact in both modes. */
int pgpu_set_exact_math(int on);
/* Deposit/advance algorithm: 1 (default) = 2D CC1 species take the specialised fused
 * kernel (single-segment closed forms, per-warp shared-memory run sums, one fp64 RED
 * per node and run; fastest on a cell-sorted species, correct on any order), with the
 * generic visitor kernel for the particles it defers; 0 = generic kernel only (one
 * global fp64 RED per particle and node).  Exact-math mode always uses the generic one. */
int pgpu_set_deposit_mode(int mode);

/* ---- grid: DomainGrid + the box owned by this device ----------------------- */
typedef struct {
  int D;               /* SpaceDim (1 or 2) */
  int ncell[2];        /* grid.num_cells (global) */
  double xmin[2];      /* DomainGrid::getXmin */
  double dx[2];        /* DomainGrid::getdX */
  int nghost;          /* grid.num_ghosts */
  int periodic[2];     /* grid.is_periodic */
  int box_lo[2];       /* cells owned by this device, inclusive, global indices */
  int box_hi[2];
  double volume_scale; /* DomainGrid::getVolumeScale */
} pgpu_grid_desc;

int pgpu_grid_create(const pgpu_grid_desc *desc, pgpu_grid_t *out);
int pgpu_grid_destroy(pgpu_grid_t g);
/* Upload one E/B component (EMFields E on edges/nodes, B on faces/cells) for the
 * whole ghosted box: what PicChargedSpecies::interpolateFieldsToParticles reads
 * (PicChargedSpecies.cpp:3814-3917).  lo/hi must equal the box's ghosted bounds
 * for that centring. */
int pgpu_fields_set(pgpu_grid_t g, int comp, const double *data, const int *lo, const int *hi);
/* Up to 4 resident field sets per grid: pgpu_fields_set writes the selected slot and
 * the particle kernels read it.  Lets the host shim double-buffer the E/B upload of
 * the next nonlinear iteration (preRHSOp, PicSpeciesInterface.cpp:899-994) behind the
 * particle work of the current one.  Slot 0 is selected at creation. */
int pgpu_fields_select(pgpu_grid_t g, int slot);
/* One transfer per preRHSOp each way (PicSpeciesInterface.cpp:899-994 hands the particles six field arrays and takes three
 * current arrays back): the six components Ex Ey Ez Bx By Bz back to back in one host buffer, each in the layout
 * pgpu_fields_set takes; and Jx Jy Jz of pgpu_current_finalize likewise.  pgpu_current_get_packed_async returns before the
 * copy has finished: read the buffer after pgpu_synchronize. */
int pgpu_fields_packed_size(pgpu_grid_t g, long *ndoubles);
int pgpu_fields_set_packed(pgpu_grid_t g, const double *data);
int pgpu_current_packed_size(pgpu_grid_t g, long *ndoubles);
int pgpu_current_get_packed_async(pgpu_grid_t g, double *data);
/* Page-lock a caller-owned host array (e.g. a Chombo FArrayBox dataPtr) so that the
 * H2D/D2H copies of pgpu_fields_set / pgpu_current_get run at full PCIe/C2C rate. */
int pgpu_host_register(void *ptr, size_t bytes);
int pgpu_host_unregister(void *ptr);
/* Bounds of component `comp` (PGPU_EX..BZ; J uses the E bounds) */
int pgpu_field_bounds(pgpu_grid_t g, int comp, int *lo, int *hi);
/* Total current of all species: PicSpeciesInterface::m_currentDensity[_virtual] */
int pgpu_current_zero(pgpu_grid_t g);                         /* SpaceUtils::zero, PicSpeciesInterface.cpp:831-834 */
int pgpu_current_add_species(pgpu_grid_t g, pgpu_species_t s);/* FArrayBox::plus, :841-855 */
/* Ghost ADD-exchange of the total J, then ghost refresh: the exchange part of
 * PicSpeciesInterface::finalizeSettingJ (:766-772).  Single device: periodic fold. */
int pgpu_current_finalize(pgpu_grid_t g);
int pgpu_current_get(pgpu_grid_t g, int comp, double *data, const int *lo, const int *hi);
/* The same without the wait: the copy is only enqueued (page-lock `data` with pgpu_host_register); the
 * array is valid after pgpu_synchronize().  Lets getCurrentDensity + getCurrentDensity_virtual of one
 * nonlinear evaluation cost one synchronisation instead of three. */
int pgpu_current_get_async(pgpu_grid_t g, int comp, double *data, const int *lo, const int *hi);

/* ---- species: PicChargedSpecies ------------------------------------------- */
typedef struct {
  double mass;         /* m_mass   (units of me) */
  double charge;       /* m_charge (units of |qe|) */
  double fnorm_const;  /* m_fnorm_const, PicChargedSpecies.cpp:1953-1958 */
  double cvac_norm;    /* m_cvac_norm,   PicChargedSpecies.cpp:1960 */
  int interp_N;        /* m_interpRhoToGrid (CIC|TSC) */
  int interp_J;        /* m_interpJToGrid */
  int interp_E;        /* m_interpEToParts */
  double rtol;         /* pic_species.rtol_particles */
  int iter_max;        /* pic_species.iter_max_particles */
  int order_swap;      /* pic_species.part_order_swap */
  int bc_check_lo[2];  /* interp_bc_check -> m_bc_check_lo/hi */
  int bc_check_hi[2];
  int motion;          /* m_motion */
  int forces;          /* m_forces */
  /* the reference's compile-time switch RELATIVISTIC_PARTICLES as a per-species flag: Boris with the
   * time-centred gamma (PicSpeciesUtils.cpp:55-78; higuera_cary = pic_species.N.higuera_cary), positions and
   * the Picard step norm with getImplicitGamma (PicChargedSpecies.cpp:496-498,548-553,693-708), current deposit
   * with w/gamma (MeshInterpI.H:72-91), setStableDt and globalMoments (:1890-1893, 4095-4098).  Such species
   * take the generic kernels; pgpu_collide_ta and pgpu_collide_coulomb switch to their LorentzScatter
   * (TakizukaAbe.cpp:580-659, Coulomb.cpp:1694-1793), Elastic stays Galilean. */
  int relativistic;
  int higuera_cary;
} pgpu_species_desc;

int pgpu_species_create(pgpu_grid_t g, const pgpu_species_desc *desc, pgpu_species_t *out);
int pgpu_species_destroy(pgpu_species_t s);
/* PicChargedSpecies::setParticleSolverParams (PicChargedSpecies.H:125-141) */
int pgpu_species_set_solver_params(pgpu_species_t s, int order_swap, int iter_max, double rtol);
/* Replace the particle set (PicChargedSpecies::initialize / partData() sync). */
int pgpu_species_upload(pgpu_species_t s, long n, const double *x, const double *xold,
                        const double *v, const double *vold, const double *w,
                        const uint64_t *id);
/* Any output pointer may be NULL. */
int pgpu_species_download(pgpu_species_t s, double *x, double *xold, double *v,
                          double *vold, double *w, uint64_t *id);
long pgpu_species_count(pgpu_species_t s);            /* numParticles() */
/* The same through the particle object's own linear record, for hosts that keep a List<JustinsParticle> (partData()
 * users: I/O, Piston, out-of-scope scattering models): JustinsParticle::linearOut / linearIn
 * (src/particle_tools/JustinsParticle.cpp:339-378, 410-450), per particle
 *   [ w | x[D] | x_old[D] | pos_virt[2] | v[3] | v_old[3] | (Real) ID ],  pgpu_particle_linear_size() = (2 D + 10) * 8 bytes
 * (JustinsParticle::size()); pos_virt is written as zero and ignored on input (planar push).  `records` holds
 * pgpu_species_count() resp. n records back to back, i.e. what ListBox::linearOut / linearIn move for a box. */
long pgpu_particle_linear_size(pgpu_species_t s);
int pgpu_species_download_linear(pgpu_species_t s, void *records);
int pgpu_species_upload_linear(pgpu_species_t s, long n, const void *records);

/* push: same-named PicChargedSpecies methods */
int pgpu_advance_positions_explicit(pgpu_species_t s, double full_dt, int half_step); /* :463-504 */
int pgpu_advance_positions_implicit(pgpu_species_t s, double full_dt);               /* :506-561 */
int pgpu_advance_positions_2nd_half(pgpu_species_t s);                               /* :997-1025 */
/* External fields: EMFields::getExternalE/B (src/fields/EMFields.H:176-198) evaluate six GridFunction objects at the
 * particle position (x_bar during the implicit solve); PicChargedSpecies::addExternalFieldsToParticles
 * (PicChargedSpecies.cpp:3948-3996) adds them to E_p, B_p after every gather of the particle loop (:1606, :1652, :1669).
 * Supported here: Constant (ibc/grid_functions/Constant.H), Cosine (Cosine.H:31-45), Heavyside (Heavyside.H:40-54). */
enum { PGPU_EXT_NONE = 0, PGPU_EXT_CONSTANT = 1, PGPU_EXT_COSINE = 2, PGPU_EXT_HEAVYSIDE = 3 };
typedef struct {
  int type;                        /* PGPU_EXT_* */
  double value;                    /* Constant: value; Cosine: amplitude */
  double constant;                 /* Cosine: constant */
  double L[2], mode[2], phase[2];  /* Cosine: value = amplitude * prod_d cos(fmod(2 pi mode_d x_d / L_d + phase_d pi, 2 pi)) + constant */
  double C[2], A[2], X0[2], eps[2];/* Heavyside: prod_d (C_d + A_d H(x_d - X0_d)), H = 1/2 within eps_d */
} pgpu_ext_fn;
/* six functions in the order Ex Ey Ez Bx By Bz; NULL switches the external fields off (the default) */
int pgpu_grid_set_external_fields(pgpu_grid_t g, const pgpu_ext_fn *six_or_null);
/* addExternalFieldsToParticles on the stored E_p, B_p of pgpu_interpolate_fields_to_particles; the fused advance
 * entry points add them by themselves */
int pgpu_add_external_fields_to_particles(pgpu_species_t s);
int pgpu_interpolate_fields_to_particles(pgpu_species_t s);                          /* :3814-3917 */
int pgpu_advance_velocities(pgpu_species_t s, double full_dt, int half_step);        /* :1116-1134 */
int pgpu_advance_velocities_2nd_half(pgpu_species_t s);                              /* :1136-1244 */
int pgpu_average_velocities(pgpu_species_t s);                                       /* :1202-1226 */
int pgpu_update_old_particle_positions(pgpu_species_t s);                            /* :1821-1845 */
int pgpu_update_old_particle_velocities(pgpu_species_t s);                           /* :1847-1867 */
int pgpu_reset_particles(pgpu_species_t s);                                          /* :1791-1819 */
/* Ep/Bp of the last pgpu_interpolate_fields_to_particles (test/diagnostic hook) */
int pgpu_species_download_fields(pgpu_species_t s, double *Ep, double *Bp);
/* Set Ep/Bp directly: JustinsParticle::setElectricField / setMagneticField
 * (JustinsParticle.H:127-151); lets pgpu_advance_velocities be driven -- and checked against
 * the reference's PicSpeciesUtils::applyForces -- without a gather. */
int pgpu_species_upload_fields(pgpu_species_t s, const double *Ep, const double *Bp);

typedef struct {
  long num_parts_its;   /* m_num_parts_its increment  (:1646) */
  long num_apply_its;   /* m_num_apply_its increment  (:1654,1671) */
  long num_unconverged; /* particles left at the iteration cap (:1680) */
} pgpu_picard_stats;
/* advanceParticles (:1594-1612) */
int pgpu_advance_particles(pgpu_species_t s, double dt);
/* advanceParticlesIteratively (:1614-1716).  deposit_J != 0 fuses this species'
 * setCurrentDensity(dt,false) (:3184-3253) into the same kernel. */
int pgpu_advance_particles_iteratively(pgpu_species_t s, double dt, int deposit_J,
                                       pgpu_picard_stats *stats);

/* The curvilinear velocity pushes PicChargedSpecies::applyForces dispatches to (PicChargedSpecies.cpp:341-355):
 * PicSpeciesUtils::applyForces_CYL_CYL / _SPH_SPH / _CYL_HYB / _SPH_HYB (src/species/pic/PicSpeciesUtils.cpp:103-473), on
 * the stored particle fields (pgpu_interpolate_fields_to_particles) like pgpu_advance_velocities.  r_old is x_old[0];
 * the particle's position_virt (dtheta, dphi) lives in two device arrays ([2][n] on the host side): zero selects the
 * predictor-corrector branch of CYL_CYL / SPH_SPH, which stores the corrected angle.  anticyclic: components stored
 * {X, Z, Y} (cyl_RZ).  The HYB types return the time-centred velocity (the reference has no byHalfDt there).  Only the
 * velocity update of the curvilinear geometries is covered: their position advance, virtual-position re-basing and
 * Jacobian-weighted deposits are not. */
enum { PGPU_PUSH_CYL_CYL = 1, PGPU_PUSH_SPH_SPH = 2, PGPU_PUSH_CYL_HYB = 3, PGPU_PUSH_SPH_HYB = 4 };
int pgpu_apply_forces_curvilinear(pgpu_species_t s, int push_type, double full_dt, int by_half_dt, int anticyclic);
int pgpu_species_virtual_positions_set(pgpu_species_t s, const double *virt /* [2][n] */);
int pgpu_species_virtual_positions_get(pgpu_species_t s, double *virt /* [2][n] */);

/* PIC_EM_EXPLICIT: the particle side of one leap-frog step (PICTimeIntegrator_EM_Explicit::timeStep,
 * src/time/PICTimeIntegrator_EM_Explicit.cpp:92-170, default branch) as ONE pass over the particles:
 * interpolateFieldsToParticles + addExternalFieldsToParticles + advanceVelocities(dt, false) (:94-111),
 * advancePositionsExplicit(dt/2) + applyBCs (:126-128), setCurrentDensity(dt, true) (:137) into the species current and,
 * if second_half != 0, advancePositions_2ndHalf + applyBCs (:166-168; only legal to fuse when no scattering runs between
 * the deposit and the second half).  Expects x_old == x and u_old == u (updateOldParticle*), like the separate calls. */
int pgpu_explicit_step(pgpu_species_t s, double dt, const int *bc_lo, const int *bc_hi, int second_half);

/* Sub-orbit model (pic_species.N.use_suborbit_model, .suborbit_fast_particles): a second particle container per species
 * (m_data_suborbit) for the particles the particle Picard loop leaves unconverged at iter_max_particles
 * (PicChargedSpecies.cpp:1699-1706: moved there by pgpu_advance_particles_iteratively, with two sub-orbits, depositing
 * nothing in that call) and for "fast" particles (transferFastParticles, :894-956).
 * pgpu_advance_suborbit_particles_and_set_J = advanceSubOrbitParticlesAndSetJ (:3324-3669), bulk container, PLANAR push:
 * every sub-orbit particle takes the step in nsub equal implicit sub-steps (one more whenever a sub-step does not
 * converge), ends at the NEW-time x, u with x_old, u_old of the step start, and the species' sub-orbit current
 * (m_suborbitJ) = sum over particles of (sum over sub-orbits of the deposit) / nsub, x charge / volume_scale.
 * pgpu_current_add_suborbit = PicSpeciesInterface::addSubOrbitJ (PicSpeciesInterface.cpp:1538-1590);
 * pgpu_merge_suborbit_particles = mergeSubOrbitParticles (:1718-1747), to be called after the second half-step of the
 * main container, as PICTimeIntegrator_EM_ThetaImplicit.cpp:313-318 does. */
int pgpu_species_set_suborbit_model(pgpu_species_t s, int use_suborbit_model, int suborbit_fast_particles);
long pgpu_species_suborbit_count(pgpu_species_t s);
int pgpu_transfer_fast_particles(pgpu_species_t s);
int pgpu_advance_suborbit_particles_and_set_J(pgpu_species_t s, double dt, int from_emjacobian);
int pgpu_species_suborbit_current_get(pgpu_species_t s, int comp, double *data, const int *lo, const int *hi);
int pgpu_current_add_suborbit(pgpu_grid_t g, pgpu_species_t s);
int pgpu_merge_suborbit_particles(pgpu_species_t s);
/* test / I-O hook: the sub-orbit container (component-major arrays of pgpu_species_suborbit_count entries) */
int pgpu_species_suborbit_download(pgpu_species_t s, double *x, double *xold, double *v, double *vold, double *w,
                                   uint64_t *id, int *nsub);

/* Outflow and inflow boundaries (PicChargedSpeciesBC, species BC types "outflow" and "inflow_outflow"):
 * pgpu_apply_bcs with PGPU_BC_OUTFLOW / PGPU_BC_INFLOW_OUTFLOW moves the particles beyond that boundary (x < Xmin,
 * x >= Xmax; outflow_Lo/Hi, PicChargedSpeciesBC.cpp:872-918) from the species into its outflow lists, kept on the device
 * and tagged boundary = 2 dir + side.  pgpu_set_current_density(.., from_explicit_solver = 1) adds their current
 * (depositInflowOutflowJ, :667-736).  The host reads them for the surface charge it keeps (pgpu_species_outflow_download),
 * takes the flux diagnostics m_delta_{Mass,MomX,MomY,MomZ,Energy}Out per boundary (flux[boundary * 5 + k]) and empties the
 * lists (removeOutflowParticles, :508-545).  Inflow particles are made by the host's InflowBC objects
 * (createInflowParticles, :467-506) and enter through pgpu_species_append (injectInflowParticles, :563-665; x_old / u_old
 * NULL = same as x / u; id NULL = made up). */
long pgpu_species_outflow_count(pgpu_species_t s);
int pgpu_species_outflow_download(pgpu_species_t s, double *x, double *xold, double *v, double *vold, double *w,
                                  uint64_t *id, int *boundary);
int pgpu_species_outflow_fluxes(pgpu_species_t s, double *flux20);
int pgpu_remove_outflow_particles(pgpu_species_t s);
int pgpu_species_append(pgpu_species_t s, long n, const double *x, const double *xold, const double *v, const double *vold,
                        const double *w, const uint64_t *id);

/* Inflow lists with pic_species.N.suborbit_inflow_J = true (the implicit shock decks): the particles the host's
 * createInflowParticles made for boundary (bdry_dir, bdry_side) -- positions OUTSIDE the domain, moving in -- are not
 * injected at once (PicChargedSpecies::injectInflowParticles returns, PicChargedSpecies.cpp:1787) but wait in the
 * boundary's inflow list (PicChargedSpeciesBC::m_inflow_list_vector): pgpu_species_inflow_append (x_old = x, u_old = u,
 * one sub-orbit).  Every nonlinear evaluation then calls pgpu_advance_inflow_particles_and_set_J =
 * PicChargedSpecies::advanceInflowParticlesAndSetJ (:3255-3322): free streaming to the boundary plane
 * (advanceInflowPartToBdry, :958-995), the rest of the step as sub-orbits (advanceSubOrbitParticlesAndSetJ with
 * is_inflow_list, :3376-3669), current into the species' inflow J (x charge / volume_scale), which
 * pgpu_current_add_inflow adds to the total (PicSpeciesInterface::addInflowJ, PicSpeciesInterface.cpp:1499-1536).  The
 * particles stay in the list, time-centred against their original old state; pgpu_apply_bcs with
 * PGPU_BC_INFLOW_OUTFLOW on that side then does PicChargedSpeciesBC::inflow_Lo / inflow_Hi
 * (PicChargedSpeciesBC.cpp:961-1001, 1047-1086): 2 x - x_old inside the boundary joins the species, anything else (a
 * particle the fields turned around) is dropped, and the m_delta_*In probes accumulate (pgpu_species_inflow_fluxes:
 * [2 dir + side][w, w ux, w uy, w uz, w |u|^2 / (gamma + 1)] of u_old; reading resets). */
int pgpu_species_inflow_append(pgpu_species_t s, long n, const double *x, const double *v, const double *w,
                               const uint64_t *id, int bdry_dir, int bdry_side);
long pgpu_species_inflow_count(pgpu_species_t s);
/* nsub_boundary[i] = 8 * numSubOrbits + (2 dir + side) */
int pgpu_species_inflow_download(pgpu_species_t s, double *x, double *xold, double *v, double *vold, double *w,
                                 uint64_t *id, int *nsub_boundary);
int pgpu_species_inflow_clear(pgpu_species_t s);
int pgpu_advance_inflow_particles_and_set_J(pgpu_species_t s, double dt, int from_emjacobian);
int pgpu_species_inflow_current_get(pgpu_species_t s, int comp, double *data, const int *lo, const int *hi);
int pgpu_current_add_inflow(pgpu_grid_t g, pgpu_species_t s);
int pgpu_species_inflow_fluxes(pgpu_species_t s, double *flux20);

/* deposit */
int pgpu_set_current_density(pgpu_species_t s, double dt, int from_explicit_solver); /* :3184-3253 */
int pgpu_species_current_get(pgpu_species_t s, int comp, double *data, const int *lo, const int *hi);
/* setChargeDensityOnNodes (:3134-3182, stag = 1,1), setChargeDensity (:3049-3086,
 * stag = 0,0) and one face direction (:3088-3132): deposit, x charge/volume_scale,
 * ghost add-exchange.  Result returned in data (bounds for that centring). */
int pgpu_set_charge_density(pgpu_species_t s, const int *stag, double *data, const int *lo, const int *hi);
/* The two halves of pgpu_set_charge_density for a domain of several boxes: _deposit leaves the species' scaled charge
 * density in the grid's resident array of that centring, self-periodic directions folded; a pgpu_halo_create_rho
 * plan adds the ghost layers of neighbouring boxes into each other on the device; pgpu_charge_density_filter
 * optionally smooths it; _get copies the array out (bounds = the ghosted box of the centring). */
/* SpaceUtils::applyBinomialFilter (SpaceUtils.cpp:9-112) on the device arrays: the [1 2 1]/4 (1D) or
 * [1 2 1; 2 4 2; 1 2 1]/16 (2D) smoothing over the box's own edges / nodes, same operation order as the reference.
 * pgpu_current_filter is the filter step of PicSpeciesInterface::filterJ (PicSpeciesInterface.cpp:996-1012;
 * in_plane = filterE_inPlane, virtual_comps = filterE_virtual) and runs after pgpu_current_finalize / the halo
 * add-exchange, which leave the neighbour sums in the first ghost layer; pgpu_charge_density_filter is the
 * a_use_filtering tail of setChargeDensityOnNodes (PicChargedSpecies.cpp:3176-3180), between the exchange and
 * pgpu_charge_density_get.  Ghosts at a physical (non-periodic) boundary are used as they are in the device array:
 * the field-BC fill the reference does there (applyEdgeBC / applyNodeBC / applyToRhoInGhosts) is host-side grid work. */
int pgpu_current_filter(pgpu_grid_t g, int in_plane, int virtual_comps);
int pgpu_charge_density_filter(pgpu_grid_t g, const int *stag);
int pgpu_charge_density_deposit(pgpu_species_t s, const int *stag);
int pgpu_charge_density_get(pgpu_grid_t g, const int *stag, double *data, const int *lo, const int *hi);

/* cell sort + cell moments: binTheParticles (:1913-1947), set*DensityFromBinFab
 * (:2881-3047), PicSpeciesInterface::setDebyeLength (PicSpeciesInterface.cpp:1627-1721) */
int pgpu_bin_particles(pgpu_species_t s);
/* The engine's own locality sort for collisionless steps (no reference counterpart: PICNIC bins only for
 * scattering): orders the particles by the cell of the half-shifted grid that the CC1 deposit of
 * MeshInterpChargeConservingF.ChF:1517-1625 segments on, so that all particles that touch the same 21 nodes
 * are contiguous.  Any particle order is correct; this one is the fast one for the fused advance + deposit
 * and for the mass-matrix deposit.  Invalidates the per-cell lists of pgpu_bin_particles. */
int pgpu_sort_for_locality(pgpu_species_t s);
int pgpu_species_cell_index(pgpu_species_t s, int *cell /* [D][n] */);
int pgpu_species_cell_offsets(pgpu_species_t s, long *offsets /* ncell+1 */);
int pgpu_set_moments_from_bins(pgpu_species_t s);
int pgpu_species_moments_get(pgpu_species_t s, double *dens, double *mom, double *ene);
int pgpu_debye_length(pgpu_grid_t g, pgpu_species_t *species, int nspecies, double *LDe_or_null);

/* boundary conditions + ownership: applyBCs (:1080-1114) for periodic / symmetry
 * (PicChargedSpeciesBC.cpp:738-765, 808-870).  bc_lo/bc_hi[d] = PGPU_BC_*. */
int pgpu_apply_bcs(pgpu_species_t s, const int *bc_lo, const int *bc_hi);

/* advanceVelocities_2ndHalf + advancePositions_2ndHalf + applyBCs in one pass over the
 * particles: the tail of every implicit step (PICTimeIntegrator_EM_ThetaImplicit.cpp:311-319).
 * Bit-identical to the three separate calls; falls back to them for symmetry walls. */
int pgpu_finish_implicit_step(pgpu_species_t s, const int *bc_lo, const int *bc_hi);

/* ---- multi-box exchanges (one box per device) ----------------------------
 * The collectives are issued by the host plumbing on DEVICE buffers (suffix _d = device
 * pointer); these calls move data between those buffers and the library's arrays. */
enum { PGPU_FAB_JTOTAL = 0, PGPU_FAB_FIELD = 1 };
/* Copy the index box lo..hi (global indices, inclusive) of a grid array into / out of a
 * contiguous column-major device buffer; add != 0 accumulates: the two halves of the ghost
 * ADD-exchange of PicSpeciesInterface::finalizeSettingJ (PicSpeciesInterface.cpp:766-772). */
int pgpu_fab_pack_d(pgpu_grid_t g, int kind, int comp, const int *lo, const int *hi, double *buf_d);
int pgpu_fab_unpack_d(pgpu_grid_t g, int kind, int comp, const int *lo, const int *hi, const double *buf_d,
                      int add);
/* Particle migration = ParticleData::gatherOutcast + remapOutcast (ParticleDataI.H:405-547).
 * mark: counts[(d0+1)+3*(d1+1)] = particles now owned by the neighbour box in direction
 * (d0,d1) (counts[4] = 0), counts[9] = particles outside the decomposition.  pack: writes the
 * leavers' wire records (pgpu_wire_doubles() doubles each: x[D] xold[D] v[3] vold[3] w id),
 * grouped by direction code in ascending order, to buf_d and removes them from the species.
 * append: adds n_add records received from neighbours. */
int pgpu_wire_doubles(pgpu_grid_t g);
int pgpu_species_mark_leavers(pgpu_species_t s, long *counts /* [10] */);
/* The same without a host synchronisation: the counts go to a device int64[10] (so that they
 * can be all-gathered on the device and read back once for all boxes and species); the caller
 * hands this box's counts back with _set_leaver_counts before packing. */
int pgpu_species_mark_leavers_d(pgpu_species_t s, long long *counts_d);
int pgpu_species_set_leaver_counts(pgpu_species_t s, const long *counts /* [10] */);
int pgpu_species_pack_leavers_d(pgpu_species_t s, double *buf_d);
int pgpu_species_append_d(pgpu_species_t s, long n_add, const double *buf_d);

/* Ghost ADD-exchange of the total J over peer memory (NVLink/NVSwitch), no collective library on the
 * data path: LevelData::exchange with an add op in finalizeSettingJ (PicSpeciesInterface.cpp:766-772).
 * A plan is a list of messages; message = the index boxes (global indices, inclusive) of Jx, Jy, Jz that
 * this box shares with one neighbour in one phase (direction).  The same boxes are sent (this box's
 * copy goes into the neighbour's inbox) and received (the neighbour's copy, delivered into inbox area
 * recv_area of this box, is added).  Inboxes are device buffers of the owning process: same-process
 * peers pass pgpu_halo_inbox() pointers around, other processes the 64-byte CUDA IPC handle.
 * One exchange = pgpu_halo_begin, then for every phase pgpu_halo_send followed by pgpu_halo_recv_add
 * (all boxes of a process send before any receives).  Collective; nothing synchronises with the host. */
typedef struct pgpu_halo_s *pgpu_halo_t;
typedef struct {
  int phase;         /* 0 .. 7, executed in ascending order */
  int recv_area;     /* 0 .. nmsg-1, unique: where the neighbour's copy of these boxes lands */
  int lo[3][2], hi[3][2];
} pgpu_halo_msg;
int pgpu_halo_create(pgpu_grid_t g, int nmsg, const pgpu_halo_msg *msgs, pgpu_halo_t *out);
/* The same plan for one resident charge-density array (the centring `stag`, see pgpu_charge_density_deposit): only
 * lo[0]/hi[0] of each message are read.  It replaces the rho.exchange / addGhosts pass at the end of
 * PicChargedSpecies::setChargeDensity, setChargeDensityOnFaces and setChargeDensityOnNodes
 * (PicChargedSpecies.cpp:3083, :3120 and :3141-3143, the reverseCopier add-exchanges). */
int pgpu_halo_create_rho(pgpu_grid_t g, const int *stag, int nmsg, const pgpu_halo_msg *msgs, pgpu_halo_t *out);
int pgpu_halo_destroy(pgpu_halo_t h);
int pgpu_halo_phases(pgpu_halo_t h);
/* inbox area of message msg: offset in doubles behind the flag block, and its length */
int pgpu_halo_area_offset(pgpu_halo_t h, int msg, long *offset_doubles, long *count);
int pgpu_halo_inbox(pgpu_halo_t h, void **inbox_d, size_t *bytes);
int pgpu_halo_ipc_handle(pgpu_halo_t h, void *handle64);
int pgpu_halo_ipc_open(pgpu_halo_t h, const void *handle64, void **inbox_d);
/* message msg is delivered into area peer_area (at peer_area_offset doubles) of the inbox peer_inbox_d */
int pgpu_halo_connect(pgpu_halo_t h, int msg, void *peer_inbox_d, int peer_area, long peer_area_offset);
int pgpu_halo_begin(pgpu_halo_t h);
int pgpu_halo_send(pgpu_halo_t h, int phase);
int pgpu_halo_recv_add(pgpu_halo_t h, int phase);

/* Particle migration over peer memory: gatherOutcast + remapOutcast (ParticleDataI.H:405-547) without a
 * message layer and with device-side counts.  Every species of every box owns an inbox of 9 areas (one per
 * direction the particles arrive from) of capacity_records wire records, double buffered by step parity.
 * send: marks the leavers, stores their records straight into the owning neighbours' inboxes (NVLink),
 * fills the holes from the tail, posts counts + arrival flags.  recv: waits for the flags of all connected
 * neighbours and appends what arrived.  finish: the step's only host synchronisation; updates the particle
 * count.  With several species: send all, recv all, finish all (one wait).  More than capacity_records
 * leavers towards one neighbour in one step is an error (PGPU_ERR_STATE at finish). */
typedef struct pgpu_migrator_s *pgpu_migrator_t;
int pgpu_migrator_create(pgpu_species_t s, long capacity_records, pgpu_migrator_t *out);
int pgpu_migrator_destroy(pgpu_migrator_t m);
int pgpu_migrator_inbox(pgpu_migrator_t m, void **inbox_d, size_t *bytes);
int pgpu_migrator_ipc_handle(pgpu_migrator_t m, void *handle64);
int pgpu_migrator_ipc_open(pgpu_migrator_t m, const void *handle64, void **inbox_d);
/* the box in direction code = (d0+1) + 3 (d1+1) of this one owns inbox peer_inbox_d */
int pgpu_migrator_connect(pgpu_migrator_t m, int code, void *peer_inbox_d);
int pgpu_migrate_send(pgpu_migrator_t m);
int pgpu_migrate_recv(pgpu_migrator_t m);
int pgpu_migrate_finish(pgpu_migrator_t m, long *n_arrived, long *n_left, long *n_lost);

/* reductions: setStableDt (:1869-1911), globalMoments (:4067-4130) */
int pgpu_stable_dt(pgpu_species_t s, double *dt_out);
int pgpu_global_moments(pgpu_species_t s, double *out /* [w, wux,wuy,wuz, wuu_x,wuu_y,wuu_z] */);

/* ---- mass matrices (use_mass_matrices decks; CC1, planar push) -------------------------------
 * The nine sigma containers and J0 / E0 of PicSpeciesInterface live on the device, on the boxes of the
 * J component of their row (sigma_x*: Jx box, sigma_y*: Jy box, sigma_z*: Jz box), CHF_FRA layout
 * (component index slowest).  Order everywhere: xx xy xz yx yy yz zx zy zz. */
/* PicSpeciesInterface::initializeMassMatrices (PicSpeciesInterface.cpp:225-417): allocates; ncomp_or_null
 * receives the components per direction [9][2] (m_ncomp_xx ... m_ncomp_zz). */
int pgpu_mass_matrices_init(pgpu_grid_t g, int interp, int *ncomp_or_null);
int pgpu_mass_matrices_ncomp(pgpu_grid_t g, int *ncomp /* [9][2] */);
/* PicSpeciesInterface::setMassMatrices (:1038-1095) = zero, accumulate every species, save_E0. */
int pgpu_mass_matrices_zero(pgpu_grid_t g);
/* PicChargedSpecies::accumulateMassMatrices (PicChargedSpecies.cpp:3671-3761) ->
 * cc1_{1,2}d_deposit_mass_matrix + compute_mm_kernals (MeshInterpMassMatrixF.ChF:835-2076); B = the
 * selected field slot.  Asynchronous; a crossing / bounds error surfaces at the next *_get or
 * pgpu_picard_totals. */
int pgpu_accumulate_mass_matrices(pgpu_species_t s, double dt);
/* m_E0 <- E of the selected field slot (:1076-1091) */
int pgpu_mass_matrices_save_E0(pgpu_grid_t g);
/* PicSpeciesInterface::computeJfromMassMatrices (:567-753) -> compute_J{x,y,z}_from_mass_matrix
 * (src/fields/FieldsF.ChF:3-415): total J = J0 + sigma (E - E0) with E = the selected field slot, written
 * over the whole ghosted box of the grid's total current (pgpu_current_finalize / pgpu_current_get follow,
 * like finalizeSettingJ after it in the reference). */
int pgpu_compute_J_from_mass_matrices(pgpu_grid_t g);
int pgpu_mass_matrix_get(pgpu_grid_t g, int which, double *data, const int *lo, const int *hi, int ncomp);
int pgpu_mass_matrix_J0_get(pgpu_grid_t g, int comp, double *data, const int *lo, const int *hi);

/* ---- collisions: Scattering subclasses ------------------------------------- */
/* TakizukaAbe::applyScattering (TakizukaAbe.cpp:240-536).  Species must be binned
 * and have number densities set.  sA == sB selects self-scattering.  The Philox
 * counter is keyed by (seed, step, cell, pair). */
int pgpu_collide_ta(pgpu_species_t sA, pgpu_species_t sB, double Clog, double dt_sec,
                    uint64_t seed, uint64_t step, long *npairs);
/* TakizukaAbe::computeDeltaU (:538-578) for explicit random numbers (test hook). */
int pgpu_ta_delta_u(long n, const double *vp1, const double *den1, const double *vp2,
                    const double *den2, double b90_fact, double Clog, double dt_sec,
                    const double *gauss, const double *u_theta, const double *u_phi,
                    double *dU);

/* TakizukaAbe::LorentzScatter (:580-659) for explicit random numbers (test hook): the collision that
 * pgpu_collide_ta applies when a species is relativistic (pgpu_species_desc.relativistic), through the
 * centre-of-momentum frame.  up1/up2/out1/out2 are [3n] component-major; gauss is used where s12 < 2. */
int pgpu_ta_lorentz_scatter(long n, const double *up1, const double *up2, double mass1, double mass2,
                            const double *den2, double dt_sec, double b90_fact, double Clog,
                            const double *gauss, const double *u_theta, const double *u_phi,
                            double *out1, double *out2);

/* Coulomb::applyScattering, PROBABILISTIC weight method (Coulomb.cpp:358-592, 919-1180): weighted
 * particles, lighter-weight particle always scatters, heavier with probability wmin/wmax; pairing
 * O(N) or all pairs (NxN / cells below NxN_Nthresh); sigma limited by the atomic spacing; b_max =
 * the Debye length set by pgpu_debye_length.  Species must be binned with moments set. */
enum { PGPU_ANG_TAKIZUKA = 0, PGPU_ANG_NANBU = 1, PGPU_ANG_BOBYLEV = 2, PGPU_ANG_NANBU_FAS = 3, PGPU_ANG_NANBU_FAS_V2 = 4,
       PGPU_ANG_ISOTROPIC = 5 };   /* the reference's enum order (Coulomb.H:132-139) */
typedef struct {
  double Clog;            /* coulomb_logarithm; 0 = per pair from b_max / b_min (Coulomb.cpp:1664-1672) */
  int angular_scattering; /* PGPU_ANG_*; the host applies exclude_electron_fas (Coulomb.cpp:65-68: NANBU_FAS(_v2) -> NANBU
                           * when one of the species is the electron) */
  int NxN;                /* Coulomb.NxN */
  int NxN_Nthresh;        /* Coulomb.NxN_Nthresh (11) */
  int num_subcycles;      /* Coulomb.num_subcycles (1) */
  /* scattering.coulomb.enforce_conservations and companions (Coulomb.H:286-293; Coulomb.cpp:596-714, 1182-1430): after
   * the weight-rejection update the cell's momentum change is taken back out of every particle and the energy change is
   * absorbed by ScatteringUtils::modEnergyPairwise sweeps.  0 = off (the remaining fields are then ignored). */
  int enforce_conservations;
  double energy_fraction;       /* 0.05 */
  double energy_fraction_max;   /* 0.5 */
  int beta_weight_exponent;     /* 1 */
  int sort_weighted_particles;  /* must be 0: the sweeps pair particles in storage order */
  int conservation_Nmin_save;   /* 100000 (0 = that default) */
  /* weight_method: 0 = PROBABILISTIC (above), 1 = CONSERVATIVE = Coulomb::applyIntra/InterScattering_SK08
   * (Coulomb.cpp:730-917, 1439-1640; Sentoku & Kemp 2008): O(N) pairs; the lighter-weight particle scatters, the heavier
   * one takes the fraction w_min / w_max of its scattered change plus a transverse kick that conserves the pair's
   * weighted energy exactly (Coulomb::enforceEnergyConservation, Coulomb.H:796-823).  Galilean build only. */
  int weight_method;
  /* scattering.coulomb.include_large_angle_scattering (Coulomb::SetPolarScattering, Coulomb.cpp:1801-1863): single
   * Rutherford events below a cutoff impact parameter on top of the cumulative small-angle model.  The host applies
   * exclude_electron_fas (Coulomb.cpp:69-71: off when one of the species is the electron).  test_large_angle_draw is
   * the event's uniform draw in the explicit-draw test entry points (pgpu_coulomb_delta_u, _lorentz_scatter); the
   * collision kernels draw it from Philox. */
  int include_large_angle_scattering;
  double test_large_angle_draw;
  /* NANBU_FAS / NANBU_FAS_v2 (Coulomb.H:365-718, full-angle scattering after Higginson JCP 2017) draw up to three uniforms
   * per pair, one after the other: u_polar and these two in the explicit-draw test entry points; the collision kernels
   * take them from Philox. */
  double test_fas_draw2, test_fas_draw3;
} pgpu_coulomb_params;
int pgpu_collide_coulomb(pgpu_species_t sA, pgpu_species_t sB, const pgpu_coulomb_params *prm, double dt_sec,
                         uint64_t seed, uint64_t step, long *npairs);
/* Coulomb::GalileanScatter + SetPolarScattering for explicit random numbers (test hook); per pair:
 * EF_norm, den12, bmax, sigma_max, gauss (N(0,1)), u_polar, u_phi (U[0,1)); outputs dU[3n], s12[n]. */
int pgpu_coulomb_delta_u(long n, const double *vp1, const double *vp2, double charge1, double charge2,
                         double mass1, double mass2, const pgpu_coulomb_params *prm, double dt_sec,
                         const double *EF_norm, const double *den12, const double *bmax, const double *sigma_max,
                         const double *gauss, const double *u_polar, const double *u_phi, double *dU, double *s12);
/* HardSphere::applySelfScattering / applyInterScattering, PROBABILISTIC weight method (HardSphere.cpp:223-665):
 * no-time-counter pair selection with gmax = 5 thermal speeds of the cell, isotropic scattering.  sigmaT =
 * pi (r1 + r2)^2 (HardSphere.cpp:52).  Species must be binned with their cell moments set. */
int pgpu_collide_hard_sphere(pgpu_species_t sA, pgpu_species_t sB, double sigmaT, double dt_sec, uint64_t seed,
                             uint64_t step, long *ncollisions);
/* The same with the deck's weight_method: 0 = PROBABILISTIC, 1 = CONSERVATIVE (HardSphere.cpp:357-392: for unequal
 * weights the heavier particle, its scattered fraction and a third particle of the cell are merged into two equally
 * weighted ones by ScatteringUtils::collapseThreeToTwo; the weights change).  Between two species (:594-636) the third
 * particle comes from the heavier particle's species, and -- as in the reference -- the partners move by 0.5 deltaU, not
 * mu/m deltaU: the pair's momentum is conserved only for equal masses. */
int pgpu_collide_hard_sphere_wm(pgpu_species_t sA, pgpu_species_t sB, double sigmaT, int weight_method, double dt_sec,
                                uint64_t seed, uint64_t step, long *ncollisions);
/* VariableHardSphere::applySelfScattering (VariableHardSphere.cpp:217-412; the reference has no inter-species VHS):
 * sigmaT(g) = 4 pi A g^(-4/alpha) with alpha = 4/(2 eta - 1) and A from the viscosity mu0 [Pa s] at T0 [K]
 * (:28-47); both partners of an accepted pair scatter. */
int pgpu_collide_vhs(pgpu_species_t s, double eta, double T0, double mu0, double dt_sec, uint64_t seed, uint64_t step,
                     long *ncollisions);
/* VariableHardSphere::setMeanFreeTime (VariableHardSphere.cpp:50-125): box maximum of n sigmaT(VTeff) VTeff */
int pgpu_scatter_nu_max_vhs(pgpu_species_t s, double eta, double T0, double mu0, double *nu_max);
/* HardSphere::setMeanFreeTime (HardSphere.cpp:65-194): box maximum of n sigmaT sqrt(Teff/m) */
int pgpu_scatter_nu_max_hard_sphere(pgpu_species_t sA, pgpu_species_t sB, double sigmaT, double *nu_max);
/* Coulomb::LorentzScatter (Coulomb.cpp:1694-1793) for n pairs with explicit draws (test hook of the relativistic
 * pair update that pgpu_collide_coulomb applies when a species is relativistic): particle 1 always scatters,
 * particle 2 where scatter2[i] != 0. */
int pgpu_coulomb_lorentz_scatter(long n, const double *up1, const double *up2, const int *scatter2, double charge1,
                                 double charge2, double mass1, double mass2, const pgpu_coulomb_params *prm,
                                 double dt_sec, const double *EF_norm, const double *den12, const double *bmax,
                                 const double *sigma_max, const double *gauss, const double *u_polar,
                                 const double *u_phi, double *out1, double *out2, double *s12);

/* Elastic::electronImpact (Elastic.cpp:225-388): every particle of sA picks a random partner of sB in its cell; sigma
 * constant (ntab = 0) or tabulated (E [eV] ascending, Q, xi) with the reference's interpolation; angular 0 = ISOTROPIC
 * (Q = momentum-transfer), 1 = OKHRIMOVSKYY.  weight_method 0 = PROBABILISTIC (:361-369: each partner is updated with
 * probability w_other / w_self), 1 = CONSERVATIVE (:334-356: where the projectile is the lighter one it scatters and the
 * target, its scattered fraction and a second target of the cell are merged by ScatteringUtils::collapseThreeToTwo into
 * two equally weighted particles -- the weights of sB change). */
typedef struct {
  double const_sigma;     /* [m^2] */
  int ntab;
  const double *E, *Q, *xi;   /* host arrays of length ntab */
  int angular_scattering;
  int use_loglog_interp;
  int weight_method;
} pgpu_elastic_params;
int pgpu_collide_elastic(pgpu_species_t sA, pgpu_species_t sB, const pgpu_elastic_params *prm, double dt_sec,
                         uint64_t seed, uint64_t step, long *ncollisions);

/* Scattering::setMeanFreeTime (Scattering.H:60): nu_max = box maximum of the per-cell collision
 * frequency [Hz]; the model's m_scatter_dt is 1/nu_max (the caller min-reduces the dt over ranks as
 * ScatteringInterface.cpp:324-345 does).  TakizukaAbe::setIntraMFT/setInterMFT (TakizukaAbe.cpp:80-238),
 * Coulomb::setIntraMFT/setInterMFT (Coulomb.cpp:108-356; needs pgpu_debye_length),
 * Elastic::setInterMFT (Elastic.cpp:146-202, incl. its use of species 1's moments for both species).
 * Species need pgpu_set_moments_from_bins.  nu_max = 0 when no cell holds both species. */
int pgpu_scatter_nu_max_ta(pgpu_species_t sA, pgpu_species_t sB, double Clog, double *nu_max);
int pgpu_scatter_nu_max_coulomb(pgpu_species_t sA, pgpu_species_t sB, const pgpu_coulomb_params *prm,
                                double *nu_max);
int pgpu_scatter_nu_max_elastic(pgpu_species_t sA, pgpu_species_t sB, const pgpu_elastic_params *prm,
                                double *nu_max);

/* ScatteringUtils::computeDeltaU (ScatteringUtils.H:84-111) for explicit angles (test hook):
 * u[3n] relative velocities (component-major), one angle set per pair, dU[3n] out. */
int pgpu_scatter_delta_u(long n, const double *u, const double *costh, const double *sinth,
                         const double *cosphi, const double *sinphi, double *dU);

/* ---- instrumentation ---------------------------------------------------------- */
/* CUDA-event timing of the library's own kernels on the launch stream. */
int pgpu_profile_enable(int on);
int pgpu_profile_reset(void);
/* total milliseconds and launch count of kernels whose name starts with prefix */
int pgpu_profile_query(const char *prefix, double *ms, long *launches);
long pgpu_launch_count(void);
/* Running totals over pgpu_advance_particles_iteratively calls since the last reset:
 * particles advanced, Boris applications (m_num_apply_its) and particles left
 * unconverged -- picardParams / avg_picard_its (PicChargedSpecies.cpp:4280-4337). */
int pgpu_picard_totals(long *advances, long *apply_its, long *unconverged, int reset);
/* How many particles of `s` the specialised kernel of its last advance (or explicit step) handed to the generic
 * kernel (orbits that cross a face of the half-shifted grid, stencils at the array edge); 0 if it did not run.
 * Diagnostic: synchronises. */
int pgpu_species_deferred_count(pgpu_species_t s, long *count);

#ifdef __cplusplus
}
#endif
#endif
