// pgpu_internal.h -- host-side state of the particle engine and the launchers that
// the translation units share.  Not part of the public ABI (include/picnic_gpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/picnic_gpu.h"
#include "pgpu_device.cuh"

namespace pgpu {

// ---- kernel argument bundles (passed by value) ---------------------------------
struct GeoAny {  // runtime-D geometry; converted to Geo<D> by the launchers
  int D;
  double le[2], re[2], dx[2], rdx[2];
  int ghosts;
  int bc_lo[2], bc_hi[2];
};
template <int D>
inline Geo<D> make_geo(const GeoAny &a) {
  Geo<D> g;
  for (int d = 0; d < D; ++d) {
    g.le[d] = a.le[d];
    g.re[d] = a.re[d];
    g.dx[d] = a.dx[d];
    g.rdx[d] = a.rdx[d];
    g.bc_lo[d] = a.bc_lo[d];
    g.bc_hi[d] = a.bc_hi[d];
  }
  g.ghosts = a.ghosts;
  return g;
}

struct FieldSet {  // Ex,Ey,Ez,Bx,By,Bz in the order MeshInterp receives them
  FabView f[6];
};
struct CurrentSet {
  FabView j[3];
};
struct PartPtrs {
  double *x[2], *xold[2], *v[3], *vold[3], *w;
  double *Ep[3], *Bp[3];
};

struct AdvanceParams {
  double alpha;     // fnorm*cnormDt/2
  double cnormDt;
  double rtol;
  int iter_max;     // <0 : single pass (advanceParticles)
  int order_swap;
  double volume;    // cell volume prod(dx) (deposit)
  double rvolume;   // 1/volume
  int rel, hc;      // RELATIVISTIC_PARTICLES build of the push; Higuera-Cary gamma
  ExtFields ext;    // external fields added after every gather (addExternalFieldsToParticles)
  double fnorm;     // m_fnorm_const (the sub-orbit model forms its own alpha per sub-step)
  int suborbit;     // m_use_suborbit_model: particles left unconverged are listed (unconv_list) and deposit nothing
  int explicit_step;   // PIC_EM_EXPLICIT leap-frog step through the CC1 tile kernel (iter_max < 0, u_new = 2 ubar - u_old)
  int *unconv_list;
  unsigned *unconv_count;
};

struct Counters {          // device-resident, 64-bit
  unsigned long long apply_its;
  unsigned long long unconverged;
  unsigned long long npairs;    // collision pairs of the last collide call
  unsigned long long maxbits;   // bit pattern of a non-negative double maximum
  unsigned int err;        // ERRBIT_*
  unsigned int pad;
};

// ---- device array with capacity ---------------------------------------------------
struct DArr {
  double *p = nullptr;
  size_t cap = 0;
};

struct DeviceFab {
  double *p = nullptr;
  int lo[2] = {0, 0}, hi[2] = {0, 0};
  int n0 = 1, n1 = 1;
  int stag[2] = {0, 0};
  size_t size() const { return (size_t)n0 * n1; }
  FabView view() const { return FabView{p, lo[0], lo[1], n0, n1}; }
};

struct Context {
  int device = -1;
  bool inited = false;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;   // H2D of packed field sets (env PGPU_COPY_STREAM=0 puts them on `stream`)
  cudaEvent_t copy_fence = nullptr;     // "everything enqueued on `stream` so far", waited on by the copy stream
  int use_copy_stream = 1;
  bool exact = false;
  int deposit_mode = 1;
  int cc1_tma = 3;           // CC1 kernel: 3 = table-driven two-phase TMA tile kernel, 2 = the same reading the raw field arrays (env PGPU_CC1_TMA)
  int cc1_rsteps = 3;        // shuffle-reduction steps before the REDs: 2, 3 or 4 (env PGPU_CC1_RSTEPS)
  int cc1_minblocks = 5;     // blocks of 128 threads per SM the table kernel is compiled for: 4 (128 regs) or 5 (96 regs, needs the 228 KB carve-out); env PGPU_CC1_MINB
  int ta_staged = 1;         // Takizuka-Abe: cells of <= 128 particles per species through the shared-memory kernels (env PGPU_TA_STAGED)
  int cc1_pair = 0;          // CC1 kernel: two particles of a dual cell in lockstep through the Picard loop (env PGPU_CC1_PAIR)
  int cc1_prefetch = 0;      // L2 prefetch of a block's next particle tile (env PGPU_CC1_PREFETCH)
  int cc1_version = 2;       // table kernel generation: 1 = round-1 map (4 consecutive particles per thread in both phases), 2 = lane map in the push phase (env PGPU_CC1_V)
  int cc1_rec_per_pass = 1;  // v2: reload the dual-cell record every Picard pass instead of holding it in 32 registers (env PGPU_CC1_REC)
  int cc1_multiseg = 1;      // crossing particles of the tile kernel through the table-driven multi-segment kernel (env PGPU_CC1_MULTISEG)
  int cc1_nodecache = 0;     // v1: keep the node records of the last (dual cell, half cell) in registers too (env PGPU_CC1_NODECACHE)
  int cc1_waves = 64;        // tile kernel grid = min(tiles, SMs*4*waves) (env PGPU_CC1_WAVES)
  bool use_fast_cc1 = true;  // pgpu_set_deposit_mode(0) turns the specialised CC1 kernel off
  Counters *d_counters = nullptr;
  int *ta_list = nullptr;    // cells the staged Takizuka-Abe kernels left to the general ones: [count, cells...]
  int ta_list_cap = 0;
  Counters *h_counters = nullptr;  // pinned
  int sticky_error = 0;
  int sm_count = 148;
  // profiling
  bool profile = false;
  struct Rec {
    std::string name;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::map<std::string, std::pair<double, long>> prof;
  long launches = 0;
  // running totals since the last pgpu_picard_totals(reset): advances launched, passes applied
  long total_advances = 0, total_apply_its = 0, total_unconverged = 0;
};
Context &ctx();
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);
#define PGPU_CUDA(x)                                     \
  do {                                                   \
    cudaError_t e__ = (x);                               \
    if (e__ != cudaSuccess) return cuda_fail(e__, #x);   \
  } while (0)

// scoped kernel timer: records events around a launch when profiling is on
struct KTimer {
  const char *name;
  cudaEvent_t a = nullptr, b = nullptr;
  explicit KTimer(const char *n);
  void stop();     // close the interval now (the destructor then does nothing)
  ~KTimer();
};

}  // namespace pgpu

struct pgpu_grid_s {
  pgpu_grid_desc desc;
  pgpu::GeoAny geo;  // bc flags are per species; copied in at launch
  pgpu::DeviceFab field[6];             // the selected slot (aliases field_slot[cur_slot])
  pgpu::DeviceFab field_slot[4][6];     // resident field sets (slot 0 always allocated)
  int cur_slot = 0;
  pgpu::ExtFields ext = {};             // pgpu_grid_set_external_fields
  // pgpu_fields_set_packed copies on the library's copy stream (so that the upload of one box overlaps the particle
  // kernels of another); every reader of the field arrays orders itself behind it with fields_wait()
  cudaEvent_t upload_done = nullptr;
  mutable bool upload_pending = false;
  // per-cell coefficient tables of each field slot for the 2D CC1 kernel (pgpu_advance_cc1.cu)
  double *tab_dual[4] = {nullptr, nullptr, nullptr, nullptr};
  double *tab_node[4] = {nullptr, nullptr, nullptr, nullptr};
  bool tab_dirty[4] = {true, true, true, true};
  size_t tab_cells = 0;
  pgpu::DeviceFab jtot[3];
  pgpu::DeviceFab scratch_rho;  // (unused since the resident rho arrays; kept for pgpu_grid_destroy)
  // resident charge-density arrays, one per centring (index stag0 + 2*stag1), filled by pgpu_charge_density_deposit
  // and exchanged between boxes by a halo plan made with pgpu_halo_create_rho
  pgpu::DeviceFab rho[4];
  double *filter_tmp = nullptr;   // Q2 of the binomial filter
  size_t filter_cap = 0;
  double *debye = nullptr;      // [ncell_box]
  void *mm = nullptr;           // pgpu::MassMatrices (pgpu_massmatrix.cu), created by pgpu_mass_matrices_init
  long ncell_box = 0;
  int nbox[2] = {1, 1};
};

struct pgpu_species_s {
  pgpu_grid_t grid = nullptr;
  pgpu_species_desc desc;
  unsigned serial = 0;   // creation order of the species in this process (random-stream separation, default particle ids)
  long n = 0;
  size_t cap = 0;
  double *x[2] = {nullptr, nullptr}, *xold[2] = {nullptr, nullptr};
  double *v[3] = {nullptr, nullptr, nullptr}, *vold[3] = {nullptr, nullptr, nullptr};
  double *w = nullptr;
  uint64_t *id = nullptr;
  double *Ep[3] = {nullptr, nullptr, nullptr}, *Bp[3] = {nullptr, nullptr, nullptr};
  double *tmp = nullptr;  // permutation scratch (one array)
  pgpu::DeviceFab J[3];
  // binning
  int *cell_key = nullptr;     // [n] linear cell id inside the box (or ncell_box for outcasts)
  int *perm = nullptr;         // [n]
  int *cell_count = nullptr;   // [ncell_box+1]
  int *cell_start = nullptr;   // [ncell_box+2]
  bool binned = false;
  // lazy permutation of the *old* arrays: the cell sort gathers x, v, w, id at once and leaves
  // xold / vold in the previous order until somebody reads them (materialize_old); an
  // updateOldParticle* call in between simply overwrites them and drops the pending gather
  int *old_perm = nullptr;
  size_t old_perm_cap = 0;
  bool pos_old_pending = false, vel_old_pending = false;
  // updateOldParticlePositions/Velocities without the copy: "xold == x" ("vold == v") is only recorded.
  // The CC1 tile kernel consumes the flag (reads the old state from x / v, writes xbar / ubar into the
  // stale old arrays, then the array pointers are swapped); everybody else gets the copy from
  // materialize_old before touching either array
  bool xold_alias = false, vold_alias = false;
  int dep_from_explicit = 0;   // setCurrentDensity(a_from_explicit_solver): which gamma divides the weight (relativistic)
  int *key_sorted = nullptr;        // [n] sorted (4*cell+quadrant) keys of the last bin
  double *spare[4] = {nullptr, nullptr, nullptr, nullptr};  // gather targets of the cell sort
  size_t sort_cap = 0;
  void *cub_tmp = nullptr;
  size_t cub_bytes = 0;
  int *bin_count = nullptr;         // counting sort: [bins + 1] counts / cursors, then [bins + 1] starts
  size_t bin_count_cap = 0;
  // sub-orbit model (PicChargedSpecies m_data_suborbit, m_suborbitJ): a second, small particle container
  int use_suborbit_model = 0, suborbit_fast_particles = 0;
  double *sub[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double *sub_w = nullptr;
  uint64_t *sub_id = nullptr;
  int *sub_nsub = nullptr;
  long n_sub = 0;
  size_t sub_cap = 0;
  pgpu::DeviceFab Jsub[3];
  // outflow lists of PicChargedSpeciesBC (m_outflow_list_vector): one container, tag = 2 dir + side
  double *out[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double *out_w = nullptr;
  uint64_t *out_id = nullptr;
  int *out_tag = nullptr;
  long n_out = 0;
  size_t out_cap = 0;
  int *out_listtag = nullptr;
  size_t out_listtag_cap = 0;
  // JustinsParticle::m_pos_virt of the curvilinear pushes (dtheta, dphi); allocated by the first curvilinear call.  The
  // cell sort and the migration do not carry them: the reference re-bases them every step (rebaseVirtualPositions)
  double *virt[2] = {nullptr, nullptr};
  size_t virt_cap = 0;
  // inflow lists of PicChargedSpeciesBC (m_inflow_list_vector): one container; inf_code = 8 * numSubOrbits + (2 dir + side)
  double *inf[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  double *inf_w = nullptr;
  uint64_t *inf_id = nullptr;
  int *inf_code = nullptr;
  long n_inf = 0;
  size_t inf_cap = 0;
  pgpu::DeviceFab Jinf[3];          // m_inflowJ / m_inflowJ_virtual
  double flux_in[20] = {0};         // m_delta_{Mass,MomX,MomY,MomZ,Energy}In per boundary since the last read
  unsigned long next_id = 0;        // ids made up by pgpu_species_append
  int *unconv_list = nullptr;       // particles the last advance left unconverged (suborbit model)
  unsigned *unconv_count = nullptr;
  size_t unconv_cap = 0;
  double *enf_save = nullptr;       // Coulomb enforce_conservations: velocities before the collisions [3][enf_cap]
  size_t enf_cap = 0;
  int *defer_list = nullptr;        // particles the CC1 fast kernel left to the generic one
  unsigned *defer_count = nullptr;
  size_t defer_cap = 0;
  int *defer_list2 = nullptr;       // second list: what the table-driven multi-segment kernel hands on (the two swap roles)
  unsigned *defer_count2 = nullptr;
  size_t defer_cap2 = 0;
  unsigned *defer_count_first = nullptr;   // the count the tile kernel of the last advance wrote (diagnostic)
  // migration scratch (pgpu_exchange.cu)
  void *mig = nullptr;
  bool mig_marked = false;
  long mig_leave = 0, mig_count[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int *mig_list = nullptr;
  size_t mig_list_cap = 0;
  void *tile_box = nullptr;         // int4 per particle tile: table window of the CC1 kernel
  size_t tile_box_cap = 0;
  double *dens = nullptr, *mom = nullptr, *ene = nullptr;  // [ncell],[3 ncell],[3 ncell]
  pgpu::PartPtrs ptrs() const {
    pgpu::PartPtrs p;
    for (int d = 0; d < 2; ++d) {
      p.x[d] = x[d];
      p.xold[d] = xold[d];
    }
    for (int c = 0; c < 3; ++c) {
      p.v[c] = v[c];
      p.vold[c] = vold[c];
      p.Ep[c] = Ep[c];
      p.Bp[c] = Bp[c];
    }
    p.w = w;
    return p;
  }
};

namespace pgpu {
GeoAny species_geo(const pgpu_species_s *s);
FieldSet grid_fields(const pgpu_grid_s *g);   // waits for a pending packed upload (fields_wait)
int fields_wait(const pgpu_grid_s *g);        // order the library stream behind a pending pgpu_fields_set_packed
CurrentSet species_current(const pgpu_species_s *s);

int scale_fab(const DeviceFab &f, double s);
int fold_periodic(const pgpu_grid_s *g, const DeviceFab &f);
// the grid's resident charge-density array of one centring (allocated on first use)
int grid_rho_fab(pgpu_grid_s *g, const int *stag, DeviceFab **out);
int copy_fab_to_host(const DeviceFab &f, int D, double *data, const int *lo, const int *hi, bool sync = true);

// launchers implemented in the kernel translation units
// keep: bit 0 = leave an aliased xold alone, bit 1 = leave an aliased vold alone 
// bit 2 = leave the pending gathers of the last cell sort pending (the caller touches no old array)
enum { KEEP_XOLD_ALIAS = 1, KEEP_VOLD_ALIAS = 2, KEEP_OLD_ALIASES = 3, KEEP_PENDING = 4 };
int materialize_old(pgpu_species_s *s, int keep = 0);
int grow_capacity(pgpu_species_s *s, long n);   // keeps the particles (pgpu_api.cu)
int launch_gather(pgpu_species_s *s);
int launch_add_external(pgpu_species_s *s);
int ensure_unconv_list(pgpu_species_s *s);                               // pgpu_suborbit.cu
int transfer_listed_to_suborbit(pgpu_species_s *s, unsigned count);      // pgpu_suborbit.cu
int inject_inflow(pgpu_species_s *s, const int *bc_lo, const int *bc_hi);
int transfer_outflow(pgpu_species_s *s, const int *bc_lo, const int *bc_hi);   // pgpu_suborbit.cu
PartPtrs outflow_part_ptrs(pgpu_species_s *s);
int launch_suborbit(pgpu_species_s *s, const AdvanceParams &prm, int from_jac, const DeviceFab *Jsub, unsigned *nfail,
                    int inflow = 0);
int launch_explicit_step(pgpu_species_s *s, const AdvanceParams &prm, const int *periodic, bool second_half,
                         bool deferred = false);
int launch_deposit_current(pgpu_species_s *s, double cnormDt);
int launch_deposit_outflow(pgpu_species_s *s);   // the outflow lists into s->J (explicit solver)
int launch_advance(pgpu_species_s *s, const AdvanceParams &prm, bool fuse_deposit);
int launch_advance_cc1_fast(pgpu_species_s *s, const AdvanceParams &prm, bool deposit);
int launch_advance_cc1_1d_fast(pgpu_species_s *s, const AdvanceParams &prm, bool deposit);
void mm_destroy(pgpu_grid_s *g);   // pgpu_massmatrix.cu
}  // namespace pgpu
