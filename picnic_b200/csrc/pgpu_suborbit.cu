// pgpu_suborbit.cu -- the sub-orbit model of the implicit particle advance.
//
// PICNIC keeps a second particle container per species, m_data_suborbit, for the particles the particle Picard loop
// cannot converge (PicChargedSpecies::advanceParticlesIteratively, PicChargedSpecies.cpp:1699-1706) and for "fast"
// particles that cross more cell faces than the CC1 mass matrices allow (transferFastParticles, :894-956).  Those
// particles take the time step in 2, 3, ... equal sub-steps (advanceSubOrbitParticlesAndSetJ, :3324-3669, the kernel is in
// pgpu_push.cu), deposit into their own current m_suborbitJ, which PicSpeciesInterface::addSubOrbitJ adds to the total
// (PicSpeciesInterface.cpp:1538-1590), and return to the main container at the end of the step (mergeSubOrbitParticles,
// :1718-1747).  Here the container is a second, small SoA next to the species' main arrays; moving particles between
// the two is a gather into the tail of one and a hole fill from the tail of the other.
#include <algorithm>
#include <vector>

#include "pgpu_internal.h"

using namespace pgpu;

#define NEED_INIT()                                  \
  if (!ctx().inited) {                               \
    set_error("pgpu_init has not been called");      \
    return PGPU_ERR_STATE;                           \
  }

namespace {
inline unsigned nb(long n) { return (unsigned)((n + 255) / 256); }

struct MainPtrs {
  double *a[10];   // x0 x1 xold0 xold1 v0 v1 v2 vold0 vold1 vold2 (entries of unused directions are null)
  double *w;
  uint64_t *id;
};
MainPtrs main_ptrs(pgpu_species_s *s) {
  MainPtrs P;
  for (int d = 0; d < 2; ++d) {
    P.a[d] = s->x[d];
    P.a[2 + d] = s->xold[d];
  }
  for (int c = 0; c < 3; ++c) {
    P.a[4 + c] = s->v[c];
    P.a[7 + c] = s->vold[c];
  }
  P.w = s->w;
  P.id = s->id;
  return P;
}
MainPtrs sub_ptrs(pgpu_species_s *s) {
  MainPtrs P;
  for (int k = 0; k < 10; ++k) P.a[k] = s->sub[k];
  P.w = s->sub_w;
  P.id = s->sub_id;
  return P;
}
MainPtrs out_ptrs(pgpu_species_s *s) {
  MainPtrs P;
  for (int k = 0; k < 10; ++k) P.a[k] = s->out[k];
  P.w = s->out_w;
  P.id = s->out_id;
  return P;
}
// a side container of a species: the sub-orbit container, or the outflow lists of PicChargedSpeciesBC
struct Aux {
  double **a;
  double **w;
  uint64_t **id;
  int **tag;
  long *n;
  size_t *cap;
};
Aux sub_aux(pgpu_species_s *s) { return Aux{s->sub, &s->sub_w, &s->sub_id, &s->sub_nsub, &s->n_sub, &s->sub_cap}; }
Aux out_aux(pgpu_species_s *s) { return Aux{s->out, &s->out_w, &s->out_id, &s->out_tag, &s->n_out, &s->out_cap}; }
Aux inf_aux(pgpu_species_s *s) { return Aux{s->inf, &s->inf_w, &s->inf_id, &s->inf_code, &s->n_inf, &s->inf_cap}; }
MainPtrs inf_ptrs(pgpu_species_s *s) {
  MainPtrs P;
  for (int k = 0; k < 10; ++k) P.a[k] = s->inf[k];
  P.w = s->inf_w;
  P.id = s->inf_id;
  return P;
}
// PicChargedSpeciesBC::inflow_Lo / inflow_Hi, final (not intermediate) advance (:961-1001, 1047-1086): the inflow-list
// particle is time-centred; its new-time state 2 x - x_old, 2 u - u_old decides: inside the boundary plane it joins the
// species (slot base + rank in `to`), otherwise (turned around) it is dropped.  Flux probes m_delta_*In from u_old.
__global__ void k_inflow_inject(MainPtrs from, const int *code, long n, MainPtrs to, long base, int D, double l0, double r0,
                                double l1, double r1, unsigned mask /* bit (2 dir + side): boundary is inflow_outflow */,
                                int rel, unsigned *count, double *flux /* [4][5] */) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int tag = code[i] & 7, dir = tag >> 1, side = tag & 1;
  if (!((mask >> tag) & 1u)) return;   // not an inflow boundary in this call: dropped with the rest of the list
  double xn[2] = {0.0, 0.0}, un[3];
  for (int d = 0; d < D; ++d) xn[d] = __dsub_rn(__dmul_rn(2.0, from.a[d][i]), from.a[2 + d][i]);
  for (int c = 0; c < 3; ++c) un[c] = __dsub_rn(__dmul_rn(2.0, from.a[4 + c][i]), from.a[7 + c][i]);
  const double edge = dir == 0 ? (side == 0 ? l0 : r0) : (side == 0 ? l1 : r1);
  const bool inside = side == 0 ? (xn[dir] >= edge) : (xn[dir] < edge);
  if (!inside) return;
  const long o = base + atomicAdd(count, 1u);
  for (int d = 0; d < D; ++d) {
    to.a[d][o] = xn[d];
    to.a[2 + d][o] = from.a[2 + d][i];
  }
  for (int c = 0; c < 3; ++c) {
    to.a[4 + c][o] = un[c];
    to.a[7 + c][o] = from.a[7 + c][i];
  }
  const double w = from.w[i];
  to.w[o] = w;
  to.id[o] = from.id[i];
  const double a = from.a[7][i], b = from.a[8][i], c = from.a[9][i];
  const double gbsq = a * a + b * b + c * c, gamma = rel ? sqrt(1.0 + gbsq) : 1.0;
  atomicAdd(flux + tag * 5 + 0, w);
  atomicAdd(flux + tag * 5 + 1, w * a);
  atomicAdd(flux + tag * 5 + 2, w * b);
  atomicAdd(flux + tag * 5 + 3, w * c);
  atomicAdd(flux + tag * 5 + 4, w * gbsq / (gamma + 1.0));
}
__global__ void k_fill_int(int *p, long n, int v) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// listed particles of `from` -> positions base .. base+count-1 of `to`
__global__ void k_copy_listed(MainPtrs from, MainPtrs to, const int *list, unsigned count, long base, int *nsub, int D,
                              int *dead, const int *listtag) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const long i = list[t], o = base + t;
  for (int k = 0; k < 10; ++k) {
    if ((k < 4) && ((k & 1) >= D)) continue;   // x1 / xold1 in 1D
    to.a[k][o] = from.a[k][i];
  }
  to.w[o] = from.w[i];
  to.id[o] = from.id[i];
  if (nsub) nsub[o] = listtag ? listtag[t] : 2;   // setNumSubOrbits(2); outflow lists: 2 dir + side
  if (dead) dead[i] = 1;
}
// the usual hole fill: survivors of the tail [new_n, n) move into the holes below new_n
__global__ void k_list_holes_movers(const int *dead, long n, long new_n, const int *list, unsigned count, int *holes,
                                    int *movers, unsigned *nh, unsigned *nm) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  const long i = list[t];
  if (i < new_n) holes[atomicAdd(nh, 1u)] = (int)i;
  const long j = new_n + t;
  if (j < n && !dead[j]) movers[atomicAdd(nm, 1u)] = (int)j;
}
__global__ void k_fill(MainPtrs P, const int *holes, const int *movers, const unsigned *nh, int D) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *nh) return;
  const long dst = holes[t], src = movers[t];
  for (int k = 0; k < 10; ++k) {
    if ((k < 4) && ((k & 1) >= D)) continue;
    P.a[k][dst] = P.a[k][src];
  }
  P.w[dst] = P.w[src];
  P.id[dst] = P.id[src];
}
__global__ void k_clear_dead(int *dead, const int *list, unsigned count) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) dead[list[t]] = 0;
}
// PicChargedSpecies::transferFastParticles (:894-956): more than ghosts - D face crossings of the half-shifted grid
__global__ void k_flag_fast(const double *x0, const double *x1, const double *xo0, const double *xo1, long n, int D,
                            double dx0, double dx1, int max_crossings, int *list, unsigned *count) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool fast = false;
  for (int d = 0; d < D; ++d) {
    const double xb = d ? x1[i] : x0[i], xo = d ? xo1[i] : xo0[i], dx = d ? dx1 : dx0;
    const double xn = __dsub_rn(__dmul_rn(2.0, xb), xo);
    // as the reference writes it (:930-931): no domain left edge in the index
    const int io = __double2int_rd(__ddiv_rn(__dsub_rn(xo, __dmul_rn(0.5, dx)), dx));
    const int in = __double2int_rd(__ddiv_rn(__dsub_rn(xn, __dmul_rn(0.5, dx)), dx));
    if (abs(in - io) > max_crossings) fast = true;
  }
  if (fast) list[atomicAdd(count, 1u)] = (int)i;
}

int ensure_aux_cap(const Aux &A, long need) {
  if ((size_t)need <= *A.cap) return 0;
  const size_t cap = (size_t)(need + need / 4 + 4096);
  cudaStream_t st = ctx().stream;
  PGPU_CUDA(cudaStreamSynchronize(st));
  const long have = *A.n;
  auto re = [&](void **p, size_t elem) -> int {
    void *q = nullptr;
    if (cudaMalloc(&q, cap * elem) != cudaSuccess) return PGPU_ERR_CUDA;
    if (*p) {
      cudaMemcpy(q, *p, (size_t)have * elem, cudaMemcpyDeviceToDevice);
      cudaFree(*p);
    }
    *p = q;
    return 0;
  };
  for (int k = 0; k < 10; ++k)
    if (re(reinterpret_cast<void **>(&A.a[k]), sizeof(double))) return PGPU_ERR_CUDA;
  if (re(reinterpret_cast<void **>(A.w), sizeof(double))) return PGPU_ERR_CUDA;
  if (re(reinterpret_cast<void **>(A.id), sizeof(uint64_t))) return PGPU_ERR_CUDA;
  if (re(reinterpret_cast<void **>(A.tag), sizeof(int))) return PGPU_ERR_CUDA;
  *A.cap = cap;
  return 0;
}
// first boundary (dir 0 lo, dir 0 hi, dir 1 lo, dir 1 hi) with an outflow BC the particle is beyond:
// PicChargedSpeciesBC::outflow_Lo / outflow_Hi (PicChargedSpeciesBC.cpp:872-918): x < Xmin, x >= Xmax
__global__ void k_flag_outflow(const double *x0, const double *x1, long n, int D, double l0, double r0, double l1, double r1,
                               int lo0, int hi0, int lo1, int hi1, int *list, int *listtag, unsigned *count) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int tag = -1;
  const double a = x0[i];
  if (lo0 && a < l0) tag = 0;
  else if (hi0 && a >= r0) tag = 1;
  else if (D == 2) {
    const double b = x1[i];
    if (lo1 && b < l1) tag = 2;
    else if (hi1 && b >= r1) tag = 3;
  }
  if (tag >= 0) {
    const unsigned slot = atomicAdd(count, 1u);
    list[slot] = (int)i;
    listtag[slot] = tag;
  }
}
// PicChargedSpeciesBC::outflow / inflow flux diagnostics (:935-947 pattern): sums of w, w u_old, w |u_old|^2 / (gamma + 1) per tag
__global__ void k_aux_flux(MainPtrs P, const int *tag, long n, int rel, double *out /* [4][5] */) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tag[i];
  if (t < 0 || t > 3) return;
  const double w = P.w[i], a = P.a[7][i], b = P.a[8][i], c = P.a[9][i];
  const double gbsq = a * a + b * b + c * c, gamma = rel ? sqrt(1.0 + gbsq) : 1.0;
  atomicAdd(out + t * 5 + 0, w);
  atomicAdd(out + t * 5 + 1, w * a);
  atomicAdd(out + t * 5 + 2, w * b);
  atomicAdd(out + t * 5 + 3, w * c);
  atomicAdd(out + t * 5 + 4, w * gbsq / (gamma + 1.0));
}
}  // namespace

namespace pgpu {
// the inflow-list part of PicChargedSpeciesBC::apply (:199-224), final advance: see k_inflow_inject.  The list is emptied
// (joined or dropped), as the reference leaves it.  The order of the injected particles among themselves is the order of
// the atomics, i.e. not reproducible; the next cell sort removes it.
int inject_inflow(pgpu_species_s *s, const int *bc_lo, const int *bc_hi) {
  if (s->n_inf == 0) return 0;
  const pgpu_grid_s *g = s->grid;
  const int D = g->desc.D;
  unsigned mask = 0;
  for (int d = 0; d < D; ++d) {
    if (bc_lo[d] == PGPU_BC_INFLOW_OUTFLOW) mask |= 1u << (2 * d);
    if (bc_hi[d] == PGPU_BC_INFLOW_OUTFLOW) mask |= 1u << (2 * d + 1);
  }
  if (!mask) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  int rc = grow_capacity(s, s->n + s->n_inf);
  if (rc) return rc;
  cudaStream_t st = ctx().stream;
  double *dflux = nullptr;
  PGPU_CUDA(cudaMalloc(&dflux, 20 * sizeof(double) + sizeof(unsigned)));
  PGPU_CUDA(cudaMemsetAsync(dflux, 0, 20 * sizeof(double) + sizeof(unsigned), st));
  unsigned *dcount = reinterpret_cast<unsigned *>(dflux + 20);
  k_inflow_inject<<<nb(s->n_inf), 256, 0, st>>>(inf_ptrs(s), s->inf_code, s->n_inf, main_ptrs(s), s->n, D, g->geo.le[0],
                                                g->geo.re[0], g->geo.le[1], g->geo.re[1], mask, s->desc.relativistic, dcount,
                                                dflux);
  double hflux[20];
  unsigned joined = 0;
  PGPU_CUDA(cudaMemcpyAsync(hflux, dflux, sizeof(hflux), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(&joined, dcount, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(dflux);
  for (int k = 0; k < 20; ++k) s->flux_in[k] += hflux[k];
  s->n += joined;
  s->n_inf = 0;
  s->binned = false;
  return 0;
}

// scratch of the unconverged / fast list: list [cap] + holes [cap] + movers [cap] ints, two counters behind the count
int ensure_unconv_list(pgpu_species_s *s) {
  if (s->unconv_list && s->unconv_cap >= s->cap) return 0;
  if (s->unconv_list) cudaFree(s->unconv_list);
  if (!s->unconv_count) PGPU_CUDA(cudaMalloc(&s->unconv_count, 4 * sizeof(unsigned)));
  PGPU_CUDA(cudaMalloc(&s->unconv_list, 3 * s->cap * sizeof(int)));
  s->unconv_cap = s->cap;
  return 0;
}

// move the particles listed in s->unconv_list[0 .. count) from the main container to a side container
static int transfer_listed(pgpu_species_s *s, const Aux &A, const MainPtrs &to, unsigned count, const int *listtag) {
  if (count == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  if (ensure_aux_cap(A, *A.n + (long)count)) return PGPU_ERR_CUDA;
  cudaStream_t st = ctx().stream;
  const int D = s->grid->desc.D;
  int *dead = s->cell_key;   // [cap] ints, free between cell sorts (the sort is invalidated below)
  int *list = s->unconv_list, *holes = list + s->unconv_cap, *movers = holes + s->unconv_cap;
  unsigned *nh = s->unconv_count + 1, *nm = s->unconv_count + 2;
  const long new_n = s->n - (long)count;
  // only the tail [new_n, n) is ever tested for "dead": clear just that
  PGPU_CUDA(cudaMemsetAsync(dead + new_n, 0, (size_t)count * sizeof(int), st));
  PGPU_CUDA(cudaMemsetAsync(nh, 0, 2 * sizeof(unsigned), st));
  KTimer t("suborbit_transfer");
  MainPtrs dst = to;   // the arrays may just have been re-allocated
  for (int k = 0; k < 10; ++k) dst.a[k] = A.a[k];
  dst.w = *A.w;
  dst.id = *A.id;
  k_copy_listed<<<nb(count), 256, 0, st>>>(main_ptrs(s), dst, list, count, *A.n, *A.tag, D, dead, listtag);
  k_list_holes_movers<<<nb(count), 256, 0, st>>>(dead, s->n, new_n, list, count, holes, movers, nh, nm);
  k_fill<<<nb(count), 256, 0, st>>>(main_ptrs(s), holes, movers, nh, D);
  s->n = new_n;
  *A.n += (long)count;
  s->binned = false;
  return 0;
}
int transfer_listed_to_suborbit(pgpu_species_s *s, unsigned count) {
  return transfer_listed(s, sub_aux(s), sub_ptrs(s), count, nullptr);
}
// PicChargedSpeciesBC::apply, the outflow part (:187-224): particles beyond an outflow (or inflow_outflow) boundary leave
// the main container for the outflow lists (one container, tagged 2 dir + side)
int transfer_outflow(pgpu_species_s *s, const int *bc_lo, const int *bc_hi) {
  const pgpu_grid_s *g = s->grid;
  const int D = g->desc.D;
  auto is_out = [](int bc) { return bc == PGPU_BC_OUTFLOW || bc == PGPU_BC_INFLOW_OUTFLOW; };
  const int lo0 = is_out(bc_lo[0]), hi0 = is_out(bc_hi[0]);
  const int lo1 = D == 2 && is_out(bc_lo[1]), hi1 = D == 2 && is_out(bc_hi[1]);
  if (!(lo0 || hi0 || lo1 || hi1) || s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  if (ensure_unconv_list(s)) return PGPU_ERR_CUDA;
  cudaStream_t st = ctx().stream;
  int *listtag = s->unconv_list + 2 * s->unconv_cap;   // the "movers" third of the scratch is free until the transfer
  if (!s->out_listtag || s->out_listtag_cap < s->cap) {
    if (s->out_listtag) cudaFree(s->out_listtag);
    PGPU_CUDA(cudaMalloc(&s->out_listtag, s->cap * sizeof(int)));
    s->out_listtag_cap = s->cap;
  }
  listtag = s->out_listtag;
  PGPU_CUDA(cudaMemsetAsync(s->unconv_count, 0, sizeof(unsigned), st));
  {
    KTimer t("bc_outflow");
    k_flag_outflow<<<nb(s->n), 256, 0, st>>>(s->x[0], s->x[1], s->n, D, g->geo.le[0], g->geo.re[0], g->geo.le[1], g->geo.re[1],
                                             lo0, hi0, lo1, hi1, s->unconv_list, listtag, s->unconv_count);
  }
  unsigned count = 0;
  PGPU_CUDA(cudaMemcpyAsync(&count, s->unconv_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  return transfer_listed(s, out_aux(s), out_ptrs(s), count, listtag);
}
// setCurrentDensity(dt, from_explicit_solver = true) adds the current of the outflow lists before the scaling
// (PicChargedSpecies.cpp:3232-3235 -> PicChargedSpeciesBC::depositInflowOutflowJ, PicChargedSpeciesBC.cpp:667-736)
PartPtrs outflow_part_ptrs(pgpu_species_s *s) {
  PartPtrs p;
  for (int d = 0; d < 2; ++d) {
    p.x[d] = s->out[d];
    p.xold[d] = s->out[2 + d];
  }
  for (int c = 0; c < 3; ++c) {
    p.v[c] = s->out[4 + c];
    p.vold[c] = s->out[7 + c];
    p.Ep[c] = p.Bp[c] = nullptr;
  }
  p.w = s->out_w;
  return p;
}
}  // namespace pgpu

extern "C" {

int pgpu_species_set_suborbit_model(pgpu_species_t s, int use_suborbit_model, int suborbit_fast_particles) {
  if (!s) return PGPU_ERR_ARG;
  s->use_suborbit_model = use_suborbit_model ? 1 : 0;
  s->suborbit_fast_particles = suborbit_fast_particles ? 1 : 0;
  return 0;
}

long pgpu_species_suborbit_count(pgpu_species_t s) { return s ? s->n_sub : -1; }

int pgpu_transfer_fast_particles(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->suborbit_fast_particles || s->desc.interp_J != CC1 || s->n == 0) return 0;   // :896-897
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  if (ensure_unconv_list(s)) return PGPU_ERR_CUDA;
  cudaStream_t st = ctx().stream;
  const pgpu_grid_s *g = s->grid;
  PGPU_CUDA(cudaMemsetAsync(s->unconv_count, 0, sizeof(unsigned), st));
  k_flag_fast<<<nb(s->n), 256, 0, st>>>(s->x[0], s->x[1], s->xold[0], s->xold[1], s->n, g->desc.D, g->geo.dx[0], g->geo.dx[1],
                                        g->desc.nghost - g->desc.D, s->unconv_list, s->unconv_count);
  unsigned count = 0;
  PGPU_CUDA(cudaMemcpyAsync(&count, s->unconv_count, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  return transfer_listed_to_suborbit(s, count);
}

int pgpu_advance_suborbit_particles_and_set_J(pgpu_species_t s, double dt, int from_emjacobian) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (s->desc.relativistic) {
    set_error("the sub-orbit model is implemented for the non-relativistic planar push");
    return PGPU_ERR_STATE;
  }
  cudaStream_t st = ctx().stream;
  for (int c = 0; c < 3; ++c) {
    if (!s->Jsub[c].p) {
      s->Jsub[c] = s->J[c];
      s->Jsub[c].p = nullptr;
      PGPU_CUDA(cudaMalloc(&s->Jsub[c].p, s->J[c].size() * sizeof(double)));
    }
    PGPU_CUDA(cudaMemsetAsync(s->Jsub[c].p, 0, s->Jsub[c].size() * sizeof(double), st));   // SpaceUtils::zero (:3330-3331)
  }
  if (s->n_sub == 0) return 0;
  if (!s->unconv_count) PGPU_CUDA(cudaMalloc(&s->unconv_count, 4 * sizeof(unsigned)));
  unsigned *nfail = s->unconv_count + 3;
  PGPU_CUDA(cudaMemsetAsync(nfail, 0, sizeof(unsigned), st));
  AdvanceParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.cnormDt = dt * s->desc.cvac_norm;
  prm.fnorm = s->desc.fnorm_const;
  prm.alpha = s->desc.fnorm_const * prm.cnormDt / 2.0;
  prm.rtol = s->desc.rtol;
  prm.iter_max = s->desc.iter_max;
  const GeoAny &g = s->grid->geo;
  prm.volume = (g.D == 1) ? g.dx[0] : g.dx[0] * g.dx[1];
  prm.rvolume = 1.0 / prm.volume;
  prm.ext = s->grid->ext;
  int rc = launch_suborbit(s, prm, from_emjacobian ? 1 : 0, s->Jsub, nfail);
  if (rc) return rc;
  // multiply by charge / volume_scale (:3366-3372)
  const double f = s->desc.charge / s->grid->desc.volume_scale;
  for (int c = 0; c < 3; ++c) {
    rc = scale_fab(s->Jsub[c], f);
    if (rc) return rc;
  }
  unsigned failed = 0;
  PGPU_CUDA(cudaMemcpyAsync(&failed, nfail, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  if (failed) {
    set_error("sub-orbit model: %u particles did not converge with 512 sub-orbits or left the ghosted arrays", failed);
    return PGPU_ERR_STATE;
  }
  return 0;
}

int pgpu_species_suborbit_current_get(pgpu_species_t s, int comp, double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!s || comp < 0 || comp >= 3 || !s->Jsub[comp].p) return PGPU_ERR_ARG;
  int rc = copy_fab_to_host(s->Jsub[comp], s->grid->desc.D, data, lo, hi);
  if (rc) return rc;
  return pgpu_synchronize();
}

// PicSpeciesInterface::addSubOrbitJ (PicSpeciesInterface.cpp:1538-1590): total J += this species' sub-orbit J
__global__ void k_add_arr(double *a, const double *b, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = __dadd_rn(a[i], b[i]);
}
int pgpu_current_add_suborbit(pgpu_grid_t g, pgpu_species_t s) {
  NEED_INIT();
  if (!g || !s) return PGPU_ERR_ARG;
  if (!s->use_suborbit_model || s->desc.charge == 0.0 || !s->Jsub[0].p) return 0;
  for (int c = 0; c < 3; ++c) {
    const long n = (long)g->jtot[c].size();
    k_add_arr<<<nb(n), 256, 0, ctx().stream>>>(g->jtot[c].p, s->Jsub[c].p, n);
  }
  return 0;
}

int pgpu_merge_suborbit_particles(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->use_suborbit_model || s->n_sub == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  int rc = grow_capacity(s, s->n + s->n_sub);
  if (rc) return rc;
  cudaStream_t st = ctx().stream;
  const int D = s->grid->desc.D;
  const MainPtrs M = main_ptrs(s), S = sub_ptrs(s);
  for (int k = 0; k < 10; ++k) {
    if ((k < 4) && ((k & 1) >= D)) continue;
    PGPU_CUDA(cudaMemcpyAsync(M.a[k] + s->n, S.a[k], (size_t)s->n_sub * sizeof(double), cudaMemcpyDeviceToDevice, st));
  }
  PGPU_CUDA(cudaMemcpyAsync(M.w + s->n, S.w, (size_t)s->n_sub * sizeof(double), cudaMemcpyDeviceToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(M.id + s->n, S.id, (size_t)s->n_sub * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
  s->n += s->n_sub;
  s->n_sub = 0;
  s->binned = false;
  return 0;
}

static int aux_download(pgpu_species_t s, const Aux &A, double *x, double *xold, double *v, double *vold, double *w,
                        uint64_t *id, int *tag) {
  const long n = *A.n;
  if (n == 0) return 0;
  cudaStream_t st = ctx().stream;
  const int D = s->grid->desc.D;
  for (int d = 0; d < D; ++d) {
    if (x) PGPU_CUDA(cudaMemcpyAsync(x + d * n, A.a[d], n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (xold) PGPU_CUDA(cudaMemcpyAsync(xold + d * n, A.a[2 + d], n * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  for (int c = 0; c < 3; ++c) {
    if (v) PGPU_CUDA(cudaMemcpyAsync(v + c * n, A.a[4 + c], n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (vold) PGPU_CUDA(cudaMemcpyAsync(vold + c * n, A.a[7 + c], n * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  if (w) PGPU_CUDA(cudaMemcpyAsync(w, *A.w, n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (id) PGPU_CUDA(cudaMemcpyAsync(id, *A.id, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (tag) PGPU_CUDA(cudaMemcpyAsync(tag, *A.tag, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  return 0;
}
int pgpu_species_suborbit_download(pgpu_species_t s, double *x, double *xold, double *v, double *vold, double *w,
                                   uint64_t *id, int *nsub) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  return aux_download(s, sub_aux(s), x, xold, v, vold, w, id, nsub);
}

// ---- outflow lists (PicChargedSpeciesBC m_outflow_list_vector) ------------------------------------------------------
long pgpu_species_outflow_count(pgpu_species_t s) { return s ? s->n_out : -1; }
int pgpu_species_outflow_download(pgpu_species_t s, double *x, double *xold, double *v, double *vold, double *w,
                                  uint64_t *id, int *boundary) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  return aux_download(s, out_aux(s), x, xold, v, vold, w, id, boundary);
}
// m_delta_{Mass,MomX,MomY,MomZ,Energy}Out per boundary (2 dir + side): sums of w, w u_old, w |u_old|^2 / (gamma + 1)
int pgpu_species_outflow_fluxes(pgpu_species_t s, double *flux20) {
  NEED_INIT();
  if (!s || !flux20) return PGPU_ERR_ARG;
  for (int k = 0; k < 20; ++k) flux20[k] = 0.0;
  if (s->n_out == 0) return 0;
  double *d = nullptr;
  PGPU_CUDA(cudaMalloc(&d, 20 * sizeof(double)));
  PGPU_CUDA(cudaMemsetAsync(d, 0, 20 * sizeof(double), ctx().stream));
  k_aux_flux<<<nb(s->n_out), 256, 0, ctx().stream>>>(out_ptrs(s), s->out_tag, s->n_out, s->desc.relativistic, d);
  PGPU_CUDA(cudaMemcpyAsync(flux20, d, 20 * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
  PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
  cudaFree(d);
  return 0;
}
// PicChargedSpeciesBC::removeOutflowParticles (:508-545): the lists are emptied (the surface charge they leave behind is
// the host's: download the lists first)
int pgpu_remove_outflow_particles(pgpu_species_t s) {
  if (!s) return PGPU_ERR_ARG;
  s->n_out = 0;
  return 0;
}
// inflow: the host's InflowBC objects create the particles (PicChargedSpeciesBC::createInflowParticles, :467-506, host
// logic with the reference's own generator); injectInflowParticles (:563-665) puts them into the main container
int pgpu_species_append(pgpu_species_t s, long n, const double *x, const double *xold, const double *v, const double *vold,
                        const double *w, const uint64_t *id) {
  NEED_INIT();
  if (!s || n < 0 || (n && (!x || !v || !w))) return PGPU_ERR_ARG;
  if (n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  int rc = grow_capacity(s, s->n + n);
  if (rc) return rc;
  cudaStream_t st = ctx().stream;
  const int D = s->grid->desc.D;
  const long o = s->n;
  for (int d = 0; d < D; ++d) {
    PGPU_CUDA(cudaMemcpyAsync(s->x[d] + o, x + d * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaMemcpyAsync(s->xold[d] + o, (xold ? xold : x) + d * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  for (int c = 0; c < 3; ++c) {
    PGPU_CUDA(cudaMemcpyAsync(s->v[c] + o, v + c * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaMemcpyAsync(s->vold[c] + o, (vold ? vold : v) + c * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  PGPU_CUDA(cudaMemcpyAsync(s->w + o, w, n * sizeof(double), cudaMemcpyHostToDevice, st));
  if (id) PGPU_CUDA(cudaMemcpyAsync(s->id + o, id, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  else {
    std::vector<uint64_t> tmp((size_t)n);
    for (long k = 0; k < n; ++k) tmp[k] = ((uint64_t)s->serial << 40) + (uint64_t)(s->next_id++) + (1ull << 39);
    PGPU_CUDA(cudaMemcpyAsync(s->id + o, tmp.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaStreamSynchronize(st));
  }
  PGPU_CUDA(cudaStreamSynchronize(st));   // the host buffers are borrowed for the call only
  s->n += n;
  s->binned = false;
  return 0;
}

// ---- inflow lists with pic_species.N.suborbit_inflow_J (PicChargedSpeciesBC m_inflow_list_vector) --------------------
// createInflowParticles (:467-506) stays on the host (it draws from the host generator); with suborbit_inflow_J the
// particles are NOT injected at once (PicChargedSpecies::injectInflowParticles returns, :1787): they wait in the inflow list
// of their boundary, every nonlinear evaluation advances them from there (advanceInflowParticlesAndSetJ), and the applyBCs
// at the end of the step lets them join the species (inflow_Lo / inflow_Hi).
int pgpu_species_inflow_append(pgpu_species_t s, long n, const double *x, const double *v, const double *w,
                               const uint64_t *id, int bdry_dir, int bdry_side) {
  NEED_INIT();
  if (!s || n < 0 || (n && (!x || !v || !w))) return PGPU_ERR_ARG;
  const int D = s->grid->desc.D;
  if (bdry_dir < 0 || bdry_dir >= D || bdry_side < 0 || bdry_side > 1) return PGPU_ERR_ARG;
  if (n == 0) return 0;
  const Aux A = inf_aux(s);
  int rc = ensure_aux_cap(A, s->n_inf + n);
  if (rc) return rc;
  cudaStream_t st = ctx().stream;
  const long o = s->n_inf;
  for (int d = 0; d < D; ++d) {
    PGPU_CUDA(cudaMemcpyAsync(s->inf[d] + o, x + d * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaMemcpyAsync(s->inf[2 + d] + o, x + d * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  for (int c = 0; c < 3; ++c) {
    PGPU_CUDA(cudaMemcpyAsync(s->inf[4 + c] + o, v + c * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaMemcpyAsync(s->inf[7 + c] + o, v + c * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  PGPU_CUDA(cudaMemcpyAsync(s->inf_w + o, w, n * sizeof(double), cudaMemcpyHostToDevice, st));
  std::vector<uint64_t> tmp;
  if (!id) {
    tmp.resize((size_t)n);
    for (long k = 0; k < n; ++k) tmp[k] = ((uint64_t)s->serial << 40) + (uint64_t)(s->next_id++) + (1ull << 39);
    id = tmp.data();
  }
  PGPU_CUDA(cudaMemcpyAsync(s->inf_id + o, id, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  k_fill_int<<<nb(n), 256, 0, st>>>(s->inf_code + o, n, 8 + 2 * bdry_dir + bdry_side);   // one sub-orbit
  PGPU_CUDA(cudaStreamSynchronize(st));
  s->n_inf += n;
  return 0;
}
long pgpu_species_inflow_count(pgpu_species_t s) { return s ? s->n_inf : -1; }
// nsub_boundary[i] = 8 * numSubOrbits + (2 dir + side)
int pgpu_species_inflow_download(pgpu_species_t s, double *x, double *xold, double *v, double *vold, double *w,
                                 uint64_t *id, int *nsub_boundary) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  return aux_download(s, inf_aux(s), x, xold, v, vold, w, id, nsub_boundary);
}
int pgpu_species_inflow_clear(pgpu_species_t s) {
  if (!s) return PGPU_ERR_ARG;
  s->n_inf = 0;
  return 0;
}

int pgpu_advance_inflow_particles_and_set_J(pgpu_species_t s, double dt, int from_emjacobian) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (s->desc.relativistic) {
    set_error("the inflow sub-orbit path is implemented for the non-relativistic planar push");
    return PGPU_ERR_STATE;
  }
  cudaStream_t st = ctx().stream;
  for (int c = 0; c < 3; ++c) {
    if (!s->Jinf[c].p) {
      s->Jinf[c] = s->J[c];
      s->Jinf[c].p = nullptr;
      PGPU_CUDA(cudaMalloc(&s->Jinf[c].p, s->J[c].size() * sizeof(double)));
    }
    PGPU_CUDA(cudaMemsetAsync(s->Jinf[c].p, 0, s->Jinf[c].size() * sizeof(double), st));   // SpaceUtils::zero (:3262-3263)
  }
  if (s->n_inf == 0) return 0;
  if (!s->unconv_count) PGPU_CUDA(cudaMalloc(&s->unconv_count, 4 * sizeof(unsigned)));
  unsigned *nfail = s->unconv_count + 3;
  PGPU_CUDA(cudaMemsetAsync(nfail, 0, sizeof(unsigned), st));
  AdvanceParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.cnormDt = dt * s->desc.cvac_norm;
  prm.fnorm = s->desc.fnorm_const;
  prm.alpha = s->desc.fnorm_const * prm.cnormDt / 2.0;
  prm.rtol = s->desc.rtol;
  prm.iter_max = s->desc.iter_max;
  const GeoAny &g = s->grid->geo;
  prm.volume = (g.D == 1) ? g.dx[0] : g.dx[0] * g.dx[1];
  prm.rvolume = 1.0 / prm.volume;
  prm.ext = s->grid->ext;
  int rc = launch_suborbit(s, prm, from_emjacobian ? 1 : 0, s->Jinf, nfail, 1);
  if (rc) return rc;
  const double f = s->desc.charge / s->grid->desc.volume_scale;   // :3308-3315
  for (int c = 0; c < 3; ++c) {
    rc = scale_fab(s->Jinf[c], f);
    if (rc) return rc;
  }
  unsigned failed = 0;
  PGPU_CUDA(cudaMemcpyAsync(&failed, nfail, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  if (failed) {
    set_error("inflow sub-orbits: %u particles did not converge with 512 sub-orbits or left the ghosted arrays", failed);
    return PGPU_ERR_STATE;
  }
  return 0;
}
int pgpu_species_inflow_current_get(pgpu_species_t s, int comp, double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!s || comp < 0 || comp >= 3 || !s->Jinf[comp].p) return PGPU_ERR_ARG;
  int rc = copy_fab_to_host(s->Jinf[comp], s->grid->desc.D, data, lo, hi);
  if (rc) return rc;
  return pgpu_synchronize();
}
// PicSpeciesInterface::addInflowJ (PicSpeciesInterface.cpp:1499-1536): total J += this species' inflow J
int pgpu_current_add_inflow(pgpu_grid_t g, pgpu_species_t s) {
  NEED_INIT();
  if (!g || !s) return PGPU_ERR_ARG;
  if (s->desc.charge == 0.0 || !s->Jinf[0].p) return 0;
  for (int c = 0; c < 3; ++c) {
    const long n = (long)g->jtot[c].size();
    k_add_arr<<<nb(n), 256, 0, ctx().stream>>>(g->jtot[c].p, s->Jinf[c].p, n);
  }
  return 0;
}
// m_delta_{Mass,MomX,MomY,MomZ,Energy}In per boundary (2 dir + side), accumulated by the inflow part of pgpu_apply_bcs
// since the last call; reading resets them (PicChargedSpeciesBC::zeroDeltas)
int pgpu_species_inflow_fluxes(pgpu_species_t s, double *flux20) {
  if (!s || !flux20) return PGPU_ERR_ARG;
  for (int k = 0; k < 20; ++k) {
    flux20[k] = s->flux_in[k];
    s->flux_in[k] = 0.0;
  }
  return 0;
}

}  // extern "C"
