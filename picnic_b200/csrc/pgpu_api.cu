// pgpu_api.cu -- C ABI (include/picnic_gpu.h): lifecycle, grid and species state,
// the streaming per-particle passes, grid-side helpers of the deposit, reductions.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pgpu_internal.h"

#include <algorithm>

namespace pgpu {

static char g_err[512] = "";
static Context g_ctx;
Context &ctx() { return g_ctx; }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char *what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return PGPU_ERR_CUDA;
}

KTimer::KTimer(const char *n) : name(n) {
  Context &c = ctx();
  c.launches += 1;
  if (c.profile) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, c.stream);
  }
}
void KTimer::stop() {
  Context &c = ctx();
  if (a) {
    cudaEventRecord(b, c.stream);
    c.recs.push_back(Context::Rec{std::string(name), a, b});
    a = b = nullptr;
  }
}
KTimer::~KTimer() { stop(); }

static void profile_drain() {
  Context &c = ctx();
  for (auto &r : c.recs) {
    cudaEventSynchronize(r.b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    auto &e = c.prof[r.name];
    e.first += ms;
    e.second += 1;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  c.recs.clear();
}

GeoAny species_geo(const pgpu_species_s *s) {
  GeoAny g = s->grid->geo;
  for (int d = 0; d < 2; ++d) {
    g.bc_lo[d] = s->desc.bc_check_lo[d];
    g.bc_hi[d] = s->desc.bc_check_hi[d];
  }
  return g;
}
int fields_wait(const pgpu_grid_s *g) {
  if (g->upload_pending) {
    PGPU_CUDA(cudaStreamWaitEvent(ctx().stream, g->upload_done, 0));
    g->upload_pending = false;
  }
  return 0;
}

FieldSet grid_fields(const pgpu_grid_s *g) {
  fields_wait(g);
  FieldSet F;
  for (int c = 0; c < 6; ++c) F.f[c] = g->field[c].view();
  return F;
}
CurrentSet species_current(const pgpu_species_s *s) {
  CurrentSet J;
  for (int c = 0; c < 3; ++c) J.j[c] = s->J[c].view();
  return J;
}

// centring of the six field components: 1 = nodal
static void comp_stag(int D, int comp, int *stag) {
  static const int e1[3][2] = {{0, 0}, {1, 0}, {1, 0}};
  static const int b1[3][2] = {{1, 0}, {0, 0}, {0, 0}};
  static const int e2[3][2] = {{0, 1}, {1, 0}, {1, 1}};
  static const int b2[3][2] = {{1, 0}, {0, 1}, {0, 0}};
  const int(*t)[2] = (D == 1) ? (comp < 3 ? e1 : b1) : (comp < 3 ? e2 : b2);
  stag[0] = t[comp % 3][0];
  stag[1] = t[comp % 3][1];
}

static int alloc_fab(const pgpu_grid_desc &d, const int *stag, DeviceFab *f) {
  for (int k = 0; k < 2; ++k) {
    if (k < d.D) {
      f->lo[k] = d.box_lo[k] - d.nghost;
      f->hi[k] = d.box_hi[k] + d.nghost + stag[k];
    } else {
      f->lo[k] = f->hi[k] = 0;
    }
    f->stag[k] = (k < d.D) ? stag[k] : 0;
  }
  f->n0 = f->hi[0] - f->lo[0] + 1;
  f->n1 = f->hi[1] - f->lo[1] + 1;
  PGPU_CUDA(cudaMalloc(&f->p, f->size() * sizeof(double)));
  PGPU_CUDA(cudaMemsetAsync(f->p, 0, f->size() * sizeof(double), ctx().stream));
  return 0;
}

// n arrays of one centring set in ONE allocation, back to back: a packed host buffer in the same order moves with a
// single cudaMemcpyAsync (pgpu_fields_set_packed / pgpu_current_get_packed_async)
static int alloc_fab_arena(const pgpu_grid_desc &d, int first_comp, int n, DeviceFab *f) {
  size_t total = 0;
  for (int c = 0; c < n; ++c) {
    int stag[2];
    comp_stag(d.D, first_comp + c, stag);
    DeviceFab &o = f[c];
    for (int k = 0; k < 2; ++k) {
      if (k < d.D) {
        o.lo[k] = d.box_lo[k] - d.nghost;
        o.hi[k] = d.box_hi[k] + d.nghost + stag[k];
      } else {
        o.lo[k] = o.hi[k] = 0;
      }
      o.stag[k] = (k < d.D) ? stag[k] : 0;
    }
    o.n0 = o.hi[0] - o.lo[0] + 1;
    o.n1 = o.hi[1] - o.lo[1] + 1;
    total += o.size();
  }
  double *base = nullptr;
  PGPU_CUDA(cudaMalloc(&base, total * sizeof(double)));
  PGPU_CUDA(cudaMemsetAsync(base, 0, total * sizeof(double), ctx().stream));
  size_t off = 0;
  for (int c = 0; c < n; ++c) {
    f[c].p = base + off;
    off += f[c].size();
  }
  return 0;
}

// read back and clear the device counters (synchronises)
static int fetch_counters(Counters *out) {
  Context &c = ctx();
  PGPU_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c.stream));
  PGPU_CUDA(cudaMemsetAsync(c.d_counters, 0, sizeof(Counters), c.stream));
  PGPU_CUDA(cudaStreamSynchronize(c.stream));
  *out = *c.h_counters;
  c.total_apply_its += (long)out->apply_its;
  c.total_unconverged += (long)out->unconverged;
  return 0;
}
static int check_err_bits(unsigned err) {
  if (err & ERRBIT_SEGMENTS) {
    set_error("particle crossing more cells than allowed: num_segments > ghosts+1 "
              "(decrease the time step or increase grid.num_ghosts)");
    ctx().sticky_error = PGPU_ERR_SEGMENTS;
    return PGPU_ERR_SEGMENTS;
  }
  if (err & ERRBIT_BOUNDS) {
    set_error("a particle stencil left the ghosted field arrays");
    ctx().sticky_error = PGPU_ERR_BOUNDS;
    return PGPU_ERR_BOUNDS;
  }
  return 0;
}

// ---- small streaming kernels --------------------------------------------------------
// out = a*in1 + b*in2 evaluated as the reference writes it (no contraction):
// mode 0: out = in2 + in1*a          (xp = xpold + up*cnormDt)
// mode 1: out = 2*in1 - in2          (2nd-half updates)
// mode 2: out = (in1 + in2)/2        (averageVelocities)
__global__ void k_stream(double *out, const double *in1, const double *in2, long n, double a, int mode) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double u = in1[i], v = in2[i];
  double r;
  if (mode == 0) r = __dadd_rn(v, __dmul_rn(u, a));
  else if (mode == 1) r = __dsub_rn(__dmul_rn(2.0, u), v);
  else r = __ddiv_rn(__dadd_rn(u, v), 2.0);
  out[i] = r;
}

// advancePositionsExplicit / advancePositionsImplicit of the RELATIVISTIC_PARTICLES build
// (PicChargedSpecies.cpp:496-498, 548-553): xp = xpold + up/gammap*a, gammap of up (explicit) or getImplicitGamma
__global__ void k_positions_rel(PartPtrs p, long n, int D, double a, int implicit_gamma) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double u[3] = {p.v[0][i], p.v[1][i], p.v[2][i]};
  double gammap;
  if (implicit_gamma) {
    const double uo[3] = {p.vold[0][i], p.vold[1][i], p.vold[2][i]};
    gammap = gamma_implicit<true>(uo, u);
  } else {
    gammap = gamma_explicit<true>(u);
  }
  for (int d = 0; d < D; ++d) p.x[d][i] = __dadd_rn(p.xold[d][i], __dmul_rn(__ddiv_rn(u[d], gammap), a));
}

__global__ void k_boris_stored(PartPtrs p, long n, double alpha, int byHalfDt, int exact, int rel, int hc) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double uo[3], E[3], B[3], u[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uo[c] = p.vold[c][i];
    E[c] = p.Ep[c][i];
    B[c] = p.Bp[c][i];
  }
  if (exact) boris<true>(uo, E, B, alpha, byHalfDt != 0, u, rel, hc);
  else boris<false>(uo, E, B, alpha, byHalfDt != 0, u, rel, hc);
#pragma unroll
  for (int c = 0; c < 3; ++c) p.v[c][i] = u[c];
}

// The curvilinear velocity pushes of PicChargedSpecies::applyForces (PicChargedSpecies.cpp:341-355) on the stored particle
// fields: type 1 applyForces_CYL_CYL, 2 _SPH_SPH, 3 _CYL_HYB, 4 _SPH_HYB (PicSpeciesUtils.cpp:103-473).  A streaming pass:
// every operation is the reference's, individually rounded (no contraction), so the result differs from the reference's only
// through sin / cos (1-2 ulp against glibc).  r_old = x_old[0]; virt0 / virt1 = position_virt (dtheta, dphi).
__global__ void k_boris_curvilinear(PartPtrs p, const double *r_old, double *virt0, double *virt1, long n, int type,
                                    double alpha, double cnormDt, int byHalfDt, int anticyclic, int rel) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto A = [](double a, double b) { return __dadd_rn(a, b); };
  auto S = [](double a, double b) { return __dsub_rn(a, b); };
  auto M_ = [](double a, double b) { return __dmul_rn(a, b); };
  auto D_ = [](double a, double b) { return __ddiv_rn(a, b); };
  int dirp[3] = {0, 1, 2};
  if (anticyclic && (type == 1 || type == 3)) {
    dirp[1] = 2;
    dirp[2] = 1;
  }
  const double uo[3] = {p.vold[dirp[0]][i], p.vold[dirp[1]][i], p.vold[dirp[2]][i]};
  const double E[3] = {p.Ep[dirp[0]][i], p.Ep[dirp[1]][i], p.Ep[dirp[2]][i]};
  const double B[3] = {p.Bp[dirp[0]][i], p.Bp[dirp[1]][i], p.Bp[dirp[2]][i]};
  double vm0 = A(uo[0], M_(alpha, E[0])), vm1 = A(uo[1], M_(alpha, E[1])), vm2 = A(uo[2], M_(alpha, E[2]));
  double bp0, bp1, bp2, gammap = 1.0, up[3];
  const double hdt = D_(cnormDt, 2.0);
  auto rotate = [&]() {
    const double denom = A(A(A(1.0, M_(bp0, bp0)), M_(bp1, bp1)), M_(bp2, bp2));
    const double vpr0 = S(A(vm0, M_(vm1, bp2)), M_(vm2, bp1));
    const double vpr1 = S(A(vm1, M_(vm2, bp0)), M_(vm0, bp2));
    const double vpr2 = S(A(vm2, M_(vm0, bp1)), M_(vm1, bp0));
    up[0] = A(vm0, D_(S(M_(vpr1, bp2), M_(vpr2, bp1)), denom));
    up[1] = A(vm1, D_(S(M_(vpr2, bp0), M_(vpr0, bp2)), denom));
    up[2] = A(vm2, D_(S(M_(vpr0, bp1), M_(vpr1, bp0)), denom));
  };
  auto gamma_vm = [&]() { return __dsqrt_rn(A(A(A(1.0, M_(vm0, vm0)), M_(vm1, vm1)), M_(vm2, vm2))); };
  if (type == 1) {
    bp0 = M_(alpha, B[0]);
    bp1 = M_(alpha, B[1]);
    if (rel) {
      gammap = gamma_vm();
      bp0 = D_(bp0, gammap);
      bp1 = D_(bp1, gammap);
    }
    const double r = r_old[i];
    double dtheta = virt0[i];
    bool set_dtheta = false;
    if (dtheta == 0.0) {
      dtheta = D_(D_(M_(hdt, uo[1]), r), gammap);
      set_dtheta = true;
    }
    bp2 = A(D_(M_(alpha, B[2]), gammap), sin(dtheta));
    rotate();
    if (set_dtheta) {
      const double rpbar = A(r, M_(hdt, up[0]));
      virt0[i] = D_(D_(M_(hdt, up[1]), rpbar), gammap);
    }
    if (!byHalfDt)
      for (int k = 0; k < 3; ++k) up[k] = S(M_(2.0, up[k]), uo[k]);
  } else if (type == 2) {
    if (rel) gammap = gamma_vm();
    bp0 = D_(M_(alpha, B[0]), gammap);
    bp1 = D_(M_(alpha, B[1]), gammap);
    bp2 = D_(M_(alpha, B[2]), gammap);
    const double r = r_old[i];
    double dtheta = virt0[i], dphi = virt1[i];
    bool set_dtheta = false;
    if (dtheta == 0.0) {
      dtheta = sin(D_(D_(M_(hdt, uo[1]), r), gammap));
      dphi = sin(D_(D_(M_(hdt, uo[2]), r), gammap));
      set_dtheta = true;
    }
    bp0 = A(bp0, M_(sin(dphi), dtheta));
    bp1 = S(bp1, dphi);
    bp2 = A(bp2, M_(cos(dphi), dtheta));
    rotate();
    if (set_dtheta) {
      const double rpbar = A(r, M_(hdt, up[0]));
      dphi = sin(D_(D_(M_(hdt, up[2]), rpbar), gammap));
      dtheta = sin(D_(D_(D_(M_(hdt, up[1]), rpbar), gammap), cos(dphi)));
      virt0[i] = dtheta;
      virt1[i] = dphi;
    }
    if (!byHalfDt)
      for (int k = 0; k < 3; ++k) up[k] = S(M_(2.0, up[k]), uo[k]);
  } else {
    bp0 = M_(alpha, B[0]);
    bp1 = M_(alpha, B[1]);
    bp2 = M_(alpha, B[2]);
    if (rel) {
      gammap = gamma_vm();
      bp0 = D_(bp0, gammap);
      bp1 = D_(bp1, gammap);
      bp2 = D_(bp2, gammap);
    }
    const double thp = virt0[i];
    const double costhp = cos(thp), sinthp = sin(thp);
    if (type == 3) {
      vm0 = A(A(M_(costhp, uo[0]), M_(sinthp, uo[1])), M_(alpha, E[0]));
      vm1 = A(A(M_(-sinthp, uo[0]), M_(costhp, uo[1])), M_(alpha, E[1]));
    } else {
      const double php = virt1[i];
      const double cosphp = cos(php), sinphp = sin(php);
      const double h = A(M_(costhp, uo[0]), M_(sinthp, uo[1]));
      vm0 = A(A(M_(cosphp, h), M_(sinphp, uo[2])), M_(alpha, E[0]));
      vm1 = A(A(M_(-sinthp, uo[0]), M_(costhp, uo[1])), M_(alpha, E[1]));
      vm2 = A(A(M_(-sinphp, h), M_(cosphp, uo[2])), M_(alpha, E[2]));
    }
    rotate();
  }
  p.v[dirp[0]][i] = up[0];
  p.v[dirp[1]][i] = up[1];
  p.v[dirp[2]][i] = up[2];
}

__global__ void k_scale(double *a, long n, double s) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = __dmul_rn(a[i], s);
}
__global__ void k_add(double *a, const double *b, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = __dadd_rn(a[i], b[i]);
}

// the three components of a current in one launch (blockIdx.y = component): these grid kernels are a few
// microseconds each, so the launch count is what they cost
struct Ptr3 {
  double *a[3];
  const double *b[3];
  long n[3];
};
__global__ void k_scale3(Ptr3 P, double s) {
  const int c = blockIdx.y;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n[c]) P.a[c][i] = __dmul_rn(P.a[c][i], s);
}
__global__ void k_add3(Ptr3 P) {
  const int c = blockIdx.y;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < P.n[c]) P.a[c][i] = __dadd_rn(P.a[c][i], P.b[c][i]);
}
struct Fab3 {
  FabView f[3];
};

// Periodic ghost fold of one direction: pass 0 adds every non-owned entry onto its
// owned image (atomics: several images can map to one entry), pass 1 refreshes the
// images.  Owned index range in direction `dir` is own_lo..own_hi (cells; for nodal
// data node own_hi+1 is the image of node own_lo).
__global__ void k_fold(Fab3 F, int dir, int own_lo, int own_hi, int pass) {
  const FabView f = F.f[blockIdx.y];
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)f.n0 * f.n1;
  if (t >= total) return;
  const int a = (int)(t % f.n0), b = (int)(t / f.n0);
  const int idx = (dir == 0) ? a + f.lo0 : b + f.lo1;
  if (idx >= own_lo && idx <= own_hi) return;
  const int N = own_hi - own_lo + 1;
  int im = idx;
  while (im < own_lo) im += N;
  while (im > own_hi) im -= N;
  const long src = (dir == 0) ? (long)(im - f.lo0) + (long)b * f.n0 : (long)a + (long)(im - f.lo1) * f.n0;
  if (pass == 0) atomicAdd(f.p + src, f.p[t]);
  else f.p[t] = f.p[src];
}

static inline unsigned nb(long n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

int scale_fab(const DeviceFab &f, double s) {
  KTimer t("fab_scale");
  k_scale<<<nb((long)f.size()), 256, 0, ctx().stream>>>(f.p, (long)f.size(), s);
  return 0;
}

// nf arrays of one grid (same periodic directions; sizes may differ by the centring)
static int fold_periodic_n(const pgpu_grid_s *g, const DeviceFab *const *f, int nf) {
  Context &c = ctx();
  Fab3 F;
  long total = 0;
  for (int k = 0; k < nf; ++k) {
    F.f[k] = f[k]->view();
    total = std::max(total, (long)f[k]->size());
  }
  for (int k = nf; k < 3; ++k) F.f[k] = F.f[0];
  for (int dir = 0; dir < g->desc.D; ++dir) {
    if (!g->desc.periodic[dir]) continue;
    // only a box spanning the whole periodic direction folds onto itself
    if (g->desc.box_lo[dir] != 0 || g->desc.box_hi[dir] != g->desc.ncell[dir] - 1) continue;
    for (int pass = 0; pass < 2; ++pass) {
      KTimer t("fold_periodic");
      k_fold<<<dim3(nb(total), nf), 256, 0, c.stream>>>(F, dir, g->desc.box_lo[dir], g->desc.box_hi[dir], pass);
    }
  }
  return 0;
}
// SpaceUtils::applyBinomialFilter(FArrayBox&, const Box&) (SpaceUtils.cpp:54-112): Q2 = the [1 2 1] (x) [1 2 1] sum of
// the neighbours over the box's own edges / nodes, then Q = (2^D Q + Q2) / 4^D -- in that operation order.  Two
// passes (Q2 from the unmodified array, then the update); entries outside the grid box keep their values.
__global__ void k_binomial_q2(FabView f, double *q2, int D, int lo0, int hi0, int lo1, int hi1) {
  const int m0 = hi0 - lo0 + 1, m1 = hi1 - lo1 + 1;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)m0 * m1) return;
  const int i = lo0 + (int)(t % m0), j = lo1 + (int)(t / m0);
  const double *p = f.p + (i - f.lo0) + (long)(j - f.lo1) * f.n0;
  if (D == 1) {
    q2[t] = __dadd_rn(p[1], p[-1]);
    return;
  }
  const long n0 = f.n0;
  double a = __dmul_rn(2.0, __dadd_rn(p[1], p[-1]));
  a = __dadd_rn(a, __dmul_rn(2.0, __dadd_rn(p[n0], p[-n0])));
  a = __dadd_rn(a, p[1 + n0]);
  a = __dadd_rn(a, p[1 - n0]);
  a = __dadd_rn(a, p[-1 + n0]);
  a = __dadd_rn(a, p[-1 - n0]);
  q2[t] = a;
}
__global__ void k_binomial_apply(FabView f, const double *q2, int D, int lo0, int hi0, int lo1, int hi1) {
  const int m0 = hi0 - lo0 + 1, m1 = hi1 - lo1 + 1;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)m0 * m1) return;
  const int i = lo0 + (int)(t % m0), j = lo1 + (int)(t / m0);
  double *p = f.p + (i - f.lo0) + (long)(j - f.lo1) * f.n0;
  const double f0 = D == 1 ? 2.0 : 4.0, f1 = D == 1 ? 4.0 : 16.0;
  *p = __ddiv_rn(__dadd_rn(__dmul_rn(*p, f0), q2[t]), f1);
}

int binomial_filter(pgpu_grid_s *g, const DeviceFab &f) {
  Context &c = ctx();
  const int D = g->desc.D;
  if (g->desc.nghost < 1) {
    set_error("the binomial filter needs one ghost layer");
    return PGPU_ERR_STATE;
  }
  int lo[2] = {0, 0}, hi[2] = {0, 0};
  for (int k = 0; k < D; ++k) lo[k] = g->desc.box_lo[k], hi[k] = g->desc.box_hi[k] + f.stag[k];
  const long n = (long)(hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1);
  if (g->filter_cap < (size_t)n) {
    if (g->filter_tmp) cudaFree(g->filter_tmp);
    g->filter_tmp = nullptr;
    PGPU_CUDA(cudaMalloc(&g->filter_tmp, n * sizeof(double)));
    g->filter_cap = n;
  }
  {
    KTimer t("binomial_filter");
    k_binomial_q2<<<nb(n), 256, 0, c.stream>>>(f.view(), g->filter_tmp, D, lo[0], hi[0], lo[1], hi[1]);
    k_binomial_apply<<<nb(n), 256, 0, c.stream>>>(f.view(), g->filter_tmp, D, lo[0], hi[0], lo[1], hi[1]);
  }
  // keep the periodic images of a self-periodic direction equal to their owners
  Fab3 F;
  F.f[0] = F.f[1] = F.f[2] = f.view();
  for (int dir = 0; dir < D; ++dir) {
    if (!g->desc.periodic[dir]) continue;
    if (g->desc.box_lo[dir] != 0 || g->desc.box_hi[dir] != g->desc.ncell[dir] - 1) continue;
    k_fold<<<dim3(nb((long)f.size()), 1), 256, 0, c.stream>>>(F, dir, g->desc.box_lo[dir], g->desc.box_hi[dir], 1);
  }
  return 0;
}

int grid_rho_fab(pgpu_grid_s *g, const int *stag, DeviceFab **out) {
  const int D = g->desc.D;
  const int st2[2] = {stag[0] ? 1 : 0, (D == 2 && stag[1]) ? 1 : 0};
  DeviceFab &f = g->rho[st2[0] + 2 * st2[1]];
  if (!f.p) {
    for (int k = 0; k < 2; ++k) {
      if (k < D) {
        f.lo[k] = g->desc.box_lo[k] - g->desc.nghost;
        f.hi[k] = g->desc.box_hi[k] + g->desc.nghost + st2[k];
      }
      f.stag[k] = st2[k];
    }
    f.n0 = f.hi[0] - f.lo[0] + 1;
    f.n1 = f.hi[1] - f.lo[1] + 1;
    PGPU_CUDA(cudaMalloc(&f.p, f.size() * sizeof(double)));
    PGPU_CUDA(cudaMemsetAsync(f.p, 0, f.size() * sizeof(double), ctx().stream));
  }
  *out = &f;
  return 0;
}

int fold_periodic(const pgpu_grid_s *g, const DeviceFab &f) {
  const DeviceFab *one[1] = {&f};
  return fold_periodic_n(g, one, 1);
}

int copy_fab_to_host(const DeviceFab &f, int D, double *data, const int *lo, const int *hi, bool sync) {
  for (int d = 0; d < D; ++d)
    if (lo[d] != f.lo[d] || hi[d] != f.hi[d]) {
      set_error("array bounds [%d:%d] in dir %d do not match the device box [%d:%d]", lo[d], hi[d], d,
                f.lo[d], f.hi[d]);
      return PGPU_ERR_ARG;
    }
  Context &c = ctx();
  PGPU_CUDA(cudaMemcpyAsync(data, f.p, f.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  if (sync) PGPU_CUDA(cudaStreamSynchronize(c.stream));
  return 0;
}

static int ensure_capacity(pgpu_species_s *s, long n) {
  if ((size_t)n <= s->cap) return 0;
  const size_t cap = (size_t)(n + n / 16 + 1024);
  const int D = s->grid->desc.D;
  auto re = [&](double *&p) -> int {
    if (p) cudaFree(p);
    p = nullptr;
    PGPU_CUDA(cudaMalloc(&p, cap * sizeof(double)));
    return 0;
  };
  for (int d = 0; d < D; ++d) {
    if (re(s->x[d])) return PGPU_ERR_CUDA;
    if (re(s->xold[d])) return PGPU_ERR_CUDA;
  }
  for (int c = 0; c < 3; ++c) {
    if (re(s->v[c])) return PGPU_ERR_CUDA;
    if (re(s->vold[c])) return PGPU_ERR_CUDA;
  }
  if (re(s->w)) return PGPU_ERR_CUDA;
  if (re(s->tmp)) return PGPU_ERR_CUDA;
  if (s->id) cudaFree(s->id);
  PGPU_CUDA(cudaMalloc(&s->id, cap * sizeof(uint64_t)));
  for (int c = 0; c < 3; ++c) {  // Ep/Bp are allocated lazily
    if (s->Ep[c]) cudaFree(s->Ep[c]);
    if (s->Bp[c]) cudaFree(s->Bp[c]);
    s->Ep[c] = s->Bp[c] = nullptr;
  }
  if (s->cell_key) cudaFree(s->cell_key);
  if (s->perm) cudaFree(s->perm);
  PGPU_CUDA(cudaMalloc(&s->cell_key, cap * sizeof(int)));
  PGPU_CUDA(cudaMalloc(&s->perm, cap * sizeof(int)));
  s->cap = cap;
  return 0;
}

// Capacity for n particles, keeping the first s->n of every array (arrivals of a migration).
int grow_capacity(pgpu_species_s *s, long n) {
  if ((size_t)n <= s->cap) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  const size_t cap = (size_t)(n + n / 16 + 1024);
  const int D = s->grid->desc.D;
  cudaStream_t st = ctx().stream;
  auto re = [&](double *&p) -> int {
    double *q = nullptr;
    PGPU_CUDA(cudaMalloc(&q, cap * sizeof(double)));
    if (p) {
      PGPU_CUDA(cudaMemcpyAsync(q, p, s->n * sizeof(double), cudaMemcpyDeviceToDevice, st));
      PGPU_CUDA(cudaStreamSynchronize(st));
      cudaFree(p);
    }
    p = q;
    return 0;
  };
  for (int d = 0; d < D; ++d) {
    if (re(s->x[d])) return PGPU_ERR_CUDA;
    if (re(s->xold[d])) return PGPU_ERR_CUDA;
  }
  for (int c = 0; c < 3; ++c) {
    if (re(s->v[c])) return PGPU_ERR_CUDA;
    if (re(s->vold[c])) return PGPU_ERR_CUDA;
  }
  if (re(s->w)) return PGPU_ERR_CUDA;
  {
    double *idp = reinterpret_cast<double *>(s->id);
    if (re(idp)) return PGPU_ERR_CUDA;
    s->id = reinterpret_cast<uint64_t *>(idp);
  }
  if (s->tmp) cudaFree(s->tmp);
  PGPU_CUDA(cudaMalloc(&s->tmp, cap * sizeof(double)));
  for (int c = 0; c < 3; ++c) {
    if (s->Ep[c]) cudaFree(s->Ep[c]);
    if (s->Bp[c]) cudaFree(s->Bp[c]);
    s->Ep[c] = s->Bp[c] = nullptr;
  }
  if (s->cell_key) cudaFree(s->cell_key);
  if (s->perm) cudaFree(s->perm);
  PGPU_CUDA(cudaMalloc(&s->cell_key, cap * sizeof(int)));
  PGPU_CUDA(cudaMalloc(&s->perm, cap * sizeof(int)));
  s->cap = cap;
  s->binned = false;
  return 0;
}

static int ensure_epbp(pgpu_species_s *s) {
  for (int c = 0; c < 3; ++c) {
    if (!s->Ep[c]) PGPU_CUDA(cudaMalloc(&s->Ep[c], s->cap * sizeof(double)));
    if (!s->Bp[c]) PGPU_CUDA(cudaMalloc(&s->Bp[c], s->cap * sizeof(double)));
  }
  return 0;
}

}  // namespace pgpu

using namespace pgpu;

#define NEED_INIT()                                        \
  do {                                                     \
    if (!ctx().inited) {                                   \
      set_error("pgpu_init has not been called");          \
      return PGPU_ERR_STATE;                               \
    }                                                      \
  } while (0)

extern "C" {

const char *pgpu_last_error(void) { return g_err; }
int pgpu_abi_version(void) { return 1; }

int pgpu_init(int device) {
  Context &c = ctx();
  if (c.inited) return 0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return PGPU_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    set_error("device %d out of range (0..%d)", device, ndev - 1);
    return PGPU_ERR_ARG;
  }
  PGPU_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PGPU_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return PGPU_ERR_CUDA;
  }
  c.sm_count = prop.multiProcessorCount;
  c.device = device;
  PGPU_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  c.own_stream = true;
  PGPU_CUDA(cudaMalloc(&c.d_counters, sizeof(Counters)));
  PGPU_CUDA(cudaMemset(c.d_counters, 0, sizeof(Counters)));
  PGPU_CUDA(cudaMallocHost(&c.h_counters, sizeof(Counters)));
  c.inited = true;
  c.sticky_error = 0;
  if (const char *e = getenv("PGPU_CC1_TMA")) c.cc1_tma = atoi(e);
  if (const char *e = getenv("PGPU_CC1_MINB")) c.cc1_minblocks = atoi(e);
  if (const char *e = getenv("PGPU_CC1_PREFETCH")) c.cc1_prefetch = atoi(e);
  if (const char *e = getenv("PGPU_CC1_PAIR")) c.cc1_pair = atoi(e);
  if (const char *e = getenv("PGPU_CC1_RSTEPS")) c.cc1_rsteps = atoi(e);
  if (const char *e = getenv("PGPU_COPY_STREAM")) c.use_copy_stream = atoi(e);
  if (const char *e = getenv("PGPU_CC1_V")) c.cc1_version = atoi(e);
  if (const char *e = getenv("PGPU_CC1_NODECACHE")) c.cc1_nodecache = atoi(e);
  if (const char *e = getenv("PGPU_CC1_MULTISEG")) c.cc1_multiseg = atoi(e);
  if (const char *e = getenv("PGPU_TA_STAGED")) c.ta_staged = atoi(e);
  if (const char *e = getenv("PGPU_CC1_REC")) c.cc1_rec_per_pass = atoi(e);
  if (const char *e = getenv("PGPU_CC1_WAVES")) c.cc1_waves = atoi(e) > 0 ? atoi(e) : 1;
  return 0;
}

int pgpu_finalize(void) {
  Context &c = ctx();
  if (!c.inited) return 0;
  if (c.copy_stream) {
    cudaStreamSynchronize(c.copy_stream);
    cudaStreamDestroy(c.copy_stream);
    cudaEventDestroy(c.copy_fence);
  }
  cudaStreamSynchronize(c.stream);
  profile_drain();
  if (c.own_stream) cudaStreamDestroy(c.stream);
  cudaFree(c.d_counters);
  if (c.ta_list) cudaFree(c.ta_list);
  cudaFreeHost(c.h_counters);
  c = Context();
  return 0;
}

int pgpu_set_stream(void *cuda_stream) {
  NEED_INIT();
  Context &c = ctx();
  cudaStreamSynchronize(c.stream);
  if (cuda_stream == nullptr) {
    if (!c.own_stream) {
      PGPU_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
      c.own_stream = true;
    }
  } else {
    if (c.own_stream) cudaStreamDestroy(c.stream);
    c.stream = (cudaStream_t)cuda_stream;
    c.own_stream = false;
  }
  return 0;
}

int pgpu_synchronize(void) {
  NEED_INIT();
  Context &c = ctx();
  if (c.copy_stream) PGPU_CUDA(cudaStreamSynchronize(c.copy_stream));
  Counters k;
  if (fetch_counters(&k)) return PGPU_ERR_CUDA;
  int rc = check_err_bits(k.err);
  if (rc) return rc;
  PGPU_CUDA(cudaGetLastError());
  return 0;
}

int pgpu_set_exact_math(int on) {
  ctx().exact = on != 0;
  return 0;
}
int pgpu_set_deposit_mode(int mode) {
  if (mode < 0 || mode > 1) return PGPU_ERR_ARG;
  ctx().deposit_mode = mode;
  ctx().use_fast_cc1 = mode == 1;
  return 0;
}

// ---- grid -------------------------------------------------------------------------
int pgpu_grid_create(const pgpu_grid_desc *d, pgpu_grid_t *out) {
  NEED_INIT();
  if (!d || !out || (d->D != 1 && d->D != 2) || d->nghost < 1) {
    set_error("bad grid descriptor");
    return PGPU_ERR_ARG;
  }
  for (int k = 0; k < d->D; ++k)
    if (d->dx[k] <= 0 || d->box_hi[k] < d->box_lo[k] || d->box_lo[k] < 0 || d->box_hi[k] >= d->ncell[k]) {
      set_error("bad grid descriptor in direction %d", k);
      return PGPU_ERR_ARG;
    }
  pgpu_grid_s *g = new pgpu_grid_s();
  g->desc = *d;
  g->geo.D = d->D;
  for (int k = 0; k < 2; ++k) {
    const bool on = k < d->D;
    g->geo.le[k] = on ? d->xmin[k] : 0.0;
    g->geo.dx[k] = on ? d->dx[k] : 1.0;
    g->geo.rdx[k] = 1.0 / g->geo.dx[k];
    // DomainGrid Xmax = Xmin + ncell*dX in exact arithmetic; decks give X_max directly
    g->geo.re[k] = on ? d->xmin[k] + d->ncell[k] * d->dx[k] : 1.0;
    g->geo.bc_lo[k] = g->geo.bc_hi[k] = 0;
    g->nbox[k] = on ? d->box_hi[k] - d->box_lo[k] + 1 : 1;
  }
  g->geo.ghosts = d->nghost;
  g->ncell_box = (long)g->nbox[0] * g->nbox[1];
  if (alloc_fab_arena(*d, 0, 6, g->field)) return PGPU_ERR_CUDA;
  for (int c = 0; c < 6; ++c) g->field_slot[0][c] = g->field[c];
  if (alloc_fab_arena(*d, 0, 3, g->jtot)) return PGPU_ERR_CUDA;   // J has the centring of E
  PGPU_CUDA(cudaMalloc(&g->debye, g->ncell_box * sizeof(double)));
  *out = g;
  return 0;
}

int pgpu_grid_destroy(pgpu_grid_t g) {
  if (!g) return 0;
  if (ctx().copy_stream) cudaStreamSynchronize(ctx().copy_stream);
  cudaStreamSynchronize(ctx().stream);
  if (g->upload_done) cudaEventDestroy(g->upload_done);
  for (int k = 0; k < 4; ++k)
    if (g->field_slot[k][0].p) cudaFree(g->field_slot[k][0].p);   // one arena per slot (alloc_fab_arena)
  cudaFree(g->jtot[0].p);
  for (int k = 0; k < 4; ++k) {
    if (g->tab_dual[k]) cudaFree(g->tab_dual[k]);
    if (g->tab_node[k]) cudaFree(g->tab_node[k]);
  }
  if (g->scratch_rho.p) cudaFree(g->scratch_rho.p);
  for (auto &f : g->rho)
    if (f.p) cudaFree(f.p);
  if (g->filter_tmp) cudaFree(g->filter_tmp);
  cudaFree(g->debye);
  mm_destroy(g);
  delete g;
  return 0;
}

int pgpu_field_bounds(pgpu_grid_t g, int comp, int *lo, int *hi) {
  if (!g || comp < 0 || comp >= 6) return PGPU_ERR_ARG;
  for (int d = 0; d < 2; ++d) {
    lo[d] = g->field[comp].lo[d];
    hi[d] = g->field[comp].hi[d];
  }
  return 0;
}

int pgpu_fields_set(pgpu_grid_t g, int comp, const double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!g || comp < 0 || comp >= 6 || !data) return PGPU_ERR_ARG;
  const DeviceFab &f = g->field[comp];
  for (int d = 0; d < g->desc.D; ++d)
    if (lo[d] != f.lo[d] || hi[d] != f.hi[d]) {
      set_error("field %d: bounds [%d:%d] in dir %d do not match the device box [%d:%d]", comp, lo[d], hi[d],
                d, f.lo[d], f.hi[d]);
      return PGPU_ERR_ARG;
    }
  if (fields_wait(g)) return PGPU_ERR_CUDA;
  PGPU_CUDA(cudaMemcpyAsync(f.p, data, f.size() * sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
  g->tab_dirty[g->cur_slot] = true;
  return 0;
}

int pgpu_fields_select(pgpu_grid_t g, int slot) {
  NEED_INIT();
  if (!g || slot < 0 || slot >= 4) return PGPU_ERR_ARG;
  if (!g->field_slot[slot][0].p && alloc_fab_arena(g->desc, 0, 6, g->field_slot[slot])) return PGPU_ERR_CUDA;
  for (int c = 0; c < 6; ++c) g->field[c] = g->field_slot[slot][c];
  g->cur_slot = slot;
  return 0;
}

// The six field components in one host buffer, back to back in the order Ex Ey Ez Bx By Bz, each in the layout
// pgpu_fields_set takes (ghosted box of the component, column major): one H2D copy per preRHSOp instead of six.
int pgpu_fields_packed_size(pgpu_grid_t g, long *ndoubles) {
  if (!g || !ndoubles) return PGPU_ERR_ARG;
  long n = 0;
  for (int c = 0; c < 6; ++c) n += (long)g->field[c].size();
  *ndoubles = n;
  return 0;
}
int pgpu_fields_set_packed(pgpu_grid_t g, const double *data) {
  NEED_INIT();
  if (!g || !data) return PGPU_ERR_ARG;
  long n = 0;
  for (int c = 0; c < 6; ++c) n += (long)g->field[c].size();
  Context &c = ctx();
  if (c.use_copy_stream) {
    // the copy is ordered behind everything already enqueued on the library stream (earlier readers of this slot) but
    // not behind what is enqueued after this call: box B's fields move while box A's particles run
    if (!c.copy_stream) {
      PGPU_CUDA(cudaStreamCreateWithFlags(&c.copy_stream, cudaStreamNonBlocking));
      PGPU_CUDA(cudaEventCreateWithFlags(&c.copy_fence, cudaEventDisableTiming));
    }
    if (!g->upload_done) PGPU_CUDA(cudaEventCreateWithFlags(&g->upload_done, cudaEventDisableTiming));
    PGPU_CUDA(cudaEventRecord(c.copy_fence, c.stream));
    PGPU_CUDA(cudaStreamWaitEvent(c.copy_stream, c.copy_fence, 0));
    PGPU_CUDA(cudaMemcpyAsync(g->field[0].p, data, n * sizeof(double), cudaMemcpyHostToDevice, c.copy_stream));
    PGPU_CUDA(cudaEventRecord(g->upload_done, c.copy_stream));
    g->upload_pending = true;
  } else {
    PGPU_CUDA(cudaMemcpyAsync(g->field[0].p, data, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  }
  g->tab_dirty[g->cur_slot] = true;
  return 0;
}
// Jx Jy Jz of pgpu_current_finalize likewise in one buffer; asynchronous: the data is valid after pgpu_synchronize
int pgpu_current_packed_size(pgpu_grid_t g, long *ndoubles) {
  if (!g || !ndoubles) return PGPU_ERR_ARG;
  long n = 0;
  for (int c = 0; c < 3; ++c) n += (long)g->jtot[c].size();
  *ndoubles = n;
  return 0;
}
int pgpu_current_get_packed_async(pgpu_grid_t g, double *data) {
  NEED_INIT();
  if (!g || !data) return PGPU_ERR_ARG;
  long n = 0;
  for (int c = 0; c < 3; ++c) n += (long)g->jtot[c].size();
  PGPU_CUDA(cudaMemcpyAsync(data, g->jtot[0].p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
  return 0;
}

int pgpu_host_register(void *ptr, size_t bytes) {
  NEED_INIT();
  PGPU_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault));
  return 0;
}
int pgpu_host_unregister(void *ptr) {
  NEED_INIT();
  PGPU_CUDA(cudaHostUnregister(ptr));
  return 0;
}

int pgpu_current_zero(pgpu_grid_t g) {
  NEED_INIT();
  if (!g) return PGPU_ERR_ARG;
  size_t n = 0;
  for (int c = 0; c < 3; ++c) n += g->jtot[c].size();
  PGPU_CUDA(cudaMemsetAsync(g->jtot[0].p, 0, n * sizeof(double), ctx().stream));   // one arena (alloc_fab_arena)
  return 0;
}

int pgpu_current_add_species(pgpu_grid_t g, pgpu_species_t s) {
  NEED_INIT();
  if (!g || !s) return PGPU_ERR_ARG;
  Ptr3 P;
  long mx = 0;
  for (int c = 0; c < 3; ++c) {
    P.a[c] = g->jtot[c].p;
    P.b[c] = s->J[c].p;
    P.n[c] = (long)g->jtot[c].size();
    mx = std::max(mx, P.n[c]);
  }
  KTimer t("current_add");
  k_add3<<<dim3(nb(mx), 3), 256, 0, ctx().stream>>>(P);
  return 0;
}

int pgpu_current_finalize(pgpu_grid_t g) {
  NEED_INIT();
  if (!g) return PGPU_ERR_ARG;
  const DeviceFab *all[3] = {&g->jtot[0], &g->jtot[1], &g->jtot[2]};
  return fold_periodic_n(g, all, 3);
}

int pgpu_current_filter(pgpu_grid_t g, int in_plane, int virtual_comps) {
  NEED_INIT();
  if (!g) return PGPU_ERR_ARG;
  const int D = g->desc.D;
  for (int c = 0; c < 3; ++c) {
    const bool virt = c >= D;
    if (virt ? !virtual_comps : !in_plane) continue;
    int rc = binomial_filter(g, g->jtot[c]);
    if (rc) return rc;
  }
  return 0;
}

int pgpu_charge_density_filter(pgpu_grid_t g, const int *stag) {
  NEED_INIT();
  if (!g || !stag) return PGPU_ERR_ARG;
  DeviceFab *f = nullptr;
  int rc = grid_rho_fab(g, stag, &f);
  if (rc) return rc;
  return binomial_filter(g, *f);
}

int pgpu_current_get(pgpu_grid_t g, int comp, double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!g || comp < 0 || comp >= 3) return PGPU_ERR_ARG;
  return copy_fab_to_host(g->jtot[comp], g->desc.D, data, lo, hi);
}

int pgpu_current_get_async(pgpu_grid_t g, int comp, double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!g || comp < 0 || comp >= 3) return PGPU_ERR_ARG;
  return copy_fab_to_host(g->jtot[comp], g->desc.D, data, lo, hi, false);
}

// ---- species ----------------------------------------------------------------------
int pgpu_species_create(pgpu_grid_t g, const pgpu_species_desc *d, pgpu_species_t *out) {
  NEED_INIT();
  if (!g || !d || !out) return PGPU_ERR_ARG;
  if (d->interp_E < 0 || d->interp_E > 3 || d->interp_J < 0 || d->interp_J > 3 || d->interp_N < 0 ||
      d->interp_N > 1) {
    set_error("bad interpolation type (interp_N must be CIC or TSC)");
    return PGPU_ERR_ARG;
  }
  pgpu_species_s *s = new pgpu_species_s();
  static unsigned next_serial = 1;
  s->serial = next_serial++;
  s->grid = g;
  s->desc = *d;
  for (int c = 0; c < 3; ++c) {
    int stag[2];
    comp_stag(g->desc.D, c, stag);
    if (alloc_fab(g->desc, stag, &s->J[c])) return PGPU_ERR_CUDA;
  }
  PGPU_CUDA(cudaMalloc(&s->cell_count, (g->ncell_box + 2) * sizeof(int)));
  PGPU_CUDA(cudaMalloc(&s->cell_start, (g->ncell_box + 2) * sizeof(int)));
  PGPU_CUDA(cudaMalloc(&s->dens, g->ncell_box * sizeof(double)));
  PGPU_CUDA(cudaMalloc(&s->mom, 3 * g->ncell_box * sizeof(double)));
  PGPU_CUDA(cudaMalloc(&s->ene, 3 * g->ncell_box * sizeof(double)));
  *out = s;
  return 0;
}

int pgpu_species_set_solver_params(pgpu_species_t s, int order_swap, int iter_max, double rtol) {
  if (!s || iter_max < 0 || !(rtol > 0.0)) return PGPU_ERR_ARG;
  s->desc.order_swap = order_swap ? 1 : 0;
  s->desc.iter_max = iter_max;
  s->desc.rtol = rtol;
  return 0;
}

int pgpu_species_destroy(pgpu_species_t s) {
  if (!s) return 0;
  cudaStreamSynchronize(ctx().stream);
  for (int d = 0; d < 2; ++d) {
    cudaFree(s->x[d]);
    cudaFree(s->xold[d]);
  }
  for (int c = 0; c < 3; ++c) {
    cudaFree(s->v[c]);
    cudaFree(s->vold[c]);
    cudaFree(s->Ep[c]);
    cudaFree(s->Bp[c]);
    cudaFree(s->J[c].p);
  }
  cudaFree(s->w);
  cudaFree(s->id);
  cudaFree(s->tmp);
  cudaFree(s->cell_key);
  cudaFree(s->key_sorted);
  cudaFree(s->old_perm);
  cudaFree(s->cub_tmp);
  for (int k = 0; k < 4; ++k) cudaFree(s->spare[k]);
  cudaFree(s->defer_list);
  cudaFree(s->defer_list2);
  cudaFree(s->defer_count2);
  cudaFree(s->bin_count);
  cudaFree(s->enf_save);
  for (int k = 0; k < 10; ++k) cudaFree(s->out[k]);
  cudaFree(s->out_w);
  cudaFree(s->out_id);
  cudaFree(s->out_tag);
  cudaFree(s->out_listtag);
  for (int k = 0; k < 2; ++k) cudaFree(s->virt[k]);
  for (int k = 0; k < 10; ++k) cudaFree(s->inf[k]);
  cudaFree(s->inf_w);
  cudaFree(s->inf_id);
  cudaFree(s->inf_code);
  for (int c = 0; c < 3; ++c) cudaFree(s->Jinf[c].p);
  for (int k = 0; k < 10; ++k) cudaFree(s->sub[k]);
  cudaFree(s->sub_w);
  cudaFree(s->sub_id);
  cudaFree(s->sub_nsub);
  for (int c = 0; c < 3; ++c) cudaFree(s->Jsub[c].p);
  cudaFree(s->unconv_list);
  cudaFree(s->unconv_count);
  cudaFree(s->tile_box);
  cudaFree(s->mig);
  cudaFree(s->mig_list);
  cudaFree(s->defer_count);
  cudaFree(s->perm);
  cudaFree(s->cell_count);
  cudaFree(s->cell_start);
  cudaFree(s->dens);
  cudaFree(s->mom);
  cudaFree(s->ene);
  delete s;
  return 0;
}

long pgpu_species_count(pgpu_species_t s) { return s ? s->n : -1; }

static __global__ void k_iota_ids(uint64_t *id, long n, uint64_t base) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) id[i] = base + (uint64_t)i;
}

int pgpu_species_upload(pgpu_species_t s, long n, const double *x, const double *xold, const double *v,
                        const double *vold, const double *w, const uint64_t *id) {
  NEED_INIT();
  if (!s || n < 0 || (n > 0 && (!x || !v || !w))) return PGPU_ERR_ARG;
  if (ensure_capacity(s, n)) return PGPU_ERR_CUDA;
  const int D = s->grid->desc.D;
  cudaStream_t st = ctx().stream;
  const size_t nb8 = (size_t)n * sizeof(double);
  for (int d = 0; d < D; ++d) {
    PGPU_CUDA(cudaMemcpyAsync(s->x[d], x + (size_t)d * n, nb8, cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaMemcpyAsync(s->xold[d], (xold ? xold : x) + (size_t)d * n, nb8, cudaMemcpyHostToDevice, st));
  }
  for (int c = 0; c < 3; ++c) {
    PGPU_CUDA(cudaMemcpyAsync(s->v[c], v + (size_t)c * n, nb8, cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaMemcpyAsync(s->vold[c], (vold ? vold : v) + (size_t)c * n, nb8, cudaMemcpyHostToDevice, st));
  }
  PGPU_CUDA(cudaMemcpyAsync(s->w, w, nb8, cudaMemcpyHostToDevice, st));
  if (id) PGPU_CUDA(cudaMemcpyAsync(s->id, id, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  else {
    // no ids given: unique ones are made up (serial of the species in the high bits, index below).  The collision
    // kernels key their per-particle draws (shuffle order, weight rejection, partner pick) on the id: identical ids
    // would give every particle of a cell identical draws
    if (n > 0) k_iota_ids<<<nb(n), 256, 0, st>>>(s->id, n, (uint64_t)s->serial << 40);
  }
  PGPU_CUDA(cudaStreamSynchronize(st));
  s->n = n;
  s->binned = false;
  s->pos_old_pending = s->vel_old_pending = false;
  s->xold_alias = s->vold_alias = false;
  return 0;
}

int pgpu_species_download(pgpu_species_t s, double *x, double *xold, double *v, double *vold, double *w,
                          uint64_t *id) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  const int D = s->grid->desc.D;
  const long n = s->n;
  cudaStream_t st = ctx().stream;
  const size_t nb8 = (size_t)n * sizeof(double);
  for (int d = 0; d < D; ++d) {
    if (x) PGPU_CUDA(cudaMemcpyAsync(x + (size_t)d * n, s->x[d], nb8, cudaMemcpyDeviceToHost, st));
    if (xold) PGPU_CUDA(cudaMemcpyAsync(xold + (size_t)d * n, s->xold[d], nb8, cudaMemcpyDeviceToHost, st));
  }
  for (int c = 0; c < 3; ++c) {
    if (v) PGPU_CUDA(cudaMemcpyAsync(v + (size_t)c * n, s->v[c], nb8, cudaMemcpyDeviceToHost, st));
    if (vold) PGPU_CUDA(cudaMemcpyAsync(vold + (size_t)c * n, s->vold[c], nb8, cudaMemcpyDeviceToHost, st));
  }
  if (w) PGPU_CUDA(cudaMemcpyAsync(w, s->w, nb8, cudaMemcpyDeviceToHost, st));
  if (id) PGPU_CUDA(cudaMemcpyAsync(id, s->id, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- List<JustinsParticle> adapter: the particle object's own linear record ---------------------------------------------
// JustinsParticle::linearOut / linearIn (src/particle_tools/JustinsParticle.cpp:339-378, 410-450) for CH_SPACEDIM = D < 3:
//   [ w | x[D] | x_old[D] | pos_virt[2] | v[3] | v_old[3] | (Real) ID ]   = 2 D + 10 doubles per particle
// (pos_virt: out-of-plane coordinates of the curvilinear pushes; zero in the planar path this library covers).  The
// transpose runs on the device, the record array crosses the bus in one copy.
static __global__ void k_linear_out(PartPtrs p, const uint64_t *id, long n, int D, double *rec) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double *r = rec + (size_t)i * (2 * D + 10);
  *r++ = p.w[i];
  for (int d = 0; d < D; ++d) *r++ = p.x[d][i];
  for (int d = 0; d < D; ++d) *r++ = p.xold[d][i];
  *r++ = 0.0;
  *r++ = 0.0;
  for (int c = 0; c < 3; ++c) *r++ = p.v[c][i];
  for (int c = 0; c < 3; ++c) *r++ = p.vold[c][i];
  *r = (double)id[i];
}
static __global__ void k_linear_in(PartPtrs p, uint64_t *id, long n, int D, const double *rec) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *r = rec + (size_t)i * (2 * D + 10);
  p.w[i] = *r++;
  for (int d = 0; d < D; ++d) p.x[d][i] = *r++;
  for (int d = 0; d < D; ++d) p.xold[d][i] = *r++;
  r += 2;
  for (int c = 0; c < 3; ++c) p.v[c][i] = *r++;
  for (int c = 0; c < 3; ++c) p.vold[c][i] = *r++;
  id[i] = (uint64_t)*r;
}

long pgpu_particle_linear_size(pgpu_species_t s) {
  return s ? (long)((2 * s->grid->desc.D + 10) * sizeof(double)) : -1;
}

int pgpu_species_download_linear(pgpu_species_t s, void *records) {
  NEED_INIT();
  if (!s || (s->n > 0 && !records)) return PGPU_ERR_ARG;
  if (s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  const int D = s->grid->desc.D;
  const size_t bytes = (size_t)s->n * (2 * D + 10) * sizeof(double);
  double *d = nullptr;
  PGPU_CUDA(cudaMalloc(&d, bytes));
  cudaStream_t st = ctx().stream;
  k_linear_out<<<nb(s->n), 256, 0, st>>>(s->ptrs(), s->id, s->n, D, d);
  PGPU_CUDA(cudaMemcpyAsync(records, d, bytes, cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

int pgpu_species_upload_linear(pgpu_species_t s, long n, const void *records) {
  NEED_INIT();
  if (!s || n < 0 || (n > 0 && !records)) return PGPU_ERR_ARG;
  if (ensure_capacity(s, n)) return PGPU_ERR_CUDA;
  const int D = s->grid->desc.D;
  s->n = n;
  s->binned = false;
  s->pos_old_pending = s->vel_old_pending = false;
  s->xold_alias = s->vold_alias = false;
  if (n == 0) return 0;
  const size_t bytes = (size_t)n * (2 * D + 10) * sizeof(double);
  double *d = nullptr;
  PGPU_CUDA(cudaMalloc(&d, bytes));
  cudaStream_t st = ctx().stream;
  PGPU_CUDA(cudaMemcpyAsync(d, records, bytes, cudaMemcpyHostToDevice, st));
  k_linear_in<<<nb(n), 256, 0, st>>>(s->ptrs(), s->id, n, D, d);
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

int pgpu_species_download_fields(pgpu_species_t s, double *Ep, double *Bp) {
  NEED_INIT();
  if (!s || !s->Ep[0]) {
    set_error("no particle fields stored; call pgpu_interpolate_fields_to_particles first");
    return PGPU_ERR_STATE;
  }
  const long n = s->n;
  cudaStream_t st = ctx().stream;
  for (int c = 0; c < 3; ++c) {
    if (Ep) PGPU_CUDA(cudaMemcpyAsync(Ep + (size_t)c * n, s->Ep[c], n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (Bp) PGPU_CUDA(cudaMemcpyAsync(Bp + (size_t)c * n, s->Bp[c], n * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  PGPU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int pgpu_species_upload_fields(pgpu_species_t s, const double *Ep, const double *Bp) {
  NEED_INIT();
  if (!s || !Ep || !Bp) return PGPU_ERR_ARG;
  if (ensure_epbp(s)) return PGPU_ERR_CUDA;
  const long n = s->n;
  cudaStream_t st = ctx().stream;
  for (int c = 0; c < 3; ++c) {
    PGPU_CUDA(cudaMemcpyAsync(s->Ep[c], Ep + (size_t)c * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
    PGPU_CUDA(cudaMemcpyAsync(s->Bp[c], Bp + (size_t)c * n, n * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  PGPU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- streaming passes ---------------------------------------------------------------
static int stream_pass(const char *name, double *out, const double *a, const double *b, long n, double s,
                       int mode) {
  if (n == 0) return 0;
  KTimer t(name);
  k_stream<<<nb(n), 256, 0, ctx().stream>>>(out, a, b, n, s, mode);
  return 0;
}

int pgpu_advance_positions_explicit(pgpu_species_t s, double full_dt, int half_step) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.motion) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  const double cnormDt = s->desc.cvac_norm * full_dt;
  const double dt_factor = half_step ? 0.5 : 1.0;
  if (s->desc.relativistic) {
    if (s->n == 0) return 0;
    KTimer t("advance_positions");
    k_positions_rel<<<nb(s->n), 256, 0, ctx().stream>>>(s->ptrs(), s->n, s->grid->desc.D, cnormDt * dt_factor, 0);
    s->binned = false;
    return 0;
  }
  for (int d = 0; d < s->grid->desc.D; ++d)
    stream_pass("advance_positions", s->x[d], s->v[d], s->xold[d], s->n, cnormDt * dt_factor, 0);
  s->binned = false;
  return 0;
}

int pgpu_advance_positions_implicit(pgpu_species_t s, double full_dt) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.motion) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  const double cnormDt = s->desc.cvac_norm * full_dt;
  const double cnormHalfDt = cnormDt * 0.5;
  if (s->desc.relativistic) {
    if (s->n == 0) return 0;
    KTimer t("advance_positions");
    k_positions_rel<<<nb(s->n), 256, 0, ctx().stream>>>(s->ptrs(), s->n, s->grid->desc.D, cnormHalfDt, 1);
    s->binned = false;
    return 0;
  }
  for (int d = 0; d < s->grid->desc.D; ++d)
    stream_pass("advance_positions", s->x[d], s->v[d], s->xold[d], s->n, cnormHalfDt, 0);
  s->binned = false;
  return 0;
}

int pgpu_advance_positions_2nd_half(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.motion) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  for (int d = 0; d < s->grid->desc.D; ++d)
    stream_pass("second_half", s->x[d], s->x[d], s->xold[d], s->n, 0.0, 1);
  s->binned = false;
  return 0;
}

int pgpu_advance_velocities_2nd_half(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.forces) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  for (int c = 0; c < 3; ++c) stream_pass("second_half", s->v[c], s->v[c], s->vold[c], s->n, 0.0, 1);
  return 0;
}

int pgpu_average_velocities(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.forces) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  for (int c = 0; c < 3; ++c) stream_pass("average_velocities", s->v[c], s->v[c], s->vold[c], s->n, 0.0, 2);
  return 0;
}

int pgpu_update_old_particle_positions(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.motion) return 0;
  s->pos_old_pending = false;   // every xold entry is overwritten: a pending gather is moot
  s->xold_alias = true;         // the copy itself is deferred (materialize_old) or never happens (CC1 tile kernel)
  return 0;
}

int pgpu_update_old_particle_velocities(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.forces) return 0;
  s->vel_old_pending = false;
  s->vold_alias = true;
  return 0;
}

int pgpu_reset_particles(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  for (int d = 0; d < s->grid->desc.D; ++d)
    PGPU_CUDA(cudaMemcpyAsync(s->x[d], s->xold[d], s->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
  for (int c = 0; c < 3; ++c)
    PGPU_CUDA(cudaMemcpyAsync(s->v[c], s->vold[c], s->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
  s->binned = false;
  return 0;
}

int pgpu_grid_set_external_fields(pgpu_grid_t g, const pgpu_ext_fn *six) {
  NEED_INIT();
  if (!g) return PGPU_ERR_ARG;
  memset(&g->ext, 0, sizeof(g->ext));
  if (!six) return 0;
  for (int c = 0; c < 6; ++c) {
    const pgpu_ext_fn &f = six[c];
    if (f.type < PGPU_EXT_NONE || f.type > PGPU_EXT_HEAVYSIDE) {
      set_error("external field %d: unknown grid function type %d", c, f.type);
      return PGPU_ERR_ARG;
    }
    ExtFn &o = g->ext.f[c];
    o.type = f.type;
    o.value = f.value;
    o.constant = f.constant;
    for (int d = 0; d < 2; ++d) {
      o.L[d] = f.L[d]; o.mode[d] = f.mode[d]; o.phase[d] = f.phase[d];
      o.C[d] = f.C[d]; o.A[d] = f.A[d]; o.X0[d] = f.X0[d]; o.eps[d] = f.eps[d];
    }
    if (f.type == PGPU_EXT_COSINE)
      for (int d = 0; d < g->desc.D; ++d)
        if (!(f.L[d] != 0.0)) {
          set_error("external field %d: Cosine needs L[%d] != 0", c, d);
          return PGPU_ERR_ARG;
        }
  }
  g->ext.on = 1;   // EMFields::externalFields(): components of type NONE contribute zero
  return 0;
}

int pgpu_add_external_fields_to_particles(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.forces || s->desc.charge == 0.0 || !s->grid->ext.on) return 0;   // PicChargedSpecies.cpp:3950-3952
  if (!s->Ep[0]) {
    set_error("addExternalFieldsToParticles needs particle fields: call pgpu_interpolate_fields_to_particles first");
    return PGPU_ERR_STATE;
  }
  return launch_add_external(s);
}

int pgpu_interpolate_fields_to_particles(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.forces || s->desc.charge == 0.0) return 0;
  if (ensure_epbp(s)) return PGPU_ERR_CUDA;
  int rc = launch_gather(s);
  if (rc) return rc;
  return pgpu_synchronize();
}

int pgpu_advance_velocities(pgpu_species_t s, double full_dt, int half_step) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.forces) return 0;
  if (!s->Ep[0]) {
    set_error("advanceVelocities needs particle fields: call pgpu_interpolate_fields_to_particles first");
    return PGPU_ERR_STATE;
  }
  if (s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  const double cnormDt = full_dt * s->desc.cvac_norm;
  const double alpha = s->desc.fnorm_const * cnormDt / 2.0;
  KTimer t("boris");
  k_boris_stored<<<nb(s->n), 256, 0, ctx().stream>>>(s->ptrs(), s->n, alpha, half_step, ctx().exact ? 1 : 0,
                                                     s->desc.relativistic, s->desc.higuera_cary);
  return 0;
}

static int ensure_virt(pgpu_species_t s) {
  if (s->virt[0] && s->virt_cap >= s->cap) return 0;
  for (int k = 0; k < 2; ++k) {
    double *q = nullptr;
    PGPU_CUDA(cudaMalloc(&q, s->cap * sizeof(double)));
    PGPU_CUDA(cudaMemsetAsync(q, 0, s->cap * sizeof(double), ctx().stream));
    if (s->virt[k]) {
      PGPU_CUDA(cudaMemcpyAsync(q, s->virt[k], (size_t)std::min<size_t>(s->virt_cap, s->cap) * sizeof(double),
                                cudaMemcpyDeviceToDevice, ctx().stream));
      PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
      cudaFree(s->virt[k]);
    }
    s->virt[k] = q;
  }
  s->virt_cap = s->cap;
  return 0;
}

int pgpu_apply_forces_curvilinear(pgpu_species_t s, int push_type, double full_dt, int by_half_dt, int anticyclic) {
  NEED_INIT();
  if (!s || push_type < PGPU_PUSH_CYL_CYL || push_type > PGPU_PUSH_SPH_HYB) return PGPU_ERR_ARG;
  if (!s->desc.forces) return 0;
  if (!s->Ep[0]) {
    set_error("applyForces needs particle fields: call pgpu_interpolate_fields_to_particles first");
    return PGPU_ERR_STATE;
  }
  if (s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  if (ensure_virt(s)) return PGPU_ERR_CUDA;
  const double cnormDt = full_dt * s->desc.cvac_norm;
  const double alpha = s->desc.fnorm_const * cnormDt / 2.0;
  KTimer t("boris_curvilinear");
  k_boris_curvilinear<<<nb(s->n), 256, 0, ctx().stream>>>(s->ptrs(), s->xold[0], s->virt[0], s->virt[1], s->n, push_type,
                                                          alpha, cnormDt, by_half_dt, anticyclic, s->desc.relativistic);
  return 0;
}

int pgpu_species_virtual_positions_set(pgpu_species_t s, const double *virt) {
  NEED_INIT();
  if (!s || (s->n > 0 && !virt)) return PGPU_ERR_ARG;
  if (s->n == 0) return 0;
  if (ensure_virt(s)) return PGPU_ERR_CUDA;
  for (int k = 0; k < 2; ++k)
    PGPU_CUDA(cudaMemcpyAsync(s->virt[k], virt + (size_t)k * s->n, (size_t)s->n * sizeof(double), cudaMemcpyHostToDevice,
                              ctx().stream));
  PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

int pgpu_species_virtual_positions_get(pgpu_species_t s, double *virt) {
  NEED_INIT();
  if (!s || (s->n > 0 && !virt)) return PGPU_ERR_ARG;
  if (s->n == 0) return 0;
  if (ensure_virt(s)) return PGPU_ERR_CUDA;
  for (int k = 0; k < 2; ++k)
    PGPU_CUDA(cudaMemcpyAsync(virt + (size_t)k * s->n, s->virt[k], (size_t)s->n * sizeof(double), cudaMemcpyDeviceToHost,
                              ctx().stream));
  PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

static AdvanceParams make_params(pgpu_species_t s, double dt, bool iterative) {
  AdvanceParams p;
  p.cnormDt = dt * s->desc.cvac_norm;
  p.alpha = s->desc.fnorm_const * p.cnormDt / 2.0;
  p.rtol = s->desc.rtol;
  p.iter_max = iterative ? s->desc.iter_max : -1;
  p.order_swap = s->desc.order_swap;
  const GeoAny &g = s->grid->geo;
  p.volume = (g.D == 1) ? g.dx[0] : g.dx[0] * g.dx[1];
  p.rvolume = 1.0 / p.volume;
  p.rel = s->desc.relativistic;
  p.hc = s->desc.higuera_cary;
  p.ext = s->grid->ext;
  p.fnorm = s->desc.fnorm_const;
  p.suborbit = 0;
  p.explicit_step = 0;
  p.unconv_list = nullptr;
  p.unconv_count = nullptr;
  return p;
}

int pgpu_advance_particles(pgpu_species_t s, double dt) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.motion || !s->desc.forces || s->desc.charge == 0.0) {
    // degenerate switches: compose the reference sequence from the separate passes
    if (s->desc.order_swap) pgpu_advance_positions_implicit(s, dt);
    if (s->desc.forces && s->desc.charge != 0.0) {
      int rc = pgpu_interpolate_fields_to_particles(s);
      if (rc) return rc;
      pgpu_advance_velocities(s, dt, 1);
    }
    if (!s->desc.order_swap) pgpu_advance_positions_implicit(s, dt);
    return 0;
  }
  s->binned = false;
  return launch_advance(s, make_params(s, dt, false), false);
}

static int scale_species_current(pgpu_species_t s) {
  const double f = s->desc.charge / s->grid->desc.volume_scale;
  Ptr3 P;
  long mx = 0;
  for (int c = 0; c < 3; ++c) {
    P.a[c] = s->J[c].p;
    P.b[c] = nullptr;
    P.n[c] = (long)s->J[c].size();
    mx = std::max(mx, P.n[c]);
  }
  KTimer t("current_scale");
  k_scale3<<<dim3(nb(mx), 3), 256, 0, ctx().stream>>>(P, f);
  return 0;
}

int pgpu_advance_particles_iteratively(pgpu_species_t s, double dt, int deposit_J, pgpu_picard_stats *stats) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  const bool iterative = !(s->desc.iter_max == 0 || !s->desc.motion || !s->desc.forces || s->desc.charge == 0.0);
  int rc = 0;
  if (!iterative && (!s->desc.motion || !s->desc.forces || s->desc.charge == 0.0)) {
    rc = pgpu_advance_particles(s, dt);
    if (rc) return rc;
    if (deposit_J && s->desc.charge != 0.0) rc = pgpu_set_current_density(s, dt, 0);
    if (stats) memset(stats, 0, sizeof(*stats));
    return rc;
  }
  s->binned = false;
  ctx().total_advances += s->n;
  const bool fuse = deposit_J && s->desc.interp_J == s->desc.interp_E && !ctx().exact;
  if (deposit_J)
    for (int c = 0; c < 3; ++c)
      PGPU_CUDA(cudaMemsetAsync(s->J[c].p, 0, s->J[c].size() * sizeof(double), ctx().stream));
  AdvanceParams prm = make_params(s, dt, iterative);
  const bool sub = iterative && s->use_suborbit_model;
  if (sub) {
    // m_use_suborbit_model (:1699-1706): particles the loop leaves unconverged are listed, deposit nothing here and move
    // to the sub-orbit container
    if (ensure_unconv_list(s)) return PGPU_ERR_CUDA;
    PGPU_CUDA(cudaMemsetAsync(s->unconv_count, 0, sizeof(unsigned), ctx().stream));
    prm.suborbit = 1;
    prm.unconv_list = s->unconv_list;
    prm.unconv_count = s->unconv_count;
  }
  rc = launch_advance(s, prm, fuse);
  if (rc) return rc;
  if (sub) {
    unsigned count = 0;
    PGPU_CUDA(cudaMemcpyAsync(&count, s->unconv_count, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx().stream));
    PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
    if (count && !fuse && deposit_J) {
      set_error("sub-orbit model: the separate deposit pass would include the particles that moved to the sub-orbit container");
    }
    rc = transfer_listed_to_suborbit(s, count);
    if (rc) return rc;
  }
  if (deposit_J) {
    if (!fuse) {
      rc = launch_deposit_current(s, dt * s->desc.cvac_norm);
      if (rc) return rc;
    }
    scale_species_current(s);
  }
  if (stats) {
    Counters k;
    if (fetch_counters(&k)) return PGPU_ERR_CUDA;
    stats->num_parts_its = iterative ? s->n : 0;
    stats->num_apply_its = iterative ? (long)k.apply_its : 0;
    stats->num_unconverged = (long)k.unconverged;
    return check_err_bits(k.err);
  }
  return 0;
}

int pgpu_set_current_density(pgpu_species_t s, double dt, int from_explicit_solver) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  s->dep_from_explicit = from_explicit_solver ? 1 : 0;
  for (int c = 0; c < 3; ++c)
    PGPU_CUDA(cudaMemsetAsync(s->J[c].p, 0, s->J[c].size() * sizeof(double), ctx().stream));
  int rc = launch_deposit_current(s, dt * s->desc.cvac_norm);
  if (rc) return rc;
  if (from_explicit_solver && s->n_out > 0) {   // depositInflowOutflowJ (PicChargedSpecies.cpp:3232-3235)
    rc = launch_deposit_outflow(s);
    if (rc) return rc;
  }
  return scale_species_current(s);
}

// The particle side of one PIC_EM_EXPLICIT leap-frog step (PICTimeIntegrator_EM_Explicit.cpp:92-170) in one pass:
// see k_explicit_step (pgpu_push.cu).  Composes the separate calls where the fused kernel does not apply (relativistic
// species, non-periodic particle BCs, switched-off motion or forces).
int pgpu_explicit_step(pgpu_species_t s, double dt, const int *bc_lo, const int *bc_hi, int second_half) {
  NEED_INIT();
  if (!s || !bc_lo || !bc_hi) return PGPU_ERR_ARG;
  const int D = s->grid->desc.D;
  bool fused = !s->desc.relativistic && s->desc.motion && s->desc.forces && s->desc.charge != 0.0;
  int periodic[2] = {0, 0};
  for (int d = 0; d < D; ++d) {
    if (bc_lo[d] != bc_hi[d] || (bc_lo[d] != PGPU_BC_PERIODIC && bc_lo[d] != PGPU_BC_NONE)) fused = false;
    periodic[d] = bc_lo[d] == PGPU_BC_PERIODIC;
  }
  for (int d = 0; d < D; ++d)
    if (s->desc.bc_check_lo[d] || s->desc.bc_check_hi[d]) fused = false;
  if (!fused) {
    int rc = pgpu_interpolate_fields_to_particles(s);
    if (!rc) rc = pgpu_add_external_fields_to_particles(s);
    if (!rc) rc = pgpu_advance_velocities(s, dt, 0);
    if (!rc) rc = pgpu_advance_positions_explicit(s, dt, 1);
    if (!rc) rc = pgpu_apply_bcs(s, bc_lo, bc_hi);
    if (!rc) rc = pgpu_set_current_density(s, dt, 1);
    if (!rc && second_half) {
      rc = pgpu_advance_positions_2nd_half(s);
      if (!rc) rc = pgpu_apply_bcs(s, bc_lo, bc_hi);
    }
    return rc;
  }
  s->binned = false;
  s->dep_from_explicit = 1;
  for (int c = 0; c < 3; ++c)
    PGPU_CUDA(cudaMemsetAsync(s->J[c].p, 0, s->J[c].size() * sizeof(double), ctx().stream));
  AdvanceParams prm = make_params(s, dt, false);
  if (D == 2 && s->desc.interp_E == CC1 && s->desc.interp_J == CC1 && !prm.ext.on && !ctx().exact && periodic[0] &&
      periodic[1]) {
    // 2D CC1: the tile kernel of the implicit advance does the step for the particles whose orbit stays in one dual
    // cell (single pass, u_new = 2 ubar - u_old, deposit of u_new); the one-pass visitor kernel redoes the ones it lists.
    // The periodic wrap and the second half then run as the streaming passes of the separate calls.
    prm.explicit_step = 1;
    prm.order_swap = 0;
    const int fast = launch_advance_cc1_fast(s, prm, true);
    if (fast < 0) return fast;
    if (fast > 0) {
      int rc = launch_explicit_step(s, prm, periodic, false, true);
      if (!rc) rc = scale_species_current(s);
      if (!rc) rc = pgpu_apply_bcs(s, bc_lo, bc_hi);
      if (!rc && second_half) {
        rc = pgpu_advance_positions_2nd_half(s);
        if (!rc) rc = pgpu_apply_bcs(s, bc_lo, bc_hi);
      }
      return rc;
    }
    prm.explicit_step = 0;
  }
  int rc = launch_explicit_step(s, prm, periodic, second_half != 0);
  if (rc) return rc;
  return scale_species_current(s);
}

int pgpu_species_current_get(pgpu_species_t s, int comp, double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!s || comp < 0 || comp >= 3) return PGPU_ERR_ARG;
  int rc = copy_fab_to_host(s->J[comp], s->grid->desc.D, data, lo, hi);
  if (rc) return rc;
  return pgpu_synchronize();
}

// ---- instrumentation ------------------------------------------------------------------
int pgpu_picard_totals(long *advances, long *apply_its, long *unconverged, int reset) {
  NEED_INIT();
  Counters k;
  if (fetch_counters(&k)) return PGPU_ERR_CUDA;
  Context &c = ctx();
  if (advances) *advances = c.total_advances;
  if (apply_its) *apply_its = c.total_apply_its;
  if (unconverged) *unconverged = c.total_unconverged;
  if (reset) c.total_advances = c.total_apply_its = c.total_unconverged = 0;
  return check_err_bits(k.err);
}

int pgpu_species_deferred_count(pgpu_species_t s, long *count) {
  NEED_INIT();
  if (!s || !count) return PGPU_ERR_ARG;
  unsigned n = 0;
  const unsigned *src = s->defer_count_first ? s->defer_count_first : s->defer_count;
  if (src) {
    PGPU_CUDA(cudaMemcpyAsync(&n, src, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx().stream));
    PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
  }
  *count = (long)n;
  return 0;
}

int pgpu_profile_enable(int on) {
  ctx().profile = on != 0;
  return 0;
}
int pgpu_profile_reset(void) {
  profile_drain();
  ctx().prof.clear();
  ctx().launches = 0;
  return 0;
}
int pgpu_profile_query(const char *prefix, double *ms, long *launches) {
  profile_drain();
  double t = 0;
  long k = 0;
  const size_t len = prefix ? strlen(prefix) : 0;
  for (auto &e : ctx().prof)
    if (len == 0 || e.first.compare(0, len, prefix) == 0) {
      t += e.second.first;
      k += e.second.second;
    }
  if (ms) *ms = t;
  if (launches) *launches = k;
  return 0;
}
long pgpu_launch_count(void) { return ctx().launches; }

}  // extern "C"
