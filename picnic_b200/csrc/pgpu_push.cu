// pgpu_push.cu -- gather, deposit and the fused implicit advance kernels.
//
// One thread per particle, SoA arrays, coalesced 8-byte loads.  The fused kernel
// keeps a particle in registers through every particle-Picard pass
// (PicChargedSpecies::advanceParticlesIteratively, PicChargedSpecies.cpp:1614-1716)
// and, when asked, deposits its current before retiring it
// (PicChargedSpecies::setCurrentDensity, :3184-3253), so the compulsory HBM traffic
// is one read of (x_old, u_old, xbar, w) and one write of (xbar, ubar).
#include "pgpu_internal.h"

namespace pgpu {

template <int D, bool X>
struct GatherOp {
  const FieldSet &F;
  double acc[6];
  bool oob;
  __device__ __forceinline__ GatherOp(const FieldSet &f) : F(f), oob(false) {
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[c] = 0.0;
  }
  __device__ __forceinline__ void operator()(int c, int i, int j, double w) {
    const FabView &f = F.f[c];
    const unsigned a = (unsigned)(i - f.lo0);
    const unsigned b = (D == 2) ? (unsigned)(j - f.lo1) : 0u;
    if (a < (unsigned)f.n0 && b < (unsigned)f.n1) {
      const double val = __ldg(f.p + (a + (size_t)b * f.n0));
      acc[c] = M<X>::mad(w, val, acc[c]);
    } else {
      oob = true;
    }
  }
};

// deposit straight into HBM/L2 with native fp64 reductions (RED.E.ADD.F64)
template <int D, bool X>
struct DepositOpGlobal {
  const CurrentSet &J;
  double val[3];  // v_c * (w/volume)
  bool oob;
  bool aggregate;  // false for the particles the CC1 tile kernel deferred: one in twenty, no common addresses
  __device__ __forceinline__ DepositOpGlobal(const CurrentSet &j, bool agg = true) : J(j), oob(false), aggregate(agg) {}
  __device__ __forceinline__ void operator()(int c, int i, int j, double w) {
    const FabView &f = J.j[c];
    const unsigned a = (unsigned)(i - f.lo0);
    const unsigned b = (D == 2) ? (unsigned)(j - f.lo1) : 0u;
    const bool in = a < (unsigned)f.n0 && b < (unsigned)f.n1;
    if (!in) oob = true;
    // exact mode keeps one atomic per particle and stencil point (the summation order of the reference's loop
    // is not reproduced either way, but the per-particle products are)
    if (X || !aggregate) {
      if (in) atomicAdd(f.p + (a + (size_t)b * f.n0), M<X>::mul(val[c], w));
    } else {
      warp_aggregated_add(f.p, in ? (long)(a + (size_t)b * f.n0) : -1L, val[c] * w);
    }
  }
};

__device__ __forceinline__ void flush_counters(Counters *cnt, unsigned apply, unsigned unconv,
                                               unsigned err) {
  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  err = __reduce_or_sync(0xffffffffu, err);
  if ((threadIdx.x & 31) == 0) {
    if (apply) atomicAdd(&cnt->apply_its, (unsigned long long)apply);
    if (unconv) atomicAdd(&cnt->unconverged, (unsigned long long)unconv);
    if (err) atomicOr(&cnt->err, err);
  }
}

// ---- interpolateFieldsToParticles: store Ep,Bp ------------------------------------
template <int D, int IE, bool X>
__global__ void __launch_bounds__(256)
k_gather(PartPtrs p, long n, Geo<D> g, FieldSet F, Counters *cnt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned err = 0;
  if (i < n) {
    double xp[D], xo[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xp[d] = p.x[d][i];
      xo[d] = p.xold[d][i];
    }
    GatherOp<D, X> op(F);
    if (!gather_visit<D, IE, X>(g, xp, xo, op)) err |= ERRBIT_SEGMENTS;
    if (op.oob) err |= ERRBIT_BOUNDS;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      p.Ep[c][i] = op.acc[c];
      p.Bp[c][i] = op.acc[3 + c];
    }
  }
  flush_counters(cnt, 0, 0, err);
}

// ---- addExternalFieldsToParticles on the stored Ep, Bp (PicChargedSpecies.cpp:3967-3996) ----------------------------
template <int D>
__global__ void __launch_bounds__(256) k_add_external(PartPtrs p, long n, ExtFields ext) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x[D], acc[6];
#pragma unroll
  for (int d = 0; d < D; ++d) x[d] = p.x[d][i];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    acc[c] = p.Ep[c][i];
    acc[3 + c] = p.Bp[c][i];
  }
  add_external<D>(ext, x, acc);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    p.Ep[c][i] = acc[c];
    p.Bp[c][i] = acc[3 + c];
  }
}
int launch_add_external(pgpu_species_s *s) {
  if (s->n == 0 || !s->grid->ext.on) return 0;
  KTimer t("add_external");
  const unsigned nb = (unsigned)((s->n + 255) / 256);
  if (s->grid->desc.D == 1) k_add_external<1><<<nb, 256, 0, ctx().stream>>>(s->ptrs(), s->n, s->grid->ext);
  else k_add_external<2><<<nb, 256, 0, ctx().stream>>>(s->ptrs(), s->n, s->grid->ext);
  return 0;
}

// ---- setCurrentDensity: deposit only ------------------------------------------------
template <int D, int IJ, bool X>
__global__ void __launch_bounds__(256)
k_deposit(PartPtrs p, long n, Geo<D> g, CurrentSet J, double volume, double rvolume, Counters *cnt, int rel,
          int from_explicit) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned err = 0;
  if (i < n) {
    double xp[D], xo[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xp[d] = p.x[d][i];
      xo[d] = p.xold[d][i];
    }
    double wp = p.w[i];
    if (rel) {   // wpog = wp / gammap (MeshInterpI.H:72-91)
      const double u[3] = {p.v[0][i], p.v[1][i], p.v[2][i]};
      double gammap;
      if (from_explicit) {
        gammap = gamma_sum_first<X>(u);
      } else {
        const double uo[3] = {p.vold[0][i], p.vold[1][i], p.vold[2][i]};
        gammap = gamma_implicit<X>(uo, u);
      }
      wp = __ddiv_rn(wp, gammap);
    }
    const double rhop = X ? __ddiv_rn(wp, volume) : wp * rvolume;
    DepositOpGlobal<D, X> op(J);
#pragma unroll
    for (int c = 0; c < 3; ++c) op.val[c] = M<X>::mul(p.v[c][i], rhop);
    if (!deposit_visit<D, IJ, X>(g, xp, xo, op)) err |= ERRBIT_SEGMENTS;
    if (op.oob) err |= ERRBIT_BOUNDS;
  }
  flush_counters(cnt, 0, 0, err);
}

// ---- advanceParticles / advanceParticlesIteratively (+ optional fused deposit) ------
#ifndef PGPU_ADV_MINB
#define PGPU_ADV_MINB 2
#endif
template <int D, int IE, bool X, bool DEP>
__global__ void __launch_bounds__(256, PGPU_ADV_MINB)
k_advance(PartPtrs p, long n, Geo<D> g, FieldSet F, CurrentSet J, AdvanceParams prm, Counters *cnt,
          const int *list, const unsigned *list_count) {
  typedef M<X> m;
  // list == nullptr: thread t handles particle t (the grid covers n).  Otherwise the
  // particles to do are list[0 .. *list_count) -- the ones the CC1 fast kernel deferred
  // (pgpu_advance_cc1.cu) -- walked with a grid stride.
  const long total = list ? (long)*list_count : n;
  const long stride = (long)gridDim.x * blockDim.x;
  unsigned err = 0, apply = 0, unconv = 0;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long i = list ? (long)list[t] : t;
    double xb[D], xo[D], uo[3], ub[3];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xb[d] = p.x[d][i];
      xo[d] = p.xold[d][i];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uo[c] = p.vold[c][i];
      ub[c] = uo[c];  // only survives if the first gather fails
    }
    const double hdt = m::mul(prm.cnormDt, 0.5);  // cnormHalfDt
    bool left_unconverged = false;

    if (prm.iter_max < 0) {
      // advanceParticles (PicChargedSpecies.cpp:1594-1612)
      if (prm.order_swap) {
        if (prm.rel) {   // advancePositionsImplicit of the relativistic build (:548-553)
          const double us[3] = {p.v[0][i], p.v[1][i], p.v[2][i]};
          const double gammap = gamma_implicit<X>(uo, us);
#pragma unroll
          for (int d = 0; d < D; ++d) xb[d] = m::add(xo[d], m::mul(__ddiv_rn(us[d], gammap), hdt));
        } else {
#pragma unroll
          for (int d = 0; d < D; ++d) xb[d] = m::mad(p.v[d][i], hdt, xo[d]);
        }
      }
      GatherOp<D, X> op(F);
      if (!gather_visit<D, IE, X>(g, xb, xo, op)) err |= ERRBIT_SEGMENTS;
      if (op.oob) err |= ERRBIT_BOUNDS;
      if (prm.ext.on) add_external<D>(prm.ext, xb, op.acc);   // :1606
      boris<X>(uo, op.acc, op.acc + 3, prm.alpha, true, ub, prm.rel, prm.hc);
      apply += 1;
      if (!prm.order_swap) {
        if (prm.rel) {
          const double gammap = gamma_implicit<X>(uo, ub);
#pragma unroll
          for (int d = 0; d < D; ++d) xb[d] = m::add(xo[d], m::mul(__ddiv_rn(ub[d], gammap), hdt));
        } else {
#pragma unroll
          for (int d = 0; d < D; ++d) xb[d] = m::mad(ub[d], hdt, xo[d]);
        }
      }
    } else {
      // advanceParticlesIteratively (:1614-1716) with stepNormTransfer (:658-733)
      bool converged = false;
      int iter = 0;  // 0 = the initial, non-reverse pass
      while (true) {
        GatherOp<D, X> op(F);
        if (!gather_visit<D, IE, X>(g, xb, xo, op)) {
          err |= ERRBIT_SEGMENTS;
          break;  // the reference aborts here (Fortran STOP)
        }
        if (op.oob) {
          err |= ERRBIT_BOUNDS;
          break;
        }
        if (prm.ext.on) add_external<D>(prm.ext, xb, op.acc);   // :1652, :1669
        boris<X>(uo, op.acc, op.acc + 3, prm.alpha, true, ub, prm.rel, prm.hc);
        apply += 1;
        double dxp[D];
        double rel_diff_max = 0.0;
        const double gammap = prm.rel ? gamma_implicit<X>(uo, ub) : 1.0;   // stepNormTransfer :693-708
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const double dxp0 = m::sub(xb[d], xo[d]);
          dxp[d] = prm.rel ? m::mul(__ddiv_rn(ub[d], gammap), hdt) : m::mul(ub[d], hdt);
          const double rel = __ddiv_rn(fabs(m::sub(dxp0, dxp[d])), g.dx[d]);
          rel_diff_max = fmax(rel_diff_max, rel);
        }
        if (iter == 0) {
#pragma unroll
          for (int d = 0; d < D; ++d) xb[d] = m::add(xo[d], dxp[d]);
          converged = !(rel_diff_max >= prm.rtol);
        } else {
          if (rel_diff_max < prm.rtol) converged = true;
          else {
#pragma unroll
            for (int d = 0; d < D; ++d) xb[d] = m::add(xo[d], dxp[d]);
          }
        }
        if (converged) break;
        if (iter >= prm.iter_max) {
          // iter counts reverse passes done; cap reached (:1678-1693)
          unconv += 1;
          left_unconverged = true;
          break;
        }
        iter += 1;
      }
    }
    if (left_unconverged && prm.suborbit) {
      // m_use_suborbit_model (:1699-1706): off to the sub-orbit container with two sub-orbits; it deposits there
      const unsigned slot = atomicAdd(prm.unconv_count, 1u);
      prm.unconv_list[slot] = (int)i;
    }
#pragma unroll
    for (int d = 0; d < D; ++d) p.x[d][i] = xb[d];
#pragma unroll
    for (int c = 0; c < 3; ++c) p.v[c][i] = ub[c];

    if (DEP && !(err & (ERRBIT_SEGMENTS | ERRBIT_BOUNDS)) && !(left_unconverged && prm.suborbit)) {
      double wp = p.w[i];
      if (prm.rel) wp = __ddiv_rn(wp, gamma_implicit<X>(uo, ub));   // MeshInterpI.H:78-89
      const double rhop = X ? __ddiv_rn(wp, prm.volume) : wp * prm.rvolume;
      DepositOpGlobal<D, X> dop(J, list == nullptr);
#pragma unroll
      for (int c = 0; c < 3; ++c) dop.val[c] = m::mul(ub[c], rhop);
      if (!deposit_visit<D, IE, X>(g, xb, xo, dop)) err |= ERRBIT_SEGMENTS;
      if (dop.oob) err |= ERRBIT_BOUNDS;
    }
  }
  flush_counters(cnt, apply, unconv, err);
}

// ---------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------
static inline unsigned nblocks(long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

template <int D, int IE>
static int launch_gather_t(pgpu_species_s *s) {
  Context &c = ctx();
  const Geo<D> g = make_geo<D>(species_geo(s));
  const FieldSet F = grid_fields(s->grid);
  KTimer t("gather");
  if (c.exact)
    k_gather<D, IE, true><<<nblocks(s->n, 256), 256, 0, c.stream>>>(s->ptrs(), s->n, g, F, c.d_counters);
  else
    k_gather<D, IE, false><<<nblocks(s->n, 256), 256, 0, c.stream>>>(s->ptrs(), s->n, g, F, c.d_counters);
  return 0;
}

template <int D>
static int launch_gather_d(pgpu_species_s *s) {
  switch (s->desc.interp_E) {
    case CIC: return launch_gather_t<D, CIC>(s);
    case TSC: return launch_gather_t<D, TSC>(s);
    case CC0: return launch_gather_t<D, CC0>(s);
    case CC1: return launch_gather_t<D, CC1>(s);
  }
  return PGPU_ERR_ARG;
}

int launch_gather(pgpu_species_s *s) {
  if (s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  return s->grid->desc.D == 1 ? launch_gather_d<1>(s) : launch_gather_d<2>(s);
}

template <int D, int IJ>
static int launch_deposit_t(pgpu_species_s *s) {
  Context &c = ctx();
  const GeoAny ga = species_geo(s);
  const Geo<D> g = make_geo<D>(ga);
  const CurrentSet J = species_current(s);
  const double volume = (D == 1) ? ga.dx[0] : ga.dx[0] * ga.dx[1];
  KTimer t("deposit_current");
  if (c.exact)
    k_deposit<D, IJ, true><<<nblocks(s->n, 256), 256, 0, c.stream>>>(s->ptrs(), s->n, g, J, volume, 1.0 / volume,
                                                                      c.d_counters, s->desc.relativistic,
                                                                      s->dep_from_explicit);
  else
    k_deposit<D, IJ, false><<<nblocks(s->n, 256), 256, 0, c.stream>>>(s->ptrs(), s->n, g, J, volume, 1.0 / volume,
                                                                       c.d_counters, s->desc.relativistic,
                                                                       s->dep_from_explicit);
  return 0;
}

template <int D, int IJ>
static int launch_deposit_out_t(pgpu_species_s *s) {
  Context &c = ctx();
  const GeoAny ga = species_geo(s);
  const Geo<D> g = make_geo<D>(ga);
  const CurrentSet J = species_current(s);
  const double volume = (D == 1) ? ga.dx[0] : ga.dx[0] * ga.dx[1];
  KTimer t("deposit_outflow");
  if (c.exact)
    k_deposit<D, IJ, true><<<nblocks(s->n_out, 256), 256, 0, c.stream>>>(outflow_part_ptrs(s), s->n_out, g, J, volume,
                                                                          1.0 / volume, c.d_counters, s->desc.relativistic, 1);
  else
    k_deposit<D, IJ, false><<<nblocks(s->n_out, 256), 256, 0, c.stream>>>(outflow_part_ptrs(s), s->n_out, g, J, volume,
                                                                           1.0 / volume, c.d_counters, s->desc.relativistic, 1);
  return 0;
}
template <int D>
static int launch_deposit_out_d(pgpu_species_s *s) {
  switch (s->desc.interp_J) {
    case CIC: return launch_deposit_out_t<D, CIC>(s);
    case TSC: return launch_deposit_out_t<D, TSC>(s);
    case CC0: return launch_deposit_out_t<D, CC0>(s);
    case CC1: return launch_deposit_out_t<D, CC1>(s);
  }
  return PGPU_ERR_ARG;
}
int launch_deposit_outflow(pgpu_species_s *s) {
  if (s->n_out == 0) return 0;
  return s->grid->desc.D == 1 ? launch_deposit_out_d<1>(s) : launch_deposit_out_d<2>(s);
}

template <int D>
static int launch_deposit_d(pgpu_species_s *s) {
  switch (s->desc.interp_J) {
    case CIC: return launch_deposit_t<D, CIC>(s);
    case TSC: return launch_deposit_t<D, TSC>(s);
    case CC0: return launch_deposit_t<D, CC0>(s);
    case CC1: return launch_deposit_t<D, CC1>(s);
  }
  return PGPU_ERR_ARG;
}

int launch_deposit_current(pgpu_species_s *s, double /*cnormDt*/) {
  if (s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  return s->grid->desc.D == 1 ? launch_deposit_d<1>(s) : launch_deposit_d<2>(s);
}

template <int D, int IE>
static int launch_advance_t(pgpu_species_s *s, const AdvanceParams &prm, bool fuse, bool deferred) {
  Context &c = ctx();
  const Geo<D> g = make_geo<D>(species_geo(s));
  const FieldSet F = grid_fields(s->grid);
  const CurrentSet J = species_current(s);
  // deferred list: its length lives on the device, so walk it with a fixed grid
  const unsigned nb = deferred ? (unsigned)(c.sm_count * 4) : nblocks(s->n, 256);
  const int *list = deferred ? s->defer_list : nullptr;
  const unsigned *cnt = deferred ? s->defer_count : nullptr;
  if (fuse) {
    // the fused deposit exists in fast arithmetic only; exact mode runs the two
    // kernels back to back (the API never asks for exact+fused)
    if (c.exact) return PGPU_ERR_STATE;
    KTimer t(deferred ? "advance_deferred" : "advance_deposit_fused");
    k_advance<D, IE, false, true><<<nb, 256, 0, c.stream>>>(s->ptrs(), s->n, g, F, J, prm, c.d_counters, list, cnt);
  } else {
    KTimer t(deferred ? "advance_deferred" : "advance");
    if (c.exact)
      k_advance<D, IE, true, false><<<nb, 256, 0, c.stream>>>(s->ptrs(), s->n, g, F, J, prm, c.d_counters, list, cnt);
    else
      k_advance<D, IE, false, false><<<nb, 256, 0, c.stream>>>(s->ptrs(), s->n, g, F, J, prm, c.d_counters, list, cnt);
  }
  return 0;
}

template <int D>
static int launch_advance_d(pgpu_species_s *s, const AdvanceParams &prm, bool fuse, bool deferred) {
  switch (s->desc.interp_E) {
    case CIC: return launch_advance_t<D, CIC>(s, prm, fuse, deferred);
    case TSC: return launch_advance_t<D, TSC>(s, prm, fuse, deferred);
    case CC0: return launch_advance_t<D, CC0>(s, prm, fuse, deferred);
    case CC1: return launch_advance_t<D, CC1>(s, prm, fuse, deferred);
  }
  return PGPU_ERR_ARG;
}

// fuse_deposit requires interp_J == interp_E (the fused kernel reuses one visitor type).
// 2D CC1 species in fast arithmetic take the specialised kernel (pgpu_advance_cc1.cu)
// first; this generic visitor kernel then handles only the particles it deferred.
// ---- PIC_EM_EXPLICIT: the particle side of one leap-frog step in ONE pass ----------------------------------------------
// PICTimeIntegrator_EM_Explicit::timeStep (src/time/PICTimeIntegrator_EM_Explicit.cpp:92-170), default branch (no Strang
// splitting, no averaged-v deposit, no scattering between the calls):
//   interpolateFieldsToParticles + addExternalFieldsToParticles + advanceVelocities(dt, false)        (:94-111)
//   advancePositionsExplicit(dt/2) + applyBCs                                                        (:126-128)
//   setCurrentDensity(dt, from_explicit = true)                                                      (:137)
//   advancePositions_2ndHalf + applyBCs  (optional: `second_half`)                                    (:166-168)
// The separate calls move 48 B of stored E_p, B_p per particle out and back in and read the particle arrays four times;
// here a particle is read once (x, x_old, u_old, w) and written once (x, u).  Same arithmetic, same order per particle.
struct ExplicitArgs {
  int periodic[2];
  double left[2], right[2];
  int second_half;
  int deferred;   // host side only: walk the species' deferred list
};
template <int D>
__device__ __forceinline__ void wrap_periodic(const ExplicitArgs &e, double *xp, double *xo) {
#pragma unroll
  for (int d = 0; d < D; ++d) {
    if (!e.periodic[d]) continue;
    const double Lbox = __dsub_rn(e.right[d], e.left[d]);   // PicChargedSpeciesBC::enforcePeriodic (:738-765)
    if (xp[d] < e.left[d]) {
      xp[d] = __dadd_rn(xp[d], Lbox);
      xo[d] = __dadd_rn(xo[d], Lbox);
    }
    if (xp[d] >= e.right[d]) {
      xp[d] = __dsub_rn(xp[d], Lbox);
      xo[d] = __dsub_rn(xo[d], Lbox);
    }
  }
}
template <int D, int IE, int IJ, bool X>
__global__ void __launch_bounds__(256)
k_explicit_step(PartPtrs p, long n, Geo<D> g, FieldSet F, CurrentSet J, AdvanceParams prm, ExplicitArgs e, Counters *cnt,
                const int *list, const unsigned *list_count) {
  typedef M<X> m;
  // list != nullptr: the particles the CC1 tile kernel deferred (grid stride over list[0 .. *list_count))
  const long total = list ? (long)*list_count : n;
  const long stride = (long)gridDim.x * blockDim.x;
  unsigned err = 0;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long i = list ? (long)list[t] : t;
    double xp[D], xo[D], uo[3], u[3];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xp[d] = p.x[d][i];
      xo[d] = p.xold[d][i];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) uo[c] = p.vold[c][i];
    GatherOp<D, X> op(F);
    if (!gather_visit<D, IE, X>(g, xp, xo, op)) err |= ERRBIT_SEGMENTS;
    if (op.oob) err |= ERRBIT_BOUNDS;
    if (prm.ext.on) add_external<D>(prm.ext, xp, op.acc);
    boris<X>(uo, op.acc, op.acc + 3, prm.alpha, false, u, 0, 0);   // advanceVelocities(dt, byHalfDt = false)
    const double hdt = m::mul(prm.cnormDt, 0.5);
#pragma unroll
    for (int d = 0; d < D; ++d) xp[d] = m::mad(u[d], hdt, xo[d]);   // advancePositionsExplicit(dt/2): x = x_old + u cnormHalfDt
    wrap_periodic<D>(e, xp, xo);
    if (!(err & (ERRBIT_SEGMENTS | ERRBIT_BOUNDS))) {
      const double rhop = X ? __ddiv_rn(p.w[i], prm.volume) : p.w[i] * prm.rvolume;
      DepositOpGlobal<D, X> dop(J);
#pragma unroll
      for (int c = 0; c < 3; ++c) dop.val[c] = m::mul(u[c], rhop);
      if (!deposit_visit<D, IJ, X>(g, xp, xo, dop)) err |= ERRBIT_SEGMENTS;
      if (dop.oob) err |= ERRBIT_BOUNDS;
    }
    if (e.second_half) {   // advancePositions_2ndHalf: x = 2 x - x_old, then applyBCs
#pragma unroll
      for (int d = 0; d < D; ++d) xp[d] = m::sub(m::mul(2.0, xp[d]), xo[d]);
      wrap_periodic<D>(e, xp, xo);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) {
      p.x[d][i] = xp[d];
      p.xold[d][i] = xo[d];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) p.v[c][i] = u[c];
  }
  flush_counters(cnt, 0, 0, err);
}

template <int D, int IE, int IJ>
static int launch_explicit_t(pgpu_species_s *s, const AdvanceParams &prm, const ExplicitArgs &e) {
  Context &c = ctx();
  const Geo<D> g = make_geo<D>(species_geo(s));
  const FieldSet F = grid_fields(s->grid);
  const CurrentSet J = species_current(s);
  const bool deferred = e.deferred != 0;
  const unsigned nb = deferred ? (unsigned)(c.sm_count * 4) : nblocks(s->n, 256);
  const int *list = deferred ? s->defer_list : nullptr;
  const unsigned *cnt = deferred ? s->defer_count : nullptr;
  KTimer t(deferred ? "explicit_step_deferred" : "explicit_step_fused");
  if (c.exact)
    k_explicit_step<D, IE, IJ, true><<<nb, 256, 0, c.stream>>>(s->ptrs(), s->n, g, F, J, prm, e, c.d_counters, list, cnt);
  else
    k_explicit_step<D, IE, IJ, false><<<nb, 256, 0, c.stream>>>(s->ptrs(), s->n, g, F, J, prm, e, c.d_counters, list, cnt);
  return 0;
}
template <int D, int IE>
static int launch_explicit_e(pgpu_species_s *s, const AdvanceParams &prm, const ExplicitArgs &e) {
  switch (s->desc.interp_J) {
    case CIC: return launch_explicit_t<D, IE, CIC>(s, prm, e);
    case TSC: return launch_explicit_t<D, IE, TSC>(s, prm, e);
    case CC0: return launch_explicit_t<D, IE, CC0>(s, prm, e);
    case CC1: return launch_explicit_t<D, IE, CC1>(s, prm, e);
  }
  return PGPU_ERR_ARG;
}
template <int D>
static int launch_explicit_d(pgpu_species_s *s, const AdvanceParams &prm, const ExplicitArgs &e) {
  switch (s->desc.interp_E) {
    case CIC: return launch_explicit_e<D, CIC>(s, prm, e);
    case TSC: return launch_explicit_e<D, TSC>(s, prm, e);
    case CC0: return launch_explicit_e<D, CC0>(s, prm, e);
    case CC1: return launch_explicit_e<D, CC1>(s, prm, e);
  }
  return PGPU_ERR_ARG;
}
// bc: 1 = periodic in that direction (the only particle BC the fused pass applies)
int launch_explicit_step(pgpu_species_s *s, const AdvanceParams &prm, const int *periodic, bool second_half, bool deferred) {
  if (s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  ExplicitArgs e;
  const GeoAny &ga = s->grid->geo;
  for (int d = 0; d < 2; ++d) {
    e.periodic[d] = (d < ga.D) ? periodic[d] : 0;
    e.left[d] = ga.le[d];
    e.right[d] = ga.re[d];
  }
  e.second_half = second_half ? 1 : 0;
  e.deferred = deferred ? 1 : 0;
  return ga.D == 1 ? launch_explicit_d<1>(s, prm, e) : launch_explicit_d<2>(s, prm, e);
}

// ---- sub-orbit model ------------------------------------------------------------------------------------
// PicChargedSpecies::advanceSubOrbitParticlesAndSetJ (PicChargedSpecies.cpp:3376-3669), bulk container (no inflow list),
// PLANAR push: one thread per sub-orbit particle.  The reference accumulates a particle's current in a private array
// and divides it by the final number of sub-orbits; atomics cannot be taken back, so the thread first runs the
// sub-orbit loop WITHOUT depositing until it has the number of sub-orbits that converges, then runs that (deterministic)
// sequence again depositing value/nsub.
struct SubPtrs {
  double *x[2], *xold[2], *v[3], *vold[3], *w;
  int *nsub;
};
// Inflow lists (advanceInflowParticlesAndSetJ, :3255-3322 = the same loop with is_inflow_list): the particle first streams
// freely from where createInflowParticles put it, outside the domain, to the boundary plane (advanceInflowPartToBdry,
// :958-995) and takes the REST of the step in sub-orbits; one the fields turn around before it is inside is handed back
// at rest normal to the boundary (:3486-3509).  bdry < 0: the bulk sub-orbit container.
struct SubRun {
  int bdry;          // 2 * dir + side of the inflow boundary, -1 for the bulk container
  bool restarted;    // in: the sub-orbit count was raised in this call (the reference then uses cnormDt / num, :3534)
  bool reflected;    // out
};
template <int D, int IE, int IJ, bool X, bool DEP>
__device__ bool suborbit_run(const Geo<D> &g, const FieldSet &F, const CurrentSet &J, const AdvanceParams &prm,
                             int iter_max, bool from_jac, int max_sub, const double (&x0)[D], const double (&u0)[3],
                             double wp, int &nsub, double (&xp)[D], double (&vp)[3], unsigned &err, unsigned &apply,
                             SubRun &R) {
  typedef M<X> m;
  double xo[D], vo[3], xs[D];
  int num = nsub;
  double cdt = __ddiv_rn(prm.cnormDt, (double)num);
  const bool inflow = R.bdry >= 0;
  const int bdir = inflow ? (R.bdry >> 1) : 0, bside = R.bdry & 1;
#pragma unroll
  for (int d = 0; d < D; ++d) xs[d] = x0[d];
  if (inflow) {
    const double X0 = bside == 0 ? g.le[bdir] : g.re[bdir];
    const double cdt0 = __ddiv_rn(__dsub_rn(X0, x0[bdir]), u0[bdir]);
#pragma unroll
    for (int d = 0; d < D; ++d) xs[d] = (d == bdir) ? X0 : __dadd_rn(x0[d], __dmul_rn(u0[d], cdt0));
    cdt = R.restarted ? __ddiv_rn(prm.cnormDt, (double)num) : __ddiv_rn(__dsub_rn(prm.cnormDt, cdt0), (double)num);
  }
  R.reflected = false;
#pragma unroll
  for (int d = 0; d < D; ++d) xp[d] = xo[d] = xs[d];
#pragma unroll
  for (int c = 0; c < 3; ++c) vp[c] = vo[c] = u0[c];
  for (int nv = 0; nv < num; nv++) {
    int iter = 0;
    bool restart = false;
    while (true) {
      GatherOp<D, X> op(F);
      if (!gather_visit<D, IE, X>(g, xp, xo, op)) {
        err |= ERRBIT_SEGMENTS;
        return false;
      }
      if (op.oob) {
        err |= ERRBIT_BOUNDS;
        return false;
      }
      if (prm.ext.on) add_external<D>(prm.ext, xp, op.acc);
      // alpha = fnorm * cnormDt_sub / 2 (PicSpeciesUtils.cpp:20)
      boris<X>(vo, op.acc, op.acc + 3, __ddiv_rn(__dmul_rn(prm.fnorm, cdt), 2.0), true, vp, 0, 0);
      apply += 1;
      const double hdt = m::mul(cdt, 0.5);
      double dxp[D], rel_max = 0.0;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const double dxp0 = m::sub(xp[d], xo[d]);
        dxp[d] = m::mul(vp[d], hdt);
        rel_max = fmax(rel_max, __ddiv_rn(fabs(m::sub(dxp0, dxp[d])), g.dx[d]));
      }
      if (rel_max < prm.rtol) break;
#pragma unroll
      for (int d = 0; d < D; ++d) xp[d] = m::add(xo[d], dxp[d]);
      if (inflow) {
        const double xn = __dadd_rn(xo[bdir], __dmul_rn(vp[bdir], cdt));
        if ((bside == 0 && xn < g.le[bdir]) || (bside == 1 && xn > g.re[bdir])) {
          R.reflected = true;
          return true;
        }
      }
      iter += 1;
      if (iter >= iter_max) {
        if (!from_jac) {
          num++;
          if (num > max_sub) return false;
#pragma unroll
          for (int d = 0; d < D; ++d) xp[d] = xo[d] = xs[d];
#pragma unroll
          for (int c = 0; c < 3; ++c) vp[c] = vo[c] = u0[c];
          cdt = __ddiv_rn(prm.cnormDt, (double)num);
          restart = true;
        }
        break;
      }
    }
    if (restart) {
      if (DEP) return false;   // the depositing run starts from the converged count: cannot happen
      R.restarted = true;
      nv = -1;
      continue;
    }
    if (DEP) {
      const double rhop = X ? __ddiv_rn(wp, prm.volume) : wp * prm.rvolume;
      DepositOpGlobal<D, X> dop(J, false);
      // bulk: the particle's current over its number of sub-orbits; inflow: times cnormDt_sub / cnormDt (:3647-3654)
      const double scale = inflow ? __ddiv_rn(cdt, prm.cnormDt) : 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        dop.val[c] = inflow ? m::mul(m::mul(vp[c], rhop), scale) : __ddiv_rn(m::mul(vp[c], rhop), (double)num);
      if (!deposit_visit<D, IJ, X>(g, xp, xo, dop)) err |= ERRBIT_SEGMENTS;
      if (dop.oob) err |= ERRBIT_BOUNDS;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) vp[c] = m::sub(m::mul(2.0, vp[c]), vo[c]);
#pragma unroll
    for (int d = 0; d < D; ++d) xp[d] = m::sub(m::mul(2.0, xp[d]), xo[d]);
    if (nv < num - 1) {
#pragma unroll
      for (int d = 0; d < D; ++d) xo[d] = xp[d];
#pragma unroll
      for (int c = 0; c < 3; ++c) vo[c] = vp[c];
    }
  }
  nsub = num;
  return true;
}
// p.nsub[i]: the particle's number of sub-orbits; for the inflow container 8 * nsub + (2 * dir + side)
template <int D, int IE, int IJ, bool X>
__global__ void __launch_bounds__(128)
k_suborbit(SubPtrs p, long n, Geo<D> g, FieldSet F, CurrentSet J, AdvanceParams prm, int iter_max, int from_jac, int max_sub,
           Counters *cnt, unsigned *nfail, int inflow) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned err = 0, apply = 0;
  if (i < n) {
    double x0[D], u0[3], xp[D], vp[3];
#pragma unroll
    for (int d = 0; d < D; ++d) x0[d] = p.xold[d][i];
#pragma unroll
    for (int c = 0; c < 3; ++c) u0[c] = p.vold[c][i];
    const int code = p.nsub[i];
    int nsub = inflow ? (code >> 3) : code;
    SubRun R;
    R.bdry = inflow ? (code & 7) : -1;
    R.restarted = false;
    R.reflected = false;
    const double wp = p.w[i];
    bool ok = suborbit_run<D, IE, IJ, X, false>(g, F, J, prm, iter_max, from_jac != 0, max_sub, x0, u0, wp, nsub, xp, vp, err, apply, R);
    if (ok && R.reflected) {
      // x = x_old as created, u = u_old with no normal component, one sub-orbit, no current (:3497-3506)
      p.nsub[i] = 8 + R.bdry;
#pragma unroll
      for (int d = 0; d < D; ++d) p.x[d][i] = x0[d];
#pragma unroll
      for (int c = 0; c < 3; ++c) p.v[c][i] = (c == (R.bdry >> 1)) ? 0.0 : u0[c];
    } else {
      if (ok) ok = suborbit_run<D, IE, IJ, X, true>(g, F, J, prm, iter_max, from_jac != 0, max_sub, x0, u0, wp, nsub, xp, vp, err, apply, R);
      if (ok && !R.reflected) {
        p.nsub[i] = inflow ? (nsub * 8 + R.bdry) : nsub;
        // inflow particles go back time-centred against their original old state (:3608-3617): inflow_Lo/Hi form 2 x - x_old
#pragma unroll
        for (int d = 0; d < D; ++d) p.x[d][i] = inflow ? __ddiv_rn(__dadd_rn(xp[d], x0[d]), 2.0) : xp[d];
#pragma unroll
        for (int c = 0; c < 3; ++c) p.v[c][i] = inflow ? __ddiv_rn(__dadd_rn(vp[c], u0[c]), 2.0) : vp[c];
      } else {
        atomicAdd(nfail, 1u);
      }
    }
  }
  flush_counters(cnt, 0, 0, err);
}
template <int D, int IE, int IJ>
static int launch_suborbit_t(pgpu_species_s *s, const SubPtrs &P, long n, const CurrentSet &J, const AdvanceParams &prm,
                             int iter_max, int from_jac, unsigned *nfail, int inflow) {
  Context &c = ctx();
  const Geo<D> g = make_geo<D>(species_geo(s));
  const FieldSet F = grid_fields(s->grid);
  KTimer t("suborbit");
  const unsigned nb = (unsigned)((n + 127) / 128);
  if (c.exact) k_suborbit<D, IE, IJ, true><<<nb, 128, 0, c.stream>>>(P, n, g, F, J, prm, iter_max, from_jac, 512, c.d_counters, nfail, inflow);
  else k_suborbit<D, IE, IJ, false><<<nb, 128, 0, c.stream>>>(P, n, g, F, J, prm, iter_max, from_jac, 512, c.d_counters, nfail, inflow);
  return 0;
}
template <int D, int IE>
static int launch_suborbit_e(pgpu_species_s *s, const SubPtrs &P, long n, const CurrentSet &J, const AdvanceParams &prm,
                             int iter_max, int from_jac, unsigned *nfail, int inflow) {
  switch (s->desc.interp_J) {
    case CIC: return launch_suborbit_t<D, IE, CIC>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
    case TSC: return launch_suborbit_t<D, IE, TSC>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
    case CC0: return launch_suborbit_t<D, IE, CC0>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
    case CC1: return launch_suborbit_t<D, IE, CC1>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
  }
  return PGPU_ERR_ARG;
}
template <int D>
static int launch_suborbit_d(pgpu_species_s *s, const SubPtrs &P, long n, const CurrentSet &J, const AdvanceParams &prm,
                             int iter_max, int from_jac, unsigned *nfail, int inflow) {
  switch (s->desc.interp_E) {
    case CIC: return launch_suborbit_e<D, CIC>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
    case TSC: return launch_suborbit_e<D, TSC>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
    case CC0: return launch_suborbit_e<D, CC0>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
    case CC1: return launch_suborbit_e<D, CC1>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
  }
  return PGPU_ERR_ARG;
}
// the sub-orbit container of s: arrays sub[0..9] = x0 x1 xold0 xold1 v0 v1 v2 vold0 vold1 vold2, sub_w, sub_nsub; with
// inflow != 0 the inflow container inf[0..9], inf_w, inf_code instead
int launch_suborbit(pgpu_species_s *s, const AdvanceParams &prm, int from_jac, const DeviceFab *Jsub, unsigned *nfail,
                    int inflow) {
  const long n = inflow ? s->n_inf : s->n_sub;
  if (n == 0) return 0;
  double **a = inflow ? s->inf : s->sub;
  SubPtrs P;
  for (int d = 0; d < 2; ++d) {
    P.x[d] = a[d];
    P.xold[d] = a[2 + d];
  }
  for (int c = 0; c < 3; ++c) {
    P.v[c] = a[4 + c];
    P.vold[c] = a[7 + c];
  }
  P.w = inflow ? s->inf_w : s->sub_w;
  P.nsub = inflow ? s->inf_code : s->sub_nsub;
  CurrentSet J;
  for (int c = 0; c < 3; ++c) J.j[c] = Jsub[c].view();
  int iter_max = s->desc.iter_max;
  if (from_jac) iter_max += iter_max;
  return s->grid->desc.D == 1 ? launch_suborbit_d<1>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow)
                              : launch_suborbit_d<2>(s, P, n, J, prm, iter_max, from_jac, nfail, inflow);
}

int launch_advance(pgpu_species_s *s, const AdvanceParams &prm, bool fuse_deposit) {
  if (s->n == 0) return 0;
  if (materialize_old(s, KEEP_OLD_ALIASES)) return PGPU_ERR_CUDA;   // pending gathers; an aliased old group is left to the CC1 kernel
  bool deferred = false;
  if (ctx().use_fast_cc1 && !prm.ext.on) {   // the specialised CC1 kernels gather from the grid arrays only
    int fr = launch_advance_cc1_fast(s, prm, fuse_deposit);
    if (fr == 0) fr = launch_advance_cc1_1d_fast(s, prm, fuse_deposit);
    if (fr < 0) return fr;
    deferred = fr == 1;
  }
  if (materialize_old(s)) return PGPU_ERR_CUDA;         // no-op if the CC1 kernel consumed the alias
  return s->grid->desc.D == 1 ? launch_advance_d<1>(s, prm, fuse_deposit, deferred)
                              : launch_advance_d<2>(s, prm, fuse_deposit, deferred);
}

}  // namespace pgpu
