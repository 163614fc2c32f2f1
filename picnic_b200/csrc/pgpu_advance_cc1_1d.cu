// pgpu_advance_cc1_1d.cu -- fused particle-Picard advance + CC1 current deposit in ONE dimension.
//
// Same job as the 2D tile kernel (pgpu_advance_cc1.cu) for D = 1:
//   PicChargedSpecies::advanceParticlesIteratively (PicChargedSpecies.cpp:1614-1716), stepNormTransfer (:658-733),
//   cc1_1d_interpolate_fields / cc1_1d_deposit_current (MeshInterpChargeConservingF.ChF:1118-1299, 941-1109),
//   CIC of the other components at xbar (MeshInterpF.ChF:497-569), Boris (PicSpeciesUtils.cpp:8-101).
// The generic visitor kernel spends ~1200 instructions per particle on this path (general segment walk, true
// divides, bounds-checked stencil functors); on the 1e8-particle 1D deck it is issue bound at 16 % of the HBM
// roofline.  Here the common case -- the orbit xold -> xnew = 2 xbar - xold stays inside one cell of the
// half-shifted ("dual") grid, so it is a single segment with seg_factor = 1 -- is closed form:
//   d = offset from the dual-cell centre in cells, del = d_bar + 1/2 in [0, 1)
//   Ex, By, Bz (cell centred):  (1 - del) F[i0] + del F[i0+1]
//   Ey, Ez, Bx (nodal):         node pair i0+s, i0+s+1 with s = (del >= 1/2), fraction del -+ 1/2
//   Jx[i0], Jx[i0+1] += jx (1 - del), jx del;   Jy, Jz: the same nodal pair, i.e. nodes i0 .. i0+2
// Everything else (crossing a dual-cell face, stencil outside the arrays) is deferred to the generic kernel through
// the same index list as in 2D.  A thread owns 4 consecutive particles (two 128-bit loads per array), accumulates
// the 8 node contributions while the dual cell stays the same, then runs of equal cells are reduced across the
// warp with shuffles: one fp64 RED per node and run.  The 1D field arrays (a few MB) are read through L1/L2.
#include "pgpu_internal.h"
#include "pgpu_device.cuh"

namespace pgpu {
namespace {

#ifndef PGPU_1D_TP
#define PGPU_1D_TP 4
#endif
constexpr int TP1 = PGPU_1D_TP;   // consecutive particles per thread (even)
constexpr int BLOCK1 = 256;
constexpr int NS1 = 8;          // Jx(i0), Jx(i0+1), Jy(i0..i0+2), Jz(i0..i0+2)
constexpr int NOCELL = 0x7fffffff;

struct Args1D {
  const double *xo, *xb, *uo[3], *w;   // inputs (xo == xb when xold is aliased to x)
  double *xbo, *ub[3];                 // outputs
  long n;
  double le, dx, rdx, hdx;
  int i_lo, i_hi;                      // dual cells whose stencil i0 .. i0+2 lies inside every array
  const double *F[6];                  // Ex Ey Ez Bx By Bz, indexable by the global cell / node index
  double *J[3];
  double alpha, hdt, tol, rvolume;
  int iter_max;
  int suborbit;
  int *list;
  unsigned *list_count;
  Counters *cnt;
};

__device__ __forceinline__ unsigned hi_abs1(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }
constexpr unsigned HI_HALF_BAND1 = 0x3fdfffffu;   // |x| with a smaller high word is < 0.5 - 2.3e-7

__device__ __forceinline__ double rcp_ge1_1d(double den) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  double e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  return r;
}

__device__ __forceinline__ void flush8(const Args1D &A, int cell, const double (&acc)[NS1]) {
  double *jx = A.J[0] + cell, *jy = A.J[1] + cell, *jz = A.J[2] + cell;
  atomicAdd(jx, acc[0]), atomicAdd(jx + 1, acc[1]);
  atomicAdd(jy, acc[2]), atomicAdd(jy + 1, acc[3]), atomicAdd(jy + 2, acc[4]);
  atomicAdd(jz, acc[5]), atomicAdd(jz + 1, acc[6]), atomicAdd(jz + 2, acc[7]);
}

// One particle.  Returns false if it has to be deferred (nothing may be stored then).
__device__ __forceinline__ bool push_1d(const Args1D &A, double xo, double &xb, const double (&uo)[3], double (&ub)[3],
                                        int &cell, double &del_out, unsigned &apply, unsigned &unconv) {
  const double xr = __dsub_rn(xo, A.le);
  const int i0 = floor_div_fast(__dsub_rn(xr, A.hdx), A.dx, A.rdx);
  if (i0 < A.i_lo || i0 > A.i_hi) return false;
  const double dO = fma(xr, A.rdx, -(double)(i0 + 1));
  const double *ex = A.F[0] + i0, *ey = A.F[1] + i0, *ez = A.F[2] + i0;
  const double *bx = A.F[3] + i0, *by = A.F[4] + i0, *bz = A.F[5] + i0;
  // cell-centred pairs do not move during the iteration
  const double ex0 = __ldg(ex), ex1 = __ldg(ex + 1) - ex0;
  const double by0 = __ldg(by), by1 = __ldg(by + 1) - by0;
  const double bz0 = __ldg(bz), bz1 = __ldg(bz + 1) - bz0;
  int iter = 0;
  bool done = false;
  unsigned napply = 0, nunconv = 0;
  double del = 0.0;
  while (true) {
    const double dxp0 = xb - xo;
    const double dB = fma(dxp0, A.rdx, dO);
    const double dN = fma(2.0, dB, -dO);
    if (!(hi_abs1(dN) < HI_HALF_BAND1)) {
      // near (or past) a dual-cell face: the reference's own floor decides
      const double xn = fma(2.0, xb, -xo);
      const int in = floor_div_exact(__dsub_rn(__dsub_rn(xn, A.le), A.hdx), A.dx);
      if (in != i0) return false;
    }
    del = dB + 0.5;
    if (done) break;
    const bool s = del >= 0.5;
    const double f = del + (s ? -0.5 : 0.5);
    const int o = s ? 1 : 0;
    const double ey0 = __ldg(ey + o), ey1 = __ldg(ey + o + 1);
    const double ez0 = __ldg(ez + o), ez1 = __ldg(ez + o + 1);
    const double bx0 = __ldg(bx + o), bx1 = __ldg(bx + o + 1);
    const double E0 = fma(del, ex1, ex0), E1 = fma(f, ey1 - ey0, ey0), E2 = fma(f, ez1 - ez0, ez0);
    const double B0 = fma(f, bx1 - bx0, bx0), B1 = fma(del, by1, by0), B2 = fma(del, bz1, bz0);
    // Boris half step (PicSpeciesUtils.cpp:8-101)
    const double vm0 = fma(A.alpha, E0, uo[0]), vm1 = fma(A.alpha, E1, uo[1]), vm2 = fma(A.alpha, E2, uo[2]);
    const double b0 = A.alpha * B0, b1 = A.alpha * B1, b2 = A.alpha * B2;
    const double den = fma(b2, b2, fma(b1, b1, fma(b0, b0, 1.0)));
    const double p0 = fma(-vm2, b1, fma(vm1, b2, vm0));
    const double p1 = fma(-vm0, b2, fma(vm2, b0, vm1));
    const double p2 = fma(-vm1, b0, fma(vm0, b1, vm2));
    const double rden = rcp_ge1_1d(den);
    const double r0 = b0 * rden, r1 = b1 * rden, r2 = b2 * rden;
    ub[0] = fma(-p2, r1, fma(p1, r2, vm0));
    ub[1] = fma(-p0, r2, fma(p2, r0, vm1));
    ub[2] = fma(-p1, r0, fma(p0, r1, vm2));
    napply += 1;
    if (A.iter_max < 0) {   // advanceParticles (:1594-1612), part_order_swap == false
      xb = fma(ub[0], A.hdt, xo);
      done = true;
      continue;
    }
    // stepNormTransfer (:658-733): |dxp0 - dxp| / dX against rtol, as |dxp0 - dxp| against rtol*dX
    const double dxp = ub[0] * A.hdt;
    const double e0 = fabs(dxp0 - dxp);
    if (iter == 0) {
      xb = xo + dxp;
      if (!(e0 >= A.tol)) done = true;
    } else {
      if (e0 < A.tol) break;   // reverse pass: xbar unchanged, its orbit was checked above
      xb = xo + dxp;
    }
    if (!done && iter >= A.iter_max) {
      if (A.suborbit) return false;   // sub-orbit model: left to the generic kernel, which lists it
      nunconv = 1;
      done = true;
    }
    iter += 1;
  }
  cell = i0;
  del_out = del;
  apply += napply;
  unconv += nunconv;
  return true;
}

#ifndef PGPU_1D_MINB
#define PGPU_1D_MINB 2
#endif
template <bool DEP>
__global__ void __launch_bounds__(BLOCK1, PGPU_1D_MINB) k_advance_cc1_1d(const Args1D A) {
  const int lane = threadIdx.x & 31;
  const long base = ((long)blockIdx.x * BLOCK1 + threadIdx.x) * TP1;
  unsigned apply = 0, unconv = 0, defer_mask = 0;
  int acc_cell = NOCELL;
  double acc[NS1];
#pragma unroll
  for (int j = 0; j < NS1; ++j) acc[j] = 0.0;

  if (base < A.n) {
    double xo[TP1], xb[TP1], uo[3][TP1], wp[TP1];
    const bool full = base + TP1 <= A.n;
    if (full) {
      // two 128-bit loads per array
#pragma unroll
      for (int h = 0; h < TP1 / 2; ++h) {
        const double2 a = reinterpret_cast<const double2 *>(A.xo + base)[h];
        const double2 b = reinterpret_cast<const double2 *>(A.xb + base)[h];
        xo[2 * h] = a.x, xo[2 * h + 1] = a.y;
        xb[2 * h] = b.x, xb[2 * h + 1] = b.y;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double2 u = reinterpret_cast<const double2 *>(A.uo[c] + base)[h];
          uo[c][2 * h] = u.x, uo[c][2 * h + 1] = u.y;
        }
        if (DEP) {
          const double2 ww = reinterpret_cast<const double2 *>(A.w + base)[h];
          wp[2 * h] = ww.x, wp[2 * h + 1] = ww.y;
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < TP1; ++q) {
        const bool in = base + q < A.n;
        xo[q] = in ? A.xo[base + q] : 0.0;
        xb[q] = in ? A.xb[base + q] : 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) uo[c][q] = in ? A.uo[c][base + q] : 0.0;
        wp[q] = (DEP && in) ? A.w[base + q] : 0.0;
      }
    }
    double xbn[TP1], ubn[3][TP1];
#pragma unroll
    for (int q = 0; q < TP1; ++q) {
      xbn[q] = xb[q];
#pragma unroll
      for (int c = 0; c < 3; ++c) ubn[c][q] = uo[c][q];   // a deferred particle's ubar slot keeps u_old
      if (base + q >= A.n) continue;
      const double u3[3] = {uo[0][q], uo[1][q], uo[2][q]};
      double ub[3] = {0.0, 0.0, 0.0}, x = xb[q], del = 0.0;
      int cell = NOCELL;
      if (!push_1d(A, xo[q], x, u3, ub, cell, del, apply, unconv)) {
        defer_mask |= 1u << q;
        continue;
      }
      xbn[q] = x;
      ubn[0][q] = ub[0], ubn[1][q] = ub[1], ubn[2][q] = ub[2];
      if (DEP) {
        if (cell != acc_cell) {
          if (acc_cell != NOCELL) {
            flush8(A, acc_cell, acc);
#pragma unroll
            for (int j = 0; j < NS1; ++j) acc[j] = 0.0;
          }
          acc_cell = cell;
        }
        const double rhop = wp[q] * A.rvolume;
        const double jx = ub[0] * rhop, jy = ub[1] * rhop, jz = ub[2] * rhop;
        const bool s = del >= 0.5;
        const double lo = 0.5 - del, hi = del - 0.5;   // one of them is the (positive) end weight
        const double n0 = s ? 0.0 : lo, n2 = s ? hi : 0.0, n1 = s ? 1.0 - hi : 1.0 - lo;
        const double jx1 = jx * del;
        acc[0] += jx - jx1;
        acc[1] += jx1;
        acc[2] = fma(jy, n0, acc[2]), acc[3] = fma(jy, n1, acc[3]), acc[4] = fma(jy, n2, acc[4]);
        acc[5] = fma(jz, n0, acc[5]), acc[6] = fma(jz, n1, acc[6]), acc[7] = fma(jz, n2, acc[7]);
      }
    }
    // results (a deferred particle is written back unchanged: needed when the outputs are other arrays)
    if (full) {
#pragma unroll
      for (int h = 0; h < TP1 / 2; ++h) {
        reinterpret_cast<double2 *>(A.xbo + base)[h] = make_double2(xbn[2 * h], xbn[2 * h + 1]);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          reinterpret_cast<double2 *>(A.ub[c] + base)[h] = make_double2(ubn[c][2 * h], ubn[c][2 * h + 1]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < TP1; ++q)
        if (base + q < A.n) {
          A.xbo[base + q] = xbn[q];
#pragma unroll
          for (int c = 0; c < 3; ++c) A.ub[c][base + q] = ubn[c][q];
        }
    }
    if (defer_mask) {
      unsigned slot = atomicAdd(A.list_count, (unsigned)__popc(defer_mask));
#pragma unroll
      for (int q = 0; q < TP1; ++q)
        if (defer_mask & (1u << q)) A.list[slot++] = (int)(base + q);
    }
  }

  if (DEP) {
    // runs of equal cells across the warp: three shuffle steps, then one RED per node and run head
    const unsigned any = __ballot_sync(0xffffffffu, acc_cell != NOCELL);
    if (any) {
      const int prev = __shfl_up_sync(0xffffffffu, acc_cell, 1);
      const bool head = (lane == 0) || (prev != acc_cell);
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      const unsigned above = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
      const int run_end = above ? (__ffs(above) - 1) : 32;
      const int run_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
      for (int sidx = 0; sidx < 5; ++sidx) {
        const int off = 1 << sidx;
        const double take = (lane + off < run_end) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < NS1; ++j) acc[j] = fma(__shfl_down_sync(0xffffffffu, acc[j], off), take, acc[j]);
      }
      // after 5 steps the run head holds the sum over its run
      if (acc_cell != NOCELL && lane == run_start) flush8(A, acc_cell, acc);
    }
  }
  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  if (lane == 0) {
    if (apply) atomicAdd(&A.cnt->apply_its, (unsigned long long)apply);
    if (unconv) atomicAdd(&A.cnt->unconverged, (unsigned long long)unconv);
  }
}

}  // namespace

// Returns 1 if the kernel was launched (deferred particles are then in s->defer_list), 0 if this
// species / configuration is not eligible, < 0 on error.
int launch_advance_cc1_1d_fast(pgpu_species_s *s, const AdvanceParams &prm, bool deposit) {
  Context &c = ctx();
  const pgpu_grid_s *g = s->grid;
  if (c.exact || g->desc.D != 1 || s->desc.interp_E != CC1 || s->desc.relativistic) return 0;
  if (deposit && s->desc.interp_J != CC1) return 0;
  if (prm.iter_max < 0 && prm.order_swap) return 0;
  if (s->desc.bc_check_lo[0] || s->desc.bc_check_hi[0]) return 0;
  if (s->n == 0) return 1;
  if ((size_t)s->n >= (size_t)0x7fffffff) return 0;   // the defer list holds int indices
  if (!s->defer_list || s->defer_cap < (size_t)s->n) {
    if (s->defer_list) cudaFree(s->defer_list);
    if (!s->defer_count) PGPU_CUDA(cudaMalloc(&s->defer_count, sizeof(unsigned)));
    PGPU_CUDA(cudaMalloc(&s->defer_list, s->cap * sizeof(int)));
    s->defer_cap = s->cap;
  }
  PGPU_CUDA(cudaMemsetAsync(s->defer_count, 0, sizeof(unsigned), c.stream));

  Args1D A;
  const bool xa = s->xold_alias, va = s->vold_alias;   // see launch_advance_cc1_fast: old == new recorded as an alias
  A.xo = xa ? s->x[0] : s->xold[0];
  A.xb = s->x[0];
  A.xbo = xa ? s->xold[0] : s->x[0];
  for (int k = 0; k < 3; ++k) {
    A.uo[k] = va ? s->v[k] : s->vold[k];
    A.ub[k] = va ? s->vold[k] : s->v[k];
  }
  A.w = s->w;
  A.n = s->n;
  A.le = g->geo.le[0];
  A.dx = g->geo.dx[0];
  A.rdx = g->geo.rdx[0];
  A.hdx = 0.5 * g->geo.dx[0];
  int lo = -(1 << 30), hi = 1 << 30;
  if (fields_wait(g)) return PGPU_ERR_CUDA;
  for (int k = 0; k < 6; ++k) {
    const DeviceFab &f = g->field[k];
    lo = std::max(lo, f.lo[0]);
    hi = std::min(hi, f.hi[0] - 2);
    A.F[k] = f.p - f.lo[0];
  }
  for (int k = 0; k < 3; ++k) {
    const DeviceFab &f = s->J[k];
    lo = std::max(lo, f.lo[0]);
    hi = std::min(hi, f.hi[0] - 2);
    A.J[k] = f.p - f.lo[0];
  }
  if (hi < lo) return 0;
  A.i_lo = lo;
  A.i_hi = hi;
  A.alpha = prm.alpha;
  A.hdt = prm.cnormDt * 0.5;
  A.tol = prm.rtol * A.dx;
  A.rvolume = prm.rvolume;
  A.iter_max = prm.iter_max;
  A.suborbit = prm.suborbit;
  A.list = s->defer_list;
  A.list_count = s->defer_count;
  A.cnt = c.d_counters;
  const unsigned blocks = (unsigned)((s->n + (long)BLOCK1 * TP1 - 1) / ((long)BLOCK1 * TP1));
  {
    KTimer t(deposit ? "advance_cc1_1d_fused" : "advance_cc1_1d");
    if (deposit) k_advance_cc1_1d<true><<<blocks, BLOCK1, 0, c.stream>>>(A);
    else k_advance_cc1_1d<false><<<blocks, BLOCK1, 0, c.stream>>>(A);
  }
  if (xa) std::swap(s->x[0], s->xold[0]);
  if (va)
    for (int k = 0; k < 3; ++k) std::swap(s->v[k], s->vold[k]);
  s->xold_alias = s->vold_alias = false;
  return 1;
}

}  // namespace pgpu
