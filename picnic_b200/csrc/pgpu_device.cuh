// pgpu_device.cuh -- device-side shape functions of the particle engine.
//
// Gather and deposit share one set of "visitors": a visitor walks the stencil of a
// particle and calls op(component, i, j, weight) for every grid entry it touches.
// A gather op accumulates weight*F(i,j); a deposit op adds value*weight into J(i,j).
// The visitors follow the arithmetic of the reference kernels:
//   CIC/TSC : src/particle_tools/MeshInterpF.ChF:323-673
//   CC0/CC1 : src/particle_tools/MeshInterpChargeConservingF.ChF:9-2084
//
// Two arithmetic modes, selected by the template flag X:
//   X = true  ("exact"): the reference's operation order with no FMA contraction
//             and true divides -- bit-faithful to the CPU restatement.
//   X = false ("fast") : contraction allowed, l/dx evaluated as l*(1/dx).
// Cell and node INDICES are computed identically (bit-exact) in both modes: the
// fast floor falls back to the true divide whenever the product is within a few
// ulp of an integer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgpu {

enum { CIC = 0, TSC = 1, CC0 = 2, CC1 = 3 };
enum { ERRBIT_SEGMENTS = 1, ERRBIT_BOUNDS = 2 };

struct FabView {
  double *p;  // element (lo0, lo1)
  int lo0, lo1;
  int n0, n1;
};

template <int D>
struct Geo {
  double le[D], re[D], dx[D], rdx[D];
  int ghosts;
  int bc_lo[D], bc_hi[D];
};

// One fp64 reduction per distinct address of a converged warp instead of one per lane.  Cell-sorted particles
// send all 32 lanes of a warp to one or two grid entries per stencil point, and same-address atomics
// serialise in L2: on the 1D 1e8-particle deck (200 ppc) the plain version spent 3/4 of the fused kernel there.
// Only fully converged warps aggregate (the stencil loops are uniform except at cell crossings); more than four
// distinct addresses (unsorted input) fall back to one atomic per lane.
__device__ __forceinline__ void warp_aggregated_add(double *base, long idx, double val) {
  const unsigned act = __activemask();
  if (act != 0xffffffffu) {
    if (idx >= 0) atomicAdd(base + idx, val);
    return;
  }
  const int lane = threadIdx.x & 31;
  const unsigned grp = __match_any_sync(0xffffffffu, idx);
  const bool leader = (__ffs(grp) - 1) == lane;
  unsigned leaders = __ballot_sync(0xffffffffu, leader);
  if (__popc(leaders) > 4) {
    if (idx >= 0) atomicAdd(base + idx, val);
    return;
  }
  while (leaders) {
    const int src = __ffs(leaders) - 1;
    leaders &= leaders - 1;
    const unsigned g = __shfl_sync(0xffffffffu, grp, src);
    double v = ((g >> lane) & 1u) ? val : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == src && idx >= 0) atomicAdd(base + idx, v);
  }
}

// ---- external fields (EMFields::getExternalE/B; Constant.H, Cosine.H:31-45, Heavyside.H:40-54) -------------------
struct ExtFn {
  int type;
  double value, constant;
  double L[2], mode[2], phase[2];
  double C[2], A[2], X0[2], eps[2];
};
struct ExtFields {
  int on;
  ExtFn f[6];
};
template <int D>
__device__ __forceinline__ double ext_value(const ExtFn &f, const double *x) {
  if (f.type == 1) return f.value;
  if (f.type == 2) {
    const double PI = 3.14159265358979323846, TWOPI = 2.0 * PI;   // PicnicConstants.H:15-17
    double value = f.value;
#pragma unroll
    for (int dir = 0; dir < D; ++dir) {
      double arg = __dadd_rn(__ddiv_rn(__dmul_rn(__dmul_rn(TWOPI, f.mode[dir]), x[dir]), f.L[dir]), __dmul_rn(f.phase[dir], PI));
      arg = fmod(arg, TWOPI);
      value = __dmul_rn(value, cos(arg));
    }
    return __dadd_rn(value, f.constant);
  }
  if (f.type == 3) {
    double prod = 1.0;
#pragma unroll
    for (int dir = 0; dir < D; ++dir) {
      const double arg = __dsub_rn(x[dir], f.X0[dir]);
      double H = (arg < 0.0) ? 0.0 : 1.0;
      if (fabs(arg) < f.eps[dir]) H = 0.5;
      const double v = __dadd_rn(f.C[dir], __dmul_rn(f.A[dir], H));
      prod = (dir == 0) ? v : __dmul_rn(prod, v);
    }
    return prod;
  }
  return 0.0;
}
// E_p += extE(x), B_p += extB(x) (PicChargedSpecies.cpp:3984-3991); acc = E_p[3], B_p[3]
template <int D>
__device__ __forceinline__ void add_external(const ExtFields &ext, const double *x, double *acc) {
#pragma unroll
  for (int c = 0; c < 6; ++c)
    if (ext.f[c].type) acc[c] = __dadd_rn(acc[c], ext_value<D>(ext.f[c], x));
}

// ---- arithmetic helpers ------------------------------------------------------
template <bool X>
struct M;
template <>
struct M<true> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b, double) { return __ddiv_rn(a, b); }
  // acc + a*b, uncontracted
  static __device__ __forceinline__ double mad(double a, double b, double acc) {
    return __dadd_rn(acc, __dmul_rn(a, b));
  }
};
template <>
struct M<false> {
  static __device__ __forceinline__ double add(double a, double b) { return a + b; }
  static __device__ __forceinline__ double sub(double a, double b) { return a - b; }
  static __device__ __forceinline__ double mul(double a, double b) { return a * b; }
  static __device__ __forceinline__ double div(double a, double, double rb) { return a * rb; }
  static __device__ __forceinline__ double mad(double a, double b, double acc) { return fma(a, b, acc); }
};

// floor(a/dx) as the reference computes it (true divide, then floor).  Bit exact.
__device__ __forceinline__ int floor_div_exact(double a, double dx) {
  return __double2int_rd(__ddiv_rn(a, dx));
}
// Same integer, usually without the divide: a*rdx differs from RN(a/dx) by at most
// ~3 ulp, so floor() can only disagree when the product is that close to an integer.
__device__ __forceinline__ int floor_div_fast(double a, double dx, double rdx) {
  const double q = __dmul_rn(a, rdx);
  const double r = rint(q);
  if (fabs(q - r) <= fabs(q) * 1.0e-15) return floor_div_exact(a, dx);
  return __double2int_rd(q);
}
template <bool X>
__device__ __forceinline__ int floor_div(double a, double dx, double rdx) {
  if (X) return floor_div_exact(a, dx);
  return floor_div_fast(a, dx, rdx);
}

// TSC shape for r = |l/dx| (MeshInterpF.ChF:431-435)
template <bool X>
__device__ __forceinline__ double tsc_w(double r) {
  if (r < 0.5) return M<X>::sub(0.75, M<X>::mul(r, r));
  const double t = M<X>::sub(1.5, r);
  return M<X>::mul(0.5, M<X>::mul(t, t));
}

// ---- CIC ------------------------------------------------------------------------
// cic_interpolate_fields (MeshInterpF.ChF:497-569) / cic_deposit_current (:323-390).
// Components: 0..2 = E-like (Ex/Jx, Ey/Jy, Ez/Jz), 3..5 = B (only if WITH_B).
// FIRST = first E-like component visited (CC0/CC1 tails skip the in-plane ones).
template <int D, bool X, bool WITH_B, int FIRST, class Op>
__device__ __forceinline__ void cic_visit(const Geo<D> &g, const double *xp, Op &op) {
  typedef M<X> m;
  int index[D], index_stag[D];
  double w[D][2], ws[D][2];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double xr = __dsub_rn(xp[d], g.le[d]);
    const double hdx = 0.5 * g.dx[d];
    index[d] = floor_div<X>(__dsub_rn(xr, hdx), g.dx[d], g.rdx[d]);
    index_stag[d] = floor_div<X>(xr, g.dx[d], g.rdx[d]);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ii = index[d] + a;
      // l = ii*dx + 0.5*dx - xp + le
      const double l = m::add(m::sub(m::add(m::mul((double)ii, g.dx[d]), hdx), xp[d]), g.le[d]);
      const int is = index_stag[d] + a;
      const double ls = m::add(m::sub(m::mul((double)is, g.dx[d]), xp[d]), g.le[d]);
      w[d][a] = m::sub(1.0, fabs(m::div(l, g.dx[d], g.rdx[d])));
      ws[d][a] = m::sub(1.0, fabs(m::div(ls, g.dx[d], g.rdx[d])));
    }
  }
  if (D == 1) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ii = index[0] + a, is = index_stag[0] + a;
      if (FIRST <= 0) op(0, ii, 0, w[0][a]);
      if (FIRST <= 1) op(1, is, 0, ws[0][a]);
      op(2, is, 0, ws[0][a]);
      if (WITH_B) {
        op(3, is, 0, ws[0][a]);
        op(4, ii, 0, w[0][a]);
        op(5, ii, 0, w[0][a]);
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int ii = index[0] + a, is = index_stag[0] + a;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int jj = index[D - 1] + b, js = index_stag[D - 1] + b;
        if (FIRST <= 0) op(0, ii, js, m::mul(w[0][a], ws[D - 1][b]));
        if (FIRST <= 1) op(1, is, jj, m::mul(ws[0][a], w[D - 1][b]));
        op(2, is, js, m::mul(ws[0][a], ws[D - 1][b]));
        if (WITH_B) {
          op(3, is, jj, m::mul(ws[0][a], w[D - 1][b]));
          op(4, ii, js, m::mul(w[0][a], ws[D - 1][b]));
          op(5, ii, jj, m::mul(w[0][a], w[D - 1][b]));
        }
      }
    }
  }
}

// ---- TSC ------------------------------------------------------------------------
// tsc_interpolate_fields (MeshInterpF.ChF:577-673) / tsc_deposit_current (:398-489)
template <int D, bool X, bool WITH_B, class Op>
__device__ __forceinline__ void tsc_visit(const Geo<D> &g, const double *xp, Op &op) {
  typedef M<X> m;
  int index[D], index_stag[D];
  double w[D][3], ws[D][3];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double xr = __dsub_rn(xp[d], g.le[d]);
    const double hdx = 0.5 * g.dx[d];
    index[d] = floor_div<X>(__dsub_rn(xr, g.dx[d]), g.dx[d], g.rdx[d]);
    index_stag[d] = floor_div<X>(__dsub_rn(xr, hdx), g.dx[d], g.rdx[d]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ii = index[d] + a;
      const double l = m::add(m::sub(m::add(m::mul((double)ii, g.dx[d]), hdx), xp[d]), g.le[d]);
      const int is = index_stag[d] + a;
      const double ls = m::add(m::sub(m::mul((double)is, g.dx[d]), xp[d]), g.le[d]);
      w[d][a] = tsc_w<X>(fabs(m::div(l, g.dx[d], g.rdx[d])));
      ws[d][a] = tsc_w<X>(fabs(m::div(ls, g.dx[d], g.rdx[d])));
    }
  }
  if (D == 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ii = index[0] + a, is = index_stag[0] + a;
      op(0, ii, 0, w[0][a]);
      op(1, is, 0, ws[0][a]);
      op(2, is, 0, ws[0][a]);
      if (WITH_B) {
        op(3, is, 0, ws[0][a]);
        op(4, ii, 0, w[0][a]);
        op(5, ii, 0, w[0][a]);
      }
    }
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ii = index[0] + a, is = index_stag[0] + a;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int jj = index[D - 1] + b, js = index_stag[D - 1] + b;
        op(0, ii, js, m::mul(w[0][a], ws[D - 1][b]));
        op(1, is, jj, m::mul(ws[0][a], w[D - 1][b]));
        op(2, is, js, m::mul(ws[0][a], ws[D - 1][b]));
        if (WITH_B) {
          op(3, is, jj, m::mul(ws[0][a], w[D - 1][b]));
          op(4, ii, js, m::mul(w[0][a], ws[D - 1][b]));
          op(5, ii, jj, m::mul(w[0][a], w[D - 1][b]));
        }
      }
    }
  }
}

// ---- nodal CIC of the out-of-plane J at xbar (CC0/CC1 deposits) --------------------
// 1D: Jy,Jz (cc0 :131-156, cc1 :1082-1107); 2D: Jz (cc0 :344-366, cc1 :1720-1740)
template <int D, bool X, class Op>
__device__ __forceinline__ void cc_virtual_deposit_visit(const Geo<D> &g, const double *xp, Op &op) {
  typedef M<X> m;
  int is[D];
  double ws[D][2];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    is[d] = floor_div<X>(__dsub_rn(xp[d], g.le[d]), g.dx[d], g.rdx[d]);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const double l = m::add(m::sub(m::mul((double)(is[d] + a), g.dx[d]), xp[d]), g.le[d]);
      ws[d][a] = m::sub(1.0, fabs(m::div(l, g.dx[d], g.rdx[d])));
    }
  }
  if (D == 1) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      op(1, is[0] + a, 0, ws[0][a]);
      op(2, is[0] + a, 0, ws[0][a]);
    }
  } else {
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) op(2, is[0] + a, is[D - 1] + b, m::mul(ws[0][a], ws[D - 1][b]));
  }
}

// ---- boundary truncation (MeshInterpChargeConservingF.ChF:2021-2084; 1D inline) ----
template <int D, bool X>
__device__ __forceinline__ void truncate_boundaries(const Geo<D> &g, double *xpold, double *xpnew,
                                                    double slope, double slope_inv) {
  typedef M<X> m;
  if (D == 1) {
    if (g.bc_lo[0] == 1) {
      if (xpold[0] < g.le[0]) xpold[0] = g.le[0];
      if (xpnew[0] < g.le[0]) xpnew[0] = g.le[0];
    }
    if (g.bc_hi[0] == 1) {
      if (xpold[0] > g.re[0]) xpold[0] = g.re[0];
      if (xpnew[0] > g.re[0]) xpnew[0] = g.re[0];
    }
    return;
  }
  const int i0 = 0, i1 = D - 1;
  if (!(g.bc_lo[i0] | g.bc_hi[i0] | g.bc_lo[i1] | g.bc_hi[i1])) return;
  const double so0 = xpold[i0], so1 = xpold[i1];
  if (g.bc_lo[i0] == 1) {
    if (xpold[i0] < g.le[i0]) {
      xpold[i0] = g.le[i0];
      xpold[i1] = m::add(so1, m::mul(slope, m::sub(xpold[i0], so0)));
    }
    if (xpnew[i0] < g.le[i0]) {
      xpnew[i0] = g.le[i0];
      xpnew[i1] = m::add(so1, m::mul(slope, m::sub(xpnew[i0], so0)));
    }
  }
  if (g.bc_hi[i0] == 1) {
    if (xpold[i0] > g.re[i0]) {
      xpold[i0] = g.re[i0];
      xpold[i1] = m::add(so1, m::mul(slope, m::sub(xpold[i0], so0)));
    }
    if (xpnew[i0] > g.re[i0]) {
      xpnew[i0] = g.re[i0];
      xpnew[i1] = m::add(so1, m::mul(slope, m::sub(xpnew[i0], so0)));
    }
  }
  if (g.bc_lo[i1] == 1) {
    if (xpold[i1] < g.le[i1]) {
      xpold[i1] = g.le[i1];
      xpold[i0] = m::add(so0, m::mul(slope_inv, m::sub(xpold[i1], so1)));
    }
    if (xpnew[i1] < g.le[i1]) {
      xpnew[i1] = g.le[i1];
      xpnew[i0] = m::add(so0, m::mul(slope_inv, m::sub(xpnew[i1], so1)));
    }
  }
  if (g.bc_hi[i1] == 1) {
    if (xpold[i1] > g.re[i1]) {
      xpold[i1] = g.re[i1];
      xpold[i0] = m::add(so0, m::mul(slope_inv, m::sub(xpold[i1], so1)));
    }
    if (xpnew[i1] > g.re[i1]) {
      xpnew[i1] = g.re[i1];
      xpnew[i0] = m::add(so0, m::mul(slope_inv, m::sub(xpnew[i1], so1)));
    }
  }
}

// ---- CC0 / CC1 in 1D ------------------------------------------------------------
// cc0_1d_* (:9-158, :377-545), cc1_1d_* (:941-1109, :1118-1299).  DEPOSIT selects
// the last-segment guard of the deposit (:1042) vs the gather (:1222).
// Returns false when a CC1 particle needs more than ghosts+1 segments.
template <bool X, bool IS_CC1, bool DEPOSIT, class Op>
__device__ __forceinline__ bool cc_1d_inplane_visit(const Geo<1> &g, double xpold_save, double xpbar,
                                                    Op &op) {
  typedef M<X> m;
  double xpold[1] = {xpold_save};
  double xpnew[1] = {m::sub(m::mul(2.0, xpbar), xpold_save)};
  const double dXp = m::sub(xpnew[0], xpold[0]);
  int sign = 1;
  double seg_factor = 1.0;
  truncate_boundaries<1, X>(g, xpold, xpnew, 0.0, 0.0);
  const double dx = g.dx[0], hdx = 0.5 * g.dx[0];
  const double shift = IS_CC1 ? hdx : 0.0;
  const int index_old = floor_div<X>(__dsub_rn(__dsub_rn(xpold[0], g.le[0]), shift), dx, g.rdx[0]);
  const int index_new = floor_div<X>(__dsub_rn(__dsub_rn(xpnew[0], g.le[0]), shift), dx, g.rdx[0]);
  if (index_new < index_old) sign = -1;
  const int num_segments = 1 + abs(index_new - index_old);
  if (IS_CC1 && num_segments > g.ghosts + 1) return false;
  // Xcell = le + (index_old + half*(1-sign) [+ 0.5])*dx
  double cidx = __dadd_rn((double)index_old, 0.5 * (double)(1 - sign));
  if (IS_CC1) cidx = __dadd_rn(cidx, 0.5);
  double Xcell = m::add(g.le[0], m::mul(cidx, dx));
  int ii_next = index_old;
  double xpold0 = xpold[0], xpnew0 = 0.0, dXp_sub = 0.0;
  for (int nn = 0; nn < num_segments; ++nn) {
    const int ii = ii_next;
    if (nn == num_segments - 1) {
      xpnew0 = xpnew[0];
      dXp_sub = m::sub(xpnew0, xpold0);
      if (IS_CC1) {
        if (DEPOSIT) {
          if (fabs(dXp_sub) > 0.0) seg_factor = __ddiv_rn(dXp_sub, dXp);
        } else {
          if (dXp != 0.0) seg_factor = __ddiv_rn(dXp_sub, dXp);
        }
      }
    } else {
      ii_next = ii + sign;
      Xcell = m::add(Xcell, (double)sign * dx);
      xpnew0 = Xcell;
      dXp_sub = m::sub(xpnew0, xpold0);
      if (IS_CC1) seg_factor = __ddiv_rn(dXp_sub, dXp);
    }
    if (IS_CC1) {
      const double xpbar0 = m::mul(0.5, m::add(xpnew0, xpold0));
      const double l0 = m::sub(m::add(m::add(g.le[0], m::mul((double)ii, dx)), hdx), xpbar0);
      const double w0 = m::sub(1.0, fabs(m::div(l0, dx, g.rdx[0])));
      op(0, ii, 0, m::mul(w0, seg_factor));
      op(0, ii + 1, 0, m::mul(m::sub(1.0, w0), seg_factor));
    } else {
      double sf;
      if (dXp != 0.0) sf = __ddiv_rn(dXp_sub, dXp);
      else sf = 1.0;
      op(0, ii, 0, sf);
    }
    xpold0 = xpnew0;
  }
  return true;
}

// ---- CC0 / CC1 in 2D ------------------------------------------------------------
// Orbit segmentation cc0_2d_* (:204-289) / cc1_2d_* (:1524-1614) and the per-segment
// weights (cc0 :291-333 / :680-714, cc1 :1616-1709 / :1885-1960).
template <bool X, bool IS_CC1, class Op>
__device__ __forceinline__ bool cc_2d_inplane_visit(const Geo<2> &g, const double *xpold_save,
                                                    const double *xpbar, Op &op) {
  typedef M<X> m;
  double xpold[2] = {xpold_save[0], xpold_save[1]};
  double xpnew[2], dXp[2];
  int sign[2] = {1, 1};
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    xpnew[d] = m::sub(m::mul(2.0, xpbar[d]), xpold[d]);
    dXp[d] = m::sub(xpnew[d], xpold[d]);
  }
  const double slope = __ddiv_rn(dXp[1], dXp[0]);
  const double slope_inv = __ddiv_rn(1.0, slope);
  truncate_boundaries<2, X>(g, xpold, xpnew, slope, slope_inv);

  int index_old[2], cell_crossings[2];
  int num_segments = 1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double shift = IS_CC1 ? 0.5 * g.dx[d] : 0.0;
    index_old[d] = floor_div<X>(__dsub_rn(__dsub_rn(xpold[d], g.le[d]), shift), g.dx[d], g.rdx[d]);
    const int index_new = floor_div<X>(__dsub_rn(__dsub_rn(xpnew[d], g.le[d]), shift), g.dx[d], g.rdx[d]);
    if (index_new < index_old[d]) sign[d] = -1;
    cell_crossings[d] = abs(index_new - index_old[d]);
    num_segments += cell_crossings[d];
  }
  if (IS_CC1 && num_segments > g.ghosts + 1) return false;

  double Xcell[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    double cidx = __dadd_rn((double)index_old[d], 0.5 * (double)(1 - sign[d]));
    if (IS_CC1) cidx = __dadd_rn(cidx, 0.5);
    Xcell[d] = m::add(g.le[d], m::mul(cidx, g.dx[d]));
  }
  double xpold0[2] = {xpold[0], xpold[1]};
  double xpnew0[2] = {0.0, 0.0}, dXp_sub[2] = {0.0, 0.0};
  int ii_next = index_old[0], jj_next = index_old[1];

  for (int nn = 0; nn < num_segments; ++nn) {
    const int ii = ii_next, jj = jj_next;
    if (nn == num_segments - 1) {
      xpnew0[0] = xpnew[0];
      xpnew0[1] = xpnew[1];
      dXp_sub[0] = m::sub(xpnew0[0], xpold0[0]);
      dXp_sub[1] = m::sub(xpnew0[1], xpold0[1]);
    } else if (cell_crossings[0] == 0) {
      jj_next = jj + sign[1];
      Xcell[1] = m::add(Xcell[1], (double)sign[1] * g.dx[1]);
      xpnew0[1] = Xcell[1];
      dXp_sub[1] = m::sub(xpnew0[1], xpold0[1]);
      dXp_sub[0] = __dmul_rn(slope_inv, dXp_sub[1]);
      xpnew0[0] = m::add(xpold0[0], dXp_sub[0]);
    } else if (cell_crossings[1] == 0) {
      ii_next = ii + sign[0];
      Xcell[0] = m::add(Xcell[0], (double)sign[0] * g.dx[0]);
      xpnew0[0] = Xcell[0];
      dXp_sub[0] = m::sub(xpnew0[0], xpold0[0]);
      dXp_sub[1] = __dmul_rn(slope, dXp_sub[0]);
      xpnew0[1] = m::add(xpold0[1], dXp_sub[1]);
    } else {
      xpnew0[0] = m::add(Xcell[0], (double)sign[0] * g.dx[0]);
      xpnew0[1] = m::add(Xcell[1], (double)sign[1] * g.dx[1]);
      dXp_sub[0] = m::sub(xpnew0[0], xpold0[0]);
      dXp_sub[1] = m::sub(xpnew0[1], xpold0[1]);
      const double dXp_sub02 = __dmul_rn(slope_inv, dXp_sub[1]);
      if (fabs(dXp_sub[0]) < fabs(dXp_sub02)) {
        dXp_sub[1] = __dmul_rn(slope, dXp_sub[0]);
        xpnew0[1] = m::add(xpold0[1], dXp_sub[1]);
        Xcell[0] = xpnew0[0];
        ii_next = ii + sign[0];
        cell_crossings[0] -= 1;
      } else {
        dXp_sub[0] = __dmul_rn(slope_inv, dXp_sub[1]);
        xpnew0[0] = m::add(xpold0[0], dXp_sub[0]);
        Xcell[1] = xpnew0[1];
        jj_next = jj + sign[1];
        cell_crossings[1] -= 1;
      }
    }

    double seg_factor[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      if (dXp[d] != 0.0) seg_factor[d] = __ddiv_rn(dXp_sub[d], dXp[d]);
      else seg_factor[d] = 1.0;
    }
    if (!IS_CC1) {
      // l = half*(xpold0+xpnew0) - (le + idx*dx)
      const double l0 = m::sub(m::mul(0.5, m::add(xpold0[0], xpnew0[0])),
                               m::add(g.le[0], m::mul((double)ii, g.dx[0])));
      const double l1 = m::sub(m::mul(0.5, m::add(xpold0[1], xpnew0[1])),
                               m::add(g.le[1], m::mul((double)jj, g.dx[1])));
      double w0 = seg_factor[0];
      double w1 = m::sub(1.0, m::div(l1, g.dx[1], g.rdx[1]));
      op(0, ii, jj, m::mul(w0, w1));
      op(0, ii, jj + 1, m::mul(w0, m::sub(1.0, w1)));
      w0 = m::sub(1.0, m::div(l0, g.dx[0], g.rdx[0]));
      w1 = seg_factor[1];
      op(1, ii, jj, m::mul(w0, w1));
      op(1, ii + 1, jj, m::mul(m::sub(1.0, w0), w1));
    } else {
      double xpbar0[2], delta[2];
      int index_start[2];
      const int cidx[2] = {ii, jj};
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        xpbar0[d] = m::mul(0.5, m::add(xpold0[d], xpnew0[d]));
        // delta = (xpbar0 - (le + (idx+0.5)*dx))/dx
        const double xc = m::add(g.le[d], m::mul(__dadd_rn((double)cidx[d], 0.5), g.dx[d]));
        delta[d] = m::div(m::sub(xpbar0[d], xc), g.dx[d], g.rdx[d]);
        index_start[d] = floor_div<X>(__dsub_rn(__dsub_rn(xpbar0[d], g.le[d]), 0.5 * g.dx[d]),
                                      g.dx[d], g.rdx[d]);
      }
      // x current: TSC-averaged weights along y at the segment end points
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int js = index_start[1] + b;
        const double base = m::mul((double)js, g.dx[1]);
        double l = m::add(m::sub(base, xpold0[1]), g.le[1]);
        double r = fabs(m::div(l, g.dx[1], g.rdx[1]));
        double ws = tsc_w<X>(r);
        l = m::add(m::sub(base, xpnew0[1]), g.le[1]);
        r = fabs(m::div(l, g.dx[1], g.rdx[1]));
        if (r < 0.5) ws = m::sub(m::add(ws, 0.75), m::mul(r, r));
        else {
          const double t = m::sub(1.5, r);
          ws = m::add(ws, m::mul(0.5, m::mul(t, t)));
        }
        ws = m::mul(0.5, ws);
        op(0, ii, js, m::mul(m::mul(m::sub(1.0, delta[0]), ws), seg_factor[0]));
        op(0, ii + 1, js, m::mul(m::mul(delta[0], ws), seg_factor[0]));
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int is = index_start[0] + a;
        const double base = m::mul((double)is, g.dx[0]);
        double l = m::add(m::sub(base, xpold0[0]), g.le[0]);
        double r = fabs(m::div(l, g.dx[0], g.rdx[0]));
        double ws = tsc_w<X>(r);
        l = m::add(m::sub(base, xpnew0[0]), g.le[0]);
        r = fabs(m::div(l, g.dx[0], g.rdx[0]));
        if (r < 0.5) ws = m::sub(m::add(ws, 0.75), m::mul(r, r));
        else {
          const double t = m::sub(1.5, r);
          ws = m::add(ws, m::mul(0.5, m::mul(t, t)));
        }
        ws = m::mul(0.5, ws);
        op(1, is, jj, m::mul(m::mul(ws, m::sub(1.0, delta[1])), seg_factor[1]));
        op(1, is, jj + 1, m::mul(m::mul(ws, delta[1]), seg_factor[1]));
      }
    }
    xpold0[0] = xpnew0[0];
    xpold0[1] = xpnew0[1];
  }
  return true;
}

// ---- dispatch: one particle --------------------------------------------------------
// MeshInterp::interpolateEMfieldsToPart (MeshInterpI.H:537-690).  op visits E (0..2)
// and B (3..5).  Returns false on the CC1 segment-limit error.
template <int D, int INTERP, bool X, class Op>
__device__ __forceinline__ bool gather_visit(const Geo<D> &g, const double *xp, const double *xpold,
                                             Op &op) {
  bool ok = true;
  if (INTERP == CIC) {
    cic_visit<D, X, true, 0>(g, xp, op);
  } else if (INTERP == TSC) {
    tsc_visit<D, X, true>(g, xp, op);
  } else {
    if constexpr (D == 1) {
      ok = cc_1d_inplane_visit<X, INTERP == CC1, false>(g, xpold[0], xp[0], op);
      cic_visit<1, X, true, 1>(g, xp, op);
    } else {
      ok = cc_2d_inplane_visit<X, INTERP == CC1>(g, xpold, xp, op);
      cic_visit<2, X, true, 2>(g, xp, op);
    }
  }
  return ok;
}

// MeshInterp::depositCurrent (MeshInterpI.H:48-228).  op visits J (0..2) with the
// pure shape weight; the caller multiplies by v_c * (w/volume).
template <int D, int INTERP, bool X, class Op>
__device__ __forceinline__ bool deposit_visit(const Geo<D> &g, const double *xp, const double *xpold,
                                              Op &op) {
  bool ok = true;
  if (INTERP == CIC) {
    cic_visit<D, X, false, 0>(g, xp, op);
  } else if (INTERP == TSC) {
    tsc_visit<D, X, false>(g, xp, op);
  } else {
    if constexpr (D == 1) {
      ok = cc_1d_inplane_visit<X, INTERP == CC1, true>(g, xpold[0], xp[0], op);
    } else {
      ok = cc_2d_inplane_visit<X, INTERP == CC1>(g, xpold, xp, op);
    }
    cc_virtual_deposit_visit<D, X>(g, xp, op);
  }
  return ok;
}

// gamma of a proper velocity as advancePositionsExplicit writes it (PicChargedSpecies.cpp:497):
// sqrt(1 + u0 u0 + u1 u1 + u2 u2), summed left to right
template <bool X>
__device__ __forceinline__ double gamma_explicit(const double *u) {
  typedef M<X> m;
  return sqrt(m::mad(u[2], u[2], m::mad(u[1], u[1], m::mad(u[0], u[0], 1.0))));
}
// ... and as depositCurrent / setStableDt write it (MeshInterpI.H:74-76): gammap = 1; gammap += (u0 u0 + u1 u1 + u2 u2)
template <bool X>
__device__ __forceinline__ double gamma_sum_first(const double *u) {
  typedef M<X> m;
  return sqrt(m::add(1.0, m::mad(u[2], u[2], m::mad(u[1], u[1], m::mul(u[0], u[0])))));
}
// PicSpeciesUtils::getImplicitGamma (PicSpeciesUtils.H:43-52) of upold and upnew = 2 upbar - upold
template <bool X>
__device__ __forceinline__ double gamma_implicit(const double *upold, const double *upbar) {
  typedef M<X> m;
  double un[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) un[c] = m::sub(m::mul(2.0, upbar[c]), upold[c]);
  const double go = m::mad(upold[2], upold[2], m::mad(upold[1], upold[1], m::mul(upold[0], upold[0])));
  const double gn = m::mad(un[2], un[2], m::mad(un[1], un[1], m::mul(un[0], un[0])));
  return m::mul(0.5, m::add(sqrt(m::add(1.0, go)), sqrt(m::add(1.0, gn))));
}

// PicSpeciesUtils::applyForces (PicSpeciesUtils.cpp:8-101), planar; rel = the RELATIVISTIC_PARTICLES build
// (time-centred gamma of Boris or, hc, of Higuera-Cary, :55-78).
template <bool X>
__device__ __forceinline__ void boris(const double *upold, const double *Ep, const double *Bp,
                                      double alpha, bool byHalfDt, double *up, int rel = 0, int hc = 0) {
  typedef M<X> m;
  const double vm0 = m::mad(alpha, Ep[0], upold[0]);
  const double vm1 = m::mad(alpha, Ep[1], upold[1]);
  const double vm2 = m::mad(alpha, Ep[2], upold[2]);
  double bp0 = m::mul(alpha, Bp[0]);
  double bp1 = m::mul(alpha, Bp[1]);
  double bp2 = m::mul(alpha, Bp[2]);
  if (rel) {
    double root;
    if (hc) {
      const double vmsq = m::mad(vm2, vm2, m::mad(vm1, vm1, m::mul(vm0, vm0)));
      const double vmdbp = m::mad(vm2, bp2, m::mad(vm1, bp1, m::mul(vm0, bp0)));
      const double bpsq = m::mad(bp2, bp2, m::mad(bp1, bp1, m::mul(bp0, bp0)));
      const double c1 = m::sub(m::add(1.0, vmsq), bpsq);
      const double c2 = m::mad(vmdbp, vmdbp, bpsq);
      root = m::mul(0.5, m::add(c1, sqrt(m::mad(4.0, c2, m::mul(c1, c1)))));
    } else {
      root = m::mad(vm2, vm2, m::mad(vm1, vm1, m::mad(vm0, vm0, 1.0)));
    }
    const double gammap = sqrt(root);
    bp0 = __ddiv_rn(bp0, gammap);
    bp1 = __ddiv_rn(bp1, gammap);
    bp2 = __ddiv_rn(bp2, gammap);
  }
  // denom = 1 + bp0*bp0 + bp1*bp1 + bp2*bp2 (left to right)
  const double denom = m::mad(bp2, bp2, m::mad(bp1, bp1, m::mad(bp0, bp0, 1.0)));
  // vpr = vm + vm x bp
  const double vpr0 = m::sub(m::mad(vm1, bp2, vm0), m::mul(vm2, bp1));
  const double vpr1 = m::sub(m::mad(vm2, bp0, vm1), m::mul(vm0, bp2));
  const double vpr2 = m::sub(m::mad(vm0, bp1, vm2), m::mul(vm1, bp0));
  double u0, u1, u2;
  if (X) {
    u0 = m::add(vm0, __ddiv_rn(m::sub(m::mul(vpr1, bp2), m::mul(vpr2, bp1)), denom));
    u1 = m::add(vm1, __ddiv_rn(m::sub(m::mul(vpr2, bp0), m::mul(vpr0, bp2)), denom));
    u2 = m::add(vm2, __ddiv_rn(m::sub(m::mul(vpr0, bp1), m::mul(vpr1, bp0)), denom));
  } else {
    const double rden = 1.0 / denom;
    u0 = fma(vpr1 * bp2 - vpr2 * bp1, rden, vm0);
    u1 = fma(vpr2 * bp0 - vpr0 * bp2, rden, vm1);
    u2 = fma(vpr0 * bp1 - vpr1 * bp0, rden, vm2);
  }
  if (!byHalfDt) {
    u0 = m::sub(m::mul(2.0, u0), upold[0]);
    u1 = m::sub(m::mul(2.0, u1), upold[1]);
    u2 = m::sub(m::mul(2.0, u2), upold[2]);
  }
  up[0] = u0;
  up[1] = u1;
  up[2] = u2;
}

}  // namespace pgpu
