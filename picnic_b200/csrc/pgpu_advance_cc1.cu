// pgpu_advance_cc1.cu -- the headline kernel: fused implicit advance + current deposit
// for 2D CC1/CC1 species (PicChargedSpecies::advanceParticlesIteratively,
// PicChargedSpecies.cpp:1614-1716, + setCurrentDensity, :3184-3253, with the CC1 shape
// of MeshInterpChargeConservingF.ChF:1482-2013).
//
// Design (DESIGN.md "k_advance_cc1"):
//  * SoA arrays, coalesced 128-bit loads; a particle stays in registers for
//    all particle-Picard passes, so HBM sees one read of (x_old, u_old, xbar, w) and one
//    write of (xbar, ubar).
//  * FAST PATH = the orbit x_old -> x_new = 2 xbar - x_old stays inside one cell of the
//    half-shifted ("dual") grid that CC1 segments on, which is the case for all but the
//    few percent of particles that cross a dual-cell face in a step.  Then CC1 has one
//    segment with seg_factor == 1 exactly, all node indices are fixed by x_old, and the
//    weights reduce to closed forms in the normalised offsets d = xi - (i0+1), |d| <= 1/2:
//        W_0 = ((1/2-d_o)^2 + (1/2-d_n)^2)/4,  W_2 = ((1/2+d_o)^2 + (1/2+d_n)^2)/4,
//        W_1 = 1 - W_0 - W_2,   delta = dbar + 1/2          
//    The in-plane E stencil (12 values) and Bz (4 values) are loaded once per particle.
//    The same-cell test |d_n| < 1/2 is decided in normalised coordinates with a 1e-9 guard
//    band; inside the band the reference's own floor((x-le-dx/2)/dx) (true divide) decides,
//    so the segment count is the reference's.
//  * Particles that leave the fast path (a face crossing in any pass, or a stencil that
//    touches the array edge) are NOT written: their indices are appended to a deferred
//    list and the generic visitor kernel (pgpu_push.cu) redoes them from their untouched
//    state.  No CPU fallback is involved; both kernels are device code.
//  * DEPOSIT without per-particle atomics: a thread owns 2*PAIRS consecutive particles
//    (128-bit loads/stores) and adds their 21 contributions (Jx 2x3, Jy 3x2, Jz 3x3 nodes
//    relative to the dual cell) in registers while the dual cell stays the same -- the
//    cell sort orders particles by (cell, half-cell quadrant), so it nearly always does.
//    A warp then runs a segmented shuffle reduction over runs of lanes with equal dual
//    cells and issues ONE fp64 RED per (node, run): ~1 RED per particle instead of 22.
//    Correct for any particle order; only the number of REDs depends on the order.
#include <algorithm>

#include "pgpu_internal.h"

namespace pgpu {

namespace {

constexpr int NSLOT = 21;
constexpr int BLOCK = 128;
constexpr double BAND = 1.0e-9;
constexpr unsigned NOKEY = 0xffffffffu;

struct FastArgs {
  const double *xo[2];
  double *xb[2];
  const double *uo[3];
  double *ub[3];
  const double *w;
  long n;
  double le[2], dx[2], rdx[2], hdx[2];
  int i_lo[2], i_hi[2];     // dual-cell index range whose whole stencil is inside every array
  const double *F[6];       // origin-shifted: F[c][i + j*fn0[c]] is component c at global (i,j)
  int fn0[6];
  double *J[3];             // origin-shifted likewise
  int jn0[3];
  double alpha, hdt, rtol, rvolume;
  int iter_max;
  int *list;                // deferred particle indices
  unsigned *list_count;
  Counters *cnt;
};

// One particle through all its particle-Picard passes on the single-segment fast path.
// Returns false (and leaves xb/ub untouched in memory terms: the caller does not store)
// when the particle has to be redone by the generic kernel.  On success c[0..20] holds
// its current contributions relative to its dual cell `key`.
template <bool DEP>
__device__ __forceinline__ bool advance_one(const FastArgs &A, const double (&xo)[2], double (&xb)[2],
                                            const double (&uo)[3], double (&ub)[3], double wp,
                                            unsigned &apply, unsigned &unconv, unsigned &key,
                                            double (&c)[NSLOT]) {
  // dual cell of x_old (bit-exact index_old of the reference) and normalised offset
  int i0[2];
  double dO[2];
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double xr = __dsub_rn(xo[d], A.le[d]);
    i0[d] = floor_div_fast(__dsub_rn(xr, A.hdx[d]), A.dx[d], A.rdx[d]);
    dO[d] = fma(xr, A.rdx[d], -(double)(i0[d] + 1));
    if (i0[d] < A.i_lo[d] || i0[d] > A.i_hi[d]) ok = false;
  }
  if (!ok) return false;

  // ---- per-particle loads: in-plane E stencil and Bz ----------------------------------
  double ex[3], dex[3], ey[3], dey[3];
  {
    const double *p = A.F[0] + (i0[0] + i0[1] * A.fn0[0]);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double v0 = __ldg(p + b * A.fn0[0]);
      const double v1 = __ldg(p + b * A.fn0[0] + 1);
      ex[b] = v0;
      dex[b] = v1 - v0;
    }
  }
  {
    const double *p = A.F[1] + (i0[0] + i0[1] * A.fn0[1]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v0 = __ldg(p + a);
      const double v1 = __ldg(p + a + A.fn0[1]);
      ey[a] = v0;
      dey[a] = v1 - v0;
    }
  }
  double bz0, bz1, bz2, bz3;
  {
    const double *p = A.F[5] + (i0[0] + i0[1] * A.fn0[5]);
    const double v00 = __ldg(p), v10 = __ldg(p + 1);
    const double v01 = __ldg(p + A.fn0[5]), v11 = __ldg(p + A.fn0[5] + 1);
    bz0 = v00;
    bz1 = v10 - v00;
    bz2 = v01 - v00;
    bz3 = (v11 - v01) - bz1;
  }
  const double *pEz = A.F[2] + (i0[0] + i0[1] * A.fn0[2]);
  const double *pBx = A.F[3] + (i0[0] + i0[1] * A.fn0[3]);
  const double *pBy = A.F[4] + (i0[0] + i0[1] * A.fn0[4]);
  double pO[2][2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double a = 0.5 - dO[d], b = 0.5 + dO[d];
    pO[d][0] = a * a;
    pO[d][1] = b * b;
  }

  // ---- particle-Picard loop (stepNormTransfer semantics, :658-733) -----------------------
  double dB[2];
  bool recheck = true;   // xbar moved after the last gather
  int iter = 0;
  unsigned napply = 0, nunconv = 0;
  while (true) {
    double dxp0[2], dN[2];
    bool same = true;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      dxp0[d] = xb[d] - xo[d];
      dB[d] = fma(dxp0[d], A.rdx[d], dO[d]);
      dN[d] = fma(2.0, dB[d], -dO[d]);
      if (!(fabs(dN[d]) < 0.5 - BAND)) {
        // guard band: let the reference's floor decide (and catch real crossings)
        const double xn = fma(2.0, xb[d], -xo[d]);
        const int in = floor_div_exact(__dsub_rn(__dsub_rn(xn, A.le[d]), A.hdx[d]), A.dx[d]);
        if (in != i0[d]) same = false;
      }
    }
    if (!same) return false;
    const double del0 = dB[0] + 0.5, del1 = dB[1] + 0.5;
    double Wx[3], Wy[3];
    {
      const double a0 = 0.5 - dN[0], b0 = 0.5 + dN[0];
      Wx[0] = 0.25 * fma(a0, a0, pO[0][0]);
      Wx[2] = 0.25 * fma(b0, b0, pO[0][1]);
      Wx[1] = (1.0 - Wx[0]) - Wx[2];
      const double a1 = 0.5 - dN[1], b1 = 0.5 + dN[1];
      Wy[0] = 0.25 * fma(a1, a1, pO[1][0]);
      Wy[2] = 0.25 * fma(b1, b1, pO[1][1]);
      Wy[1] = (1.0 - Wy[0]) - Wy[2];
    }
    double E[3], B[3];
    E[0] = Wy[0] * fma(del0, dex[0], ex[0]);
    E[0] = fma(Wy[1], fma(del0, dex[1], ex[1]), E[0]);
    E[0] = fma(Wy[2], fma(del0, dex[2], ex[2]), E[0]);
    E[1] = Wx[0] * fma(del1, dey[0], ey[0]);
    E[1] = fma(Wx[1], fma(del1, dey[1], ey[1]), E[1]);
    E[1] = fma(Wx[2], fma(del1, dey[2], ey[2]), E[1]);
    // nodal CIC at xbar: node pair (i0+s, i0+s+1), fraction f
    const int sx = del0 >= 0.5 ? 1 : 0, sy = del1 >= 0.5 ? 1 : 0;
    const double fx = del0 + (sx ? -0.5 : 0.5), fy = del1 + (sy ? -0.5 : 0.5);
    {
      const double *p = pEz + (sx + sy * A.fn0[2]);
      const double v00 = __ldg(p), v10 = __ldg(p + 1);
      const double v01 = __ldg(p + A.fn0[2]), v11 = __ldg(p + A.fn0[2] + 1);
      const double t0 = fma(fx, v10 - v00, v00), t1 = fma(fx, v11 - v01, v01);
      E[2] = fma(fy, t1 - t0, t0);
    }
    {  // Bx: nodal in x, cell-centred in y
      const double *p = pBx + sx;
      const double v00 = __ldg(p), v10 = __ldg(p + 1);
      const double v01 = __ldg(p + A.fn0[3]), v11 = __ldg(p + A.fn0[3] + 1);
      const double t0 = fma(fx, v10 - v00, v00), t1 = fma(fx, v11 - v01, v01);
      B[0] = fma(del1, t1 - t0, t0);
    }
    {  // By: cell-centred in x, nodal in y
      const double *p = pBy + sy * A.fn0[4];
      const double v00 = __ldg(p), v10 = __ldg(p + 1);
      const double v01 = __ldg(p + A.fn0[4]), v11 = __ldg(p + A.fn0[4] + 1);
      const double t0 = fma(del0, v10 - v00, v00), t1 = fma(del0, v11 - v01, v01);
      B[1] = fma(fy, t1 - t0, t0);
    }
    B[2] = fma(del1, fma(del0, bz3, bz2), fma(del0, bz1, bz0));

    // Boris half step (PicSpeciesUtils.cpp:8-101)
    {
      const double vm0 = fma(A.alpha, E[0], uo[0]), vm1 = fma(A.alpha, E[1], uo[1]),
                   vm2 = fma(A.alpha, E[2], uo[2]);
      const double b0 = A.alpha * B[0], b1 = A.alpha * B[1], b2 = A.alpha * B[2];
      const double den = fma(b2, b2, fma(b1, b1, fma(b0, b0, 1.0)));
      const double p0 = fma(vm1, b2, vm0) - vm2 * b1;
      const double p1 = fma(vm2, b0, vm1) - vm0 * b2;
      const double p2 = fma(vm0, b1, vm2) - vm1 * b0;
      const double rden = 1.0 / den;
      ub[0] = fma(fma(p1, b2, -(p2 * b1)), rden, vm0);
      ub[1] = fma(fma(p2, b0, -(p0 * b2)), rden, vm1);
      ub[2] = fma(fma(p0, b1, -(p1 * b0)), rden, vm2);
    }
    napply += 1;
    if (A.iter_max < 0) {  // advanceParticles (:1594-1612), part_order_swap == false
      xb[0] = fma(ub[0], A.hdt, xo[0]);
      xb[1] = fma(ub[1], A.hdt, xo[1]);
      break;
    }
    const double dxp_0 = ub[0] * A.hdt, dxp_1 = ub[1] * A.hdt;
    const double rel = fmax(fabs(dxp0[0] - dxp_0) * A.rdx[0], fabs(dxp0[1] - dxp_1) * A.rdx[1]);
    if (iter == 0) {
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
      if (!(rel >= A.rtol)) break;
    } else {
      if (rel < A.rtol) {
        recheck = false;   // reverse pass: xbar is the one the weights were built from
        break;
      }
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
    }
    if (iter >= A.iter_max) {
      nunconv = 1;
      break;
    }
    iter += 1;
  }

  if (recheck) {
    // xbar changed after the last gather: the orbit must still be single-segment
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const double dxp0 = xb[d] - xo[d];
      dB[d] = fma(dxp0, A.rdx[d], dO[d]);
      const double dN = fma(2.0, dB[d], -dO[d]);
      if (!(fabs(dN) < 0.5 - BAND)) {
        const double xn = fma(2.0, xb[d], -xo[d]);
        const int in = floor_div_exact(__dsub_rn(__dsub_rn(xn, A.le[d]), A.hdx[d]), A.dx[d]);
        if (in != i0[d]) ok = false;
      }
    }
    if (!ok) return false;
  }
  apply += napply;
  unconv += nunconv;

  if (DEP) {
    key = ((unsigned)(i0[1] + 32768) << 16) | (unsigned)(i0[0] + 32768);
    const double del0 = dB[0] + 0.5, del1 = dB[1] + 0.5;
    const double dN0 = fma(2.0, dB[0], -dO[0]), dN1 = fma(2.0, dB[1], -dO[1]);
    double Wx[3], Wy[3];
    {
      const double a0 = 0.5 - dN0, b0 = 0.5 + dN0;
      Wx[0] = 0.25 * fma(a0, a0, pO[0][0]);
      Wx[2] = 0.25 * fma(b0, b0, pO[0][1]);
      Wx[1] = (1.0 - Wx[0]) - Wx[2];
      const double a1 = 0.5 - dN1, b1 = 0.5 + dN1;
      Wy[0] = 0.25 * fma(a1, a1, pO[1][0]);
      Wy[2] = 0.25 * fma(b1, b1, pO[1][1]);
      Wy[1] = (1.0 - Wy[0]) - Wy[2];
    }
    const double rhop = wp * A.rvolume;
    const double jx = ub[0] * rhop, jy = ub[1] * rhop, jz = ub[2] * rhop;
    // Jx(i0+a, j0+b), a<2, b<3  -> slot a + 2 b
    const double jx1 = jx * del0, jx0 = jx - jx1;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      c[0 + 2 * b] = jx0 * Wy[b];
      c[1 + 2 * b] = jx1 * Wy[b];
    }
    // Jy(i0+a, j0+b), a<3, b<2  -> slot 6 + a + 3 b
    const double jy1 = jy * del1, jy0 = jy - jy1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      c[6 + a] = jy0 * Wx[a];
      c[9 + a] = jy1 * Wx[a];
    }
    // Jz nodal CIC at xbar over nodes i0..i0+2 x j0..j0+2 -> slot 12 + a + 3 b
    double nx[3], ny[3];
    if (del0 >= 0.5) { nx[0] = 0.0; nx[1] = 1.5 - del0; nx[2] = del0 - 0.5; }
    else             { nx[0] = 0.5 - del0; nx[1] = del0 + 0.5; nx[2] = 0.0; }
    if (del1 >= 0.5) { ny[0] = 0.0; ny[1] = 1.5 - del1; ny[2] = del1 - 0.5; }
    else             { ny[0] = 0.5 - del1; ny[1] = del1 + 0.5; ny[2] = 0.0; }
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double t = jz * ny[b];
#pragma unroll
      for (int a = 0; a < 3; ++a) c[12 + a + 3 * b] = t * nx[a];
    }
  }
  return true;
}

// 21 fp64 REDs of one accumulator set into the J arrays at dual cell `key`
__device__ __forceinline__ void flush_direct(const FastArgs &A, unsigned key, const double (&acc)[NSLOT]) {
  const int ci = (int)(key & 0xffffu) - 32768, cj = (int)(key >> 16) - 32768;
  {
    double *p = A.J[0] + (ci + cj * A.jn0[0]);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      atomicAdd(p + b * A.jn0[0], acc[2 * b]);
      atomicAdd(p + b * A.jn0[0] + 1, acc[2 * b + 1]);
    }
  }
  {
    double *p = A.J[1] + (ci + cj * A.jn0[1]);
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int a = 0; a < 3; ++a) atomicAdd(p + a + b * A.jn0[1], acc[6 + a + 3 * b]);
  }
  {
    double *p = A.J[2] + (ci + cj * A.jn0[2]);
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int a = 0; a < 3; ++a) atomicAdd(p + a + b * A.jn0[2], acc[12 + a + 3 * b]);
  }
}

// PAIRS x 2 consecutive particles per thread (128-bit loads/stores), processed one after the
// other; their contributions accumulate in registers while the dual cell stays the same.
template <bool DEP, int PAIRS, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB) k_advance_cc1_2d(const FastArgs A) {
  constexpr int P = 2 * PAIRS;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long base = t * P;
  const int lane = threadIdx.x & 31;
  unsigned apply = 0, unconv = 0;
  unsigned acc_key = NOKEY;
  double acc[NSLOT];
#pragma unroll
  for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
  unsigned defer_mask = 0;

#pragma unroll 1
  for (int pr = 0; pr < PAIRS; ++pr) {
    const long i = base + 2 * pr;
    if (i >= A.n) break;
    const bool two = i + 1 < A.n;
    double2 xo0, xo1, xb0, xb1, u0, u1, u2, w2;
    if (two) {
      xo0 = *reinterpret_cast<const double2 *>(A.xo[0] + i);
      xo1 = *reinterpret_cast<const double2 *>(A.xo[1] + i);
      xb0 = *reinterpret_cast<const double2 *>(A.xb[0] + i);
      xb1 = *reinterpret_cast<const double2 *>(A.xb[1] + i);
      u0 = *reinterpret_cast<const double2 *>(A.uo[0] + i);
      u1 = *reinterpret_cast<const double2 *>(A.uo[1] + i);
      u2 = *reinterpret_cast<const double2 *>(A.uo[2] + i);
      w2 = *reinterpret_cast<const double2 *>(A.w + i);
    } else {
      xo0 = make_double2(A.xo[0][i], 0.0);
      xo1 = make_double2(A.xo[1][i], 0.0);
      xb0 = make_double2(A.xb[0][i], 0.0);
      xb1 = make_double2(A.xb[1][i], 0.0);
      u0 = make_double2(A.uo[0][i], 0.0);
      u1 = make_double2(A.uo[1][i], 0.0);
      u2 = make_double2(A.uo[2][i], 0.0);
      w2 = make_double2(A.w[i], 0.0);
    }
    double2 ob0 = xb0, ob1 = xb1, ov0 = u0, ov1 = u1, ov2 = u2;   // outputs (xbar, ubar)
    bool okA = false, okB = false;
    {
      const double xo[2] = {xo0.x, xo1.x};
      double xb[2] = {xb0.x, xb1.x};
      const double uo[3] = {u0.x, u1.x, u2.x};
      double ub[3] = {0.0, 0.0, 0.0};
      unsigned key = NOKEY;
      double c[NSLOT];
      okA = advance_one<DEP>(A, xo, xb, uo, ub, w2.x, apply, unconv, key, c);
      if (okA) {
        ob0.x = xb[0]; ob1.x = xb[1]; ov0.x = ub[0]; ov1.x = ub[1]; ov2.x = ub[2];
        if (DEP) {
          if (key != acc_key) {
            if (acc_key != NOKEY) flush_direct(A, acc_key, acc);
            acc_key = key;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] = c[j];
          } else {
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] += c[j];
          }
        }
      } else {
        defer_mask |= 1u << (2 * pr);
      }
    }
    if (two) {
      const double xo[2] = {xo0.y, xo1.y};
      double xb[2] = {xb0.y, xb1.y};
      const double uo[3] = {u0.y, u1.y, u2.y};
      double ub[3] = {0.0, 0.0, 0.0};
      unsigned key = NOKEY;
      double c[NSLOT];
      okB = advance_one<DEP>(A, xo, xb, uo, ub, w2.y, apply, unconv, key, c);
      if (okB) {
        ob0.y = xb[0]; ob1.y = xb[1]; ov0.y = ub[0]; ov1.y = ub[1]; ov2.y = ub[2];
        if (DEP) {
          if (key != acc_key) {
            if (acc_key != NOKEY) flush_direct(A, acc_key, acc);
            acc_key = key;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] = c[j];
          } else {
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] += c[j];
          }
        }
      } else {
        defer_mask |= 1u << (2 * pr + 1);
      }
    }
    // a deferred particle keeps its stored (xbar, ubar): the generic kernel restarts from them
    if (two && okA && okB) {
      *reinterpret_cast<double2 *>(A.xb[0] + i) = ob0;
      *reinterpret_cast<double2 *>(A.xb[1] + i) = ob1;
      *reinterpret_cast<double2 *>(A.ub[0] + i) = ov0;
      *reinterpret_cast<double2 *>(A.ub[1] + i) = ov1;
      *reinterpret_cast<double2 *>(A.ub[2] + i) = ov2;
    } else {
      if (okA) {
        A.xb[0][i] = ob0.x; A.xb[1][i] = ob1.x;
        A.ub[0][i] = ov0.x; A.ub[1][i] = ov1.x; A.ub[2][i] = ov2.x;
      }
      if (okB) {
        A.xb[0][i + 1] = ob0.y; A.xb[1][i + 1] = ob1.y;
        A.ub[0][i + 1] = ov0.y; A.ub[1][i + 1] = ov1.y; A.ub[2][i + 1] = ov2.y;
      }
    }
  }

  // ---- deferred list ---------------------------------------------------------------------
  if (defer_mask) {
    unsigned slot = atomicAdd(A.list_count, (unsigned)__popc(defer_mask));
#pragma unroll
    for (int q = 0; q < P; ++q)
      if (defer_mask & (1u << q)) A.list[slot++] = (int)(base + q);
  }

  // ---- deposit: segmented warp reduction over runs of equal dual cells, one RED per node
  //      and run (the particle arrays are cell sorted, so a warp holds a handful of runs)
  if (DEP) {
    const unsigned any = __ballot_sync(0xffffffffu, acc_key != NOKEY);
    if (any) {
      // runs = maximal stretches of consecutive lanes with the same key (any particle order
      // is handled: a key that reappears later simply forms another run)
      const unsigned prev = __shfl_up_sync(0xffffffffu, acc_key, 1);
      const bool head = (lane == 0) || (prev != acc_key);
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      const unsigned above = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
      const int run_end = above ? (__ffs(above) - 1) : 32;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const bool take = lane + off < run_end;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const double v = __shfl_down_sync(0xffffffffu, acc[j], off);
          if (take) acc[j] += v;
        }
      }
      if (head && acc_key != NOKEY) flush_direct(A, acc_key, acc);
    }
  }

  // ---- counters ----------------------------------------------------------------------------
  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  if (lane == 0) {
    if (apply) atomicAdd(&A.cnt->apply_its, (unsigned long long)apply);
    if (unconv) atomicAdd(&A.cnt->unconverged, (unsigned long long)unconv);
  }
}


// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UBLKCP, SYNCS) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy (contiguous bytes, 16-byte aligned and sized), completion on mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int TP = 4;                 // consecutive particles per thread
constexpr int TILE = BLOCK * TP;      // particles per tile
constexpr int NIN = 8;                // xo0 xo1 xb0 xb1 uo0 uo1 uo2 w
constexpr int STAGES = 2;
constexpr size_t TMA_SMEM = (size_t)STAGES * NIN * TILE * sizeof(double) + 64;

// Persistent, double-buffered version: each block walks tiles of TILE consecutive particles.
// One thread issues eight 1D TMA bulk copies per tile (the SoA slices of that tile) into a
// shared-memory stage and arms an mbarrier; while the block computes tile k the copies of
// tile k+1 are in flight, so no warp ever waits on HBM latency and the particle data never
// occupies registers before it is used.  Results (xbar in place, ubar over the u_old slots)
// leave through TMA bulk stores.  Thread t owns particles 4t..4t+3 of the tile (visited in a
// lane-rotated order so that the 8-byte shared loads are bank-conflict free).
template <bool DEP>
__global__ void __launch_bounds__(BLOCK, 3) k_advance_cc1_2d_tma(const FastArgs A, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);                       // [STAGES][NIN][TILE]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sm + STAGES * NIN * TILE);  // [STAGES]
  const int tid = threadIdx.x, lane = tid & 31;
  const double *src[NIN] = {A.xo[0], A.xo[1], A.xb[0], A.xb[1], A.uo[0], A.uo[1], A.uo[2], A.w};

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue_load = [&](int tile, int stage) {
    mbar_expect_tx(&bars[stage], NIN * TILE * (unsigned)sizeof(double));
    const long off = (long)tile * TILE;
#pragma unroll
    for (int a = 0; a < NIN; ++a)
      bulk_g2s(sm + ((size_t)stage * NIN + a) * TILE, src[a] + off, TILE * (unsigned)sizeof(double), &bars[stage]);
  };
  int tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) issue_load(tile, 0);

  unsigned apply = 0, unconv = 0;
  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int stage = it & 1;
    const unsigned parity = (it >> 1) & 1;
    if (tid == 0) {
      const int next = tile + gridDim.x;
      if (next < ntiles) {
        bulk_wait_read0();   // the stores that read the other stage have drained
        issue_load(next, stage ^ 1);
      }
    }
    mbar_wait(&bars[stage], parity);
    double *st = sm + (size_t)stage * NIN * TILE;
    const long tbase = (long)tile * TILE;
    const int nvalid = (A.n - tbase) < TILE ? (int)(A.n - tbase) : TILE;

    unsigned acc_key = NOKEY;
    double acc[NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
    unsigned defer_mask = 0;
#pragma unroll 1
    for (int qq = 0; qq < TP; ++qq) {
      const int q = (qq + (lane >> 2)) & (TP - 1);
      const int k = tid * TP + q;
      if (k >= nvalid) continue;
      const double xo[2] = {st[0 * TILE + k], st[1 * TILE + k]};
      double xb[2] = {st[2 * TILE + k], st[3 * TILE + k]};
      const double uo[3] = {st[4 * TILE + k], st[5 * TILE + k], st[6 * TILE + k]};
      const double wp = st[7 * TILE + k];
      double ub[3] = {0.0, 0.0, 0.0};
      unsigned key = NOKEY;
      double c[NSLOT];
      if (advance_one<DEP>(A, xo, xb, uo, ub, wp, apply, unconv, key, c)) {
        st[2 * TILE + k] = xb[0];
        st[3 * TILE + k] = xb[1];
        st[4 * TILE + k] = ub[0];
        st[5 * TILE + k] = ub[1];
        st[6 * TILE + k] = ub[2];
        if (DEP) {
          if (key != acc_key) {
            if (acc_key != NOKEY) flush_direct(A, acc_key, acc);
            acc_key = key;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] = c[j];
          } else {
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] += c[j];
          }
        }
      } else {
        // deferred: xbar stays as stored (the generic kernel restarts from it); the ubar slot
        // keeps u_old, which that kernel overwrites
        defer_mask |= 1u << q;
      }
    }
    // results -> global
    if (nvalid == TILE) {
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int a = 0; a < 2; ++a) bulk_s2g(A.xb[a] + tbase, st + (2 + a) * TILE, TILE * (unsigned)sizeof(double));
#pragma unroll
        for (int a = 0; a < 3; ++a) bulk_s2g(A.ub[a] + tbase, st + (4 + a) * TILE, TILE * (unsigned)sizeof(double));
        bulk_commit();
      }
    } else {
      // ragged last tile: plain stores of the valid particles only
#pragma unroll 1
      for (int q = 0; q < TP; ++q) {
        const int k = tid * TP + q;
        if (k < nvalid && !(defer_mask & (1u << q))) {
          A.xb[0][tbase + k] = st[2 * TILE + k];
          A.xb[1][tbase + k] = st[3 * TILE + k];
          A.ub[0][tbase + k] = st[4 * TILE + k];
          A.ub[1][tbase + k] = st[5 * TILE + k];
          A.ub[2][tbase + k] = st[6 * TILE + k];
        }
      }
      __syncthreads();
    }
    if (defer_mask) {
      unsigned slot = atomicAdd(A.list_count, (unsigned)__popc(defer_mask));
#pragma unroll
      for (int q = 0; q < TP; ++q)
        if (defer_mask & (1u << q)) A.list[slot++] = (int)(tbase + tid * TP + q);
    }
    if (DEP) {
      const unsigned any = __ballot_sync(0xffffffffu, acc_key != NOKEY);
      if (any) {
        const unsigned prev = __shfl_up_sync(0xffffffffu, acc_key, 1);
        const bool head = (lane == 0) || (prev != acc_key);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned above = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
        const int run_end = above ? (__ffs(above) - 1) : 32;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const bool take = lane + off < run_end;
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) {
            const double v = __shfl_down_sync(0xffffffffu, acc[j], off);
            if (take) acc[j] += v;
          }
        }
        if (head && acc_key != NOKEY) flush_direct(A, acc_key, acc);
      }
    }
  }
  if (tid == 0) bulk_wait0();   // all bulk stores complete before the block retires

  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  if (lane == 0) {
    if (apply) atomicAdd(&A.cnt->apply_its, (unsigned long long)apply);
    if (unconv) atomicAdd(&A.cnt->unconverged, (unsigned long long)unconv);
  }
}


// =============================================================================================
// Two-phase tile kernel.  Phase 1 pushes the tile's particles (all particle-Picard passes) and
// leaves (xbar, ubar) in the shared-memory tile; phase 2 re-reads them and builds the current
// contributions.  Splitting the phases keeps the 21 register accumulators out of the Picard
// loop's live range, which is what lets four blocks (16 warps) share an SM.
// =============================================================================================
struct CellRef {
  int i0[2];
  double dO[2];
};

__device__ __forceinline__ bool locate_dual(const FastArgs &A, const double (&xo)[2], CellRef &r) {
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double xr = __dsub_rn(xo[d], A.le[d]);
    r.i0[d] = floor_div_fast(__dsub_rn(xr, A.hdx[d]), A.dx[d], A.rdx[d]);
    r.dO[d] = fma(xr, A.rdx[d], -(double)(r.i0[d] + 1));
    if (r.i0[d] < A.i_lo[d] || r.i0[d] > A.i_hi[d]) ok = false;
  }
  return ok;
}

// normalised offsets of xbar and x_new; false if the orbit leaves the dual cell of x_old
__device__ __forceinline__ bool orbit_offsets(const FastArgs &A, const CellRef &r, const double (&xo)[2],
                                              const double (&xb)[2], double (&dxp0)[2], double (&dB)[2],
                                              double (&dN)[2]) {
  bool same = true;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    dxp0[d] = xb[d] - xo[d];
    dB[d] = fma(dxp0[d], A.rdx[d], r.dO[d]);
    dN[d] = fma(2.0, dB[d], -r.dO[d]);
    if (!(fabs(dN[d]) < 0.5 - BAND)) {
      // guard band: let the reference's floor decide (and catch real crossings)
      const double xn = fma(2.0, xb[d], -xo[d]);
      const int in = floor_div_exact(__dsub_rn(__dsub_rn(xn, A.le[d]), A.hdx[d]), A.dx[d]);
      if (in != r.i0[d]) same = false;
    }
  }
  return same;
}

// Phase 1: returns false if the particle must be redone by the generic kernel.
__device__ __forceinline__ bool push_one(const FastArgs &A, const double (&xo)[2], double (&xb)[2],
                                         const double (&uo)[3], double (&ub)[3], unsigned &apply,
                                         unsigned &unconv) {
  CellRef r;
  if (!locate_dual(A, xo, r)) return false;
  const int(&i0)[2] = r.i0;
  const double(&dO)[2] = r.dO;
  double ex[3], dex[3], ey[3], dey[3];
  {
    const double *p = A.F[0] + (i0[0] + i0[1] * A.fn0[0]);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double v0 = __ldg(p + b * A.fn0[0]);
      const double v1 = __ldg(p + b * A.fn0[0] + 1);
      ex[b] = v0;
      dex[b] = v1 - v0;
    }
  }
  {
    const double *p = A.F[1] + (i0[0] + i0[1] * A.fn0[1]);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v0 = __ldg(p + a);
      const double v1 = __ldg(p + a + A.fn0[1]);
      ey[a] = v0;
      dey[a] = v1 - v0;
    }
  }
  double bz0, bz1, bz2, bz3;
  {
    const double *p = A.F[5] + (i0[0] + i0[1] * A.fn0[5]);
    const double v00 = __ldg(p), v10 = __ldg(p + 1);
    const double v01 = __ldg(p + A.fn0[5]), v11 = __ldg(p + A.fn0[5] + 1);
    bz0 = v00;
    bz1 = v10 - v00;
    bz2 = v01 - v00;
    bz3 = (v11 - v01) - bz1;
  }
  const int oEz = i0[0] + i0[1] * A.fn0[2];
  const int oBx = i0[0] + i0[1] * A.fn0[3];
  const int oBy = i0[0] + i0[1] * A.fn0[4];
  double pO[2][2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double a = 0.5 - dO[d], b = 0.5 + dO[d];
    pO[d][0] = a * a;
    pO[d][1] = b * b;
  }

  int iter = 0;
  bool done = false;
  unsigned napply = 0, nunconv = 0;
  while (true) {
    double dxp0[2], dB[2], dN[2];
    if (!orbit_offsets(A, r, xo, xb, dxp0, dB, dN)) return false;
    if (done) break;
    const double del0 = dB[0] + 0.5, del1 = dB[1] + 0.5;
    double Wx[3], Wy[3];
    {
      const double a0 = 0.5 - dN[0], b0 = 0.5 + dN[0];
      Wx[0] = 0.25 * fma(a0, a0, pO[0][0]);
      Wx[2] = 0.25 * fma(b0, b0, pO[0][1]);
      Wx[1] = (1.0 - Wx[0]) - Wx[2];
      const double a1 = 0.5 - dN[1], b1 = 0.5 + dN[1];
      Wy[0] = 0.25 * fma(a1, a1, pO[1][0]);
      Wy[2] = 0.25 * fma(b1, b1, pO[1][1]);
      Wy[1] = (1.0 - Wy[0]) - Wy[2];
    }
    double E[3], B[3];
    E[0] = Wy[0] * fma(del0, dex[0], ex[0]);
    E[0] = fma(Wy[1], fma(del0, dex[1], ex[1]), E[0]);
    E[0] = fma(Wy[2], fma(del0, dex[2], ex[2]), E[0]);
    E[1] = Wx[0] * fma(del1, dey[0], ey[0]);
    E[1] = fma(Wx[1], fma(del1, dey[1], ey[1]), E[1]);
    E[1] = fma(Wx[2], fma(del1, dey[2], ey[2]), E[1]);
    const int sx = del0 >= 0.5 ? 1 : 0, sy = del1 >= 0.5 ? 1 : 0;
    const double fx = del0 + (sx ? -0.5 : 0.5), fy = del1 + (sy ? -0.5 : 0.5);
    {
      const double *p = A.F[2] + (oEz + sx + sy * A.fn0[2]);
      const double v00 = __ldg(p), v10 = __ldg(p + 1);
      const double v01 = __ldg(p + A.fn0[2]), v11 = __ldg(p + A.fn0[2] + 1);
      const double t0 = fma(fx, v10 - v00, v00), t1 = fma(fx, v11 - v01, v01);
      E[2] = fma(fy, t1 - t0, t0);
    }
    {
      const double *p = A.F[3] + (oBx + sx);
      const double v00 = __ldg(p), v10 = __ldg(p + 1);
      const double v01 = __ldg(p + A.fn0[3]), v11 = __ldg(p + A.fn0[3] + 1);
      const double t0 = fma(fx, v10 - v00, v00), t1 = fma(fx, v11 - v01, v01);
      B[0] = fma(del1, t1 - t0, t0);
    }
    {
      const double *p = A.F[4] + (oBy + sy * A.fn0[4]);
      const double v00 = __ldg(p), v10 = __ldg(p + 1);
      const double v01 = __ldg(p + A.fn0[4]), v11 = __ldg(p + A.fn0[4] + 1);
      const double t0 = fma(del0, v10 - v00, v00), t1 = fma(del0, v11 - v01, v01);
      B[1] = fma(fy, t1 - t0, t0);
    }
    B[2] = fma(del1, fma(del0, bz3, bz2), fma(del0, bz1, bz0));
    {
      const double vm0 = fma(A.alpha, E[0], uo[0]), vm1 = fma(A.alpha, E[1], uo[1]),
                   vm2 = fma(A.alpha, E[2], uo[2]);
      const double b0 = A.alpha * B[0], b1 = A.alpha * B[1], b2 = A.alpha * B[2];
      const double den = fma(b2, b2, fma(b1, b1, fma(b0, b0, 1.0)));
      const double p0 = fma(vm1, b2, vm0) - vm2 * b1;
      const double p1 = fma(vm2, b0, vm1) - vm0 * b2;
      const double p2 = fma(vm0, b1, vm2) - vm1 * b0;
      const double rden = 1.0 / den;
      ub[0] = fma(fma(p1, b2, -(p2 * b1)), rden, vm0);
      ub[1] = fma(fma(p2, b0, -(p0 * b2)), rden, vm1);
      ub[2] = fma(fma(p0, b1, -(p1 * b0)), rden, vm2);
    }
    napply += 1;
    if (A.iter_max < 0) {  // advanceParticles (:1594-1612), part_order_swap == false
      xb[0] = fma(ub[0], A.hdt, xo[0]);
      xb[1] = fma(ub[1], A.hdt, xo[1]);
      done = true;
      continue;
    }
    const double dxp_0 = ub[0] * A.hdt, dxp_1 = ub[1] * A.hdt;
    const double rel = fmax(fabs(dxp0[0] - dxp_0) * A.rdx[0], fabs(dxp0[1] - dxp_1) * A.rdx[1]);
    if (iter == 0) {
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
      if (!(rel >= A.rtol)) done = true;
    } else {
      if (rel < A.rtol) break;  // reverse pass: xbar unchanged, its orbit was checked above
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
    }
    if (!done && iter >= A.iter_max) {
      nunconv = 1;
      done = true;
    }
    iter += 1;
  }
  apply += napply;
  unconv += nunconv;
  return true;
}

// Phase 2: the 21 node contributions of a pushed, single-segment particle.
__device__ __forceinline__ void deposit_one(const FastArgs &A, const double (&xo)[2], const double (&xb)[2],
                                            const double (&ub)[3], double wp, unsigned &key,
                                            double (&c)[NSLOT]) {
  CellRef r;
  locate_dual(A, xo, r);
  key = ((unsigned)(r.i0[1] + 32768) << 16) | (unsigned)(r.i0[0] + 32768);
  double dB[2], dN[2], Wx[3], Wy[3];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    dB[d] = fma(xb[d] - xo[d], A.rdx[d], r.dO[d]);
    dN[d] = fma(2.0, dB[d], -r.dO[d]);
  }
  {
    const double a0 = 0.5 - dN[0], b0 = 0.5 + dN[0], c0 = 0.5 - r.dO[0], e0 = 0.5 + r.dO[0];
    Wx[0] = 0.25 * fma(a0, a0, c0 * c0);
    Wx[2] = 0.25 * fma(b0, b0, e0 * e0);
    Wx[1] = (1.0 - Wx[0]) - Wx[2];
    const double a1 = 0.5 - dN[1], b1 = 0.5 + dN[1], c1 = 0.5 - r.dO[1], e1 = 0.5 + r.dO[1];
    Wy[0] = 0.25 * fma(a1, a1, c1 * c1);
    Wy[2] = 0.25 * fma(b1, b1, e1 * e1);
    Wy[1] = (1.0 - Wy[0]) - Wy[2];
  }
  const double del0 = dB[0] + 0.5, del1 = dB[1] + 0.5;
  const double rhop = wp * A.rvolume;
  const double jx = ub[0] * rhop, jy = ub[1] * rhop, jz = ub[2] * rhop;
  const double jx1 = jx * del0, jx0 = jx - jx1;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    c[0 + 2 * b] = jx0 * Wy[b];
    c[1 + 2 * b] = jx1 * Wy[b];
  }
  const double jy1 = jy * del1, jy0 = jy - jy1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    c[6 + a] = jy0 * Wx[a];
    c[9 + a] = jy1 * Wx[a];
  }
  double nx[3], ny[3];
  if (del0 >= 0.5) { nx[0] = 0.0; nx[1] = 1.5 - del0; nx[2] = del0 - 0.5; }
  else             { nx[0] = 0.5 - del0; nx[1] = del0 + 0.5; nx[2] = 0.0; }
  if (del1 >= 0.5) { ny[0] = 0.0; ny[1] = 1.5 - del1; ny[2] = del1 - 0.5; }
  else             { ny[0] = 0.5 - del1; ny[1] = del1 + 0.5; ny[2] = 0.0; }
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const double t = jz * ny[b];
#pragma unroll
    for (int a = 0; a < 3; ++a) c[12 + a + 3 * b] = t * nx[a];
  }
}

constexpr size_t TILE2_SMEM = (size_t)NIN * TILE * sizeof(double) + 64;

template <bool DEP, int RSTEPS>
__global__ void __launch_bounds__(BLOCK, 4) k_advance_cc1_2d_tile(const FastArgs A, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *st = reinterpret_cast<double *>(smem_raw);                 // [NIN][TILE]
  uint64_t *bar = reinterpret_cast<uint64_t *>(st + NIN * TILE);
  const int tid = threadIdx.x, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  unsigned apply = 0, unconv = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const long tbase = (long)tile * TILE;
    if (tid == 0) {
      bulk_wait_read0();   // the previous tile's stores have finished reading the buffer
      mbar_expect_tx(bar, NIN * TILE * (unsigned)sizeof(double));
      constexpr unsigned BYTES = TILE * (unsigned)sizeof(double);
      bulk_g2s(st + 0 * TILE, A.xo[0] + tbase, BYTES, bar);
      bulk_g2s(st + 1 * TILE, A.xo[1] + tbase, BYTES, bar);
      bulk_g2s(st + 2 * TILE, A.xb[0] + tbase, BYTES, bar);
      bulk_g2s(st + 3 * TILE, A.xb[1] + tbase, BYTES, bar);
      bulk_g2s(st + 4 * TILE, A.uo[0] + tbase, BYTES, bar);
      bulk_g2s(st + 5 * TILE, A.uo[1] + tbase, BYTES, bar);
      bulk_g2s(st + 6 * TILE, A.uo[2] + tbase, BYTES, bar);
      bulk_g2s(st + 7 * TILE, A.w + tbase, BYTES, bar);
    }
    mbar_wait(bar, (unsigned)(it & 1));
    const int nvalid = (A.n - tbase) < TILE ? (int)(A.n - tbase) : TILE;

    // ---- phase 1: push ------------------------------------------------------------------
    unsigned defer_mask = 0;
#pragma unroll 1
    for (int qq = 0; qq < TP; ++qq) {
      const int q = (qq + (lane >> 2)) & (TP - 1);
      const int k = tid * TP + q;
      if (k >= nvalid) {
        defer_mask |= 16u << q;   // not a particle
        continue;
      }
      const double xo[2] = {st[0 * TILE + k], st[1 * TILE + k]};
      double xb[2] = {st[2 * TILE + k], st[3 * TILE + k]};
      const double uo[3] = {st[4 * TILE + k], st[5 * TILE + k], st[6 * TILE + k]};
      double ub[3] = {0.0, 0.0, 0.0};
      if (push_one(A, xo, xb, uo, ub, apply, unconv)) {
        st[2 * TILE + k] = xb[0];
        st[3 * TILE + k] = xb[1];
        st[4 * TILE + k] = ub[0];
        st[5 * TILE + k] = ub[1];
        st[6 * TILE + k] = ub[2];
      } else {
        // deferred: xbar stays as stored (the generic kernel restarts from it); the ubar slot
        // keeps u_old, which that kernel overwrites
        defer_mask |= 1u << q;
      }
    }

    // ---- phase 2: deposit ---------------------------------------------------------------
    if (DEP) {
      unsigned acc_key = NOKEY;
      double acc[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
#pragma unroll 1
      for (int qq = 0; qq < TP; ++qq) {
        const int q = (qq + (lane >> 2)) & (TP - 1);
        if (defer_mask & (17u << q)) continue;
        const int k = tid * TP + q;
        const double xo[2] = {st[0 * TILE + k], st[1 * TILE + k]};
        const double xb[2] = {st[2 * TILE + k], st[3 * TILE + k]};
        const double ub[3] = {st[4 * TILE + k], st[5 * TILE + k], st[6 * TILE + k]};
        const double wp = st[7 * TILE + k];
        unsigned key;
        double c[NSLOT];
        deposit_one(A, xo, xb, ub, wp, key, c);
        if (key != acc_key) {
          if (acc_key != NOKEY) flush_direct(A, acc_key, acc);
          acc_key = key;
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) acc[j] = c[j];
        } else {
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) acc[j] += c[j];
        }
      }
      const unsigned any = __ballot_sync(0xffffffffu, acc_key != NOKEY);
      if (any) {
        const unsigned prev = __shfl_up_sync(0xffffffffu, acc_key, 1);
        const bool head = (lane == 0) || (prev != acc_key);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned above = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
        const int run_end = above ? (__ffs(above) - 1) : 32;
        const int run_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int sidx = 0; sidx < RSTEPS; ++sidx) {
          const int off = 1 << sidx;
          const bool take = lane + off < run_end;
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) {
            const double v = __shfl_down_sync(0xffffffffu, acc[j], off);
            if (take) acc[j] += v;
          }
        }
        // after RSTEPS steps lane l holds the sum over [l, min(l + 2^RSTEPS, run_end))
        if (acc_key != NOKEY && (((lane - run_start) & ((1 << RSTEPS) - 1)) == 0)) flush_direct(A, acc_key, acc);
      }
    }

    // ---- results -> global ----------------------------------------------------------------
    defer_mask &= 15u;
    if (nvalid == TILE) {
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) {
        constexpr unsigned BYTES = TILE * (unsigned)sizeof(double);
        bulk_s2g(A.xb[0] + tbase, st + 2 * TILE, BYTES);
        bulk_s2g(A.xb[1] + tbase, st + 3 * TILE, BYTES);
        bulk_s2g(A.ub[0] + tbase, st + 4 * TILE, BYTES);
        bulk_s2g(A.ub[1] + tbase, st + 5 * TILE, BYTES);
        bulk_s2g(A.ub[2] + tbase, st + 6 * TILE, BYTES);
        bulk_commit();
      }
    } else {
#pragma unroll 1
      for (int q = 0; q < TP; ++q) {
        const int k = tid * TP + q;
        if (k < nvalid && !(defer_mask & (1u << q))) {
          A.xb[0][tbase + k] = st[2 * TILE + k];
          A.xb[1][tbase + k] = st[3 * TILE + k];
          A.ub[0][tbase + k] = st[4 * TILE + k];
          A.ub[1][tbase + k] = st[5 * TILE + k];
          A.ub[2][tbase + k] = st[6 * TILE + k];
        }
      }
      __syncthreads();
    }
    if (defer_mask) {
      unsigned slot = atomicAdd(A.list_count, (unsigned)__popc(defer_mask));
#pragma unroll
      for (int q = 0; q < TP; ++q)
        if (defer_mask & (1u << q)) A.list[slot++] = (int)(tbase + tid * TP + q);
    }
  }
  if (tid == 0) bulk_wait0();

  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  if (lane == 0) {
    if (apply) atomicAdd(&A.cnt->apply_its, (unsigned long long)apply);
    if (unconv) atomicAdd(&A.cnt->unconverged, (unsigned long long)unconv);
  }
}

}  // namespace

// Returns 1 if the fast kernel was launched (deferred particles are then in s->defer_list),
// 0 if this species/configuration is not eligible, <0 on error.
int launch_advance_cc1_fast(pgpu_species_s *s, const AdvanceParams &prm, bool deposit) {
  Context &c = ctx();
  const pgpu_grid_s *g = s->grid;
  if (c.exact || g->desc.D != 2 || s->desc.interp_E != CC1) return 0;
  if (deposit && s->desc.interp_J != CC1) return 0;
  if (prm.iter_max < 0 && prm.order_swap) return 0;
  for (int d = 0; d < 2; ++d)
    if (s->desc.bc_check_lo[d] || s->desc.bc_check_hi[d]) return 0;
  if (g->nbox[0] + 2 * g->desc.nghost >= 32768 || g->nbox[1] + 2 * g->desc.nghost >= 32768) return 0;
  if (s->n == 0) return 1;
  if (!s->defer_list || s->defer_cap < (size_t)s->n) {
    if (s->defer_list) cudaFree(s->defer_list);
    if (!s->defer_count) PGPU_CUDA(cudaMalloc(&s->defer_count, sizeof(unsigned)));
    PGPU_CUDA(cudaMalloc(&s->defer_list, s->cap * sizeof(int)));
    s->defer_cap = s->cap;
  }
  PGPU_CUDA(cudaMemsetAsync(s->defer_count, 0, sizeof(unsigned), c.stream));

  FastArgs A;
  for (int d = 0; d < 2; ++d) {
    A.xo[d] = s->xold[d];
    A.xb[d] = s->x[d];
    A.le[d] = g->geo.le[d];
    A.dx[d] = g->geo.dx[d];
    A.rdx[d] = g->geo.rdx[d];
    A.hdx[d] = 0.5 * g->geo.dx[d];
  }
  for (int k = 0; k < 3; ++k) {
    A.uo[k] = s->vold[k];
    A.ub[k] = s->v[k];
  }
  A.w = s->w;
  A.n = s->n;
  // dual-cell range whose stencil (i0..i0+2 in either direction) lies inside all nine arrays
  int lo[2] = {-(1 << 30), -(1 << 30)}, hi[2] = {1 << 30, 1 << 30};
  auto clamp = [&](const DeviceFab &f) {
    for (int d = 0; d < 2; ++d) {
      lo[d] = lo[d] > f.lo[d] ? lo[d] : f.lo[d];
      hi[d] = hi[d] < f.hi[d] - 2 ? hi[d] : f.hi[d] - 2;
    }
  };
  for (int k = 0; k < 6; ++k) {
    const DeviceFab &f = g->field[k];
    clamp(f);
    A.F[k] = f.p - f.lo[0] - (long)f.lo[1] * f.n0;
    A.fn0[k] = f.n0;
  }
  for (int k = 0; k < 3; ++k) {
    const DeviceFab &f = s->J[k];
    clamp(f);
    A.J[k] = f.p - f.lo[0] - (long)f.lo[1] * f.n0;
    A.jn0[k] = f.n0;
  }
  for (int d = 0; d < 2; ++d) {
    A.i_lo[d] = lo[d];
    A.i_hi[d] = hi[d];
  }
  A.alpha = prm.alpha;
  A.hdt = prm.cnormDt * 0.5;
  A.rtol = prm.rtol;
  A.rvolume = prm.rvolume;
  A.iter_max = prm.iter_max;
  A.list = s->defer_list;
  A.list_count = s->defer_count;
  A.cnt = c.d_counters;
  KTimer t(deposit ? "advance_cc1_fused" : "advance_cc1");
  const bool tile_ok = s->cap >= (size_t)(((s->n + TILE - 1) / TILE) * TILE);
  if (c.cc1_tma == 2 && tile_ok) {
    const int ntiles = (int)((s->n + TILE - 1) / TILE);
    const int grid = std::min(ntiles, c.sm_count * 4 * c.cc1_waves);
#define PGPU_TILE_LAUNCH(DEPV, RS)                                                                        \
  do {                                                                                                    \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tile<DEPV, RS>,                                       \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE2_SMEM));       \
    k_advance_cc1_2d_tile<DEPV, RS><<<grid, BLOCK, TILE2_SMEM, c.stream>>>(A, ntiles);                    \
  } while (0)
    if (!deposit) PGPU_TILE_LAUNCH(false, 0);
    else if (c.cc1_rsteps == 0) PGPU_TILE_LAUNCH(true, 0);
    else if (c.cc1_rsteps == 1) PGPU_TILE_LAUNCH(true, 1);
    else if (c.cc1_rsteps == 2) PGPU_TILE_LAUNCH(true, 2);
    else if (c.cc1_rsteps == 3) PGPU_TILE_LAUNCH(true, 3);
    else if (c.cc1_rsteps == 4) PGPU_TILE_LAUNCH(true, 4);
    else PGPU_TILE_LAUNCH(true, 5);
    return 1;
  }
  if (c.cc1_tma == 1 && tile_ok) {
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)TMA_SMEM));
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)TMA_SMEM));
    const int ntiles = (int)((s->n + TILE - 1) / TILE);
    const int grid = std::min(ntiles, c.sm_count * 3);
    if (deposit) k_advance_cc1_2d_tma<true><<<grid, BLOCK, TMA_SMEM, c.stream>>>(A, ntiles);
    else k_advance_cc1_2d_tma<false><<<grid, BLOCK, TMA_SMEM, c.stream>>>(A, ntiles);
    return 1;
  }
  const int pairs = c.cc1_pairs;
  const long per_block = (long)BLOCK * 2 * pairs;
  const unsigned nb = (unsigned)((s->n + per_block - 1) / per_block);
  const int minb = c.cc1_minblocks;
#define PGPU_LAUNCH_CC1(DEPV, PV, MB) k_advance_cc1_2d<DEPV, PV, MB><<<nb, BLOCK, 0, c.stream>>>(A)
#define PGPU_PICK_MB(DEPV, PV)                            \
  do {                                                    \
    if (minb == 4) PGPU_LAUNCH_CC1(DEPV, PV, 4);          \
    else if (minb == 5) PGPU_LAUNCH_CC1(DEPV, PV, 5);     \
    else PGPU_LAUNCH_CC1(DEPV, PV, 3);                    \
  } while (0)
  if (deposit) {
    if (pairs == 1) PGPU_PICK_MB(true, 1);
    else if (pairs == 2) PGPU_PICK_MB(true, 2);
    else PGPU_PICK_MB(true, 4);
  } else {
    if (pairs == 1) PGPU_PICK_MB(false, 1);
    else if (pairs == 2) PGPU_PICK_MB(false, 2);
    else PGPU_PICK_MB(false, 4);
  }
  return 1;
}

}  // namespace pgpu
