// pgpu_massmatrix.cu -- mass-matrix deposit of the implicit solver's preconditioner / JFNK shortcut.
//
// Replaces, for CC1 and the planar push:
//   PicSpeciesInterface::initializeMassMatrices   src/species/pic/PicSpeciesInterface.cpp:225-417   (component tables)
//   PicSpeciesInterface::setMassMatrices          :1038-1095  (zero, accumulate per species, remember E0)
//   PicChargedSpecies::accumulateMassMatrices     src/species/pic/charged/PicChargedSpecies.cpp:3671-3761
//   MeshInterp::depositMassMatrices               src/particle_tools/MeshInterpI.H:231-475
//   cc1_1d_deposit_mass_matrix / cc1_2d_deposit_mass_matrix / compute_mm_kernals
//                                                 src/particle_tools/MeshInterpMassMatrixF.ChF:835-1220, 1228-1862, 1869-2076
//   PicSpeciesInterface::computeJfromMassMatrices :567-753 -> compute_J{x,y,z}_from_mass_matrix, src/fields/FieldsF.ChF:3-415
//
// The nine sigma containers live on the device next to the fields (231 components of a 512^2 box with 3 ghost
// layers = 0.5 GB); they never travel to the host inside a solve: the contraction J = J0 + sigma (E - E0) runs here,
// on the E of the selected field slot, into the grid's total current (pgpu_current_finalize / pgpu_current_get follow).
//
// Three deposit kernels:
//   k_mm<D>          one thread per particle, the reference's loops as they stand, one fp64 reduction per distinct
//                    address of a converged warp.  Any D, any number of segments, any particle order.
//   k_mm_cc1_2d_run  2D fast path for cell-sorted single-segment particles (the 95 % case): a warp stages the twelve
//                    kernel values and sixteen shape products of its 32 particles in shared memory, then every lane
//                    owns nine of the 272 (row point, column point) products of a particle and sums them in registers
//                    over the run of particles that share the dual cell and the node cell -- across the consecutive
//                    tiles of the warp's chunk -- with one RED per product and run instead of one per product and
//                    particle, and no shuffles.  Other particles go to a list for k_mm.
//   k_mm_cc1_1d_run  the same scheme in 1D (42 products, two per lane).
//
// This file is compiled with -fmad=false: every per-particle product is then the IEEE product the reference's
// Fortran computes, and the index decisions (true divide + floor) are bit exact; only the summation order differs.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pgpu_internal.h"

namespace pgpu {

enum { XX = 0, XY, XZ, YX, YY, YZ, ZX, ZY, ZZ };

struct MMView {
  double *p;
  int lo0, lo1, n0, n1, ncomp;
  long plane;
};
struct MMSet {
  MMView s[9];
  FabView J[3];
  FabView B[3];
};
struct MMParams {
  double qovs, alphas, volume;
  int anticyclic;  // +1 / -1
  int rel;
  int mX;          // maxXings
};

struct MassMatrices {
  int interp = -1;
  int ncomp[9][2];
  int mX = 0;
  DeviceFab row_box[3];       // box of the J component of a row (sigma arrays share it)
  double *arena = nullptr;     // J0[3] then sigma[9], one allocation (32-bit element offsets in the run kernel)
  size_t arena_elems = 0;
  double *sigma[9] = {nullptr};
  DeviceFab J0[3];
  DeviceFab E0[3];
  bool have_E0 = false;
  // fast path
  int *defer_list = nullptr;
  unsigned *defer_count = nullptr;
  size_t defer_cap = 0;
  void *table_d = nullptr;    // MMEntry[288]
  void *flush_d = nullptr;    // MMFlush[288]
};

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
struct MMTarget {
  const MMSet &T;
  unsigned err;
  bool aggregate;
  __device__ __forceinline__ MMTarget(const MMSet &t, bool agg) : T(t), err(0), aggregate(agg) {}
  __device__ __forceinline__ void add(double *base, long idx, double v) {
    if (aggregate) warp_aggregated_add(base, idx, v);
    else if (idx >= 0) atomicAdd(base + idx, v);
  }
  __device__ __forceinline__ void addJ(int c, int i, int j, double v) {
    const FabView &f = T.J[c];
    const unsigned a = (unsigned)(i - f.lo0), b = (unsigned)(j - f.lo1);
    const bool in = a < (unsigned)f.n0 && b < (unsigned)f.n1;
    if (!in) err |= ERRBIT_BOUNDS;
    add(f.p, in ? (long)(a + (size_t)b * f.n0) : -1L, v);
  }
  __device__ __forceinline__ void addS(int k, int i, int j, int nc, double v) {
    const MMView &f = T.s[k];
    const unsigned a = (unsigned)(i - f.lo0), b = (unsigned)(j - f.lo1);
    const bool in = a < (unsigned)f.n0 && b < (unsigned)f.n1 && (unsigned)nc < (unsigned)f.ncomp;
    if (!in) err |= ERRBIT_BOUNDS;
    add(f.p, in ? (long)(a + (size_t)b * f.n0) + (long)nc * f.plane : -1L, v);
  }
};

__device__ __forceinline__ int ifloor(double a) { return __double2int_rd(a); }

__device__ __forceinline__ bool fab_in(const FabView &f, int i, int j) {
  return (unsigned)(i - f.lo0) < (unsigned)f.n0 && (unsigned)(j - f.lo1) < (unsigned)f.n1;
}
__device__ __forceinline__ double fab_at(const FabView &f, int i, int j) {
  return __ldg(f.p + ((i - f.lo0) + (size_t)(j - f.lo1) * f.n0));
}

// compute_mm_kernals, planar branch (MeshInterpMassMatrixF.ChF:1941-2074); Bp is scaled in place
__device__ __forceinline__ void mm_kernels(double *fp, double (*f)[3], double *Bp, double qp, const MMParams &prm,
                                           const double *upold, const double *upbar) {
  const double alphas = prm.alphas;
  double gammap_bar = 1.0, gammap_new = 1.0, gammap_tilde = 1.0;
  double upnew[3] = {0.0, 0.0, 0.0};
  double rhop;
  if (prm.rel) {
#pragma unroll
    for (int c = 0; c < 3; ++c) upnew[c] = 2.0 * upbar[c] - upold[c];
    gammap_bar = sqrt(1.0 + upbar[0] * upbar[0] + upbar[1] * upbar[1] + upbar[2] * upbar[2]);
    const double gammap_old = sqrt(1.0 + upold[0] * upold[0] + upold[1] * upold[1] + upold[2] * upold[2]);
    gammap_new = sqrt(1.0 + upnew[0] * upnew[0] + upnew[1] * upnew[1] + upnew[2] * upnew[2]);
    gammap_tilde = 0.5 * (gammap_old + gammap_new);
    rhop = qp / prm.volume / gammap_tilde;
#pragma unroll
    for (int c = 0; c < 3; ++c) Bp[c] = alphas * Bp[c] / gammap_bar;
  } else {
    rhop = qp / prm.volume;
#pragma unroll
    for (int c = 0; c < 3; ++c) Bp[c] = alphas * Bp[c];
  }
  const double Bpx = Bp[0], Bpy = Bp[1], Bpz = Bp[2];
  const double ac = (double)prm.anticyclic;
  const double Bpsq = Bpx * Bpx + Bpy * Bpy + Bpz * Bpz;
  const double arogp = alphas * rhop / (1.0 + Bpsq);
  f[0][0] = arogp * (Bpx * Bpx + 1.0);
  f[0][1] = arogp * (Bpx * Bpy + ac * Bpz);
  f[0][2] = arogp * (Bpx * Bpz - ac * Bpy);
  f[1][0] = arogp * (Bpy * Bpx - ac * Bpz);
  f[1][1] = arogp * (Bpy * Bpy + 1.0);
  f[1][2] = arogp * (Bpy * Bpz + ac * Bpx);
  f[2][0] = arogp * (Bpz * Bpx + ac * Bpy);
  f[2][1] = arogp * (Bpz * Bpy - ac * Bpx);
  f[2][2] = arogp * (Bpz * Bpz + 1.0);
  if (prm.rel && gammap_bar > 1.01) {
    const double upBp = upbar[0] * Bpx + upbar[1] * Bpy + upbar[2] * Bpz;
    double gp_denom = gammap_bar * gammap_bar + Bpsq + upBp * upBp;
    double gp[3], upf[3];
    gp[0] = (Bpsq * upbar[0] - upBp * Bpx - ac * (upbar[1] * Bpz - upbar[2] * Bpy)) / gp_denom;
    gp[1] = (Bpsq * upbar[1] - upBp * Bpy - ac * (upbar[2] * Bpx - upbar[0] * Bpz)) / gp_denom;
    gp[2] = (Bpsq * upbar[2] - upBp * Bpz - ac * (upbar[0] * Bpy - upbar[1] * Bpx)) / gp_denom;
#pragma unroll
    for (int e = 0; e < 3; ++e) upf[e] = upbar[0] * f[0][e] + upbar[1] * f[1][e] + upbar[2] * f[2][e];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int e = 0; e < 3; ++e) f[j][e] = f[j][e] + gp[j] * upf[e];
    gp_denom = gammap_tilde * gammap_new;
#pragma unroll
    for (int j = 0; j < 3; ++j) gp[j] = -upbar[j] / gp_denom;
#pragma unroll
    for (int e = 0; e < 3; ++e) upf[e] = upnew[0] * f[0][e] + upnew[1] * f[1][e] + upnew[2] * f[2][e];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int e = 0; e < 3; ++e) f[j][e] = f[j][e] + gp[j] * upf[e];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) fp[c] = rhop * upbar[c];
}

// ---- cc1_1d_deposit_mass_matrix (:835-1220) ---------------------------------------------------------------
__device__ bool mm_cc1_1d(const Geo<1> &g, MMTarget &T, const MMParams &prm, const double *upold, const double *upbar,
                          double qp, double xpold, double xpbar) {
  const double dx = g.dx[0], le = g.le[0], re = g.re[0];
  const int index = ifloor((xpbar - le - 0.5 * dx) / dx);
  const int index_stag = ifloor((xpbar - le) / dx);
  double Bp[3] = {0.0, 0.0, 0.0};
  for (int ii = index; ii <= index + 1; ++ii) {
    const double l0 = ii * dx + 0.5 * dx - xpbar + le;
    const int ii_stag = ii - index + index_stag;
    const double l0_stag = ii_stag * dx - xpbar + le;
    const double w0 = 1.0 - fabs(l0 / dx);
    const double w0_stag = 1.0 - fabs(l0_stag / dx);
    if (!fab_in(T.T.B[0], ii_stag, 0) || !fab_in(T.T.B[1], ii, 0) || !fab_in(T.T.B[2], ii, 0)) {
      T.err |= ERRBIT_BOUNDS;
      return true;
    }
    Bp[0] = Bp[0] + w0_stag * fab_at(T.T.B[0], ii_stag, 0);
    Bp[1] = Bp[1] + w0 * fab_at(T.T.B[1], ii, 0);
    Bp[2] = Bp[2] + w0 * fab_at(T.T.B[2], ii, 0);
  }
  double fp[3], f[3][3];
  mm_kernels(fp, f, Bp, qp, prm, upold, upbar);

  // the crossing analysis first: a particle with too many segments deposits nothing at all here (the reference
  // stops the run at that point, :1069-1076)
  double xpnew = 2.0 * xpbar - xpold;
  const double dXp = fabs(xpnew - xpold);
  double bc_seg_factor = 1.0;
  double xpold0 = xpold;
  const int shift = (index == index_stag) ? 0 : 1;
  if (g.bc_lo[0] == 1) {
    if (xpold0 < le) {
      xpold0 = le;
      bc_seg_factor = fabs(xpnew - xpold0) / dXp;
    }
    if (xpnew < le) {
      xpnew = le;
      bc_seg_factor = fabs(xpnew - xpold0) / dXp;
    }
  }
  if (g.bc_hi[0] == 1) {
    if (xpold0 > re) {
      xpold0 = re;
      bc_seg_factor = fabs(xpnew - xpold0) / dXp;
    }
    if (xpnew > re) {
      xpnew = re;
      bc_seg_factor = fabs(xpnew - xpold0) / dXp;
    }
  }
  const double l0_stag = xpbar - index_stag * dx - le;
  const double wx_up_stag = l0_stag / dx;
  const double wx_dn_stag = 1.0 - wx_up_stag;
  const double l0 = xpbar - (index + 0.5) * dx - le;
  const double wx_up = l0 / dx;
  const double wx_dn = 1.0 - wx_up;
  const int index_old = ifloor((xpold0 - le - 0.5 * dx) / dx);
  const int index_new = ifloor((xpnew - le - 0.5 * dx) / dx);
  const int num_segments = 1 + abs(index_new - index_old);
  const int index_min = min(index_old, index_new);
  const int maxXings = prm.mX;
  if (num_segments > maxXings + 1 || num_segments > 3) return false;

  for (int ii_stag = index_stag; ii_stag <= index_stag + 1; ++ii_stag) {
    const double ls = ii_stag * dx - xpbar + le;
    const double weight = 1.0 - fabs(ls / dx);
    const int off_diag_comp = (ii_stag == index_stag) ? 2 : 0;
    T.addS(YY, ii_stag, 0, 1, f[1][1] * weight * weight);
    T.addS(YY, ii_stag, 0, off_diag_comp, f[1][1] * weight * (1.0 - weight));
    T.addS(YZ, ii_stag, 0, 1, f[1][2] * weight * weight);
    T.addS(YZ, ii_stag, 0, off_diag_comp, f[1][2] * weight * (1.0 - weight));
    T.addJ(1, ii_stag, 0, fp[1] * weight);
    T.addS(ZY, ii_stag, 0, 1, f[2][1] * weight * weight);
    T.addS(ZY, ii_stag, 0, off_diag_comp, f[2][1] * weight * (1.0 - weight));
    T.addS(ZZ, ii_stag, 0, 1, f[2][2] * weight * weight);
    T.addS(ZZ, ii_stag, 0, off_diag_comp, f[2][2] * weight * (1.0 - weight));
    T.addJ(2, ii_stag, 0, fp[2] * weight);
  }

  int SegNumX[3] = {1, 0, 0};
  double dn[3] = {0.0, wx_dn * bc_seg_factor, 0.0};
  double up[3] = {0.0, wx_up * bc_seg_factor, 0.0};
  const double xmin = fmin(xpold0, xpnew), xmax = fmax(xpold0, xpnew);
  if (num_segments == 2) {
    if (index_min < index) {
      const double Xcell = le + (index + 0.5) * dx;
      double dXp_sub = Xcell - xmin;
      dn[0] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      up[0] = dXp_sub / dXp - dn[0];
      dXp_sub = xmax - Xcell;
      up[1] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      dn[1] = dXp_sub / dXp - up[1];
      SegNumX[0] = 0;
      SegNumX[1] = 1;
    } else {
      const double Xcell = le + (index + 1.5) * dx;
      double dXp_sub = Xcell - xmin;
      dn[1] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      up[1] = dXp_sub / dXp - dn[1];
      dXp_sub = xmax - Xcell;
      up[2] = dXp_sub / dXp * dXp_sub / 2.0 / dx;
      dn[2] = dXp_sub / dXp - up[2];
      SegNumX[0] = 1;
      SegNumX[1] = 2;
    }
  }
  if (num_segments == 3) {
    double Xcell = le + (index + 0.5) * dx;
    double dXp_sub = Xcell - xmin;
    dn[0] = dXp_sub * dXp_sub / 2.0 / dXp / dx;
    up[0] = dXp_sub / dXp - dn[0];
    up[1] = dx / dXp / 2.0;
    dn[1] = up[1];
    Xcell = Xcell + dx;
    dXp_sub = xmax - Xcell;
    up[2] = dXp_sub * dXp_sub / 2.0 / dXp / dx;
    dn[2] = dXp_sub / dXp - up[2];
    SegNumX[0] = 0;
    SegNumX[1] = 1;
    SegNumX[2] = 2;
  }
#pragma unroll 1
  for (int nJ = 0; nJ < num_segments; ++nJ) {
    const int llJ = SegNumX[nJ];
    const int ii = index - 1 + llJ;
    for (int iiJ = 0; iiJ < 2; ++iiJ) {
      const double weight_J = iiJ ? up[llJ] : dn[llJ];
      T.addJ(0, ii + iiJ, 0, fp[0] * weight_J);
#pragma unroll 1
      for (int nE = 0; nE < num_segments; ++nE) {
        const int llE = SegNumX[nE];
        for (int iiE = 0; iiE < 2; ++iiE) {
          const double weight_E = iiE ? up[llE] : dn[llE];
          const int Nc = 1 + maxXings + iiE - iiJ + llE - llJ;
          T.addS(XX, ii + iiJ, 0, Nc, f[0][0] * weight_J * weight_E);
        }
      }
      for (int iiE = 0; iiE < 2; ++iiE) {
        const double weight_E = (1 - iiE) * wx_dn_stag + iiE * wx_up_stag;
        const int Nc = 1 + maxXings + shift + iiE - iiJ - llJ;
        T.addS(XY, ii + iiJ, 0, Nc, f[0][1] * weight_J * weight_E);
        T.addS(XZ, ii + iiJ, 0, Nc, f[0][2] * weight_J * weight_E);
      }
    }
  }
  for (int iiJ = 0; iiJ < 2; ++iiJ) {
    const double weight_J = (1 - iiJ) * wx_dn_stag + iiJ * wx_up_stag;
#pragma unroll 1
    for (int nE = 0; nE < num_segments; ++nE) {
      const int llE = SegNumX[nE];
      for (int iiE = 0; iiE < 2; ++iiE) {
        const double weight_E = iiE ? up[llE] : dn[llE];
        const int Nc = maxXings - shift + iiE - iiJ + llE;
        T.addS(YX, index_stag + iiJ, 0, Nc, f[1][0] * weight_J * weight_E);
        T.addS(ZX, index_stag + iiJ, 0, Nc, f[2][0] * weight_J * weight_E);
      }
    }
  }
  return true;
}

// ---- cc1_2d_deposit_mass_matrix (:1228-1862) ----------------------------------------------------------------
// a / dx[d] as the reference computes it (IEEE divide).  When dx is a power of two -- 0.25 in the reference's decks and
// in every BASELINE configuration -- 1/dx is exact and a * (1/dx) is the same double bit for bit, without the ~15
// instructions of a DDIV; the test is on the mantissa of dx and warp-uniform.  -DMM_FASTDIV takes the product for any dx
// in the shape arguments (at most one ulp off in a weight); index decisions never do.
__device__ __forceinline__ bool mm_pow2(double dx) {
  return (__double_as_longlong(dx) & 0x000fffffffffffffLL) == 0;
}
template <class G>
__device__ __forceinline__ double mm_div_exact(const G &g, double a, int d) {
  return mm_pow2(g.dx[d]) ? a * g.rdx[d] : a / g.dx[d];
}
#ifdef MM_FASTDIV
#define MM_DIVDX(a, d) ((a) * g.rdx[d])
#else
#define MM_DIVDX(a, d) mm_div_exact(g, (a), (d))
#endif
#define MM_FLOORDX(a, d) ifloor(mm_div_exact(g, (a), (d)))
// Per-particle set-up shared by the two 2D kernels: indices, CIC weights at xbar, B gather, kernels.
struct MM2DHead {
  int index[2], index_stag[2];
  double wv[2][2], wsv[2][2];
  double fp[3], f[3][3];
};
__device__ __forceinline__ bool mm_2d_head(const Geo<2> &g, const MMSet &T, const MMParams &prm, const double *upold,
                                           const double *upbar, double qp, const double *xpbar, MM2DHead &h) {
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    h.index[d] = MM_FLOORDX(xpbar[d] - g.le[d] - 0.5 * g.dx[d], d);
    h.index_stag[d] = MM_FLOORDX(xpbar[d] - g.le[d], d);
    const double l = xpbar[d] - ((h.index[d] + 0.5) * g.dx[d] + g.le[d]);
    h.wv[d][1] = MM_DIVDX(l, d);
    h.wv[d][0] = 1.0 - h.wv[d][1];
    const double ls = xpbar[d] - (h.index_stag[d] * g.dx[d] + g.le[d]);
    h.wsv[d][1] = MM_DIVDX(ls, d);
    h.wsv[d][0] = 1.0 - h.wsv[d][1];
  }
  double Bp[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int iiJ = 0; iiJ < 2; ++iiJ) {
    const int ii = h.index[0] + iiJ, ii_stag = h.index_stag[0] + iiJ;
#pragma unroll
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const int jj = h.index[1] + jjJ, jj_stag = h.index_stag[1] + jjJ;
      if (!fab_in(T.B[0], ii_stag, jj) || !fab_in(T.B[1], ii, jj_stag) || !fab_in(T.B[2], ii, jj)) return false;
      double weight = h.wsv[0][iiJ] * h.wv[1][jjJ];
      Bp[0] = Bp[0] + weight * fab_at(T.B[0], ii_stag, jj);
      weight = h.wv[0][iiJ] * h.wsv[1][jjJ];
      Bp[1] = Bp[1] + weight * fab_at(T.B[1], ii, jj_stag);
      weight = h.wv[0][iiJ] * h.wv[1][jjJ];
      Bp[2] = Bp[2] + weight * fab_at(T.B[2], ii, jj);
    }
  }
  mm_kernels(h.fp, h.f, Bp, qp, prm, upold, upbar);
  return true;
}

// The segment walk and the weights of every segment (:1417-1597).  Returns the number of segments, 0 when a
// direction crosses more than maxXings faces.
enum { MM_MAXSEG = 6 };
struct MM2DSegs {
  int SegNumX[MM_MAXSEG], SegNumY[MM_MAXSEG];
  double cicX[MM_MAXSEG][2], cicY[MM_MAXSEG][2], tscX[MM_MAXSEG][3], tscY[MM_MAXSEG][3];
};
__device__ int mm_2d_segments(const Geo<2> &g, const MMParams &prm, const int *index, const double *xpold_in,
                              const double *xpbar, MM2DSegs &S) {
  double xpold[2] = {xpold_in[0], xpold_in[1]};
  double xpnew[2], dXp[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    xpnew[d] = 2.0 * xpbar[d] - xpold[d];
    dXp[d] = xpnew[d] - xpold[d];
  }
  const double slope = dXp[1] / dXp[0];
  const double slope_inv = 1.0 / slope;
  truncate_boundaries<2, true>(g, xpold, xpnew, slope, slope_inv);
  int index_old[2], sign[2], cell_crossings[2];
  int num_segments = 1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    index_old[d] = ifloor((xpold[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
    const int index_new = ifloor((xpnew[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
    sign[d] = (index_new < index_old[d]) ? -1 : 1;
    cell_crossings[d] = abs(index_new - index_old[d]);
    num_segments += cell_crossings[d];
    if (cell_crossings[d] > prm.mX) return 0;
  }
  if (num_segments > MM_MAXSEG) return 0;
  double Xcell[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) Xcell[d] = g.le[d] + (index_old[d] + 0.5 * (1 - sign[d]) + 0.5) * g.dx[d];
  double xpold0[2] = {xpold[0], xpold[1]}, xpnew0[2] = {0.0, 0.0}, dXp_sub[2] = {0.0, 0.0};
  int ii_next = index_old[0], jj_next = index_old[1];
#pragma unroll 1
  for (int nn = 0; nn < num_segments; ++nn) {
    const int ii = ii_next, jj = jj_next;
    if (nn == num_segments - 1) {
      xpnew0[0] = xpnew[0];
      xpnew0[1] = xpnew[1];
      dXp_sub[0] = xpnew0[0] - xpold0[0];
      dXp_sub[1] = xpnew0[1] - xpold0[1];
    } else if (cell_crossings[0] == 0) {
      jj_next = jj + sign[1];
      Xcell[1] = Xcell[1] + sign[1] * g.dx[1];
      xpnew0[1] = Xcell[1];
      dXp_sub[1] = xpnew0[1] - xpold0[1];
      dXp_sub[0] = slope_inv * dXp_sub[1];
      xpnew0[0] = xpold0[0] + dXp_sub[0];
    } else if (cell_crossings[1] == 0) {
      ii_next = ii + sign[0];
      Xcell[0] = Xcell[0] + sign[0] * g.dx[0];
      xpnew0[0] = Xcell[0];
      dXp_sub[0] = xpnew0[0] - xpold0[0];
      dXp_sub[1] = slope * dXp_sub[0];
      xpnew0[1] = xpold0[1] + dXp_sub[1];
    } else {
      xpnew0[0] = Xcell[0] + sign[0] * g.dx[0];
      xpnew0[1] = Xcell[1] + sign[1] * g.dx[1];
      dXp_sub[0] = xpnew0[0] - xpold0[0];
      dXp_sub[1] = xpnew0[1] - xpold0[1];
      const double dXp_sub02 = slope_inv * dXp_sub[1];
      if (fabs(dXp_sub[0]) < fabs(dXp_sub02)) {
        dXp_sub[1] = slope * dXp_sub[0];
        xpnew0[1] = xpold0[1] + dXp_sub[1];
        Xcell[0] = xpnew0[0];
        ii_next = ii + sign[0];
        cell_crossings[0] -= 1;
      } else {
        dXp_sub[0] = slope_inv * dXp_sub[1];
        xpnew0[0] = xpold0[0] + dXp_sub[0];
        Xcell[1] = xpnew0[1];
        jj_next = jj + sign[1];
        cell_crossings[1] -= 1;
      }
    }
    double seg_factor[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) seg_factor[d] = (dXp[d] != 0.0) ? dXp_sub[d] / dXp[d] : 1.0;
    double xpbar0[2];
    int index_start[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      xpbar0[d] = 0.5 * (xpold0[d] + xpnew0[d]);
      index_start[d] = ifloor((xpbar0[d] - g.le[d] - 0.5 * g.dx[d]) / g.dx[d]);
    }
    S.SegNumX[nn] = 1 + index_start[0] - index[0];
    S.SegNumY[nn] = 1 + index_start[1] - index[1];
    const double delta0 = (xpbar0[0] - (g.le[0] + (ii + 0.5) * g.dx[0])) / g.dx[0];
    const double delta1 = (xpbar0[1] - (g.le[1] + (jj + 0.5) * g.dx[1])) / g.dx[1];
    S.cicX[nn][0] = (1.0 - delta0) * seg_factor[0];
    S.cicX[nn][1] = delta0 * seg_factor[0];
    S.cicY[nn][0] = (1.0 - delta1) * seg_factor[1];
    S.cicY[nn][1] = delta1 * seg_factor[1];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        double l = (index_start[d] + b) * g.dx[d] - xpold0[d] + g.le[d];
        double delta = fabs(l / g.dx[d]);
        double t = 1.5 - delta;
        const double w_old = (b == 1) ? 0.75 - delta * delta : 0.5 * (t * t);
        l = (index_start[d] + b) * g.dx[d] - xpnew0[d] + g.le[d];
        delta = fabs(l / g.dx[d]);
        t = 1.5 - delta;
        const double w_new = (b == 1) ? 0.75 - delta * delta : 0.5 * (t * t);
        if (d == 0) S.tscX[nn][b] = 0.5 * (w_old + w_new);
        else S.tscY[nn][b] = 0.5 * (w_old + w_new);
      }
    }
    xpold0[0] = xpnew0[0];
    xpold0[1] = xpnew0[1];
  }
  return num_segments;
}

__device__ bool mm_cc1_2d(const Geo<2> &g, MMTarget &T, const MMParams &prm, const double *upold, const double *upbar,
                          double qp, const double *xpold, const double *xpbar) {
  MM2DHead h;
  if (!mm_2d_head(g, T.T, prm, upold, upbar, qp, xpbar, h)) {
    T.err |= ERRBIT_BOUNDS;
    return true;
  }
  MM2DSegs S;
  const int num_segments = mm_2d_segments(g, prm, h.index, xpold, xpbar, S);
  if (num_segments == 0) return false;
  const int mX = prm.mX;
  int shift[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) shift[d] = (h.index[d] == h.index_stag[d]) ? 0 : 1;

  // Jz and sigma_zz (:1375-1411)
  for (int iiJ = 0; iiJ < 2; ++iiJ)
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const int ii_stag = h.index_stag[0] + iiJ, jj_stag = h.index_stag[1] + jjJ;
      const double weight_J = h.wsv[0][iiJ] * h.wsv[1][jjJ];
      T.addJ(2, ii_stag, jj_stag, h.fp[2] * weight_J);
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE) {
          const double weight_E = h.wsv[0][iiE] * h.wsv[1][jjE];
          const int Nc = 1 + iiE - iiJ + 3 * (1 + jjE - jjJ);
          T.addS(ZZ, ii_stag, jj_stag, Nc, h.f[2][2] * weight_J * weight_E);
        }
    }
  // segments (:1603-1793)
#pragma unroll 1
  for (int nJ = 0; nJ < num_segments; ++nJ) {
    const int llJ = S.SegNumX[nJ], mmJ = S.SegNumY[nJ];
    const int ii = h.index[0] - 1 + llJ, jj = h.index[1] - 1 + mmJ;
    for (int iiJ = 0; iiJ < 2; ++iiJ)
      for (int jjJ = 0; jjJ < 3; ++jjJ) {
        const double weight_J = S.cicX[nJ][iiJ] * S.tscY[nJ][jjJ];
        T.addJ(0, ii + iiJ, jj + jjJ, h.fp[0] * weight_J);
#pragma unroll 1
        for (int nE = 0; nE < num_segments; ++nE) {
          const int llE = S.SegNumX[nE], mmE = S.SegNumY[nE];
          for (int iiE = 0; iiE < 2; ++iiE)
            for (int jjE = 0; jjE < 3; ++jjE) {
              const int Nc = 1 + mX + llE - llJ + iiE - iiJ + (3 + 2 * mX) * (2 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = S.cicX[nE][iiE] * S.tscY[nE][jjE];
              T.addS(XX, ii + iiJ, jj + jjJ, Nc, h.f[0][0] * weight_J * weight_E);
            }
          for (int iiE = 0; iiE < 3; ++iiE)
            for (int jjE = 0; jjE < 2; ++jjE) {
              const int Nc = 1 + mX + llE - llJ + iiE - iiJ + (4 + 2 * mX) * (2 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = S.tscX[nE][iiE] * S.cicY[nE][jjE];
              T.addS(XY, ii + iiJ, jj + jjJ, Nc, h.f[0][1] * weight_J * weight_E);
            }
        }
        for (int iiE = 0; iiE < 2; ++iiE)
          for (int jjE = 0; jjE < 2; ++jjE) {
            const int Nc = 1 + mX + shift[0] - llJ + iiE - iiJ + (2 + 2 * mX) * (2 + mX + shift[1] - mmJ + jjE - jjJ);
            const double weight_E = h.wsv[0][iiE] * h.wsv[1][jjE];
            T.addS(XZ, ii + iiJ, jj + jjJ, Nc, h.f[0][2] * weight_J * weight_E);
          }
      }
    for (int iiJ = 0; iiJ < 3; ++iiJ)
      for (int jjJ = 0; jjJ < 2; ++jjJ) {
        const double weight_J = S.tscX[nJ][iiJ] * S.cicY[nJ][jjJ];
        T.addJ(1, ii + iiJ, jj + jjJ, h.fp[1] * weight_J);
#pragma unroll 1
        for (int nE = 0; nE < num_segments; ++nE) {
          const int llE = S.SegNumX[nE], mmE = S.SegNumY[nE];
          for (int iiE = 0; iiE < 2; ++iiE)
            for (int jjE = 0; jjE < 3; ++jjE) {
              const int Nc = 2 + mX + llE - llJ + iiE - iiJ + (4 + 2 * mX) * (1 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = S.cicX[nE][iiE] * S.tscY[nE][jjE];
              T.addS(YX, ii + iiJ, jj + jjJ, Nc, h.f[1][0] * weight_J * weight_E);
            }
          for (int iiE = 0; iiE < 3; ++iiE)
            for (int jjE = 0; jjE < 2; ++jjE) {
              const int Nc = 2 + mX + llE - llJ + iiE - iiJ + (5 + 2 * mX) * (1 + mX + mmE - mmJ + jjE - jjJ);
              const double weight_E = S.tscX[nE][iiE] * S.cicY[nE][jjE];
              T.addS(YY, ii + iiJ, jj + jjJ, Nc, h.f[1][1] * weight_J * weight_E);
            }
        }
        for (int iiE = 0; iiE < 2; ++iiE)
          for (int jjE = 0; jjE < 2; ++jjE) {
            const int Nc = 2 + mX + shift[0] - llJ + iiE - iiJ + (3 + 2 * mX) * (1 + mX + shift[1] - mmJ + jjE - jjJ);
            const double weight_E = h.wsv[0][iiE] * h.wsv[1][jjE];
            T.addS(YZ, ii + iiJ, jj + jjJ, Nc, h.f[1][2] * weight_J * weight_E);
          }
      }
  }
  // Jz rows against Ex, Ey (:1798-1858)
  for (int iiJ = 0; iiJ < 2; ++iiJ)
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const double weight_J = h.wsv[0][iiJ] * h.wsv[1][jjJ];
#pragma unroll 1
      for (int nE = 0; nE < num_segments; ++nE) {
        const int llE = S.SegNumX[nE], mmE = S.SegNumY[nE];
        for (int iiE = 0; iiE < 2; ++iiE)
          for (int jjE = 0; jjE < 3; ++jjE) {
            const int Nc = mX - shift[0] + llE + iiE - iiJ + (2 + 2 * mX) * (mX - shift[1] + mmE + jjE - jjJ);
            const double weight_E = S.cicX[nE][iiE] * S.tscY[nE][jjE];
            T.addS(ZX, h.index_stag[0] + iiJ, h.index_stag[1] + jjJ, Nc, h.f[2][0] * weight_J * weight_E);
          }
        for (int iiE = 0; iiE < 3; ++iiE)
          for (int jjE = 0; jjE < 2; ++jjE) {
            const int Nc = mX - shift[0] + llE + iiE - iiJ + (3 + 2 * mX) * (mX - shift[1] + mmE + jjE - jjJ);
            const double weight_E = S.tscX[nE][iiE] * S.cicY[nE][jjE];
            T.addS(ZY, h.index_stag[0] + iiJ, h.index_stag[1] + jjJ, Nc, h.f[2][1] * weight_J * weight_E);
          }
      }
    }
  return true;
}

// ---- the generic kernel: one thread per particle --------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(128)
k_mm(PartPtrs p, long n, Geo<D> g, MMSet T, MMParams prm, Counters *cnt, const int *list, const unsigned *list_count) {
  const long total = list ? (long)*list_count : n;
  const long stride = (long)gridDim.x * blockDim.x;
  unsigned err = 0;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long i = list ? (long)list[t] : t;
    double xb[2], xo[2], uo[3], ub[3];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xb[d] = p.x[d][i];
      xo[d] = p.xold[d][i];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      uo[c] = p.vold[c][i];
      ub[c] = p.v[c][i];
    }
    const double qp = p.w[i] * prm.qovs;
    MMTarget tgt(T, list == nullptr);
    bool ok;
    if constexpr (D == 1) ok = mm_cc1_1d(g, tgt, prm, uo, ub, qp, xo[0], xb[0]);
    else ok = mm_cc1_2d(g, tgt, prm, uo, ub, qp, xo, xb);
    if (!ok) err |= ERRBIT_SEGMENTS;
    err |= tgt.err;
  }
  if (err) atomicOr(&cnt->err, err);
}

// ---- the 2D run kernel -------------------------------------------------------------------------------------------
// A single-segment particle makes 272 products  f * weight_J * weight_E  (256 into the sigmas, 16 into J0 without
// weight_E).  Both weights come from three families of the particle's shape products,
//   P[0..5]   cicX[i]*tscY[j]   (Jx / Ex points, i*3+j)
//   P[6..11]  tscX[i]*cicY[j]   (Jy / Ey points, i*2+j)
//   P[12..15] wsx[i]*wsy[j]     (Jz / Ez points, i*2+j)
// and f from F[0..8] = f[row][col], F[9..11] = fp[row] (F[12] = 0 for padding), so a particle is a record of 13 + 16
// doubles + its key (index, index_stag): 31 doubles.  Product e of a particle is (F[fi] * P[pj]) * P[pe] -- the
// reference's f * weight_J * weight_E, bit for bit.
// Phase 1: every lane writes the record of its particle to shared memory (7.9 KB per warp).
// Phase 2: lane L owns nine products that share row factors (slots 0..5: one F[fi]*P[pj] and six consecutive column
// weights; slots 6,7: one row factor and two column weights; slot 8: a J0 product), walks the runs of equal keys four
// particles at a time and adds into nine register accumulators; at the end of a run (warp-uniform) every lane flushes
// its accumulators with one RED each.
struct MMEntry {
  unsigned char arr;   // 0..8 sigma, 9..11 J0, 255 = padding
  unsigned char fj;    // group id: (array, row point)
  unsigned char pe;    // column weight: index into P
  unsigned char base;  // 0: (index0, index1), 1: (index_stag0, index_stag1)
  unsigned char fi;    // index into F
  unsigned char pj;    // row weight: index into P
  unsigned char pad[2];
  signed char di, dj;  // point = base + (di, dj)
  signed char ncs0, ncs1;
  int nc0;             // component = nc0 + ncs0*shift0 + ncs1*shift1
};
// where product e of a run goes, as an element offset into the arena that holds J0 and the nine sigmas:
//   off0 + bi + bj*n0(row) + (shift0*ncs0 + shift1*ncs1)*plane(row),  (bi, bj) = the run's index (base 0) or
//   index_stag (base 1).  off0 folds in the array's place in the arena, the box origin, the point offset (di, dj) and
//   the component nc0.  meta: bit 0 base, bits 1-2 row, bits 8-15 ncs0, bits 16-23 ncs1 (signed), < 0 = padding.
//   Built on the host once per grid; every lane keeps its nine descriptors in registers.
struct MMFlush {
  unsigned off0;
  int meta;
};
template <int IMM>
__device__ __forceinline__ double lds_f64(unsigned addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
  return v;
}
// one particle (record at byte offset IMM from the lane's base addresses) into the lane's nine accumulators
struct MMLane {
  unsigned f6, j6, e6, f2, j2, e2, f1, j1;   // byte offsets into a record
};
template <int IMM>
__device__ __forceinline__ void mm_accumulate(double *acc, unsigned qb, const MMLane &o) {
  const double r6 = lds_f64<IMM>(qb + o.f6) * lds_f64<IMM>(qb + o.j6);
  const unsigned b6 = qb + o.e6;
  acc[0] += r6 * lds_f64<IMM>(b6);
  acc[1] += r6 * lds_f64<IMM + 8>(b6);
  acc[2] += r6 * lds_f64<IMM + 16>(b6);
  acc[3] += r6 * lds_f64<IMM + 24>(b6);
  acc[4] += r6 * lds_f64<IMM + 32>(b6);
  acc[5] += r6 * lds_f64<IMM + 40>(b6);
  const double r2 = lds_f64<IMM>(qb + o.f2) * lds_f64<IMM>(qb + o.j2);
  const unsigned b2 = qb + o.e2;
  acc[6] += r2 * lds_f64<IMM>(b2);
  acc[7] += r2 * lds_f64<IMM + 8>(b2);
  acc[8] += lds_f64<IMM>(qb + o.f1) * lds_f64<IMM>(qb + o.j1);   // J0: fp * weight_J (padding lanes read F[12] = 0)
}
#ifndef MM_MINBLOCKS
#define MM_MINBLOCKS 6
#endif
enum { MM_CHUNK_MAX = 32 };   // consecutive 32-particle tiles per warp and chunk (fewer on small problems)
enum { MM_NENT = 272, MM_NF = 13, MM_NP = 16, MM_KEY = MM_NF + MM_NP, MM_REC = 31, MM_EPL = 9, MM_WARPS = 2 };

__global__ void __launch_bounds__(32 * MM_WARPS, MM_MINBLOCKS)
k_mm_cc1_2d_run(PartPtrs p, long n, Geo<2> g, MMSet T, MMParams prm, const MMEntry *__restrict__ table,
                const MMFlush *__restrict__ flush, Counters *cnt, int *defer_list, unsigned *defer_count, int chunk) {
  extern __shared__ double mm_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *rec = mm_smem + (size_t)wid * 32 * MM_REC;
  const long nwarps = (long)gridDim.x * MM_WARPS;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(rec);
  // this lane's products (build_table): slots 0..5 share the row factor FJ[fj6] and take six consecutive column
  // weights, slots 6,7 share FJ[fj2] and take two, slot 8 is a J0 product (column weight 1).  Byte offsets into a
  // record, and where the run sums go.
  unsigned foff[MM_EPL];
#pragma unroll
  for (int e = 0; e < MM_EPL; ++e) foff[e] = flush[e * 32 + lane].off0;
  const int fmeta6 = flush[lane].meta, fmeta2 = flush[6 * 32 + lane].meta, fmeta1 = flush[8 * 32 + lane].meta;
  MMLane o;
  {
    const MMEntry g6 = table[lane], g2 = table[6 * 32 + lane], g1 = table[8 * 32 + lane];
    o.f6 = 8u * g6.fi;
    o.j6 = 8u * (MM_NF + g6.pj);
    o.e6 = 8u * (MM_NF + g6.pe);
    o.f2 = 8u * g2.fi;
    o.j2 = 8u * (MM_NF + g2.pj);
    o.e2 = 8u * (MM_NF + g2.pe);
    o.f1 = 8u * g1.fi;
    o.j1 = 8u * (MM_NF + g1.pj);
  }
  double *const arena = T.J[0].p;   // J0 of row x opens the arena
  const int rn0[3] = {T.J[0].n0, T.J[1].n0, T.J[2].n0};
  const int rplane[3] = {T.J[0].n0 * T.J[0].n1, T.J[1].n0 * T.J[1].n1, T.J[2].n0 * T.J[2].n1};
  unsigned err = 0;
  // A warp takes chunks of `chunk` consecutive tiles and keeps a run open across tile boundaries (and across
  // deferred particles): the reductions are what bounds this kernel (4.5e8 REDs = 3.6 of 6.1 ms when every tile
  // flushed its own fragments), so a run is flushed only when the key changes or the chunk ends.
  double acc[MM_EPL];
  long long r01 = LLONG_MIN, r23 = 0;   // key of the open run (LLONG_MIN: none)
  auto flush_run = [&]() {
    const int k0 = (int)(r01 >> 32), k1 = (int)r01, k2 = (int)(r23 >> 32), k3 = (int)r23;
    // every point of the stencils inside the arrays? (the row boxes are the J boxes; components are in range by
    // construction of the table)
    const bool in = fab_in(T.J[0], k0, k1) && fab_in(T.J[0], k0 + 1, k1 + 2) && fab_in(T.J[1], k0, k1) &&
                    fab_in(T.J[1], k0 + 2, k1 + 1) && fab_in(T.J[2], k2, k3) && fab_in(T.J[2], k2 + 1, k3 + 1);
    if (!in) {
      err |= ERRBIT_BOUNDS;
      return;
    }
    const int s0 = (k0 == k2) ? 0 : 1, s1 = (k1 == k3) ? 0 : 1;
    // the products of a slot group go to one array: the run-dependent part of the address once per group (32-bit:
    // the arena holds fewer than 2^32 doubles), the product-dependent part is foff[e]
    auto dyn = [&](int meta) -> long long {
      const int row = (meta >> 1) & 3;
      const int n0 = row == 0 ? rn0[0] : (row == 1 ? rn0[1] : rn0[2]);
      const int pl = row == 0 ? rplane[0] : (row == 1 ? rplane[1] : rplane[2]);
      const int bi = (meta & 1) ? k2 : k0, bj = (meta & 1) ? k3 : k1;
      const int sh = s0 * (int)(signed char)(meta >> 8) + s1 * (int)(signed char)(meta >> 16);
      return (long long)(bi + bj * n0 + sh * pl);
    };
    double *const t6 = arena + dyn(fmeta6), *const t2 = arena + dyn(fmeta2);
#pragma unroll
    for (int e = 0; e < 6; ++e) atomicAdd(t6 + foff[e], acc[e]);
    atomicAdd(t2 + foff[6], acc[6]);
    atomicAdd(t2 + foff[7], acc[7]);
    if (fmeta1 >= 0) atomicAdd(arena + dyn(fmeta1) + foff[8], acc[8]);
  };
  const long ntiles = (n + 31) / 32;
  for (long tile0 = ((long)blockIdx.x * MM_WARPS + wid) * chunk; tile0 < ntiles; tile0 += nwarps * chunk) {
   const long tile1 = tile0 + chunk < ntiles ? tile0 + chunk : ntiles;
#pragma unroll 1
   for (long tile = tile0; tile < tile1; ++tile) {
    const long base = tile * 32;
    const long i = base + lane;
    // ---- phase 1 -------------------------------------------------------------------------------------------------
    double *R = rec + lane * MM_REC;
    long long key01 = LLONG_MIN, key23 = lane;   // not on the fast path
    if (i < n) {
      const double xb[2] = {p.x[0][i], p.x[1][i]}, xo[2] = {p.xold[0][i], p.xold[1][i]};
      const double uo[3] = {p.vold[0][i], p.vold[1][i], p.vold[2][i]};
      const double ub[3] = {p.v[0][i], p.v[1][i], p.v[2][i]};
      const double qp = p.w[i] * prm.qovs;
      MM2DHead h;
      bool fast = mm_2d_head(g, T, prm, uo, ub, qp, xb, h);
      const bool oob = !fast;
      if (oob) err |= ERRBIT_BOUNDS;
      double cic[2][2], tsc[2][3];
      if (fast) {
        // single segment <=> xold and xnew lie in the dual cell of xbar (no boundary truncation on this path)
        fast = !(g.bc_lo[0] | g.bc_hi[0] | g.bc_lo[1] | g.bc_hi[1]);
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const double xn = 2.0 * xb[d] - xo[d];
          const int io = MM_FLOORDX(xo[d] - g.le[d] - 0.5 * g.dx[d], d);
          const int in = MM_FLOORDX(xn - g.le[d] - 0.5 * g.dx[d], d);
          // the weights of the only segment (:1517-1592 with nn = 0 = num_segments-1)
          const double dXp = xn - xo[d];
          const double seg_factor = (dXp != 0.0) ? dXp / dXp : 1.0;   // = 1 (folded by the compiler unless dXp is inf/nan)
          const double xpbar0 = 0.5 * (xo[d] + xn);
          const int index_start = MM_FLOORDX(xpbar0 - g.le[d] - 0.5 * g.dx[d], d);
          fast = fast && io == h.index[d] && in == h.index[d] && index_start == h.index[d];
          const double delta = MM_DIVDX(xpbar0 - (g.le[d] + (h.index[d] + 0.5) * g.dx[d]), d);
          cic[d][0] = (1.0 - delta) * seg_factor;
          cic[d][1] = delta * seg_factor;
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            double l = (index_start + b) * g.dx[d] - xo[d] + g.le[d];
            double dl = fabs(MM_DIVDX(l, d));
            double t = 1.5 - dl;
            const double w_old = (b == 1) ? 0.75 - dl * dl : 0.5 * (t * t);
            l = (index_start + b) * g.dx[d] - xn + g.le[d];
            dl = fabs(MM_DIVDX(l, d));
            t = 1.5 - dl;
            const double w_new = (b == 1) ? 0.75 - dl * dl : 0.5 * (t * t);
            tsc[d][b] = 0.5 * (w_old + w_new);
          }
        }
      }
      if (fast) {
        double *P = R + MM_NF;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 3; ++b) {
            P[a * 3 + b] = cic[0][a] * tsc[1][b];       // Jx / Ex point (a, b)
            P[6 + b * 2 + a] = tsc[0][b] * cic[1][a];   // Jy / Ey point (b, a)
          }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) P[12 + a * 2 + b] = h.wsv[0][a] * h.wsv[1][b];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int e = 0; e < 3; ++e) R[3 * j + e] = h.f[j][e];
          R[9 + j] = h.fp[j];
        }
        R[12] = 0.0;
        key01 = ((long long)h.index[0] << 32) | (unsigned)h.index[1];
        key23 = ((long long)h.index_stag[0] << 32) | (unsigned)h.index_stag[1];
      } else if (!oob) {
        defer_list[atomicAdd(defer_count, 1u)] = (int)i;
      }
    }
    reinterpret_cast<long long *>(R + MM_KEY)[0] = key01;
    reinterpret_cast<long long *>(R + MM_KEY)[1] = key23;
    __syncwarp();
    // ---- phase 2 -------------------------------------------------------------------------------------------------
    // runs of the tile: maximal stretches of consecutive fast-path lanes with one key (warp-uniform bookkeeping)
    const unsigned valid = __ballot_sync(0xffffffffu, key01 != LLONG_MIN);
    const long long p01 = __shfl_up_sync(0xffffffffu, key01, 1), p23 = __shfl_up_sync(0xffffffffu, key23, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key01 != p01 || key23 != p23);
    unsigned rem = valid;
#pragma unroll 1
    while (rem) {
      const int qs = __ffs(rem) - 1;
      const unsigned stop = (heads | ~valid) & ~((2u << qs) - 1u);
      const int qe = stop ? __ffs(stop) - 1 : 32;
      rem &= (qe == 32) ? 0u : ~((1u << qe) - 1u);
      const long long c01 = __shfl_sync(0xffffffffu, key01, qs), c23 = __shfl_sync(0xffffffffu, key23, qs);
      if (c01 != r01 || c23 != r23) {     // warp-uniform
        if (r01 != LLONG_MIN) flush_run();
        r01 = c01;
        r23 = c23;
#pragma unroll
        for (int e = 0; e < MM_EPL; ++e) acc[e] = 0.0;
      }
      int q = qs;
#pragma unroll 1
      for (; q + 4 <= qe; q += 4) {
        const unsigned qb = sbase + (unsigned)q * (MM_REC * 8);
        mm_accumulate<0>(acc, qb, o);
        mm_accumulate<MM_REC * 8>(acc, qb, o);
        mm_accumulate<2 * MM_REC * 8>(acc, qb, o);
        mm_accumulate<3 * MM_REC * 8>(acc, qb, o);
      }
#pragma unroll 1
      for (; q < qe; ++q) mm_accumulate<0>(acc, sbase + (unsigned)q * (MM_REC * 8), o);
    }
    __syncwarp();
   }
   if (r01 != LLONG_MIN) flush_run();
   r01 = LLONG_MIN;
  }
  if (err) atomicOr(&cnt->err, err);
}

// ---- the 1D run kernel ---------------------------------------------------------------------------------------------
// Same scheme as k_mm_cc1_2d_run for cc1_1d_deposit_mass_matrix (:835-1220) and one segment: 42 products per particle,
// (F[fi] * P[pj]) * P[pe] with
//   P[0..1] = Deltap_dn, Deltap_up (the CIC pair of the dual cell), P[2..3] = w0_stag of the two nodes (:966-967),
//   P[4..5] = 1 - w0_stag (:976), P[6..7] = wx_dn_stag, wx_up_stag (:1053-1055), P[8] = 1 (J0 products).
// The reference forms the node weights in two ways (1 - |l/dx| for the y/z rows, l/dx and its complement for the
// mixed terms); both are kept so that every product is the reference's.  Lane L owns products L and L + 32.
enum { MM1_NENT = 42, MM1_NP = 9, MM1_KEY = MM_NF + MM1_NP, MM1_REC = 23, MM1_EPL = 2 };

__global__ void __launch_bounds__(32 * MM_WARPS)
k_mm_cc1_1d_run(PartPtrs p, long n, Geo<1> g, MMSet T, MMParams prm, const MMEntry *__restrict__ table,
                const MMFlush *__restrict__ flush, Counters *cnt, int *defer_list, unsigned *defer_count, int chunk) {
  extern __shared__ double mm_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *rec = mm_smem + (size_t)wid * 32 * MM1_REC;
  const long nwarps = (long)gridDim.x * MM_WARPS;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(rec);
  unsigned of[MM1_EPL], oj[MM1_EPL], oe[MM1_EPL], foff[MM1_EPL];
  int fmeta[MM1_EPL];
#pragma unroll
  for (int e = 0; e < MM1_EPL; ++e) {
    const MMEntry en = table[e * 32 + lane];
    of[e] = 8u * en.fi;
    oj[e] = 8u * (MM_NF + en.pj);
    oe[e] = 8u * (MM_NF + en.pe);
    const MMFlush fl = flush[e * 32 + lane];
    foff[e] = fl.off0;
    fmeta[e] = fl.meta;
  }
  double *const arena = T.J[0].p;
  const int rplane[3] = {T.J[0].n0, T.J[1].n0, T.J[2].n0};
  unsigned err = 0;
  double acc[MM1_EPL];
  long long rkey = LLONG_MIN;
  auto flush_run = [&]() {
    const int k0 = (int)(rkey >> 32), k2 = (int)rkey;
    const bool in = fab_in(T.J[0], k0, 0) && fab_in(T.J[0], k0 + 1, 0) && fab_in(T.J[1], k2, 0) &&
                    fab_in(T.J[1], k2 + 1, 0) && fab_in(T.J[2], k2, 0) && fab_in(T.J[2], k2 + 1, 0);
    if (!in) {
      err |= ERRBIT_BOUNDS;
      return;
    }
    const int s0 = (k0 == k2) ? 0 : 1;
#pragma unroll
    for (int e = 0; e < MM1_EPL; ++e) {
      const int meta = fmeta[e];
      if (meta < 0) continue;
      const int row = (meta >> 1) & 3;
      const int pl = row == 0 ? rplane[0] : (row == 1 ? rplane[1] : rplane[2]);
      const int bi = (meta & 1) ? k2 : k0;
      const int sh = s0 * (int)(signed char)(meta >> 8);
      atomicAdd(arena + ((long long)foff[e] + (long long)(bi + sh * pl)), acc[e]);
    }
  };
  const long ntiles = (n + 31) / 32;
  for (long tile0 = ((long)blockIdx.x * MM_WARPS + wid) * chunk; tile0 < ntiles; tile0 += nwarps * chunk) {
   const long tile1 = tile0 + chunk < ntiles ? tile0 + chunk : ntiles;
#pragma unroll 1
   for (long tile = tile0; tile < tile1; ++tile) {
    const long i = tile * 32 + lane;
    double *R = rec + lane * MM1_REC;
    long long key = LLONG_MIN;
    if (i < n) {
      const double xpbar = p.x[0][i], xpold = p.xold[0][i];
      const double uo[3] = {p.vold[0][i], p.vold[1][i], p.vold[2][i]};
      const double ub[3] = {p.v[0][i], p.v[1][i], p.v[2][i]};
      const double qp = p.w[i] * prm.qovs;
      const double dx = g.dx[0], le = g.le[0];
      const int index = MM_FLOORDX(xpbar - le - 0.5 * dx, 0);
      const int index_stag = MM_FLOORDX(xpbar - le, 0);
      // magnetic field at the particle (:899-943) and the node weights of the y/z rows (:961-967)
      double Bp[3] = {0.0, 0.0, 0.0}, wsa[2];
      bool ok = true;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int ii = index + a, ii_stag = index_stag + a;
        const double l0 = ii * dx + 0.5 * dx - xpbar + le;
        const double l0_stag = ii_stag * dx - xpbar + le;
        const double w0 = 1.0 - fabs(mm_div_exact(g, l0, 0));
        const double w0_stag = 1.0 - fabs(mm_div_exact(g, l0_stag, 0));
        wsa[a] = w0_stag;
        if (!fab_in(T.B[0], ii_stag, 0) || !fab_in(T.B[1], ii, 0) || !fab_in(T.B[2], ii, 0)) {
          ok = false;
        } else {
          Bp[0] = Bp[0] + w0_stag * fab_at(T.B[0], ii_stag, 0);
          Bp[1] = Bp[1] + w0 * fab_at(T.B[1], ii, 0);
          Bp[2] = Bp[2] + w0 * fab_at(T.B[2], ii, 0);
        }
      }
      if (!ok) err |= ERRBIT_BOUNDS;
      bool fast = ok && !(g.bc_lo[0] | g.bc_hi[0]);
      if (fast) {
        const double xpnew = 2.0 * xpbar - xpold;
        const int index_old = MM_FLOORDX(xpold - le - 0.5 * dx, 0);
        const int index_new = MM_FLOORDX(xpnew - le - 0.5 * dx, 0);
        fast = index_old == index && index_new == index;   // one segment, in the dual cell of xbar
      }
      if (fast) {
        double fp[3], f[3][3];
        mm_kernels(fp, f, Bp, qp, prm, uo, ub);
        // :1052-1058, :1099-1104 with bc_seg_factor = 1
        const double l0_stag = xpbar - index_stag * dx - le;
        const double wx_up_stag = mm_div_exact(g, l0_stag, 0);
        const double wx_dn_stag = 1.0 - wx_up_stag;
        const double l0 = xpbar - (index + 0.5) * dx - le;
        const double wx_up = mm_div_exact(g, l0, 0);
        const double wx_dn = 1.0 - wx_up;
        double *P = R + MM_NF;
        P[0] = wx_dn * 1.0;
        P[1] = wx_up * 1.0;
        P[2] = wsa[0];
        P[3] = wsa[1];
        P[4] = 1.0 - wsa[0];
        P[5] = 1.0 - wsa[1];
        P[6] = wx_dn_stag;
        P[7] = wx_up_stag;
        P[8] = 1.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int e = 0; e < 3; ++e) R[3 * j + e] = f[j][e];
          R[9 + j] = fp[j];
        }
        R[12] = 0.0;
        key = ((long long)index << 32) | (unsigned)index_stag;
      } else if (ok) {
        defer_list[atomicAdd(defer_count, 1u)] = (int)i;
      }
    }
    reinterpret_cast<long long *>(R + MM1_KEY)[0] = key;
    __syncwarp();
    const unsigned valid = __ballot_sync(0xffffffffu, key != LLONG_MIN);
    const long long pk = __shfl_up_sync(0xffffffffu, key, 1);
    const unsigned heads = __ballot_sync(0xffffffffu, lane == 0 || key != pk);
    unsigned rem = valid;
#pragma unroll 1
    while (rem) {
      const int qs = __ffs(rem) - 1;
      const unsigned stop = (heads | ~valid) & ~((2u << qs) - 1u);
      const int qe = stop ? __ffs(stop) - 1 : 32;
      rem &= (qe == 32) ? 0u : ~((1u << qe) - 1u);
      const long long ck = __shfl_sync(0xffffffffu, key, qs);
      if (ck != rkey) {
        if (rkey != LLONG_MIN) flush_run();
        rkey = ck;
#pragma unroll
        for (int e = 0; e < MM1_EPL; ++e) acc[e] = 0.0;
      }
      int q = qs;
#pragma unroll 1
      for (; q + 4 <= qe; q += 4) {
        const unsigned qb = sbase + (unsigned)q * (MM1_REC * 8);
#pragma unroll
        for (int e = 0; e < MM1_EPL; ++e) {
          const unsigned a = qb + of[e], b = qb + oj[e], c = qb + oe[e];
          acc[e] += (lds_f64<0>(a) * lds_f64<0>(b)) * lds_f64<0>(c);
          acc[e] += (lds_f64<MM1_REC * 8>(a) * lds_f64<MM1_REC * 8>(b)) * lds_f64<MM1_REC * 8>(c);
          acc[e] += (lds_f64<2 * MM1_REC * 8>(a) * lds_f64<2 * MM1_REC * 8>(b)) * lds_f64<2 * MM1_REC * 8>(c);
          acc[e] += (lds_f64<3 * MM1_REC * 8>(a) * lds_f64<3 * MM1_REC * 8>(b)) * lds_f64<3 * MM1_REC * 8>(c);
        }
      }
#pragma unroll 1
      for (; q < qe; ++q) {
        const unsigned qb = sbase + (unsigned)q * (MM1_REC * 8);
#pragma unroll
        for (int e = 0; e < MM1_EPL; ++e)
          acc[e] += (lds_f64<0>(qb + of[e]) * lds_f64<0>(qb + oj[e])) * lds_f64<0>(qb + oe[e]);
      }
    }
    __syncwarp();
   }
   if (rkey != LLONG_MIN) flush_run();
   rkey = LLONG_MIN;
  }
  if (err) atomicOr(&cnt->err, err);
}

// ---- J = J0 + sigma (E - E0)  (FieldsF.ChF:3-415) ---------------------------------------------------------------
struct ContractArgs {
  MMView s[3];            // the three sigma arrays of the row
  FabView E[3], E0[3];
  FabView J0, J;
  int N0[3], N1[3], o0[3], o1[3];
  int D;
};
__global__ void __launch_bounds__(256) k_mm_contract(ContractArgs A) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)A.J.n0 * A.J.n1;
  if (t >= total) return;
  const int a = (int)(t % A.J.n0), b = (int)(t / A.J.n0);
  const int i = a + A.J.lo0, j = b + A.J.lo1;
  double part[3];
#pragma unroll
  for (int e = 0; e < 3; ++e) {
    const MMView &S = A.s[e];
    const FabView &Ev = A.E[e], &E0v = A.E0[e];
    const int ehi0 = Ev.lo0 + Ev.n0 - 1, ehi1 = Ev.lo1 + Ev.n1 - 1;
    const int ii_min = max(0, A.o0[e] + Ev.lo0 - i), ii_max = min(A.N0[e] - 1, A.o0[e] + ehi0 - i);
    int jj_min = 0, jj_max = 0;
    if (A.D >= 2) {
      jj_min = max(0, A.o1[e] + Ev.lo1 - j);
      jj_max = min(A.N1[e] - 1, A.o1[e] + ehi1 - j);
    }
    double acc = 0.0;
    for (int ii = ii_min; ii <= ii_max; ++ii)
      for (int jj = jj_min; jj <= jj_max; ++jj) {
        const int ei = i + ii - A.o0[e], ej = j + jj - A.o1[e];
        const size_t eo = (size_t)(ei - Ev.lo0) + (size_t)(ej - Ev.lo1) * Ev.n0;
        const double dE = Ev.p[eo] - E0v.p[eo];
        const int Nc = ii + A.N0[e] * jj;
        acc = acc + S.p[(size_t)a + (size_t)b * S.n0 + (size_t)Nc * S.plane] * dE;
      }
    part[e] = acc;
  }
  const double sigdE = part[0] + part[1] + part[2];
  A.J.p[t] = A.J0.p[t] + sigdE;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// PicSpeciesInterface.cpp:256-350
static int mm_ncomp_table(int D, int interp, int ghosts, int nc[9][2]) {
  const int tsc = (interp == TSC) ? 2 : 0;
  static const int d0[9] = {3, 4, 4, 4, 3, 3, 4, 3, 3};
  static const int d1[9] = {3, 4, 3, 4, 3, 4, 3, 4, 3};
  for (int k = 0; k < 9; ++k) {
    nc[k][0] = d0[k] + tsc;
    nc[k][1] = (D >= 2) ? d1[k] + tsc : 1;
  }
  if (interp == CC1 && D == 1) {
    if (ghosts < 2) return -1;
    const int m = ghosts - 1;
    nc[XX][0] = 3 + 2 * m;
    nc[XY][0] = nc[XZ][0] = nc[YX][0] = nc[ZX][0] = 2 + 2 * m;
    return 0;
  }
  if (interp == CC1 && D == 2) {
    if (ghosts < 3) return -1;
    const int m = ghosts - 2;
    static const int c0[9] = {3, 4, 2, 4, 5, 3, 2, 3, 3};
    static const int c1[9] = {5, 4, 3, 4, 3, 2, 3, 2, 3};
    for (int k = 0; k < 8; ++k) {
      nc[k][0] = c0[k] + 2 * m;
      nc[k][1] = c1[k] + 2 * m;
    }
    nc[ZZ][0] = nc[ZZ][1] = 3;
    return 0;
  }
  return -2;
}

// the 272 products of a single-segment 2D particle (llJ = mmJ = llE = mmE = 1) following the reference's loops
// (:1375-1411, :1603-1858), then dealt to the lanes (9 slots x 32 lanes)
static int build_table(int mX, MMEntry *tab_out) {
  MMEntry tab_buf[MM_EPL * 32 + 32];
  MMEntry *tab = tab_buf;
  struct CopyBack {
    MMEntry *from, *to;
    ~CopyBack() { memcpy(to, from, sizeof(MMEntry) * MM_EPL * 32); }
  } copy_back{tab_buf, tab_out};
  int n = 0;
  auto put = [&](int arr, int fj, int pe, int di, int dj, int nc0, int ncs0, int ncs1, int base) {
    MMEntry e;
    memset(&e, 0, sizeof(e));
    e.arr = (unsigned char)arr;
    e.fj = (unsigned char)fj;
    e.pe = (unsigned char)pe;
    // row factor of group fj: F index and row-weight index (families PX = 0, PY = 6, PZ = 12)
    if (fj < 18) { e.fi = (unsigned char)(0 + fj / 6); e.pj = (unsigned char)(0 + fj % 6); }
    else if (fj < 36) { e.fi = (unsigned char)(3 + (fj - 18) / 6); e.pj = (unsigned char)(6 + (fj - 18) % 6); }
    else if (fj < 48) { e.fi = (unsigned char)(6 + (fj - 36) / 4); e.pj = (unsigned char)(12 + (fj - 36) % 4); }
    else if (fj < 54) { e.fi = 9; e.pj = (unsigned char)(0 + fj - 48); }
    else if (fj < 60) { e.fi = 10; e.pj = (unsigned char)(6 + fj - 54); }
    else { e.fi = 11; e.pj = (unsigned char)(12 + fj - 60); }
    e.base = (unsigned char)base;
    e.di = (signed char)di;
    e.dj = (signed char)dj;
    e.ncs0 = (signed char)ncs0;
    e.ncs1 = (signed char)ncs1;
    e.nc0 = nc0;
    tab[n++] = e;
  };
  const int PX = 0, PY = 6, PZ = 12, ONE = 16;   // families of P
  // Jz, sigma_zz (row z, column z)
  for (int iiJ = 0; iiJ < 2; ++iiJ)
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const int pt = iiJ * 2 + jjJ;
      put(9 + 2, 60 + pt, ONE, iiJ, jjJ, 0, 0, 0, 1);
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE)
          put(ZZ, 36 + 2 * 4 + pt, PZ + iiE * 2 + jjE, iiJ, jjJ, 1 + iiE - iiJ + 3 * (1 + jjE - jjJ), 0, 0, 1);
    }
  // Jx rows
  for (int iiJ = 0; iiJ < 2; ++iiJ)
    for (int jjJ = 0; jjJ < 3; ++jjJ) {
      const int pt = iiJ * 3 + jjJ;
      put(9 + 0, 48 + pt, ONE, iiJ, jjJ, 0, 0, 0, 0);
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 3; ++jjE)
          put(XX, 0 + pt, PX + iiE * 3 + jjE, iiJ, jjJ, 1 + mX + iiE - iiJ + (3 + 2 * mX) * (2 + mX + jjE - jjJ), 0, 0, 0);
      for (int iiE = 0; iiE < 3; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE)
          put(XY, 6 + pt, PY + iiE * 2 + jjE, iiJ, jjJ, 1 + mX + iiE - iiJ + (4 + 2 * mX) * (2 + mX + jjE - jjJ), 0, 0, 0);
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE)
          put(XZ, 12 + pt, PZ + iiE * 2 + jjE, iiJ, jjJ,
              1 + mX - 1 + iiE - iiJ + (2 + 2 * mX) * (2 + mX - 1 + jjE - jjJ), 1, 2 + 2 * mX, 0);
    }
  // Jy rows
  for (int iiJ = 0; iiJ < 3; ++iiJ)
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const int pt = iiJ * 2 + jjJ;
      put(9 + 1, 54 + pt, ONE, iiJ, jjJ, 0, 0, 0, 0);
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 3; ++jjE)
          put(YX, 18 + pt, PX + iiE * 3 + jjE, iiJ, jjJ, 2 + mX + iiE - iiJ + (4 + 2 * mX) * (1 + mX + jjE - jjJ), 0, 0, 0);
      for (int iiE = 0; iiE < 3; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE)
          put(YY, 24 + pt, PY + iiE * 2 + jjE, iiJ, jjJ, 2 + mX + iiE - iiJ + (5 + 2 * mX) * (1 + mX + jjE - jjJ), 0, 0, 0);
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE)
          put(YZ, 30 + pt, PZ + iiE * 2 + jjE, iiJ, jjJ,
              2 + mX - 1 + iiE - iiJ + (3 + 2 * mX) * (1 + mX - 1 + jjE - jjJ), 1, 3 + 2 * mX, 0);
    }
  // Jz rows against Ex, Ey
  for (int iiJ = 0; iiJ < 2; ++iiJ)
    for (int jjJ = 0; jjJ < 2; ++jjJ) {
      const int pt = iiJ * 2 + jjJ;
      for (int iiE = 0; iiE < 2; ++iiE)
        for (int jjE = 0; jjE < 3; ++jjE)
          put(ZX, 36 + pt, PX + iiE * 3 + jjE, iiJ, jjJ, mX + 1 + iiE - iiJ + (2 + 2 * mX) * (mX + 1 + jjE - jjJ), -1,
              -(2 + 2 * mX), 1);
      for (int iiE = 0; iiE < 3; ++iiE)
        for (int jjE = 0; jjE < 2; ++jjE)
          put(ZY, 40 + pt, PY + iiE * 2 + jjE, iiJ, jjJ, mX + 1 + iiE - iiJ + (3 + 2 * mX) * (mX + 1 + jjE - jjJ), -1,
              -(3 + 2 * mX), 1);
    }
  const int real = n;
  if (real != MM_NENT) return real;
  // Lane assignment: products that share the row factor FJ[fj] sit on one lane, so that a lane loads it once per
  // particle.  32 groups of six column points (xx xy yx yy zx zy) -> slots 0..5 of lane L; 16 groups of four (xz yz zz)
  // in halves -> slots 6,7; the 16 J0 products -> slot 8 of lanes 0..15 (padding elsewhere).
  MMEntry flat[MM_NENT];
  memcpy(flat, tab, sizeof(flat));
  int first6[32], n6 = 0, first4[16], n4 = 0, j0[16], n1 = 0;
  for (int k = 0; k < MM_NENT; ++k) {
    if (flat[k].arr >= 9) {
      if (n1 < 16) j0[n1] = k;
      ++n1;
      continue;
    }
    bool seen = false;
    int cnt = 0;
    for (int q = 0; q < MM_NENT; ++q)
      if (flat[q].arr < 9 && flat[q].fj == flat[k].fj) {
        if (q < k) seen = true;
        ++cnt;
      }
    if (seen) continue;
    if (cnt == 6 && n6 < 32) first6[n6++] = k;
    else if (cnt == 4 && n4 < 16) first4[n4++] = k;
    else return -1;
  }
  if (n6 != 32 || n4 != 16 || n1 != 16) return -1;
  MMEntry pad;
  memset(&pad, 0, sizeof(pad));
  pad.arr = 255;
  pad.fi = 12;   // F[12] = 0
  pad.pj = 0;
  pad.pe = 0;
  auto member = [&](int fj, int rank) -> int {   // the rank-th product (by column weight index) of group fj
    int idx[6], m = 0;
    for (int q = 0; q < MM_NENT; ++q)
      if (flat[q].arr < 9 && flat[q].fj == fj) idx[m++] = q;
    for (int a = 0; a < m; ++a)
      for (int b = a + 1; b < m; ++b)
        if (flat[idx[b]].pe < flat[idx[a]].pe) std::swap(idx[a], idx[b]);
    return idx[rank];
  };
  for (int L = 0; L < 32; ++L) {
    for (int e = 0; e < 6; ++e) tab[e * 32 + L] = flat[member(flat[first6[L]].fj, e)];
    for (int e = 0; e < 2; ++e) tab[(6 + e) * 32 + L] = flat[member(flat[first4[L / 2]].fj, 2 * (L % 2) + e)];
    tab[8 * 32 + L] = (L < 16) ? flat[j0[L]] : pad;
    // the kernel relies on consecutive column weights within a slot group
    for (int e = 1; e < 6; ++e)
      if (tab[e * 32 + L].pe != tab[L].pe + e) return -1;
    if (tab[7 * 32 + L].pe != tab[6 * 32 + L].pe + 1) return -1;
  }
  return real;
}

// the 42 products of a single-segment 1D particle (cc1_1d_deposit_mass_matrix :961-1220 with llJ = llE = 1), 2 x 32 slots
static int build_table_1d(int mX, MMEntry *tab) {
  int n = 0;
  MMEntry pad;
  memset(&pad, 0, sizeof(pad));
  pad.arr = 255;
  pad.fi = 12;
  for (int k = 0; k < MM1_EPL * 32; ++k) tab[k] = pad;
  auto put = [&](int arr, int fi, int pj, int pe, int base, int di, int nc0, int ncs0) {
    MMEntry e;
    memset(&e, 0, sizeof(e));
    e.arr = (unsigned char)arr;
    e.fi = (unsigned char)fi;
    e.pj = (unsigned char)pj;
    e.pe = (unsigned char)pe;
    e.base = (unsigned char)base;
    e.di = (signed char)di;
    e.nc0 = nc0;
    e.ncs0 = (signed char)ncs0;
    tab[n++] = e;
  };
  const int DN = 0, WSA = 2, WSC = 4, WSB = 6, ONE = 8;
  for (int p = 0; p < 2; ++p) {                       // y and z rows against Ey, Ez (:961-1009)
    const int off = (p == 0) ? 2 : 0;
    const int arrs[4] = {YY, YZ, ZY, ZZ}, fis[4] = {4, 5, 7, 8};
    for (int a = 0; a < 4; ++a) {
      put(arrs[a], fis[a], WSA + p, WSA + p, 1, p, 1, 0);
      put(arrs[a], fis[a], WSA + p, WSC + p, 1, p, off, 0);
    }
    put(9 + 1, 10, WSA + p, ONE, 1, p, 0, 0);
    put(9 + 2, 11, WSA + p, ONE, 1, p, 0, 0);
  }
  for (int iiJ = 0; iiJ < 2; ++iiJ) {                 // x row (:1153-1195)
    put(9 + 0, 9, DN + iiJ, ONE, 0, iiJ, 0, 0);
    for (int iiE = 0; iiE < 2; ++iiE) put(XX, 0, DN + iiJ, DN + iiE, 0, iiJ, 1 + mX + iiE - iiJ, 0);
    for (int iiE = 0; iiE < 2; ++iiE) {
      put(XY, 1, DN + iiJ, WSB + iiE, 0, iiJ, 1 + mX + iiE - iiJ - 1, 1);
      put(XZ, 2, DN + iiJ, WSB + iiE, 0, iiJ, 1 + mX + iiE - iiJ - 1, 1);
    }
  }
  for (int iiJ = 0; iiJ < 2; ++iiJ)                   // y and z rows against Ex (:1200-1220)
    for (int iiE = 0; iiE < 2; ++iiE) {
      put(YX, 3, WSB + iiJ, DN + iiE, 1, iiJ, mX + iiE - iiJ + 1, -1);
      put(ZX, 6, WSB + iiJ, DN + iiE, 1, iiJ, mX + iiE - iiJ + 1, -1);
    }
  return n;
}

static MassMatrices *mm_of(pgpu_grid_s *g) { return static_cast<MassMatrices *>(g->mm); }

void mm_destroy(pgpu_grid_s *g) {
  MassMatrices *m = mm_of(g);
  if (!m) return;
  if (m->arena) cudaFree(m->arena);
  for (int c = 0; c < 3; ++c)
    if (m->E0[c].p) cudaFree(m->E0[c].p);
  if (m->defer_list) cudaFree(m->defer_list);
  if (m->defer_count) cudaFree(m->defer_count);
  if (m->table_d) cudaFree(m->table_d);
  if (m->flush_d) cudaFree(m->flush_d);
  delete m;
  g->mm = nullptr;
}

static MMSet make_set(pgpu_grid_s *g, MassMatrices *m) {
  MMSet T;
  for (int k = 0; k < 9; ++k) {
    const DeviceFab &b = m->row_box[k / 3];
    MMView v;
    v.p = m->sigma[k];
    v.lo0 = b.lo[0];
    v.lo1 = b.lo[1];
    v.n0 = b.n0;
    v.n1 = b.n1;
    v.ncomp = m->ncomp[k][0] * m->ncomp[k][1];
    v.plane = (long)b.n0 * b.n1;
    T.s[k] = v;
  }
  for (int c = 0; c < 3; ++c) {
    T.J[c] = m->J0[c].view();
    T.B[c] = g->field[3 + c].view();   // callers have run fields_wait(g)
  }
  return T;
}

}  // namespace pgpu

using namespace pgpu;

#define NEED_MM(g)                                                              \
  if (!ctx().inited) {                                                          \
    set_error("pgpu_init has not been called");                                 \
    return PGPU_ERR_STATE;                                                      \
  }                                                                             \
  if (!(g) || !(g)->mm) {                                                       \
    set_error("mass matrices are not initialised (pgpu_mass_matrices_init)");   \
    return PGPU_ERR_STATE;                                                      \
  }                                                                             \
  if (fields_wait(g)) return PGPU_ERR_CUDA;

extern "C" {

int pgpu_mass_matrices_init(pgpu_grid_t g, int interp, int *ncomp_out) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  if (!g) return PGPU_ERR_ARG;
  int nc[9][2];
  const int rc = mm_ncomp_table(g->desc.D, interp, g->desc.nghost, nc);
  if (rc == -1) {
    set_error("mass matrices with CC1 need grid.num_ghosts >= %d (PicSpeciesInterface.cpp:310,321)",
              g->desc.D == 1 ? 2 : 3);
    return PGPU_ERR_ARG;
  }
  if (rc) {
    set_error("mass matrices are implemented for CC1 interpolation only");
    return PGPU_ERR_ARG;
  }
  if (g->mm) mm_destroy(g);
  MassMatrices *m = new MassMatrices();
  g->mm = m;
  m->interp = interp;
  memcpy(m->ncomp, nc, sizeof(nc));
  m->mX = g->desc.D == 1 ? g->desc.nghost - 1 : g->desc.nghost - 2;
  cudaStream_t st = ctx().stream;
  size_t elems = 0;
  for (int c = 0; c < 3; ++c) {
    // the J component of the row: same box as the grid's total current
    m->row_box[c] = g->jtot[c];
    m->row_box[c].p = nullptr;
    m->J0[c] = g->jtot[c];
    m->E0[c] = g->jtot[c];
    PGPU_CUDA(cudaMalloc(&m->E0[c].p, m->E0[c].size() * sizeof(double)));
    PGPU_CUDA(cudaMemsetAsync(m->E0[c].p, 0, m->E0[c].size() * sizeof(double), st));
    elems += m->J0[c].size();
  }
  for (int k = 0; k < 9; ++k) elems += m->row_box[k / 3].size() * (size_t)(nc[k][0] * nc[k][1]);
  if (elems >= 0xffffffffull) {
    set_error("mass matrices of this box need %zu doubles; the run kernel addresses them with 32 bits", elems);
    return PGPU_ERR_ARG;
  }
  m->arena_elems = elems;
  PGPU_CUDA(cudaMalloc(&m->arena, elems * sizeof(double)));
  PGPU_CUDA(cudaMemsetAsync(m->arena, 0, elems * sizeof(double), st));
  {
    size_t at = 0;
    for (int c = 0; c < 3; ++c) {
      m->J0[c].p = m->arena + at;
      at += m->J0[c].size();
    }
    for (int k = 0; k < 9; ++k) {
      m->sigma[k] = m->arena + at;
      at += m->row_box[k / 3].size() * (size_t)(nc[k][0] * nc[k][1]);
    }
  }
  {
    const bool two = g->desc.D == 2;
    const int nslot = (two ? (int)MM_EPL : (int)MM1_EPL) * 32;
    MMEntry tab[MM_EPL * 32];
    if ((two ? build_table(m->mX, tab) : build_table_1d(m->mX, tab)) != (two ? (int)MM_NENT : (int)MM1_NENT)) {
      set_error("internal: mass-matrix product table has the wrong length");
      return PGPU_ERR_STATE;
    }
    PGPU_CUDA(cudaMalloc(&m->table_d, sizeof(MMEntry) * nslot));
    PGPU_CUDA(cudaMemcpy(m->table_d, tab, sizeof(MMEntry) * nslot, cudaMemcpyHostToDevice));
    {
      const MMSet T = make_set(g, m);
      MMFlush fl[MM_EPL * 32];
      for (int e = 0; e < nslot; ++e) {
        const MMEntry &en = tab[e];
        MMFlush f;
        f.off0 = 0;
        f.meta = -1;
        if (en.arr < 12) {
          long off;
          int row;
          if (en.arr < 9) {
            const MMView &v = T.s[en.arr];
            row = en.arr / 3;
            off = (v.p - m->arena) + (long)(en.di - v.lo0) + (long)(en.dj - v.lo1) * v.n0 + (long)en.nc0 * v.plane;
          } else {
            const FabView &v = T.J[en.arr - 9];
            row = en.arr - 9;
            off = (v.p - m->arena) + (long)(en.di - v.lo0) + (long)(en.dj - v.lo1) * v.n0;
          }
          f.off0 = (unsigned)off;
          f.meta = (en.base & 1) | (row << 1) | ((int)(unsigned char)en.ncs0 << 8) | ((int)(unsigned char)en.ncs1 << 16);
        }
        fl[e] = f;
      }
      PGPU_CUDA(cudaMalloc(&m->flush_d, sizeof(MMFlush) * nslot));
      PGPU_CUDA(cudaMemcpy(m->flush_d, fl, sizeof(MMFlush) * nslot, cudaMemcpyHostToDevice));
    }
    PGPU_CUDA(cudaMalloc(&m->defer_count, sizeof(unsigned)));
    if (two)
      PGPU_CUDA(cudaFuncSetAttribute(k_mm_cc1_2d_run, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(MM_WARPS * 32 * MM_REC * sizeof(double))));
  }
  if (ncomp_out)
    for (int k = 0; k < 9; ++k) {
      ncomp_out[2 * k] = nc[k][0];
      ncomp_out[2 * k + 1] = nc[k][1];
    }
  return 0;
}

int pgpu_mass_matrices_zero(pgpu_grid_t g) {
  NEED_MM(g);
  MassMatrices *m = mm_of(g);
  cudaStream_t st = ctx().stream;
  PGPU_CUDA(cudaMemsetAsync(m->arena, 0, m->arena_elems * sizeof(double), st));
  return 0;
}

int pgpu_accumulate_mass_matrices(pgpu_species_t s, double dt) {
  if (!s) return PGPU_ERR_ARG;
  pgpu_grid_s *g = s->grid;
  NEED_MM(g);
  if (s->desc.charge == 0.0 || s->n == 0) return 0;   // PicChargedSpecies.cpp:3687
  if (s->desc.interp_J != CC1) {
    set_error("mass matrices are implemented for CC1 current interpolation only");
    return PGPU_ERR_ARG;
  }
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  MassMatrices *m = mm_of(g);
  Context &c = ctx();
  const GeoAny ga = species_geo(s);
  const double cnormDt = dt * s->desc.cvac_norm;
  MMParams prm;
  prm.qovs = s->desc.charge / g->desc.volume_scale;
  prm.alphas = s->desc.fnorm_const * cnormDt / 2.0;
  prm.volume = (g->desc.D == 1) ? ga.dx[0] : ga.dx[0] * ga.dx[1];
  prm.anticyclic = 1;
  prm.rel = s->desc.relativistic;
  prm.mX = m->mX;
  const MMSet T = make_set(g, m);
  const bool generic_only = c.deposit_mode == 0 || !c.use_fast_cc1;
  auto chunk_for = [&](long ntiles, int minblocks) {
    // chunk length: long enough that few runs are cut (a cut run costs a full set of REDs), short enough that every
    // SM gets several chunks per resident warp
    const long resident = (long)c.sm_count * minblocks * MM_WARPS;
    int chunk = (int)std::max<long>(1, std::min<long>(MM_CHUNK_MAX, ntiles / (resident * 4)));
    if (const char *e = getenv("PGPU_MM_CHUNK")) chunk = std::max(1, std::min((int)MM_CHUNK_MAX, atoi(e)));   // tests, tuning
    return chunk;
  };
  if (!generic_only) {
    if (m->defer_cap < (size_t)s->n) {
      if (m->defer_list) cudaFree(m->defer_list);
      m->defer_cap = s->cap ? s->cap : (size_t)s->n;
      PGPU_CUDA(cudaMalloc(&m->defer_list, m->defer_cap * sizeof(int)));
    }
    PGPU_CUDA(cudaMemsetAsync(m->defer_count, 0, sizeof(unsigned), c.stream));
  }
  if (g->desc.D == 1) {
    const Geo<1> g1 = make_geo<1>(ga);
    if (generic_only) {
      KTimer t("mass_matrix_generic");
      const unsigned nb = (unsigned)((s->n + 127) / 128);
      k_mm<1><<<nb, 128, 0, c.stream>>>(s->ptrs(), s->n, g1, T, prm, c.d_counters, nullptr, nullptr);
      return 0;
    }
    {
      KTimer t("mass_matrix_run");
      const long ntiles = (s->n + 31) / 32;
      const int chunk = chunk_for(ntiles, 8);
      const long chunks = (ntiles + chunk - 1) / chunk;
      const unsigned nb = (unsigned)std::min<long>((chunks + MM_WARPS - 1) / MM_WARPS, (long)c.sm_count * 64);
      const size_t smem = MM_WARPS * 32 * MM1_REC * sizeof(double);
      k_mm_cc1_1d_run<<<nb, 32 * MM_WARPS, smem, c.stream>>>(s->ptrs(), s->n, g1, T, prm,
                                                             static_cast<const MMEntry *>(m->table_d),
                                                             static_cast<const MMFlush *>(m->flush_d), c.d_counters,
                                                             m->defer_list, m->defer_count, chunk);
    }
    {
      KTimer t("mass_matrix_deferred");
      k_mm<1><<<(unsigned)(c.sm_count * 4), 128, 0, c.stream>>>(s->ptrs(), s->n, g1, T, prm, c.d_counters,
                                                                  m->defer_list, m->defer_count);
    }
    return 0;
  }
  const Geo<2> g2 = make_geo<2>(ga);
  if (generic_only) {
    KTimer t("mass_matrix_generic");
    const unsigned nb = (unsigned)((s->n + 127) / 128);
    k_mm<2><<<nb, 128, 0, c.stream>>>(s->ptrs(), s->n, g2, T, prm, c.d_counters, nullptr, nullptr);
    return 0;
  }
  {
    KTimer t("mass_matrix_run");
    const long ntiles = (s->n + 31) / 32;
    const int chunk = chunk_for(ntiles, MM_MINBLOCKS);
    const long chunks = (ntiles + chunk - 1) / chunk;
    const unsigned nb = (unsigned)std::min<long>((chunks + MM_WARPS - 1) / MM_WARPS, (long)c.sm_count * 64);
    const size_t smem = MM_WARPS * 32 * MM_REC * sizeof(double);
    k_mm_cc1_2d_run<<<nb, 32 * MM_WARPS, smem, c.stream>>>(s->ptrs(), s->n, g2, T, prm,
                                                           static_cast<const MMEntry *>(m->table_d),
                                                           static_cast<const MMFlush *>(m->flush_d), c.d_counters,
                                                           m->defer_list, m->defer_count, chunk);
  }
  {
    KTimer t("mass_matrix_deferred");
    k_mm<2><<<(unsigned)(c.sm_count * 4), 128, 0, c.stream>>>(s->ptrs(), s->n, g2, T, prm, c.d_counters,
                                                                m->defer_list, m->defer_count);
  }
  return 0;
}

int pgpu_mass_matrices_save_E0(pgpu_grid_t g) {
  if (!g) return PGPU_ERR_ARG;
  NEED_MM(g);
  MassMatrices *m = mm_of(g);
  if (fields_wait(g)) return PGPU_ERR_CUDA;
  for (int c = 0; c < 3; ++c)
    PGPU_CUDA(cudaMemcpyAsync(m->E0[c].p, g->field[c].p, m->E0[c].size() * sizeof(double), cudaMemcpyDeviceToDevice,
                              ctx().stream));
  m->have_E0 = true;
  return 0;
}

int pgpu_compute_J_from_mass_matrices(pgpu_grid_t g) {
  NEED_MM(g);
  MassMatrices *m = mm_of(g);
  if (!m->have_E0) {
    set_error("pgpu_mass_matrices_save_E0 has not been called");
    return PGPU_ERR_STATE;
  }
  const MMSet T = make_set(g, m);
  const int D = g->desc.D;
  for (int row = 0; row < 3; ++row) {
    ContractArgs A;
    A.D = D;
    for (int e = 0; e < 3; ++e) {
      const int k = 3 * row + e;
      A.s[e] = T.s[k];
      A.E[e] = g->field[e].view();
      A.E0[e] = m->E0[e].view();
      A.N0[e] = m->ncomp[k][0];
      A.N1[e] = (D >= 2) ? m->ncomp[k][1] : 1;
      A.o0[e] = (m->ncomp[k][0] - 1) / 2;
      A.o1[e] = (D >= 2) ? (m->ncomp[k][1] - 1) / 2 : 0;
    }
    // FieldsF.ChF:34-39, 172-177, 310-315: Nc/2 instead of (Nc-1)/2 where the row and the column differ in centring
    if (row == 0 && D >= 2) A.o1[1] = m->ncomp[XY][1] / 2;
    if (row == 1) A.o0[0] = m->ncomp[YX][0] / 2;
    if (row == 2) {
      A.o0[0] = m->ncomp[ZX][0] / 2;
      if (D >= 2) A.o1[1] = m->ncomp[ZY][1] / 2;
    }
    A.J0 = m->J0[row].view();
    A.J = g->jtot[row].view();
    const long total = (long)A.J.n0 * A.J.n1;
    KTimer t("mass_matrix_contract");
    k_mm_contract<<<(unsigned)((total + 255) / 256), 256, 0, ctx().stream>>>(A);
  }
  return 0;
}

int pgpu_mass_matrices_ncomp(pgpu_grid_t g, int *ncomp_out) {
  NEED_MM(g);
  MassMatrices *m = mm_of(g);
  for (int k = 0; k < 9; ++k) {
    ncomp_out[2 * k] = m->ncomp[k][0];
    ncomp_out[2 * k + 1] = m->ncomp[k][1];
  }
  return 0;
}

int pgpu_mass_matrix_get(pgpu_grid_t g, int which, double *data, const int *lo, const int *hi, int ncomp) {
  if (!g) return PGPU_ERR_ARG;
  NEED_MM(g);
  MassMatrices *m = mm_of(g);
  if (which < 0 || which >= 9 || !data) return PGPU_ERR_ARG;
  const DeviceFab &b = m->row_box[which / 3];
  for (int d = 0; d < g->desc.D; ++d)
    if (lo[d] != b.lo[d] || hi[d] != b.hi[d]) {
      set_error("sigma %d: bounds [%d:%d] in dir %d do not match the device box [%d:%d]", which, lo[d], hi[d], d,
                b.lo[d], b.hi[d]);
      return PGPU_ERR_ARG;
    }
  if (ncomp != m->ncomp[which][0] * m->ncomp[which][1]) {
    set_error("sigma %d has %d components, not %d", which, m->ncomp[which][0] * m->ncomp[which][1], ncomp);
    return PGPU_ERR_ARG;
  }
  PGPU_CUDA(cudaMemcpyAsync(data, m->sigma[which], b.size() * (size_t)ncomp * sizeof(double), cudaMemcpyDeviceToHost,
                            ctx().stream));
  // synchronises and reports a crossing / bounds error of the deposit kernels
  return pgpu_picard_totals(nullptr, nullptr, nullptr, 0);
}

int pgpu_mass_matrix_J0_get(pgpu_grid_t g, int comp, double *data, const int *lo, const int *hi) {
  if (!g) return PGPU_ERR_ARG;
  NEED_MM(g);
  if (comp < 0 || comp >= 3) return PGPU_ERR_ARG;
  int rc = copy_fab_to_host(mm_of(g)->J0[comp], g->desc.D, data, lo, hi);
  if (rc) return rc;
  return pgpu_picard_totals(nullptr, nullptr, nullptr, 0);
}

}  // extern "C"
