// pgpu_bins.cu -- cell sort (binTheParticles), cell moments, Debye length, charge
// deposit, boundary conditions and the particle reductions.
//
// The reference rebuilds per-cell linked lists of particle pointers every scatter
// call (PicChargedSpecies::binTheParticles, PicChargedSpecies.cpp:1913-1947 +
// BinFab::locateBin, BinFabImplem.H:582-594).  Here the same cell index is the key
// of a counting sort that physically reorders the SoA arrays, so that the particles
// of a cell are contiguous (warp-contiguous for the collision and moment kernels)
// and the gather/deposit kernels see cell-coherent warps.
#include <algorithm>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "pgpu_internal.h"

namespace pgpu {

static inline unsigned nb(long n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

struct BoxInfo {
  int D;
  double le[2], dx[2];
  int lo[2], n[2];
  int ncell;
};

static BoxInfo box_info(const pgpu_grid_s *g) {
  BoxInfo b;
  b.D = g->desc.D;
  for (int d = 0; d < 2; ++d) {
    b.le[d] = g->geo.le[d];
    b.dx[d] = g->geo.dx[d];
    b.lo[d] = (d < b.D) ? g->desc.box_lo[d] : 0;
    b.n[d] = g->nbox[d];
  }
  b.ncell = (int)g->ncell_box;
  return b;
}

// BinFab::locateBin: thisPos -= origin; thisPos /= dx; (int)floor(thisPos)
__device__ __forceinline__ int locate_bin(double x, double le, double dx) {
  return __double2int_rd(__ddiv_rn(__dsub_rn(x, le), dx));
}

// Sort key = 4*cell + quadrant.  `cell` is the reference's bin (locate_bin); the quadrant
// says which half of the cell the particle sits in per direction, i.e. which cell of the
// half-shifted grid that the CC1 deposit segments on -- particles of one cell stay
// contiguous (what the collision and moment kernels need) and, inside the cell, particles
// that deposit onto the same node set are contiguous too (what pgpu_advance_cc1.cu sums
// over).  Outcasts (outside the box) get the last key.
// dual != 0: key = index of the DUAL cell (the cell of the half-shifted grid CC1 segments on) = (c0 + q0) + (c1 + q1)
// (n0 + 1): all particles that share the 21 deposit nodes are then ONE run (four times longer than the runs of
// 4*cell + quadrant, where they sit in four primal cells); pgpu_sort_for_locality
// count != nullptr: the histogram of the counting sort in the same pass (warp-aggregated); iota may be nullptr
__global__ void k_cell_key(const double *x0, const double *x1, long n, BoxInfo b, int *key, int *iota, int dual,
                           int *count) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int k = -1;
  if (i < n) {
    const double xa = x0[i];
    const int g0 = locate_bin(xa, b.le[0], b.dx[0]);
    int q = (__dsub_rn(xa, b.le[0]) - ((double)g0 + 0.5) * b.dx[0]) >= 0.0 ? 1 : 0;
    int c0 = g0 - b.lo[0], c1 = 0;
    if (b.D == 2) {
      const double xc = x1[i];
      const int g1 = locate_bin(xc, b.le[1], b.dx[1]);
      q |= (__dsub_rn(xc, b.le[1]) - ((double)g1 + 0.5) * b.dx[1]) >= 0.0 ? 2 : 0;
      c1 = g1 - b.lo[1];
    }
    const bool inside = c0 >= 0 && c0 < b.n[0] && c1 >= 0 && c1 < b.n[1];
    if (dual) {
      const int m0 = b.n[0] + 1, m1 = (b.D == 2) ? b.n[1] + 1 : 1;
      k = inside ? (c0 + (q & 1)) + (c1 + (q >> 1)) * m0 : m0 * m1;
    } else {
      k = inside ? 4 * (c0 + c1 * b.n[0]) + q : 4 * b.ncell;   // else the outcast bin
    }
    key[i] = k;
    if (iota) iota[i] = (int)i;
  }
  if (count) {
    const unsigned grp = __match_any_sync(0xffffffffu, k);
    if (k >= 0 && (__ffs(grp) - 1) == (int)(threadIdx.x & 31)) atomicAdd(count + k, __popc(grp));
  }
}

__global__ void k_cell_ijk(const double *x0, const double *x1, long n, BoxInfo b, int *out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = locate_bin(x0[i], b.le[0], b.dx[0]);
  if (b.D == 2) out[n + i] = locate_bin(x1[i], b.le[1], b.dx[1]);
}

// ---- counting sort of the bin keys -------------------------------------------------------------------------------
// Particles move at most a cell or so between two sorts, and there are far fewer bins than particles (tens to hundreds
// of particles per bin): a histogram, a scan over the bins and one scatter of the particle indices replace the radix
// sort of (key, index) pairs (four passes over 8 bytes per particle plus its scratch).  Consecutive particles mostly
// share their bin, so the lanes of a warp that hit one bin are counted by their leader (__match_any_sync): one
// atomic per distinct bin of a warp, and lanes of a bin keep their relative order.
// perm[dst] = i, key_sorted[dst] = key[i] with dst = start[key] + (position among the particles of the bin seen so far)
__global__ void k_bin_scatter(const int *key, long n, const int *start, int *cursor, int *perm, int *key_sorted) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int k = i < n ? key[i] : -1;
  const unsigned grp = __match_any_sync(0xffffffffu, k);
  const int leader = __ffs(grp) - 1;
  int base = 0;
  if (k >= 0 && leader == lane) base = atomicAdd(cursor + k, __popc(grp));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (k >= 0) {
    const int dst = start[k] + base + __popc(grp & ((1u << lane) - 1u));
    perm[dst] = (int)i;
    if (key_sorted) key_sorted[dst] = k;
  }
}

// The scatter hands out the slots of a bin in arrival order, which differs from run to run.  The collision kernels pair
// the particles of a bin by storage order, so the bin sort puts every bin back into ascending source index (= what the
// stable radix sort gave).  Bins hold tens of particles: one thread sorts one bin of <= 32 entries by insertion (entries
// of one source warp already arrive in order, so few moves); larger bins go on a list for k_bin_canon_big.
__global__ void k_bin_canon(const int *start, int nbins, int *perm, int *biglist, int *nbig) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbins) return;
  const int s = start[b], n = start[b + 1] - s;
  if (n < 2) return;
  if (n > 32) {
    biglist[atomicAdd(nbig, 1)] = b;
    return;
  }
  int a[32];
  for (int i = 0; i < n; ++i) a[i] = perm[s + i];
  bool moved = false;
  for (int i = 1; i < n; ++i) {
    const int v = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > v) {
      a[j + 1] = a[j];
      --j;
      moved = true;
    }
    a[j + 1] = v;
  }
  if (moved)
    for (int i = 0; i < n; ++i) perm[s + i] = a[i];
}
// ascending bitonic sort of up to 32 R distinct ints held R per lane (element e = lane + 32 r), padded with INT_MAX
template <int R>
__device__ __forceinline__ void warp_sort_ints(int *p, int n, int lane) {
  int v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int k = lane + 32 * r;
    v[r] = k < n ? p[k] : 0x7fffffff;
  }
  // already ascending (particles that did not change bin since the last sort mostly are): nothing to do
  bool sorted = true;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    int nxt = __shfl_down_sync(0xffffffffu, v[r], 1);
    if (lane == 31) nxt = 0x7fffffff;
    sorted = sorted && v[r] <= nxt;
  }
#pragma unroll
  for (int r = 0; r + 1 < R; ++r) {
    const int first_next = __shfl_sync(0xffffffffu, v[r + 1], 0);
    if (lane == 31) sorted = sorted && v[r] <= first_next;
  }
  if (__all_sync(0xffffffffu, sorted)) return;
#pragma unroll
  for (int k2 = 2; k2 <= 32 * R; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int dr = j >> 5;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (r & dr) continue;
          const bool up = (((lane + 32 * r) & k2) == 0);
          const int x = v[r], y = v[r | dr];
          const int lo = min(x, y), hi = max(x, y);
          v[r] = up ? lo : hi;
          v[r | dr] = up ? hi : lo;
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool up = (((lane + 32 * r) & k2) == 0);
          v[r] = (lower == up) ? min(v[r], o) : max(v[r], o);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int k = lane + 32 * r;
    if (k < n) p[k] = v[r];
  }
}
// listed bins of 33 .. 4096 entries, one warp each: up to 256 entries in registers through a bitonic network, beyond that
// the rank of every entry against the bin into a scratch copy, copied back (bigger bins -- everything in one cell -- keep
// the arrival order)
__global__ void k_bin_canon_big(const int *start, const int *biglist, const int *nbig, int *perm, int *scratch) {
  const int lane = threadIdx.x & 31, nwarps = (int)((gridDim.x * blockDim.x) >> 5), nl = *nbig;
  for (int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); l < nl; l += nwarps) {
    const int b = biglist[l];
    const int s = start[b], n = start[b + 1] - s;
    if (n <= 64) {
      warp_sort_ints<2>(perm + s, n, lane);
      continue;
    }
    if (n <= 128) {
      warp_sort_ints<4>(perm + s, n, lane);
      continue;
    }
    if (n <= 256) {
      warp_sort_ints<8>(perm + s, n, lane);
      continue;
    }
    if (n > 4096) continue;
    for (int e = lane; e < n; e += 32) {
      const int v = perm[s + e];
      int rank = 0;
      for (int q = 0; q < n; ++q) rank += (perm[s + q] < v) ? 1 : 0;
      scratch[s + rank] = v;
    }
    __syncwarp();
    for (int e = lane; e < n; e += 32) perm[s + e] = scratch[s + e];
    __syncwarp();
  }
}

// cell_start[c] = first sorted position whose cell is >= c, for c = 0..nbins (nbins = ncell+1
// bins incl. the outcast bin; cell_start[nbins] = n).  One thread per sorted position fills
// the entries of every cell that begins at its position (empty cells included).
__global__ void k_cell_starts(const int *sorted_key, long n, int nbins, int *cell_start) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int cur = (i < n) ? (sorted_key[i] >> 2) : nbins;
  const int prev = (i > 0) ? (sorted_key[i - 1] >> 2) : -1;
  for (int c = prev + 1; c <= cur; ++c) cell_start[c] = (int)i;
}

// the same from the bin starts of the counting sort (bins = 4 quadrants per cell + the outcast bin; start[nb_bins] = n)
__global__ void k_cell_starts_from_bins(const int *start, int nb_bins, int nbins, int *cell_start) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= nbins) cell_start[c] = start[min(4 * c, nb_bins)];
}

// out[a][i] = in[a][perm[i]] for up to 4 arrays per launch (perm is read once per particle)
struct PermuteSet {
  const double *in[4];
  double *out[4];
  int count;
};
// four particles per thread, a block-width apart: all the loads of a thread are in flight before its first store
constexpr int PERMUTE_UNROLL = 4;
__global__ void __launch_bounds__(256) k_permute(PermuteSet ps, const int *perm, long n) {
  const long base = (long)blockIdx.x * (256 * PERMUTE_UNROLL) + threadIdx.x;
  int src[PERMUTE_UNROLL];
#pragma unroll
  for (int u = 0; u < PERMUTE_UNROLL; ++u) {
    const long i = base + u * 256;
    src[u] = i < n ? perm[i] : -1;
  }
  double val[4][PERMUTE_UNROLL];
#pragma unroll
  for (int a = 0; a < 4; ++a)
    if (a < ps.count) {
#pragma unroll
      for (int u = 0; u < PERMUTE_UNROLL; ++u)
        if (src[u] >= 0) val[a][u] = ps.in[a][src[u]];
    }
#pragma unroll
  for (int a = 0; a < 4; ++a)
    if (a < ps.count) {
#pragma unroll
      for (int u = 0; u < PERMUTE_UNROLL; ++u)
        if (src[u] >= 0) __stcs(ps.out[a] + base + u * 256, val[a][u]);
    }
}

// set{Number,Momentum,Energy}DensityFromBinFab (PicChargedSpecies.cpp:2881-3047):
// one warp per cell over the cell-sorted arrays
__global__ void k_cell_moments(const int *cell_start, int ncell, const double *w, const double *v0,
                               const double *v1, const double *v2, double kn, double km, double ke,
                               double *dens, double *mom, double *ene) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const int s = cell_start[cell], e = cell_start[cell + 1];
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int k = s + lane; k < e; k += 32) {
    const double wp = w[k], u0 = v0[k], u1 = v1[k], u2 = v2[k];
    a[0] += wp;
    a[1] += wp * u0;
    a[2] += wp * u1;
    a[3] += wp * u2;
    a[4] += wp * u0 * u0;
    a[5] += wp * u1 * u1;
    a[6] += wp * u2 * u2;
  }
#pragma unroll
  for (int q = 0; q < 7; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
  if (lane == 0) {
    dens[cell] = a[0] * kn;
    for (int q = 0; q < 3; ++q) {
      mom[q * ncell + cell] = a[1 + q] * km;
      ene[q * ncell + cell] = a[4 + q] * ke;
    }
  }
}

// PicSpeciesInterface::setDebyeLength (PicSpeciesInterface.cpp:1627-1721), one species
__global__ void k_debye_accumulate(int ncell, const double *dens, const double *mom, const double *ene,
                                   double mass, double charge, double *sum_inv) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const double PI = 3.14159265358979323846, CVAC = 2.99792458e+08, ME = 9.10938370e-31, QE = 1.60217663e-19;
  const double MU0 = 4.0 * PI * 1.0e-7, EP0 = 1.0 / CVAC / CVAC / MU0;
  const double HBAR = 6.62607015e-34 / (2.0 * PI);
  const double EV_PER_JOULE = 1.0 / QE;
  const double mcSq_eV = ME * CVAC * CVAC * EV_PER_JOULE;
  const double Aconst = EP0 / QE / (charge * charge);
  const double N = dens[c];
  if (N == 0.0) return;
  const double R = 1.0 / cbrt(4.0 / 3.0 * PI * N);
  const double rho = N * mass;
  const double ux = mom[c], uy = mom[ncell + c], uz = mom[2 * ncell + c];
  const double meanE = (ux * ux + uy * uy + uz * uz) / rho / 2.0;
  const double EF_eV = HBAR * HBAR / (2.0 * ME * mass) * pow(3.0 * PI * PI * N, 2.0 / 3.0) * EV_PER_JOULE;
  double rhoE = 0.0;
  for (int dir = 0; dir < 3; ++dir) rhoE += ene[dir * ncell + c];
  double T_eV = 2.0 / 3.0 * (rhoE - meanE) / N * mcSq_eV;
  T_eV = fmax(T_eV, 0.01);
  const double LDe_sq = fmax(Aconst * (T_eV + 2.0 / 3.0 * EF_eV) / N, R * R);
  sum_inv[c] += 1.0 / LDe_sq;
}
__global__ void k_debye_finish(int ncell, double *a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < ncell) a[c] = 1.0 / sqrt(a[c]);
}

// MeshInterp::deposit -> cic_deposit / tsc_deposit (MeshInterpF.ChF:206-255, 739-800)
template <int D, bool X>
__global__ void k_deposit_rho(const double *x0, const double *x1, const double *w, long n, Geo<D> g,
                              int interp, int stag0, int stag1, FabView rho, double volume, Counters *cnt) {
  typedef M<X> m;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned err = 0;
  if (i < n) {
    const double xp[2] = {x0[i], D == 2 ? x1[i] : 0.0};
    const int stag[2] = {stag0, stag1};
    const double particle_rho = __ddiv_rn(w[i], volume);
    const int npt = (interp == TSC) ? 3 : 2;
    int index[D];
    double wt[D][3];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      // off = 0.5*dx*(1-stag)
      const double off = m::mul(m::mul(0.5, g.dx[d]), __dsub_rn(1.0, (double)stag[d]));
      double a = __dsub_rn(__dsub_rn(xp[d], g.le[d]), (interp == TSC) ? 0.5 * g.dx[d] : off);
      if (interp == TSC) a = __dsub_rn(a, off);
      index[d] = floor_div<X>(a, g.dx[d], g.rdx[d]);
      for (int q = 0; q < npt; ++q) {
        const double l = m::add(m::sub(m::add(m::mul((double)(index[d] + q), g.dx[d]), off), xp[d]), g.le[d]);
        const double r = fabs(m::div(l, g.dx[d], g.rdx[d]));
        wt[d][q] = (interp == TSC) ? tsc_w<X>(r) : m::sub(1.0, r);
      }
    }
    for (int q0 = 0; q0 < npt; ++q0) {
      const int ii = index[0] + q0;
      const unsigned a = (unsigned)(ii - rho.lo0);
      if (D == 1) {
        if (a < (unsigned)rho.n0) atomicAdd(rho.p + a, m::mul(particle_rho, wt[0][q0]));
        else err |= ERRBIT_BOUNDS;
      } else {
        for (int q1 = 0; q1 < npt; ++q1) {
          const unsigned b = (unsigned)(index[D - 1] + q1 - rho.lo1);
          if (a < (unsigned)rho.n0 && b < (unsigned)rho.n1)
            atomicAdd(rho.p + (a + (size_t)b * rho.n0), m::mul(particle_rho, m::mul(wt[0][q0], wt[D - 1][q1])));
          else err |= ERRBIT_BOUNDS;
        }
      }
    }
  }
  err = __reduce_or_sync(0xffffffffu, err);
  if ((threadIdx.x & 31) == 0 && err) atomicOr(&cnt->err, err);
}

// advanceVelocities_2ndHalf + advancePositions_2ndHalf + periodic applyBCs in one pass
// (PicChargedSpecies.cpp:1136-1244, 997-1025; PicChargedSpeciesBC.cpp:738-765): the three
// calls always follow each other at the end of an implicit step
// (PICTimeIntegrator_EM_ThetaImplicit.cpp:311-319).  Same arithmetic as the separate kernels.
struct FinishArgs {
  double *x[2], *xold[2], *v[3];
  const double *vold[3];
  long n;
  int D, motion, forces;
  int periodic[2];
  double left[2], right[2];
};
__global__ void k_finish_step(FinishArgs a) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  if (a.forces) {
#pragma unroll
    for (int c = 0; c < 3; ++c) a.v[c][i] = __dsub_rn(__dmul_rn(2.0, a.v[c][i]), a.vold[c][i]);
  }
  if (a.motion) {
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      if (d >= a.D) break;
      double xo = a.xold[d][i];
      double xp = __dsub_rn(__dmul_rn(2.0, a.x[d][i]), xo);
      bool ch = false;
      if (a.periodic[d]) {
        const double Lbox = __dsub_rn(a.right[d], a.left[d]);
        if (xp < a.left[d]) {
          xp = __dadd_rn(xp, Lbox);
          xo = __dadd_rn(xo, Lbox);
          ch = true;
        }
        if (xp >= a.right[d]) {
          xp = __dsub_rn(xp, Lbox);
          xo = __dsub_rn(xo, Lbox);
          ch = true;
        }
      }
      a.x[d][i] = xp;
      if (ch) a.xold[d][i] = xo;
    }
  }
}

// PicChargedSpeciesBC::enforcePeriodic (PicChargedSpeciesBC.cpp:738-765)
__global__ void k_bc_periodic(double *x, double *xold, long n, double left, double right) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double Lbox = __dsub_rn(right, left);
  double xp = x[i], xo = xold[i];
  bool ch = false;
  if (xp < left) {
    xp = __dadd_rn(xp, Lbox);
    xo = __dadd_rn(xo, Lbox);
    ch = true;
  }
  if (xp >= right) {
    xp = __dsub_rn(xp, Lbox);
    xo = __dsub_rn(xo, Lbox);
    ch = true;
  }
  if (ch) {
    x[i] = xp;
    xold[i] = xo;
  }
}

// PicChargedSpeciesBC::symmetry_Lo / symmetry_Hi (PicChargedSpeciesBC.cpp:808-870)
__global__ void k_bc_symmetry(double *x, double *xold, double *v, double *vold, long n, double left,
                              double right, int do_lo, int do_hi) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double xp = x[i], xo = xold[i], vp = v[i], vo = vold[i];
  bool ch = false;
  if (do_lo && xp <= left) {
    xp = __dsub_rn(__dmul_rn(2., left), xp);
    vp = -vp;
    xo = __dsub_rn(__dmul_rn(2., left), xo);
    vo = -vo;
    ch = true;
  }
  if (do_hi && xp >= right) {
    xp = __dsub_rn(__dmul_rn(2., right), xp);
    vp = -vp;
    xo = __dsub_rn(__dmul_rn(2., right), xo);
    vo = -vo;
    if (xp == right) xp = __dmul_rn(0.999999999, right);
    ch = true;
  }
  if (ch) {
    x[i] = xp;
    xold[i] = xo;
    v[i] = vp;
    vold[i] = vo;
  }
}

// block reduction helpers for the global reductions
__global__ void k_max_dtinv(const double *v0, const double *v1, const double *v2, long n, int D, double dx0,
                            double dx1, int rel, unsigned long long *out_bits) {
  double mx = 0.0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    double gammap = 1.0;
    if (rel) {   // setStableDt :1890-1895: |up/gammap| / dX
      const double u[3] = {v0[i], v1[i], v2[i]};
      gammap = gamma_sum_first<true>(u);
    }
    mx = fmax(mx, __ddiv_rn(fabs(__ddiv_rn(v0[i], gammap)), dx0));
    if (D == 2) mx = fmax(mx, __ddiv_rn(fabs(__ddiv_rn(v1[i], gammap)), dx1));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(mx));
}

__global__ void k_global_moments(const double *w, const double *v0, const double *v1, const double *v2,
                                 long n, double *out7, int rel) {
  double a[7] = {0, 0, 0, 0, 0, 0, 0};
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const double wp = w[i], u0 = v0[i], u1 = v1[i], u2 = v2[i];
    a[0] += wp;
    a[1] += wp * u0;
    a[2] += wp * u1;
    a[3] += wp * u2;
    // relativistic build (:4095-4098): energy = sum wp*gbsq*2/(gammap+1), split over the components here
    double ke = 1.0;
    if (rel) ke = 2.0 / (sqrt(1.0 + (u0 * u0 + u1 * u1 + u2 * u2)) + 1.0);
    a[4] += wp * u0 * u0 * ke;
    a[5] += wp * u1 * u1 * ke;
    a[6] += wp * u2 * u2 * ke;
  }
#pragma unroll
  for (int q = 0; q < 7; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out7 + q, a[q]);
  }
}

}  // namespace pgpu

using namespace pgpu;

#define NEED_INIT()                               \
  do {                                            \
    if (!ctx().inited) {                          \
      set_error("pgpu_init has not been called"); \
      return PGPU_ERR_STATE;                      \
    }                                             \
  } while (0)

namespace pgpu {
// out-of-place gather of a list of particle arrays through the spare pool
static int gather_arrays(pgpu_species_s *s, std::vector<double **> &arrs, const int *perm) {
  const long n = s->n;
  for (size_t a0 = 0; a0 < arrs.size(); a0 += 4) {
    PermuteSet ps;
    ps.count = (int)std::min<size_t>(4, arrs.size() - a0);
    for (int a = 0; a < ps.count; ++a) {
      ps.in[a] = *arrs[a0 + a];
      ps.out[a] = s->spare[a];
    }
    KTimer t("bin_permute");
    k_permute<<<nb(n, 256 * PERMUTE_UNROLL), 256, 0, ctx().stream>>>(ps, perm, n);
    for (int a = 0; a < ps.count; ++a) {
      double *old = *arrs[a0 + a];
      *arrs[a0 + a] = s->spare[a];
      s->spare[a] = old;
    }
  }
  return 0;
}

int materialize_old(pgpu_species_s *s, int keep) {
  // the deferred copy of updateOldParticlePositions / Velocities (an aliased group has no pending gather)
  if (s->xold_alias && !(keep & KEEP_XOLD_ALIAS)) {
    for (int d = 0; d < s->grid->desc.D; ++d)
      PGPU_CUDA(cudaMemcpyAsync(s->xold[d], s->x[d], s->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
    s->xold_alias = false;
  }
  if (s->vold_alias && !(keep & KEEP_VOLD_ALIAS)) {
    for (int q = 0; q < 3; ++q)
      PGPU_CUDA(cudaMemcpyAsync(s->vold[q], s->v[q], s->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx().stream));
    s->vold_alias = false;
  }
  if ((keep & KEEP_PENDING) || (!s->pos_old_pending && !s->vel_old_pending)) return 0;
  std::vector<double **> arrs;
  if (s->pos_old_pending)
    for (int d = 0; d < s->grid->desc.D; ++d) arrs.push_back(&s->xold[d]);
  if (s->vel_old_pending)
    for (int q = 0; q < 3; ++q) arrs.push_back(&s->vold[q]);
  s->pos_old_pending = s->vel_old_pending = false;
  if (s->n == 0) return 0;
  return gather_arrays(s, arrs, s->old_perm);
}
}  // namespace pgpu

extern "C" {

static int bin_impl(pgpu_species_t s, bool dual);
int pgpu_bin_particles(pgpu_species_t s) { return bin_impl(s, false); }
int pgpu_sort_for_locality(pgpu_species_t s) { return bin_impl(s, true); }

static int bin_impl(pgpu_species_t s, bool dual) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  Context &c = ctx();
  const pgpu_grid_s *g = s->grid;
  const BoxInfo b = box_info(g);
  const long n = s->n;
  const int nbins = b.ncell + 1;  // + outcast bin
  if (n == 0) {
    PGPU_CUDA(cudaMemsetAsync(s->cell_start, 0, (nbins + 1) * sizeof(int), c.stream));
    s->binned = !dual;
    return 0;
  }
  // scratch: sorted keys + the pool of spare particle arrays the gather writes into
  if (s->sort_cap < s->cap) {
    if (s->key_sorted) cudaFree(s->key_sorted);
    for (double *&p : s->spare)
      if (p) { cudaFree(p); p = nullptr; }
    PGPU_CUDA(cudaMalloc(&s->key_sorted, s->cap * sizeof(int)));
    for (int k = 0; k < 4; ++k) PGPU_CUDA(cudaMalloc(&s->spare[k], s->cap * sizeof(double)));
    s->sort_cap = s->cap;
  }
  int *iota = reinterpret_cast<int *>(s->tmp);  // tmp holds >= n doubles
  int nbits = 1;
  const long maxkey = dual ? (long)(b.n[0] + 1) * (b.D == 2 ? b.n[1] + 1 : 1) : 4L * b.ncell;
  while ((1L << nbits) <= maxkey) ++nbits;
  static const int sort_mode = [] { const char *e = getenv("PGPU_SORT"); return (e && !strcmp(e, "radix")) ? 1 : 0; }();
  if (sort_mode != 0) {
    KTimer t("bin_key");
    k_cell_key<<<nb(n), 256, 0, c.stream>>>(s->x[0], s->x[1], n, b, s->cell_key, iota, dual ? 1 : 0, nullptr);
  }
  if (sort_mode == 0) {
    // counting sort: bins = maxkey + 1 (the last one is the outcast bin)
    const long nb_bins = maxkey + 1;
    if (s->bin_count_cap < (size_t)(2 * nb_bins + 2)) {
      if (s->bin_count) cudaFree(s->bin_count);
      PGPU_CUDA(cudaMalloc(&s->bin_count, (size_t)(2 * nb_bins + 2) * sizeof(int)));
      s->bin_count_cap = (size_t)(2 * nb_bins + 2);
    }
    int *count = s->bin_count, *start = s->bin_count + nb_bins + 1;
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, count, start, (int)nb_bins + 1, c.stream);
    if (need > s->cub_bytes) {
      if (s->cub_tmp) cudaFree(s->cub_tmp);
      PGPU_CUDA(cudaMalloc(&s->cub_tmp, need));
      s->cub_bytes = need;
    }
    PGPU_CUDA(cudaMemsetAsync(count, 0, (size_t)(nb_bins + 1) * sizeof(int), c.stream));
    {
      KTimer t("bin_key");     // keys + histogram in one pass
      k_cell_key<<<nb(n), 256, 0, c.stream>>>(s->x[0], s->x[1], n, b, s->cell_key, nullptr, dual ? 1 : 0, count);
    }
    KTimer t("bin_sort");
    PGPU_CUDA(cub::DeviceScan::ExclusiveSum(s->cub_tmp, need, count, start, (int)nb_bins + 1, c.stream));   // start[nb_bins] = n
    // the cursors of the scatter, and in the slot behind them the counter of the big-bin list
    PGPU_CUDA(cudaMemsetAsync(count, 0, (size_t)(nb_bins + 1) * sizeof(int), c.stream));
    k_bin_scatter<<<nb(n), 256, 0, c.stream>>>(s->cell_key, n, start, count, s->perm, nullptr);
    if (!dual) {
      k_cell_starts_from_bins<<<nb(nbins + 1), 256, 0, c.stream>>>(start, (int)nb_bins, nbins, s->cell_start);
      // tmp holds n doubles: ints [0, n) are the scratch copy, ints [n, 2n) the list of big bins
      int *biglist = iota + n, *nbig = count + nb_bins;
      k_bin_canon<<<nb(nb_bins), 256, 0, c.stream>>>(start, (int)nb_bins, s->perm, biglist, nbig);
      k_bin_canon_big<<<c.sm_count * 8, 256, 0, c.stream>>>(start, biglist, nbig, s->perm, iota);
    }
  } else {
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, s->cell_key, s->key_sorted, iota, s->perm, (int)n, 0, nbits,
                                    c.stream);
    if (need > s->cub_bytes) {
      // sized for the capacity, not for today's count: with migration n drifts from step to step and a
      // cudaFree + cudaMalloc in the middle of a step stalls the whole device
      cub::DeviceRadixSort::SortPairs(nullptr, need, s->cell_key, s->key_sorted, iota, s->perm, (int)s->cap, 0, nbits,
                                      c.stream);
      if (s->cub_tmp) cudaFree(s->cub_tmp);
      PGPU_CUDA(cudaMalloc(&s->cub_tmp, need));
      s->cub_bytes = need;
    }
    KTimer t("bin_sort");
    PGPU_CUDA(cub::DeviceRadixSort::SortPairs(s->cub_tmp, need, s->cell_key, s->key_sorted, iota, s->perm, (int)n, 0,
                                              nbits, c.stream));
  }
  if (!dual && sort_mode != 0) {
    KTimer t("bin_starts");
    k_cell_starts<<<nb(n + 1), 256, 0, c.stream>>>(s->key_sorted, n, nbins, s->cell_start);
  }
  // gather the particle arrays into the sorted order, four arrays per launch (the old arrays
  // become the next spares); xold / vold are gathered lazily (materialize_old)
  // a still-pending gather of an earlier sort; an aliased old group stays aliased (x == xold in any order)
  if (materialize_old(s, KEEP_OLD_ALIASES)) return PGPU_ERR_CUDA;
  std::vector<double **> arrs;
  const int D = g->desc.D;
  for (int d = 0; d < D; ++d) arrs.push_back(&s->x[d]);
  for (int q = 0; q < 3; ++q) arrs.push_back(&s->v[q]);
  arrs.push_back(&s->w);
  arrs.push_back(reinterpret_cast<double **>(&s->id));
  if (gather_arrays(s, arrs, s->perm)) return PGPU_ERR_CUDA;
  if (!s->old_perm || s->old_perm_cap < s->cap) {
    if (s->old_perm) cudaFree(s->old_perm);
    PGPU_CUDA(cudaMalloc(&s->old_perm, s->cap * sizeof(int)));
    s->old_perm_cap = s->cap;
  }
  std::swap(s->perm, s->old_perm);
  s->pos_old_pending = !s->xold_alias;
  s->vel_old_pending = !s->vold_alias;
  s->binned = !dual;   // the per-cell lists (collisions, moments) exist only after the primal-cell sort
  return 0;
}

int pgpu_species_cell_index(pgpu_species_t s, int *cell) {
  NEED_INIT();
  if (!s || !cell) return PGPU_ERR_ARG;
  const long n = s->n;
  if (n == 0) return 0;
  const int D = s->grid->desc.D;
  int *d_out = nullptr;
  PGPU_CUDA(cudaMalloc(&d_out, (size_t)D * n * sizeof(int)));
  k_cell_ijk<<<nb(n), 256, 0, ctx().stream>>>(s->x[0], s->x[1], n, box_info(s->grid), d_out);
  PGPU_CUDA(cudaMemcpyAsync(cell, d_out, (size_t)D * n * sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
  PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
  cudaFree(d_out);
  return 0;
}

int pgpu_species_cell_offsets(pgpu_species_t s, long *offsets) {
  NEED_INIT();
  if (!s || !offsets) return PGPU_ERR_ARG;
  if (!s->binned) {
    set_error("species is not binned; call pgpu_bin_particles first");
    return PGPU_ERR_STATE;
  }
  const long nc = s->grid->ncell_box + 2;
  std::vector<int> h(nc);
  PGPU_CUDA(cudaMemcpyAsync(h.data(), s->cell_start, nc * sizeof(int), cudaMemcpyDeviceToHost, ctx().stream));
  PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
  for (long i = 0; i < nc - 1; ++i) offsets[i] = h[i];
  return 0;
}

int pgpu_set_moments_from_bins(pgpu_species_t s) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->binned) {
    set_error("species is not binned; call pgpu_bin_particles first");
    return PGPU_ERR_STATE;
  }
  const pgpu_grid_s *g = s->grid;
  const double dV_mapped = (g->desc.D == 1) ? g->geo.dx[0] : g->geo.dx[0] * g->geo.dx[1];
  const double dV_phys = dV_mapped * g->desc.volume_scale;
  const int ncell = (int)g->ncell_box;
  KTimer t("cell_moments");
  k_cell_moments<<<nb((long)ncell * 32), 256, 0, ctx().stream>>>(
      s->cell_start, ncell, s->w, s->v[0], s->v[1], s->v[2], 1.0 / dV_phys, s->desc.mass / dV_phys,
      0.5 * s->desc.mass / dV_phys, s->dens, s->mom, s->ene);
  return 0;
}

int pgpu_species_moments_get(pgpu_species_t s, double *dens, double *mom, double *ene) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  const long nc = s->grid->ncell_box;
  cudaStream_t st = ctx().stream;
  if (dens) PGPU_CUDA(cudaMemcpyAsync(dens, s->dens, nc * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (mom) PGPU_CUDA(cudaMemcpyAsync(mom, s->mom, 3 * nc * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (ene) PGPU_CUDA(cudaMemcpyAsync(ene, s->ene, 3 * nc * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int pgpu_debye_length(pgpu_grid_t g, pgpu_species_t *species, int nspecies, double *LDe) {
  NEED_INIT();
  if (!g) return PGPU_ERR_ARG;
  const int nc = (int)g->ncell_box;
  cudaStream_t st = ctx().stream;
  PGPU_CUDA(cudaMemsetAsync(g->debye, 0, nc * sizeof(double), st));
  for (int k = 0; k < nspecies; ++k) {
    pgpu_species_t s = species[k];
    if (s->desc.charge == 0.0) continue;
    KTimer t("debye");
    k_debye_accumulate<<<nb(nc), 256, 0, st>>>(nc, s->dens, s->mom, s->ene, s->desc.mass, s->desc.charge, g->debye);
  }
  {
    KTimer t("debye");
    k_debye_finish<<<nb(nc), 256, 0, st>>>(nc, g->debye);
  }
  if (LDe) {
    PGPU_CUDA(cudaMemcpyAsync(LDe, g->debye, nc * sizeof(double), cudaMemcpyDeviceToHost, st));
    PGPU_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

int pgpu_charge_density_deposit(pgpu_species_t s, const int *stag) {
  NEED_INIT();
  if (!s || !stag) return PGPU_ERR_ARG;
  pgpu_grid_s *g = s->grid;
  const int D = g->desc.D;
  Context &c = ctx();
  DeviceFab *fp = nullptr;
  int rc = grid_rho_fab(g, stag, &fp);
  if (rc) return rc;
  const DeviceFab &f = *fp;
  const int st2[2] = {f.stag[0], f.stag[1]};
  PGPU_CUDA(cudaMemsetAsync(f.p, 0, f.size() * sizeof(double), c.stream));
  const GeoAny ga = species_geo(s);
  const double volume = (D == 1) ? ga.dx[0] : ga.dx[0] * ga.dx[1];
  if (s->n > 0) {
    KTimer t("deposit_rho");
    if (D == 1) {
      if (c.exact)
        k_deposit_rho<1, true><<<nb(s->n), 256, 0, c.stream>>>(s->x[0], s->x[1], s->w, s->n, make_geo<1>(ga), s->desc.interp_N, st2[0], st2[1], f.view(), volume, c.d_counters);
      else
        k_deposit_rho<1, false><<<nb(s->n), 256, 0, c.stream>>>(s->x[0], s->x[1], s->w, s->n, make_geo<1>(ga), s->desc.interp_N, st2[0], st2[1], f.view(), volume, c.d_counters);
    } else {
      if (c.exact)
        k_deposit_rho<2, true><<<nb(s->n), 256, 0, c.stream>>>(s->x[0], s->x[1], s->w, s->n, make_geo<2>(ga), s->desc.interp_N, st2[0], st2[1], f.view(), volume, c.d_counters);
      else
        k_deposit_rho<2, false><<<nb(s->n), 256, 0, c.stream>>>(s->x[0], s->x[1], s->w, s->n, make_geo<2>(ga), s->desc.interp_N, st2[0], st2[1], f.view(), volume, c.d_counters);
    }
  }
  // this_rho.mult(m_charge/volume_scale); (cartesian Jacobian == 1)
  scale_fab(f, s->desc.charge / g->desc.volume_scale);
  // the ghost add-exchange of a box that is its own periodic neighbour; between boxes it is the halo plan's job
  if (fold_periodic(g, f)) return PGPU_ERR_CUDA;
  return 0;
}

int pgpu_charge_density_get(pgpu_grid_t g, const int *stag, double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!g || !stag || !data || !lo || !hi) return PGPU_ERR_ARG;
  const int D = g->desc.D;
  DeviceFab *fp = nullptr;
  int rc = grid_rho_fab(g, stag, &fp);
  if (rc) return rc;
  for (int k = 0; k < D; ++k)
    if (lo[k] != fp->lo[k] || hi[k] != fp->hi[k]) {
      set_error("charge density bounds do not match the ghosted box for this centring");
      return PGPU_ERR_ARG;
    }
  rc = copy_fab_to_host(*fp, D, data, lo, hi);
  if (rc) return rc;
  return pgpu_synchronize();
}

int pgpu_set_charge_density(pgpu_species_t s, const int *stag, double *data, const int *lo, const int *hi) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  DeviceFab *fp = nullptr;
  int rc = grid_rho_fab(s->grid, stag, &fp);
  if (rc) return rc;
  for (int k = 0; k < s->grid->desc.D; ++k)
    if (lo[k] != fp->lo[k] || hi[k] != fp->hi[k]) {
      set_error("charge density bounds do not match the ghosted box for this centring");
      return PGPU_ERR_ARG;
    }
  rc = pgpu_charge_density_deposit(s, stag);
  if (rc) return rc;
  return pgpu_charge_density_get(s->grid, stag, data, lo, hi);
}

int pgpu_apply_bcs(pgpu_species_t s, const int *bc_lo, const int *bc_hi) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  if (!s->desc.motion) return 0;
  const pgpu_grid_s *g = s->grid;
  const long n = s->n;
  if (n == 0 && s->n_inf == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  cudaStream_t st = ctx().stream;
  // the order of PicChargedSpeciesBC::apply (:161-239): symmetry, then outflow / inflow_outflow, then periodic
  for (int d = 0; d < g->desc.D; ++d) {
    const double left = g->geo.le[d], right = g->geo.re[d];
    const int do_lo = bc_lo[d] == PGPU_BC_SYMMETRY, do_hi = bc_hi[d] == PGPU_BC_SYMMETRY;
    if ((do_lo || do_hi) && n > 0) {
      KTimer t("bc_symmetry");
      k_bc_symmetry<<<nb(n), 256, 0, st>>>(s->x[d], s->xold[d], s->v[d], s->vold[d], n, left, right, do_lo, do_hi);
    }
  }
  s->binned = false;
  // the leavers of outflow boundaries move to the outflow lists (:187-224)
  int rc = s->n > 0 ? transfer_outflow(s, bc_lo, bc_hi) : 0;
  if (rc) return rc;
  // the inflow lists of inflow_outflow boundaries join the species (inflow_Lo / inflow_Hi, :199-224)
  rc = inject_inflow(s, bc_lo, bc_hi);
  if (rc) return rc;
  if (s->n == 0) return 0;
  for (int d = 0; d < g->desc.D; ++d) {
    if (bc_lo[d] == PGPU_BC_PERIODIC || bc_hi[d] == PGPU_BC_PERIODIC) {
      KTimer t("bc_periodic");
      k_bc_periodic<<<nb(s->n), 256, 0, st>>>(s->x[d], s->xold[d], s->n, g->geo.le[d], g->geo.re[d]);
    }
  }
  return 0;
}

int pgpu_finish_implicit_step(pgpu_species_t s, const int *bc_lo, const int *bc_hi) {
  NEED_INIT();
  if (!s || !bc_lo || !bc_hi) return PGPU_ERR_ARG;
  const pgpu_grid_s *g = s->grid;
  bool fusable = true;
  for (int d = 0; d < g->desc.D; ++d) {
    if (bc_lo[d] == PGPU_BC_SYMMETRY || bc_hi[d] == PGPU_BC_SYMMETRY) fusable = false;
    if ((bc_lo[d] == PGPU_BC_PERIODIC) != (bc_hi[d] == PGPU_BC_PERIODIC)) fusable = false;
    // outflow / inflow_outflow boundaries move particles between the species and its side lists: the separate calls
    if (bc_lo[d] == PGPU_BC_OUTFLOW || bc_hi[d] == PGPU_BC_OUTFLOW || bc_lo[d] == PGPU_BC_INFLOW_OUTFLOW ||
        bc_hi[d] == PGPU_BC_INFLOW_OUTFLOW)
      fusable = false;
  }
  if (!fusable) {
    int rc = pgpu_advance_velocities_2nd_half(s);
    if (!rc) rc = pgpu_advance_positions_2nd_half(s);
    if (!rc) rc = pgpu_apply_bcs(s, bc_lo, bc_hi);
    return rc;
  }
  if (s->n == 0) return 0;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  FinishArgs a;
  for (int d = 0; d < 2; ++d) {
    a.x[d] = s->x[d];
    a.xold[d] = s->xold[d];
    a.periodic[d] = d < g->desc.D && bc_lo[d] == PGPU_BC_PERIODIC;
    a.left[d] = g->geo.le[d];
    a.right[d] = g->geo.re[d];
  }
  for (int c = 0; c < 3; ++c) {
    a.v[c] = s->v[c];
    a.vold[c] = s->vold[c];
  }
  a.n = s->n;
  a.D = g->desc.D;
  a.motion = s->desc.motion;
  a.forces = s->desc.forces;
  KTimer t("finish_step");
  k_finish_step<<<nb(s->n), 256, 0, ctx().stream>>>(a);
  s->binned = false;
  return 0;
}

int pgpu_stable_dt(pgpu_species_t s, double *dt_out) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  Context &c = ctx();
  unsigned long long *d_bits = &c.d_counters->maxbits;  // zero between calls
  Counters k;
  if (s->n > 0) {
    KTimer t("stable_dt");
    k_max_dtinv<<<c.sm_count * 4, 256, 0, c.stream>>>(s->v[0], s->v[1], s->v[2], s->n, s->grid->desc.D,
                                                       s->grid->geo.dx[0], s->grid->geo.dx[1], s->desc.relativistic,
                                                       d_bits);
  }
  PGPU_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c.stream));
  PGPU_CUDA(cudaMemsetAsync(&c.d_counters->maxbits, 0, sizeof(unsigned long long), c.stream));
  PGPU_CUDA(cudaStreamSynchronize(c.stream));
  k = *c.h_counters;
  double maxDtinv;
  memcpy(&maxDtinv, &k.maxbits, sizeof(double));
  // local_stable_dt = 1.0/maxDtinv/m_cvac_norm (PicChargedSpecies.cpp:1904)
  *dt_out = 1.0 / maxDtinv / s->desc.cvac_norm;
  return 0;
}

int pgpu_global_moments(pgpu_species_t s, double *out) {
  NEED_INIT();
  if (!s) return PGPU_ERR_ARG;
  Context &c = ctx();
  double *d_out = nullptr;
  PGPU_CUDA(cudaMalloc(&d_out, 7 * sizeof(double)));
  PGPU_CUDA(cudaMemsetAsync(d_out, 0, 7 * sizeof(double), c.stream));
  if (s->n > 0) {
    KTimer t("global_moments");
    k_global_moments<<<c.sm_count * 4, 256, 0, c.stream>>>(s->w, s->v[0], s->v[1], s->v[2], s->n, d_out,
                                                            s->desc.relativistic);
  }
  PGPU_CUDA(cudaMemcpyAsync(out, d_out, 7 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  PGPU_CUDA(cudaStreamSynchronize(c.stream));
  cudaFree(d_out);
  return 0;
}

}  // extern "C"
