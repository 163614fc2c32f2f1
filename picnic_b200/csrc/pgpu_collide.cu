// pgpu_collide.cu -- intra-cell Monte-Carlo binary Coulomb collisions.
//
// TakizukaAbe (src/scattering/TakizukaAbe.cpp:240-578): per cell, shuffle the
// particle list, pair neighbours, rotate the relative velocity by a random angle.
// The reference draws from one global std::mt19937 (std::shuffle + MathUtils::rand/
// randn); that stream cannot be reproduced in parallel.  Here every random number is
// a pure function of (seed, step, particle id) or (seed, step, global cell, pair) via
// Philox4x32-10, so results do not depend on the box decomposition, the number of
// GPUs or the storage order of the particles.
//
// One warp owns one cell of the cell-sorted arrays: lanes build the random order
// (rank of a per-particle Philox key), then each lane scatters pairs.
#include <vector>

#include "pgpu_internal.h"

namespace pgpu {

static inline unsigned nb(long n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// ---- Philox4x32-10 (Salmon et al., SC'11) -------------------------------------------
struct u4 {
  unsigned x, y, z, w;
};
__device__ __forceinline__ u4 philox4x32_10(u4 ctr, unsigned k0, unsigned k1) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    u4 n;
    n.x = hi1 ^ ctr.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ ctr.w ^ k1;
    n.w = lo0;
    ctr = n;
    k0 += W0;
    k1 += W1;
  }
  return ctr;
}
// uniform in (0,1): never 0 or 1
__device__ __forceinline__ double u01(unsigned a) { return ((double)a + 0.5) * 2.3283064365386963e-10; }

enum { STREAM_SHUFFLE = 0x5348u, STREAM_PAIR = 0x5041u };

// ScatteringUtils::computeDeltaU (ScatteringUtils.H:78-105)
__device__ __forceinline__ void scatter_delta_u(double ux, double uy, double uz, double costh, double sinth,
                                                double cosphi, double sinphi, double *dU) {
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  const double uperp = sqrt(ux * ux + uy * uy);
  if (uperp == 0.0) {
    dU[0] = u * sinth * cosphi;
    dU[1] = u * sinth * sinphi;
    dU[2] = u * costh - u;
  } else {
    // one reciprocal instead of the reference's four divisions by uperp (differs from it by rounding only)
    const double iu = 1.0 / uperp, sc = sinth * cosphi, ss = sinth * sinphi, omc = 1. - costh;
    dU[0] = ux * uz * iu * sc - uy * u * iu * ss - ux * omc;
    dU[1] = uy * uz * iu * sc + ux * u * iu * ss - uy * omc;
    dU[2] = -uperp * sc - uz * omc;
  }
}

// TakizukaAbe::computeDeltaU (TakizukaAbe.cpp:538-578) with the draws made explicit
__device__ __forceinline__ void ta_delta_u(const double *vp1, double den1, const double *vp2, double den2,
                                           double b90_fact, double Clog, double dt_sec, double gauss,
                                           double u_theta, double u_phi, double *dU) {
  const double PI = 3.14159265358979323846, TWOPI = 2.0 * PI, CVAC = 2.99792458e+08;
  const double ux = vp1[0] - vp2[0], uy = vp1[1] - vp2[1], uz = vp1[2] - vp2[2];
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  const double den = fmin(den1, den2);
  const double b90 = b90_fact / (u * u);
  const double deltasq_var = TWOPI * b90 * b90 * den * Clog * u * CVAC * dt_sec;
  double sinth, costh;
  if (deltasq_var < 1.0) {
    const double delta = sqrt(deltasq_var) * gauss;
    const double deltasq = delta * delta;
    const double inv = 1.0 / (1.0 + deltasq);
    sinth = 2.0 * delta * inv;
    costh = 1.0 - 2.0 * deltasq * inv;
  } else {
    sincospi(u_theta, &sinth, &costh);          // theta = pi u_theta
  }
  double sinphi, cosphi;
  sincospi(2.0 * u_phi, &sinphi, &cosphi);      // phi = 2 pi u_phi
  scatter_delta_u(ux, uy, uz, costh, sinth, cosphi, sinphi, dU);
}

// ScatteringUtils::rotateVelocity (ScatteringUtils.H:49-75)
__device__ __forceinline__ void rotate_velocity(double *a_u, double costh, double sinth, double cosphi, double sinphi) {
  const double ux = a_u[0], uy = a_u[1], uz = a_u[2];
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  const double uperp = sqrt(ux * ux + uy * uy);
  if (uperp == 0.0) {
    a_u[0] = u * sinth * cosphi;
    a_u[1] = u * sinth * sinphi;
    a_u[2] = u * costh;
  } else {
    const double iu = 1.0 / uperp, sc = sinth * cosphi, ss = sinth * sinphi;
    a_u[0] = ux * uz * iu * sc - uy * u * iu * ss + ux * costh;
    a_u[1] = uy * uz * iu * sc + ux * u * iu * ss + uy * costh;
    a_u[2] = -uperp * sc + uz * costh;
  }
}

// TakizukaAbe::LorentzScatter (TakizukaAbe.cpp:580-659): relativistic binary collision of two equal-weight
// particles through the centre-of-momentum frame, draws made explicit.  The reference keeps the scalars in
// long double; here they are doubles (differences at round-off).
__device__ __forceinline__ void ta_lorentz_scatter(double *up1, double *up2, double m1, double m2, double den2,
                                                   double dt_sec, double b90_fact, double Clog, double gauss,
                                                   double u_theta, double u_phi) {
  const double PI = 3.14159265358979323846, TWOPI = 2.0 * PI, CVAC = 2.99792458e+08;
  const double gamma1 = sqrt(1.0 + up1[0] * up1[0] + up1[1] * up1[1] + up1[2] * up1[2]);
  const double gamma2 = sqrt(1.0 + up2[0] * up2[0] + up2[1] * up2[1] + up2[2] * up2[2]);
  const double Etot = gamma1 * m1 + gamma2 * m2;
  double vcm[3], upst[3];
#pragma unroll
  for (int n = 0; n < 3; ++n) vcm[n] = (m1 * up1[n] + m2 * up2[n]) / Etot;
  const double gammacm = 1.0 / sqrt(1.0 - vcm[0] * vcm[0] - vcm[1] * vcm[1] - vcm[2] * vcm[2]);
  double vcmdotup = vcm[0] * up2[0] + vcm[1] * up2[1] + vcm[2] * up2[2];
  const double gamma2st = gammacm * (gamma2 - vcmdotup);
  vcmdotup = vcm[0] * up1[0] + vcm[1] * up1[1] + vcm[2] * up1[2];
  const double gamma1st = gammacm * (gamma1 - vcmdotup);
  const double gfac = gammacm / (1.0 + gammacm);
  double upst_fact = (gfac * vcmdotup - gamma1) * gammacm;
#pragma unroll
  for (int n = 0; n < 3; ++n) upst[n] = up1[n] + upst_fact * vcm[n];
  const double muRst = gamma1st * m1 * gamma2st * m2 / (m2 * gamma2st + m1 * gamma1st);
  const double upstsq = upst[0] * upst[0] + upst[1] * upst[1] + upst[2] * upst[2];
  const double denom = 1.0 + upstsq * m1 / m2 / gamma1st / gamma2st;
  const double vrelst = sqrt(upstsq) * m1 / muRst / denom;
  double s12 = PI * b90_fact * b90_fact * den2 * Clog * vrelst * CVAC * dt_sec;
  s12 *= gamma1st * gamma2st / gamma1 / gamma2;
  const double mv2 = muRst * vrelst * vrelst;
  s12 /= mv2 * mv2;
  double sinth, costh;
  if (s12 < 2.0) {
    const double delta = sqrt(s12 / 2.0) * gauss;
    const double deltasq = delta * delta;
    const double inv = 1.0 / (1.0 + deltasq);
    sinth = 2.0 * delta * inv;
    costh = 1.0 - 2.0 * deltasq * inv;
  } else {
    sincos(PI * u_theta, &sinth, &costh);
  }
  double sinphi, cosphi;
  sincos(TWOPI * u_phi, &sinphi, &cosphi);
  rotate_velocity(upst, costh, sinth, cosphi, sinphi);
  vcmdotup = vcm[0] * upst[0] + vcm[1] * upst[1] + vcm[2] * upst[2];
  upst_fact = (gfac * vcmdotup + gamma1st) * gammacm;
#pragma unroll
  for (int n = 0; n < 3; ++n) up1[n] = upst[n] + upst_fact * vcm[n];
  const double r = -m1 / m2;     // p2* = -p1*
#pragma unroll
  for (int n = 0; n < 3; ++n) upst[n] *= r;
  vcmdotup *= r;
  upst_fact = (gfac * vcmdotup + gamma2st) * gammacm;
#pragma unroll
  for (int n = 0; n < 3; ++n) up2[n] = upst[n] + upst_fact * vcm[n];
}

__global__ void k_ta_lorentz(long n, const double *u1, const double *u2, double m1, double m2, const double *den2,
                             double dt_sec, double b90_fact, double Clog, const double *gauss, const double *uth,
                             const double *uphi, double *o1, double *o2) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[3] = {u1[i], u1[n + i], u1[2 * n + i]}, b[3] = {u2[i], u2[n + i], u2[2 * n + i]};
  ta_lorentz_scatter(a, b, m1, m2, den2[i], dt_sec, b90_fact, Clog, gauss[i], uth[i], uphi[i]);
  for (int c = 0; c < 3; ++c) {
    o1[c * n + i] = a[c];
    o2[c * n + i] = b[c];
  }
}

__global__ void k_ta_delta_u(long n, const double *vp1, const double *den1, const double *vp2,
                             const double *den2, double b90_fact, double Clog, double dt_sec,
                             const double *gauss, const double *uth, const double *uphi, double *dU) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a[3] = {vp1[i], vp1[n + i], vp1[2 * n + i]};
  const double b[3] = {vp2[i], vp2[n + i], vp2[2 * n + i]};
  double d[3];
  ta_delta_u(a, den1[i], b, den2[i], b90_fact, Clog, dt_sec, gauss[i], uth[i], uphi[i], d);
  dU[i] = d[0];
  dU[n + i] = d[1];
  dU[2 * n + i] = d[2];
}

__global__ void k_scatter_delta_u(long n, const double *u, const double *ct, const double *st, const double *cp,
                                  const double *sp, double *dU) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double d[3];
  scatter_delta_u(u[i], u[n + i], u[2 * n + i], ct[i], st[i], cp[i], sp[i], d);
  dU[i] = d[0];
  dU[n + i] = d[1];
  dU[2 * n + i] = d[2];
}

struct TAParams {
  double b90_fact, Clog, dt_sec;
  double f1, f2;  // mu/m1, mu/m2
  double m1, m2;  // masses (LorentzScatter)
  int rel;        // RELATIVISTIC_PARTICLES build: LorentzScatter instead of computeDeltaU
  unsigned seed_lo, seed_hi, step_lo, step_hi;
  int box_lo0, box_lo1, nbox0, ncell_glob0;  // to form the global cell id
};

__device__ __forceinline__ unsigned global_cell(const TAParams &P, int cell) {
  const int i = cell % P.nbox0 + P.box_lo0, j = cell / P.nbox0 + P.box_lo1;
  return (unsigned)(i + j * P.ncell_glob0);
}

// Cells of up to 32 R particles (R = 2, 4, 8): bitonic network over R registers per lane and warp shuffles.
// The sorted word is (Philox bits << IB) | local index with IB = log2(32 R) index bits, so the sort needs no
// payload and the key is unique; two particles whose 31 - IB random bits collide (6e-5 per 64-particle cell)
// keep their storage order.
template <int R>
__device__ __forceinline__ void warp_bitonic_order(int s, int n, const uint64_t *id, const TAParams &P, unsigned salt,
                                                   int *out, int lane) {   // out[pos], pos = 0..n-1 (global or shared)
  constexpr unsigned IDX_MASK = 32u * R - 1u;
  unsigned v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int k = lane + 32 * r;
    v[r] = 0xffffffffu;
    if (k < n) {
      const uint64_t pid = id[s + k];
      u4 c;
      c.x = (unsigned)pid;
      c.y = (unsigned)(pid >> 32);
      c.z = P.step_lo;
      c.w = P.step_hi ^ (STREAM_SHUFFLE << 16) ^ salt;
      v[r] = ((philox4x32_10(c, P.seed_lo, P.seed_hi).x >> 1) & ~IDX_MASK) | (unsigned)k;
    }
  }
  // element e = lane + 32 r; stage (k2, j): e and e ^ j are ordered ascending iff (e & k2) == 0
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
          const unsigned a = v[r], b = v[r | dr];
          const unsigned lo = min(a, b), hi = max(a, b);
          v[r] = up ? lo : hi;
          v[r | dr] = up ? hi : lo;
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const unsigned o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool up = (((lane + 32 * r) & k2) == 0);
          v[r] = (lower == up) ? min(v[r], o) : max(v[r], o);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int pos = lane + 32 * r;
    if (pos < n) out[pos] = (int)(v[r] & IDX_MASK);
  }
  __syncwarp();
}

// random order of the n particles of a cell: order[r] = local index of the particle
// with the r-th smallest (key, id) where key = Philox(seed, step, id)
__device__ __forceinline__ void warp_shuffle_order(int s, int n, const uint64_t *id, const TAParams &P,
                                                   unsigned salt, unsigned *key, int *order, int lane) {
  if (n <= 64) return warp_bitonic_order<2>(s, n, id, P, salt, order + s, lane);
  if (n <= 128) return warp_bitonic_order<4>(s, n, id, P, salt, order + s, lane);
  if (n <= 256) return warp_bitonic_order<8>(s, n, id, P, salt, order + s, lane);
  // larger cells: rank of every key by counting, keys in global scratch
  for (int k = lane; k < n; k += 32) {
    const uint64_t pid = id[s + k];
    u4 c;
    c.x = (unsigned)pid;
    c.y = (unsigned)(pid >> 32);
    c.z = P.step_lo;
    c.w = P.step_hi ^ (STREAM_SHUFFLE << 16) ^ salt;
    key[s + k] = philox4x32_10(c, P.seed_lo, P.seed_hi).x;
  }
  __syncwarp();
  for (int k = lane; k < n; k += 32) {
    const unsigned mine = key[s + k];
    int r = 0;
    for (int j = 0; j < n; ++j) {
      const unsigned o = key[s + j];
      r += (o < mine || (o == mine && j < k)) ? 1 : 0;
    }
    order[s + r] = k;
  }
  __syncwarp();
}

__device__ __forceinline__ void pair_randoms(const TAParams &P, unsigned gcell, unsigned pair, unsigned salt,
                                             double &gauss, double &uth, double &uphi) {
  u4 c;
  c.x = pair;
  c.y = gcell;
  c.z = P.step_lo;
  c.w = P.step_hi ^ (STREAM_PAIR << 16) ^ salt;
  const u4 r = philox4x32_10(c, P.seed_lo, P.seed_hi);
  // Box-Muller in single precision: the two uniforms carry 32 random bits each, so an fp64 log / cos buys
  // nothing for the N(0,1) draw (it only enters TakizukaAbe::computeDeltaU as a number)
  const float f1 = ((float)(r.x >> 8) + 0.5f) * 5.9604644775390625e-8f;    // (0,1), 24 bits
  const float f2 = ((float)(r.y >> 8) + 0.5f) * 5.9604644775390625e-8f;
  gauss = (double)(sqrtf(-2.0f * logf(f1)) * cospif(2.0f * f2));
  uth = u01(r.z);
  uphi = u01(r.w);
}

template <int REL = -1>   // -1: P.rel decides at run time; 0 / 1: compiled for one build
__device__ __forceinline__ void scatter_pair(double *v0, double *v1, double *v2, int pa, double *w0,
                                             double *w1, double *w2, int pb, double dena, double denb,
                                             const TAParams &P, double gauss, double uth, double uphi,
                                             bool inter = false) {
  double a[3] = {v0[pa], v1[pa], v2[pa]};
  double b[3] = {w0[pb], w1[pb], w2[pb]};
  if (REL < 0 ? P.rel != 0 : REL != 0) {
    // TakizukaAbe.cpp:336-337 (self) and :503-505 (between species: the one with the lower density goes second)
    if (inter && dena <= denb) ta_lorentz_scatter(b, a, P.m2, P.m1, dena, P.dt_sec, P.b90_fact, P.Clog, gauss, uth, uphi);
    else ta_lorentz_scatter(a, b, P.m1, P.m2, denb, P.dt_sec, P.b90_fact, P.Clog, gauss, uth, uphi);
    v0[pa] = a[0], v1[pa] = a[1], v2[pa] = a[2];
    w0[pb] = b[0], w1[pb] = b[1], w2[pb] = b[2];
    return;
  }
  double dU[3];
  ta_delta_u(a, dena, b, denb, P.b90_fact, P.Clog, P.dt_sec, gauss, uth, uphi, dU);
  v0[pa] = a[0] + P.f1 * dU[0];
  v1[pa] = a[1] + P.f1 * dU[1];
  v2[pa] = a[2] + P.f1 * dU[2];
  w0[pb] = b[0] - P.f2 * dU[0];
  w1[pb] = b[1] - P.f2 * dU[1];
  w2[pb] = b[2] - P.f2 * dU[2];
}

// TakizukaAbe::applySelfScattering (TakizukaAbe.cpp:263-402), one cell by one warp
__device__ __forceinline__ void ta_self_cell(int cell, int lane, const int *cell_start, double *v0, double *v1, double *v2,
                                             const uint64_t *id, const double *dens, const TAParams &P, unsigned *key,
                                             int *order, unsigned long long *npairs) {
  const int s = cell_start[cell], n = cell_start[cell + 1] - s;
  if (n < 2) return;
  const double numDen = dens[cell];
  if (numDen == 0.0) return;
  warp_shuffle_order(s, n, id, P, 0u, key, order, lane);
  const unsigned gcell = global_cell(P, cell);
  const int pstart = (n % 2 == 0) ? 0 : 3;
  const int nmain = (n - pstart) / 2;
  for (int q = lane; q < nmain; q += 32) {
    const int pa = s + order[s + pstart + 2 * q], pb = s + order[s + pstart + 2 * q + 1];
    double g, ut, up;
    pair_randoms(P, gcell, (unsigned)(pstart + 2 * q), 0u, g, ut, up);
    scatter_pair(v0, v1, v2, pa, v0, v1, v2, pb, numDen, numDen, P, g, ut, up);
  }
  if (pstart == 3 && lane == 0) {
    // particles 0,1,2 scatter as (0,1), (1,2), (0,2) with half the density (:353-388)
    const int t0 = s + order[s], t1 = s + order[s + 1], t2 = s + order[s + 2];
    for (int p = 0; p < 3; ++p) {
      double g, ut, up;
      pair_randoms(P, gcell, (unsigned)p, 1u, g, ut, up);
      // the relativistic build passes the full density to the three pairs of an odd cell (:372 vs :377)
      const double dh = P.rel ? numDen : numDen / 2.0;
      // (no small arrays here: indexed by the loop counter they would live in local memory)
      scatter_pair(v0, v1, v2, p == 1 ? t1 : t0, v0, v1, v2, p == 0 ? t1 : t2, dh, dh, P, g, ut, up);
    }
  }
  if (lane == 0) atomicAdd(npairs, (unsigned long long)(nmain + (pstart == 3 ? 3 : 0)));
}

// every cell (list == nullptr), or the cells the staged kernel left in `list` (list[0] = how many, then the cells)
__global__ void __launch_bounds__(256)
k_ta_self(const int *cell_start, int ncell, double *v0, double *v1, double *v2, const uint64_t *id,
          const double *dens, TAParams P, unsigned *key, int *order, unsigned long long *npairs, const int *list) {
  const int gw = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), nw = (int)(((long)gridDim.x * blockDim.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int count = list ? list[0] : ncell;
  for (int i = gw; i < count; i += nw)
    ta_self_cell(list ? list[1 + i] : i, lane, cell_start, v0, v1, v2, id, dens, P, key, order, npairs);
}

// TakizukaAbe::applyInterScattering (TakizukaAbe.cpp:404-536).  Pair p couples
// particle p%n1 with p (or p with p%n2): the particle of the shorter list collides
// repeatedly, in order of p, so one lane owns it and walks its partners.
__device__ __forceinline__ void ta_inter_cell(int cell, int lane, const int *cs1, const int *cs2, double *a0, double *a1,
                                              double *a2, const uint64_t *id1, const double *dens1, double *b0, double *b1,
                                              double *b2, const uint64_t *id2, const double *dens2, const TAParams &P,
                                              unsigned *key1, int *order1, unsigned *key2, int *order2,
                                              unsigned long long *npairs) {
  const int s1 = cs1[cell], n1 = cs1[cell + 1] - s1;
  const int s2 = cs2[cell], n2 = cs2[cell + 1] - s2;
  if ((long)n1 * n2 < 2) return;
  const double numDen1 = dens1[cell], numDen2 = dens2[cell];
  if (numDen1 * numDen2 == 0.0) return;
  warp_shuffle_order(s1, n1, id1, P, 2u, key1, order1, lane);
  warp_shuffle_order(s2, n2, id2, P, 3u, key2, order2, lane);
  const unsigned gcell = global_cell(P, cell);
  const int pMin = min(n1, n2), pMax = max(n1, n2);
  const bool first_short = (pMin == n1);
  for (int r = lane; r < pMin; r += 32) {
    for (int p = r; p < pMax; p += pMin) {
      const int i1 = s1 + order1[s1 + (first_short ? r : p)];
      const int i2 = s2 + order2[s2 + (first_short ? p : r)];
      double g, ut, up;
      pair_randoms(P, gcell, (unsigned)p, 2u, g, ut, up);
      scatter_pair(a0, a1, a2, i1, b0, b1, b2, i2, numDen1, numDen2, P, g, ut, up, true);
    }
  }
  if (lane == 0) atomicAdd(npairs, (unsigned long long)pMax);
}

__global__ void __launch_bounds__(256)
k_ta_inter(const int *cs1, const int *cs2, int ncell, double *a0, double *a1, double *a2, const uint64_t *id1,
           const double *dens1, double *b0, double *b1, double *b2, const uint64_t *id2, const double *dens2,
           TAParams P, unsigned *key1, int *order1, unsigned *key2, int *order2, unsigned long long *npairs,
           const int *list) {
  const int gw = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), nw = (int)(((long)gridDim.x * blockDim.x) >> 5);
  const int lane = threadIdx.x & 31;
  const int count = list ? list[0] : ncell;
  for (int i = gw; i < count; i += nw)
    ta_inter_cell(list ? list[1 + i] : i, lane, cs1, cs2, a0, a1, a2, id1, dens1, b0, b1, b2, id2, dens2, P, key1, order1,
                  key2, order2, npairs);
}


// ---- the staged Takizuka-Abe kernels ------------------------------------------------------------------------------
// Cells of up to TA_NMAX particles per species (the usual case: decks run 16 - 200 per cell): the warp copies the
// cell's velocities into shared memory with cp.async while it draws and sorts the shuffle keys, pairs the particles
// there, and writes the velocities back in storage order.  Same keys, same pair draws, same arithmetic as
// k_ta_self / k_ta_inter, which keep the larger cells: what changes is that the three dependent global round trips of
// a cell (ids -> order -> velocities) become one.
constexpr int TA_NMAX = 128;
constexpr int TA_SPECIES_BYTES = TA_NMAX * (3 * 8 + 4);     // v[3][TA_NMAX] doubles + order[TA_NMAX]
constexpr int TA_SELF_WARPS = 8, TA_INTER_WARPS = 4;
// blocks per SM the staged kernels are compiled for: 4 x 8 warps at 64 registers (self), 7 x 4 warps at 72 (inter; 201 KB of
// shared memory) -- measured against 3 / 6: kernels of a C2 step 0.396 -> 0.366 ms
#ifndef PGPU_TA_SELF_MINB
#define PGPU_TA_SELF_MINB 4
#endif
#ifndef PGPU_TA_INTER_MINB
#define PGPU_TA_INTER_MINB 7
#endif

__device__ __forceinline__ void cp_async8(void *smem, const void *g) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ta_stage_in(double *sv, const double *v0, const double *v1, const double *v2, int s, int n,
                                            int lane) {
  for (int k = lane; k < n; k += 32) {
    cp_async8(sv + k, v0 + s + k);
    cp_async8(sv + TA_NMAX + k, v1 + s + k);
    cp_async8(sv + 2 * TA_NMAX + k, v2 + s + k);
  }
}
__device__ __forceinline__ void ta_stage_out(const double *sv, double *v0, double *v1, double *v2, int s, int n, int lane) {
  for (int k = lane; k < n; k += 32) {
    v0[s + k] = sv[k];
    v1[s + k] = sv[TA_NMAX + k];
    v2[s + k] = sv[2 * TA_NMAX + k];
  }
}
__device__ __forceinline__ void ta_staged_order(int s, int n, const uint64_t *id, const TAParams &P, unsigned salt, int *so,
                                                int lane) {
  if (n <= 64) warp_bitonic_order<2>(s, n, id, P, salt, so, lane);
  else warp_bitonic_order<4>(s, n, id, P, salt, so, lane);
}

template <int REL>
__global__ void __launch_bounds__(32 * TA_SELF_WARPS, PGPU_TA_SELF_MINB)
k_ta_self_staged(const int *cell_start, int ncell, double *v0, double *v1, double *v2, const uint64_t *id,
                 const double *dens, TAParams P, unsigned long long *npairs, int *list) {
  extern __shared__ __align__(16) unsigned char ta_smem[];
  __shared__ unsigned block_pairs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) block_pairs = 0;
  __syncthreads();
  const int cell = blockIdx.x * TA_SELF_WARPS + warp;
  int s = 0, n = 0;
  double numDen = 0.0;
  if (cell < ncell) {
    s = cell_start[cell];
    n = cell_start[cell + 1] - s;
    numDen = dens[cell];
  }
  if (numDen != 0.0 && n >= 2 && n <= TA_NMAX) {
    double *sv = reinterpret_cast<double *>(ta_smem + warp * TA_SPECIES_BYTES);
    int *so = reinterpret_cast<int *>(sv + 3 * TA_NMAX);
    ta_stage_in(sv, v0, v1, v2, s, n, lane);
    ta_staged_order(s, n, id, P, 0u, so, lane);
    cp_async_wait_all();
    __syncwarp();
    const unsigned gcell = global_cell(P, cell);
    const int pstart = (n % 2 == 0) ? 0 : 3;
    const int nmain = (n - pstart) / 2;
    double *w0 = sv, *w1 = sv + TA_NMAX, *w2 = sv + 2 * TA_NMAX;
    for (int q = lane; q < nmain; q += 32) {
      const int pa = so[pstart + 2 * q], pb = so[pstart + 2 * q + 1];
      double g, ut, up;
      pair_randoms(P, gcell, (unsigned)(pstart + 2 * q), 0u, g, ut, up);
      scatter_pair<REL>(w0, w1, w2, pa, w0, w1, w2, pb, numDen, numDen, P, g, ut, up);
    }
    if (pstart == 3 && lane == 0) {
      // particles 0,1,2 scatter as (0,1), (1,2), (0,2) with half the density (TakizukaAbe.cpp:353-388)
      const int t0 = so[0], t1 = so[1], t2 = so[2];
      for (int p = 0; p < 3; ++p) {
        double g, ut, up;
        pair_randoms(P, gcell, (unsigned)p, 1u, g, ut, up);
        const double dh = REL ? numDen : numDen / 2.0;
        scatter_pair<REL>(w0, w1, w2, p == 1 ? t1 : t0, w0, w1, w2, p == 0 ? t1 : t2, dh, dh, P, g, ut, up);
      }
    }
    __syncwarp();
    ta_stage_out(sv, v0, v1, v2, s, n, lane);
    if (lane == 0) atomicAdd(&block_pairs, (unsigned)(nmain + (pstart == 3 ? 3 : 0)));
  } else if (n > TA_NMAX && lane == 0) {
    list[1 + atomicAdd(list, 1)] = cell;       // left to k_ta_self
  }
  __syncthreads();
  if (threadIdx.x == 0 && block_pairs) atomicAdd(npairs, (unsigned long long)block_pairs);
}

template <int REL>
__global__ void __launch_bounds__(32 * TA_INTER_WARPS, PGPU_TA_INTER_MINB)
k_ta_inter_staged(const int *cs1, const int *cs2, int ncell, double *a0, double *a1, double *a2, const uint64_t *id1,
                  const double *dens1, double *b0, double *b1, double *b2, const uint64_t *id2, const double *dens2,
                  TAParams P, unsigned long long *npairs, int *list) {
  extern __shared__ __align__(16) unsigned char ta_smem[];
  __shared__ unsigned block_pairs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) block_pairs = 0;
  __syncthreads();
  const int cell = blockIdx.x * TA_INTER_WARPS + warp;
  int s1 = 0, n1 = 0, s2 = 0, n2 = 0;
  double numDen1 = 0.0, numDen2 = 0.0;
  if (cell < ncell) {
    s1 = cs1[cell], n1 = cs1[cell + 1] - s1;
    s2 = cs2[cell], n2 = cs2[cell + 1] - s2;
    numDen1 = dens1[cell], numDen2 = dens2[cell];
  }
  if (numDen1 * numDen2 != 0.0 && (long)n1 * n2 >= 2 && max(n1, n2) <= TA_NMAX) {
    double *sa = reinterpret_cast<double *>(ta_smem + warp * 2 * TA_SPECIES_BYTES);
    int *oa = reinterpret_cast<int *>(sa + 3 * TA_NMAX);
    double *sb = reinterpret_cast<double *>(ta_smem + warp * 2 * TA_SPECIES_BYTES + TA_SPECIES_BYTES);
    int *ob = reinterpret_cast<int *>(sb + 3 * TA_NMAX);
    ta_stage_in(sa, a0, a1, a2, s1, n1, lane);
    ta_stage_in(sb, b0, b1, b2, s2, n2, lane);
    ta_staged_order(s1, n1, id1, P, 2u, oa, lane);
    ta_staged_order(s2, n2, id2, P, 3u, ob, lane);
    cp_async_wait_all();
    __syncwarp();
    const unsigned gcell = global_cell(P, cell);
    const int pMin = min(n1, n2), pMax = max(n1, n2);
    const bool first_short = (pMin == n1);
    for (int r = lane; r < pMin; r += 32) {
      for (int p = r; p < pMax; p += pMin) {
        const int i1 = oa[first_short ? r : p], i2 = ob[first_short ? p : r];
        double g, ut, up;
        pair_randoms(P, gcell, (unsigned)p, 2u, g, ut, up);
        scatter_pair<REL>(sa, sa + TA_NMAX, sa + 2 * TA_NMAX, i1, sb, sb + TA_NMAX, sb + 2 * TA_NMAX, i2, numDen1, numDen2, P, g,
                     ut, up, true);
      }
    }
    __syncwarp();
    ta_stage_out(sa, a0, a1, a2, s1, n1, lane);
    ta_stage_out(sb, b0, b1, b2, s2, n2, lane);
    if (lane == 0) atomicAdd(&block_pairs, (unsigned)pMax);
  } else if (max(n1, n2) > TA_NMAX && lane == 0) {
    list[1 + atomicAdd(list, 1)] = cell;       // left to k_ta_inter
  }
  __syncthreads();
  if (threadIdx.x == 0 && block_pairs) atomicAdd(npairs, (unsigned long long)block_pairs);
}


// =============================================================================================
// Coulomb, PROBABILISTIC weight method (src/scattering/Coulomb.cpp:400-592, 919-1180,
// 1642-1692, 1795-1903; Coulomb.H:339-363), non-relativistic (Galilean) build.
// One warp per cell.  Pairs that share a particle are ordered by construction:
//  * O(N) pairing: disjoint neighbour pairs of the shuffled list, one lane each;
//  * NxN pairing (small cells, or Coulomb.NxN = true): a round-robin tournament -- every round
//    is a set of disjoint pairs done by the lanes in parallel, rounds are separated by a warp
//    barrier.  The reference visits the same pairs in lexicographic order; the order of
//    independent random kicks is statistically irrelevant.
//  * inter-species: a lane owns one particle of the shorter list and walks its partners.
// =============================================================================================
enum { STREAM_WEIGHT = 0x5754u, STREAM_ELA = 0x454cu };
enum { ANG_TAKIZUKA = 0, ANG_NANBU = 1, ANG_BOBYLEV = 2, ANG_NANBU_FAS = 3, ANG_NANBU_FAS_V2 = 4, ANG_ISOTROPIC = 5 };

struct CoulParams {
  double b90_fact, bqm_fact, EF_fact, mu, f1, f2, Clog, dt_sec, cellV_SI;
  int angular, NxN, NxN_Nthresh;
  unsigned seed_lo, seed_hi, step_lo, step_hi;
  int box_lo0, box_lo1, nbox0, ncell_glob0;
  int rel;               // RELATIVISTIC_PARTICLES build: pairs go through LorentzScatter
  int sk08;              // weight_method = CONSERVATIVE (Sentoku-Kemp 2008 update of the heavier-weight particle)
  double mass1, mass2;
  int large_angle;       // include_large_angle_scattering (Coulomb::SetPolarScattering, first half)
  double large_draw;     // its uniform RL in the explicit-draw test entry points
  double fas_draw2, fas_draw3;   // second / third uniform of NANBU_FAS(_v2) in the explicit-draw test entry points
};
__device__ __forceinline__ unsigned global_cell(const CoulParams &P, int cell) {
  const int i = cell % P.nbox0 + P.box_lo0, j = cell / P.nbox0 + P.box_lo1;
  return (unsigned)(i + j * P.ncell_glob0);
}

// Coulomb::setNANBUcosthsinth (Coulomb.H:339-363)
__device__ __forceinline__ void nanbu_costh_sinth(double s12, double U, double &costh, double &sinth) {
  double A12, c;
  if (s12 < 0.1466) {
    A12 = 1.0 / (s12 * (1.0 - s12 / 2.0 + s12 * s12 / 6.0));
    c = 1.0 + 1.0 / A12 * log(1.0 - U * (1.0 - exp(-2.0 * A12)));
  } else if (s12 < 3.0) {
    const double s12sq = s12 * s12, s12cu = s12 * s12sq;
    A12 = 1.0 / (0.0056958 + 0.9560202 * s12 - 0.508139 * s12sq + 0.47913906 * s12cu - 0.12788975 * s12sq * s12sq +
                 0.02389567 * s12cu * s12sq);
    c = 1.0 + 1.0 / A12 * log(1.0 - U * (1.0 - exp(-2.0 * A12)));
  } else if (s12 < 6.0) {
    A12 = 3.0 * exp(-s12);
    c = 1.0 + 1.0 / A12 * log(1.0 - U * (1.0 - exp(-2.0 * A12)));
  } else {
    c = 2.0 * U - 1.0;
  }
  costh = c;
  sinth = sqrt(1.0 - c * c);
}

// angular_scattering = NANBU_FAS / NANBU_FAS_v2 (Higginson, JCP 2017): Coulomb::setNANBUFAScosthsinth (Coulomb.H:365-428),
// setNANBUFAS_v2_costhsinth (:430-571), setFAScoefficients (:573-618), setFAS_v2_coefficients (:620-678),
// getTransitionX_NANBU (:680-718).  The reference draws its uniforms one after the other and only those a branch takes;
// u[0..2] are handed out in that order.  A failed v2 solve (the reference exits) falls back to Nanbu's model.
struct FasDraws {
  double u[3];
  int pos;
  __device__ __forceinline__ double next() {
    const double v = pos == 0 ? u[0] : (pos == 1 ? u[1] : u[2]);
    ++pos;
    return v;
  }
};
__device__ __noinline__ double fas_transition_x(double Clog, double s12, double alpha_g, double sA) {
  double xc = 2.0;
  const double C0 = s12 / (8.0 * Clog * alpha_g * sA);
  if (C0 > exp(-2.0)) return 1.0;
  double error = 1.0;
  int iter = 0;
  while (error > 1.0e-4) {
    const double xold = xc;
    const double y0 = xc * xc * exp(-2.0 * xc) - C0;
    const double dy0dx = 2.0 * xc * (1.0 - xc) * exp(-2.0 * xc);
    xc = xc - y0 / dy0dx;
    error = fabs(1.0 - xold / xc);
    iter += 1;
    if (iter > 20) break;
  }
  if (sA * xc > 1.0) xc = 1.0 / sA;
  return xc;
}
__device__ __noinline__ int fas_coefficients(double &alpha_g, double &sA, double &muc, double Clog, double s12,
                                             double mu_max) {
  alpha_g = 1.0;
  sA = s12 / 2.0;
  double Xc = fas_transition_x(Clog, s12, alpha_g, sA);
  muc = sA * Xc;
  int iter = 0, success = 1;
  double error = 1.0;
  while (error > 1.0e-4) {
    const double sAold = sA;
    const double f1 = 1.0 - exp(-s12) + s12 / Clog * 0.5 * log(Xc * sAold / mu_max);
    const double f2 = (1.0 - exp(-2.0 * Xc)) / (1.0 - (1.0 + Xc) * exp(-2.0 * Xc));
    sA = (4.0 * mu_max * Clog / (4.0 * mu_max * Clog + s12)) * (s12 / (4.0 * Xc * Clog) + f1 * f2);
    Xc = fas_transition_x(Clog, s12, alpha_g, sA);
    muc = sA * Xc;
    alpha_g = (1.0 - s12 / (4.0 * Clog) * (mu_max - muc) / mu_max / muc) / (1.0 - exp(-2.0 * muc / sA));
    error = fabs(1.0 - sAold / sA);
    iter += 1;
    if (iter > 20) {
      success = -1;
      break;
    }
  }
  return success;
}
__device__ __noinline__ int fas_v2_coefficients(double &alpha_g, double &sA, double Clog, double s12, double mu_max,
                                                double mu_tr) {
  const double S_L = s12 / (4.0 * Clog) * mu_max / mu_tr / (mu_max + mu_tr);
  const double mu_L = s12 / (2.0 * Clog) * log((mu_max + mu_tr) / mu_tr);
  const double mu_N97 = 1.0 - exp(-s12);
  if (mu_L > mu_N97 || S_L > 1.0) {
    alpha_g = 0.0;
    sA = mu_tr;
    return -1;
  }
  sA = s12 / 2.0;
  alpha_g = fmax(0.0, (1.0 - S_L) / (1.0 - exp(-2.0 * mu_max / sA)));
  int iter = 0, success = 1;
  double error = 1.0;
  while (error > 1.0e-4) {
    const double sAold = sA;
    const double f1 = mu_N97 - mu_L + alpha_g * mu_max * exp(-2.0 * mu_max / sAold);
    const double f2 = 1.0 - S_L;
    sA = f1 / f2;
    alpha_g = fmax(0.0, (1.0 - S_L) / (1.0 - exp(-2.0 * mu_max / sA)));
    error = fabs(1.0 - sAold / sA);
    iter += 1;
    if (iter > 20) {
      success = -1;
      break;
    }
  }
  return success;
}
__device__ __noinline__ void nanbu_fas(double s12, double Clog, double b0, double bmin_qm, double sigma_eff, FasDraws D,
                                       double &costh, double &sinth) {
  const double PI = 3.14159265358979323846;
  const double bperp_sq = b0 * b0 / 4.0, bmin_sq = bmin_qm * bmin_qm;
  const double bmax_sq = exp(2.0 * Clog) * (bperp_sq + bmin_sq) - bperp_sq;
  const double s12_min = 1.33 * 4.0 * Clog / (exp(2.0 * Clog) - 1.0);
  costh = 1.0;
  sinth = 0.0;
  if (s12 < s12_min) {
    const double N12 = s12 / sigma_eff * PI * (bmax_sq - bmin_sq);
    const double PL = 1.0 - exp(-N12);
    if (D.next() < PL) {
      const double RL = D.next();
      const double bsq = bmax_sq - RL * (bmax_sq - bmin_sq);
      costh = (bsq - bperp_sq) / (bsq + bperp_sq);
      sinth = sqrt(1.0 - costh * costh);
    }
  } else if (s12 < 0.5) {
    double alpha_g, sA, muc;
    const double costhmax = (bmin_sq - bperp_sq) / (bmin_sq + bperp_sq);
    const double mu_max = (1.0 - costhmax) / 2.0;
    const int success = fas_coefficients(alpha_g, sA, muc, Clog, s12, mu_max);
    if (success < 0 || muc != muc) {
      nanbu_costh_sinth(s12, D.next(), costh, sinth);
    } else {
      const double Uc = 1.0 - s12 / (4.0 * Clog) * (mu_max - muc) / (mu_max * muc);
      const double R = D.next();
      if (R < Uc) {
        costh = 1.0 + sA * log(1.0 - R / Uc * (1.0 - exp(-2.0 * muc / sA)));
        sinth = sqrt(1.0 - costh * costh);
      } else {
        const double costhc = 1.0 - 2.0 * muc;
        const double R2 = D.next();
        const double numer = R2 * (costhc - costhmax) - costhc * (1.0 - costhmax);
        const double denom = R2 * (costhc - costhmax) - (1.0 - costhmax);
        costh = numer / denom;
        sinth = sqrt(1.0 - costh * costh);
      }
    }
  } else {
    nanbu_costh_sinth(s12, D.next(), costh, sinth);
  }
}
__device__ __noinline__ void nanbu_fas_v2(double s12, double Clog, double b0, double bmin_qm, FasDraws D, double &costh,
                                          double &sinth) {
  const double bperp_sq = b0 * b0 / 4.0, bmin_sq = bmin_qm * bmin_qm;
  const double bmax_sq = exp(2.0 * Clog) * (bperp_sq + bmin_sq) - bperp_sq;
  const double costhmax = (bmin_sq - bperp_sq) / (bmin_sq + bperp_sq);
  const double mu_max = (1.0 - costhmax) / 2.0;
  if (s12 > 0.6) {
    nanbu_costh_sinth(s12, D.next(), costh, sinth);
    return;
  }
  double Nmax = 2.0 * Clog;
  const double mutr_factor = Clog;
  const double sig_ratio = bperp_sq / (bmax_sq - bmin_sq);
  const double s12_Nmin = 4.0 * Clog * sig_ratio;
  double s12_Nmax = Nmax * s12_Nmin;
  const double Nmax_min = 2.0 * mutr_factor * mu_max / (exp(2.0 * Clog) - 1.0) / s12_Nmin;
  if (Nmax < Nmax_min) {
    Nmax = Nmax_min;
    s12_Nmax = Nmax * s12_Nmin;
  }
  double alpha_g, sA, S_N97, mu_tr;
  if (s12 < s12_Nmax) {
    const double Ntot = s12 / s12_Nmin;
    const double Pscatter = 1.0 - exp(-Ntot);
    if (D.next() > Pscatter) {
      costh = 1.0;
      sinth = 0.0;
      return;
    }
    mu_tr = s12_Nmax / mutr_factor;
    const int success = fas_v2_coefficients(alpha_g, sA, Clog, s12_Nmax, mu_max, mu_tr);
    if (success < 0 || sA != sA) {
      nanbu_costh_sinth(s12, D.next(), costh, sinth);
      return;
    }
    alpha_g = fmax(0.0, alpha_g * (s12 - s12_Nmin) / (s12_Nmax - s12_Nmin));
    if (alpha_g == 0.0) sA = mu_max;
    else sA *= s12 / s12_Nmax;
    double mu0;
    const double coefc = 4.0 * sig_ratio / mu_max;
    if (coefc < 1.0e-10) mu0 = sig_ratio;
    else mu0 = mu_max * (-1.0 + sqrt(1.0 + coefc)) / 2.0;
    S_N97 = alpha_g * (1.0 - exp(-2.0 * mu_max / sA));
    if (s12 <= s12_Nmin) {
      mu_tr = mu0;
    } else {
      const double C0 = (1.0 - S_N97) * 4.0 * Clog / s12;
      const double coef = 4.0 / mu_max / C0;
      if (coef < 1.0e-10) mu_tr = 1.0 / C0;
      else mu_tr = mu_max * (-1.0 + sqrt(1.0 + coef)) / 2.0;
      mu_tr = fmax(mu_tr, mu0);
    }
  } else {
    mu_tr = s12 / mutr_factor;
    const int success = fas_v2_coefficients(alpha_g, sA, Clog, s12, mu_max, mu_tr);
    if (success < 0 || sA != sA) {
      nanbu_costh_sinth(s12, D.next(), costh, sinth);
      return;
    }
    const double S_L = s12 / (4.0 * Clog) * mu_max / mu_tr / (mu_max + mu_tr);
    S_N97 = 1.0 - S_L;
  }
  const double R = D.next();
  if (R < S_N97) {
    costh = 1.0 + sA * log(1.0 - R / S_N97 * (1.0 - exp(-2.0 * mu_max / sA)));
  } else {
    const double RL = D.next();
    costh = 1.0 - 2.0 * RL * mu_tr * mu_max / (mu_max * (1.0 - RL) + mu_tr);
  }
  sinth = sqrt(1.0 - costh * costh);
}

// Coulomb::SetPolarScattering (:1795-1903), small-angle part, with explicit draws
__device__ __forceinline__ void coulomb_polar(int angular, double s12, double gauss, double upol, double &costh,
                                              double &sinth, double Clog = 0.0, double b0 = 0.0, double bmin_qm = 0.0,
                                              double sigma_eff = 0.0, double ufas2 = 0.5, double ufas3 = 0.5,
                                              const bool fas = true) {   // fas = false: compiled without the full-angle models
  const double PI = 3.14159265358979323846;
  costh = 1.0;
  sinth = 0.0;
  switch (angular) {
    case ANG_TAKIZUKA:
      if (s12 < 2.0) {
        const double delta = sqrt(s12 / 2.0) * fabs(gauss);
        const double deltasq = delta * delta;
        sinth = 2.0 * delta / (1.0 + deltasq);
        costh = 1.0 - 2.0 * deltasq / (1.0 + deltasq);
      } else {
        sincospi(upol, &sinth, &costh);        // theta = pi upol
      }
      break;
    case ANG_NANBU:
      nanbu_costh_sinth(s12, upol, costh, sinth);
      break;
    case ANG_BOBYLEV:
      costh = 1.0 - fmin(s12, 2.0);
      sinth = sin(acos(costh));
      break;
    case ANG_NANBU_FAS:
      if (fas) {
        FasDraws D = {{upol, ufas2, ufas3}, 0};
        nanbu_fas(s12, Clog, b0, bmin_qm, sigma_eff, D, costh, sinth);
      }
      break;
    case ANG_NANBU_FAS_V2:
      if (fas) {
        FasDraws D = {{upol, ufas2, ufas3}, 0};
        nanbu_fas_v2(s12, Clog, b0, bmin_qm, D, costh, sinth);
      }
      break;
    default:
      sincospi(upol, &sinth, &costh);        // theta = pi upol
  }
}

// scattering.coulomb.include_large_angle_scattering, Coulomb::SetPolarScattering (:1801-1863): with probability SL the pair
// makes ONE Rutherford event with an impact parameter below the cutoff b_c; the variance of the cumulative small-angle part
// shrinks so that the total stays s12.  RL = the reference's uniform draw.  true = no small-angle part follows.
__device__ __forceinline__ bool coulomb_large_angle(double &s12, double Clog, double b0, double bmin_qm, double sigma_eff,
                                                    double RL, double &costh, double &sinth) {
  const double PI = 3.14159265358979323846;
  const double bperp_sq = b0 * b0 / 4.0, bmin_sq = bmin_qm * bmin_qm;
  const double bmax_sq = exp(2.0 * Clog) * (bperp_sq + bmin_sq) - bperp_sq;
  const double N12 = s12 / sigma_eff * PI * (bmax_sq - bmin_sq);
  double bc_sq = bperp_sq + bmin_sq;
  const double N12_min = 0.1;
  double N12_tr = 80.0;
  const double N12_tr0 = N12_min / 2.0 * (bmax_sq - bmin_sq) / (bc_sq - bmin_sq);
  if (N12_tr0 < N12_tr) N12_tr = N12_tr0;
  double SL;
  if (N12 <= N12_min) {
    SL = N12;
  } else if (N12 <= N12_tr) {
    const double SL_tr = N12_tr * (bc_sq - bmin_sq) / (bmax_sq - bmin_sq);
    SL = (N12 - N12_min) / (N12_tr - N12_min) * SL_tr + (N12_tr - N12) / (N12_tr - N12_min) * N12_min;
  } else {
    SL = fmin(0.1, N12 * (bc_sq - bmin_sq) / (bmax_sq - bmin_sq));
  }
  bc_sq = bmin_sq + SL / N12 * (bmax_sq - bmin_sq);
  const double ClogM = 0.5 * log((bperp_sq + bmax_sq) / (bperp_sq + bc_sq));
  s12 *= ClogM / Clog / (1.0 - SL);
  costh = 1.0;
  sinth = 0.0;
  if (RL < SL) {
    const double bsq = bc_sq - RL / SL * (bc_sq - bmin_sq);
    costh = (bsq - bperp_sq) / (bsq + bperp_sq);
    sinth = sqrt(1.0 - costh * costh);
    return true;
  }
  return N12 <= N12_min;
}

// Coulomb::GalileanScatter (:1642-1692) + SetPolarScattering (:1795-1903) with explicit draws.
// false (dU = 0) where the reference returns early.
__device__ __forceinline__ bool coulomb_delta_u(const CoulParams &P, const double *vp1, const double *vp2,
                                                double EF_norm, double den12, double bmax, double sigma_max,
                                                double gauss, double upol, double uphi, double *dU, double *s12o,
                                                double ularge = 0.5, double ufas2 = 0.5, double ufas3 = 0.5,
                                                const bool fas = true, const bool lean = false) {
  const double PI = 3.14159265358979323846, CVAC = 2.99792458e+08;
  dU[0] = dU[1] = dU[2] = 0.0;
  const double ux = vp1[0] - vp2[0], uy = vp1[1] - vp2[1], uz = vp1[2] - vp2[2];
  const double u = sqrt(ux * ux + uy * uy + uz * uz);
  if (u <= 2.2250738585072014e-308) return false;
  const double vsum = sqrt(vp1[0] * vp1[0] + vp1[1] * vp1[1] + vp1[2] * vp1[2]) +
                      sqrt(vp2[0] * vp2[0] + vp2[1] * vp2[1] + vp2[2] * vp2[2]);
  if (u <= 1.0e-14 * vsum) return false;
  double b0 = P.b90_fact / (P.mu * u * u + 2.0 * EF_norm);
  const double bmin_qm = P.bqm_fact / (P.mu * u + sqrt(2.0 * EF_norm * P.mu));
  double Clog = P.Clog;
  if (Clog == 0.0) {
    Clog = 0.5 * log((b0 * b0 / 4.0 + bmax * bmax) / (b0 * b0 / 4.0 + bmin_qm * bmin_qm));
    Clog = fmax(2.0, Clog);
  }
  b0 = P.b90_fact / (P.mu * u * u);
  double sigma_eff = PI * b0 * b0 * Clog;
  sigma_eff = fmin(sigma_eff, sigma_max);
  double s12 = sigma_eff * den12 * u * CVAC * P.dt_sec;
  double costh, sinth;
  bool skip_small = false;
  if (!lean && P.large_angle) skip_small = coulomb_large_angle(s12, Clog, b0, bmin_qm, sigma_eff, ularge, costh, sinth);
  if (s12o) *s12o = skip_small ? -1.0 : s12;
  if (!skip_small) coulomb_polar(P.angular, s12, gauss, upol, costh, sinth, Clog, b0, bmin_qm, sigma_eff, ufas2, ufas3, fas);
  double sinphi, cosphi;
  sincospi(2.0 * uphi, &sinphi, &cosphi);       // phi = 2 pi uphi
  scatter_delta_u(ux, uy, uz, costh, sinth, cosphi, sinphi, dU);
  return true;
}

// Coulomb::LorentzScatter (:1694-1793): collision through the centre-of-momentum frame.  Particle 1 (mass m1)
// always scatters, particle 2 only if scatter2 (momentum conservation then gives its proper velocity).  The reference
// keeps the scalars in long double; fp64 here (parity with the oracle < 1e-12, tests/test_gpu_coulomb_elastic.py).
__device__ __forceinline__ bool coulomb_lorentz_scatter(const CoulParams &P, double *up1, double *up2, bool scatter2,
                                                        double m1, double m2, double EF_norm, double den12,
                                                        double bmax, double sigma_max, double gauss, double upol,
                                                        double uphi, double *s12o, double ularge = 0.5,
                                                        double ufas2 = 0.5, double ufas3 = 0.5, const bool fas = true) {
  const double PI = 3.14159265358979323846, CVAC = 2.99792458e+08;
  if (s12o) *s12o = 0.0;
  const double gb1sq = up1[0] * up1[0] + up1[1] * up1[1] + up1[2] * up1[2];
  const double gb2sq = up2[0] * up2[0] + up2[1] * up2[1] + up2[2] * up2[2];
  const double g1 = sqrt(1.0 + gb1sq), g2 = sqrt(1.0 + gb2sq);
  const double Etot = g1 * m1 + g2 * m2;
  double ptot[3], vcm[3], upst[3];
#pragma unroll
  for (int n = 0; n < 3; ++n) {
    ptot[n] = m1 * up1[n] + m2 * up2[n];
    vcm[n] = ptot[n] / Etot;
  }
  const double vcmsq = vcm[0] * vcm[0] + vcm[1] * vcm[1] + vcm[2] * vcm[2];
  const double gcm = 1.0 / sqrt(1.0 - vcmsq);
  double ucmdotup = gcm * (vcm[0] * up2[0] + vcm[1] * up2[1] + vcm[2] * up2[2]);
  const double g2st = gcm * g2 - ucmdotup;
  ucmdotup = gcm * (vcm[0] * up1[0] + vcm[1] * up1[1] + vcm[2] * up1[2]);
  const double g1st = gcm * g1 - ucmdotup;
  double upst_fact = gcm * (ucmdotup / (1.0 + gcm) - g1);
#pragma unroll
  for (int n = 0; n < 3; ++n) upst[n] = up1[n] + upst_fact * vcm[n];
  const double muRst = g1st * m1 * g2st * m2 / (g1st * m1 + g2st * m2);
  const double upstsq = upst[0] * upst[0] + upst[1] * upst[1] + upst[2] * upst[2];
  const double vrelst = sqrt(upstsq) * m1 / muRst;
  if (vrelst <= 2.2250738585072014e-308) return false;
  const double vsum = sqrt(gb1sq) / g1 + sqrt(gb2sq) / g2;
  if (vrelst <= 1.0e-14 * vsum) return false;
  const double denom = 1.0 + upstsq * m1 / m2 / g1st / g2st;
  const double vrelst_invar = vrelst / denom;
  double b0 = P.b90_fact / (muRst * vrelst * vrelst_invar + 2.0 * EF_norm);
  const double bmin_qm = P.bqm_fact / (muRst * vrelst + sqrt(2.0 * EF_norm * muRst));
  double Clog = P.Clog;
  if (Clog == 0.0 && upstsq > 0.0) {
    Clog = 0.5 * log((b0 * b0 / 4.0 + bmax * bmax) / (b0 * b0 / 4.0 + bmin_qm * bmin_qm));
    Clog = fmax(2.0, Clog);
  }
  b0 = P.b90_fact / (muRst * vrelst * vrelst_invar);
  double sigma_eff = PI * b0 * b0 * Clog;
  sigma_eff = fmin(sigma_eff, sigma_max);
  double s12 = sigma_eff * den12 * vrelst * CVAC * P.dt_sec;
  s12 *= g1st * g2st / g1 / g2;
  double costh, sinth, sinphi, cosphi;
  bool skip_small = false;
  if (P.large_angle) skip_small = coulomb_large_angle(s12, Clog, b0, bmin_qm, sigma_eff, ularge, costh, sinth);
  if (s12o) *s12o = skip_small ? -1.0 : s12;
  if (!skip_small) coulomb_polar(P.angular, s12, gauss, upol, costh, sinth, Clog, b0, bmin_qm, sigma_eff, ufas2, ufas3, fas);
  sincos(2.0 * PI * uphi, &sinphi, &cosphi);
  rotate_velocity(upst, costh, sinth, cosphi, sinphi);
  ucmdotup = gcm * (vcm[0] * upst[0] + vcm[1] * upst[1] + vcm[2] * upst[2]);
  upst_fact = gcm * (ucmdotup / (1.0 + gcm) + g1st);
#pragma unroll
  for (int n = 0; n < 3; ++n) up1[n] = upst[n] + upst_fact * vcm[n];
  if (scatter2) {
#pragma unroll
    for (int n = 0; n < 3; ++n) up2[n] = (ptot[n] - m1 * up1[n]) / m2;
  }
  return true;
}

__global__ void k_coulomb_lorentz(long n, CoulParams P, const double *vp1, const double *vp2, const int *scatter2,
                                  const double *EF, const double *den12, const double *bmax, const double *smax,
                                  const double *gauss, const double *upol, const double *uphi, double *o1, double *o2,
                                  double *s12) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[3] = {vp1[i], vp1[n + i], vp1[2 * n + i]}, b[3] = {vp2[i], vp2[n + i], vp2[2 * n + i]};
  double s = 0.0;
  coulomb_lorentz_scatter(P, a, b, scatter2[i] != 0, P.mass1, P.mass2, EF[i], den12[i], bmax[i], smax[i], gauss[i],
                          upol[i], uphi[i], &s, P.large_draw, P.fas_draw2, P.fas_draw3);
  for (int c = 0; c < 3; ++c) {
    o1[c * n + i] = a[c];
    o2[c * n + i] = b[c];
  }
  s12[i] = s;
}

__global__ void k_coulomb_delta_u(long n, CoulParams P, const double *vp1, const double *vp2, const double *EF,
                                  const double *den12, const double *bmax, const double *smax, const double *gauss,
                                  const double *upol, const double *uphi, double *dU, double *s12) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a[3] = {vp1[i], vp1[n + i], vp1[2 * n + i]}, b[3] = {vp2[i], vp2[n + i], vp2[2 * n + i]};
  double d[3], s = 0.0;
  coulomb_delta_u(P, a, b, EF[i], den12[i], bmax[i], smax[i], gauss[i], upol[i], uphi[i], d, &s, P.large_draw, P.fas_draw2,
                  P.fas_draw3);
  dU[i] = d[0];
  dU[n + i] = d[1];
  dU[2 * n + i] = d[2];
  s12[i] = s;
}

struct CellCtx {
  double EF_norm, bmax, sigma_max;
  unsigned gcell;
};

// one pair: draws from Philox(cell, pair), weight rejection as Coulomb.cpp:561-584 / 1149-1172
// LEAN: compiled for the plain case (Galilean build, PROBABILISTIC weights, no large-angle events): fewer registers
template <bool FAS, bool LEAN = false>
__device__ __forceinline__ void coulomb_pair(const CoulParams &P, const CellCtx &C, double *a0, double *a1, double *a2,
                                             const double *wa, int pa, double *b0, double *b1, double *b2,
                                             const double *wb, int pb, double den_fact, unsigned pair_id,
                                             unsigned salt) {
  u4 c;
  c.x = pair_id;
  c.y = C.gcell;
  c.z = P.step_lo;
  c.w = P.step_hi ^ (STREAM_PAIR << 16) ^ salt;
  const u4 r = philox4x32_10(c, P.seed_lo, P.seed_hi);
  const double TWOPI = 6.28318530717958647692;
  // the N(0,1) draw only enters TAKIZUKA's small-angle branch (SetPolarScattering :1868-1874)
  const double gauss = (P.angular == ANG_TAKIZUKA) ? sqrt(-2.0 * log(u01(r.x))) * cos(TWOPI * u01(r.y)) : 0.0;
  const double w1 = wa[pa], w2 = wb[pb];
  const double den12 = fmax(w1, w2) * den_fact;
  double va[3] = {a0[pa], a1[pa], a2[pa]}, vb[3] = {b0[pb], b1[pb], b2[pb]}, dU[3];
  double ularge = 0.5, ufas2 = 0.5, ufas3 = 0.5;
  if ((!LEAN && P.large_angle) || (FAS && (P.angular == ANG_NANBU_FAS || P.angular == ANG_NANBU_FAS_V2))) {
    // the extra uniforms of SetPolarScattering come from the pair's weight stream: word 1 the large-angle event, words 2
    // and 3 the second and third draw of the full-angle models
    c.w = P.step_hi ^ (STREAM_WEIGHT << 16) ^ salt;
    const u4 wr = philox4x32_10(c, P.seed_lo, P.seed_hi);
    ularge = u01(wr.y);
    ufas2 = u01(wr.z);
    ufas3 = u01(wr.w);
  }
  if (!LEAN && P.rel) {
    // Coulomb.cpp:548-559 / 1139-1150: the lighter-weight particle goes first and always scatters, the other one
    // with probability w_min / w_max
    bool other = true;
    if ((float)w1 != (float)w2) {
      c.w = P.step_hi ^ (STREAM_WEIGHT << 16) ^ salt;
      const double ur = u01(philox4x32_10(c, P.seed_lo, P.seed_hi).x);
      other = !(ur > fmin(w1, w2) / fmax(w1, w2));
    }
    if ((float)w2 < (float)w1)
      coulomb_lorentz_scatter(P, vb, va, other, P.mass2, P.mass1, C.EF_norm, den12, C.bmax, C.sigma_max, gauss,
                              u01(r.z), u01(r.w), nullptr, ularge, ufas2, ufas3, FAS);
    else
      coulomb_lorentz_scatter(P, va, vb, other, P.mass1, P.mass2, C.EF_norm, den12, C.bmax, C.sigma_max, gauss,
                              u01(r.z), u01(r.w), nullptr, ularge, ufas2, ufas3, FAS);
    a0[pa] = va[0];
    a1[pa] = va[1];
    a2[pa] = va[2];
    b0[pb] = vb[0];
    b1[pb] = vb[1];
    b2[pb] = vb[2];
    return;
  }
  coulomb_delta_u(P, va, vb, C.EF_norm, den12, C.bmax, C.sigma_max, gauss, u01(r.z), u01(r.w), dU, nullptr, ularge, ufas2,
                  ufas3, FAS, LEAN);
  if (!LEAN && P.sk08 && (float)w1 != (float)w2) {
    // weight_method = CONSERVATIVE: Sentoku & Kemp, JCP 227 (2008) (Coulomb.cpp:849-897 / 1575-1621).  The lighter
    // particle scatters; the heavier one moves by the fraction w_min / w_max of its scattered change, and a kick normal to
    // its velocity (random azimuth) makes up the energy E_before + ratio (E_scattered - E_before) exactly
    // (Coulomb::enforceEnergyConservation, Coulomb.H:796-823)
    c.w = P.step_hi ^ (STREAM_WEIGHT << 16) ^ salt;
    const double phi2 = 2.0 * u01(philox4x32_10(c, P.seed_lo, P.seed_hi).x);   // phi / pi
    const bool first_light = (float)w1 < (float)w2;
    double *vl = first_light ? va : vb, *vh = first_light ? vb : va;
    const double fl = first_light ? P.f1 : -P.f2, fh = first_light ? -P.f2 : P.f1;
    const double mh = first_light ? P.mass2 : P.mass1;
    const double ratio = first_light ? w1 / w2 : w2 / w1;
    double vhp[3], nb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      vl[k] += fl * dU[k];
      vhp[k] = vh[k] + fh * dU[k];
    }
    const double Ebefore = mh * (vh[0] * vh[0] + vh[1] * vh[1] + vh[2] * vh[2]) / 2.0;
    const double Escatter = mh * (vhp[0] * vhp[0] + vhp[1] * vhp[1] + vhp[2] * vhp[2]) / 2.0;
    const double Eafter = Ebefore + ratio * (Escatter - Ebefore);
#pragma unroll
    for (int k = 0; k < 3; ++k) nb[k] = vh[k] + ratio * (vhp[k] - vh[k]);
    double br = nb[0] * nb[0] + nb[1] * nb[1], bm = br + nb[2] * nb[2];
    const double Eafter2 = mh / 2.0 * bm;
    bm = sqrt(bm);
    br = sqrt(br);
    if (!(Eafter < Eafter2)) {
      const double dmag = sqrt(2.0 / mh * (Eafter - Eafter2));
      double sphi, cphi;
      sincospi(phi2, &sphi, &cphi);
      const double d0 = (nb[2] * nb[0] * cphi - bm * nb[1] * sphi) / br * dmag / bm;
      const double d1 = (nb[2] * nb[1] * cphi + bm * nb[0] * sphi) / br * dmag / bm;
      const double d2 = -br * cphi * dmag / bm;
      nb[0] += d0, nb[1] += d1, nb[2] += d2;
    }
    vh[0] = nb[0], vh[1] = nb[1], vh[2] = nb[2];
    a0[pa] = va[0], a1[pa] = va[1], a2[pa] = va[2];
    b0[pb] = vb[0], b1[pb] = vb[1], b2[pb] = vb[2];
    return;
  }
  bool s1 = true, s2 = true;
  if ((float)w1 != (float)w2) {
    c.w = P.step_hi ^ (STREAM_WEIGHT << 16) ^ salt;
    const double ur = u01(philox4x32_10(c, P.seed_lo, P.seed_hi).x);
    if ((float)w1 < (float)w2) s2 = ur < w1 / w2;
    else s1 = ur < w2 / w1;
  }
  if (s1) {
    a0[pa] = va[0] + P.f1 * dU[0];
    a1[pa] = va[1] + P.f1 * dU[1];
    a2[pa] = va[2] + P.f1 * dU[2];
  }
  if (s2) {
    b0[pb] = vb[0] - P.f2 * dU[0];
    b1[pb] = vb[1] - P.f2 * dU[1];
    b2[pb] = vb[2] - P.f2 * dU[2];
  }
}

__device__ __forceinline__ void ta_from_coul(const CoulParams &P, TAParams &T) {
  T.seed_lo = P.seed_lo;
  T.seed_hi = P.seed_hi;
  T.step_lo = P.step_lo;
  T.step_hi = P.step_hi;
}

#ifndef PGPU_COUL_LEAN_MINB
#define PGPU_COUL_LEAN_MINB 5
#endif
// Coulomb::applyIntraScattering_PROB (:400-592)
template <bool FAS, bool LEAN>
__global__ void __launch_bounds__(LEAN ? 128 : 256, LEAN ? PGPU_COUL_LEAN_MINB : 1)
k_coulomb_intra(const int *cell_start, int ncell, double *v0, double *v1, double *v2, const double *w,
                const uint64_t *id, const double *dens, const double *LDe, CoulParams P, unsigned *key, int *order,
                unsigned long long *npairs) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const double numDen = dens[cell];
  const int s = cell_start[cell], n = cell_start[cell + 1] - s;
  if (numDen == 0.0 || n < 2) return;
  const double PI = 3.14159265358979323846;
  CellCtx C;
  C.EF_norm = P.EF_fact * pow(numDen, 2.0 / 3.0);
  C.bmax = LDe[cell];
  C.sigma_max = 1.0 / (numDen * (1.0 / cbrt(4.0 / 3.0 * PI * numDen)));
  C.gcell = global_cell(P, cell);
  TAParams T;
  ta_from_coul(P, T);
  warp_shuffle_order(s, n, id, T, 4u, key, order, lane);
  const bool NxN = P.NxN || n < P.NxN_Nthresh;
  const double Naa = (double)(n - 1);
  unsigned long long mine = 0;
  if (NxN) {
    // round-robin tournament over M = n (+1 if odd) players: M-1 rounds of M/2 disjoint pairs
    const int M = n + (n & 1), R = M - 1;
    const double den_fact = 1.0 / P.cellV_SI;
    for (int r = 0; r < R; ++r) {
      for (int i = lane; i < M / 2; i += 32) {
        int pa = (i == 0) ? M - 1 : (r + i) % R;
        int pb = (i == 0) ? r : (r - i + R) % R;
        if (pa < n && pb < n) {
          if (pa > pb) {
            const int t = pa;
            pa = pb;
            pb = t;
          }
          coulomb_pair<FAS, LEAN>(P, C, v0, v1, v2, w, s + order[s + pa], v0, v1, v2, w, s + order[s + pb], den_fact,
                       (unsigned)(pa * 65536 + pb), 0u);
          ++mine;
        }
      }
      __syncwarp();
    }
  } else {
    const int pstart = (n & 1) ? 3 : 0;
    if (pstart == 3 && lane == 0) {
      // odd cell: (0,1), (0,2), (1,2) with half the density (:468-470, 536-538)
      for (int q = 0; q < 3; ++q) {
        const int p1 = q == 2 ? 1 : 0, p2 = q == 0 ? 1 : 2;
        coulomb_pair<FAS, LEAN>(P, C, v0, v1, v2, w, s + order[s + p1], v0, v1, v2, w, s + order[s + p2],
                                Naa / P.cellV_SI / 2.0, (unsigned)(p1 * 65536 + p2), 1u);
        ++mine;
      }
    }
    const int nmain = (n - pstart) / 2;
    for (int q = lane; q < nmain; q += 32) {
      const int pa = pstart + 2 * q, pb = pa + 1;
      coulomb_pair<FAS, LEAN>(P, C, v0, v1, v2, w, s + order[s + pa], v0, v1, v2, w, s + order[s + pb], Naa / P.cellV_SI,
                   (unsigned)(pa * 65536 + pb), 0u);
      ++mine;
    }
  }
  mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
  if (lane == 0) atomicAdd(npairs, mine);
}

// Coulomb::applyInterScattering_PROB (:919-1180)
template <bool FAS, bool LEAN>
__global__ void __launch_bounds__(LEAN ? 128 : 256, LEAN ? PGPU_COUL_LEAN_MINB : 1)
k_coulomb_inter(const int *cs1, const int *cs2, int ncell, double *a0, double *a1, double *a2, const double *wa,
                const uint64_t *id1, const double *dens1, double *b0, double *b1, double *b2, const double *wb,
                const uint64_t *id2, const double *dens2, const double *LDe, CoulParams P, unsigned *key1, int *order1,
                unsigned *key2, int *order2, unsigned long long *npairs) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const double numDen1 = dens1[cell], numDen2 = dens2[cell];
  if (numDen1 * numDen2 == 0.0) return;
  const int s1 = cs1[cell], n1 = cs1[cell + 1] - s1;
  const int s2 = cs2[cell], n2 = cs2[cell + 1] - s2;
  if ((long)n1 * n2 < 2) return;
  const double PI = 3.14159265358979323846;
  const double minn = fmin(numDen1, numDen2), maxn = fmax(numDen1, numDen2);
  CellCtx C;
  C.EF_norm = P.EF_fact * pow(maxn, 2.0 / 3.0);
  C.bmax = LDe[cell];
  C.sigma_max = 1.0 / (minn * (1.0 / cbrt(4.0 / 3.0 * PI * minn)));
  C.gcell = global_cell(P, cell);
  TAParams T;
  ta_from_coul(P, T);
  warp_shuffle_order(s1, n1, id1, T, 5u, key1, order1, lane);
  warp_shuffle_order(s2, n2, id2, T, 6u, key2, order2, lane);
  const int Nmin = min(n1, n2), Nmax = max(n1, n2);
  const bool first_short = (Nmin == n1);
  const bool NxN = P.NxN || Nmin < P.NxN_Nthresh;
  unsigned long long mine = 0;
  if (NxN) {
    // all Nmin x Nmax pairs: chunks of Nmin long-list particles, Latin-square rounds inside a chunk
    const double den_fact = 1.0 / P.cellV_SI;
    for (int cb = 0; cb < Nmax; cb += Nmin) {
      const int nc = min(Nmin, Nmax - cb);
      for (int rr = 0; rr < Nmin; ++rr) {
        for (int t = lane; t < nc; t += 32) {
          const int pl = cb + t, ps = (t + rr) % Nmin;
          const int i1 = s1 + order1[s1 + (first_short ? ps : pl)];
          const int i2 = s2 + order2[s2 + (first_short ? pl : ps)];
          coulomb_pair<FAS, LEAN>(P, C, a0, a1, a2, wa, i1, b0, b1, b2, wb, i2, den_fact, (unsigned)(pl * 65536 + ps), 2u);
          ++mine;
        }
        __syncwarp();
      }
    }
  } else {
    const double den_fact = (double)Nmin / P.cellV_SI;
    for (int r = lane; r < Nmin; r += 32) {
      for (int p = r; p < Nmax; p += Nmin) {
        const int i1 = s1 + order1[s1 + (first_short ? r : p)];
        const int i2 = s2 + order2[s2 + (first_short ? p : r)];
        coulomb_pair<FAS, LEAN>(P, C, a0, a1, a2, wa, i1, b0, b1, b2, wb, i2, den_fact, (unsigned)p, 2u);
        ++mine;
      }
    }
  }
  mine = __reduce_add_sync(0xffffffffu, (unsigned)mine);
  if (lane == 0) atomicAdd(npairs, mine);
}

// =============================================================================================
// Elastic::electronImpact (src/scattering/Elastic.cpp:225-388), PROBABILISTIC weights.
// Every species-1 particle of a cell picks a random species-2 partner; lanes that picked the
// same partner in one batch go one after the other (__match_any_sync), as the reference's
// sequential loop would.
// =============================================================================================
// HardSphere, PROBABILISTIC weight method (src/scattering/HardSphere.cpp:223-418 self, 419-665 inter):
// no-time-counter pairs.  The candidates of a cell are sequential by construction (a collision changes the
// velocities the next candidate sees), so one thread owns a cell; the parallelism is the number of cells.
// Draws: Philox keyed by (candidate, global cell, step); the reference's "redraw until different" for the
// second index of a self pair is taken as a uniform pick among the other N-1 particles (same distribution).
// =============================================================================================
enum { STREAM_HS = 0x4853u };
struct HSParams {
  double sigmaT, dt_sec, mass1, mass2, mu, Vc;
  int vhs;                          // VariableHardSphere: sigmaT(g) = fourPiA * g^(-fourOverAlpha), self only
  double fourPiA, fourOverAlpha;
  int conservative;                 // HardSphere weight method CONSERVATIVE (collapseThreeToTwo), needs wmut (species 1)
  double *wmut;
  double *wmut2;                    // species 2 (inter-species CONSERVATIVE)
  unsigned seed_lo, seed_hi, step_lo, step_hi;
  int box_lo0, box_lo1, nbox0, ncell_glob0;
};
__device__ __forceinline__ u4 hs_draw(const HSParams &P, unsigned k, unsigned gcell, unsigned sub) {
  u4 c;
  c.x = k;
  c.y = gcell;
  c.z = P.step_lo;
  c.w = P.step_hi ^ (STREAM_HS << 16) ^ sub;
  return philox4x32_10(c, P.seed_lo, P.seed_hi);
}
__global__ void __launch_bounds__(128)
k_hard_sphere(const int *cs1, const int *cs2, int ncell, double *a0, double *a1, double *a2, const double *wa,
              const double *dens1, const double *ene1, double *b0, double *b1, double *b2, const double *wb,
              const double *dens2, const double *ene2, HSParams P, int self, unsigned long long *ncoll) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned mine = 0;
  if (cell < ncell) {
    const double CVAC = 2.99792458e+08, TWOPI = 6.28318530717958647692, cvacSq = CVAC * CVAC;
    const int s1 = cs1[cell], n1 = cs1[cell + 1] - s1, s2 = cs2[cell], n2 = cs2[cell + 1] - s2;
    const double nd1 = dens1[cell], nd2 = dens2[cell];
    const unsigned gcell = (unsigned)((cell % P.nbox0 + P.box_lo0) + (cell / P.nbox0 + P.box_lo1) * P.ncell_glob0);
    double gmax = 0.0, Nmax = 0.0, sigmaTmax = P.sigmaT;
    bool go;
    if (self) {
      go = nd1 != 0.0 && n1 >= 2;
      if (go) {
        double e = 0.0;
        for (int d = 0; d < 3; ++d) e = e + ene1[(size_t)d * ncell + cell];
        const double Teff = 2.0 / 3.0 * e / nd1 * cvacSq;
        gmax = 5.0 * sqrt(Teff / P.mass1);
        if (P.vhs) {   // VariableHardSphere.cpp:279-289: sigmaT at gmax, no cap on nuMax*dt
          sigmaTmax = P.fourPiA * pow(gmax, -P.fourOverAlpha);
          Nmax = 0.5 * (n1 - 1) * (nd1 * sigmaTmax * gmax * P.dt_sec);
        } else {
          const double nuMaxDt = nd1 * P.sigmaT * gmax * P.dt_sec;
          Nmax = 0.5 * (n1 - 1) * fmin(nuMaxDt, 1.0);
        }
      }
    } else {
      go = nd1 * nd2 != 0.0 && !(n1 < 2 && n2 < 2) && n1 >= 1 && n2 >= 1;
      if (go) {
        double e1 = 0.0, e2 = 0.0;
        for (int d = 0; d < 3; ++d) {
          e1 = e1 + ene1[(size_t)d * ncell + cell];
          e2 = e2 + ene2[(size_t)d * ncell + cell];
        }
        const double Teff1 = 2.0 / 3.0 * e1 / nd1 * cvacSq, Teff2 = 2.0 / 3.0 * e2 / nd2 * cvacSq;
        gmax = 2.5 * sqrt(2.0 * fmax(Teff1, Teff2) / P.mu);
        const double W1 = nd1 / n1 * P.Vc, W2 = nd2 / n2 * P.Vc;
        Nmax = fmax(W1, W2) * n1 * n2 / P.Vc * P.sigmaT * gmax * P.dt_sec;
      }
    }
    if (go) {
      const double whole = floor(Nmax), rem = Nmax - whole;
      const double u0 = u01(hs_draw(P, 0xffffffffu, gcell, 2u).x);
      int Nint = (int)whole;
      if ((self && !P.vhs) ? (u0 <= rem) : (u0 < rem)) Nint += 1;
      const double f1 = self ? 0.5 : P.mu / P.mass1, f2 = self ? 0.5 : P.mu / P.mass2;
      for (int k = 0; k < Nint; ++k) {
        const u4 r0 = hs_draw(P, (unsigned)k, gcell, 0u);
        int q1 = min(n1 - 1, (int)(u01(r0.x) * n1)), q2;
        if (self) {
          q2 = min(n1 - 2, (int)(u01(r0.y) * (n1 - 1)));
          if (q2 >= q1) q2 += 1;
        } else {
          q2 = min(n2 - 1, (int)(u01(r0.y) * n2));
        }
        const int i1 = s1 + q1, i2 = s2 + q2;
        const double v1[3] = {a0[i1], a1[i1], a2[i1]}, v2[3] = {b0[i2], b1[i2], b2[i2]};
        const double ux = v1[0] - v2[0], uy = v1[1] - v2[1], uz = v1[2] - v2[2];
        const double g12 = sqrt(ux * ux + uy * uy + uz * uz) * CVAC;
        const double sigT = P.vhs ? P.fourPiA * pow(g12, -P.fourOverAlpha) : P.sigmaT;
        const double q12 = g12 * sigT / (gmax * sigmaTmax);
        if (u01(r0.z) > q12) continue;
        mine += 1;
        const u4 r1 = hs_draw(P, (unsigned)k, gcell, 1u);
        const double costh = 1.0 - 2.0 * u01(r0.w);
        const double sinth = sqrt(1.0 - costh * costh);
        double sinphi, cosphi, dU[3];
        sincos(TWOPI * u01(r1.x), &sinphi, &cosphi);
        scatter_delta_u(ux, uy, uz, costh, sinth, cosphi, sinphi, dU);
        const double wp1 = wa[i1], wp2 = wb[i2], u3 = P.vhs ? 0.0 : u01(r1.y);   // VHS updates both partners
        if (P.conservative && wp1 != wp2 && !self) {
          // HardSphere.cpp:594-636, between species: the lighter particle moves by 0.5 deltaU (the reference's factor
          // here, not mu/m: its quirk, kept), the heavier one's scattered copy by -0.5 deltaU; the heavier particle,
          // that copy and a third particle of the heavier particle's species are merged (collapseThreeToTwo)
          const bool first_light = wp1 < wp2;
          const int nh = first_light ? n2 : n1, sh = first_light ? s2 : s1, qh = first_light ? q2 : q1;
          if (nh < 2) continue;
          int q3 = min(nh - 2, (int)(u01(r1.z) * (nh - 1)));
          if (q3 >= qh) q3 += 1;
          double *h0 = first_light ? b0 : a0, *h1 = first_light ? b1 : a1, *h2 = first_light ? b2 : a2;
          double *hw = first_light ? P.wmut2 : P.wmut;
          const int ih = sh + qh, i3 = sh + q3;
          const double wl = first_light ? wp1 : wp2, wh = first_light ? wp2 : wp1, w3 = hw[i3];
          double vh[3], vhp[3], v3[3] = {h0[i3], h1[i3], h2[i3]};
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            vh[d] = first_light ? v2[d] : v1[d];
            vhp[d] = first_light ? v2[d] - 0.5 * dU[d] : v1[d] + 0.5 * dU[d];
          }
          const double wp23 = 0.5 * (wh + w3);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double c23 = (wl * vhp[d] + (wh - wl) * vh[d] + w3 * v3[d]) / wp23;
            const double d23 = (wl * vhp[d] * vhp[d] + (wh - wl) * vh[d] * vh[d] + w3 * v3[d] * v3[d]) / wp23;
            const double root = sqrt(fmax(2.0 * d23 - c23 * c23, 0.0));
            vh[d] = 0.5 * (c23 + root);
            v3[d] = 0.5 * (c23 - root);
          }
          if (first_light) {
            a0[i1] = v1[0] + 0.5 * dU[0]; a1[i1] = v1[1] + 0.5 * dU[1]; a2[i1] = v1[2] + 0.5 * dU[2];
          } else {
            b0[i2] = v2[0] - 0.5 * dU[0]; b1[i2] = v2[1] - 0.5 * dU[1]; b2[i2] = v2[2] - 0.5 * dU[2];
          }
          h0[ih] = vh[0]; h1[ih] = vh[1]; h2[ih] = vh[2];
          h0[i3] = v3[0]; h1[i3] = v3[1]; h2[i3] = v3[2];
          hw[ih] = wp23;
          hw[i3] = wp23;
          continue;
        }
        if (P.conservative && wp1 != wp2) {
          // HardSphere.cpp:357-392: the lighter particle scatters; the heavier one, its scattered fraction and a third
          // particle of the cell are merged into two equally weighted particles (ScatteringUtils::collapseThreeToTwo)
          if (n1 < 3) continue;
          int q3 = min(n1 - 3, (int)(u01(r1.z) * (n1 - 2)));
          const int qlo = min(q1, q2), qhi = max(q1, q2);
          if (q3 >= qlo) q3 += 1;
          if (q3 >= qhi) q3 += 1;
          const int i3 = s1 + q3;
          const bool first_light = wp1 < wp2;
          const int ih = first_light ? i2 : i1;                    // the heavier particle
          const double wl = first_light ? wp1 : wp2;
          double vh[3], vhp[3], vl[3], v3[3] = {a0[i3], a1[i3], a2[i3]};
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            if (first_light) {
              vl[d] = v1[d] + 0.5 * dU[d];
              vh[d] = v2[d];
              vhp[d] = v2[d] - 0.5 * dU[d];
            } else {
              vh[d] = v1[d];
              vhp[d] = v1[d] + 0.5 * dU[d];
              vl[d] = v2[d] - 0.5 * dU[d];
            }
          }
          const double wh = first_light ? wp2 : wp1, w3 = P.wmut[i3];
          const double wp23 = 0.5 * (wh + w3);
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const double c23 = (wl * vhp[d] + (wh - wl) * vh[d] + w3 * v3[d]) / wp23;
            const double d23 = (wl * vhp[d] * vhp[d] + (wh - wl) * vh[d] * vh[d] + w3 * v3[d] * v3[d]) / wp23;
            const double root = sqrt(fmax(2.0 * d23 - c23 * c23, 0.0));   // the reference asserts arg >= 0
            vh[d] = 0.5 * (c23 + root);
            v3[d] = 0.5 * (c23 - root);
          }
          const int il = first_light ? i1 : i2;
          a0[il] = vl[0]; a1[il] = vl[1]; a2[il] = vl[2];
          a0[ih] = vh[0]; a1[ih] = vh[1]; a2[ih] = vh[2];
          a0[i3] = v3[0]; a1[i3] = v3[1]; a2[i3] = v3[2];
          P.wmut[ih] = wp23;
          P.wmut[i3] = wp23;
          continue;
        }
        if (P.vhs || u3 <= wp2 / wp1) {
          a0[i1] = v1[0] + f1 * dU[0];
          a1[i1] = v1[1] + f1 * dU[1];
          a2[i1] = v1[2] + f1 * dU[2];
        }
        if (P.vhs || u3 <= wp1 / wp2) {
          b0[i2] = v2[0] - f2 * dU[0];
          b1[i2] = v2[1] - f2 * dU[1];
          b2[i2] = v2[2] - f2 * dU[2];
        }
      }
    }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(ncoll, (unsigned long long)mine);
}

// =============================================================================================
struct ElaParams {
  double mu, f1, f2, const_sigma, dt_sec, mcSq;
  int ntab, angular, loglog;
  int conservative;   // weight_method = CONSERVATIVE (Elastic.cpp:334-356): wmut = the weights of species 2, rewritten
  double *wmut;
  const double *E, *Q, *XI;
  unsigned seed_lo, seed_hi, step_lo, step_hi;
};

// Elastic::getSigma / getTextSigma (:390-476) incl. the reference's interpolation formulas
__device__ __forceinline__ double elastic_sigma(const ElaParams &P, double g12, double &xi) {
  xi = 0.0;
  if (P.ntab == 0) return P.const_sigma;
  const double KE = P.mu * P.mcSq * g12 * g12 / 2.0;
  const double *E = P.E, *Q = P.Q, *XI = P.XI;
  const int N = P.ntab;
  double sigma = 0.0;
  if (KE >= E[N - 1]) {
    if (P.angular == 0) sigma = Q[N - 1] * log(KE) / log(E[N - 1]) * E[N - 1] / KE;
    else {
      sigma = Q[N - 1] * E[N - 1] / KE;
      xi = XI[N - 1];
    }
    return sigma;
  }
  int i = N / 2;
  while (KE < E[i]) i--;
  while (KE > E[i + 1]) i++;
  const double l0 = log10(KE), lu = log10(E[i]), ld = log10(E[i + 1]);
  if (P.loglog && Q[i] * E[i] > 0.0) sigma = pow(10.0, (log10(Q[i + 1]) * (l0 - ld) + log10(Q[i]) * (lu - l0)) / (lu - ld));
  else sigma = (Q[i + 1] * (KE - E[i]) + Q[i] * (E[i + 1] - KE)) / (E[i + 1] - E[i]);
  if (P.angular == 1) {
    if (E[i] * KE > 0.0) xi = (XI[i + 1] * (l0 - ld) + XI[i] * (lu - l0)) / (lu - ld);
    else xi = (XI[i + 1] * (KE - E[i]) + XI[i] * (E[i + 1] - KE)) / (E[i + 1] - E[i]);
  }
  return sigma;
}

__global__ void __launch_bounds__(256)
k_elastic(const int *cs1, const int *cs2, int ncell, double *a0, double *a1, double *a2, const double *wa,
          const uint64_t *id1, double *b0, double *b1, double *b2, const double *wb, const double *dens2, ElaParams P,
          unsigned long long *ncoll) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const int s1 = cs1[cell], n1 = cs1[cell + 1] - s1;
  const int s2 = cs2[cell], n2 = cs2[cell + 1] - s2;
  if (n1 < 1 || n2 < 1) return;
  const double numDen2 = dens2[cell];
  const double TWOPI = 6.28318530717958647692, CVAC = 2.99792458e+08;
  unsigned mine = 0;
  for (int base = 0; base < n1; base += 32) {
    const int q = base + lane;
    const bool have = q < n1;
    int i1 = 0, i2 = -1 - lane;   // distinct dummies for idle lanes
    u4 r0, r1;
    if (have) {
      i1 = s1 + q;
      const uint64_t pid = id1[i1];
      u4 c;
      c.x = (unsigned)pid;
      c.y = (unsigned)(pid >> 32);
      c.z = P.step_lo;
      c.w = P.step_hi ^ (STREAM_ELA << 16);
      r0 = philox4x32_10(c, P.seed_lo, P.seed_hi);
      c.w ^= 1u;
      r1 = philox4x32_10(c, P.seed_lo, P.seed_hi);
      i2 = s2 + min(n2 - 1, (int)(u01(r0.x) * n2));     // MathUtils::randInt(0, n2-1)
    }
    bool pending = have;
    while (__any_sync(0xffffffffu, pending)) {
      bool turn;
      if (P.conservative) {
        // a merge also rewrites a third particle of species 2: the projectiles of a cell go one at a time
        const unsigned pend = __ballot_sync(0xffffffffu, pending);
        turn = pending && (__ffs(pend) - 1 == lane);
      } else {
        const unsigned same = __match_any_sync(0xffffffffu, pending ? i2 : -1 - lane);
        turn = pending && (__ffs(same) - 1 == lane);
      }
      if (turn) {
        const double va[3] = {a0[i1], a1[i1], a2[i1]}, vb[3] = {b0[i2], b1[i2], b2[i2]};
        const double ux = va[0] - vb[0], uy = va[1] - vb[1], uz = va[2] - vb[2];
        const double g12 = sqrt(ux * ux + uy * uy + uz * uz);
        double xi;
        const double sigma = elastic_sigma(P, g12, xi);
        if (sigma != 0.0) {
          const double q12 = 1.0 - exp(-(g12 * CVAC * sigma * numDen2 * P.dt_sec));
          if (u01(r0.y) <= q12) {
            ++mine;
            double sinphi, cosphi;
            sincos(TWOPI * u01(r0.z), &sinphi, &cosphi);
            const double R = u01(r0.w);
            const double costh = 1.0 - 2.0 * R * (1.0 - xi) / (1.0 + xi * (1.0 - 2.0 * R));   // getScatteringCos
            const double sinth = sqrt(1.0 - costh * costh);
            double dU[3];
            scatter_delta_u(ux, uy, uz, costh, sinth, cosphi, sinphi, dU);
            const double r2 = u01(r1.x), w1 = wa[i1], w2 = P.conservative ? P.wmut[i2] : wb[i2];
            if (P.conservative && w1 < w2) {
              // Elastic.cpp:334-356: the lighter projectile scatters; the target, its scattered fraction w1 and a second
              // target of the cell become two equally weighted particles (ScatteringUtils::collapseThreeToTwo, :20-47)
              if (n2 >= 2) {
                const int q2 = i2 - s2;
                int q3 = min(n2 - 2, (int)(u01(r1.y) * (n2 - 1)));   // uniform over the other n2 - 1 (the reference redraws)
                if (q3 >= q2) q3 += 1;
                const int i3 = s2 + q3;
                const double w3 = P.wmut[i3], wp23 = 0.5 * (w2 + w3);
                const double v3[3] = {b0[i3], b1[i3], b2[i3]};
                double nh[3], n3[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                  const double vhp = vb[d] - P.f2 * dU[d];
                  const double c23 = (w1 * vhp + (w2 - w1) * vb[d] + w3 * v3[d]) / wp23;
                  const double d23 = (w1 * vhp * vhp + (w2 - w1) * vb[d] * vb[d] + w3 * v3[d] * v3[d]) / wp23;
                  const double root = sqrt(fmax(2.0 * d23 - c23 * c23, 0.0));   // the reference asserts arg >= 0
                  nh[d] = 0.5 * (c23 + root);
                  n3[d] = 0.5 * (c23 - root);
                }
                a0[i1] = va[0] + P.f1 * dU[0];
                a1[i1] = va[1] + P.f1 * dU[1];
                a2[i1] = va[2] + P.f1 * dU[2];
                b0[i2] = nh[0]; b1[i2] = nh[1]; b2[i2] = nh[2];
                b0[i3] = n3[0]; b1[i3] = n3[1]; b2[i3] = n3[2];
                P.wmut[i2] = wp23;
                P.wmut[i3] = wp23;
              }
            } else {
            if (r2 <= w2 / w1) {
              a0[i1] = va[0] + P.f1 * dU[0];
              a1[i1] = va[1] + P.f1 * dU[1];
              a2[i1] = va[2] + P.f1 * dU[2];
            }
            if (r2 <= w1 / w2) {
              b0[i2] = vb[0] - P.f2 * dU[0];
              b1[i2] = vb[1] - P.f2 * dU[1];
              b2[i2] = vb[2] - P.f2 * dU[2];
            }
            }
          }
        }
        pending = false;
      }
      __syncwarp();
    }
  }
  mine = __reduce_add_sync(0xffffffffu, mine);
  if (lane == 0 && mine) atomicAdd(ncoll, (unsigned long long)mine);
}


// ---- Scattering::setMeanFreeTime: box maximum of the per-cell collision frequency -----------
// (TakizukaAbe.cpp:80-238, Coulomb.cpp:108-356, Elastic.cpp:146-202).  One thread per cell; the
// maximum of the non-negative frequencies is taken on their bit patterns.
__device__ __forceinline__ void max_bits(unsigned long long *out, double v) {
  if (!(v > 0.0)) v = 0.0;   // also drops NaN (a cell with zero temperature); every lane takes part
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(v));
}

// MathUtils::gammainc (MathUtils.cpp:65-95), a = 3/2
__device__ double gammainc_3half(double x) {
  double soln = 0.88622692545275801365;   // tgamma(1.5)
  if (x < 10.0) {
    double sign = -1.0, factorial = 1.0;
    soln = 0.0;
    for (int p = 1; p < 41; p++) {
      const double coef = 0.5 + p;
      if (p > 1) factorial = factorial * (p - 1);
      sign = -sign;
      soln = soln + sign * pow(x, coef) / coef / factorial;
    }
  }
  return soln;
}

struct NuConsts {
  double CVAC, ME, QE, EP0, PI;
};
__device__ __forceinline__ NuConsts nu_consts() {
  NuConsts k;
  k.PI = 3.14159265358979323846;
  k.CVAC = 2.99792458e+08;
  k.ME = 9.10938370e-31;
  k.QE = 1.60217663e-19;
  k.EP0 = 1.0 / k.CVAC / k.CVAC / (4.0 * k.PI * 1.0e-7);
  return k;
}

__global__ void k_nu_max_ta(int ncell, const double *dens1, const double *ene1, const double *dens2,
                            const double *ene2, double charge1, double charge2, double mass1, double mass2,
                            double Clog, int intra, unsigned long long *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const NuConsts k = nu_consts();
  const double cvacSq = k.CVAC * k.CVAC, EV_PER_JOULE = 1.0 / k.QE;
  double nu = 0.0;
  if (c < ncell) {
    if (intra) {
      const double n = dens1[c];
      if (n != 0.0) {
        double e = 0.0;
        for (int d = 0; d < 3; ++d) e = e + ene1[(size_t)d * ncell + c];
        double T = k.ME * 2.0 / 3.0 * e / n * cvacSq;
        T = EV_PER_JOULE * T;
        double tau = 3.44e5 * pow(T, 1.5) / (n * (1.0 / 1.0e+06)) / Clog;
        tau = tau * sqrt(mass1 / 2.0) / pow(charge1 * charge2, 2.0);
        nu = 1.0 / tau;
      }
    } else {
      const double n1 = dens1[c], n2 = dens2[c];
      if (n1 * n2 != 0.0) {
        const double q = k.QE * charge1 * k.QE * charge2 / k.EP0;
        const double factor = q * q / (4.0 * k.PI);
        double e1 = 0.0, e2 = 0.0;
        for (int d = 0; d < 3; ++d) {
          e1 = e1 + ene1[(size_t)d * ncell + c];
          e2 = e2 + ene2[(size_t)d * ncell + c];
        }
        const double energy1 = k.ME * e1 / n1 * cvacSq, energy2 = k.ME * e2 / n2 * cvacSq;
        const double T1 = EV_PER_JOULE * 2.0 / 3.0 * energy1, T2 = EV_PER_JOULE * 2.0 / 3.0 * energy2;
        const double VT1 = sqrt(k.QE * T1 / (k.ME * mass1)), VT2 = sqrt(k.QE * T2 / (k.ME * mass2));
        const double x12 = (T1 / mass1) / (T2 / mass2), x21 = 1. / x12;
        const double psi12 = 2.0 / sqrt(k.PI) * gammainc_3half(x12);
        const double psi21 = 2.0 / sqrt(k.PI) * gammainc_3half(x21);
        const double nu012 = factor * Clog * n2 / (energy1 * energy1) * VT1;
        const double nu021 = factor * Clog * n1 / (energy2 * energy2) * VT2;
        nu = fmax((1.0 + mass1 / mass2) * psi12 * nu012, (1.0 + mass2 / mass1) * psi21 * nu021);
      }
    }
  }
  max_bits(out, nu);
}

__global__ void k_nu_max_coulomb(int ncell, const double *LDe, const double *dens1, const double *mom1,
                                 const double *ene1, const double *dens2, const double *mom2, const double *ene2,
                                 double mass1, double mass2, CoulParams P, int intra, unsigned long long *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const NuConsts k = nu_consts();
  const double cvacSq = k.CVAC * k.CVAC, mcSq_eV = k.ME * (1.0 / k.QE) * cvacSq;
  double nu = 0.0;
  if (c < ncell) {
    bool live = false;
    double g12sq = 0.0, nfreq = 0.0, sigma_max = 0.0, EF_norm = 0.0;
    const size_t N = (size_t)ncell;
    if (intra) {
      const double n = dens1[c];
      if (n != 0.0) {
        live = true;
        const double rho = mass1 * n;
        const double spacing = 1.0 / cbrt(4.0 / 3.0 * k.PI * n);
        sigma_max = 1.0 / (n * spacing);
        EF_norm = P.EF_fact * pow(n, 2.0 / 3.0);
        const double ux = mom1[c], uy = mom1[N + c], uz = mom1[2 * N + c];
        const double meanE = (ux * ux + uy * uy + uz * uz) / rho / 2.0;
        double e = 0.0;
        for (int d = 0; d < 3; ++d) e += ene1[d * N + c];
        double T = 2.0 / 3.0 * (e - meanE) / n * mcSq_eV;
        T = fmax(T, 0.01);
        g12sq = 6.0 * k.QE / k.ME * T / mass1;
        nfreq = n;
      }
    } else {
      const double n1 = dens1[c], n2 = dens2[c];
      if (n1 * n2 != 0.0) {
        live = true;
        const double rho1 = mass1 * n1, rho2 = mass2 * n2;
        const double minn = fmin(n1, n2), maxn = fmax(n1, n2);
        const double spacing = 1.0 / cbrt(4.0 / 3.0 * k.PI * minn);
        sigma_max = 1.0 / (minn * spacing);
        EF_norm = P.EF_fact * pow(maxn, 2.0 / 3.0);
        const double ux1 = mom1[c], uy1 = mom1[N + c], uz1 = mom1[2 * N + c];
        const double ux2 = mom2[c], uy2 = mom2[N + c], uz2 = mom2[2 * N + c];
        const double meanE1 = (ux1 * ux1 + uy1 * uy1 + uz1 * uz1) / rho1 / 2.0;
        const double meanE2 = (ux2 * ux2 + uy2 * uy2 + uz2 * uz2) / rho2 / 2.0;
        double e1 = 0.0, e2 = 0.0;
        for (int d = 0; d < 3; ++d) e1 += ene1[d * N + c];
        for (int d = 0; d < 3; ++d) e2 += ene2[d * N + c];
        double T1 = 2.0 / 3.0 * (e1 - meanE1) / n1 * mcSq_eV, T2 = 2.0 / 3.0 * (e2 - meanE2) / n2 * mcSq_eV;
        T1 = fmax(T1, 0.01);
        T2 = fmax(T2, 0.01);
        const double VT1 = sqrt(k.QE * T1 / (k.ME * mass1)), VT2 = sqrt(k.QE * T2 / (k.ME * mass2));
        g12sq = (3.0 * VT1 * VT1 + 3.0 * VT2 * VT2);
        const double dx = ux1 / rho1 - ux2 / rho2, dy = uy1 / rho1 - uy2 / rho2, dz = uz1 / rho1 - uz2 / rho2;
        g12sq += dx * dx * cvacSq;
        g12sq += dy * dy * cvacSq;
        g12sq += dz * dz * cvacSq;
        nfreq = maxn;
      }
    }
    if (live) {
      const double g12sq_norm = g12sq / cvacSq;
      const double b90 = P.b90_fact / (P.mu * g12sq_norm + 2.0 * EF_norm);
      double Clog = P.Clog;
      if (Clog == 0.0 && g12sq > 0.0) {
        const double bmax = LDe[c];
        const double bmin_qm = P.bqm_fact / (P.mu * sqrt(g12sq_norm));
        const double bmin = fmax(b90 / 2.0, bmin_qm);
        Clog = 0.5 * log(1.0 + bmax * bmax / bmin / bmin);
        Clog = fmax(2.0, Clog);
      }
      double sigma90 = 8.0 / k.PI * b90 * b90 * Clog;
      sigma90 = fmin(sigma90, sigma_max);
      nu = sqrt(g12sq) * nfreq * sigma90;
    }
  }
  max_bits(out, nu);
}

__global__ void k_nu_max_elastic(int ncell, const double *dens1, const double *ene1, const double *dens2,
                                 const double *ene2, double mass1, double mass2, ElaParams P,
                                 unsigned long long *out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  double nu = 0.0;
  if (c < ncell) {
    const double n1 = dens1[c], n2 = dens2[c];
    if (n1 * n2 != 0.0) {
      const size_t N = (size_t)ncell;
      double b1 = 0.0, b2 = 0.0;
      for (int d = 0; d < 3; ++d) {
        b1 += 2.0 * ene1[d * N + c];
        b2 += 2.0 * ene2[d * N + c];
      }
      b1 /= n1 * mass1;
      b2 /= n2 * mass2;
      const double g12 = sqrt(b1 + b2);
      double xi;
      const double sigma = elastic_sigma(P, g12, xi);
      nu = n2 * sigma * g12 * 2.99792458e+08;
    }
  }
  max_bits(out, nu);
}
}  // namespace pgpu

using namespace pgpu;

extern "C" {

int pgpu_ta_delta_u(long n, const double *vp1, const double *den1, const double *vp2, const double *den2,
                    double b90_fact, double Clog, double dt_sec, const double *gauss, const double *u_theta,
                    const double *u_phi, double *dU) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  cudaStream_t st = ctx().stream;
  double *d = nullptr;
  const size_t N = (size_t)n;
  // layout: vp1[3n] vp2[3n] den1[n] den2[n] gauss[n] uth[n] uphi[n] dU[3n]
  PGPU_CUDA(cudaMalloc(&d, 14 * N * sizeof(double)));
  double *d_v1 = d, *d_v2 = d + 3 * N, *d_d1 = d + 6 * N, *d_d2 = d + 7 * N, *d_g = d + 8 * N, *d_t = d + 9 * N,
         *d_p = d + 10 * N, *d_o = d + 11 * N;
  PGPU_CUDA(cudaMemcpyAsync(d_v1, vp1, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_v2, vp2, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_d1, den1, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_d2, den2, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_g, gauss, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_t, u_theta, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_p, u_phi, N * sizeof(double), cudaMemcpyHostToDevice, st));
  {
    KTimer t("ta_delta_u");
    k_ta_delta_u<<<nb(n), 256, 0, st>>>(n, d_v1, d_d1, d_v2, d_d2, b90_fact, Clog, dt_sec, d_g, d_t, d_p, d_o);
  }
  PGPU_CUDA(cudaMemcpyAsync(dU, d_o, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

int pgpu_ta_lorentz_scatter(long n, const double *up1, const double *up2, double mass1, double mass2,
                            const double *den2, double dt_sec, double b90_fact, double Clog, const double *gauss,
                            const double *u_theta, const double *u_phi, double *out1, double *out2) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  cudaStream_t st = ctx().stream;
  double *d = nullptr;
  const size_t N = (size_t)n;
  PGPU_CUDA(cudaMalloc(&d, 16 * N * sizeof(double)));
  double *d_1 = d, *d_2 = d + 3 * N, *d_d = d + 6 * N, *d_g = d + 7 * N, *d_t = d + 8 * N, *d_p = d + 9 * N,
         *d_o1 = d + 10 * N, *d_o2 = d + 13 * N;
  PGPU_CUDA(cudaMemcpyAsync(d_1, up1, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_2, up2, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_d, den2, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_g, gauss, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_t, u_theta, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d_p, u_phi, N * sizeof(double), cudaMemcpyHostToDevice, st));
  {
    KTimer t("ta_lorentz");
    k_ta_lorentz<<<nb(n), 256, 0, st>>>(n, d_1, d_2, mass1, mass2, d_d, dt_sec, b90_fact, Clog, d_g, d_t, d_p, d_o1,
                                        d_o2);
  }
  PGPU_CUDA(cudaMemcpyAsync(out1, d_o1, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(out2, d_o2, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

int pgpu_scatter_delta_u(long n, const double *u, const double *costh, const double *sinth, const double *cosphi,
                         const double *sinphi, double *dU) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  cudaStream_t st = ctx().stream;
  double *d = nullptr;
  const size_t N = (size_t)n;
  PGPU_CUDA(cudaMalloc(&d, 10 * N * sizeof(double)));   // u[3n] ct st cp sp dU[3n]
  PGPU_CUDA(cudaMemcpyAsync(d, u, 3 * N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d + 3 * N, costh, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d + 4 * N, sinth, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d + 5 * N, cosphi, N * sizeof(double), cudaMemcpyHostToDevice, st));
  PGPU_CUDA(cudaMemcpyAsync(d + 6 * N, sinphi, N * sizeof(double), cudaMemcpyHostToDevice, st));
  {
    KTimer t("scatter_delta_u");
    k_scatter_delta_u<<<nb(n), 256, 0, st>>>(n, d, d + 3 * N, d + 4 * N, d + 5 * N, d + 6 * N, d + 7 * N);
  }
  PGPU_CUDA(cudaMemcpyAsync(dU, d + 7 * N, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}


static int coulomb_consts(double charge1, double charge2, double mass1, double mass2, const pgpu_coulomb_params *prm,
                          double dt_sec, CoulParams *P) {
  if (!prm) return PGPU_ERR_ARG;
  const int a = prm->angular_scattering;
  if (a < PGPU_ANG_TAKIZUKA || a > PGPU_ANG_ISOTROPIC) {
    set_error("Coulomb: angular_scattering %d is not one of TAKIZUKA, NANBU, BOBYLEV, NANBU_FAS, NANBU_FAS_v2, ISOTROPIC", a);
    return PGPU_ERR_ARG;
  }
  if (charge1 == 0.0 || charge2 == 0.0) {
    set_error("Coulomb: neutral species");
    return PGPU_ERR_ARG;
  }
  const double PI = 3.14159265358979323846, CVAC = 2.99792458e+08, ME = 9.10938370e-31, QE = 1.60217663e-19;
  const double MU0 = 4.0 * PI * 1.0e-7, EP0 = 1.0 / CVAC / CVAC / MU0, HBAR = 6.62607015e-34 / (2.0 * PI);
  P->mu = mass1 * mass2 / (mass1 + mass2);
  const double qocSq = QE * QE / (CVAC * CVAC);
  P->b90_fact = fabs(charge1 * charge2) * qocSq / (2.0 * PI * EP0 * ME);     // Coulomb.cpp:47-50
  P->bqm_fact = HBAR / (2.0 * ME * CVAC);
  P->EF_fact = 0.0;
  if (mass1 == 1.0 || mass2 == 1.0)
    P->EF_fact = HBAR * HBAR / (2.0 * ME * P->mu) * pow(3.0 * PI * PI, 2.0 / 3.0) / (ME * CVAC * CVAC);
  P->f1 = P->mu / mass1;
  P->f2 = P->mu / mass2;
  P->mass1 = mass1;
  P->mass2 = mass2;
  P->rel = 0;
  if (prm->weight_method != 0 && prm->weight_method != 1) {
    set_error("Coulomb: weight_method must be 0 (PROBABILISTIC) or 1 (CONSERVATIVE, Sentoku-Kemp)");
    return PGPU_ERR_ARG;
  }
  P->sk08 = prm->weight_method;
  P->large_angle = prm->include_large_angle_scattering ? 1 : 0;
  P->large_draw = prm->test_large_angle_draw;
  P->fas_draw2 = prm->test_fas_draw2;
  P->fas_draw3 = prm->test_fas_draw3;
  P->Clog = prm->Clog;
  P->dt_sec = dt_sec;
  P->angular = a;
  P->NxN = prm->NxN ? 1 : 0;
  P->NxN_Nthresh = prm->NxN_Nthresh;
  P->cellV_SI = 1.0;
  P->seed_lo = P->seed_hi = P->step_lo = P->step_hi = 0;
  P->box_lo0 = P->box_lo1 = 0;
  P->nbox0 = P->ncell_glob0 = 1;
  return 0;
}

int pgpu_coulomb_delta_u(long n, const double *vp1, const double *vp2, double charge1, double charge2, double mass1,
                         double mass2, const pgpu_coulomb_params *prm, double dt_sec, const double *EF_norm,
                         const double *den12, const double *bmax, const double *sigma_max, const double *gauss,
                         const double *u_polar, const double *u_phi, double *dU, double *s12) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  CoulParams P;
  const int rc = coulomb_consts(charge1, charge2, mass1, mass2, prm, dt_sec, &P);
  if (rc) return rc;
  cudaStream_t st = ctx().stream;
  const size_t N = (size_t)n;
  double *d = nullptr;
  PGPU_CUDA(cudaMalloc(&d, 17 * N * sizeof(double)));   // v1[3] v2[3] EF den bmax smax g up uphi dU[3] s12
  const double *src[9] = {vp1, vp2, EF_norm, den12, bmax, sigma_max, gauss, u_polar, u_phi};
  const size_t len[9] = {3 * N, 3 * N, N, N, N, N, N, N, N};
  size_t off[10] = {0};
  for (int k = 0; k < 9; ++k) {
    PGPU_CUDA(cudaMemcpyAsync(d + off[k], src[k], len[k] * sizeof(double), cudaMemcpyHostToDevice, st));
    off[k + 1] = off[k] + len[k];
  }
  {
    KTimer t("coulomb_delta_u");
    k_coulomb_delta_u<<<nb(n), 256, 0, st>>>(n, P, d + off[0], d + off[1], d + off[2], d + off[3], d + off[4],
                                             d + off[5], d + off[6], d + off[7], d + off[8], d + off[9],
                                             d + off[9] + 3 * N);
  }
  PGPU_CUDA(cudaMemcpyAsync(dU, d + off[9], 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(s12, d + off[9] + 3 * N, N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  return 0;
}

int pgpu_coulomb_lorentz_scatter(long n, const double *up1, const double *up2, const int *scatter2, double charge1,
                                 double charge2, double mass1, double mass2, const pgpu_coulomb_params *prm,
                                 double dt_sec, const double *EF_norm, const double *den12, const double *bmax,
                                 const double *sigma_max, const double *gauss, const double *u_polar,
                                 const double *u_phi, double *out1, double *out2, double *s12) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  CoulParams P;
  const int rc = coulomb_consts(charge1, charge2, mass1, mass2, prm, dt_sec, &P);
  if (rc) return rc;
  P.rel = 1;
  cudaStream_t st = ctx().stream;
  const size_t N = (size_t)n;
  double *d = nullptr;
  int *ds = nullptr;
  PGPU_CUDA(cudaMalloc(&d, 20 * N * sizeof(double)));   // v1[3] v2[3] EF den bmax smax g up uphi | o1[3] o2[3] s12
  PGPU_CUDA(cudaMalloc(&ds, N * sizeof(int)));
  PGPU_CUDA(cudaMemcpyAsync(ds, scatter2, N * sizeof(int), cudaMemcpyHostToDevice, st));
  const double *src[9] = {up1, up2, EF_norm, den12, bmax, sigma_max, gauss, u_polar, u_phi};
  const size_t len[9] = {3 * N, 3 * N, N, N, N, N, N, N, N};
  size_t off[10] = {0};
  for (int k = 0; k < 9; ++k) {
    PGPU_CUDA(cudaMemcpyAsync(d + off[k], src[k], len[k] * sizeof(double), cudaMemcpyHostToDevice, st));
    off[k + 1] = off[k] + len[k];
  }
  {
    KTimer t("coulomb_lorentz");
    k_coulomb_lorentz<<<nb(n), 256, 0, st>>>(n, P, d + off[0], d + off[1], ds, d + off[2], d + off[3], d + off[4],
                                             d + off[5], d + off[6], d + off[7], d + off[8], d + off[9],
                                             d + off[9] + 3 * N, d + off[9] + 6 * N);
  }
  PGPU_CUDA(cudaMemcpyAsync(out1, d + off[9], 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(out2, d + off[9] + 3 * N, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(s12, d + off[9] + 6 * N, N * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  cudaFree(d);
  cudaFree(ds);
  return 0;
}

static int fetch_pairs(long *out) {
  Context &c = ctx();
  if (out) {
    PGPU_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c.stream));
    PGPU_CUDA(cudaMemsetAsync(&c.d_counters->npairs, 0, sizeof(unsigned long long), c.stream));
    PGPU_CUDA(cudaStreamSynchronize(c.stream));
    *out = (long)c.h_counters->npairs;
  } else {
    PGPU_CUDA(cudaMemsetAsync(&c.d_counters->npairs, 0, sizeof(unsigned long long), c.stream));
  }
  return 0;
}

// =============================================================================================
// scattering.coulomb.enforce_conservations (Coulomb.cpp:486-512, 596-714 intra; 1024-1083, 1182-1430 inter): the
// weight-rejection update of unequal-weight pairs conserves momentum and energy only on average; afterwards the cell's
// weighted momentum change is taken back out of every particle and the energy change is absorbed by zero-angle inelastic
// "collisions" of neighbouring particles of one list (ScatteringUtils::modEnergyPairwise, ScatteringUtils.H:113-205).
// One warp per cell: the sums are warp reductions; the pair sweeps are sequential in the reference (each pair sees what
// the previous one left of deltaE) and run on lane 0.
// =============================================================================================
struct EnfParams {
  double energy_fraction, energy_fraction_max;
  int wexp, nmin_save, rel;
  double mass1, mass2;
};
struct EnfList {
  const int *cs;
  double *v0, *v1, *v2;
  const double *w;
  double *save;      // [3][cap] copy of the velocities before the collisions (restored if the fix-up fails)
  long cap;
};
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double ipow(double w, int e) {
  double r = 1.0;
  for (int k = 0; k < e; ++k) r *= w;
  return r;
}
// W = sum w^exp, p = sum w v, E = sum w Efact |v|^2 (Efact = 1/2, or 1/(gamma+1) in the relativistic build), wsum = sum w
__device__ __forceinline__ void cell_sums_dev(const EnfList &L, int s, int n, const EnfParams &P, int lane, double &W,
                                              double (&p)[3], double &E, double &wsum) {
  W = E = wsum = 0.0;
  p[0] = p[1] = p[2] = 0.0;
  for (int q = lane; q < n; q += 32) {
    const int i = s + q;
    const double w = L.w[i], a = L.v0[i], b = L.v1[i], c = L.v2[i];
    const double gbsq = a * a + b * b + c * c;
    const double Efact = P.rel ? 1.0 / (sqrt(1.0 + gbsq) + 1.0) : 0.5;
    W += ipow(w, P.wexp);
    wsum += w;
    p[0] += w * a;
    p[1] += w * b;
    p[2] += w * c;
    E += w * Efact * gbsq;
  }
  W = warp_sum(W);
  wsum = warp_sum(wsum);
  E = warp_sum(E);
#pragma unroll
  for (int k = 0; k < 3; ++k) p[k] = warp_sum(p[k]);
}
// before the collisions: sums of the cell list -> buf[6] = W p0 p1 p2 E wsum, and the velocities -> save
__global__ void __launch_bounds__(256) k_coul_enf_pre(EnfList L, int ncell, EnfParams P, int which, double *buf) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const int s = L.cs[cell], n = L.cs[cell + 1] - s;
  double W, p[3], E, wsum;
  cell_sums_dev(L, s, n, P, lane, W, p, E, wsum);
  if (lane == 0) {
    double *b = buf + ((size_t)cell * 2 + which) * 6;
    b[0] = W, b[1] = p[0], b[2] = p[1], b[3] = p[2], b[4] = E, b[5] = wsum;
  }
  for (int q = lane; q < n; q += 32) {
    const int i = s + q;
    L.save[i] = L.v0[i];
    L.save[L.cap + i] = L.v1[i];
    L.save[2 * L.cap + i] = L.v2[i];
  }
}
// ScatteringUtils::modEnergyPairwise in fp64 (the reference keeps the scalars in long double)
__device__ void mod_energy_pairwise_dev(double *b1, double *b2, double wpmp1, double wpmp2, double Erel_frac,
                                        double &Erel_cumm, double &a_deltaE, int rel) {
  const double sign = a_deltaE < 0.0 ? -1.0 : 1.0;
  const double ux = b1[0] - b2[0], uy = b1[1] - b2[1], uz = b1[2] - b2[2];
  const double usq = ux * ux + uy * uy + uz * uz;
  double Erel, muR = 0.0, E1 = 0.0, E2 = 0.0, Etot = 0.0, px = 0.0, py = 0.0, pz = 0.0;
  if (rel) {
    const double g1 = sqrt(1.0 + b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    const double g2 = sqrt(1.0 + b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
    E1 = wpmp1 * g1;
    E2 = wpmp2 * g2;
    Etot = E1 + E2;
    px = wpmp1 * b1[0] + wpmp2 * b2[0];
    py = wpmp1 * b1[1] + wpmp2 * b2[1];
    pz = wpmp1 * b1[2] + wpmp2 * b2[2];
    Erel = sqrt(Etot * Etot - px * px - py * py - pz * pz) - wpmp1 - wpmp2;
  } else {
    muR = wpmp1 * wpmp2 / (wpmp1 + wpmp2);
    Erel = muR / 2.0 * usq;
  }
  if (!(Erel > 0.0)) return;
  double deltaE = sign * Erel_frac * Erel;
  if (fabs(deltaE) > fabs(a_deltaE)) {
    deltaE = a_deltaE;
    a_deltaE = 0.0;
  } else {
    a_deltaE -= deltaE;
  }
  Erel_cumm += Erel - deltaE;
  if (rel) {
    const double A = Etot - deltaE, D = A * A + E2 * E2 - E1 * E1;
    const double p2dotu = wpmp2 * (b2[0] * ux + b2[1] * uy + b2[2] * uz), ptdotu = px * ux + py * uy + pz * uz;
    const double a = A * A * usq - ptdotu * ptdotu, b = D * ptdotu - 2.0 * A * A * p2dotu, c = A * A * E2 * E2 - D * D / 4.0;
    const double root = b * b - 4.0 * a * c;
    if (root < 0.0 || a == 0.0) return;
    const double alpha = (-b + sqrt(root)) / (2.0 * a), r1 = alpha / wpmp1, r2 = alpha / wpmp2;
    b1[0] += r1 * ux, b1[1] += r1 * uy, b1[2] += r1 * uz;
    b2[0] -= r2 * ux, b2[1] -= r2 * uy, b2[2] -= r2 * uz;
  } else {
    const double k = sqrt(1.0 - deltaE / Erel) - 1.0;   // u'/u - 1
    const double f1 = muR / wpmp1 * k, f2 = muR / wpmp2 * k;
    b1[0] += f1 * ux, b1[1] += f1 * uy, b1[2] += f1 * uz;
    b2[0] -= f2 * ux, b2[1] -= f2 * uy, b2[2] -= f2 * uz;
  }
}
// the pair sweeps over one list in storage order (Coulomb.cpp:643-706 / 1257-1341); false = the correction failed
__device__ bool absorb_energy_dev(const EnfList &L, int s, int N, double mass, const EnfParams &P, double &deltaE) {
  int loop_count = 0;
  double Erel_cumm = 0.0, fmult = 1.0;
  for (int p = 0; p < N; p++) {
    if (deltaE == 0.0) break;
    const int p1 = p;
    p++;
    if (p == N) {
      loop_count++;
      p = 0;
    }
    const int i1 = s + p1, i2 = s + p;
    double a[3] = {L.v0[i1], L.v1[i1], L.v2[i1]}, b[3] = {L.v0[i2], L.v1[i2], L.v2[i2]};
    mod_energy_pairwise_dev(a, b, mass * L.w[i1], mass * L.w[i2], P.energy_fraction * fmult, Erel_cumm, deltaE, P.rel);
    L.v0[i1] = a[0], L.v1[i1] = a[1], L.v2[i1] = a[2];
    L.v0[i2] = b[0], L.v1[i2] = b[1], L.v2[i2] = b[2];
    if (deltaE == 0.0) break;
    if (p == N - 1) {
      loop_count++;
      const double eff = fabs(deltaE) / Erel_cumm;
      if (eff > P.energy_fraction_max || loop_count > 10) return false;
      if (eff > P.energy_fraction) fmult = eff / P.energy_fraction;
      Erel_cumm = 0.0;
      p = -1;
    }
  }
  return true;
}
// after the collisions.  inter != 0: two lists (L1 with mass1, L2 with mass2); else L1 only
__global__ void __launch_bounds__(256)
k_coul_enf_post(EnfList L1, EnfList L2, int inter, int ncell, EnfParams P, const double *buf, unsigned *nfailed) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const int s1 = L1.cs[cell], n1 = L1.cs[cell + 1] - s1;
  const int s2 = inter ? L2.cs[cell] : 0, n2 = inter ? L2.cs[cell + 1] - s2 : 0;
  if (inter ? (n1 < 1 || n2 < 1 || (long)n1 * n2 < 2) : (n1 < 2)) return;   // cells the collision kernels skip
  const double *b1 = buf + ((size_t)cell * 2 + 0) * 6, *b2 = buf + ((size_t)cell * 2 + 1) * 6;
  double W, p[3], E, wsum, W2 = 0.0, p2[3] = {0.0, 0.0, 0.0}, E2 = 0.0, wsum2 = 0.0;
  cell_sums_dev(L1, s1, n1, P, lane, W, p, E, wsum);
  if (inter) cell_sums_dev(L2, s2, n2, P, lane, W2, p2, E2, wsum2);
  const double m1 = inter ? P.mass1 : 1.0, m2 = P.mass2;
  const double Wtot0 = inter ? m1 * b1[0] + m2 * b2[0] : b1[0];
  double dB[3], dBsq = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dB[k] = inter ? (m1 * p[k] + m2 * p2[k]) - (m1 * b1[1 + k] + m2 * b2[1 + k]) : p[k] - b1[1 + k];
    dBsq += dB[k] * dB[k];
  }
  if (!(dBsq > 0.0)) return;
#pragma unroll
  for (int k = 0; k < 3; ++k) dB[k] /= Wtot0;
  // take the momentum change back out, and sum the energies of the shifted velocities
  double Et1 = 0.0, Et2 = 0.0;
  for (int q = lane; q < n1; q += 32) {
    const int i = s1 + q;
    const double f = ipow(L1.w[i], P.wexp - 1);
    const double a = L1.v0[i] - f * dB[0], b = L1.v1[i] - f * dB[1], c = L1.v2[i] - f * dB[2];
    L1.v0[i] = a, L1.v1[i] = b, L1.v2[i] = c;
    const double gbsq = a * a + b * b + c * c;
    Et1 += L1.w[i] * (P.rel ? 1.0 / (sqrt(1.0 + gbsq) + 1.0) : 0.5) * gbsq;
  }
  for (int q = lane; q < n2; q += 32) {
    const int i = s2 + q;
    const double f = ipow(L2.w[i], P.wexp - 1);
    const double a = L2.v0[i] - f * dB[0], b = L2.v1[i] - f * dB[1], c = L2.v2[i] - f * dB[2];
    L2.v0[i] = a, L2.v1[i] = b, L2.v2[i] = c;
    const double gbsq = a * a + b * b + c * c;
    Et2 += L2.w[i] * (P.rel ? 1.0 / (sqrt(1.0 + gbsq) + 1.0) : 0.5) * gbsq;
  }
  Et1 = warp_sum(Et1);
  Et2 = warp_sum(Et2);
  __syncwarp();
  int ok = 1;
  if (lane == 0) {
    if (!inter) {
      double deltaE = P.mass1 * (Et1 - b1[4]);
      ok = absorb_energy_dev(L1, s1, n1, P.mass1, P, deltaE) ? 1 : 0;
    } else {
      Et1 *= m1;
      Et2 *= m2;
      const double deltaE = (Et1 + Et2) - (m1 * b1[4] + m2 * b2[4]);
      const double wm1 = b1[5] / n1, wm2 = b2[5] / n2, den = wm1 * Et1 + wm2 * Et2;
      double d1, d2;
      if (n1 == 1) d1 = 0.0, d2 = deltaE;
      else if (n2 == 1) d1 = deltaE, d2 = 0.0;
      else d1 = wm1 * Et1 / den * deltaE, d2 = wm2 * Et2 / den * deltaE;
      ok = absorb_energy_dev(L1, s1, n1, m1, P, d1) ? 1 : 0;
      if (ok) ok = absorb_energy_dev(L2, s2, n2, m2, P, d2) ? 1 : 0;
    }
    if (!ok) atomicAdd(nfailed, 1u);
  }
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if (!ok && (n1 <= P.nmin_save || (inter && n2 <= P.nmin_save))) {   // Coulomb.cpp:689-696 / 1409-1426
    for (int q = lane; q < n1; q += 32) {
      const int i = s1 + q;
      L1.v0[i] = L1.save[i], L1.v1[i] = L1.save[L1.cap + i], L1.v2[i] = L1.save[2 * L1.cap + i];
    }
    for (int q = lane; q < n2; q += 32) {
      const int i = s2 + q;
      L2.v0[i] = L2.save[i], L2.v1[i] = L2.save[L2.cap + i], L2.v2[i] = L2.save[2 * L2.cap + i];
    }
  }
}

// Every scattering operator draws from its own Philox stream: the caller's seed is mixed with the identity of the two
// species (mass and charge: stable when a species object is re-created, e.g. at a restart) and a per-model salt
// (splitmix64 finaliser), so that e-e and i-i self-scattering, or e-i1 and e-i2, of one step do not replay each other's
// draws when a driver hands every operator the same (seed, step).  Two species with identical mass and charge are told
// apart by the caller's seed (the C++ shim gives every operator object its own).
static uint64_t species_tag(const pgpu_species_s *s) {
  if (!s) return 0;
  uint64_t a, b;
  memcpy(&a, &s->desc.mass, 8);
  memcpy(&b, &s->desc.charge, 8);
  return a * 0x9e3779b97f4a7c15ull ^ (b + 0x7f4a7c15ull) * 0xc2b2ae3d27d4eb4full;
}
static uint64_t stream_seed(uint64_t seed, const pgpu_species_s *sA, const pgpu_species_s *sB, uint64_t salt) {
  uint64_t z = seed ^ species_tag(sA) ^ (species_tag(sB) << 1 | species_tag(sB) >> 63) ^ (salt << 8);
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

static int need_binned(pgpu_species_t sA, pgpu_species_t sB) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  if (!sA || !sB || sA->grid != sB->grid) return PGPU_ERR_ARG;
  if (!sA->binned || !sB->binned) {
    set_error("collisions need binned species: call pgpu_bin_particles + pgpu_set_moments_from_bins first");
    return PGPU_ERR_STATE;
  }
  // collisions write v: a deferred vold = v copy has to happen first (x is not touched: xold may stay aliased)
  if (materialize_old(sA, KEEP_XOLD_ALIAS | KEEP_PENDING) || materialize_old(sB, KEEP_XOLD_ALIAS | KEEP_PENDING)) return PGPU_ERR_CUDA;
  return 0;
}

int pgpu_collide_coulomb(pgpu_species_t sA, pgpu_species_t sB, const pgpu_coulomb_params *prm, double dt_sec,
                         uint64_t seed, uint64_t step, long *npairs_out) {
  if (!sA) return PGPU_ERR_ARG;
  int rc = need_binned(sA, sB);
  if (rc) return rc;
  Context &c = ctx();
  const pgpu_grid_s *g = sA->grid;
  CoulParams P;
  rc = coulomb_consts(sA->desc.charge, sB->desc.charge, sA->desc.mass, sB->desc.mass, prm, dt_sec, &P);
  if (rc) return rc;
  const int nsub = prm->num_subcycles > 0 ? prm->num_subcycles : 1;
  P.dt_sec = dt_sec / (double)nsub;                                   // Coulomb.cpp:371
  P.rel = (sA->desc.relativistic || sB->desc.relativistic) ? 1 : 0;   // the reference's compile-time switch
  if (P.sk08 && (P.rel || prm->enforce_conservations)) {
    set_error("Coulomb: weight_method CONSERVATIVE is the Galilean SK08 update (Coulomb.cpp:730-917, 1439-1640); it has no "
              "relativistic form and is not combined with enforce_conservations");
    return PGPU_ERR_ARG;
  }
  if (P.sk08) P.NxN = 0, P.NxN_Nthresh = 0;   // applyIntra/InterScattering_SK08 pair in O(N) only
  const double dV = (g->desc.D == 1) ? g->geo.dx[0] : g->geo.dx[0] * g->geo.dx[1];
  P.cellV_SI = dV * g->desc.volume_scale;
  seed = stream_seed(seed, sA, sB, 1);
  P.seed_lo = (unsigned)seed;
  P.seed_hi = (unsigned)(seed >> 32);
  P.step_lo = (unsigned)step;
  P.box_lo0 = g->desc.box_lo[0];
  P.box_lo1 = (g->desc.D == 2) ? g->desc.box_lo[1] : 0;
  P.nbox0 = g->nbox[0];
  P.ncell_glob0 = g->desc.ncell[0];
  const int ncell = (int)g->ncell_box;
  unsigned long long *d_np = &c.d_counters->npairs;
  // scattering.coulomb.enforce_conservations (Coulomb.H:286-293)
  const bool enforce = prm->enforce_conservations != 0;
  EnfParams EP;
  EnfList L1, L2;
  static double *enf_buf = nullptr;
  static size_t enf_buf_cells = 0;
  static unsigned *enf_failed = nullptr;
  if (enforce) {
    if (prm->sort_weighted_particles) {
      set_error("Coulomb: sort_weighted_particles is not implemented (the energy fix-up pairs particles in storage order)");
      return PGPU_ERR_ARG;
    }
    if (!(prm->energy_fraction > 0.0) || !(prm->energy_fraction_max > 0.0) || prm->beta_weight_exponent < 1) {
      set_error("Coulomb: enforce_conservations needs energy_fraction > 0, energy_fraction_max > 0, beta_weight_exponent >= 1");
      return PGPU_ERR_ARG;
    }
    EP.energy_fraction = prm->energy_fraction;
    EP.energy_fraction_max = prm->energy_fraction_max;
    EP.wexp = prm->beta_weight_exponent;
    EP.nmin_save = prm->conservation_Nmin_save > 0 ? prm->conservation_Nmin_save : 100000;
    EP.rel = P.rel;
    EP.mass1 = sA->desc.mass;
    EP.mass2 = sB->desc.mass;
    if ((size_t)ncell > enf_buf_cells) {
      if (enf_buf) cudaFree(enf_buf);
      PGPU_CUDA(cudaMalloc(&enf_buf, (size_t)ncell * 12 * sizeof(double)));
      enf_buf_cells = (size_t)ncell;
    }
    if (!enf_failed) {
      PGPU_CUDA(cudaMalloc(&enf_failed, sizeof(unsigned)));
      PGPU_CUDA(cudaMemsetAsync(enf_failed, 0, sizeof(unsigned), c.stream));
    }
    auto mk = [&](pgpu_species_s *sp, EnfList &L) -> int {
      if (!sp->enf_save || sp->enf_cap < sp->cap) {
        if (sp->enf_save) cudaFree(sp->enf_save);
        if (cudaMalloc(&sp->enf_save, 3 * sp->cap * sizeof(double)) != cudaSuccess) return PGPU_ERR_CUDA;
        sp->enf_cap = sp->cap;
      }
      L.cs = sp->cell_start;
      L.v0 = sp->v[0], L.v1 = sp->v[1], L.v2 = sp->v[2];
      L.w = sp->w;
      L.save = sp->enf_save;
      L.cap = (long)sp->enf_cap;
      return 0;
    };
    if (mk(sA, L1) || mk(sB, L2)) return PGPU_ERR_CUDA;
  }
  // the full-angle models live in their own instantiation: their solver calls cost the others registers
  const bool fas = P.angular == ANG_NANBU_FAS || P.angular == ANG_NANBU_FAS_V2;
  // the plain case (Galilean, PROBABILISTIC weights, no large-angle events) has a leaner instantiation in smaller blocks
  const bool lean = !fas && !P.rel && !P.sk08 && !P.large_angle;
  const int tpb = lean ? 128 : 256;
  for (int sub = 0; sub < nsub; ++sub) {
    P.step_hi = ((unsigned)(step >> 32) & 0xffu) | ((unsigned)sub << 8);
    if (enforce) {
      KTimer t("collide_coulomb_enforce");
      k_coul_enf_pre<<<nb((long)ncell * 32), 256, 0, c.stream>>>(L1, ncell, EP, 0, enf_buf);
      if (sA != sB) k_coul_enf_pre<<<nb((long)ncell * 32), 256, 0, c.stream>>>(L2, ncell, EP, 1, enf_buf);
    }
    if (sA == sB) {
      KTimer t("collide_coulomb_intra");
      auto k = fas ? k_coulomb_intra<true, false> : (lean ? k_coulomb_intra<false, true> : k_coulomb_intra<false, false>);
      k<<<nb((long)ncell * 32, tpb), tpb, 0, c.stream>>>(sA->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->w, sA->id,
                                                    sA->dens, g->debye, P, (unsigned *)sA->cell_key, sA->perm, d_np);
    } else {
      KTimer t("collide_coulomb_inter");
      auto k = fas ? k_coulomb_inter<true, false> : (lean ? k_coulomb_inter<false, true> : k_coulomb_inter<false, false>);
      k<<<nb((long)ncell * 32, tpb), tpb, 0, c.stream>>>(
          sA->cell_start, sB->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->w, sA->id, sA->dens, sB->v[0],
          sB->v[1], sB->v[2], sB->w, sB->id, sB->dens, g->debye, P, (unsigned *)sA->cell_key, sA->perm,
          (unsigned *)sB->cell_key, sB->perm, d_np);
    }
    if (enforce) {
      KTimer t("collide_coulomb_enforce");
      k_coul_enf_post<<<nb((long)ncell * 32), 256, 0, c.stream>>>(L1, L2, sA != sB ? 1 : 0, ncell, EP, enf_buf, enf_failed);
    }
  }
  return fetch_pairs(npairs_out);
}

int pgpu_collide_elastic(pgpu_species_t sA, pgpu_species_t sB, const pgpu_elastic_params *prm, double dt_sec,
                         uint64_t seed, uint64_t step, long *ncoll_out) {
  if (!sA) return PGPU_ERR_ARG;
  int rc = need_binned(sA, sB);
  if (rc) return rc;
  if (!prm || sA == sB || (prm->ntab && (!prm->E || !prm->Q || prm->ntab < 2))) return PGPU_ERR_ARG;
  if (prm->ntab && prm->angular_scattering == 1 && !prm->xi) return PGPU_ERR_ARG;
  Context &c = ctx();
  const pgpu_grid_s *g = sA->grid;
  const double CVAC = 2.99792458e+08, ME = 9.10938370e-31, QE = 1.60217663e-19;
  ElaParams P;
  const double m1 = sA->desc.mass, m2 = sB->desc.mass;
  P.mu = m1 * m2 / (m1 + m2);
  P.f1 = P.mu / m1;
  P.f2 = P.mu / m2;
  P.const_sigma = prm->const_sigma;
  P.dt_sec = dt_sec;
  P.mcSq = ME * CVAC * CVAC / QE;
  P.ntab = prm->ntab;
  P.angular = prm->angular_scattering;
  P.loglog = prm->use_loglog_interp;
  if (prm->weight_method != 0 && prm->weight_method != 1) return PGPU_ERR_ARG;
  P.conservative = prm->weight_method;
  P.wmut = sB->w;
  P.E = P.Q = P.XI = nullptr;
  if (prm->ntab) {
    // the cross-section table lives on the device between calls (no cudaMalloc / cudaFree in the middle of a step):
    // re-uploaded only when its contents change
    static double *d_tab = nullptr;
    static size_t d_cap = 0;
    static std::vector<double> h_tab;
    const size_t N = (size_t)prm->ntab;
    std::vector<double> cur(3 * N, 0.0);
    memcpy(cur.data(), prm->E, N * sizeof(double));
    memcpy(cur.data() + N, prm->Q, N * sizeof(double));
    if (prm->xi) memcpy(cur.data() + 2 * N, prm->xi, N * sizeof(double));
    if (cur != h_tab) {
      if (3 * N > d_cap) {
        if (d_tab) {
          cudaStreamSynchronize(c.stream);
          cudaFree(d_tab);
        }
        PGPU_CUDA(cudaMalloc(&d_tab, 3 * N * sizeof(double)));
        d_cap = 3 * N;
      }
      PGPU_CUDA(cudaStreamSynchronize(c.stream));   // an earlier launch may still read the old table
      PGPU_CUDA(cudaMemcpy(d_tab, cur.data(), 3 * N * sizeof(double), cudaMemcpyHostToDevice));
      h_tab.swap(cur);
    }
    P.E = d_tab;
    P.Q = d_tab + N;
    P.XI = d_tab + 2 * N;
  }
  seed = stream_seed(seed, sA, sB, 3);
  P.seed_lo = (unsigned)seed;
  P.seed_hi = (unsigned)(seed >> 32);
  P.step_lo = (unsigned)step;
  P.step_hi = (unsigned)(step >> 32) & 0xffffu;
  const int ncell = (int)g->ncell_box;
  {
    KTimer t("collide_elastic");
    k_elastic<<<nb((long)ncell * 32), 256, 0, c.stream>>>(sA->cell_start, sB->cell_start, ncell, sA->v[0], sA->v[1],
                                                          sA->v[2], sA->w, sA->id, sB->v[0], sB->v[1], sB->v[2], sB->w,
                                                          sB->dens, P, &c.d_counters->npairs);
  }
  return fetch_pairs(ncoll_out);
}

static int launch_hard_sphere(pgpu_species_t sA, pgpu_species_t sB, HSParams P, double dt_sec, uint64_t seed,
                              uint64_t step, long *ncoll_out, const char *who) {
  int rc = need_binned(sA, sB);
  if (rc) return rc;
  if (!sA->dens || !sA->ene || !sB->dens || !sB->ene) {
    set_error("%s needs the cell moments: call pgpu_set_moments_from_bins first", who);
    return PGPU_ERR_STATE;
  }
  Context &c = ctx();
  const pgpu_grid_s *g = sA->grid;
  P.dt_sec = dt_sec;
  P.mass1 = sA->desc.mass;
  P.mass2 = sB->desc.mass;
  P.mu = P.mass1 * P.mass2 / (P.mass1 + P.mass2);
  const double dV = (g->desc.D == 1) ? g->geo.dx[0] : g->geo.dx[0] * g->geo.dx[1];
  P.Vc = dV * g->desc.volume_scale;      // DomainGrid::getMappedCellVolume in SI
  seed = stream_seed(seed, sA, sB, 5);
  P.seed_lo = (unsigned)seed;
  P.seed_hi = (unsigned)(seed >> 32);
  P.step_lo = (unsigned)step;
  P.step_hi = (unsigned)(step >> 32) & 0xffffu;
  P.box_lo0 = g->desc.box_lo[0];
  P.box_lo1 = (g->desc.D == 2) ? g->desc.box_lo[1] : 0;
  P.nbox0 = g->nbox[0];
  P.ncell_glob0 = g->desc.ncell[0];
  const int ncell = (int)g->ncell_box;
  {
    KTimer t("collide_hard_sphere");
    k_hard_sphere<<<(unsigned)((ncell + 127) / 128), 128, 0, c.stream>>>(
        sA->cell_start, sB->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->w, sA->dens, sA->ene, sB->v[0], sB->v[1],
        sB->v[2], sB->w, sB->dens, sB->ene, P, sA == sB ? 1 : 0, &c.d_counters->npairs);
  }
  return fetch_pairs(ncoll_out);
}

int pgpu_collide_hard_sphere(pgpu_species_t sA, pgpu_species_t sB, double sigmaT, double dt_sec, uint64_t seed,
                             uint64_t step, long *ncoll_out) {
  if (!sA) return PGPU_ERR_ARG;
  if (!(sigmaT > 0.0)) {
    set_error("HardSphere: sigmaT must be positive");
    return PGPU_ERR_ARG;
  }
  HSParams P;
  memset(&P, 0, sizeof(P));
  P.sigmaT = sigmaT;
  return launch_hard_sphere(sA, sB, P, dt_sec, seed, step, ncoll_out, "HardSphere");
}

// weight_method: 0 = PROBABILISTIC, 1 = CONSERVATIVE (self and inter-species; the species' weights change)
int pgpu_collide_hard_sphere_wm(pgpu_species_t sA, pgpu_species_t sB, double sigmaT, int weight_method, double dt_sec,
                                uint64_t seed, uint64_t step, long *ncoll_out) {
  if (weight_method == 0) return pgpu_collide_hard_sphere(sA, sB, sigmaT, dt_sec, seed, step, ncoll_out);
  if (!sA || !sB) return PGPU_ERR_ARG;
  if (weight_method != 1) {
    set_error("HardSphere: weight_method must be 0 (PROBABILISTIC) or 1 (CONSERVATIVE)");
    return PGPU_ERR_ARG;
  }
  if (!(sigmaT > 0.0)) {
    set_error("HardSphere: sigmaT must be positive");
    return PGPU_ERR_ARG;
  }
  HSParams P;
  memset(&P, 0, sizeof(P));
  P.sigmaT = sigmaT;
  P.conservative = 1;
  P.wmut = sA->w;
  P.wmut2 = sB->w;
  return launch_hard_sphere(sA, sB, P, dt_sec, seed, step, ncoll_out, "HardSphere");
}

// VariableHardSphere::initialize (VariableHardSphere.cpp:28-47): alpha = 4/(2 eta - 1), A from the reference
// viscosity mu0 at T0
static void vhs_consts(double mass, double eta, double T0, double mu0, double *fourPiA, double *fourOverAlpha) {
  const double PI = 3.14159265358979323846, ME = 9.10938370e-31, KB = 1.380649e-23;
  const double alpha = 4. / (2. * eta - 1.);
  const double Mass_kg = mass * ME;
  const double VT0 = sqrt(KB * T0 / Mass_kg);
  const double Gamma0 = tgamma(4. - 2. / alpha);
  const double Aconst = 15. / 32. / Gamma0 / mu0 * Mass_kg / sqrt(PI) * VT0 * pow(4. * VT0 * VT0, 2. / alpha);
  *fourPiA = 4. * PI * Aconst;
  *fourOverAlpha = 4.0 / alpha;
}

int pgpu_collide_vhs(pgpu_species_t s, double eta, double T0, double mu0, double dt_sec, uint64_t seed, uint64_t step,
                     long *ncoll_out) {
  if (!s) return PGPU_ERR_ARG;
  if (!(eta > 0.5) || !(T0 > 0.0) || !(mu0 > 0.0)) {
    set_error("VariableHardSphere: need eta > 1/2, T0 > 0, mu0 > 0");
    return PGPU_ERR_ARG;
  }
  HSParams P;
  memset(&P, 0, sizeof(P));
  P.vhs = 1;
  P.sigmaT = 1.0;
  if (!s) return PGPU_ERR_ARG;
  vhs_consts(s->desc.mass, eta, T0, mu0, &P.fourPiA, &P.fourOverAlpha);
  return launch_hard_sphere(s, s, P, dt_sec, seed, step, ncoll_out, "VariableHardSphere");
}

// HardSphere::setIntraMFT / setInterMFT (HardSphere.cpp:84-194): nu_max = max over cells of n * sigmaT * sqrt(Teff/m).
// The moments are ncell-sized: read back and reduced on the host (once per step, like the reference's loop).
int pgpu_scatter_nu_max_hard_sphere(pgpu_species_t sA, pgpu_species_t sB, double sigmaT, double *nu_max) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  if (!sA || !sB || sA->grid != sB->grid || !nu_max) return PGPU_ERR_ARG;
  if (!sA->dens || !sA->ene || !sB->dens || !sB->ene) {
    set_error("HardSphere::setMeanFreeTime needs the cell moments: call pgpu_set_moments_from_bins first");
    return PGPU_ERR_STATE;
  }
  const size_t ncell = (size_t)sA->grid->ncell_box;
  std::vector<double> d1(ncell), e1(3 * ncell), d2(ncell), e2(3 * ncell);
  cudaStream_t st = ctx().stream;
  PGPU_CUDA(cudaMemcpyAsync(d1.data(), sA->dens, ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(e1.data(), sA->ene, 3 * ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(d2.data(), sB->dens, ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(e2.data(), sB->ene, 3 * ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  const double CVAC = 2.99792458e+08, cvacSq = CVAC * CVAC;
  const double m1 = sA->desc.mass, m2 = sB->desc.mass;
  double box_nuMax = 0.0;
  for (size_t c = 0; c < ncell; ++c) {
    if (sA == sB) {
      if (d1[c] == 0.0) continue;
      const double e = e1[c] + e1[ncell + c] + e1[2 * ncell + c];
      const double Teff = 2.0 / 3.0 * e / d1[c] * cvacSq;
      box_nuMax = std::max(box_nuMax, d1[c] * sigmaT * sqrt(Teff / m1));
    } else {
      if (d1[c] * d2[c] == 0.0) continue;
      const double ea = e1[c] + e1[ncell + c] + e1[2 * ncell + c], eb = e2[c] + e2[ncell + c] + e2[2 * ncell + c];
      const double Teff1 = 2.0 / 3.0 * ea / d1[c] * cvacSq, Teff2 = 2.0 / 3.0 * eb / d2[c] * cvacSq;
      box_nuMax = std::max(box_nuMax, d2[c] * sigmaT * sqrt(Teff1 / m1));
      box_nuMax = std::max(box_nuMax, d1[c] * sigmaT * sqrt(Teff2 / m2));
    }
  }
  *nu_max = box_nuMax;
  return 0;
}

// VariableHardSphere::setIntraMFT (VariableHardSphere.cpp:80-125): box maximum of n * sigmaT(VTeff) * VTeff
int pgpu_scatter_nu_max_vhs(pgpu_species_t s, double eta, double T0, double mu0, double *nu_max) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  if (!s || !nu_max || !(eta > 0.5) || !(T0 > 0.0) || !(mu0 > 0.0)) return PGPU_ERR_ARG;
  if (!s->dens || !s->ene) {
    set_error("VariableHardSphere::setMeanFreeTime needs the cell moments: call pgpu_set_moments_from_bins first");
    return PGPU_ERR_STATE;
  }
  double fourPiA, fourOverAlpha;
  vhs_consts(s->desc.mass, eta, T0, mu0, &fourPiA, &fourOverAlpha);
  const size_t ncell = (size_t)s->grid->ncell_box;
  std::vector<double> d(ncell), e(3 * ncell);
  cudaStream_t st = ctx().stream;
  PGPU_CUDA(cudaMemcpyAsync(d.data(), s->dens, ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaMemcpyAsync(e.data(), s->ene, 3 * ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
  PGPU_CUDA(cudaStreamSynchronize(st));
  const double CVAC = 2.99792458e+08, cvacSq = CVAC * CVAC;
  double box_nuMax = 0.0;
  for (size_t c = 0; c < ncell; ++c) {
    if (d[c] == 0.0) continue;
    const double en = e[c] + e[ncell + c] + e[2 * ncell + c];
    const double Teff = 2.0 / 3.0 * en / d[c] * cvacSq;
    const double VTeff = sqrt(Teff / s->desc.mass);
    const double sigmaTmax = fourPiA * pow(VTeff, -fourOverAlpha);
    box_nuMax = std::max(box_nuMax, d[c] * sigmaTmax * VTeff);
  }
  *nu_max = box_nuMax;
  return 0;
}

int pgpu_collide_ta(pgpu_species_t sA, pgpu_species_t sB, double Clog, double dt_sec, uint64_t seed,
                    uint64_t step, long *npairs_out) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  if (!sA || !sB || sA->grid != sB->grid) return PGPU_ERR_ARG;
  if (!sA->binned || !sB->binned) {
    set_error("collisions need binned species: call pgpu_bin_particles + pgpu_set_moments_from_bins first");
    return PGPU_ERR_STATE;
  }
  if (materialize_old(sA, KEEP_XOLD_ALIAS | KEEP_PENDING) || materialize_old(sB, KEEP_XOLD_ALIAS | KEEP_PENDING)) return PGPU_ERR_CUDA;
  Context &c = ctx();
  const pgpu_grid_s *g = sA->grid;
  // m_mu, m_b90_fact (TakizukaAbe.cpp:39-49); long double as in TakizukaAbe.H:137-139
  const long double m1 = sA->desc.mass, m2 = sB->desc.mass;
  const long double mu = m1 * m2 / (m1 + m2);
  const double PI = 3.14159265358979323846, CVAC = 2.99792458e+08, ME = 9.10938370e-31, QE = 1.60217663e-19;
  const double MU0 = 4.0 * PI * 1.0e-7, EP0 = 1.0 / CVAC / CVAC / MU0;
  const double b90_codeToPhys = QE * QE / (4.0 * PI * EP0 * ME);
  const int q1 = (int)sA->desc.charge, q2 = (int)sB->desc.charge;
  TAParams P;
  P.b90_fact = (double)(abs(q1 * q2) / (mu * (long double)(CVAC * CVAC)) * b90_codeToPhys);
  P.rel = (sA->desc.relativistic || sB->desc.relativistic) ? 1 : 0;
  P.m1 = sA->desc.mass;
  P.m2 = sB->desc.mass;
  // RELATIVISTIC_PARTICLES build: m_b90_codeToPhys has 2 pi (TakizukaAbe.H:27-28), m_b90_fact no 1/mu (.cpp:45-46)
  if (P.rel) P.b90_fact = abs(q1 * q2) / (CVAC * CVAC) * (QE * QE / (2.0 * PI * EP0 * ME));
  P.Clog = Clog;
  P.dt_sec = dt_sec;
  P.f1 = (double)(mu / m1);
  P.f2 = (double)(mu / m2);
  seed = stream_seed(seed, sA, sB, 7);
  P.seed_lo = (unsigned)seed;
  P.seed_hi = (unsigned)(seed >> 32);
  P.step_lo = (unsigned)step;
  P.step_hi = (unsigned)(step >> 32) & 0xffffu;
  P.box_lo0 = g->desc.box_lo[0];
  P.box_lo1 = (g->desc.D == 2) ? g->desc.box_lo[1] : 0;
  P.nbox0 = g->nbox[0];
  P.ncell_glob0 = g->desc.ncell[0];
  const int ncell = (int)g->ncell_box;
  unsigned long long *d_np = &c.d_counters->npairs;  // zero between calls
  // cells of up to TA_NMAX particles per species go through the staged kernels, which leave the others in a list for
  // the general ones
  int *list = nullptr;
  if (c.ta_staged) {
    if (c.ta_list_cap < ncell + 1) {
      if (c.ta_list) cudaFree(c.ta_list);
      c.ta_list = nullptr;
      PGPU_CUDA(cudaMalloc(&c.ta_list, (size_t)(ncell + 1) * sizeof(int)));
      c.ta_list_cap = ncell + 1;
      cudaFuncSetAttribute(k_ta_self_staged<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(k_ta_self_staged<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(k_ta_inter_staged<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaFuncSetAttribute(k_ta_inter_staged<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    list = c.ta_list;
    PGPU_CUDA(cudaMemsetAsync(list, 0, sizeof(int), c.stream));
  }
  const unsigned tail_blocks = list ? (unsigned)(2 * c.sm_count) : nb((long)ncell * 32);
  if (sA == sB) {
    KTimer t("collide_ta_self");
    if (list) {
      auto k = P.rel ? k_ta_self_staged<1> : k_ta_self_staged<0>;
      k<<<(ncell + TA_SELF_WARPS - 1) / TA_SELF_WARPS, 32 * TA_SELF_WARPS, TA_SELF_WARPS * TA_SPECIES_BYTES, c.stream>>>(
          sA->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->id, sA->dens, P, d_np, list);
    }
    k_ta_self<<<tail_blocks, 256, 0, c.stream>>>(sA->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->id, sA->dens, P,
                                                 (unsigned *)sA->cell_key, sA->perm, d_np, list);
  } else {
    KTimer t("collide_ta_inter");
    if (list) {
      auto k = P.rel ? k_ta_inter_staged<1> : k_ta_inter_staged<0>;
      k<<<(ncell + TA_INTER_WARPS - 1) / TA_INTER_WARPS, 32 * TA_INTER_WARPS, TA_INTER_WARPS * 2 * TA_SPECIES_BYTES,
          c.stream>>>(sA->cell_start, sB->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->id, sA->dens, sB->v[0],
                      sB->v[1], sB->v[2], sB->id, sB->dens, P, d_np, list);
    }
    k_ta_inter<<<tail_blocks, 256, 0, c.stream>>>(
        sA->cell_start, sB->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->id, sA->dens, sB->v[0],
        sB->v[1], sB->v[2], sB->id, sB->dens, P, (unsigned *)sA->cell_key, sA->perm, (unsigned *)sB->cell_key,
        sB->perm, d_np, list);
  }
  if (npairs_out) {
    PGPU_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c.stream));
    PGPU_CUDA(cudaMemsetAsync(&c.d_counters->npairs, 0, sizeof(unsigned long long), c.stream));
    PGPU_CUDA(cudaStreamSynchronize(c.stream));
    *npairs_out = (long)c.h_counters->npairs;
  } else {
    PGPU_CUDA(cudaMemsetAsync(&c.d_counters->npairs, 0, sizeof(unsigned long long), c.stream));
  }
  return 0;
}


// ---- Scattering::setMeanFreeTime ---------------------------------------------------------------
static int need_moments(pgpu_species_t sA, pgpu_species_t sB) {
  if (!ctx().inited) {
    set_error("pgpu_init has not been called");
    return PGPU_ERR_STATE;
  }
  if (!sA || !sB || sA->grid != sB->grid) return PGPU_ERR_ARG;
  if (!sA->dens || !sB->dens) {
    set_error("setMeanFreeTime needs the cell moments: call pgpu_set_moments_from_bins first");
    return PGPU_ERR_STATE;
  }
  return 0;
}

static int fetch_nu_max(double *nu_max) {
  Context &c = ctx();
  PGPU_CUDA(cudaMemcpyAsync(c.h_counters, c.d_counters, sizeof(Counters), cudaMemcpyDeviceToHost, c.stream));
  PGPU_CUDA(cudaMemsetAsync(&c.d_counters->maxbits, 0, sizeof(unsigned long long), c.stream));
  PGPU_CUDA(cudaStreamSynchronize(c.stream));
  memcpy(nu_max, &c.h_counters->maxbits, sizeof(double));
  return 0;
}

int pgpu_scatter_nu_max_ta(pgpu_species_t sA, pgpu_species_t sB, double Clog, double *nu_max) {
  if (!sA) return PGPU_ERR_ARG;
  int rc = need_moments(sA, sB);
  if (rc) return rc;
  if (!nu_max) return PGPU_ERR_ARG;
  Context &c = ctx();
  const int ncell = (int)sA->grid->ncell_box;
  {
    KTimer t("nu_max");
    k_nu_max_ta<<<nb(ncell), 256, 0, c.stream>>>(ncell, sA->dens, sA->ene, sB->dens, sB->ene, sA->desc.charge,
                                                 sB->desc.charge, sA->desc.mass, sB->desc.mass, Clog, sA == sB,
                                                 &c.d_counters->maxbits);
  }
  return fetch_nu_max(nu_max);
}

int pgpu_scatter_nu_max_coulomb(pgpu_species_t sA, pgpu_species_t sB, const pgpu_coulomb_params *prm,
                                double *nu_max) {
  if (!sA) return PGPU_ERR_ARG;
  int rc = need_moments(sA, sB);
  if (rc) return rc;
  if (!nu_max || !prm) return PGPU_ERR_ARG;
  Context &c = ctx();
  const pgpu_grid_s *g = sA->grid;
  if (!g->debye) {
    set_error("Coulomb::setMeanFreeTime needs the Debye length: call pgpu_debye_length first");
    return PGPU_ERR_STATE;
  }
  CoulParams P;
  rc = coulomb_consts(sA->desc.charge, sB->desc.charge, sA->desc.mass, sB->desc.mass, prm, 0.0, &P);
  if (rc) return rc;
  const int ncell = (int)g->ncell_box;
  {
    KTimer t("nu_max");
    k_nu_max_coulomb<<<nb(ncell), 256, 0, c.stream>>>(ncell, g->debye, sA->dens, sA->mom, sA->ene, sB->dens, sB->mom,
                                                      sB->ene, sA->desc.mass, sB->desc.mass, P, sA == sB,
                                                      &c.d_counters->maxbits);
  }
  return fetch_nu_max(nu_max);
}

int pgpu_scatter_nu_max_elastic(pgpu_species_t sA, pgpu_species_t sB, const pgpu_elastic_params *prm,
                                double *nu_max) {
  if (!sA) return PGPU_ERR_ARG;
  int rc = need_moments(sA, sB);
  if (rc) return rc;
  if (!nu_max || !prm || (prm->ntab && (!prm->E || !prm->Q || prm->ntab < 2))) return PGPU_ERR_ARG;
  if (prm->ntab && prm->angular_scattering == 1 && !prm->xi) return PGPU_ERR_ARG;
  Context &c = ctx();
  const double CVAC = 2.99792458e+08, ME = 9.10938370e-31, QE = 1.60217663e-19;
  ElaParams P;
  const double m1 = sA->desc.mass, m2 = sB->desc.mass;
  P.mu = m1 * m2 / (m1 + m2);
  P.f1 = P.mu / m1;
  P.f2 = P.mu / m2;
  P.const_sigma = prm->const_sigma;
  P.dt_sec = 0.0;
  P.mcSq = ME * CVAC * CVAC / QE;
  P.ntab = prm->ntab;
  P.angular = prm->angular_scattering;
  P.loglog = prm->use_loglog_interp;
  P.E = P.Q = P.XI = nullptr;
  P.seed_lo = P.seed_hi = P.step_lo = P.step_hi = 0;
  double *d_tab = nullptr;
  if (prm->ntab) {
    const size_t N = (size_t)prm->ntab;
    PGPU_CUDA(cudaMalloc(&d_tab, 3 * N * sizeof(double)));
    PGPU_CUDA(cudaMemcpyAsync(d_tab, prm->E, N * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    PGPU_CUDA(cudaMemcpyAsync(d_tab + N, prm->Q, N * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    if (prm->xi) PGPU_CUDA(cudaMemcpyAsync(d_tab + 2 * N, prm->xi, N * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    else PGPU_CUDA(cudaMemsetAsync(d_tab + 2 * N, 0, N * sizeof(double), c.stream));
    P.E = d_tab;
    P.Q = d_tab + N;
    P.XI = d_tab + 2 * N;
  }
  const int ncell = (int)sA->grid->ncell_box;
  {
    // Reference quirk kept: Elastic::setMeanFreeTime reads species 1's moments
    // for BOTH species (Elastic.cpp:130-131); the masses are the two species' own.
    KTimer t("nu_max");
    k_nu_max_elastic<<<nb(ncell), 256, 0, c.stream>>>(ncell, sA->dens, sA->ene, sA->dens, sA->ene, m1, m2, P,
                                                      &c.d_counters->maxbits);
  }
  rc = fetch_nu_max(nu_max);   // synchronises
  if (d_tab) cudaFree(d_tab);
  return rc;
}

}  // extern "C"
