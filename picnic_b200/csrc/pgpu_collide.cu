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
    dU[0] = ux * uz / uperp * sinth * cosphi - uy * u / uperp * sinth * sinphi - ux * (1. - costh);
    dU[1] = uy * uz / uperp * sinth * cosphi + ux * u / uperp * sinth * sinphi - uy * (1. - costh);
    dU[2] = -uperp * sinth * cosphi - uz * (1. - costh);
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
    sinth = 2.0 * delta / (1.0 + deltasq);
    costh = 1.0 - 2.0 * deltasq / (1.0 + deltasq);
  } else {
    const double theta = PI * u_theta;
    sincos(theta, &sinth, &costh);
  }
  double sinphi, cosphi;
  sincos(TWOPI * u_phi, &sinphi, &cosphi);
  scatter_delta_u(ux, uy, uz, costh, sinth, cosphi, sinphi, dU);
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
  unsigned seed_lo, seed_hi, step_lo, step_hi;
  int box_lo0, box_lo1, nbox0, ncell_glob0;  // to form the global cell id
};

__device__ __forceinline__ unsigned global_cell(const TAParams &P, int cell) {
  const int i = cell % P.nbox0 + P.box_lo0, j = cell / P.nbox0 + P.box_lo1;
  return (unsigned)(i + j * P.ncell_glob0);
}

// random order of the n particles of a cell: order[r] = local index of the particle
// with the r-th smallest (key, id) where key = Philox(seed, step, id)
__device__ __forceinline__ void warp_shuffle_order(int s, int n, const uint64_t *id, const TAParams &P,
                                                   unsigned salt, unsigned *key, int *order, int lane) {
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
  // Box-Muller
  const double TWOPI = 6.28318530717958647692;
  gauss = sqrt(-2.0 * log(u01(r.x))) * cos(TWOPI * u01(r.y));
  uth = u01(r.z);
  uphi = u01(r.w);
}

__device__ __forceinline__ void scatter_pair(double *v0, double *v1, double *v2, int pa, double *w0,
                                             double *w1, double *w2, int pb, double dena, double denb,
                                             const TAParams &P, double gauss, double uth, double uphi) {
  double a[3] = {v0[pa], v1[pa], v2[pa]};
  double b[3] = {w0[pb], w1[pb], w2[pb]};
  double dU[3];
  ta_delta_u(a, dena, b, denb, P.b90_fact, P.Clog, P.dt_sec, gauss, uth, uphi, dU);
  v0[pa] = a[0] + P.f1 * dU[0];
  v1[pa] = a[1] + P.f1 * dU[1];
  v2[pa] = a[2] + P.f1 * dU[2];
  w0[pb] = b[0] - P.f2 * dU[0];
  w1[pb] = b[1] - P.f2 * dU[1];
  w2[pb] = b[2] - P.f2 * dU[2];
}

// TakizukaAbe::applySelfScattering (TakizukaAbe.cpp:263-402)
__global__ void __launch_bounds__(256)
k_ta_self(const int *cell_start, int ncell, double *v0, double *v1, double *v2, const uint64_t *id,
          const double *dens, TAParams P, unsigned *key, int *order, unsigned long long *npairs) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const double numDen = dens[cell];
  const int s = cell_start[cell], n = cell_start[cell + 1] - s;
  if (numDen == 0.0 || n < 2) return;
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
    const int t[3] = {s + order[s], s + order[s + 1], s + order[s + 2]};
    const int p1[3] = {0, 1, 0}, p2[3] = {1, 2, 2};
    for (int p = 0; p < 3; ++p) {
      double g, ut, up;
      pair_randoms(P, gcell, (unsigned)p, 1u, g, ut, up);
      scatter_pair(v0, v1, v2, t[p1[p]], v0, v1, v2, t[p2[p]], numDen / 2.0, numDen / 2.0, P, g, ut, up);
    }
  }
  if (lane == 0) atomicAdd(npairs, (unsigned long long)(nmain + (pstart == 3 ? 3 : 0)));
}

// TakizukaAbe::applyInterScattering (TakizukaAbe.cpp:404-536).  Pair p couples
// particle p%n1 with p (or p with p%n2): the particle of the shorter list collides
// repeatedly, in order of p, so one lane owns it and walks its partners.
__global__ void __launch_bounds__(256)
k_ta_inter(const int *cs1, const int *cs2, int ncell, double *a0, double *a1, double *a2, const uint64_t *id1,
           const double *dens1, double *b0, double *b1, double *b2, const uint64_t *id2, const double *dens2,
           TAParams P, unsigned *key1, int *order1, unsigned *key2, int *order2, unsigned long long *npairs) {
  const int cell = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (cell >= ncell) return;
  const double numDen1 = dens1[cell], numDen2 = dens2[cell];
  if (numDen1 * numDen2 == 0.0) return;
  const int s1 = cs1[cell], n1 = cs1[cell + 1] - s1;
  const int s2 = cs2[cell], n2 = cs2[cell + 1] - s2;
  if ((long)n1 * n2 < 2) return;
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
      scatter_pair(a0, a1, a2, i1, b0, b1, b2, i2, numDen1, numDen2, P, g, ut, up);
    }
  }
  if (lane == 0) atomicAdd(npairs, (unsigned long long)pMax);
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
  P.Clog = Clog;
  P.dt_sec = dt_sec;
  P.f1 = (double)(mu / m1);
  P.f2 = (double)(mu / m2);
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
  if (sA == sB) {
    KTimer t("collide_ta_self");
    k_ta_self<<<nb((long)ncell * 32), 256, 0, c.stream>>>(sA->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2],
                                                          sA->id, sA->dens, P, (unsigned *)sA->cell_key, sA->perm,
                                                          d_np);
  } else {
    KTimer t("collide_ta_inter");
    k_ta_inter<<<nb((long)ncell * 32), 256, 0, c.stream>>>(
        sA->cell_start, sB->cell_start, ncell, sA->v[0], sA->v[1], sA->v[2], sA->id, sA->dens, sB->v[0],
        sB->v[1], sB->v[2], sB->id, sB->dens, P, (unsigned *)sA->cell_key, sA->perm, (unsigned *)sB->cell_key,
        sB->perm, d_np);
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

}  // extern "C"
