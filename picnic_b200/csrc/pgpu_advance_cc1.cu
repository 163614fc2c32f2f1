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

constexpr int ITER_EXPLICIT = -2;   // FastArgs::iter_max of the explicit leap-frog step (v2 kernel only)

struct FastArgs {
  const double *xo[2];
  double *xb[2];
  double *xbo[2];           // where xbar is written: xb, or the stale xold arrays when xold is aliased to x (tab kernel)
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
  // coefficient tables of the selected field slot (k_build_tables)
  const double *tdual, *tnode;
  int tlo[2], tn0;
  double tol[2];            // rtol * dx
  const int4 *tile_box;     // per tile: dual-cell window (i, j, columns, rows) staged in shared memory
  int prefetch;             // persistent grid: pull the block's next tile into L2 while this one is computed
  int suborbit;             // m_use_suborbit_model: a particle that hits iter_max is deferred, not counted
};

// 21 fp64 REDs of one accumulator set into the J arrays at dual cell `key` (one pointer per grid row:
// the rest are immediate offsets)
__device__ __forceinline__ void flush_direct(const FastArgs &A, unsigned key, const double (&acc)[NSLOT]) {
  const int ci = (int)(key & 0xffffu) - 32768, cj = (int)(key >> 16) - 32768;
  {
    double *p0 = A.J[0] + (ci + cj * A.jn0[0]);
    double *p1 = p0 + A.jn0[0];
    double *p2 = p1 + A.jn0[0];
    atomicAdd(p0, acc[0]), atomicAdd(p0 + 1, acc[1]);
    atomicAdd(p1, acc[2]), atomicAdd(p1 + 1, acc[3]);
    atomicAdd(p2, acc[4]), atomicAdd(p2 + 1, acc[5]);
  }
  {
    double *p0 = A.J[1] + (ci + cj * A.jn0[1]);
    double *p1 = p0 + A.jn0[1];
#pragma unroll
    for (int a = 0; a < 3; ++a) atomicAdd(p0 + a, acc[6 + a]), atomicAdd(p1 + a, acc[9 + a]);
  }
  {
    double *p0 = A.J[2] + (ci + cj * A.jn0[2]);
    double *p1 = p0 + A.jn0[2];
    double *p2 = p1 + A.jn0[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) atomicAdd(p0 + a, acc[12 + a]), atomicAdd(p1 + a, acc[15 + a]), atomicAdd(p2 + a, acc[18 + a]);
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
__device__ __forceinline__ void l2_prefetch(const void *src_gmem, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int TP = 4;                 // consecutive particles per thread
constexpr int TILE = BLOCK * TP;      // particles per tile
constexpr int NIN = 8;                // xo0 xo1 xb0 xb1 uo0 uo1 uo2 w

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
      if (A.suborbit) return false;   // sub-orbit model: the generic kernel redoes it and lists it for the sub-orbit container
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

// =============================================================================================
// Table-driven two-phase tile kernel (the default).
//
// All particles of a dual cell read the same stencil, so the stencil algebra that does not
// depend on the particle is hoisted out of the particle loop into per-cell coefficient tables,
// rebuilt by k_build_tables whenever the E/B arrays of a field slot change (one thread per
// cell; the tables of a 512^2 box are 59 MB and stay L2 resident while the sorted particles
// stream past):
//   dual record (16 doubles, one per dual cell (i0,j0)):
//     in-plane E as  E0 = g1 + W0'*G0 + W2'*G2,  g1 = e1 + del0*d1, Gk = Pk + del0*Qk,
//     with W' = 4 W (the 1/4 of the CC1 weights is folded into P,Q) and W1 = 1 - W0 - W2
//     eliminated:  [e1 d1 P0 Q0 P2 Q2] for Ex, the same six for Ey, and the four bilinear
//     coefficients of Bz.
//   node record (12 doubles, one per index pair (i,j)): bilinear coefficients
//     c0 + fx c1 + fy c2 + fx fy c3 of Ez over nodes (i..i+1, j..j+1), of Bx (nodal in x,
//     centred in y) and of By (centred in x, nodal in y).
// A particle-Picard pass is then 78 fp64 instructions and six 128-bit loads.  Phase 1 leaves
// the dual-cell key and the normalised offsets (d_old, d_bar) in the shared-memory tile, so the
// deposit phase does not locate the particle again.
// =============================================================================================
// The dual record is padded to 18 doubles (144 B) and the rows of the shared-memory window are
// offset by another 64 B: with a 128 B stride the same word of two neighbouring records falls
// into the same banks, and the 128-bit record loads of a warp that straddles two or three dual
// cells (the usual case) then replay -- ncu: 8.5 wavefronts per load instead of ~3.  The node
// record (96 B) does not need it.
constexpr int TD = 18, TN = 12, TDU = 16;   // stride / stride / doubles used of the dual record

struct TabArgs {
  const double *F[6];   // origin-shifted like FastArgs::F
  int fn0[6];
  int lo[2], n0, n1;    // records for indices lo .. lo+n-1 in either direction
  double *dual, *node;
};

__global__ void __launch_bounds__(128) k_build_tables(const TabArgs T) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T.n0 * T.n1) return;
  const int i = T.lo[0] + t % T.n0, j = T.lo[1] + t / T.n0;
  auto at = [&](int c, int a, int b) { return __ldg(T.F[c] + ((i + a) + (long)(j + b) * T.fn0[c])); };
  auto bilinear = [&](int c, double *o) {
    const double v00 = at(c, 0, 0), v10 = at(c, 1, 0), v01 = at(c, 0, 1), v11 = at(c, 1, 1);
    o[0] = v00;
    o[1] = v10 - v00;
    o[2] = v01 - v00;
    o[3] = (v11 - v01) - (v10 - v00);
  };
  double rec[TDU];
  bilinear(2, rec);        // Ez
  bilinear(3, rec + 4);    // Bx
  bilinear(4, rec + 8);    // By
  double2 *dn = reinterpret_cast<double2 *>(T.node + (size_t)t * TN);
#pragma unroll
  for (int k = 0; k < TN / 2; ++k) dn[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
  // the dual record needs one more row/column of Ex, Ey: only for true dual cells
  if (t % T.n0 == T.n0 - 1 || t / T.n0 == T.n1 - 1) return;
  {
    // Ex(i+a, j+b), a<2, b<3: value at a=0 and x-difference per row b
    double e[3], d[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      e[b] = at(0, 0, b);
      d[b] = at(0, 1, b) - e[b];
    }
    rec[0] = e[1];
    rec[1] = d[1];
    rec[2] = 0.25 * (e[0] - e[1]);
    rec[3] = 0.25 * (d[0] - d[1]);
    rec[4] = 0.25 * (e[2] - e[1]);
    rec[5] = 0.25 * (d[2] - d[1]);
  }
  {
    // Ey(i+a, j+b), a<3, b<2: value at b=0 and y-difference per column a
    double e[3], d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      e[a] = at(1, a, 0);
      d[a] = at(1, a, 1) - e[a];
    }
    rec[6] = e[1];
    rec[7] = d[1];
    rec[8] = 0.25 * (e[0] - e[1]);
    rec[9] = 0.25 * (d[0] - d[1]);
    rec[10] = 0.25 * (e[2] - e[1]);
    rec[11] = 0.25 * (d[2] - d[1]);
  }
  bilinear(5, rec + 12);   // Bz
  double2 *dd = reinterpret_cast<double2 *>(T.dual + (size_t)t * TD);
#pragma unroll
  for (int k = 0; k < TDU / 2; ++k) dd[k] = make_double2(rec[2 * k], rec[2 * k + 1]);
}

__device__ __forceinline__ unsigned hi_abs(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }
// hi word of the doubles in [0.5 - 2^-22, 0.5): |x| with a smaller hi word is < 0.5 - 2.3e-7
constexpr unsigned HI_HALF_BAND = 0x3fdfffffu;

// 1/den for den >= 1: MUFU.RCP64H seed + two Newton steps (full double precision, no
// special-case branch; den = 1 + |b|^2 is never 0, inf or denormal for finite fields)
__device__ __forceinline__ double rcp_ge1(double den) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
  double e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  return r;
}

// 128-bit load through a generic pointer (the table window lives in shared memory, or in
// global memory for particles outside the window)
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }


#ifndef PGPU_TAB_LDGSTS
#define PGPU_TAB_LDGSTS 0
#endif

struct TabWindow {
  const double *dual, *node;   // record (i0, j0) of the window origin
  int i, j, ncol, nrow;        // dual-cell window
  int drow, nrow_stride;       // doubles between rows of the dual / node records
};

// Phase 1 with the coefficient tables.  On success key/dO/dB describe the final orbit.
// Dual record of the last particle this thread pushed: the cell sort makes a thread's consecutive
// particles share their dual cell nearly always, so the eight 128-bit record loads are skipped
// unless the key changes (they were 1/4 of the kernel's shared-memory wavefronts).
struct DualRec {
  unsigned key;
  double2 x01, x23, x45, y01, y23, y45, z01, z23;
  // NODECACHE: the node records of the last (dual cell, half-cell of xbar) this thread gathered from
  unsigned nkey, nsel;
  double2 ez01, ez23, bx01, bx23, by01, by23;
};

template <bool NODECACHE>
__device__ __forceinline__ bool push_tab(const FastArgs &A, const TabWindow &Wn, DualRec &R, const double (&xo)[2],
                                         double (&xb)[2], const double (&uo)[3], double (&ub)[3],
                                         unsigned &key, double (&dO)[2], double (&dB)[2], unsigned &apply,
                                         unsigned &unconv) {
  int i0[2];
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double xr = __dsub_rn(xo[d], A.le[d]);
    i0[d] = floor_div_fast(__dsub_rn(xr, A.hdx[d]), A.dx[d], A.rdx[d]);
    dO[d] = fma(xr, A.rdx[d], -(double)(i0[d] + 1));
    if (i0[d] < A.i_lo[d] || i0[d] > A.i_hi[d]) ok = false;
  }
  if (!ok) return false;
  key = ((unsigned)(i0[1] + 32768) << 16) | (unsigned)(i0[0] + 32768);
  const double *td, *tn;
  int nrow;
  {
    const unsigned wi = (unsigned)(i0[0] - Wn.i), wj = (unsigned)(i0[1] - Wn.j);
    if (wi < (unsigned)Wn.ncol && wj < (unsigned)Wn.nrow) {
      td = Wn.dual + (wi * TD + wj * Wn.drow);
      tn = Wn.node + (wi * TN + wj * Wn.nrow_stride);
      nrow = Wn.nrow_stride;
    } else {
      const int cidx = (i0[0] - A.tlo[0]) + (i0[1] - A.tlo[1]) * A.tn0;
      td = A.tdual + (size_t)cidx * TD;
      tn = A.tnode + (size_t)cidx * TN;
      nrow = A.tn0 * TN;
    }
  }
  if (key != R.key) {
    R.key = key;
    R.x01 = ld2(td), R.x23 = ld2(td + 2), R.x45 = ld2(td + 4);       // Ex: e1 d1 | P0 Q0 | P2 Q2
    R.y01 = ld2(td + 6), R.y23 = ld2(td + 8), R.y45 = ld2(td + 10);   // Ey
    R.z01 = ld2(td + 12), R.z23 = ld2(td + 14);                        // Bz c0 c1 | c2 c3
  }
  const double2 x01 = R.x01, x23 = R.x23, x45 = R.x45, y01 = R.y01, y23 = R.y23, y45 = R.y45, z01 = R.z01,
                z23 = R.z23;
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
    double dxp0[2], dN[2];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      dxp0[d] = xb[d] - xo[d];
      dB[d] = fma(dxp0[d], A.rdx[d], dO[d]);
      dN[d] = fma(2.0, dB[d], -dO[d]);
    }
    if (!(hi_abs(dN[0]) < HI_HALF_BAND && hi_abs(dN[1]) < HI_HALF_BAND)) {
      // near (or past) a dual-cell face: the reference's own floor decides
      bool same = true;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const double xn = fma(2.0, xb[d], -xo[d]);
        const int in = floor_div_exact(__dsub_rn(__dsub_rn(xn, A.le[d]), A.hdx[d]), A.dx[d]);
        if (in != i0[d]) same = false;
      }
      if (!same) return false;
    }
    if (done) break;
    const double del0 = dB[0] + 0.5, del1 = dB[1] + 0.5;
    double E[3], B[3];
    {
      const double a0 = 0.5 - dN[0], b0 = 0.5 + dN[0], a1 = 0.5 - dN[1], b1 = 0.5 + dN[1];
      const double Wx0 = fma(a0, a0, pO[0][0]), Wx2 = fma(b0, b0, pO[0][1]);   // 4 W
      const double Wy0 = fma(a1, a1, pO[1][0]), Wy2 = fma(b1, b1, pO[1][1]);
      E[0] = fma(Wy0, fma(del0, x23.y, x23.x), fma(Wy2, fma(del0, x45.y, x45.x), fma(del0, x01.y, x01.x)));
      E[1] = fma(Wx0, fma(del1, y23.y, y23.x), fma(Wx2, fma(del1, y45.y, y45.x), fma(del1, y01.y, y01.x)));
    }
    {
      // nodal CIC at xbar: node pair (i0+s, i0+s+1), fraction f
      const bool sx = del0 >= 0.5, sy = del1 >= 0.5;
      const double fx = del0 + (sx ? -0.5 : 0.5), fy = del1 + (sy ? -0.5 : 0.5);
      const int ox = sx ? TN : 0, oy = sy ? nrow : 0;
      double2 ez01, ez23, bx01, bx23, by01, by23;
      if (NODECACHE) {
        const unsigned sel = (sx ? 1u : 0u) | (sy ? 2u : 0u);
        if (R.nkey != key || R.nsel != sel) {
          R.nkey = key;
          R.nsel = sel;
          R.ez01 = ld2(tn + ox + oy), R.ez23 = ld2(tn + ox + oy + 2);
          R.bx01 = ld2(tn + ox + 4), R.bx23 = ld2(tn + ox + 6);
          R.by01 = ld2(tn + oy + 8), R.by23 = ld2(tn + oy + 10);
        }
        ez01 = R.ez01, ez23 = R.ez23, bx01 = R.bx01, bx23 = R.bx23, by01 = R.by01, by23 = R.by23;
      } else {
        ez01 = ld2(tn + ox + oy), ez23 = ld2(tn + ox + oy + 2);
        bx01 = ld2(tn + ox + 4), bx23 = ld2(tn + ox + 6);
        by01 = ld2(tn + oy + 8), by23 = ld2(tn + oy + 10);
      }
      E[2] = fma(fy, fma(fx, ez23.y, ez23.x), fma(fx, ez01.y, ez01.x));
      B[0] = fma(del1, fma(fx, bx23.y, bx23.x), fma(fx, bx01.y, bx01.x));
      B[1] = fma(fy, fma(del0, by23.y, by23.x), fma(del0, by01.y, by01.x));
      B[2] = fma(del1, fma(del0, z23.y, z23.x), fma(del0, z01.y, z01.x));
    }
    // Boris half step (PicSpeciesUtils.cpp:8-101)
    {
      const double vm0 = fma(A.alpha, E[0], uo[0]), vm1 = fma(A.alpha, E[1], uo[1]),
                   vm2 = fma(A.alpha, E[2], uo[2]);
      const double b0 = A.alpha * B[0], b1 = A.alpha * B[1], b2 = A.alpha * B[2];
      const double den = fma(b2, b2, fma(b1, b1, fma(b0, b0, 1.0)));
      const double p0 = fma(-vm2, b1, fma(vm1, b2, vm0));
      const double p1 = fma(-vm0, b2, fma(vm2, b0, vm1));
      const double p2 = fma(-vm1, b0, fma(vm0, b1, vm2));
      const double rden = rcp_ge1(den);
      const double r0 = b0 * rden, r1 = b1 * rden, r2 = b2 * rden;
      ub[0] = fma(-p2, r1, fma(p1, r2, vm0));
      ub[1] = fma(-p0, r2, fma(p2, r0, vm1));
      ub[2] = fma(-p1, r0, fma(p0, r1, vm2));
    }
    napply += 1;
    if (A.iter_max < 0) {  // advanceParticles (:1594-1612), part_order_swap == false
      xb[0] = fma(ub[0], A.hdt, xo[0]);
      xb[1] = fma(ub[1], A.hdt, xo[1]);
      done = true;
      continue;
    }
    // stepNormTransfer (:658-733): |dxp0 - dxp| / dX against rtol, as |dxp0 - dxp| against rtol*dX
    const double dxp_0 = ub[0] * A.hdt, dxp_1 = ub[1] * A.hdt;
    const double e0 = fabs(dxp0[0] - dxp_0), e1 = fabs(dxp0[1] - dxp_1);
    if (iter == 0) {
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
      if (!(e0 >= A.tol[0]) && !(e1 >= A.tol[1])) done = true;
    } else {
      if (e0 < A.tol[0] && e1 < A.tol[1]) break;  // reverse pass: xbar unchanged, its orbit was checked above
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
    }
    if (!done && iter >= A.iter_max) {
      if (A.suborbit) return false;   // sub-orbit model: the generic kernel redoes it and lists it for the sub-orbit container
      nunconv = 1;
      done = true;
    }
    iter += 1;
  }
  apply += napply;
  unconv += nunconv;
  return true;
}

// ---- phase 1, NP particles of ONE dual cell in lockstep ------------------------------------------
// The Picard pass is a chain of dependent fp64 instructions (8 cycles each on B200) and a warp issues in
// order: with 4 warps per scheduler the chain latency, not the pipe, sets the pace.  Two particles of the same
// dual cell (neighbours in the cell-sorted order) share the cell's table record and their chains are
// independent, so interleaving them doubles the instructions in flight per warp.
struct TabCell {
  unsigned key;
  int i0[2];
  const double *td, *tn;
  int nrow;
};
__device__ __forceinline__ bool locate_cell(const FastArgs &A, const TabWindow &Wn, const double (&xo)[2], TabCell &C,
                                            double (&dO)[2]) {
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double xr = __dsub_rn(xo[d], A.le[d]);
    C.i0[d] = floor_div_fast(__dsub_rn(xr, A.hdx[d]), A.dx[d], A.rdx[d]);
    dO[d] = fma(xr, A.rdx[d], -(double)(C.i0[d] + 1));
    if (C.i0[d] < A.i_lo[d] || C.i0[d] > A.i_hi[d]) ok = false;
  }
  if (!ok) return false;
  C.key = ((unsigned)(C.i0[1] + 32768) << 16) | (unsigned)(C.i0[0] + 32768);
  const unsigned wi = (unsigned)(C.i0[0] - Wn.i), wj = (unsigned)(C.i0[1] - Wn.j);
  if (wi < (unsigned)Wn.ncol && wj < (unsigned)Wn.nrow) {
    C.td = Wn.dual + (wi * TD + wj * Wn.drow);
    C.tn = Wn.node + (wi * TN + wj * Wn.nrow_stride);
    C.nrow = Wn.nrow_stride;
  } else {
    const int cidx = (C.i0[0] - A.tlo[0]) + (C.i0[1] - A.tlo[1]) * A.tn0;
    C.td = A.tdual + (size_t)cidx * TD;
    C.tn = A.tnode + (size_t)cidx * TN;
    C.nrow = A.tn0 * TN;
  }
  return true;
}
__device__ __forceinline__ void load_record(DualRec &R, const TabCell &C) {
  if (C.key != R.key) {
    R.key = C.key;
    R.x01 = ld2(C.td), R.x23 = ld2(C.td + 2), R.x45 = ld2(C.td + 4);
    R.y01 = ld2(C.td + 6), R.y23 = ld2(C.td + 8), R.y45 = ld2(C.td + 10);
    R.z01 = ld2(C.td + 12), R.z23 = ld2(C.td + 14);
  }
}

// Same arithmetic and the same per-particle control flow as push_tab (pass order of advanceParticlesIteratively,
// PicChargedSpecies.cpp:1614-1716).  ok[p]: in = particle present, out = false if it has to be deferred.
template <int NP>
__device__ __forceinline__ void picard_same_cell(const FastArgs &A, const TabCell (&CC)[NP],
                                                 const double (&xo)[NP][2], double (&xb)[NP][2],
                                                 const double (&uo)[NP][3], double (&ub)[NP][3],
                                                 const double (&dO)[NP][2], bool (&ok)[NP], unsigned &apply,
                                                 unsigned &unconv) {
  double pO[NP][2][2];
  bool live[NP], done[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      const double a = 0.5 - dO[p][d], b = 0.5 + dO[p][d];
      pO[p][d][0] = a * a;
      pO[p][d][1] = b * b;
    }
    live[p] = ok[p];
    done[p] = false;
  }
  int iter = 0;
  unsigned napply[NP], nunconv[NP];   // counted only for particles that are not deferred (the generic kernel redoes those)
  double dxp0[NP][2], dB[NP][2], dN[NP][2];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    napply[p] = nunconv[p] = 0;
#pragma unroll
    for (int d = 0; d < 2; ++d) dxp0[p][d] = dB[p][d] = dN[p][d] = 0.0;
  }
  while (true) {
    bool any = false;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!live[p]) continue;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        dxp0[p][d] = xb[p][d] - xo[p][d];
        dB[p][d] = fma(dxp0[p][d], A.rdx[d], dO[p][d]);
        dN[p][d] = fma(2.0, dB[p][d], -dO[p][d]);
      }
      if (!(hi_abs(dN[p][0]) < HI_HALF_BAND && hi_abs(dN[p][1]) < HI_HALF_BAND)) {
        bool same = true;   // near (or past) a dual-cell face: the reference's own floor decides
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const double xn = fma(2.0, xb[p][d], -xo[p][d]);
          const int in = floor_div_exact(__dsub_rn(__dsub_rn(xn, A.le[d]), A.hdx[d]), A.dx[d]);
          if (in != CC[p].i0[d]) same = false;
        }
        if (!same) {
          ok[p] = false;
          live[p] = false;
        }
      }
      if (live[p] && done[p]) live[p] = false;   // the final iterate's orbit stays in the cell: finished
      any = any || live[p];
    }
    if (!any) break;
    double un[NP][3];
    asm volatile("" ::: "memory");   // keep the table loads inside the pass (hoisted they cost 32 registers a particle)
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      // a finished particle is recomputed from its last offsets; the result is dropped
      const TabCell &C = CC[p];
      DualRec R;
      R.x01 = ld2(C.td), R.x23 = ld2(C.td + 2), R.x45 = ld2(C.td + 4);
      R.y01 = ld2(C.td + 6), R.y23 = ld2(C.td + 8), R.y45 = ld2(C.td + 10);
      R.z01 = ld2(C.td + 12), R.z23 = ld2(C.td + 14);
      const double del0 = dB[p][0] + 0.5, del1 = dB[p][1] + 0.5;
      double E[3], B[3];
      {
        const double a0 = 0.5 - dN[p][0], b0 = 0.5 + dN[p][0], a1 = 0.5 - dN[p][1], b1 = 0.5 + dN[p][1];
        const double Wx0 = fma(a0, a0, pO[p][0][0]), Wx2 = fma(b0, b0, pO[p][0][1]);
        const double Wy0 = fma(a1, a1, pO[p][1][0]), Wy2 = fma(b1, b1, pO[p][1][1]);
        E[0] = fma(Wy0, fma(del0, R.x23.y, R.x23.x), fma(Wy2, fma(del0, R.x45.y, R.x45.x), fma(del0, R.x01.y, R.x01.x)));
        E[1] = fma(Wx0, fma(del1, R.y23.y, R.y23.x), fma(Wx2, fma(del1, R.y45.y, R.y45.x), fma(del1, R.y01.y, R.y01.x)));
      }
      {
        const bool sx = del0 >= 0.5, sy = del1 >= 0.5;
        const double fx = del0 + (sx ? -0.5 : 0.5), fy = del1 + (sy ? -0.5 : 0.5);
        const int ox = sx ? TN : 0, oy = sy ? C.nrow : 0;
        const double2 ez01 = ld2(C.tn + ox + oy), ez23 = ld2(C.tn + ox + oy + 2);
        const double2 bx01 = ld2(C.tn + ox + 4), bx23 = ld2(C.tn + ox + 6);
        const double2 by01 = ld2(C.tn + oy + 8), by23 = ld2(C.tn + oy + 10);
        E[2] = fma(fy, fma(fx, ez23.y, ez23.x), fma(fx, ez01.y, ez01.x));
        B[0] = fma(del1, fma(fx, bx23.y, bx23.x), fma(fx, bx01.y, bx01.x));
        B[1] = fma(fy, fma(del0, by23.y, by23.x), fma(del0, by01.y, by01.x));
        B[2] = fma(del1, fma(del0, R.z23.y, R.z23.x), fma(del0, R.z01.y, R.z01.x));
      }
      {
        const double vm0 = fma(A.alpha, E[0], uo[p][0]), vm1 = fma(A.alpha, E[1], uo[p][1]),
                     vm2 = fma(A.alpha, E[2], uo[p][2]);
        const double b0 = A.alpha * B[0], b1 = A.alpha * B[1], b2 = A.alpha * B[2];
        const double den = fma(b2, b2, fma(b1, b1, fma(b0, b0, 1.0)));
        const double p0 = fma(-vm2, b1, fma(vm1, b2, vm0));
        const double p1 = fma(-vm0, b2, fma(vm2, b0, vm1));
        const double p2 = fma(-vm1, b0, fma(vm0, b1, vm2));
        const double rden = rcp_ge1(den);
        const double r0 = b0 * rden, r1 = b1 * rden, r2 = b2 * rden;
        un[p][0] = fma(-p2, r1, fma(p1, r2, vm0));
        un[p][1] = fma(-p0, r2, fma(p2, r0, vm1));
        un[p][2] = fma(-p1, r0, fma(p0, r1, vm2));
      }
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      if (!live[p]) continue;
      napply[p] += 1;
      ub[p][0] = un[p][0], ub[p][1] = un[p][1], ub[p][2] = un[p][2];
      if (A.iter_max < 0) {  // advanceParticles (:1594-1612), part_order_swap == false
        xb[p][0] = fma(un[p][0], A.hdt, xo[p][0]);
        xb[p][1] = fma(un[p][1], A.hdt, xo[p][1]);
        done[p] = true;
        continue;
      }
      // stepNormTransfer (:658-733)
      const double dxp_0 = un[p][0] * A.hdt, dxp_1 = un[p][1] * A.hdt;
      const double e0 = fabs(dxp0[p][0] - dxp_0), e1 = fabs(dxp0[p][1] - dxp_1);
      if (iter == 0) {
        xb[p][0] = xo[p][0] + dxp_0;
        xb[p][1] = xo[p][1] + dxp_1;
        if (!(e0 >= A.tol[0]) && !(e1 >= A.tol[1])) done[p] = true;
      } else {
        if (e0 < A.tol[0] && e1 < A.tol[1]) {   // reverse pass: xbar unchanged, its orbit was checked above
          live[p] = false;
          continue;
        }
        xb[p][0] = xo[p][0] + dxp_0;
        xb[p][1] = xo[p][1] + dxp_1;
      }
      if (!done[p] && iter >= A.iter_max) {
        nunconv[p] = 1;
        done[p] = true;
      }
    }
    iter += 1;
  }
#pragma unroll
  for (int p = 0; p < NP; ++p)
    if (ok[p]) {
      apply += napply[p];
      unconv += nunconv[p];
    }
}

// Phase 2: add the 21 node contributions of one pushed particle into acc (FMA form).
__device__ __forceinline__ void deposit_tab(const FastArgs &A, const double (&dO)[2], const double (&dB)[2],
                                            const double (&ub)[3], double wp, double (&acc)[NSLOT]) {
  double W[2][3], n[2][3], del[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double dN = fma(2.0, dB[d], -dO[d]);
    const double a = 0.5 - dN, b = 0.5 + dN, c = 0.5 - dO[d], e = 0.5 + dO[d];
    W[d][0] = 0.25 * fma(a, a, c * c);
    W[d][2] = 0.25 * fma(b, b, e * e);
    W[d][1] = (1.0 - W[d][0]) - W[d][2];
    del[d] = dB[d] + 0.5;
    // nodal CIC at xbar over the three nodes of the dual cell
    const bool s = del[d] >= 0.5;
    const double lo = 0.5 - del[d], hi = del[d] - 0.5;     // one of them is the (positive) end weight
    n[d][0] = s ? 0.0 : lo;
    n[d][2] = s ? hi : 0.0;
    n[d][1] = s ? 1.0 - hi : 1.0 - lo;
  }
  const double rhop = wp * A.rvolume;
  const double jx = ub[0] * rhop, jy = ub[1] * rhop, jz = ub[2] * rhop;
  // Jx(i0+a, j0+b), a<2, b<3  -> slot a + 2 b
  const double jx1 = jx * del[0], jx0 = jx - jx1;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    acc[0 + 2 * b] = fma(jx0, W[1][b], acc[0 + 2 * b]);
    acc[1 + 2 * b] = fma(jx1, W[1][b], acc[1 + 2 * b]);
  }
  // Jy(i0+a, j0+b), a<3, b<2  -> slot 6 + a + 3 b
  const double jy1 = jy * del[1], jy0 = jy - jy1;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    acc[6 + a] = fma(jy0, W[0][a], acc[6 + a]);
    acc[9 + a] = fma(jy1, W[0][a], acc[9 + a]);
  }
  // Jz nodal CIC over nodes i0..i0+2 x j0..j0+2 -> slot 12 + a + 3 b
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const double t = jz * n[1][b];
#pragma unroll
    for (int a = 0; a < 3; ++a) acc[12 + a + 3 * b] = fma(t, n[0][a], acc[12 + a + 3 * b]);
  }
}

constexpr int NTAB = 8;    // xo0->dO0 xo1->dO1 xb0 xb1 u0 u1 u2 w
constexpr int WMAX = 16;   // widest dual-cell window staged in shared memory (2 rows; node records: 3 rows, +1 column)
constexpr int DROW = WMAX * TD + 8;   // doubles between the two rows of the staged dual records
constexpr int SDUAL = 2 * DROW, SNODE = 3 * (WMAX + 1) * TN;
// the same storage as a window of THREE dual rows (four node rows) of up to W3 columns: tiles of particles sorted by
// dual cell (pgpu_sort_for_locality) sit in one dual row, and the rows above and below catch the particles that
// drifted since the sort
constexpr int W3 = 10;
constexpr int DROW3 = W3 * TD + 8, NROW3 = (W3 + 1) * TN;
static_assert(3 * DROW3 <= SDUAL && 4 * NROW3 <= SNODE, "three-row window does not fit");
__device__ __forceinline__ int win_drow(int rows) { return rows == 3 ? DROW3 : DROW; }
__device__ __forceinline__ int win_nstride(int rows) { return rows == 3 ? NROW3 : (WMAX + 1) * TN; }
constexpr size_t TAB_SMEM =
    (size_t)(NTAB * TILE + SDUAL + SNODE) * sizeof(double) + TILE * sizeof(unsigned) + 64;

// Window of a tile = the dual cells its particles can sit in if the tile is cell sorted: the
// cells of its first and last particle bound the rest.  Particles outside the window (unsorted
// input, or moved since the sort) read the tables from global memory instead -- any order is
// correct, the sorted one is fast.
__global__ void k_tile_boxes(const FastArgs A, int ntiles, int4 *box) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  const long first = (long)t * TILE;
  const long last = (first + TILE <= A.n ? first + TILE : A.n) - 1;
  int c[2][2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const long p = e ? last : first;
#pragma unroll
    for (int d = 0; d < 2; ++d)   // primal cell (BinFab::locateBin)
      c[e][d] = __double2int_rd(__ddiv_rn(__dsub_rn(A.xo[d][p], A.le[d]), A.dx[d]));
  }
  int4 b = make_int4(0, 0, 0, 0);
  // dual cells of the first and last particle: floor((x - le - dx/2)/dx)
  int dc[2][2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const long p = e ? last : first;
#pragma unroll
    for (int d = 0; d < 2; ++d)
      dc[e][d] = __double2int_rd(__ddiv_rn(__dsub_rn(__dsub_rn(A.xo[d][p], A.le[d]), A.hdx[d]), A.dx[d]));
  }
  if (dc[0][1] == dc[1][1] && dc[1][0] >= dc[0][0] && dc[1][0] - dc[0][0] + 3 <= W3) {
    // one dual row (tile sorted by dual cell, or a stretch of one half-row of primal cells): three rows, +-1 column
    const int i0 = max(dc[0][0] - 1, A.i_lo[0]), i1 = min(dc[1][0] + 1, A.i_hi[0]);
    const int j0 = max(dc[0][1] - 1, A.i_lo[1]), j1 = min(dc[0][1] + 1, A.i_hi[1]);
    if (i1 >= i0 && j1 >= j0) b = make_int4(i0, j0, i1 - i0 + 1, j1 - j0 + 1);
    if (b.w == 3 || b.z == 0) {
      box[t] = b;
      return;
    }
    b = make_int4(0, 0, 0, 0);   // clipped to fewer rows at the table edge: the two-row layout below handles it
  }
  if (c[0][1] == c[1][1] && c[1][0] >= c[0][0]) {
    // dual columns i-1 .. i_last, dual rows j-1 .. j, clipped to the table
    const int i0 = max(c[0][0] - 1, A.i_lo[0]), i1 = min(c[1][0], A.i_hi[0]);
    const int j0 = max(c[0][1] - 1, A.i_lo[1]), j1 = min(c[0][1], A.i_hi[1]);
    if (i1 >= i0 && j1 >= j0 && i1 - i0 + 1 <= WMAX) b = make_int4(i0, j0, i1 - i0 + 1, j1 - j0 + 1);
  }
  box[t] = b;
}

template <bool DEP, int RSTEPS, int MINB, bool PAIR, bool NODECACHE = false, bool ALIAS = false>
__global__ void __launch_bounds__(BLOCK, MINB) k_advance_cc1_2d_tab(const FastArgs A, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *st = reinterpret_cast<double *>(smem_raw);                 // [NTAB][TILE]
  double *sdual = st + NTAB * TILE;                                   // [2][WMAX][TD]
  double *snode = sdual + SDUAL;                                      // [3][WMAX+1][TN]
  unsigned *skey = reinterpret_cast<unsigned *>(snode + SNODE);      // [TILE]
  uint64_t *bar = reinterpret_cast<uint64_t *>(skey + TILE);
  __shared__ int4 sbox;
  const int tid = threadIdx.x, lane = tid & 31;

  // window of this block's next tile, fetched one tile ahead by the issuing lanes
  int nbx = 0, nby = 0, nbz = 0, nbw = 0;
  auto fetch_box = [&](int t) {
    const int4 b = __ldg(A.tile_box + t);
    nbx = b.x;
    nby = b.y;
    nbz = b.z;
    nbw = b.w;
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (lane == 0 && (int)blockIdx.x < ntiles) fetch_box(blockIdx.x);
  __syncthreads();

  unsigned apply = 0, unconv = 0;
  int it = 0;
#ifdef PGPU_CLOCKS
  long long ck[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c0, c1;
#define CK(i) do { c1 = clock64(); ck[i] += c1 - c0; c0 = c1; } while (0)
  c0 = clock64();
#else
#define CK(i)
#endif
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const long tbase = (long)tile * TILE;
    if (lane == 0) {
      // lane 0 of each warp issues a share of the tile's bulk copies (a bulk copy is issued by one
      // thread; four threads in four warps issue in parallel): warp w loads particle arrays 2w and
      // 2w+1 -- the ones it stored for the previous tile -- plus rows of the table window
      const int wp = tid >> 5;
      if (wp) bulk_wait_read0();   // this thread's stores of the previous tile have read their buffers
      if (tid == 0) CK(0);
      constexpr unsigned BYTES = TILE * (unsigned)sizeof(double);
#if !PGPU_TAB_LDGSTS
      // the particle arrays first: their addresses do not depend on the tile's window
      if (wp == 0) {
        bulk_g2s(st + 0 * TILE, A.xo[0] + tbase, BYTES, bar);
        bulk_g2s(st + 1 * TILE, A.xo[1] + tbase, BYTES, bar);
      } else if (wp == 1) {
        if (!ALIAS) {   // first evaluation of a step: xbar == x_old, nothing to load
          bulk_g2s(st + 2 * TILE, A.xb[0] + tbase, BYTES, bar);
          bulk_g2s(st + 3 * TILE, A.xb[1] + tbase, BYTES, bar);
        }
      } else if (wp == 2) {
        bulk_g2s(st + 4 * TILE, A.uo[0] + tbase, BYTES, bar);
        bulk_g2s(st + 5 * TILE, A.uo[1] + tbase, BYTES, bar);
      } else {
        bulk_g2s(st + 6 * TILE, A.uo[2] + tbase, BYTES, bar);
        bulk_g2s(st + 7 * TILE, A.w + tbase, BYTES, bar);
      }
#endif
      if (tid == 0) CK(6);
      const int4 box = make_int4(nbx, nby, nbz, nbw);
      const unsigned dbytes = (unsigned)box.z * TD * (unsigned)sizeof(double);
      const unsigned nbytes = (unsigned)(box.z + 1) * TN * (unsigned)sizeof(double);
      const size_t cw0 = (size_t)(box.x - A.tlo[0]) + (size_t)(box.y - A.tlo[1]) * A.tn0;
      if (wp == 0) {
        sbox = box;
        if (box.z) {
          for (int r = 0; r < box.w; ++r)
            bulk_g2s(sdual + r * win_drow(box.w), A.tdual + (cw0 + (size_t)r * A.tn0) * TD, dbytes, bar);
          if (box.w == 3)   // fourth node row of the three-row window (warps 1..3 load node rows 0..2)
            bulk_g2s(snode + 3 * win_nstride(3), A.tnode + (cw0 + (size_t)3 * A.tn0) * TN, nbytes, bar);
        }
        // the one arrival of the phase, posted after this thread's copies (a copy that completes first only
        // drives the transaction count negative for a while; the phase cannot end before this arrival)
        const unsigned tabbytes = box.z ? box.w * dbytes + (box.w + 1) * nbytes : 0u;
#if PGPU_TAB_LDGSTS
        mbar_expect_tx(bar, tabbytes);
#else
        mbar_expect_tx(bar, (ALIAS ? NIN - 2 : NIN) * TILE * (unsigned)sizeof(double) + tabbytes);
#endif
      } else {
        const int r = wp - 1;   // node-record row of the window
        if (box.z && r <= box.w && r < 3)
          bulk_g2s(snode + r * win_nstride(box.w), A.tnode + (cw0 + (size_t)r * A.tn0) * TN, nbytes, bar);
      }
      if (tid == 0) CK(7);
      if (tile + (int)gridDim.x < ntiles) {
        fetch_box(tile + gridDim.x);
        if (A.prefetch) {
          const long nb = tbase + (long)gridDim.x * TILE;
          if (wp == 0) { l2_prefetch(A.xo[0] + nb, BYTES); l2_prefetch(A.xo[1] + nb, BYTES); }
          else if (wp == 1) { l2_prefetch(A.xb[0] + nb, BYTES); l2_prefetch(A.xb[1] + nb, BYTES); }
          else if (wp == 2) { l2_prefetch(A.uo[0] + nb, BYTES); l2_prefetch(A.uo[1] + nb, BYTES); }
          else { l2_prefetch(A.uo[2] + nb, BYTES); l2_prefetch(A.w + nb, BYTES); }
        }
      }
    }
    if (tid == 0) CK(1);
    __syncthreads();       // sbox written (and the previous tile's shared-memory reads are over)
#if PGPU_TAB_LDGSTS
    {
      // every thread copies 16 particle-array chunks of 16 bytes (LDGSTS) instead of eight 4 KB bulk copies
      const double *src[NIN] = {A.xo[0], A.xo[1], A.xb[0], A.xb[1], A.uo[0], A.uo[1], A.uo[2], A.w};
#pragma unroll
      for (int a = 0; a < NIN; ++a)
#pragma unroll
        for (int h = 0; h < TILE / (2 * BLOCK); ++h) {
          const int c2 = 2 * (tid + h * BLOCK);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(st + a * TILE + c2)),
                       "l"(src[a] + tbase + c2)
                       : "memory");
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    }
#endif
    mbar_wait(bar, (unsigned)(it & 1));
    if (tid == 0) CK(2);
    const int nvalid = (A.n - tbase) < TILE ? (int)(A.n - tbase) : TILE;
    TabWindow Wn;
    {
      const int4 box = sbox;
      Wn.dual = sdual;
      Wn.node = snode;
      Wn.i = box.x;
      Wn.j = box.y;
      Wn.ncol = box.z;
      Wn.nrow = box.w;
      Wn.drow = win_drow(box.w);
      Wn.nrow_stride = win_nstride(box.w);
    }

    // ---- phase 1: push ------------------------------------------------------------------
    unsigned defer_mask = 0;
    DualRec R;
    R.key = R.nkey = NOKEY;   // the window is restaged per tile, so the cached records do not outlive it
    R.nsel = 0;
    (void)R;
    if (PAIR) {
#pragma unroll 1
      for (int h = 0; h < TP / 2; ++h) {
        int kk[2];
        bool present[2], located[2];
        double xo[2][2], xb[2][2], uo[2][3], ub[2][3], dO[2][2];
        TabCell C[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int q = (2 * h + p + (lane >> 2)) & (TP - 1);
          kk[p] = tid * TP + q;
          present[p] = kk[p] < nvalid;
          located[p] = false;
          C[p].key = NOKEY;
          xo[p][0] = xo[p][1] = xb[p][0] = xb[p][1] = uo[p][0] = uo[p][1] = uo[p][2] = 0.0;
          dO[p][0] = dO[p][1] = 0.0;
          if (present[p]) {
            const int k = kk[p];
            xo[p][0] = st[0 * TILE + k], xo[p][1] = st[1 * TILE + k];
            xb[p][0] = st[2 * TILE + k], xb[p][1] = st[3 * TILE + k];
            uo[p][0] = st[4 * TILE + k], uo[p][1] = st[5 * TILE + k], uo[p][2] = st[6 * TILE + k];
            located[p] = locate_cell(A, Wn, xo[p], C[p], dO[p]);
          }
          ub[p][0] = ub[p][1] = ub[p][2] = 0.0;
        }
        bool ok[2] = {located[0], located[1]};
        if (!located[0]) C[0] = C[1];      // a missing particle is masked (ok = false) and only needs valid pointers
        if (!located[1]) C[1] = C[0];
        if (located[0] || located[1]) picard_same_cell<2>(A, C, xo, xb, uo, ub, dO, ok, apply, unconv);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int k = kk[p];
          unsigned key = NOKEY;
          if (present[p]) {
            if (ok[p]) {
              key = C[p].key;
              st[0 * TILE + k] = dO[p][0];
              st[1 * TILE + k] = dO[p][1];
              st[2 * TILE + k] = xb[p][0];
              st[3 * TILE + k] = xb[p][1];
              st[4 * TILE + k] = ub[p][0];
              st[5 * TILE + k] = ub[p][1];
              st[6 * TILE + k] = ub[p][2];
            } else {
              defer_mask |= 1u << (k - tid * TP);
            }
          }
          skey[k] = key;
        }
      }
    } else {
  #pragma unroll 1
      for (int qq = 0; qq < TP; ++qq) {
        const int q = (qq + (lane >> 2)) & (TP - 1);
        const int k = tid * TP + q;
        unsigned key = NOKEY;
        if (k < nvalid) {
          const double xo[2] = {st[0 * TILE + k], st[1 * TILE + k]};
          double xb[2];
          if (ALIAS) {
            xb[0] = xo[0];
            xb[1] = xo[1];
          } else {
            xb[0] = st[2 * TILE + k];
            xb[1] = st[3 * TILE + k];
          }
          const double uo[3] = {st[4 * TILE + k], st[5 * TILE + k], st[6 * TILE + k]};
          double ub[3] = {0.0, 0.0, 0.0}, dO[2], dB[2];
          if (push_tab<NODECACHE>(A, Wn, R, xo, xb, uo, ub, key, dO, dB, apply, unconv)) {
            st[0 * TILE + k] = dO[0];
            st[1 * TILE + k] = dO[1];
            st[2 * TILE + k] = xb[0];
            st[3 * TILE + k] = xb[1];
            st[4 * TILE + k] = ub[0];
            st[5 * TILE + k] = ub[1];
            st[6 * TILE + k] = ub[2];
          } else {
            // deferred: xbar stays as stored (the generic kernel restarts from it); the ubar slot
            // keeps u_old, which that kernel overwrites
            if (ALIAS) {
              st[2 * TILE + k] = xo[0];
              st[3 * TILE + k] = xo[1];
            }
            key = NOKEY;
            defer_mask |= 1u << q;
          }
        }
        skey[k] = key;
      }
    }

    if (tid == 0) CK(3);
    // ---- phase 2: deposit (same thread -> particle map: no block barrier needed) ---------
    if (DEP) {
      unsigned acc_key = NOKEY;
      double acc[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
#pragma unroll 1
      for (int qq = 0; qq < TP; ++qq) {
        const int q = (qq + (lane >> 2)) & (TP - 1);
        const int k = tid * TP + q;
        const unsigned key = skey[k];
        if (key == NOKEY) continue;
        if (key != acc_key) {
          if (acc_key != NOKEY) {
            flush_direct(A, acc_key, acc);
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
          }
          acc_key = key;
        }
        const double dO[2] = {st[0 * TILE + k], st[1 * TILE + k]};
        // d_bar from xbar and the dual-cell index (the same value as phase 1's up to rounding)
        const double dB[2] = {
            fma(st[2 * TILE + k] - A.le[0], A.rdx[0], -(double)((int)(key & 0xffffu) - 32767)),
            fma(st[3 * TILE + k] - A.le[1], A.rdx[1], -(double)((int)(key >> 16) - 32767))};
        const double ub[3] = {st[4 * TILE + k], st[5 * TILE + k], st[6 * TILE + k]};
        deposit_tab(A, dO, dB, ub, st[7 * TILE + k], acc);
      }
      const unsigned any = __ballot_sync(0xffffffffu, acc_key != NOKEY);
      if (any) {
        // runs = maximal stretches of consecutive lanes with the same key (any particle order
        // is handled: a key that reappears later simply forms another run)
        const unsigned prev = __shfl_up_sync(0xffffffffu, acc_key, 1);
        const bool head = (lane == 0) || (prev != acc_key);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const unsigned above = (lane == 31) ? 0u : (heads & (0xffffffffu << (lane + 1)));
        const int run_end = above ? (__ffs(above) - 1) : 32;
        const int run_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
        for (int sidx = 0; sidx < RSTEPS; ++sidx) {
          const int off = 1 << sidx;
          const double take = (lane + off < run_end) ? 1.0 : 0.0;
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) acc[j] = fma(__shfl_down_sync(0xffffffffu, acc[j], off), take, acc[j]);
        }
        // after RSTEPS steps lane l holds the sum over [l, min(l + 2^RSTEPS, run_end))
        if (acc_key != NOKEY && (((lane - run_start) & ((1 << RSTEPS) - 1)) == 0)) flush_direct(A, acc_key, acc);
      }
    }

    if (tid == 0) CK(4);
    // ---- results -> global ----------------------------------------------------------------
    if (nvalid == TILE) {
      fence_proxy_async();
      __syncthreads();
      if (tid == 0) CK(5);
      if (lane == 0 && tid) {
        constexpr unsigned BYTES = TILE * (unsigned)sizeof(double);
        const int wp = tid >> 5;
        if (wp == 1) {
          bulk_s2g(A.xbo[0] + tbase, st + 2 * TILE, BYTES);
          bulk_s2g(A.xbo[1] + tbase, st + 3 * TILE, BYTES);
        } else if (wp == 2) {
          bulk_s2g(A.ub[0] + tbase, st + 4 * TILE, BYTES);
          bulk_s2g(A.ub[1] + tbase, st + 5 * TILE, BYTES);
        } else {
          bulk_s2g(A.ub[2] + tbase, st + 6 * TILE, BYTES);
        }
        bulk_commit();
      }
    } else {
      // ragged last tile; like the bulk stores above this also writes the deferred particles' slots
      // (stored xbar, u_old) -- needed when the outputs are not the arrays the inputs came from
#pragma unroll 1
      for (int q = 0; q < TP; ++q) {
        const int k = tid * TP + q;
        if (k < nvalid) {
          A.xbo[0][tbase + k] = st[2 * TILE + k];
          A.xbo[1][tbase + k] = st[3 * TILE + k];
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
  if (lane == 0 && tid) bulk_wait0();
#ifdef PGPU_CLOCKS
  if (tid == 0 && (blockIdx.x % 97) == 5 && blockIdx.x < 600)
    printf("#blk %d tiles %d cycles/tile: wait_read %lld issue %lld (particle copies %lld, window copies %lld) mbar %lld phase1 %lld "
           "phase2 %lld endsync %lld\n",
           blockIdx.x, it, ck[0] / it, (ck[6] + ck[7] + ck[1]) / it, ck[6] / it, ck[7] / it, ck[2] / it, ck[3] / it, ck[4] / it,
           ck[5] / it);
#endif

  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  if (lane == 0) {
    if (apply) atomicAdd(&A.cnt->apply_its, (unsigned long long)apply);
    if (unconv) atomicAdd(&A.cnt->unconverged, (unsigned long long)unconv);
  }
}


// =============================================================================================
// v2 of the table-driven tile kernel (round 2).  ncu of v1 put the shared-memory data pipe at 66 % of its
// peak (above the fp64 pipe, 43 %, and the issue slots, 57 %): with v1's map of four CONSECUTIVE particles
// per thread a warp's 32 lanes sit in five or six different dual cells at any time, so every 128-bit table
// load is replayed ~3 times, and a third of the RED instructions run with one or two active lanes.
//   * Phase 1 (push) maps LANES to consecutive particles (particle = warp base + 32 r + lane): the lanes of
//     a warp then sit in one to three dual cells, the table loads are near-broadcasts, and every access to
//     the particle tile is a conflict-free 64-bit access.  The per-thread record cache of v1 is gone (its
//     hit rate would be nil with this map) and with it 32 registers.
//   * Phase 2 (deposit) keeps v1's map (four consecutive particles per thread: the register accumulators
//     need runs of equal keys per THREAD).  A warp only ever touches its own 128 particles of the tile, so
//     the phases are separated by __syncwarp(), not by a block barrier.
//   * ALIAS: the first evaluation of a step has xbar == x_old (updateOldParticlePositions recorded as an
//     alias): the two xbar arrays are then not loaded at all (-16 B per particle of HBM traffic).
// =============================================================================================
template <bool REC_PER_PASS>
__device__ __forceinline__ bool push_v2(const FastArgs &A, const TabWindow &Wn, const double (&xo)[2], double (&xb)[2],
                                        const double (&uo)[3], double (&ub)[3], unsigned &key, double (&dO)[2],
                                        unsigned &apply, unsigned &unconv) {
  int i0[2];
  bool ok = true;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double xr = __dsub_rn(xo[d], A.le[d]);
    i0[d] = floor_div_fast(__dsub_rn(xr, A.hdx[d]), A.dx[d], A.rdx[d]);
    dO[d] = fma(xr, A.rdx[d], -(double)(i0[d] + 1));
    if (i0[d] < A.i_lo[d] || i0[d] > A.i_hi[d]) ok = false;
  }
  if (!ok) return false;
  key = ((unsigned)(i0[1] + 32768) << 16) | (unsigned)(i0[0] + 32768);
  const double *td, *tn;
  int nrow;
  {
    const unsigned wi = (unsigned)(i0[0] - Wn.i), wj = (unsigned)(i0[1] - Wn.j);
    if (wi < (unsigned)Wn.ncol && wj < (unsigned)Wn.nrow) {
      td = Wn.dual + (wi * TD + wj * Wn.drow);
      tn = Wn.node + (wi * TN + wj * Wn.nrow_stride);
      nrow = Wn.nrow_stride;
    } else {
      const int cidx = (i0[0] - A.tlo[0]) + (i0[1] - A.tlo[1]) * A.tn0;
      td = A.tdual + (size_t)cidx * TD;
      tn = A.tnode + (size_t)cidx * TN;
      nrow = A.tn0 * TN;
    }
  }
  double2 x01, x23, x45, y01, y23, y45, z01, z23;
  if (!REC_PER_PASS) {
    x01 = ld2(td), x23 = ld2(td + 2), x45 = ld2(td + 4);       // Ex: e1 d1 | P0 Q0 | P2 Q2
    y01 = ld2(td + 6), y23 = ld2(td + 8), y45 = ld2(td + 10);   // Ey
    z01 = ld2(td + 12), z23 = ld2(td + 14);                      // Bz c0 c1 | c2 c3
  }
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
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      dxp0[d] = xb[d] - xo[d];
      dB[d] = fma(dxp0[d], A.rdx[d], dO[d]);
      dN[d] = fma(2.0, dB[d], -dO[d]);
    }
    if (!(hi_abs(dN[0]) < HI_HALF_BAND && hi_abs(dN[1]) < HI_HALF_BAND)) {
      // near (or past) a dual-cell face: the reference's own floor decides
      bool same = true;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const double xn = fma(2.0, xb[d], -xo[d]);
        const int in = floor_div_exact(__dsub_rn(__dsub_rn(xn, A.le[d]), A.hdx[d]), A.dx[d]);
        if (in != i0[d]) same = false;
      }
      if (!same) return false;
    }
    if (done) break;
    if (REC_PER_PASS) {
      asm volatile("" ::: "memory");   // keep the record loads inside the pass
      x01 = ld2(td), x23 = ld2(td + 2), x45 = ld2(td + 4);
      y01 = ld2(td + 6), y23 = ld2(td + 8), y45 = ld2(td + 10);
      z01 = ld2(td + 12), z23 = ld2(td + 14);
    }
    const double del0 = dB[0] + 0.5, del1 = dB[1] + 0.5;
    double E[3], B[3];
    {
      const double a0 = 0.5 - dN[0], b0 = 0.5 + dN[0], a1 = 0.5 - dN[1], b1 = 0.5 + dN[1];
      const double Wx0 = fma(a0, a0, pO[0][0]), Wx2 = fma(b0, b0, pO[0][1]);   // 4 W
      const double Wy0 = fma(a1, a1, pO[1][0]), Wy2 = fma(b1, b1, pO[1][1]);
      E[0] = fma(Wy0, fma(del0, x23.y, x23.x), fma(Wy2, fma(del0, x45.y, x45.x), fma(del0, x01.y, x01.x)));
      E[1] = fma(Wx0, fma(del1, y23.y, y23.x), fma(Wx2, fma(del1, y45.y, y45.x), fma(del1, y01.y, y01.x)));
    }
    {
      // nodal CIC at xbar: node pair (i0+s, i0+s+1), fraction f
      const bool sx = del0 >= 0.5, sy = del1 >= 0.5;
      const double fx = del0 + (sx ? -0.5 : 0.5), fy = del1 + (sy ? -0.5 : 0.5);
      const int ox = sx ? TN : 0, oy = sy ? nrow : 0;
      const double2 ez01 = ld2(tn + ox + oy), ez23 = ld2(tn + ox + oy + 2);
      const double2 bx01 = ld2(tn + ox + 4), bx23 = ld2(tn + ox + 6);
      const double2 by01 = ld2(tn + oy + 8), by23 = ld2(tn + oy + 10);
      E[2] = fma(fy, fma(fx, ez23.y, ez23.x), fma(fx, ez01.y, ez01.x));
      B[0] = fma(del1, fma(fx, bx23.y, bx23.x), fma(fx, bx01.y, bx01.x));
      B[1] = fma(fy, fma(del0, by23.y, by23.x), fma(del0, by01.y, by01.x));
      B[2] = fma(del1, fma(del0, z23.y, z23.x), fma(del0, z01.y, z01.x));
    }
    // Boris half step (PicSpeciesUtils.cpp:8-101)
    {
      const double vm0 = fma(A.alpha, E[0], uo[0]), vm1 = fma(A.alpha, E[1], uo[1]),
                   vm2 = fma(A.alpha, E[2], uo[2]);
      const double b0 = A.alpha * B[0], b1 = A.alpha * B[1], b2 = A.alpha * B[2];
      const double den = fma(b2, b2, fma(b1, b1, fma(b0, b0, 1.0)));
      const double p0 = fma(-vm2, b1, fma(vm1, b2, vm0));
      const double p1 = fma(-vm0, b2, fma(vm2, b0, vm1));
      const double p2 = fma(-vm1, b0, fma(vm0, b1, vm2));
      const double rden = rcp_ge1(den);
      const double r0 = b0 * rden, r1 = b1 * rden, r2 = b2 * rden;
      ub[0] = fma(-p2, r1, fma(p1, r2, vm0));
      ub[1] = fma(-p0, r2, fma(p2, r0, vm1));
      ub[2] = fma(-p1, r0, fma(p0, r1, vm2));
    }
    napply += 1;
    if (A.iter_max < 0) {  // advanceParticles (:1594-1612), part_order_swap == false
      if (A.iter_max == ITER_EXPLICIT) {
        // PIC_EM_EXPLICIT (pgpu_explicit_step): advanceVelocities(dt, byHalfDt = false) leaves u_new = 2 ubar - u_old,
        // advancePositionsExplicit(dt/2) moves with it, and setCurrentDensity(dt, true) deposits u_new over the orbit
        // x_old -> 2 x - x_old: the same closed forms with u_new in the place of ubar
        ub[0] = fma(2.0, ub[0], -uo[0]);
        ub[1] = fma(2.0, ub[1], -uo[1]);
        ub[2] = fma(2.0, ub[2], -uo[2]);
      }
      xb[0] = fma(ub[0], A.hdt, xo[0]);
      xb[1] = fma(ub[1], A.hdt, xo[1]);
      done = true;
      continue;
    }
    // stepNormTransfer (:658-733): |dxp0 - dxp| / dX against rtol, as |dxp0 - dxp| against rtol*dX
    const double dxp_0 = ub[0] * A.hdt, dxp_1 = ub[1] * A.hdt;
    const double e0 = fabs(dxp0[0] - dxp_0), e1 = fabs(dxp0[1] - dxp_1);
    if (iter == 0) {
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
      if (!(e0 >= A.tol[0]) && !(e1 >= A.tol[1])) done = true;
    } else {
      if (e0 < A.tol[0] && e1 < A.tol[1]) break;  // reverse pass: xbar unchanged, its orbit was checked above
      xb[0] = xo[0] + dxp_0;
      xb[1] = xo[1] + dxp_1;
    }
    if (!done && iter >= A.iter_max) {
      if (A.suborbit) return false;   // sub-orbit model: the generic kernel redoes it and lists it for the sub-orbit container
      nunconv = 1;
      done = true;
    }
    iter += 1;
  }
  apply += napply;
  unconv += nunconv;
  return true;
}

template <bool DEP, int RSTEPS, int MINB, bool ALIAS, bool REC_PER_PASS>
__global__ void __launch_bounds__(BLOCK, MINB) k_advance_cc1_2d_v2(const FastArgs A, int ntiles) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *st = reinterpret_cast<double *>(smem_raw);                 // [NTAB][TILE]
  double *sdual = st + NTAB * TILE;                                   // [2][WMAX][TD]
  double *snode = sdual + SDUAL;                                      // [3][WMAX+1][TN]
  unsigned *skey = reinterpret_cast<unsigned *>(snode + SNODE);      // [TILE]
  uint64_t *bar = reinterpret_cast<uint64_t *>(skey + TILE);
  __shared__ int4 sbox;
  const int tid = threadIdx.x, lane = tid & 31, wbase = (tid >> 5) * (32 * TP);
  constexpr unsigned NLOAD = ALIAS ? NIN - 2 : NIN;

  int nbx = 0, nby = 0, nbz = 0, nbw = 0;
  auto fetch_box = [&](int t) {
    const int4 b = __ldg(A.tile_box + t);
    nbx = b.x;
    nby = b.y;
    nbz = b.z;
    nbw = b.w;
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (lane == 0 && (int)blockIdx.x < ntiles) fetch_box(blockIdx.x);
  __syncthreads();

  unsigned apply = 0, unconv = 0;
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const long tbase = (long)tile * TILE;
    if (lane == 0) {
      const int wp = tid >> 5;
      if (wp) bulk_wait_read0();   // this thread's stores of the previous tile have read their buffers
      constexpr unsigned BYTES = TILE * (unsigned)sizeof(double);
      if (wp == 0) {
        bulk_g2s(st + 0 * TILE, A.xo[0] + tbase, BYTES, bar);
        bulk_g2s(st + 1 * TILE, A.xo[1] + tbase, BYTES, bar);
      } else if (wp == 1) {
        if (!ALIAS) {
          bulk_g2s(st + 2 * TILE, A.xb[0] + tbase, BYTES, bar);
          bulk_g2s(st + 3 * TILE, A.xb[1] + tbase, BYTES, bar);
        }
      } else if (wp == 2) {
        bulk_g2s(st + 4 * TILE, A.uo[0] + tbase, BYTES, bar);
        bulk_g2s(st + 5 * TILE, A.uo[1] + tbase, BYTES, bar);
      } else {
        bulk_g2s(st + 6 * TILE, A.uo[2] + tbase, BYTES, bar);
        bulk_g2s(st + 7 * TILE, A.w + tbase, BYTES, bar);
      }
      const int4 box = make_int4(nbx, nby, nbz, nbw);
      const unsigned dbytes = (unsigned)box.z * TD * (unsigned)sizeof(double);
      const unsigned nbytes = (unsigned)(box.z + 1) * TN * (unsigned)sizeof(double);
      const size_t cw0 = (size_t)(box.x - A.tlo[0]) + (size_t)(box.y - A.tlo[1]) * A.tn0;
      if (wp == 0) {
        sbox = box;
        if (box.z) {
          for (int r = 0; r < box.w; ++r)
            bulk_g2s(sdual + r * win_drow(box.w), A.tdual + (cw0 + (size_t)r * A.tn0) * TD, dbytes, bar);
          if (box.w == 3)   // fourth node row of the three-row window (warps 1..3 load node rows 0..2)
            bulk_g2s(snode + 3 * win_nstride(3), A.tnode + (cw0 + (size_t)3 * A.tn0) * TN, nbytes, bar);
        }
        const unsigned tabbytes = box.z ? box.w * dbytes + (box.w + 1) * nbytes : 0u;
        mbar_expect_tx(bar, NLOAD * TILE * (unsigned)sizeof(double) + tabbytes);
      } else {
        const int r = wp - 1;   // node-record row of the window
        if (box.z && r <= box.w && r < 3)
          bulk_g2s(snode + r * win_nstride(box.w), A.tnode + (cw0 + (size_t)r * A.tn0) * TN, nbytes, bar);
      }
      if (tile + (int)gridDim.x < ntiles) fetch_box(tile + gridDim.x);
    }
    __syncthreads();       // sbox written (and the previous tile's shared-memory reads are over)
    mbar_wait(bar, (unsigned)(it & 1));
    const int nvalid = (A.n - tbase) < TILE ? (int)(A.n - tbase) : TILE;
    TabWindow Wn;
    {
      const int4 box = sbox;
      Wn.dual = sdual;
      Wn.node = snode;
      Wn.i = box.x;
      Wn.j = box.y;
      Wn.ncol = box.z;
      Wn.nrow = box.w;
      Wn.drow = win_drow(box.w);
      Wn.nrow_stride = win_nstride(box.w);
    }

    // ---- phase 1: push; lane <-> consecutive particles ---------------------------------------
    unsigned defer_mask = 0;
#pragma unroll 1
    for (int r = 0; r < TP; ++r) {
      const int k = wbase + 32 * r + lane;
      unsigned key = NOKEY;
      if (k < nvalid) {
        const double xo[2] = {st[0 * TILE + k], st[1 * TILE + k]};
        double xb[2];
        if (ALIAS) {
          xb[0] = xo[0];
          xb[1] = xo[1];
        } else {
          xb[0] = st[2 * TILE + k];
          xb[1] = st[3 * TILE + k];
        }
        const double uo[3] = {st[4 * TILE + k], st[5 * TILE + k], st[6 * TILE + k]};
        double ub[3] = {0.0, 0.0, 0.0}, dO[2];
        if (push_v2<REC_PER_PASS>(A, Wn, xo, xb, uo, ub, key, dO, apply, unconv)) {
          st[0 * TILE + k] = dO[0];
          st[1 * TILE + k] = dO[1];
          st[2 * TILE + k] = xb[0];
          st[3 * TILE + k] = xb[1];
          st[4 * TILE + k] = ub[0];
          st[5 * TILE + k] = ub[1];
          st[6 * TILE + k] = ub[2];
        } else {
          // deferred: xbar stays as stored (the generic kernel restarts from it); the ubar slot keeps
          // u_old, which that kernel overwrites
          if (ALIAS) {
            st[2 * TILE + k] = xo[0];
            st[3 * TILE + k] = xo[1];
          }
          key = NOKEY;
          defer_mask |= 1u << r;
        }
      }
      skey[k] = key;
    }
    __syncwarp();

    // ---- phase 2: deposit; thread <-> four consecutive particles of its own warp -------------
    if (DEP) {
      unsigned acc_key = NOKEY;
      double acc[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
#pragma unroll 1
      for (int qq = 0; qq < TP; ++qq) {
        const int q = (qq + (lane >> 2)) & (TP - 1);
        const int k = tid * TP + q;
        const unsigned key = skey[k];
        if (key == NOKEY) continue;
        if (key != acc_key) {
          if (acc_key != NOKEY) {
            flush_direct(A, acc_key, acc);
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) acc[j] = 0.0;
          }
          acc_key = key;
        }
        const double dO[2] = {st[0 * TILE + k], st[1 * TILE + k]};
        const double dB[2] = {
            fma(st[2 * TILE + k] - A.le[0], A.rdx[0], -(double)((int)(key & 0xffffu) - 32767)),
            fma(st[3 * TILE + k] - A.le[1], A.rdx[1], -(double)((int)(key >> 16) - 32767))};
        const double ub[3] = {st[4 * TILE + k], st[5 * TILE + k], st[6 * TILE + k]};
        deposit_tab(A, dO, dB, ub, st[7 * TILE + k], acc);
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
          const double take = (lane + off < run_end) ? 1.0 : 0.0;
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) acc[j] = fma(__shfl_down_sync(0xffffffffu, acc[j], off), take, acc[j]);
        }
        if (acc_key != NOKEY && (((lane - run_start) & ((1 << RSTEPS) - 1)) == 0)) flush_direct(A, acc_key, acc);
      }
    }

    // ---- results -> global ----------------------------------------------------------------
    if (nvalid == TILE) {
      fence_proxy_async();
      __syncthreads();
      if (lane == 0 && tid) {
        constexpr unsigned BYTES = TILE * (unsigned)sizeof(double);
        const int wp = tid >> 5;
        if (wp == 1) {
          bulk_s2g(A.xbo[0] + tbase, st + 2 * TILE, BYTES);
          bulk_s2g(A.xbo[1] + tbase, st + 3 * TILE, BYTES);
        } else if (wp == 2) {
          bulk_s2g(A.ub[0] + tbase, st + 4 * TILE, BYTES);
          bulk_s2g(A.ub[1] + tbase, st + 5 * TILE, BYTES);
        } else {
          bulk_s2g(A.ub[2] + tbase, st + 6 * TILE, BYTES);
        }
        bulk_commit();
      }
    } else {
      __syncthreads();
#pragma unroll 1
      for (int q = 0; q < TP; ++q) {
        const int k = tid * TP + q;
        if (k < nvalid) {
          A.xbo[0][tbase + k] = st[2 * TILE + k];
          A.xbo[1][tbase + k] = st[3 * TILE + k];
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
      for (int r = 0; r < TP; ++r)
        if (defer_mask & (1u << r)) A.list[slot++] = (int)(tbase + wbase + 32 * r + lane);
    }
  }
  if (lane == 0 && tid) bulk_wait0();

  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  if (lane == 0) {
    if (apply) atomicAdd(&A.cnt->apply_its, (unsigned long long)apply);
    if (unconv) atomicAdd(&A.cnt->unconverged, (unsigned long long)unconv);
  }
}

// =============================================================================================
// Multi-segment orbits (the particles the tile kernel defers), table driven.
//
// CC1 splits the orbit x_old -> x_new at the faces of the half-shifted grid; per segment the in-plane weights are the
// single-segment closed forms with the SEGMENT's end points, times seg_factor = segment length / orbit length per direction
// (cc1_2d_interpolate_fields / cc1_2d_deposit_current, MeshInterpChargeConservingF.ChF:1885-1960, 1616-1709), and Ez, B
// and Jz use nodal CIC at the orbit's x_bar (no segments).  So a crossing particle needs the same table records as a
// single-segment one, of the two or three dual cells it visits: this kernel walks the reference's segments
// (walk_cc1_2d = the walker of cc_2d_inplane_visit, pgpu_device.cuh, decision for decision) and evaluates every segment
// from the global coefficient tables -- ~2.5x fewer instructions than the visitor kernel's per-point loads with bounds
// checks.  One thread per listed particle.  Anything it cannot do (a visited cell whose stencil leaves the arrays, more
// segments than ghosts + 1, a particle bound for the sub-orbit container) goes, untouched, to a second list for the visitor
// kernel, which keeps the reference's error behaviour.
// =============================================================================================
struct DeferPtrs {
  double *x[2];            // in: the stored x_bar (Picard starting guess); out: x_bar
  const double *xold[2];
  double *v[3];            // out: u_bar
  const double *vold[3];
  const double *w;
  const int *list;
  const unsigned *count;
  int *list2;              // rejects, for the visitor kernel
  unsigned *count2;
  int max_segments;        // ghosts + 1
};

// fn(ii, jj, x_start, x_end, seg_factor) per segment; returns 0, 1 (fn refused a cell) or 2 (too many segments)
template <class Fn>
__device__ __forceinline__ int walk_cc1_2d(const FastArgs &A, int max_segments, const double (&xo)[2],
                                           const double (&xb)[2], Fn &&fn) {
  double xpnew[2], dXp[2];
  int sign[2] = {1, 1}, index_old[2], cell_crossings[2], num_segments = 1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    xpnew[d] = 2.0 * xb[d] - xo[d];
    dXp[d] = xpnew[d] - xo[d];
  }
  // the reference divides (slope = dXp1 / dXp0, its inverse, one seg_factor per segment and direction); two reciprocals
  // serve them all within an ulp, which only matters for the tie of an orbit through a cell corner (either order of the
  // two zero-length-apart crossings gives the same sums)
  const double rd0 = __ddiv_rn(1.0, dXp[0]), rd1 = __ddiv_rn(1.0, dXp[1]);
  const double slope = dXp[1] * rd0;
  const double slope_inv = dXp[0] * rd1;
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    index_old[d] = floor_div_fast(__dsub_rn(__dsub_rn(xo[d], A.le[d]), A.hdx[d]), A.dx[d], A.rdx[d]);
    const int index_new = floor_div_fast(__dsub_rn(__dsub_rn(xpnew[d], A.le[d]), A.hdx[d]), A.dx[d], A.rdx[d]);
    if (index_new < index_old[d]) sign[d] = -1;
    cell_crossings[d] = abs(index_new - index_old[d]);
    num_segments += cell_crossings[d];
  }
  if (num_segments > max_segments) return 2;
  double Xcell[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    double cidx = __dadd_rn((double)index_old[d], 0.5 * (double)(1 - sign[d]));
    cidx = __dadd_rn(cidx, 0.5);
    Xcell[d] = A.le[d] + cidx * A.dx[d];
  }
  double xpold0[2] = {xo[0], xo[1]};
  double xpnew0[2] = {0.0, 0.0}, dXp_sub[2] = {0.0, 0.0};
  int ii_next = index_old[0], jj_next = index_old[1];
  for (int nn = 0; nn < num_segments; ++nn) {
    const int ii = ii_next, jj = jj_next;
    if (nn == num_segments - 1) {
      xpnew0[0] = xpnew[0];
      xpnew0[1] = xpnew[1];
      dXp_sub[0] = xpnew0[0] - xpold0[0];
      dXp_sub[1] = xpnew0[1] - xpold0[1];
    } else if (cell_crossings[0] == 0) {
      jj_next = jj + sign[1];
      Xcell[1] = Xcell[1] + (double)sign[1] * A.dx[1];
      xpnew0[1] = Xcell[1];
      dXp_sub[1] = xpnew0[1] - xpold0[1];
      dXp_sub[0] = __dmul_rn(slope_inv, dXp_sub[1]);
      xpnew0[0] = xpold0[0] + dXp_sub[0];
    } else if (cell_crossings[1] == 0) {
      ii_next = ii + sign[0];
      Xcell[0] = Xcell[0] + (double)sign[0] * A.dx[0];
      xpnew0[0] = Xcell[0];
      dXp_sub[0] = xpnew0[0] - xpold0[0];
      dXp_sub[1] = __dmul_rn(slope, dXp_sub[0]);
      xpnew0[1] = xpold0[1] + dXp_sub[1];
    } else {
      xpnew0[0] = Xcell[0] + (double)sign[0] * A.dx[0];
      xpnew0[1] = Xcell[1] + (double)sign[1] * A.dx[1];
      dXp_sub[0] = xpnew0[0] - xpold0[0];
      dXp_sub[1] = xpnew0[1] - xpold0[1];
      const double dXp_sub02 = __dmul_rn(slope_inv, dXp_sub[1]);
      if (fabs(dXp_sub[0]) < fabs(dXp_sub02)) {
        dXp_sub[1] = __dmul_rn(slope, dXp_sub[0]);
        xpnew0[1] = xpold0[1] + dXp_sub[1];
        Xcell[0] = xpnew0[0];
        ii_next = ii + sign[0];
        cell_crossings[0] -= 1;
      } else {
        dXp_sub[0] = __dmul_rn(slope_inv, dXp_sub[1]);
        xpnew0[0] = xpold0[0] + dXp_sub[0];
        Xcell[1] = xpnew0[1];
        jj_next = jj + sign[1];
        cell_crossings[1] -= 1;
      }
    }
    const double seg_factor[2] = {(dXp[0] != 0.0) ? dXp_sub[0] * rd0 : 1.0, (dXp[1] != 0.0) ? dXp_sub[1] * rd1 : 1.0};
    if (!fn(ii, jj, xpold0, xpnew0, seg_factor)) return 1;
    xpold0[0] = xpnew0[0];
    xpold0[1] = xpnew0[1];
  }
  return 0;
}

__device__ __forceinline__ bool tab_cell_ok(const FastArgs &A, int i, int j) {
  return i >= A.i_lo[0] && i <= A.i_hi[0] && j >= A.i_lo[1] && j <= A.i_hi[1];
}
// offsets of a point from the centre node of dual cell (i, j), in cells
__device__ __forceinline__ double cell_offset(const FastArgs &A, int d, double x, int idx) {
  return fma(__dsub_rn(x, A.le[d]), A.rdx[d], -(double)(idx + 1));
}

// the segments of the orbit gathered last: the deposit of a converged particle reuses them instead of walking again
constexpr int MAXSEG = 4;
struct SegStore {
  int n;
  int ii[MAXSEG], jj[MAXSEG];
  double dS[MAXSEG][2], dE[MAXSEG][2], sf[MAXSEG][2];
  int ib[2];        // dual cell of x_bar and its offsets there (nodal part)
  double dB[2];
};

// E and B at the orbit (x_old, x_bar); false = hand the particle to the visitor kernel
__device__ __forceinline__ bool gather_multiseg(const FastArgs &A, int max_segments, const double (&xo)[2],
                                                const double (&xb)[2], double (&E)[3], double (&B)[3], SegStore &S) {
  double e0 = 0.0, e1 = 0.0;
  S.n = 0;
  const int rc = walk_cc1_2d(A, max_segments < MAXSEG ? max_segments : MAXSEG, xo, xb,
                             [&](int ii, int jj, const double *xs, const double *xe, const double *sf) {
    if (!tab_cell_ok(A, ii, jj)) return false;
    const double *td = A.tdual + (size_t)((ii - A.tlo[0]) + (jj - A.tlo[1]) * A.tn0) * TD;
    const double dS0 = cell_offset(A, 0, xs[0], ii), dS1 = cell_offset(A, 1, xs[1], jj);
    const double dE0 = cell_offset(A, 0, xe[0], ii), dE1 = cell_offset(A, 1, xe[1], jj);
    {
      const int k = S.n++;
      S.ii[k] = ii, S.jj[k] = jj;
      S.dS[k][0] = dS0, S.dS[k][1] = dS1, S.dE[k][0] = dE0, S.dE[k][1] = dE1;
      S.sf[k][0] = sf[0], S.sf[k][1] = sf[1];
    }
    const double del0 = 0.5 * (dS0 + dE0) + 0.5, del1 = 0.5 * (dS1 + dE1) + 0.5;
    const double a0 = 0.5 - dE0, b0 = 0.5 + dE0, c0 = 0.5 - dS0, g0 = 0.5 + dS0;
    const double a1 = 0.5 - dE1, b1 = 0.5 + dE1, c1 = 0.5 - dS1, g1 = 0.5 + dS1;
    const double Wx0 = fma(a0, a0, c0 * c0), Wx2 = fma(b0, b0, g0 * g0);   // 4 W
    const double Wy0 = fma(a1, a1, c1 * c1), Wy2 = fma(b1, b1, g1 * g1);
    const double2 x01 = ld2(td), x23 = ld2(td + 2), x45 = ld2(td + 4);
    const double2 y01 = ld2(td + 6), y23 = ld2(td + 8), y45 = ld2(td + 10);
    const double ex = fma(Wy0, fma(del0, x23.y, x23.x), fma(Wy2, fma(del0, x45.y, x45.x), fma(del0, x01.y, x01.x)));
    const double ey = fma(Wx0, fma(del1, y23.y, y23.x), fma(Wx2, fma(del1, y45.y, y45.x), fma(del1, y01.y, y01.x)));
    e0 = fma(sf[0], ex, e0);
    e1 = fma(sf[1], ey, e1);
    return true;
  });
  if (rc) return false;
  E[0] = e0;
  E[1] = e1;
  // nodal CIC at x_bar, from the records of x_bar's own dual cell
  int ib[2];
  double dB[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const double xr = __dsub_rn(xb[d], A.le[d]);
    ib[d] = floor_div_fast(__dsub_rn(xr, A.hdx[d]), A.dx[d], A.rdx[d]);
    dB[d] = fma(xr, A.rdx[d], -(double)(ib[d] + 1));
  }
  if (!tab_cell_ok(A, ib[0], ib[1])) return false;
  S.ib[0] = ib[0], S.ib[1] = ib[1], S.dB[0] = dB[0], S.dB[1] = dB[1];
  const size_t cidx = (size_t)((ib[0] - A.tlo[0]) + (ib[1] - A.tlo[1]) * A.tn0);
  const double *td = A.tdual + cidx * TD, *tn = A.tnode + cidx * TN;
  const int nrow = A.tn0 * TN;
  const double del0 = dB[0] + 0.5, del1 = dB[1] + 0.5;
  const bool sx = del0 >= 0.5, sy = del1 >= 0.5;
  const double fx = del0 + (sx ? -0.5 : 0.5), fy = del1 + (sy ? -0.5 : 0.5);
  const int ox = sx ? TN : 0, oy = sy ? nrow : 0;
  const double2 ez01 = ld2(tn + ox + oy), ez23 = ld2(tn + ox + oy + 2);
  const double2 bx01 = ld2(tn + ox + 4), bx23 = ld2(tn + ox + 6);
  const double2 by01 = ld2(tn + oy + 8), by23 = ld2(tn + oy + 10);
  const double2 z01 = ld2(td + 12), z23 = ld2(td + 14);
  E[2] = fma(fy, fma(fx, ez23.y, ez23.x), fma(fx, ez01.y, ez01.x));
  B[0] = fma(del1, fma(fx, bx23.y, bx23.x), fma(fx, bx01.y, bx01.x));
  B[1] = fma(fy, fma(del0, by23.y, by23.x), fma(del0, by01.y, by01.x));
  B[2] = fma(del1, fma(del0, z23.y, z23.x), fma(del0, z01.y, z01.x));
  return true;
}

__device__ __forceinline__ void boris_half(const FastArgs &A, const double (&uo)[3], const double (&E)[3],
                                           const double (&B)[3], double (&ub)[3]) {
  const double vm0 = fma(A.alpha, E[0], uo[0]), vm1 = fma(A.alpha, E[1], uo[1]), vm2 = fma(A.alpha, E[2], uo[2]);
  const double b0 = A.alpha * B[0], b1 = A.alpha * B[1], b2 = A.alpha * B[2];
  const double den = fma(b2, b2, fma(b1, b1, fma(b0, b0, 1.0)));
  const double p0 = fma(-vm2, b1, fma(vm1, b2, vm0));
  const double p1 = fma(-vm0, b2, fma(vm2, b0, vm1));
  const double p2 = fma(-vm1, b0, fma(vm0, b1, vm2));
  const double rden = rcp_ge1(den);
  const double r0 = b0 * rden, r1 = b1 * rden, r2 = b2 * rden;
  ub[0] = fma(-p2, r1, fma(p1, r2, vm0));
  ub[1] = fma(-p0, r2, fma(p2, r0, vm1));
  ub[2] = fma(-p1, r0, fma(p0, r1, vm2));
}

#ifndef PGPU_MS_MINB
#define PGPU_MS_MINB 4
#endif
template <bool DEP>
__global__ void __launch_bounds__(128, PGPU_MS_MINB) k_advance_cc1_2d_multiseg(const FastArgs A, const DeferPtrs P) {
  const long total = (long)*P.count;
  const long stride = (long)gridDim.x * blockDim.x;
  unsigned apply = 0, unconv = 0;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long i = (long)P.list[t];
    const double xo[2] = {P.xold[0][i], P.xold[1][i]};
    double xb[2] = {P.x[0][i], P.x[1][i]};
    const double uo[3] = {P.vold[0][i], P.vold[1][i], P.vold[2][i]};
    double ub[3] = {uo[0], uo[1], uo[2]}, E[3], B[3];
    unsigned napply = 0;
    bool ok = true, left_unconverged = false;
    SegStore S;
    bool segs_current = false;   // S describes the orbit (x_old, x_bar as it stands)
    if (A.iter_max < 0) {   // advanceParticles (:1594-1612), part_order_swap == false
      ok = gather_multiseg(A, P.max_segments, xo, xb, E, B, S);
      if (ok) {
        boris_half(A, uo, E, B, ub);
        napply = 1;
        xb[0] = fma(ub[0], A.hdt, xo[0]);
        xb[1] = fma(ub[1], A.hdt, xo[1]);
      }
    } else {                // advanceParticlesIteratively (:1614-1716) with stepNormTransfer (:658-733)
      int iter = 0;
      while (true) {
        if (!gather_multiseg(A, P.max_segments, xo, xb, E, B, S)) {
          ok = false;
          break;
        }
        segs_current = true;
        boris_half(A, uo, E, B, ub);
        napply += 1;
        const double dxp_0 = ub[0] * A.hdt, dxp_1 = ub[1] * A.hdt;
        const double e0 = fabs((xb[0] - xo[0]) - dxp_0), e1 = fabs((xb[1] - xo[1]) - dxp_1);
        bool converged;
        if (iter == 0) {
          xb[0] = xo[0] + dxp_0;
          xb[1] = xo[1] + dxp_1;
          segs_current = false;
          converged = !(e0 >= A.tol[0]) && !(e1 >= A.tol[1]);
        } else {
          converged = e0 < A.tol[0] && e1 < A.tol[1];
          if (!converged) {
            xb[0] = xo[0] + dxp_0;
            xb[1] = xo[1] + dxp_1;
            segs_current = false;
          }
        }
        if (converged) break;
        if (iter >= A.iter_max) {
          left_unconverged = true;
          break;
        }
        iter += 1;
      }
    }
    if (ok && left_unconverged && A.suborbit) ok = false;   // the visitor kernel lists it for the sub-orbit container
    if (ok && DEP && !segs_current) {
      // the orbit moved after its last gather (converged in the first pass, or left unconverged): its segments, and
      // that every cell of them is depositable, before the first atomic goes out
      double Ed[3], Bd[3];
      ok = gather_multiseg(A, P.max_segments, xo, xb, Ed, Bd, S);
    }
    if (!ok) {
      P.list2[atomicAdd(P.count2, 1u)] = (int)i;
      continue;
    }
    apply += napply;
    unconv += left_unconverged ? 1u : 0u;
    P.x[0][i] = xb[0];
    P.x[1][i] = xb[1];
    P.v[0][i] = ub[0];
    P.v[1][i] = ub[1];
    P.v[2][i] = ub[2];
    if (DEP) {
      const double wp = P.w[i];
      for (int k = 0; k < S.n; ++k) {
        const int ii = S.ii[k], jj = S.jj[k];
        const double dO[2] = {S.dS[k][0], S.dS[k][1]};
        const double dM[2] = {0.5 * (S.dS[k][0] + S.dE[k][0]), 0.5 * (S.dS[k][1] + S.dE[k][1])};
        const double us[3] = {ub[0] * S.sf[k][0], ub[1] * S.sf[k][1], 0.0};
        double c[NSLOT];
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) c[j] = 0.0;
        deposit_tab(A, dO, dM, us, wp, c);
        {
          double *p0 = A.J[0] + (ii + jj * A.jn0[0]);
          double *p1 = p0 + A.jn0[0];
          double *p2 = p1 + A.jn0[0];
          atomicAdd(p0, c[0]), atomicAdd(p0 + 1, c[1]);
          atomicAdd(p1, c[2]), atomicAdd(p1 + 1, c[3]);
          atomicAdd(p2, c[4]), atomicAdd(p2 + 1, c[5]);
        }
        {
          double *p0 = A.J[1] + (ii + jj * A.jn0[1]);
          double *p1 = p0 + A.jn0[1];
#pragma unroll
          for (int a = 0; a < 3; ++a) atomicAdd(p0 + a, c[6 + a]), atomicAdd(p1 + a, c[9 + a]);
        }
      }
      // Jz: nodal CIC at x_bar (cc1_2d_deposit_current :1720-1740)
      const double dB[2] = {S.dB[0], S.dB[1]};
      const double uz[3] = {0.0, 0.0, ub[2]};
      double c[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) c[j] = 0.0;
      deposit_tab(A, dB, dB, uz, wp, c);
      double *p0 = A.J[2] + (S.ib[0] + S.ib[1] * A.jn0[2]);
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const double val = c[12 + a + 3 * b];
          if (val != 0.0) atomicAdd(p0 + a + b * A.jn0[2], val);
        }
    }
  }
  apply = __reduce_add_sync(0xffffffffu, apply);
  unconv = __reduce_add_sync(0xffffffffu, unconv);
  if ((threadIdx.x & 31) == 0) {
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
  if (c.exact || g->desc.D != 2 || s->desc.interp_E != CC1 || s->desc.relativistic) return 0;
  if (deposit && s->desc.interp_J != CC1) return 0;
  if (prm.iter_max < 0 && prm.order_swap) return 0;
  if (prm.explicit_step && (c.cc1_tma != 3 || c.cc1_version < 2 || !deposit || prm.iter_max >= 0)) return 0;
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
  // updateOldParticle* recorded as an alias (xold == x, vold == v): the table kernel reads the old state
  // from x / v and writes xbar / ubar into the stale old arrays; the pointers are swapped after the launch
  if (c.cc1_tma != 3 && materialize_old(s)) return PGPU_ERR_CUDA;
  const bool xa = s->xold_alias, va = s->vold_alias;
  for (int d = 0; d < 2; ++d) {
    A.xo[d] = xa ? s->x[d] : s->xold[d];
    A.xb[d] = s->x[d];
    A.xbo[d] = xa ? s->xold[d] : s->x[d];
    A.le[d] = g->geo.le[d];
    A.dx[d] = g->geo.dx[d];
    A.rdx[d] = g->geo.rdx[d];
    A.hdx[d] = 0.5 * g->geo.dx[d];
  }
  for (int k = 0; k < 3; ++k) {
    A.uo[k] = va ? s->v[k] : s->vold[k];
    A.ub[k] = va ? s->vold[k] : s->v[k];
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
  if (fields_wait(g)) return PGPU_ERR_CUDA;
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
  A.iter_max = prm.explicit_step ? ITER_EXPLICIT : prm.iter_max;
  A.suborbit = prm.suborbit;
  A.list = s->defer_list;
  A.list_count = s->defer_count;
  A.cnt = c.d_counters;
  A.tdual = A.tnode = nullptr;
  A.prefetch = c.cc1_prefetch;
  const bool tile_ok = s->cap >= (size_t)(((s->n + TILE - 1) / TILE) * TILE);
  if (!tile_ok) return 0;
  const int ntiles = (int)((s->n + TILE - 1) / TILE);
  const int grid = std::min(ntiles, c.sm_count * 4 * c.cc1_waves);
  if (c.cc1_tma == 3) {
    // coefficient tables of the selected field slot, rebuilt when its E/B arrays changed
    pgpu_grid_s *gm = s->grid;
    const int slot = gm->cur_slot;
    const int n0 = hi[0] - lo[0] + 2, n1 = hi[1] - lo[1] + 2;   // dual range + 1 (node records)
    if (n0 < 2 || n1 < 2) return 0;
    const size_t ncell = (size_t)n0 * n1;
    if (gm->tab_cells != ncell) {
      for (int k = 0; k < 4; ++k) {
        if (gm->tab_dual[k]) cudaFree(gm->tab_dual[k]);
        if (gm->tab_node[k]) cudaFree(gm->tab_node[k]);
        gm->tab_dual[k] = gm->tab_node[k] = nullptr;
        gm->tab_dirty[k] = true;
      }
      gm->tab_cells = ncell;
    }
    if (!gm->tab_dual[slot]) {
      PGPU_CUDA(cudaMalloc(&gm->tab_dual[slot], ncell * TD * sizeof(double)));
      PGPU_CUDA(cudaMalloc(&gm->tab_node[slot], ncell * TN * sizeof(double)));
      gm->tab_dirty[slot] = true;
    }
    if (gm->tab_dirty[slot]) {
      TabArgs T;
      for (int k = 0; k < 6; ++k) {
        T.F[k] = A.F[k];
        T.fn0[k] = A.fn0[k];
      }
      T.lo[0] = lo[0];
      T.lo[1] = lo[1];
      T.n0 = n0;
      T.n1 = n1;
      T.dual = gm->tab_dual[slot];
      T.node = gm->tab_node[slot];
      KTimer t("build_tables");
      k_build_tables<<<(unsigned)((ncell + 127) / 128), 128, 0, c.stream>>>(T);
      gm->tab_dirty[slot] = false;
    }
    A.tdual = gm->tab_dual[slot];
    A.tnode = gm->tab_node[slot];
    A.tlo[0] = lo[0];
    A.tlo[1] = lo[1];
    A.tn0 = n0;
    A.tol[0] = prm.rtol * A.dx[0];
    A.tol[1] = prm.rtol * A.dx[1];
    if (s->tile_box_cap < (size_t)ntiles) {
      if (s->tile_box) cudaFree(s->tile_box);
      s->tile_box_cap = (size_t)(s->cap + TILE - 1) / TILE;
      PGPU_CUDA(cudaMalloc(&s->tile_box, s->tile_box_cap * sizeof(int4)));
    }
    A.tile_box = (const int4 *)s->tile_box;
    {
      KTimer t("tile_boxes");
      k_tile_boxes<<<(unsigned)((ntiles + 127) / 128), 128, 0, c.stream>>>(A, ntiles, (int4 *)s->tile_box);
    }
    KTimer t(prm.explicit_step ? "explicit_step_cc1" : (deposit ? "advance_cc1_fused" : "advance_cc1"));
#define PGPU_TAB_LAUNCH(DEPV, RS, MB, PR)                                                                 \
  do {                                                                                                    \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tab<DEPV, RS, MB, PR>,                                \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB_SMEM));        \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tab<DEPV, RS, MB, PR>,                                \
                                   cudaFuncAttributePreferredSharedMemoryCarveout, 100));                \
    k_advance_cc1_2d_tab<DEPV, RS, MB, PR><<<gridm, BLOCK, TAB_SMEM, c.stream>>>(A, ntiles);              \
  } while (0)
    const int mb = c.cc1_minblocks == 5 ? 5 : (c.cc1_minblocks == 3 ? 3 : 4);
    const int gridm = std::min(ntiles, c.sm_count * mb * c.cc1_waves);
    if (c.cc1_version >= 2) {
#define PGPU_V2_LAUNCH(DEPV, RS, MB, AL, RP)                                                              \
  do {                                                                                                    \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_v2<DEPV, RS, MB, AL, RP>,                             \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB_SMEM));        \
    /* the driver's default carve-out (196 KB) holds four 45.7 KB blocks; five need the full 228 KB */   \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_v2<DEPV, RS, MB, AL, RP>,                             \
                                   cudaFuncAttributePreferredSharedMemoryCarveout, 100));                \
    k_advance_cc1_2d_v2<DEPV, RS, MB, AL, RP><<<gridm, BLOCK, TAB_SMEM, c.stream>>>(A, ntiles);           \
  } while (0)
#define PGPU_V2_PICK(DEPV, RS)                                                                            \
  do {                                                                                                    \
    const int sel = (mb == 5 ? 4 : 0) | (xa ? 2 : 0) | (c.cc1_rec_per_pass ? 1 : 0);                      \
    switch (sel) {                                                                                        \
      case 0: PGPU_V2_LAUNCH(DEPV, RS, 4, false, false); break;                                           \
      case 1: PGPU_V2_LAUNCH(DEPV, RS, 4, false, true); break;                                            \
      case 2: PGPU_V2_LAUNCH(DEPV, RS, 4, true, false); break;                                            \
      case 3: PGPU_V2_LAUNCH(DEPV, RS, 4, true, true); break;                                             \
      case 4: PGPU_V2_LAUNCH(DEPV, RS, 5, false, false); break;                                           \
      case 5: PGPU_V2_LAUNCH(DEPV, RS, 5, false, true); break;                                            \
      case 6: PGPU_V2_LAUNCH(DEPV, RS, 5, true, false); break;                                            \
      default: PGPU_V2_LAUNCH(DEPV, RS, 5, true, true); break;                                            \
    }                                                                                                     \
  } while (0)
      if (deposit) PGPU_V2_PICK(true, 3);
      else PGPU_V2_PICK(false, 0);
      t.stop();
      if (xa)
        for (int d = 0; d < 2; ++d) std::swap(s->x[d], s->xold[d]);
      if (va)
        for (int k = 0; k < 3; ++k) std::swap(s->v[k], s->vold[k]);
      s->xold_alias = s->vold_alias = false;
      s->defer_count_first = s->defer_count;
      if (c.cc1_multiseg && !prm.explicit_step) {
        // the listed (crossing) particles from the tables; what it cannot do lands on the second list, which becomes
        // THE list of the visitor kernel
        if (!s->defer_list2 || s->defer_cap2 < s->defer_cap) {
          if (s->defer_list2) cudaFree(s->defer_list2);
          if (!s->defer_count2) PGPU_CUDA(cudaMalloc(&s->defer_count2, sizeof(unsigned)));
          PGPU_CUDA(cudaMalloc(&s->defer_list2, s->defer_cap * sizeof(int)));
          s->defer_cap2 = s->defer_cap;
        }
        PGPU_CUDA(cudaMemsetAsync(s->defer_count2, 0, sizeof(unsigned), c.stream));
        DeferPtrs P;
        for (int d = 0; d < 2; ++d) {
          P.x[d] = s->x[d];
          P.xold[d] = s->xold[d];
        }
        for (int k = 0; k < 3; ++k) {
          P.v[k] = s->v[k];
          P.vold[k] = s->vold[k];
        }
        P.w = s->w;
        P.list = s->defer_list;
        P.count = s->defer_count;
        P.list2 = s->defer_list2;
        P.count2 = s->defer_count2;
        P.max_segments = g->desc.nghost + 1;
        {
          KTimer t2(deposit ? "advance_multiseg_fused" : "advance_multiseg");
          const unsigned gridd = (unsigned)(c.sm_count * 8);
          if (deposit) k_advance_cc1_2d_multiseg<true><<<gridd, 128, 0, c.stream>>>(A, P);
          else k_advance_cc1_2d_multiseg<false><<<gridd, 128, 0, c.stream>>>(A, P);
        }
        std::swap(s->defer_list, s->defer_list2);
        std::swap(s->defer_count, s->defer_count2);
        std::swap(s->defer_cap, s->defer_cap2);
      }
      return 1;
    }
    if (c.cc1_pair && !prm.suborbit) {
      // two particles of a dual cell in lockstep through the Picard loop (PGPU_CC1_PAIR=1)
      if (!deposit) {
        if (mb == 3) PGPU_TAB_LAUNCH(false, 0, 3, true);
        else PGPU_TAB_LAUNCH(false, 0, 4, true);
      } else {
        if (mb == 3) PGPU_TAB_LAUNCH(true, 3, 3, true);
        else PGPU_TAB_LAUNCH(true, 3, 4, true);
      }
    } else if (!deposit) {
      if (mb == 4) PGPU_TAB_LAUNCH(false, 0, 4, false);
      else PGPU_TAB_LAUNCH(false, 0, 5, false);
    } else if (c.cc1_rsteps <= 2) {
      if (mb == 4) PGPU_TAB_LAUNCH(true, 2, 4, false);
      else PGPU_TAB_LAUNCH(true, 2, 5, false);
    } else if (c.cc1_nodecache || xa) {
#define PGPU_TAB_LAUNCH6(MB, NC, AL)                                                                      \
  do {                                                                                                    \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tab<true, 3, MB, false, NC, AL>,                      \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TAB_SMEM));        \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tab<true, 3, MB, false, NC, AL>,                      \
                                   cudaFuncAttributePreferredSharedMemoryCarveout, 100));                \
    k_advance_cc1_2d_tab<true, 3, MB, false, NC, AL><<<gridm, BLOCK, TAB_SMEM, c.stream>>>(A, ntiles);    \
  } while (0)
      const int sel = (mb == 3 ? 0 : (mb == 4 ? 4 : 8)) | (c.cc1_nodecache ? 2 : 0) | (xa ? 1 : 0);
      switch (sel) {
        case 0: PGPU_TAB_LAUNCH6(3, false, false); break;
        case 1: PGPU_TAB_LAUNCH6(3, false, true); break;
        case 2: PGPU_TAB_LAUNCH6(3, true, false); break;
        case 3: PGPU_TAB_LAUNCH6(3, true, true); break;
        case 4: PGPU_TAB_LAUNCH6(4, false, false); break;
        case 5: PGPU_TAB_LAUNCH6(4, false, true); break;
        case 6: PGPU_TAB_LAUNCH6(4, true, false); break;
        case 7: PGPU_TAB_LAUNCH6(4, true, true); break;
        case 8: PGPU_TAB_LAUNCH6(5, false, false); break;
        case 9: PGPU_TAB_LAUNCH6(5, false, true); break;
        case 10: PGPU_TAB_LAUNCH6(5, true, false); break;
        default: PGPU_TAB_LAUNCH6(5, true, true); break;
      }
    } else {
      if (mb == 4) PGPU_TAB_LAUNCH(true, 3, 4, false);
      else if (mb == 3) PGPU_TAB_LAUNCH(true, 3, 3, false);
      else PGPU_TAB_LAUNCH(true, 3, 5, false);
    }
    if (xa)
      for (int d = 0; d < 2; ++d) std::swap(s->x[d], s->xold[d]);
    if (va)
      for (int k = 0; k < 3; ++k) std::swap(s->v[k], s->vold[k]);
    s->xold_alias = s->vold_alias = false;
    return 1;
  }
  KTimer t(deposit ? "advance_cc1_fused" : "advance_cc1");
#define PGPU_TILE_LAUNCH(DEPV, RS)                                                                        \
  do {                                                                                                    \
    PGPU_CUDA(cudaFuncSetAttribute(k_advance_cc1_2d_tile<DEPV, RS>,                                       \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE2_SMEM));       \
    k_advance_cc1_2d_tile<DEPV, RS><<<grid, BLOCK, TILE2_SMEM, c.stream>>>(A, ntiles);                    \
  } while (0)
  if (!deposit) PGPU_TILE_LAUNCH(false, 0);
  else if (c.cc1_rsteps <= 2) PGPU_TILE_LAUNCH(true, 2);
  else if (c.cc1_rsteps == 3) PGPU_TILE_LAUNCH(true, 3);
  else PGPU_TILE_LAUNCH(true, 4);
  return 1;
}

}  // namespace pgpu
