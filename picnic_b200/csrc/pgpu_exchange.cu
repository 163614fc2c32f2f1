// pgpu_exchange.cu -- device side of the multi-box exchanges.
//
// One process per GPU owns one Chombo box.  The collectives themselves are issued by the host
// plumbing (picnic_b200/halo.py: torch.distributed, NCCL over NVLink) on DEVICE buffers; this
// file provides what has to touch the particle and grid arrays:
//   * pack / unpack(+add) of an index box of a grid array  -- the ghost ADD-exchange of J
//     (LevelData::exchange with an add op in PicSpeciesInterface::finalizeSettingJ,
//     PicSpeciesInterface.cpp:766-772) and the ghost refresh that follows;
//   * outgoing-particle migration: ParticleData::gatherOutcast / remapOutcast
//     (src/particle_tools/ParticleDataI.H:405-547): a particle belongs to the box whose cells
//     contain it, box index = floor((x - origin) / (dx * boxSize)) per direction.  Leavers are
//     packed into wire records, the holes they leave are filled from the tail of the arrays,
//     and arrivals are appended.
#include "pgpu_internal.h"

#include <vector>

namespace pgpu {

static inline unsigned nb(long n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

// ---- grid arrays ------------------------------------------------------------------------
__global__ void k_fab_pack(FabView f, int lo0, int lo1, int m0, int m1, double *buf) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)m0 * m1) return;
  const int a = (int)(t % m0), b = (int)(t / m0);
  buf[t] = f.p[(long)(lo0 + a - f.lo0) + (long)(lo1 + b - f.lo1) * f.n0];
}
__global__ void k_fab_unpack(FabView f, int lo0, int lo1, int m0, int m1, const double *buf, int add) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)m0 * m1) return;
  const int a = (int)(t % m0), b = (int)(t / m0);
  double *p = f.p + ((long)(lo0 + a - f.lo0) + (long)(lo1 + b - f.lo1) * f.n0);
  *p = add ? __dadd_rn(*p, buf[t]) : buf[t];
}

static const DeviceFab *pick_fab(pgpu_grid_t g, int kind, int comp) {
  if (!g) return nullptr;
  if (kind == PGPU_FAB_JTOTAL && comp >= 0 && comp < 3) return &g->jtot[comp];
  if (kind == PGPU_FAB_FIELD && comp >= 0 && comp < 6) {
    fields_wait(g);
    return &g->field[comp];
  }
  return nullptr;
}

static int check_box(const DeviceFab &f, int D, const int *lo, const int *hi, int *m) {
  m[0] = m[1] = 1;
  for (int d = 0; d < D; ++d) {
    if (lo[d] < f.lo[d] || hi[d] > f.hi[d] || hi[d] < lo[d]) {
      set_error("index box [%d:%d] in dir %d is not inside the device array [%d:%d]", lo[d], hi[d], d, f.lo[d],
                f.hi[d]);
      return PGPU_ERR_ARG;
    }
    m[d] = hi[d] - lo[d] + 1;
  }
  return 0;
}

// ---- migration --------------------------------------------------------------------------
struct MigGeo {
  int D;
  double le[2], boxlen[2];   // origin, dx * boxSize
  int my[2], nbx[2];         // this box's coordinate and the number of boxes per direction
  int periodic[2];
};

// direction code of the owning box relative to this one: (d0+1) + 3 (d1+1), 4 = stays;
// 9 = left the decomposition (a host boundary condition has to deal with it)
__device__ __forceinline__ int owner_code(const MigGeo &G, double x0, double x1) {
  int code = 0, mul = 1;
  const double xs[2] = {x0, x1};
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    int diff = 0;
    if (d < G.D) {
      const int b = __double2int_rd(__ddiv_rn(__dsub_rn(xs[d], G.le[d]), G.boxlen[d]));
      diff = b - G.my[d];
      if (G.periodic[d]) {
        if (diff > 1) diff -= G.nbx[d];
        if (diff < -1) diff += G.nbx[d];
      }
      if (b < 0 || b >= G.nbx[d] || diff < -1 || diff > 1) return 9;
    }
    code += (diff + 1) * mul;
    mul *= 3;
  }
  return code;
}

struct MigCounters {
  unsigned count[10];   // per code; [9] = lost
  unsigned nleave;      // list length
  unsigned nhole, nmove;
  unsigned cursor[9];
};

__global__ void k_mark_leavers(const double *x0, const double *x1, long n, MigGeo G, int *dead, int *list,
                               MigCounters *mc) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int code = owner_code(G, x0[i], G.D == 2 ? x1[i] : 0.0);
  dead[i] = code == 4 ? 0 : 1 + code;
  if (code != 4) {
    atomicAdd(&mc->count[code], 1u);
    const unsigned slot = atomicAdd(&mc->nleave, 1u);
    list[slot] = (int)i;
  }
}

struct WirePtrs {
  double *x[2], *xold[2], *v[3], *vold[3], *w;
  uint64_t *id;
  int D;
};
__device__ __forceinline__ int wire_len(int D) { return 2 * D + 8; }

__device__ __forceinline__ void wire_out(const WirePtrs &P, long i, double *r) {
  int k = 0;
  for (int d = 0; d < P.D; ++d) r[k++] = P.x[d][i];
  for (int d = 0; d < P.D; ++d) r[k++] = P.xold[d][i];
  for (int c = 0; c < 3; ++c) r[k++] = P.v[c][i];
  for (int c = 0; c < 3; ++c) r[k++] = P.vold[c][i];
  r[k++] = P.w[i];
  r[k++] = __longlong_as_double((long long)P.id[i]);
}
__device__ __forceinline__ void wire_in(const WirePtrs &P, long i, const double *r) {
  int k = 0;
  for (int d = 0; d < P.D; ++d) P.x[d][i] = r[k++];
  for (int d = 0; d < P.D; ++d) P.xold[d][i] = r[k++];
  for (int c = 0; c < 3; ++c) P.v[c][i] = r[k++];
  for (int c = 0; c < 3; ++c) P.vold[c][i] = r[k++];
  P.w[i] = r[k++];
  P.id[i] = (uint64_t)__double_as_longlong(r[k++]);
}

// records grouped by direction code: offset[code] + running cursor
__global__ void k_pack_leavers(WirePtrs P, const int *list, const int *dead, MigCounters *mc, const unsigned *offset,
                               double *buf, long new_n, int *holes) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= mc->nleave) return;
  const long i = list[t];
  const int code = dead[i] - 1;
  if (code < 9) {
    const unsigned pos = offset[code] + atomicAdd(&mc->cursor[code], 1u);
    wire_out(P, i, buf + (size_t)pos * wire_len(P.D));
  }
  if (i < new_n) holes[atomicAdd(&mc->nhole, 1u)] = (int)i;
}
__global__ void k_list_movers(const int *dead, long new_n, long n, MigCounters *mc, int *movers) {
  const long i = new_n + (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!dead[i]) movers[atomicAdd(&mc->nmove, 1u)] = (int)i;
}
__global__ void k_fill_holes(WirePtrs P, const int *holes, const int *movers, const MigCounters *mc) {
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= mc->nhole) return;
  const long dst = holes[t], src = movers[t];
  for (int d = 0; d < P.D; ++d) {
    P.x[d][dst] = P.x[d][src];
    P.xold[d][dst] = P.xold[d][src];
  }
  for (int c = 0; c < 3; ++c) {
    P.v[c][dst] = P.v[c][src];
    P.vold[c][dst] = P.vold[c][src];
  }
  P.w[dst] = P.w[src];
  P.id[dst] = P.id[src];
}
__global__ void k_append(WirePtrs P, long n0, long nadd, const double *buf) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nadd) return;
  wire_in(P, n0 + t, buf + (size_t)t * wire_len(P.D));
}

// ---- migration over peer memory (device-side counts) --------------------------------------
// Inbox of one species of one box: [flags u64[16]] [counts u32 [2 parity][16]] pad to 1 KiB, then
// [2 parity][9 areas][cap records] wire records.  Area a = direction code the particles ARRIVE from
// (= 8 - the sender's code).
constexpr size_t MIGBOX_HEADER = 1024;
struct MigBoxView {
  unsigned long long *flags;   // [16]
  unsigned *counts;            // [2][16]
  double *recs;                // [2][9][cap][nw]
};
__host__ __device__ inline MigBoxView migbox_view(void *base) {
  MigBoxView v;
  unsigned char *b = reinterpret_cast<unsigned char *>(base);
  v.flags = reinterpret_cast<unsigned long long *>(b);
  v.counts = reinterpret_cast<unsigned *>(b + 16 * sizeof(unsigned long long));
  v.recs = reinterpret_cast<double *>(b + MIGBOX_HEADER);
  return v;
}
struct MigPeers {
  void *inbox[9];              // the neighbour's inbox per direction code (nullptr: no neighbour)
};
struct MigResult {             // written by the receiving kernel, read back by the host
  long long n_final, n_arrived, n_left, n_lost;
  unsigned overflow, pad;      // leavers that did not fit a neighbour's inbox area (dropped; reported as an error)
};

// leavers -> the neighbours' inboxes (peer stores), holes below new_n = n - nleave
__global__ void k_mig_pack_send(WirePtrs P, const int *list, const int *dead, MigCounters *mc, MigPeers peers,
                                unsigned parity, long cap, long n, int *holes, long list_cap, unsigned *overflow) {
  const unsigned nleave = mc->nleave;
  const long new_n = n - (long)nleave;
  const int nw = wire_len(P.D);
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < nleave; t += gridDim.x * blockDim.x) {
    const long i = list[t];
    const int code = dead[i] - 1;
    if (code < 9 && peers.inbox[code]) {
      const unsigned pos = atomicAdd(&mc->cursor[code], 1u);
      if (pos < (unsigned long)cap) {
        const MigBoxView v = migbox_view(peers.inbox[code]);
        wire_out(P, i, v.recs + (((size_t)parity * 9 + (8 - code)) * cap + pos) * nw);
      } else {
        atomicAdd(overflow, 1u);        // dropped: counted, reported by pgpu_migrate_finish as an error
      }
    } else if (code < 9 && code != 4) {
      atomicAdd(&mc->count[9], 1u);     // a leaver towards a direction without a connected box is lost, not sent
    }
    if (i < new_n) {
      const unsigned h = atomicAdd(&mc->nhole, 1u);
      if ((long)h < list_cap) holes[h] = (int)i;
      else atomicAdd(overflow, 1u);
    }
  }
  __threadfence_system();
}
__global__ void k_mig_list_movers(const int *dead, long n, MigCounters *mc, int *movers, long list_cap) {
  const unsigned nleave = mc->nleave;
  const long new_n = n - (long)nleave;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < nleave; t += gridDim.x * blockDim.x) {
    const long i = new_n + t;
    if (!dead[i]) {
      const unsigned m = atomicAdd(&mc->nmove, 1u);
      if ((long)m < list_cap) movers[m] = (int)i;
    }
  }
}
__global__ void k_mig_fill_holes(WirePtrs P, const int *holes, const int *movers, const MigCounters *mc, long list_cap) {
  const unsigned nh = mc->nhole < (unsigned long)list_cap ? mc->nhole : (unsigned)list_cap;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < nh; t += gridDim.x * blockDim.x) {
    const long dst = holes[t], src = movers[t];
    for (int d = 0; d < P.D; ++d) {
      P.x[d][dst] = P.x[d][src];
      P.xold[d][dst] = P.xold[d][src];
    }
    for (int c = 0; c < 3; ++c) {
      P.v[c][dst] = P.v[c][src];
      P.vold[c][dst] = P.vold[c][src];
    }
    P.w[dst] = P.w[src];
    P.id[dst] = P.id[src];
  }
}
// counts + arrival flags to the neighbours (after k_mig_pack_send in stream order)
__global__ void k_mig_post(const MigCounters *mc, MigPeers peers, unsigned parity, long cap, unsigned long long seq) {
  const int code = threadIdx.x;
  if (code >= 9 || code == 4 || !peers.inbox[code]) return;
  const MigBoxView v = migbox_view(peers.inbox[code]);
  const unsigned c = mc->cursor[code] < (unsigned long)cap ? mc->cursor[code] : (unsigned)cap;
  v.counts[parity * 16 + (8 - code)] = c;
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(v.flags + (8 - code)), "l"(seq) : "memory");
}
// wait for every neighbour, append what arrived behind the survivors
__global__ void k_mig_recv_append(WirePtrs P, void *inbox, unsigned area_mask, unsigned parity, long cap, long n,
                                  const MigCounters *mc, unsigned long long seq, MigResult *res, const unsigned *overflow) {
  __shared__ unsigned cnt[9];
  const MigBoxView v = migbox_view(inbox);
  if (threadIdx.x < 9) {
    const int a = threadIdx.x;
    unsigned c = 0;
    if (area_mask & (1u << a)) {
      unsigned long long f;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(v.flags + a) : "memory");
        if (f < seq) __nanosleep(64);
      } while (f < seq);
      c = __ldcg(v.counts + parity * 16 + a);
    }
    cnt[a] = c;
  }
  __syncthreads();
  const long new_n = n - (long)mc->nleave;
  const int nw = wire_len(P.D);
  long base = new_n, total = 0;
  for (int a = 0; a < 9; ++a) total += cnt[a];
  for (int a = 0; a < 9; ++a) {
    const long c = cnt[a];
    const double *src = v.recs + ((size_t)parity * 9 + a) * cap * nw;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < c * nw; t += (long)gridDim.x * blockDim.x) {
      // one double per thread: coalesced reads of the inbox (written by a peer: bypass L1)
      const long j = t / nw;
      const int k = (int)(t - j * nw);
      const double val = __ldcg(src + t);
      const long i = base + j;
      const int D = P.D;
      if (k < D) P.x[k][i] = val;
      else if (k < 2 * D) P.xold[k - D][i] = val;
      else if (k < 2 * D + 3) P.v[k - 2 * D][i] = val;
      else if (k < 2 * D + 6) P.vold[k - 2 * D - 3][i] = val;
      else if (k == 2 * D + 6) P.w[i] = val;
      else P.id[i] = (uint64_t)__double_as_longlong(val);
    }
    base += c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    res->n_final = new_n + total;
    res->n_arrived = total;
    res->n_left = mc->nleave;
    res->n_lost = mc->count[9];
    res->overflow = *overflow;
  }
}

static WirePtrs wire_ptrs(pgpu_species_s *s) {
  WirePtrs P;
  P.D = s->grid->desc.D;
  for (int d = 0; d < 2; ++d) {
    P.x[d] = s->x[d];
    P.xold[d] = s->xold[d];
  }
  for (int c = 0; c < 3; ++c) {
    P.v[c] = s->v[c];
    P.vold[c] = s->vold[c];
  }
  P.w = s->w;
  P.id = s->id;
  return P;
}

}  // namespace pgpu

using namespace pgpu;

extern "C" {

int pgpu_fab_pack_d(pgpu_grid_t g, int kind, int comp, const int *lo, const int *hi, double *buf_d) {
  if (!g) return PGPU_ERR_ARG;
  if (!ctx().inited) return PGPU_ERR_STATE;
  const DeviceFab *f = pick_fab(g, kind, comp);
  if (!f || !buf_d) return PGPU_ERR_ARG;
  int m[2];
  if (check_box(*f, g->desc.D, lo, hi, m)) return PGPU_ERR_ARG;
  KTimer t("halo_pack");
  k_fab_pack<<<nb((long)m[0] * m[1]), 256, 0, ctx().stream>>>(f->view(), lo[0], g->desc.D == 2 ? lo[1] : 0, m[0], m[1],
                                                             buf_d);
  return 0;
}

int pgpu_fab_unpack_d(pgpu_grid_t g, int kind, int comp, const int *lo, const int *hi, const double *buf_d, int add) {
  if (!g) return PGPU_ERR_ARG;
  if (!ctx().inited) return PGPU_ERR_STATE;
  const DeviceFab *f = pick_fab(g, kind, comp);
  if (!f || !buf_d) return PGPU_ERR_ARG;
  int m[2];
  if (check_box(*f, g->desc.D, lo, hi, m)) return PGPU_ERR_ARG;
  KTimer t("halo_unpack");
  k_fab_unpack<<<nb((long)m[0] * m[1]), 256, 0, ctx().stream>>>(f->view(), lo[0], g->desc.D == 2 ? lo[1] : 0, m[0],
                                                               m[1], buf_d, add);
  return 0;
}

int pgpu_wire_doubles(pgpu_grid_t g) { return g ? 2 * g->desc.D + 8 : 0; }

__global__ void k_counts_out(const MigCounters *mc, long long *out) {
  if (threadIdx.x < 10) out[threadIdx.x] = (long long)mc->count[threadIdx.x];
}

// Launches the marking pass; no synchronisation.  counts_d (optional): device int64[10].
static int mark_leavers_launch(pgpu_species_s *s, long long *counts_d) {
  Context &c = ctx();
  const pgpu_grid_s *g = s->grid;
  if (!s->mig) PGPU_CUDA(cudaMalloc(&s->mig, sizeof(MigCounters) + 16 * sizeof(unsigned)));
  PGPU_CUDA(cudaMemsetAsync(s->mig, 0, sizeof(MigCounters) + 16 * sizeof(unsigned), c.stream));
  s->mig_marked = false;
  s->mig_leave = 0;
  for (int k = 0; k < 10; ++k) s->mig_count[k] = 0;
  if (s->n == 0) {
    if (counts_d) k_counts_out<<<1, 32, 0, c.stream>>>((const MigCounters *)s->mig, counts_d);
    return 0;
  }
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  MigGeo G;
  G.D = g->desc.D;
  for (int d = 0; d < 2; ++d) {
    const bool on = d < G.D;
    G.le[d] = g->geo.le[d];
    G.boxlen[d] = on ? g->geo.dx[d] * g->nbox[d] : 1.0;
    G.my[d] = on ? g->desc.box_lo[d] / g->nbox[d] : 0;
    G.nbx[d] = on ? g->desc.ncell[d] / g->nbox[d] : 1;
    G.periodic[d] = on ? g->desc.periodic[d] : 0;
    if (on && (g->desc.box_lo[d] % g->nbox[d] || g->desc.ncell[d] % g->nbox[d])) {
      set_error("migration needs the square equal-box decomposition of System.cpp:169-245");
      return PGPU_ERR_ARG;
    }
  }
  {
    KTimer t("mig_mark");
    k_mark_leavers<<<nb(s->n), 256, 0, c.stream>>>(s->x[0], s->x[1], s->n, G, s->cell_key, s->perm,
                                                  (MigCounters *)s->mig);
    if (counts_d) k_counts_out<<<1, 32, 0, c.stream>>>((const MigCounters *)s->mig, counts_d);
  }
  s->binned = false;
  return 0;
}

int pgpu_species_mark_leavers(pgpu_species_t s, long *counts) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!s || !counts) return PGPU_ERR_ARG;
  Context &c = ctx();
  for (int k = 0; k < 10; ++k) counts[k] = 0;
  const int rc = mark_leavers_launch(s, nullptr);
  if (rc) return rc;
  if (s->n) {
    MigCounters h;
    PGPU_CUDA(cudaMemcpyAsync(&h, s->mig, sizeof(MigCounters), cudaMemcpyDeviceToHost, c.stream));
    PGPU_CUDA(cudaStreamSynchronize(c.stream));
    for (int k = 0; k < 10; ++k) counts[k] = s->mig_count[k] = h.count[k];
    s->mig_leave = h.nleave;
  }
  s->mig_marked = true;
  return 0;
}

int pgpu_species_mark_leavers_d(pgpu_species_t s, long long *counts_d) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!s || !counts_d) return PGPU_ERR_ARG;
  return mark_leavers_launch(s, counts_d);
}

int pgpu_species_set_leaver_counts(pgpu_species_t s, const long *counts) {
  if (!s || !counts) return PGPU_ERR_ARG;
  long tot = 0;
  for (int k = 0; k < 10; ++k) {
    s->mig_count[k] = counts[k];
    tot += counts[k];
  }
  s->mig_leave = tot;
  s->mig_marked = true;
  return 0;
}

int pgpu_species_pack_leavers_d(pgpu_species_t s, double *buf_d) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!s || !s->mig_marked) {
    set_error("pgpu_species_pack_leavers_d needs pgpu_species_mark_leavers first");
    return PGPU_ERR_STATE;
  }
  Context &c = ctx();
  const long L = s->mig_leave;
  s->mig_marked = false;
  if (L == 0) return 0;
  if (!buf_d) return PGPU_ERR_ARG;
  unsigned off[16] = {0};
  for (int k = 1; k < 10; ++k) off[k] = off[k - 1] + (k - 1 == 4 ? 0u : (unsigned)s->mig_count[k - 1]);
  unsigned *d_off = (unsigned *)((char *)s->mig + sizeof(MigCounters));
  PGPU_CUDA(cudaMemcpyAsync(d_off, off, sizeof(off), cudaMemcpyHostToDevice, c.stream));
  const long new_n = s->n - L;
  // scratch: holes and movers lists live in the second half of the key array / spare ints
  if (s->mig_list_cap < (size_t)(2 * L)) {
    if (s->mig_list) cudaFree(s->mig_list);
    s->mig_list_cap = (size_t)(4 * L + 65536);   // generous: a reallocation stalls the device
    PGPU_CUDA(cudaMalloc(&s->mig_list, s->mig_list_cap * sizeof(int)));
  }
  int *holes = s->mig_list, *movers = s->mig_list + L;
  const WirePtrs P = wire_ptrs(s);
  {
    KTimer t("mig_pack");
    k_pack_leavers<<<nb(L), 256, 0, c.stream>>>(P, s->perm, s->cell_key, (MigCounters *)s->mig, d_off, buf_d, new_n,
                                               holes);
    k_list_movers<<<nb(L), 256, 0, c.stream>>>(s->cell_key, new_n, s->n, (MigCounters *)s->mig, movers);
    k_fill_holes<<<nb(L), 256, 0, c.stream>>>(P, holes, movers, (const MigCounters *)s->mig);
  }
  s->n = new_n;
  s->binned = false;
  return 0;
}

int pgpu_species_append_d(pgpu_species_t s, long n_add, const double *buf_d) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!s || n_add < 0) return PGPU_ERR_ARG;
  if (n_add == 0) return 0;
  if (!buf_d) return PGPU_ERR_ARG;
  if (materialize_old(s)) return PGPU_ERR_CUDA;
  if (grow_capacity(s, s->n + n_add)) return PGPU_ERR_CUDA;
  {
    KTimer t("mig_append");
    k_append<<<nb(n_add), 256, 0, ctx().stream>>>(wire_ptrs(s), s->n, n_add, buf_d);
  }
  s->n += n_add;
  s->binned = false;
  return 0;
}

// ---- pgpu_migrator_*: migration over peer memory ---------------------------------------------
struct pgpu_migrator_s {
  pgpu_species_t s = nullptr;
  long cap = 0;                 // records per inbox area
  int nw = 0;
  void *inbox = nullptr;
  size_t inbox_bytes = 0;
  pgpu::MigPeers peers;
  unsigned area_mask = 0;       // areas (arrival codes) that have a neighbour
  unsigned long long seq = 0;
  pgpu::MigResult *d_res = nullptr, *h_res = nullptr;
  unsigned *d_overflow = nullptr;
  long n_at_send = 0;
  bool sent = false, received = false;
  std::vector<void *> opened;
};

int pgpu_migrator_create(pgpu_species_t s, long capacity_records, pgpu_migrator_t *out) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!s || !out || capacity_records < 1) return PGPU_ERR_ARG;
  pgpu_migrator_s *m = new pgpu_migrator_s;
  m->s = s;
  m->cap = capacity_records;
  m->nw = 2 * s->grid->desc.D + 8;
  m->inbox_bytes = MIGBOX_HEADER + (size_t)2 * 9 * m->cap * m->nw * sizeof(double);
  for (int k = 0; k < 9; ++k) m->peers.inbox[k] = nullptr;
  PGPU_CUDA(cudaMalloc(&m->inbox, m->inbox_bytes));
  PGPU_CUDA(cudaMemset(m->inbox, 0, MIGBOX_HEADER));
  PGPU_CUDA(cudaMalloc(&m->d_res, sizeof(MigResult)));
  PGPU_CUDA(cudaMallocHost(&m->h_res, sizeof(MigResult)));
  PGPU_CUDA(cudaMalloc(&m->d_overflow, sizeof(unsigned)));
  PGPU_CUDA(cudaMemset(m->d_overflow, 0, sizeof(unsigned)));
  *out = m;
  return 0;
}

int pgpu_migrator_destroy(pgpu_migrator_t m) {
  if (!m) return 0;
  cudaStreamSynchronize(ctx().stream);
  for (void *p : m->opened) cudaIpcCloseMemHandle(p);
  cudaFree(m->inbox);
  cudaFree(m->d_res);
  cudaFreeHost(m->h_res);
  cudaFree(m->d_overflow);
  delete m;
  return 0;
}

int pgpu_migrator_inbox(pgpu_migrator_t m, void **inbox_d, size_t *bytes) {
  if (!m || !inbox_d) return PGPU_ERR_ARG;
  *inbox_d = m->inbox;
  if (bytes) *bytes = m->inbox_bytes;
  return 0;
}

int pgpu_migrator_ipc_handle(pgpu_migrator_t m, void *handle64) {
  if (!m || !handle64) return PGPU_ERR_ARG;
  cudaIpcMemHandle_t hd;
  PGPU_CUDA(cudaIpcGetMemHandle(&hd, m->inbox));
  memcpy(handle64, &hd, 64);
  return 0;
}

int pgpu_migrator_ipc_open(pgpu_migrator_t m, const void *handle64, void **inbox_d) {
  if (!m || !handle64 || !inbox_d) return PGPU_ERR_ARG;
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  void *p = nullptr;
  PGPU_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  m->opened.push_back(p);
  *inbox_d = p;
  return 0;
}

int pgpu_migrator_connect(pgpu_migrator_t m, int code, void *peer_inbox_d) {
  if (!m || code < 0 || code > 8 || code == 4 || !peer_inbox_d) return PGPU_ERR_ARG;
  m->peers.inbox[code] = peer_inbox_d;
  // the neighbour in direction `code` sends me its leavers of direction 8 - code, which arrive from `code`
  m->area_mask |= 1u << code;
  return 0;
}

int pgpu_migrate_send(pgpu_migrator_t m) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!m || m->sent) return PGPU_ERR_ARG;
  Context &c = ctx();
  pgpu_species_s *s = m->s;
  // room for everything that can arrive, before anything is in flight (a reallocation synchronises)
  if (grow_capacity(s, s->n + 8 * m->cap)) return PGPU_ERR_CUDA;
  int rc = mark_leavers_launch(s, nullptr);
  if (rc) return rc;
  m->seq += 1;
  m->n_at_send = s->n;
  m->sent = true;
  const long list_cap = 8 * m->cap + 65536;
  if (s->mig_list_cap < (size_t)(2 * list_cap)) {
    if (s->mig_list) cudaFree(s->mig_list);
    s->mig_list_cap = (size_t)(2 * list_cap);
    PGPU_CUDA(cudaMalloc(&s->mig_list, s->mig_list_cap * sizeof(int)));
  }
  if (!s->mig) {   // n == 0 took the early exit of the marking pass
    PGPU_CUDA(cudaMalloc(&s->mig, sizeof(MigCounters) + 16 * sizeof(unsigned)));
    PGPU_CUDA(cudaMemsetAsync(s->mig, 0, sizeof(MigCounters) + 16 * sizeof(unsigned), c.stream));
  }
  int *holes = s->mig_list, *movers = s->mig_list + list_cap;
  const WirePtrs P = wire_ptrs(s);
  const unsigned parity = (unsigned)(m->seq & 1);
  MigCounters *mc = (MigCounters *)s->mig;
  const unsigned grid = (unsigned)(c.sm_count * 4);
  {
    KTimer t("mig_p2p_send");
    if (s->n) {
      k_mig_pack_send<<<grid, 256, 0, c.stream>>>(P, s->perm, s->cell_key, mc, m->peers, parity, m->cap, s->n, holes,
                                                 list_cap, m->d_overflow);
      k_mig_list_movers<<<grid, 256, 0, c.stream>>>(s->cell_key, s->n, mc, movers, list_cap);
      k_mig_fill_holes<<<grid, 256, 0, c.stream>>>(P, holes, movers, mc, list_cap);
    }
    k_mig_post<<<1, 32, 0, c.stream>>>(mc, m->peers, parity, m->cap, m->seq);
  }
  return 0;
}

int pgpu_migrate_recv(pgpu_migrator_t m) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!m || !m->sent || m->received) return PGPU_ERR_ARG;
  Context &c = ctx();
  pgpu_species_s *s = m->s;
  {
    KTimer t("mig_p2p_recv");
    k_mig_recv_append<<<(unsigned)c.sm_count, 256, 0, c.stream>>>(wire_ptrs(s), m->inbox, m->area_mask,
                                                                 (unsigned)(m->seq & 1), m->cap, m->n_at_send,
                                                                 (const MigCounters *)s->mig, m->seq, m->d_res,
                                                                 m->d_overflow);
  }
  PGPU_CUDA(cudaMemcpyAsync(m->h_res, m->d_res, sizeof(MigResult), cudaMemcpyDeviceToHost, c.stream));
  m->received = true;
  return 0;
}

int pgpu_migrate_finish(pgpu_migrator_t m, long *n_arrived, long *n_left, long *n_lost) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!m || !m->received) return PGPU_ERR_ARG;
  PGPU_CUDA(cudaStreamSynchronize(ctx().stream));
  m->sent = m->received = false;
  pgpu_species_s *s = m->s;
  // the species is left consistent in every case: the leavers are gone, the arrivals appended, the count updated
  s->n = (long)m->h_res->n_final;
  s->binned = false;
  s->mig_marked = false;
  if (n_arrived) *n_arrived = (long)m->h_res->n_arrived;
  if (n_left) *n_left = (long)m->h_res->n_left;
  if (n_lost) *n_lost = (long)m->h_res->n_lost + (long)m->h_res->overflow;
  if (m->h_res->overflow) {
    PGPU_CUDA(cudaMemsetAsync(m->d_overflow, 0, sizeof(unsigned), ctx().stream));
    set_error("migration inbox overflow: %u leavers did not fit an inbox area of %ld particles and were dropped (raise "
              "the capacity of pgpu_migrator_create)", m->h_res->overflow, m->cap);
    return PGPU_ERR_STATE;
  }
  return 0;
}

}  // extern "C"
