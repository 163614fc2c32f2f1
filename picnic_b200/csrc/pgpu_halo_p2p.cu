// pgpu_halo_p2p.cu -- ghost ADD-exchange of the total current J over peer memory (NVLink / NVSwitch).
//
// Replaces LevelData::exchange with an add op in PicSpeciesInterface::finalizeSettingJ
// (src/species/pic/PicSpeciesInterface.cpp:766-772) for the one-box-per-GPU decomposition.
// Instead of pack kernels + NCCL send/recv + unpack kernels (12 launches and one NCCL group per
// direction), every direction is TWO kernels:
//   k_halo_send      gathers the overlap strips of Jx, Jy, Jz and stores them straight into the
//                    neighbour's inbox (a cudaMalloc'ed buffer of the peer process, mapped here through
//                    CUDA IPC; loads/stores travel over NVLink), then -- last block, after a system
//                    fence -- stamps the neighbour's arrival flag with the exchange's sequence number;
//   k_halo_recv_add  spins until the flags of this box's own inbox carry the sequence number, then adds
//                    the received strips into J.
// The exchange stays direction by direction over the full transverse extent (ghosts included), so
// corners need no diagonal messages (picnic_b200/halo.py).  Inbox areas are double buffered by the
// parity of the sequence number: a box cannot start exchange k+2 before its neighbour has sent k+1,
// i.e. before the neighbour's reads of exchange k are over.  No host synchronisation, no collective
// library call on the data path.
#include "pgpu_internal.h"

#include <vector>

namespace pgpu {

constexpr int HALO_MAX_MSG = 2;      // messages per phase: the -side and the +side neighbour
constexpr int HALO_THREADS = 256;
constexpr size_t HALO_FLAG_BYTES = 1024;   // arrival flags (one u64 per inbox area) in front of the areas
constexpr int HALO_MAX_AREAS = (int)(HALO_FLAG_BYTES / sizeof(unsigned long long));

struct HaloPart {      // one component's index box inside a message
  double *p;           // array element (lo0, lo1)
  long n0;             // doubles between rows of the array
  int m0, m1;          // extent of the index box
  long off;            // offset of the part inside the message
};
struct HaloMsg {
  HaloPart part[3];
  long count;                      // doubles in the message
  double *remote;                  // area of this message in the peer's inbox (both parity slots)
  const double *local;             // area in this box's inbox that the peer's message lands in
  unsigned long long *remote_flag;
  const unsigned long long *local_flag;
};
struct HaloPhase {
  HaloMsg msg[HALO_MAX_MSG];
  int nmsg;
  long total;                      // doubles of all messages
};

__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(HALO_THREADS) k_halo_send(HaloPhase P, unsigned long long seq, unsigned *done,
                                                            unsigned done_target) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < P.total; t += stride) {
    long r = t;
    int m = 0;
    while (m + 1 < P.nmsg && r >= P.msg[m].count) r -= P.msg[m++].count;
    const HaloMsg &M = P.msg[m];
    int c = 0;
    while (c < 2 && r >= M.part[c + 1].off) ++c;
    const HaloPart &Q = M.part[c];
    const long k = r - Q.off;
    const int a = (int)(k % Q.m0), b = (int)(k / Q.m0);
    M.remote[(seq & 1) * M.count + r] = Q.p[a + b * Q.n0];     // peer store
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                       // this block's peer stores are visible system-wide ...
    if (atomicAdd(done, 1u) + 1u == done_target) {
      __threadfence_system();                     // ... and so are those of every block counted before
      for (int m = 0; m < P.nmsg; ++m) st_sys(P.msg[m].remote_flag, seq);
    }
  }
}

__global__ void __launch_bounds__(HALO_THREADS) k_halo_recv_add(HaloPhase P, unsigned long long seq) {
  if (threadIdx.x == 0)
    for (int m = 0; m < P.nmsg; ++m)
      while (ld_sys(P.msg[m].local_flag) < seq) __nanosleep(64);
  __syncthreads();
  const long stride = (long)gridDim.x * blockDim.x;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < P.total; t += stride) {
    long r = t;
    int m = 0;
    while (m + 1 < P.nmsg && r >= P.msg[m].count) r -= P.msg[m++].count;
    const HaloMsg &M = P.msg[m];
    int c = 0;
    while (c < 2 && r >= M.part[c + 1].off) ++c;
    const HaloPart &Q = M.part[c];
    const long k = r - Q.off;
    const int a = (int)(k % Q.m0), b = (int)(k / Q.m0);
    double *p = Q.p + (a + b * Q.n0);
    // the inbox was written by another device: read it past the (non-coherent) L1
    *p = __dadd_rn(*p, __ldcg(M.local + (seq & 1) * M.count + r));
  }
}

}  // namespace pgpu

using namespace pgpu;

struct pgpu_halo_s {
  pgpu_grid_t grid = nullptr;
  // the device arrays the plan adds into: the three J components (pgpu_halo_create) or one resident charge-density
  // array (pgpu_halo_create_rho, nparts == 1)
  const DeviceFab *target[3] = {nullptr, nullptr, nullptr};
  int nparts = 3;
  std::vector<pgpu_halo_msg> msgs;
  std::vector<long> count, area_off;     // per message: doubles, offset of its area in an inbox (doubles)
  std::vector<double *> remote_base;     // per message: the peer's inbox (nullptr until connected),
  std::vector<int> remote_area;          //   the area of that inbox the message lands in
  std::vector<long> remote_off;          //   and the area's offset in doubles
  int nphase = 0;
  unsigned char *inbox = nullptr;        // [flags: nmsg u64, padded to 256 B][areas: 2 slots each]
  size_t inbox_bytes = 0;
  unsigned *done = nullptr;              // per phase: blocks of k_halo_send that finished (monotonic)
  std::vector<unsigned> done_target;
  unsigned long long seq = 0;
  std::vector<void *> opened;            // IPC mappings to close
};

static long msg_count(pgpu_grid_t g, const pgpu_halo_msg &m, int nparts) {
  long n = 0;
  for (int c = 0; c < nparts; ++c) {
    long k = 1;
    for (int d = 0; d < g->desc.D; ++d) k *= (m.hi[c][d] - m.lo[c][d] + 1);
    n += k;
  }
  return n;
}

extern "C" {

static int halo_create(pgpu_grid_t g, const DeviceFab *const *target, int nparts, int nmsg, const pgpu_halo_msg *msgs,
                       pgpu_halo_t *out) {
  pgpu_halo_s *h = new pgpu_halo_s;
  h->grid = g;
  h->nparts = nparts;
  for (int c = 0; c < nparts; ++c) h->target[c] = target[c];
  std::vector<int> per_phase;
  long off = 0;
  for (int i = 0; i < nmsg; ++i) {
    const pgpu_halo_msg &m = msgs[i];
    if (m.phase < 0 || m.phase > 7 || m.recv_area < 0 || m.recv_area >= nmsg || nmsg > HALO_MAX_AREAS) {
      delete h;
      return PGPU_ERR_ARG;
    }
    for (int c = 0; c < nparts; ++c)
      for (int d = 0; d < g->desc.D; ++d)
        if (m.lo[c][d] < target[c]->lo[d] || m.hi[c][d] > target[c]->hi[d] || m.hi[c][d] < m.lo[c][d]) {
          set_error("halo message %d: index box of component %d is not inside the device array", i, c);
          delete h;
          return PGPU_ERR_ARG;
        }
    if ((int)per_phase.size() <= m.phase) per_phase.resize(m.phase + 1, 0);
    if (++per_phase[m.phase] > HALO_MAX_MSG) {
      set_error("more than %d halo messages in phase %d", HALO_MAX_MSG, m.phase);
      delete h;
      return PGPU_ERR_ARG;
    }
    h->msgs.push_back(m);
    h->count.push_back(msg_count(g, m, nparts));
  }
  // inbox areas in recv_area order, so that a peer with the mirrored plan knows where to write
  h->area_off.assign(nmsg, 0);
  {
    std::vector<long> by_area(nmsg, -1);
    for (int i = 0; i < nmsg; ++i) by_area[h->msgs[i].recv_area] = i;
    for (int a = 0; a < nmsg; ++a) {
      if (by_area[a] < 0) {
        set_error("halo plan: inbox area %d has no message", a);
        delete h;
        return PGPU_ERR_ARG;
      }
      h->area_off[by_area[a]] = off;
      off += 2 * ((h->count[by_area[a]] + 31) / 32 * 32);
    }
  }
  h->nphase = (int)per_phase.size();
  h->inbox_bytes = HALO_FLAG_BYTES + (size_t)off * sizeof(double);
  PGPU_CUDA(cudaMalloc(&h->inbox, h->inbox_bytes));
  PGPU_CUDA(cudaMemset(h->inbox, 0, h->inbox_bytes));
  PGPU_CUDA(cudaMalloc(&h->done, 8 * sizeof(unsigned)));
  PGPU_CUDA(cudaMemset(h->done, 0, 8 * sizeof(unsigned)));
  h->done_target.assign(8, 0);
  h->remote_base.assign(nmsg, nullptr);
  h->remote_area.assign(nmsg, 0);
  h->remote_off.assign(nmsg, 0);
  *out = h;
  return 0;
}

int pgpu_halo_create(pgpu_grid_t g, int nmsg, const pgpu_halo_msg *msgs, pgpu_halo_t *out) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!g || !out || nmsg < 0 || (nmsg && !msgs)) return PGPU_ERR_ARG;
  const DeviceFab *target[3] = {&g->jtot[0], &g->jtot[1], &g->jtot[2]};
  return halo_create(g, target, 3, nmsg, msgs, out);
}

int pgpu_halo_create_rho(pgpu_grid_t g, const int *stag, int nmsg, const pgpu_halo_msg *msgs, pgpu_halo_t *out) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!g || !stag || !out || nmsg < 0 || (nmsg && !msgs)) return PGPU_ERR_ARG;
  DeviceFab *f = nullptr;
  int rc = grid_rho_fab(g, stag, &f);
  if (rc) return rc;
  const DeviceFab *target[1] = {f};
  return halo_create(g, target, 1, nmsg, msgs, out);
}

int pgpu_halo_destroy(pgpu_halo_t h) {
  if (!h) return 0;
  cudaStreamSynchronize(ctx().stream);
  for (void *p : h->opened) cudaIpcCloseMemHandle(p);
  if (h->inbox) cudaFree(h->inbox);
  if (h->done) cudaFree(h->done);
  delete h;
  return 0;
}

int pgpu_halo_area_offset(pgpu_halo_t h, int msg, long *offset_doubles, long *count) {
  if (!h || msg < 0 || msg >= (int)h->msgs.size()) return PGPU_ERR_ARG;
  if (offset_doubles) *offset_doubles = h->area_off[msg];
  if (count) *count = h->count[msg];
  return 0;
}

int pgpu_halo_inbox(pgpu_halo_t h, void **inbox_d, size_t *bytes) {
  if (!h || !inbox_d) return PGPU_ERR_ARG;
  *inbox_d = h->inbox;
  if (bytes) *bytes = h->inbox_bytes;
  return 0;
}

int pgpu_halo_ipc_handle(pgpu_halo_t h, void *handle64) {
  if (!h || !handle64) return PGPU_ERR_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t hd;
  PGPU_CUDA(cudaIpcGetMemHandle(&hd, h->inbox));
  memcpy(handle64, &hd, 64);
  return 0;
}

int pgpu_halo_ipc_open(pgpu_halo_t h, const void *handle64, void **inbox_d) {
  if (!h || !handle64 || !inbox_d) return PGPU_ERR_ARG;
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  void *p = nullptr;
  PGPU_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  h->opened.push_back(p);
  *inbox_d = p;
  return 0;
}

int pgpu_halo_connect(pgpu_halo_t h, int msg, void *peer_inbox_d, int peer_area, long peer_area_offset) {
  if (!h || msg < 0 || msg >= (int)h->msgs.size() || !peer_inbox_d || peer_area < 0 || peer_area >= HALO_MAX_AREAS ||
      peer_area_offset < 0)
    return PGPU_ERR_ARG;
  h->remote_base[msg] = reinterpret_cast<double *>(peer_inbox_d);
  h->remote_area[msg] = peer_area;
  h->remote_off[msg] = peer_area_offset;
  return 0;
}

static int fill_phase(pgpu_halo_s *h, int phase, HaloPhase *P) {
  pgpu_grid_s *g = h->grid;
  const int D = g->desc.D;
  P->nmsg = 0;
  P->total = 0;
  for (size_t i = 0; i < h->msgs.size(); ++i) {
    const pgpu_halo_msg &m = h->msgs[i];
    if (m.phase != phase) continue;
    if (!h->remote_base[i]) {
      set_error("halo message %d is not connected to its peer's inbox (pgpu_halo_connect)", (int)i);
      return PGPU_ERR_STATE;
    }
    HaloMsg &M = P->msg[P->nmsg++];
    long off = 0;
    for (int c = 0; c < 3; ++c) {
      HaloPart &Q = M.part[c];
      if (c >= h->nparts) {               // an empty part: the kernels' part search never selects it
        Q.n0 = 1, Q.m0 = 0, Q.m1 = 0, Q.p = nullptr, Q.off = off;
        continue;
      }
      const DeviceFab &f = *h->target[c];
      Q.n0 = f.n0;
      Q.m0 = m.hi[c][0] - m.lo[c][0] + 1;
      Q.m1 = D == 2 ? m.hi[c][1] - m.lo[c][1] + 1 : 1;
      Q.p = f.p + (m.lo[c][0] - f.lo[0]) + (D == 2 ? (long)(m.lo[c][1] - f.lo[1]) * f.n0 : 0);
      Q.off = off;
      off += (long)Q.m0 * Q.m1;
    }
    M.count = off;
    unsigned char *rb = reinterpret_cast<unsigned char *>(h->remote_base[i]);
    M.remote_flag = reinterpret_cast<unsigned long long *>(rb) + h->remote_area[i];
    M.remote = reinterpret_cast<double *>(rb + HALO_FLAG_BYTES) + h->remote_off[i];
    M.local_flag = reinterpret_cast<const unsigned long long *>(h->inbox) + m.recv_area;
    M.local = reinterpret_cast<const double *>(h->inbox + HALO_FLAG_BYTES) + h->area_off[i];
    P->total += off;
  }
  return 0;
}

int pgpu_halo_begin(pgpu_halo_t h) {
  if (!h) return PGPU_ERR_ARG;
  h->seq += 1;
  return 0;
}

int pgpu_halo_send(pgpu_halo_t h, int phase) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!h || phase < 0 || phase >= h->nphase || h->seq == 0) return PGPU_ERR_ARG;
  HaloPhase P;
  int rc = fill_phase(h, phase, &P);
  if (rc) return rc;
  if (P.nmsg == 0) return 0;
  const unsigned blocks = (unsigned)std::min<long>((P.total + HALO_THREADS - 1) / HALO_THREADS, ctx().sm_count);
  h->done_target[phase] += blocks;
  KTimer t("halo_p2p_send");
  k_halo_send<<<blocks, HALO_THREADS, 0, ctx().stream>>>(P, h->seq, h->done + phase, h->done_target[phase]);
  return 0;
}

int pgpu_halo_recv_add(pgpu_halo_t h, int phase) {
  if (!ctx().inited) return PGPU_ERR_STATE;
  if (!h || phase < 0 || phase >= h->nphase || h->seq == 0) return PGPU_ERR_ARG;
  HaloPhase P;
  int rc = fill_phase(h, phase, &P);
  if (rc) return rc;
  if (P.nmsg == 0) return 0;
  const unsigned blocks = (unsigned)std::min<long>((P.total + HALO_THREADS - 1) / HALO_THREADS, ctx().sm_count);
  KTimer t("halo_p2p_recv");
  k_halo_recv_add<<<blocks, HALO_THREADS, 0, ctx().stream>>>(P, h->seq);
  return 0;
}

int pgpu_halo_phases(pgpu_halo_t h) { return h ? h->nphase : 0; }

}  // extern "C"
